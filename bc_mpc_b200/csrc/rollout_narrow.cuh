// BF16 fast-mode rollout kernel for NARROW dynamics nets (hidden <= 256: the reference's default size, train_mpc_ppo.py:75,
// and the shape of saved_data/ppo-mpc/vars.pkl), uniform shooting, ReLU.
//
// Why a second kernel: a step is a chain of dependent hand-offs (input -> layer 1 -> epilogue -> layer 2 -> epilogue ->
// layer 3 -> state update, ~400 cycles per hop) whose length does not shrink with the network.  At hidden = 256 the MMAs of a
// step take ~2000 cycles and the chain ~4500, so the tensor pipe idles 70 % of the time in rollout_tmem_kernel.  Narrow
// hidden activations leave room in TMEM for TWO independent 128-path tiles per CTA: each tile has its own issuer warp, its
// own four epilogue warps and its own TMEM half, and the two dependency chains interleave on the (in-order) tensor pipe.
// All weights (<= 88 KB per CTA of a pair) stay resident in shared memory: no ring, no producer loop.
//
// CTA pair (cta_group::2, M = 256 per MMA), 384 threads:
//   warp 0      loads the weight images once (both CTAs: their own row halves)
//   warp 1 / 3  MMA issuer of tile pipe 0 / 1 (leader CTA); in the peer CTA warp 1 tells the leader that its halves landed
//   warp 2      TMEM allocation
//   warps 4-7   epilogue of pipe 0 (one warp per TMEM lane quadrant), warps 8-11 epilogue of pipe 1
// TMEM, per pipe t (columns 256 t + ...):
//   [  0,128)  A1  hidden-1 activations (bf16 pairs), A operand of layer 2
//   [128,192)  D0  accumulator: layer-1 even chunks, every layer-2 chunk
// (Splitting the layer-2 chunks into two independently drained N = 32 halves was tried: +1.5 % at hidden 256, -5 % below.)
//   [192,256)  X   layer-1 odd chunks; afterwards A2 [192,224) (one hidden-2 chunk) and D3 [224,256) (output layer)
// Biases: layer 1 through two constant-1 input columns (as in rollout_tmem_kernel); layers 2 / 3 added in the epilogues
// (the weight images are those of the kVarBias variant, packed by weight_pack.cuh).
//
// Every barrier completes exactly once per step (parity = step & 1).  Replaces the same reference lines as
// rollout_tmem_kernel (dynamics.py:106-119, cost_functions.py:10-30,59-63, controllers.py:53,63-85).
#pragma once
#include "rollout_tmem.cuh"

namespace bcmpc {
namespace tmx {

constexpr int kNwMaxChunks = 4;
// per-pipe event barriers
enum : int {
  kNwA0Ready = 0,      // input rows of the step are in shared memory               (epilogue -> issuer)
  kNwD3Full = 1,       // output accumulator complete                               (issuer -> epilogue)
  kNwDFull1 = 2,       // [4] layer-1 accumulator of chunk c complete               (issuer -> epilogue)
  kNwDEmpty1 = 6,      // [4] ... drained                                           (epilogue -> issuer)
  kNwA1Ready = 10,     // [4] hidden-1 k-block c stored                              (epilogue -> issuer)
  kNwDFull2 = 14,      // [4] layer-2 accumulator of chunk c complete
  kNwDEmpty2 = 18,     // [4] ... drained
  kNwA2Ready = 22,     // [4] hidden-2 chunk c stored
  kNwA2Empty = 26,     // [4] layer 3 has consumed hidden-2 chunk c
  // proposal network (policy-guided shooting), once per step each
  kNwAp0Ready = 30,    // policy input rows in shared memory                        (epilogue -> issuer)
  kNwPdFull = 31,      // [4] policy accumulators: layer-1 chunks 0,1, layer-2 chunks 0,1
  kNwAp1Ready = 35,    // [2] policy hidden-1 chunk stored
  kNwAp2Ready = 37,    // [2] policy hidden-2 chunk stored
  kNwPolFull = 39,     // policy output accumulator complete
  kNwNumEv = 40
};

struct NwLayout {
  uint32_t blk_h = 0, blk_o = 0, w1_bytes = 0, w3_bytes = 0, w2_bytes = 0;
  uint32_t off_a0 = 0, off_w13 = 0, off_w2 = 0, off_vec = 0, off_pol = 0, pol_img_bytes = 0, off_misc = 0, smem_bytes = 0;
};
constexpr uint32_t kNwMiscBytes = 1024;      // [weights-landed barrier][2 x 40 events] x 8 B, tmem ptr, 2 x argmin scratch
constexpr uint32_t kNwOffEv = 8, kNwOffTmemPtr = 8 + 2 * kNwNumEv * 8, kNwOffRed = 704;
constexpr NwLayout nw_layout(int nc, bool pol = false) {
  NwLayout L;
  L.blk_h = (kChunk / 2) * 128u;
  L.blk_o = (kOutN / 2) * 128u;
  L.w1_bytes = nc * L.blk_h;
  L.w3_bytes = nc * L.blk_o;
  L.w2_bytes = nc * nc * L.blk_h;
  L.off_a0 = 0;                                   // two input tiles
  L.off_w13 = 2 * kA0Bytes;                       // [W1 chunks][W3 k-blocks]: the resident image of rollout_tmem_kernel
  L.off_w2 = L.off_w13 + L.w1_bytes + L.w3_bytes; // [chunk c][k-block]: its streamed image, all of it
  L.off_vec = L.off_w2 + L.w2_bytes;              // [b2 (64 nc)][b3 (32)]
  L.off_pol = (L.off_vec + 2u * (64u * nc + kOutN) * 4u + 1023u) / 1024u * 1024u;    // (LayerNorm: [u2][c2][u3][c3])
  L.pol_img_bytes = pol ? 6u * L.blk_h + 2u * L.blk_o : 0u;                           // proposal network images, then its biases
  L.off_misc = (L.off_pol + (pol ? L.pol_img_bytes + kPolVecFloats * 4u : 0u) + 1023u) / 1024u * 1024u;
  L.smem_bytes = L.off_misc + kNwMiscBytes;
  return L;
}
template <int NC, bool POL>
struct NwLay {
  static constexpr NwLayout value = nw_layout(NC, POL);
};

// LN: the LayerNorm variant of the dynamics net (dynamics.py:68-69), same algebra as rollout_tmem_kernel - un-normalised
// ReLU outputs r through the MMAs with W' = diag(gamma) W, then LN(r) W + b = rho (r W') - rho mu u + c in the next epilogue.
// Every row is handled by ONE thread for all chunks here, so the statistics never leave its registers.
// POL: policy-guided shooting - the proposal network (20 -> <=128 tanh -> <=128 tanh -> 6) runs at the top of every step
// on the pipe's own accumulators, its hidden activations in the (then idle) A1 columns, exactly as in rollout_tmem_kernel.
template <int NC, bool LN, bool POL, bool DBG>
__global__ void __launch_bounds__(kThreads, 1) rollout_narrow_kernel(const __grid_constant__ TmArgs a) {
  constexpr int CG = 2;
  constexpr NwLayout L = NwLay<NC, POL>::value;
  constexpr uint32_t kIdescH = ptx::umma_idesc_bf16(kRows * CG, kChunk), kIdescO = ptx::umma_idesc_bf16(kRows * CG, kOutN);
  static_assert(NC >= 1 && NC <= kNwMaxChunks && L.smem_bytes <= (uint32_t)kSmemLimit, "narrow kernel: hidden <= 256");

  extern __shared__ __align__(1024) uint8_t smem[];
  const PlanArgs& p = a.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // persistent CTA pairs (as in rollout_tmem_kernel): pair `first_unit` rolls the units first, first + stride, ...;
  // a unit = four tiles of one root state (two pipes x two CTAs)
  const int units_per_state = a.p.tiles_per_state / 4;
  const int total_units = a.p.B * units_per_state;
  const int first_unit = (int)blockIdx.x / CG, unit_stride = (int)gridDim.x / CG;
  const int n_it = first_unit < total_units ? (total_units - first_unit + unit_stride - 1) / unit_stride : 0;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t misc = base + L.off_misc;
  const uint32_t misc_leader = ptx::mapa(misc, 0);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.off_misc + kNwOffTmemPtr);
  const int H = p.H;
  const int Htot = n_it * H;
  const uint32_t hc = a.hint_chain;

  if (tid == 0) {
    if (base & 1023u) { printf("bcmpc: dynamic smem base not 1024-byte aligned\n"); __trap(); }
    ptx::mbar_init(misc, leader ? 2 : 1);            // weights landed (leader: + the peer's notice)
    for (int t = 0; t < 2; ++t) {
      const uint32_t ev = misc + kNwOffEv + (uint32_t)t * kNwNumEv * 8u;
      ptx::mbar_init(ev + 8u * kNwA0Ready, 4 * CG);
      ptx::mbar_init(ev + 8u * kNwD3Full, 1);
      for (int i = 0; i < 4; ++i) {
        ptx::mbar_init(ev + 8u * (kNwDFull1 + i), 1);
        ptx::mbar_init(ev + 8u * (kNwDEmpty1 + i), 4 * CG);
        ptx::mbar_init(ev + 8u * (kNwA1Ready + i), 4 * CG);
        ptx::mbar_init(ev + 8u * (kNwDFull2 + i), 1);
        ptx::mbar_init(ev + 8u * (kNwDEmpty2 + i), 4 * CG);
        ptx::mbar_init(ev + 8u * (kNwA2Ready + i), 4 * CG);
        ptx::mbar_init(ev + 8u * (kNwA2Empty + i), 1);
        ptx::mbar_init(ev + 8u * (kNwPdFull + i), 1);
      }
      ptx::mbar_init(ev + 8u * kNwAp0Ready, 4 * CG);
      ptx::mbar_init(ev + 8u * kNwPolFull, 1);
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(ev + 8u * (kNwAp1Ready + i), 4 * CG); ptx::mbar_init(ev + 8u * (kNwAp2Ready + i), 4 * CG); }
    }
    ptx::fence_mbar_init();
  }
  {
    float* s_vec = reinterpret_cast<float*>(smem + L.off_vec);
    for (int i = tid; i < (LN ? 2 : 1) * (64 * NC + kOutN); i += kThreads) s_vec[i] = a.lnvec[i];
    if constexpr (POL) {
      float* s_pv = reinterpret_cast<float*>(smem + L.off_pol + L.pol_img_bytes);
      for (int i = tid; i < 2 * kPolMaxHidden + kMaxAct; i += kThreads) s_pv[i] = a.polvec[i];
    }
  }
  if (warp == 2) ptx::tmem_alloc_cg<CG>(ptx::smem_u32(s_tmem), kTmemCols);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  if (*s_tmem != 0u) { if (tid == 0) printf("bcmpc: unexpected TMEM base 0x%x\n", *s_tmem); __trap(); }

  if (warp == 0) {
    // ===================== weights: every CTA loads its own row halves, once =====================
    if (ptx::elect_one()) {
      constexpr uint32_t w13 = L.w1_bytes + L.w3_bytes, stage = NC * L.blk_h;
      ptx::mbar_arrive_expect_tx(misc, w13 + L.w2_bytes + L.pol_img_bytes);
      if constexpr (POL) {
        const uint8_t* psrc = a.pol + (size_t)rank * L.pol_img_bytes;
        for (uint32_t o = 0; o < L.pol_img_bytes; o += 16384u) ptx::bulk_g2s(base + L.off_pol + o, psrc + o, min(16384u, L.pol_img_bytes - o), misc);
      }
      const uint8_t* src = a.res + (size_t)rank * w13;
      for (uint32_t o = 0; o < w13; o += 16384u) ptx::bulk_g2s(base + L.off_w13 + o, src + o, min(16384u, w13 - o), misc);
      for (int c = 0; c < NC; ++c) {           // streamed image of the pair mode: per chunk [rank][NC k-blocks]
        const uint8_t* s2 = a.stream + ((size_t)c * CG + rank) * stage;
        for (uint32_t o = 0; o < stage; o += 16384u) ptx::bulk_g2s(base + L.off_w2 + c * stage + o, s2 + o, min(16384u, stage - o), misc);
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 3) {
    const uint32_t t = warp == 1 ? 0u : 1u;     // tile pipe
    if (leader) {
      // ===================== MMA issuer of pipe t =====================
      const uint32_t ev = misc + kNwOffEv + t * kNwNumEv * 8u;
      const uint32_t cA1 = 256u * t, cD0 = cA1 + 128u, cX = cA1 + 192u, cD3 = cA1 + 224u;
      ptx::mbar_wait(misc, 0);
      const uint64_t a0desc = ptx::umma_desc_sw128(base + L.off_a0 + t * (uint32_t)kA0Bytes);
      const uint64_t w1desc = ptx::umma_desc_sw128(base + L.off_w13);
      const uint64_t w3desc = ptx::umma_desc_sw128(base + L.off_w13 + L.w1_bytes);
      const uint64_t w2desc = ptx::umma_desc_sw128(base + L.off_w2);
      for (int h = 0; h < Htot; ++h) {
        const uint32_t par = (uint32_t)h & 1u;
        if constexpr (POL) {
          // ---- proposal network: input (smem) -> tanh layer -> tanh layer -> action means (D3 columns 0..5) ----
          const uint64_t wpdesc = ptx::umma_desc_sw128(base + L.off_pol);
          const uint32_t np1 = (uint32_t)a.np1, np2 = (uint32_t)a.np2;
          ptx::mbar_wait(ev + 8u * kNwAp0Ready, par, hc);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            for (uint32_t j = 0; j < np1; ++j) {
#pragma unroll
              for (uint32_t ks = 0; ks < 2; ++ks)
                ptx::umma_ss<CG>(j ? cX : cD0, a0desc + 2u * ks, wpdesc + (uint64_t)(j * (L.blk_h >> 4) + 2u * ks), kIdescH, ks);
              ptx::umma_commit_cg<CG>(ev + 8u * (kNwPdFull + j));
            }
          }
          __syncwarp();
          if (np1 > 1) ptx::mbar_wait2(ev + 8u * kNwAp1Ready, ev + 8u * (kNwAp1Ready + 1), par, hc);
          else ptx::mbar_wait(ev + 8u * kNwAp1Ready, par, hc);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            for (uint32_t j = 0; j < np2; ++j) {
              for (uint32_t kb = 0; kb < np1; ++kb)
#pragma unroll
                for (uint32_t ks = 0; ks < 4; ++ks)
                  ptx::umma_ts<CG>(j ? cX : cD0, cA1 + kb * 32u + ks * 8u, wpdesc + (uint64_t)((2u + j * 2u + kb) * (L.blk_h >> 4) + 2u * ks), kIdescH,
                                   (kb | ks) ? 1u : 0u);
              ptx::umma_commit_cg<CG>(ev + 8u * (kNwPdFull + 2 + j));
            }
          }
          __syncwarp();
          if (np2 > 1) ptx::mbar_wait2(ev + 8u * kNwAp2Ready, ev + 8u * (kNwAp2Ready + 1), par, hc);
          else ptx::mbar_wait(ev + 8u * kNwAp2Ready, par, hc);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            for (uint32_t kb = 0; kb < np2; ++kb)
#pragma unroll
              for (uint32_t ks = 0; ks < 4; ++ks)
                ptx::umma_ts<CG>(cD3, cA1 + 64u + kb * 32u + ks * 8u, wpdesc + (uint64_t)(6u * (L.blk_h >> 4) + kb * (L.blk_o >> 4) + 2u * ks), kIdescO,
                                 (kb | ks) ? 1u : 0u);
            ptx::umma_commit_cg<CG>(ev + 8u * kNwPolFull);
          }
          __syncwarp();
        }
        ptx::mbar_wait(ev + 8u * kNwA0Ready, par, hc);
        // ---- layer 1: chunk c into D0 (even) / X (odd); chunks 2, 3 after chunks 0, 1 have been drained ----
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if (c >= 2) ptx::mbar_wait(ev + 8u * (kNwDEmpty1 + c - 2), par, hc);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            constexpr int nks = (kD + kA + 2 + 15) / 16;
#pragma unroll
            for (int ks = 0; ks < nks; ++ks)
              ptx::umma_ss<CG>((c & 1) ? cX : cD0, a0desc + 2u * ks, w1desc + (uint64_t)(c * (L.blk_h >> 4) + 2u * ks), kIdescH, ks > 0 ? 1u : 0u);
            ptx::umma_commit_cg<CG>(ev + 8u * (kNwDFull1 + c));
          }
          __syncwarp();
        }
        // ---- layer 2 chunk c into D0, then layer 3 for chunk c-1 (its hidden-2 chunk is stored by then) ----
#pragma unroll
        for (int c = 0; c <= NC; ++c) {
          if (c < NC) {
            if (c == 0) ptx::mbar_wait(ev + 8u * (kNwA1Ready + NC - 1), par, hc);   // chunks are stored in order: the last implies all
            else ptx::mbar_wait(ev + 8u * (kNwDEmpty2 + c - 1), par, hc);            // D0 drained
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int kb = 0; kb < NC; ++kb)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  ptx::umma_ts<CG>(cD0, cA1 + kb * 32 + ks * 8, w2desc + (uint64_t)((c * NC + kb) * (L.blk_h >> 4) + 2u * ks), kIdescH,
                                   (kb | ks) ? 1u : 0u);
              ptx::umma_commit_cg<CG>(ev + 8u * (kNwDFull2 + c));
            }
            __syncwarp();
          }
          if (c >= 1) {
            ptx::mbar_wait(ev + 8u * (kNwA2Ready + c - 1), par, hc);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                ptx::umma_ts<CG>(cD3, cX + ks * 8, w3desc + (uint64_t)((c - 1) * (L.blk_o >> 4) + 2u * ks), kIdescO, (c - 1 > 0 || ks > 0) ? 1u : 0u);
              ptx::umma_commit_cg<CG>(ev + 8u * (c == NC ? kNwD3Full : kNwA2Empty + c - 1));
            }
            __syncwarp();
          }
        }
      }
    } else if (warp == 1) {
      // ===================== peer CTA: tell the leader that this CTA's weight halves have landed =====================
      ptx::mbar_wait(misc, 0);
      if (ptx::elect_one()) ptx::mbar_arrive_remote(misc_leader);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== epilogue of pipe t: four warps, one per TMEM lane quadrant =====================
    const uint32_t t = (uint32_t)(warp - 4) >> 2;
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t ev = misc + kNwOffEv + t * kNwNumEv * 8u;
    const uint32_t ev_leader = misc_leader + kNwOffEv + t * kNwNumEv * 8u;
    const uint32_t cA1 = 256u * t, cD0 = cA1 + 128u, cX = cA1 + 192u, cD3 = cA1 + 224u;
    for (int it = 0; it < n_it; ++it) {      // ---- one work unit ----
    const int unit = first_unit + it * unit_stride;
    const int b = unit / units_per_state;
    const int tile = 4 * (unit % units_per_state) + 2 * (int)t + (int)rank;        // the pair's unit owns tiles 4 u + 2 t + rank
    const int k = tile * kRows + row;
    const bool valid = k < p.K;
    const int kc = valid ? k : p.K - 1;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t a0_row = base + L.off_a0 + t * (uint32_t)kA0Bytes + (uint32_t)row * 128u;
    // biases [b2 (64 NC)][b3 (32)], or with LayerNorm the correction vectors [u2][c2][u3][c3]
    const float* s_b2 = reinterpret_cast<const float*>(smem + L.off_vec);
    const float* s_b3 = s_b2 + 64 * NC;
    [[maybe_unused]] const float* s_u2 = s_b2;
    [[maybe_unused]] const float* s_c2 = s_b2 + 64 * NC;
    [[maybe_unused]] const float* s_u3 = s_b2 + 2 * 64 * NC;
    [[maybe_unused]] const float* s_c3 = s_u3 + kOutN;
    [[maybe_unused]] float ln_sum = 0.f, ln_sq = 0.f, rho = 1.f, rmu = 0.f;
    auto ln_finish = [&]() {          // statistics of the layer just completed -> rho = 1 / sqrt(var + eps), rmu = rho * mean
      const float mean = ln_sum * a.ln_inv_n;
      const float var = fmaxf(ln_sq * a.ln_inv_n - mean * mean, 0.f);
      rho = rsqrtf(var + a.ln_eps);
      rmu = rho * mean;
      ln_sum = 0.f; ln_sq = 0.f;
    };
    auto arrive = [&](uint32_t id) { ptx::mbar_arrive_remote(ev_leader + 8u * id); };
    float s[kD], z[kA], act0[kA], cost = 0.f;
#pragma unroll
    for (int d = 0; d < kD; ++d) s[d] = (float)root_state(p, (size_t)((DBG && p.per_row_states) ? kc : b), kD, d);
    get_noise<DBG, POL>(a, b, 0, kc, z);
    if constexpr (DBG) {
      if (p.states_out && valid)
        for (int d = 0; d < kD; ++d) p.states_out[(size_t)k * kD + d] = (double)s[d];
    }
    // one 64-column accumulator: TMEM -> registers (+ bias), ReLU + bf16, -> TMEM operand of the next layer
    auto chunk = [&](uint32_t par, uint32_t full_id, uint32_t empty_id, uint32_t col, uint32_t dst, uint32_t ready_id, int wait_id, const float* bias) {
      ptx::mbar_wait(ev + 8u * full_id, par, hc);
      ptx::tc_fence_after();
      uint32_t v0[32], v1[32], w0[16], w1[16];
      ptx::tmem_ld32(lane_off + col, v0);
      ptx::tmem_ld32(lane_off + col + 32u, v1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive(empty_id);
      if constexpr (LN) {
        // bias == nullptr: hidden 1 (plain ReLU of the accumulator); else hidden 2: bias points at u2 of this chunk, c2 follows
        const float4* u4 = reinterpret_cast<const float4*>(bias);
        const float4* c4 = reinterpret_cast<const float4*>(bias + 64 * NC);
        auto relu_stat = [&](float acc, float u, float cc) {
          float zv = bias ? fmaf(rho, acc, fmaf(-rmu, u, cc)) : acc;
          zv = zv < 0.f ? 0.f : zv;
          ln_sum += zv;
          ln_sq = fmaf(zv, zv, ln_sq);
          return zv;
        };
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 uu = make_float4(0.f, 0.f, 0.f, 0.f), cc = uu;
          if (bias) { uu = u4[i]; cc = c4[i]; }
          const float r0 = relu_stat(__uint_as_float(v0[4 * i]), uu.x, cc.x), r1 = relu_stat(__uint_as_float(v0[4 * i + 1]), uu.y, cc.y);
          const float r2 = relu_stat(__uint_as_float(v0[4 * i + 2]), uu.z, cc.z), r3 = relu_stat(__uint_as_float(v0[4 * i + 3]), uu.w, cc.w);
          w0[2 * i] = pack_bf16(r0, r1);
          w0[2 * i + 1] = pack_bf16(r2, r3);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 uu = make_float4(0.f, 0.f, 0.f, 0.f), cc = uu;
          if (bias) { uu = u4[8 + i]; cc = c4[8 + i]; }
          const float r0 = relu_stat(__uint_as_float(v1[4 * i]), uu.x, cc.x), r1 = relu_stat(__uint_as_float(v1[4 * i + 1]), uu.y, cc.y);
          const float r2 = relu_stat(__uint_as_float(v1[4 * i + 2]), uu.z, cc.z), r3 = relu_stat(__uint_as_float(v1[4 * i + 3]), uu.w, cc.w);
          w1[2 * i] = pack_bf16(r0, r1);
          w1[2 * i + 1] = pack_bf16(r2, r3);
        }
      } else {
        const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) bb = b4[i];
          w0[2 * i] = pack_relu_bf16(__uint_as_float(v0[4 * i]) + bb.x, __uint_as_float(v0[4 * i + 1]) + bb.y);
          w0[2 * i + 1] = pack_relu_bf16(__uint_as_float(v0[4 * i + 2]) + bb.z, __uint_as_float(v0[4 * i + 3]) + bb.w);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) bb = b4[8 + i];
          w1[2 * i] = pack_relu_bf16(__uint_as_float(v1[4 * i]) + bb.x, __uint_as_float(v1[4 * i + 1]) + bb.y);
          w1[2 * i + 1] = pack_relu_bf16(__uint_as_float(v1[4 * i + 2]) + bb.z, __uint_as_float(v1[4 * i + 3]) + bb.w);
        }
      }
      if (wait_id >= 0) {
        ptx::mbar_wait(ev + 8u * (uint32_t)wait_id, par, hc);       // layer 3 has consumed the previous hidden-2 chunk
        ptx::tc_fence_after();
      }
      ptx::tmem_st16(lane_off + dst, w0);
      ptx::tmem_st16(lane_off + dst + 16u, w1);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive(ready_id);
    };
    for (int h = 0; h < H; ++h) {
      const uint32_t par = (uint32_t)(it * H + h) & 1u;        // running step of this CTA: the barriers' phase
      float act[kA];
      if constexpr (POL) {
        // proposal network (ppo_bc_policy.py:60,70-72 / behavioral_cloning.py:21-26) on the tensor cores
        build_policy_input(a, a0_row, row, s);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive(kNwAp0Ready);
        const float* s_pv = reinterpret_cast<const float*>(smem + L.off_pol + L.pol_img_bytes);
        // one 64-unit chunk of a tanh layer: accumulator + bias -> tanh -> bf16 -> TMEM operand of the next policy layer
        auto pol_chunk = [&](int layer, int j) {
          const uint32_t col = j ? cX : cD0, dst = cA1 + (layer == 1 ? 0u : 64u) + 32u * (uint32_t)j;
          const float* bias = s_pv + (layer == 1 ? 0 : kPolMaxHidden) + 64 * j;
          ptx::mbar_wait(ev + 8u * (uint32_t)(kNwPdFull + (layer - 1) * 2 + j), par, hc);
          ptx::tc_fence_after();
          uint32_t v0[32], v1[32], w0[16], w1[16];
          ptx::tmem_ld32(lane_off + col, v0);
          ptx::tmem_ld32(lane_off + col + 32u, v1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            w0[i] = ptx::tanh_bf16x2(pack_bf16(__uint_as_float(v0[2 * i]) + bias[2 * i], __uint_as_float(v0[2 * i + 1]) + bias[2 * i + 1]));
#pragma unroll
          for (int i = 0; i < 16; ++i)
            w1[i] = ptx::tanh_bf16x2(pack_bf16(__uint_as_float(v1[2 * i]) + bias[32 + 2 * i], __uint_as_float(v1[2 * i + 1]) + bias[32 + 2 * i + 1]));
          ptx::tmem_st16(lane_off + dst, w0);
          ptx::tmem_st16(lane_off + dst + 16u, w1);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive((uint32_t)((layer == 1 ? kNwAp1Ready : kNwAp2Ready) + j));
        };
        for (int j = 0; j < a.np1; ++j) pol_chunk(1, j);
        for (int j = 0; j < a.np2; ++j) pol_chunk(2, j);
        ptx::mbar_wait(ev + 8u * kNwPolFull, par, hc);
        ptx::tc_fence_after();
        uint32_t vm[32];
        ptx::tmem_ld32(lane_off + cD3, vm);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
#pragma unroll
        for (int j = 0; j < kA; ++j) {
          const float mean = __uint_as_float(vm[j]) + s_pv[2 * kPolMaxHidden + j];
          const int prop = h == 0 ? a.proposal0 : a.proposal;           // (the tree-search planner switches the mode after its first stage)
          if (prop == BCMPC_PROP_POLICY_GAUSSIAN) act[j] = fmaf(a.pol_std[j], z[j], mean);                        // distributions.py:80
          else if (prop == BCMPC_PROP_POLICY_MIX) act[j] = fmaf(a.one_minus_explore, mean, a.explore * z[j]);   // controllers.py:208
          else if (prop == BCMPC_PROP_POLICY_MEAN) act[j] = mean;                                                // controllers.py:429,433
          else if (prop == BCMPC_PROP_UNIFORM) act[j] = z[j];                                                    // controllers.py:411 (random first stage)
          else act[j] = mean + z[j];                                                                             // controllers.pyc (BC + U[-1,1))
        }
      } else {
#pragma unroll
        for (int j = 0; j < kA; ++j) act[j] = z[j];                          // controllers.py:70: the sampled action itself
      }
      if (h == 0) {
#pragma unroll
        for (int j = 0; j < kA; ++j) act0[j] = act[j];
      }
      if constexpr (DBG) {
        if (p.actions_out && valid)
          for (int j = 0; j < kA; ++j) p.actions_out[((size_t)h * p.K + k) * kA + j] = (double)act[j];
      }
      build_input(a, a0_row, row, s, act);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive(kNwA0Ready);
#pragma unroll 1
      for (uint32_t c = 0; c < (uint32_t)NC; ++c)
        chunk(par, kNwDFull1 + c, kNwDEmpty1 + c, (c & 1u) ? cX : cD0, cA1 + 32u * c, kNwA1Ready + c, -1, nullptr);
      if constexpr (LN) ln_finish();                                        // statistics of hidden 1
      if (h + 1 < H) get_noise<DBG, POL>(a, b, h + 1, kc, z);              // next step's randomness, off the critical path
#pragma unroll 1
      for (uint32_t c = 0; c < (uint32_t)NC; ++c)
        chunk(par, kNwDFull2 + c, kNwDEmpty2 + c, cD0, cX, kNwA2Ready + c, c > 0 ? (int)(kNwA2Empty + c - 1) : -1, s_b2 + 64u * c);
      if constexpr (LN) ln_finish();                                        // statistics of hidden 2
      ptx::mbar_wait(ev + 8u * kNwD3Full, par, hc);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld32(lane_off + cD3, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      float delta[kD];
#pragma unroll
      for (int d = 0; d < kD; ++d) {
        const float y = LN ? fmaf(rho, __uint_as_float(v[d]), fmaf(-rmu, s_u3[d], s_c3[d])) : __uint_as_float(v[d]) + s_b3[d];
        delta[d] = fmaf(y, a.std_delta[d], a.mean_delta[d]);                 // dynamics.py:113
      }
      if (s[5] >= 0.2f) cost += 10.f;                                      // cost_functions.py:16-26 (pre-step state)
      if (s[6] >= 0.f) cost += 10.f;
      if (s[7] >= 0.f) cost += 10.f;
      cost -= delta[17] * 100.f;                                           // cost_functions.py:28 (/ 0.01)
#pragma unroll
      for (int d = 0; d < kD; ++d) s[d] += delta[d];                       // dynamics.py:116
      if constexpr (DBG) {
        if (p.states_out && valid)
          for (int d = 0; d < kD; ++d) p.states_out[((size_t)(h + 1) * p.K + k) * kD + d] = (double)s[d];
      }
    }
    // ---- per-tile argmin over the 128 rows (np.argmin order) ----
    if constexpr (DBG) {
      if (p.costs_out && valid) p.costs_out[(size_t)b * p.K + k] = (double)cost;
    }
    double c = (double)cost;
    int64_t idx = valid ? p.path_offset + k : INT64_MAX;
    int src = row;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double oc = __shfl_xor_sync(0xffffffffu, c, off);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, idx, off);
      const int os = __shfl_xor_sync(0xffffffffu, src, off);
      if (oi != INT64_MAX && (idx == INT64_MAX || cand_better(oc, oi, c, idx))) { c = oc; idx = oi; src = os; }
    }
    unsigned char* s_red = smem + L.off_misc + kNwOffRed + t * 128u;
    double* r_cost = reinterpret_cast<double*>(s_red);
    int64_t* r_idx = reinterpret_cast<int64_t*>(s_red + 32);
    int* r_src = reinterpret_cast<int*>(s_red + 64);
    if (lane == 0) { r_cost[q] = c; r_idx[q] = idx; r_src[q] = src; }
    ptx::named_bar_sync(1 + (int)t, 128);
    c = r_cost[0]; idx = r_idx[0]; src = r_src[0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (r_idx[i] != INT64_MAX && (idx == INT64_MAX || cand_better(r_cost[i], r_idx[i], c, idx))) { c = r_cost[i]; idx = r_idx[i]; src = r_src[i]; }
    if (row == src && p.cand) {
      Candidate& o = p.cand[(size_t)b * p.tiles_per_state + tile];
      o.cost = c;
      o.index = idx == INT64_MAX ? -1 : idx;
#pragma unroll
      for (int j = 0; j < kA; ++j) o.action[j] = (double)act0[j];
      for (int j = kA; j < kMaxAct; ++j) o.action[j] = 0.0;
    }
    if (p.out && (src >> 5) == q) fused_state_minloc(p, b, lane, src & 31);   // (src is the same in the pipe's 128 threads)
    }      // work units
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc_cg<CG>(0u, kTmemCols);
}

template <int NC, bool LN, bool POL, bool DBG>
inline cudaError_t nw_launch_t(const TmArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = rollout_narrow_kernel<NC, LN, POL, DBG>;
  constexpr NwLayout L = NwLay<NC, POL>::value;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a);
}
// grid.x = tiles / 2 (two tiles per CTA), tiles a multiple of 4 (two CTAs per pair); one translation unit per flavour
cudaError_t nw_launch_prod(int nc, bool ln, bool pol, const TmArgs& a, dim3 grid, cudaStream_t st);
cudaError_t nw_launch_dbg(int nc, bool ln, bool pol, const TmArgs& a, dim3 grid, cudaStream_t st);
inline cudaError_t nw_launch(int nc, bool ln, bool pol, const TmArgs& a, dim3 grid, cudaStream_t st) {
  return tm_needs_debug_flavour(a.p) ? nw_launch_dbg(nc, ln, pol, a, grid, st) : nw_launch_prod(nc, ln, pol, a, grid, st);
}

}  // namespace tmx
}  // namespace bcmpc
