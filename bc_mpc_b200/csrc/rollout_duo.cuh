// BF16 fast-mode rollout kernel for WIDE two-hidden-layer nets (hidden 257..510: the 2 x 500 net of BASELINE configs[0..3])
// when a launch has more work than one wave: TWO tiles per CTA take turns on ONE hidden-1 buffer.
//
// Why: in rollout_tmem_kernel a step is  layer 1 (~2500 cycles of hand-offs, tensor pipe almost idle)  ->  layer 2 (8 chunks
// x 1024 cycles, the pipe's own rate)  ->  tail: last hidden-2 chunk -> layer 3 -> state update -> next input (~2500 cycles,
// pipe idle).  A second independent tile would fill the idle stretches, but the bf16 hidden-1 activations of one 128-path
// tile are half of TMEM, and layer 2 needs all of them until its last chunk.  The hidden-1 buffer is, however, DEAD from the
// end of a tile's layer 2 until its next layer-1 epilogue - exactly the stretch in which that tile has nothing for the
// tensor pipe.  So two tiles (pipes) X, Y share it and alternate TURNS:
//
//     pipe X :  L1  [ L2 ........ ]  tail, state update, next input            L1  [ L2 ........ ]  tail ...
//     pipe Y :       tail, update, next input            L1  [ L2 ........ ]   tail, update, next input   L1 [ L2 ...
//     A1     :  X X X X X X X X X X X                    Y Y Y Y Y Y Y Y Y Y   X X X X X X X X X X X     Y Y
//
// While X runs layer 2, Y finishes its step (layer 3, cost, state update, next action, next input) - all of it off the
// tensor pipe's critical path.  What stays exposed per turn is the layer-1 phase.
//
// CTA pair (cta_group::2, M = 256 per MMA), 384 threads; per CTA:
//   TMEM  [  0,256)  A1   hidden-1 activations of the pipe whose turn it is (bf16 pairs, A operand of layer 2)
//         [256,384)  two 64-column fp32 accumulators; accumulator i is only ever written by issuer i (warp 1 / warp 3)
//         [384,448)  pipe 0: A2 (32: one hidden-2 chunk, A operand of layer 3) | D3 (32: output accumulator);  as ONE 64-column
//         [448,512)  pipe 1: the same                                            buffer also a layer-1 accumulator
//         (layer 1 also borrows three quarters of A1 itself as accumulators - see du_l1_col)
//   smem  two input tiles A0 (one per pipe), resident W1 / W3, ring of W2 stages (the images of rollout_tmem_kernel's CTA-pair
//         mode, unchanged), event barriers per pipe (each completes exactly once per pipe step: parity = pipe step & 1)
//   warp 0 weight producer; warps 1 / 3 MMA issuers (leader CTA): issuer i owns accumulator i - the layer-2 chunks c = i mod 2
//   of BOTH pipes in turn order, the layer-1 chunks that use its accumulator, and layer 3 of the chunks it issued (queued
//   right behind the NEXT chunk of the other issuer, so the single hidden-2 buffer is free again when that chunk completes);
//   warp 2 TMEM allocation; warps 4-7 epilogue of pipe 0 (one warp per TMEM lane quadrant), warps 8-11 epilogue of pipe 1.
//
// Cross-pipe ordering (everything else is per pipe):
//   * an issuer's first use of its accumulator in a turn waits for the other pipe's last drain of it (previous turn);
//   * a pipe's first hidden-1 store of a turn waits for the other pipe's last two layer-2 chunks to have COMPLETED (A1 free).
// Pipes draw their work units independently (pipe t: units t, t + 2, ... of the CTA pair's list); when the other pipe has run
// out, a pipe finishes alone (the cross-pipe waits are skipped: its own program order covers them).
//
// Results are bit-identical to rollout_tmem_kernel: same weight images, same k-order inside every accumulator, same epilogues.
// Replaces the same reference lines (dynamics.py:106-119, cost_functions.py:10-30,59-63, controllers.py:53,63-85).
#pragma once
#include "rollout_tmem.cuh"

namespace bcmpc {
namespace tmx {

constexpr int kDuColAcc = 256, kDuColOwn = 384;      // accumulator i: kDuColAcc + 64 i; pipe t's [A2 | D3]: kDuColOwn + 64 t
enum : int {
  kDuA0Ready = 0,      // input rows of the step are in shared memory               (epilogue -> issuers)
  kDuD3Full = 1,       // output accumulator complete                               (issuer -> epilogue)
  kDuA1Ready = 2,      // [8]  hidden-1 k-block c stored                            (epilogue -> issuers)
  kDuDFull = 10,       // [16] accumulator of use e complete (layer-1 chunk c -> c, layer-2 chunk c -> NC + c)
  kDuDEmpty = 26,      // [16] ... drained into registers
  kDuA2Ready = 42,     // [8]  hidden-2 chunk c stored                              (epilogue -> issuer)
  kDuA2Empty = 50,     // [8]  layer 3 has consumed hidden-2 chunk c                (issuer -> epilogue)
  kDuNumEv = 58
};
constexpr uint32_t kDuOffWFull = 0, kDuOffWEmpty = 128, kDuOffResFull = 256, kDuOffEv = 264, kDuOffTmemPtr = kDuOffEv + 2 * kDuNumEv * 8,
                   kDuOffRed = 1280, kDuMiscBytes = 1536;
static_assert(kDuOffTmemPtr + 16 <= kDuOffRed, "duo kernel: misc area");

struct DuLayout {
  uint32_t blk_h = 0, blk_o = 0, w1_bytes = 0, res_bytes = 0, stage_bytes = 0;
  int nst = 0, nst_r[2] = {0, 0}, first_r[2] = {0, 0};
  uint32_t off_a0 = 0, off_res = 0, off_vec = 0, off_ring = 0, off_misc = 0, smem_bytes = 0;
};
constexpr DuLayout du_layout(int nc, int var) {
  const Layout W = make_layout(2, nc, var & ~kVarPolicy);      // geometry of the weight images (CTA-pair mode of rollout_tmem_kernel)
  DuLayout L;
  L.blk_h = W.blk_h; L.blk_o = W.blk_o; L.w1_bytes = W.w1_bytes; L.res_bytes = W.res_bytes; L.stage_bytes = W.stage_bytes;
  L.off_a0 = 0;
  L.off_res = 2 * kA0Bytes;
  L.off_vec = (L.off_res + L.res_bytes + 1023u) / 1024u * 1024u;              // explicit biases [b2 (64 nc)][b3 (32)]
  const uint32_t vec_bytes = (var & kVarBias) ? (64u * nc + kOutN) * 4u : 0u;
  L.off_ring = (L.off_vec + vec_bytes + 1023u) / 1024u * 1024u;
  L.nst = (int)((kSmemLimit - kDuMiscBytes - L.off_ring) / L.stage_bytes);
  if (L.nst > kMaxStages) L.nst = kMaxStages;
  L.nst_r[0] = (L.nst + 1) / 2; L.nst_r[1] = L.nst / 2;
  L.first_r[0] = 0; L.first_r[1] = L.nst_r[0];
  L.off_misc = L.off_ring + L.nst * L.stage_bytes;
  L.smem_bytes = L.off_misc + kDuMiscBytes;
  return L;
}
template <int NC, int VAR>
struct DuLay {
  static constexpr DuLayout value = du_layout(NC, VAR);
};

// pipeline trace (-DBCMPC_TRACE, tools/duo_trace.py): CTA 0, per pipe and pipe step < kDuTrSteps
constexpr int kDuTrSteps = 6, kDuTrSlots = 56;      // + [16..23] L1 chunk issued, [24..31] L1 chunk seen, [32..39] L1 chunk stored, [40..47] L2 chunk seen, [48..55] L2 chunk issued
enum : int { kDuTrL1First = 0, kDuTrL1Last = 1, kDuTrL2First = 2, kDuTrL2Last = 3, kDuTrL3Last = 4, kDuTrA0Pub = 8, kDuTrA1Free = 9, kDuTrL1Seen = 10,
             kDuTrA1Done = 11, kDuTrL2Seen = 12, kDuTrL2Done = 13, kDuTrD3Seen = 14, kDuTrFinal = 15 };

// Layer 1 has eight 64-column chunks and (almost) nowhere to put them: its buffers are the pipe's own [A2 | D3] columns,
// the two accumulators, and two quarters of the A1 region itself - chunk c's bf16 output goes to A1[32c, 32c + 32), so an
// fp32 accumulator may sit in quarter Qk = A1[64k, 64k + 64) as long as it is drained before outputs 2k, 2k + 1 are stored.
// The pipe's own epilogue group converts chunks 0..4, the OTHER pipe's group (idle once it has published its next input)
// chunks 5..7: a TMEM load right after a TMEM store of the same warp takes ~250 cycles instead of ~65, so one group needs
// ~400 cycles per chunk.  The two groups work through their chunks in order, and no accumulator one group reads lies in a
// quarter the other group stores to (own group: Q0, Q1, lower Q2; helper: upper Q2, Q3).
//   chunk      0     1     2     3     4   |   5     6     7
//   buffer    own  acc0   Q1   acc1   own  |  Q3   acc0  acc1       (4, 6, 7 re-use the buffers of 0, 1, 3 once those are drained)
//   issuer     1     0     0     1     1   |   0     0     1
//   issued    EARLY: 0 as soon as the pipe's input is ready, 1 as soon as the other pipe's last use of accumulator 0 is drained -
//             both queue behind the other pipe's last layer-2 MMAs; 2 and 5 (one batch, one commit) when the other pipe's layer 2
//             has completed (A1 free); 3 when the other pipe's last use of accumulator 1 is drained
__host__ __device__ constexpr int du_l1_col(int c, int own) {
  return c == 0 || c == 4 ? own : c == 1 || c == 6 ? kDuColAcc : c == 3 || c == 7 ? kDuColAcc + 64 : c == 2 ? 64 : 192;
}
__host__ __device__ constexpr int du_last_chunk(int nc, int i) { return ((nc - 1) & 1) == i ? nc - 1 : nc - 2; }   // last layer-2 chunk of issuer i

template <int NC, int VAR, bool DBG>
__global__ void __launch_bounds__(kThreads, 1) rollout_duo_kernel(const __grid_constant__ TmArgs a) {
  constexpr int CG = 2;
  constexpr bool BIAS = (VAR & kVarBias) != 0, TANH = (VAR & kVarTanh) != 0;
  static_assert((VAR & ~(kVarBias | kVarTanh)) == 0 && (!TANH || BIAS), "duo kernel: uniform shooting, plain dynamics net");
  static_assert(NC >= 2 && NC <= 8, "duo kernel: 2..8 hidden chunks");
  constexpr DuLayout L = DuLay<NC, VAR>::value;
  static_assert(L.nst_r[1] >= 1, "duo kernel: weight ring too small");
  constexpr uint32_t kIdescH = ptx::umma_idesc_bf16(kRows * CG, kChunk), kIdescO = ptx::umma_idesc_bf16(kRows * CG, kOutN);

  extern __shared__ __align__(1024) uint8_t smem[];
  const PlanArgs& p = a.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  // work units as in rollout_tmem_kernel (a unit = the two tiles of one root state that share an MMA); this CTA pair's list is
  // first, first + stride, ...: pipe 0 takes its even entries, pipe 1 the odd ones
  const int units_per_state = p.tiles_per_state / CG;
  const int total_units = p.B * units_per_state;
  const int first_unit = (int)blockIdx.x / CG, unit_stride = (int)gridDim.x / CG;
  const int n_it = first_unit < total_units ? (total_units - first_unit + unit_stride - 1) / unit_stride : 0;
  const int H = p.H;
  const int S0 = ((n_it + 1) / 2) * H, S1 = (n_it / 2) * H;       // steps of pipe 0 / pipe 1 (S0 >= S1)
  const int Htot = S0 + S1;                                        // turns = weight passes
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t misc = base + L.off_misc;
  const uint32_t ev = misc + kDuOffEv;                              // event barriers of THIS CTA: [pipe][kDuNumEv]
  const uint32_t ev_leader = ptx::mapa(ev, 0);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.off_misc + kDuOffTmemPtr);
  const uint32_t hc = a.hint_chain;
#ifdef BCMPC_TRACE
  const bool tracing = a.trace != nullptr && blockIdx.x == 0;
#else
  constexpr bool tracing = false;
#endif
  auto stamp = [&](int t, int s, int slot) {
    if (tracing && s < kDuTrSteps) a.trace[(t * kDuTrSteps + s) * kDuTrSlots + slot] = clock64();
  };

  if (tid == 0) {
    if (base & 1023u) { printf("bcmpc: dynamic smem base not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < kMaxStages; ++i) {
      ptx::mbar_init(misc + kDuOffWFull + 8u * i, leader ? 2 : 1);
      ptx::mbar_init(misc + kDuOffWEmpty + 8u * i, 1);
    }
    ptx::mbar_init(misc + kDuOffResFull, leader ? 2 : 1);
    for (int t = 0; t < 2; ++t) {
      const uint32_t e = ev + (uint32_t)t * kDuNumEv * 8u;
      ptx::mbar_init(e + 8u * kDuA0Ready, 4 * CG);
      ptx::mbar_init(e + 8u * kDuD3Full, 1);
      for (int i = 0; i < 8; ++i) {
        ptx::mbar_init(e + 8u * (kDuA1Ready + i), 4 * CG);
        ptx::mbar_init(e + 8u * (kDuA2Ready + i), 4 * CG);
        ptx::mbar_init(e + 8u * (kDuA2Empty + i), 1);
      }
      for (int i = 0; i < 16; ++i) { ptx::mbar_init(e + 8u * (kDuDFull + i), 1); ptx::mbar_init(e + 8u * (kDuDEmpty + i), 4 * CG); }
    }
    ptx::fence_mbar_init();
  }
  if constexpr (BIAS) {
    float* s_vec = reinterpret_cast<float*>(smem + L.off_vec);
    for (int i = tid; i < 64 * NC + kOutN; i += kThreads) s_vec[i] = a.lnvec[i];
  }
  if (warp == 2) ptx::tmem_alloc_cg<CG>(ptx::smem_u32(s_tmem), kTmemCols);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  if (*s_tmem != 0u) { if (tid == 0) printf("bcmpc: unexpected TMEM base 0x%x\n", *s_tmem); __trap(); }

  if (warp == 0) {
    // ===================== weight producer (every CTA loads its own half): one pass over W2 per turn =====================
    if (ptx::elect_one()) {
      const uint8_t* src = a.res + (size_t)rank * L.res_bytes;
      ptx::mbar_arrive_expect_tx(misc + kDuOffResFull, L.res_bytes);
      for (uint32_t o = 0; o < L.res_bytes; o += 16384u)
        ptx::bulk_g2s(base + L.off_res + o, src + o, min(16384u, L.res_bytes - o), misc + kDuOffResFull);
    }
    __syncwarp();
    uint32_t stage[2] = {0, 0}, phase[2] = {0, 0};
    for (int n = 0; n < Htot; ++n) {
      const uint8_t* src = a.stream + (size_t)rank * L.stage_bytes;
      for (int c = 0; c < NC; ++c) {
        const int r = c & 1;
        const uint32_t st = (uint32_t)L.first_r[r] + stage[r];
        ptx::mbar_wait(misc + kDuOffWEmpty + 8u * st, phase[r] ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(misc + kDuOffWFull + 8u * st, L.stage_bytes);
#pragma unroll
          for (uint32_t o = 0; o < L.stage_bytes; o += 16384u)
            ptx::bulk_g2s(base + L.off_ring + st * L.stage_bytes + o, src + o, min(16384u, L.stage_bytes - o), misc + kDuOffWFull + 8u * st);
        }
        __syncwarp();
        src += (size_t)CG * L.stage_bytes;
        if (++stage[r] == (uint32_t)L.nst_r[r]) { stage[r] = 0; phase[r] ^= 1u; }
      }
    }
  } else if (warp == 1 || warp == 3) {
    if (leader) {
      // ===================== MMA issuer i: accumulator i, sub-ring i =====================
      const int i = warp == 1 ? 0 : 1;
      const uint32_t acc = (uint32_t)kDuColAcc + 64u * (uint32_t)i;
      ptx::mbar_wait(misc + kDuOffResFull, 0);
      const uint64_t w1desc = ptx::umma_desc_sw128(base + L.off_res);
      const uint64_t w3desc = ptx::umma_desc_sw128(base + L.off_res + L.w1_bytes);
      const uint32_t nst_mine = (uint32_t)L.nst_r[i], first_mine = (uint32_t)L.first_r[i];
      constexpr int kLast[2] = {du_last_chunk(NC, 0), du_last_chunk(NC, 1)};
      constexpr bool kHelper = NC > 5;                             // the other pipe's epilogue group stores part of hidden 1 (chunks 5, 6, 7)
      uint32_t gs = 0;                                             // weight stages this issuer has consumed
      constexpr int nks1 = (kD + kA + 2 + 15) / 16;
      // one layer-1 chunk of pipe t into TMEM columns d
      auto l1_chunk = [&](int t, [[maybe_unused]] int ts, int c, uint32_t d) {
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t a0desc = ptx::umma_desc_sw128(base + L.off_a0 + (uint32_t)t * (uint32_t)kA0Bytes);
#pragma unroll
          for (int ks = 0; ks < nks1; ++ks)
            ptx::umma_ss<CG>(d, a0desc + 2u * ks, w1desc + (uint64_t)((uint32_t)c * (L.blk_h >> 4) + 2u * ks), kIdescH, ks > 0 ? 1u : 0u);
          ptx::umma_commit_cg<CG>(ev + (uint32_t)t * kDuNumEv * 8u + 8u * (uint32_t)(kDuDFull + c));
          stamp(t, ts, 16 + c);
        }
        __syncwarp();
      };
      // the layer-1 chunks of turn (t, s) that can be issued EARLY, i.e. while the previous turn's layer 2 is still in the pipe
      auto l1_early = [&](int t, int s) {
        const uint32_t par = (uint32_t)s & 1u;
        const uint32_t evt = ev + (uint32_t)t * kDuNumEv * 8u, evo = ev + (uint32_t)(1 - t) * kDuNumEv * 8u;
        const bool has_prev = t == 1 || (s >= 1 && s - 1 < S1);
        const uint32_t prev_par = t == 1 ? par : par ^ 1u;
        ptx::mbar_wait(evt + 8u * kDuA0Ready, par, hc);
        if (i == 1) {
          if (ptx::elect_one()) stamp(t, s, kDuTrL1First);
          l1_chunk(t, s, 0, (uint32_t)kDuColOwn + 64u * (uint32_t)t);
        } else {
          if (has_prev) ptx::mbar_wait(evo + 8u * (uint32_t)(kDuDEmpty + NC + kLast[0]), prev_par, hc);   // the other pipe's last use of accumulator 0
          l1_chunk(t, s, 1, acc);
        }
      };
      int t = 0, s = 0;
      if (S0 > 0) l1_early(0, 0);
      while (s < S0) {
        // ---- one turn: pipe t, pipe step s ----
        const uint32_t par = (uint32_t)s & 1u;
        const uint32_t evt = ev + (uint32_t)t * kDuNumEv * 8u, evo = ev + (uint32_t)(1 - t) * kDuNumEv * 8u;
        const bool has_prev = t == 1 || (s >= 1 && s - 1 < S1);        // the turn before this one belonged to the other pipe
        const uint32_t prev_par = t == 1 ? par : par ^ 1u;
        const uint32_t own = (uint32_t)kDuColOwn + 64u * (uint32_t)t;
        // ---- the rest of layer 1 ----
        if (i == 0) {
          if (has_prev) ptx::mbar_wait2(evo + 8u * (uint32_t)(kDuDFull + NC + NC - 1), evo + 8u * (uint32_t)(kDuDFull + NC + NC - 2), prev_par, hc);   // A1 free
          ptx::tc_fence_after();
          if (ptx::elect_one()) {          // chunks 2 and 5 into quarters of A1: one batch, one commit
            const uint64_t a0desc = ptx::umma_desc_sw128(base + L.off_a0 + (uint32_t)t * (uint32_t)kA0Bytes);
#pragma unroll
            for (int c = 2; c <= 5; c += 3)
              if (c < NC) {
#pragma unroll
                for (int ks = 0; ks < nks1; ++ks)
                  ptx::umma_ss<CG>((uint32_t)du_l1_col(c, 0), a0desc + 2u * ks, w1desc + (uint64_t)((uint32_t)c * (L.blk_h >> 4) + 2u * ks), kIdescH, ks > 0 ? 1u : 0u);
              }
            ptx::umma_commit_cg<CG>(evt + 8u * (uint32_t)(kDuDFull + (NC > 5 ? 5 : 2)));
            stamp(t, s, 16 + 2);
          }
          __syncwarp();
          if constexpr (NC > 6) {
            ptx::mbar_wait(evt + 8u * (uint32_t)(kDuDEmpty + 1), par, hc);
            l1_chunk(t, s, 6, acc);
          }
        } else {
          if (has_prev) ptx::mbar_wait(evo + 8u * (uint32_t)(kDuDEmpty + NC + kLast[1]), prev_par, hc);   // the other pipe's last use of accumulator 1
          l1_chunk(t, s, 3, acc);
          ptx::mbar_wait(evt + 8u * (uint32_t)(kDuDEmpty + 0), par, hc);
          l1_chunk(t, s, 4, own);
          if constexpr (NC > 7) {
            ptx::mbar_wait(evt + 8u * (uint32_t)(kDuDEmpty + 3), par, hc);
            l1_chunk(t, s, 7, acc);
          }
          if (ptx::elect_one()) stamp(t, s, kDuTrL1Last);
          __syncwarp();
        }
        // ---- layer 2: my chunks; layer 3 of chunk c - 2 goes in right before chunk c ----
#pragma unroll 1
        for (int c = i; c < NC; c += 2) {
          const uint32_t stage = first_mine + gs % nst_mine, phase = (gs / nst_mine) & 1u;
          ++gs;
          ptx::mbar_wait(misc + kDuOffWFull + 8u * stage, phase);
          if (c < 2) {                                  // all of A1 stored (by the pipe's own group and its helper): every layer-1 accumulator drained
            if (kHelper && (t == 1 || s < S1)) ptx::mbar_wait2(evt + 8u * (uint32_t)(kDuA1Ready + NC - 1), evt + 8u * (uint32_t)(kDuA1Ready + NC - 2), par, hc);
            else ptx::mbar_wait(evt + 8u * (uint32_t)(kDuA1Ready + NC - 1), par, hc);
          }
          else ptx::mbar_wait(evt + 8u * (uint32_t)(kDuDEmpty + NC + c - 2), par, hc);
          ptx::tc_fence_after();
          const uint64_t bdesc = ptx::umma_desc_sw128(base + L.off_ring + stage * L.stage_bytes);
          constexpr int kHalf = (NC + 1) / 2;
          if (ptx::elect_one()) {
            if (c < 2) stamp(t, s, kDuTrL2First);
            stamp(t, s, 48 + c);
#pragma unroll
            for (int kb = 0; kb < kHalf; ++kb)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                ptx::umma_ts<CG>(acc, kColA1 + kb * 32 + ks * 8, bdesc + (uint64_t)(kb * (L.blk_h >> 4) + 2u * ks), kIdescH, (kb == 0 && ks == 0) ? 0u : 1u);
          }
          __syncwarp();
          if (c >= 2) {     // layer 3 of chunk c - 2 in the middle of chunk c: its hidden-2 store has had half a chunk of slack
            ptx::mbar_wait(evt + 8u * (uint32_t)(kDuA2Ready + c - 2), par, hc);
            ptx::tc_fence_after();
          }
          if (ptx::elect_one()) {
            if (c >= 2) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                ptx::umma_ts<CG>(own + 32u, own + ks * 8, w3desc + (uint64_t)((uint32_t)(c - 2) * (L.blk_o >> 4) + 2u * ks), kIdescO, (c - 2 == 0 && ks == 0) ? 0u : 1u);
              ptx::umma_commit_cg<CG>(evt + 8u * (uint32_t)(kDuA2Empty + c - 2));
            }
#pragma unroll
            for (int kb = kHalf; kb < NC; ++kb)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                ptx::umma_ts<CG>(acc, kColA1 + kb * 32 + ks * 8, bdesc + (uint64_t)(kb * (L.blk_h >> 4) + 2u * ks), kIdescH, 1u);
            ptx::umma_commit_cg<CG>(misc + kDuOffWEmpty + 8u * stage);
            ptx::umma_commit_cg<CG>(evt + 8u * (uint32_t)(kDuDFull + NC + c));
            stamp(t, s, kDuTrL2Last);
          }
          __syncwarp();
        }
        // ---- the next turn's early layer-1 chunks go in behind my last layer-2 chunk ----
        int tn = t, sn = s;
        if (tn == 0 && sn < S1) tn = 1; else { tn = 0; ++sn; }
        if (sn < S0 && tn != t) l1_early(tn, sn);
        // ---- layer 3 of my last chunk (the very last one completes the output accumulator) ----
        {
          const int cl = kLast[i];
          ptx::mbar_wait(evt + 8u * (uint32_t)(kDuA2Ready + cl), par, hc);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_ts<CG>(own + 32u, own + ks * 8, w3desc + (uint64_t)((uint32_t)cl * (L.blk_o >> 4) + 2u * ks), kIdescO, (cl == 0 && ks == 0) ? 0u : 1u);
            ptx::umma_commit_cg<CG>(evt + 8u * (uint32_t)(cl == NC - 1 ? kDuD3Full : kDuA2Empty + cl));
            stamp(t, s, kDuTrL3Last);
          }
          __syncwarp();
        }
        if (sn < S0 && tn == t) l1_early(tn, sn);      // alone: the pipe's next input needs this step's output first
        t = tn; s = sn;
      }
    } else if (warp == 1) {
      // ===================== peer CTA: tell the leader when this CTA's weight halves have landed =====================
      const uint32_t misc_leader = ptx::mapa(misc, 0);
      ptx::mbar_wait(misc + kDuOffResFull, 0);
      if (ptx::elect_one()) ptx::mbar_arrive_remote(misc_leader + kDuOffResFull);
      __syncwarp();
      uint32_t stage[2] = {0, 0}, phase[2] = {0, 0};
      for (int n = 0; n < Htot; ++n) {
        for (int c = 0; c < NC; ++c) {
          const int r = c & 1;
          const uint32_t st = (uint32_t)L.first_r[r] + stage[r];
          ptx::mbar_wait(misc + kDuOffWFull + 8u * st, phase[r]);
          if (ptx::elect_one()) ptx::mbar_arrive_remote(misc_leader + kDuOffWFull + 8u * st);
          __syncwarp();
          if (++stage[r] == (uint32_t)L.nst_r[r]) { stage[r] = 0; phase[r] ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue of pipe t: four warps, one per TMEM lane quadrant =====================
    const int t = (warp - 4) >> 2;
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t evt = ev + (uint32_t)t * kDuNumEv * 8u, evo = ev + (uint32_t)(1 - t) * kDuNumEv * 8u;
    const uint32_t evt_leader = ev_leader + (uint32_t)t * kDuNumEv * 8u;
    const uint32_t own = (uint32_t)kDuColOwn + 64u * (uint32_t)t;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t a0_row = base + L.off_a0 + (uint32_t)t * (uint32_t)kA0Bytes + (uint32_t)row * 128u;
    const bool tr = tracing && q == 0 && lane == 0;
    auto arrive = [&](uint32_t id) { ptx::mbar_arrive_remote(evt_leader + 8u * id); };
    [[maybe_unused]] const float* s_b2 = reinterpret_cast<const float*>(smem + L.off_vec);
    // Layer-1 epilogue of pipe tp (its step's parity parp), the chunks in `set` (see du_l1_col): accumulator -> activation -> bf16 ->
    // A1.  Only the drains of 0, 1, 3 (re-used buffers) and each group's last store are signalled.
    constexpr uint32_t kAllSet = (1u << NC) - 1u, kHelperSet = kAllSet & 0xE0u, kOwnerSet = kAllSet & 0x1Fu;
    auto helper_available = [&](int tp, int sp) { return tp == 0 ? sp < S1 : true; };   // is the other group at a step boundary when (tp, sp) runs?
    auto l1_epilogue = [&](int tp, uint32_t parp, uint32_t set, int ready_id, [[maybe_unused]] int tr_step) {
      const uint32_t evp = ev + (uint32_t)tp * kDuNumEv * 8u, evp_leader = ev_leader + (uint32_t)tp * kDuNumEv * 8u;
      const uint32_t ownp = (uint32_t)kDuColOwn + 64u * (uint32_t)tp;
      auto l1_cvt = [&](const uint32_t (&v)[32], uint32_t (&w)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if constexpr (TANH) w[j] = ptx::tanh_bf16x2(pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])));
          else w[j] = pack_relu_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        }
      };
      uint32_t v0[32], v1[32], w0[16], w1[16];
      int seen = -1;                                               // the last "accumulator complete" event this group has waited for
#pragma unroll 1
      for (int c = 0; c < NC; ++c) {
        if (!((set >> c) & 1u)) continue;
        const int e = (c == 2 && NC > 5) ? 5 : c;                 // (chunks 2 and 5 are one batch with one commit)
        if (e != seen) {
          if (c == 0 && ((set >> 1) & 1u)) { ptx::mbar_wait2(evp + 8u * (uint32_t)(kDuDFull + 0), evp + 8u * (uint32_t)(kDuDFull + 1), parp, hc); seen = 1; }
          else { ptx::mbar_wait(evp + 8u * (uint32_t)(kDuDFull + e), parp, hc); seen = e; }
          ptx::tc_fence_after();
        }
        const uint32_t col = (uint32_t)du_l1_col(c, (int)ownp);
        ptx::tmem_ld32(lane_off + col, v0);
        ptx::tmem_ld32(lane_off + col + 32u, v1);
        ptx::tmem_ld_wait();
        if (tr_step >= 0) stamp(tp, tr_step, 24 + c);
        if (c == 0 || (c == 1 && NC > 6) || (c == 3 && NC > 7)) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote(evp_leader + 8u * (uint32_t)(kDuDEmpty + c));
        }
        l1_cvt(v0, w0);
        l1_cvt(v1, w1);
        ptx::tmem_st16(lane_off + (uint32_t)kColA1 + 32u * (uint32_t)c, w0);
        ptx::tmem_st16(lane_off + (uint32_t)kColA1 + 32u * (uint32_t)c + 16u, w1);
        ptx::tmem_st_wait();
        if (tr_step >= 0) stamp(tp, tr_step, 32 + c);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_remote(evp_leader + 8u * (uint32_t)ready_id);
    };
    const int n_mine = t == 0 ? (n_it + 1) / 2 : n_it / 2;
    for (int itl = 0; itl < n_mine; ++itl) {      // ---- one work unit of this pipe ----
      const int unit = first_unit + (2 * itl + t) * unit_stride;
      const int b = unit / units_per_state, tile = (unit % units_per_state) * CG + (int)rank;
      const int k = tile * kRows + row;
      const bool valid = k < p.K;
      const int kc = valid ? k : p.K - 1;
      float s[kD], z[kA], act0[kA], cost = 0.f;
#pragma unroll
      for (int d = 0; d < kD; ++d) s[d] = (float)root_state(p, (size_t)((DBG && p.per_row_states) ? kc : b), kD, d);
      get_noise<DBG, false>(a, b, 0, kc, z);
      if constexpr (DBG) {
        if (p.states_out && valid)
          for (int d = 0; d < kD; ++d) p.states_out[(size_t)k * kD + d] = (double)s[d];
      }
      // one 64-column accumulator: TMEM -> registers, (bias) activation + bf16, -> TMEM operand of the next layer
      // wait_full: wait for the chunk's own "accumulator complete" event (false: an event already seen implies it - an issuer's
      // commit covers every MMA it issued before); empty_id / ready_id < 0: nobody waits for that hand-off, skip the arrive
      auto chunk = [&](uint32_t par, uint32_t e, bool wait_full, int empty_id, uint32_t col, uint32_t dst, int ready_id, int wait_id, const float* bias, [[maybe_unused]] int trs_step) {
        if (wait_full) {
          ptx::mbar_wait(evt + 8u * (kDuDFull + e), par, hc);
          ptx::tc_fence_after();
        }
        if (trs_step >= 0) stamp(t, trs_step, (e < (uint32_t)NC ? 24 : 40 - NC) + (int)e);
        uint32_t v0[32], v1[32], w0[16], w1[16];
        ptx::tmem_ld32(lane_off + col, v0);
        ptx::tmem_ld32(lane_off + col + 32u, v1);
        ptx::tmem_ld_wait();
        if (empty_id >= 0) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive((uint32_t)empty_id);
        }
        if constexpr (BIAS) {
          const float4* b4 = reinterpret_cast<const float4*>(bias);
          auto act2 = [&](float x, float y) {
            if constexpr (TANH) return ptx::tanh_bf16x2(pack_bf16(x, y));
            else return pack_relu_bf16(x, y);
          };
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) bb = b4[j];
            w0[2 * j] = act2(__uint_as_float(v0[4 * j]) + bb.x, __uint_as_float(v0[4 * j + 1]) + bb.y);
            w0[2 * j + 1] = act2(__uint_as_float(v0[4 * j + 2]) + bb.z, __uint_as_float(v0[4 * j + 3]) + bb.w);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) bb = b4[8 + j];
            w1[2 * j] = act2(__uint_as_float(v1[4 * j]) + bb.x, __uint_as_float(v1[4 * j + 1]) + bb.y);
            w1[2 * j + 1] = act2(__uint_as_float(v1[4 * j + 2]) + bb.z, __uint_as_float(v1[4 * j + 3]) + bb.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) w0[j] = pack_relu_bf16(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
#pragma unroll
          for (int j = 0; j < 16; ++j) w1[j] = pack_relu_bf16(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
        }
        if (wait_id >= 0) {
          ptx::mbar_wait(evt + 8u * (uint32_t)wait_id, par, hc);        // layer 3 has consumed the previous hidden-2 chunk
          ptx::tc_fence_after();
        }
        ptx::tmem_st16(lane_off + dst, w0);
        ptx::tmem_st16(lane_off + dst + 16u, w1);
        if (ready_id >= 0) {
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive((uint32_t)ready_id);
        }
        if (trs_step >= 0 && e < (uint32_t)NC) stamp(t, trs_step, 32 + (int)e);
      };
      for (int h = 0; h < H; ++h) {
        const int sp = itl * H + h;                                   // this pipe's running step: the barriers' phase
        const uint32_t par = (uint32_t)sp & 1u;
        const bool has_prev = t == 1 || (sp >= 1 && sp - 1 < S1);
        const uint32_t prev_par = t == 1 ? par : par ^ 1u;
        const bool trs = tr && sp < kDuTrSteps;
        if (h == 0) {
#pragma unroll
          for (int j = 0; j < kA; ++j) act0[j] = z[j];
        }
        if constexpr (DBG) {
          if (p.actions_out && valid)
            for (int j = 0; j < kA; ++j) p.actions_out[((size_t)h * p.K + k) * kA + j] = (double)z[j];
        }
        build_input(a, a0_row, row, s, z);                            // controllers.py:70: the sampled action itself
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive(kDuA0Ready);
        if (trs) stamp(t, sp, kDuTrA0Pub);
        // lend a hand with the layer-1 epilogue of the turn before mine (the other pipe's): pipe 1 at the top of its step sp helps
        // pipe 0's step sp, pipe 0 at the top of its step sp >= 1 helps pipe 1's step sp - 1
        if (kHelperSet != 0u) {
          if (t == 1) l1_epilogue(0, par, kHelperSet, kDuA1Ready + NC - 2, -1);
          else if (sp >= 1 && sp - 1 < S1) l1_epilogue(1, par ^ 1u, kHelperSet, kDuA1Ready + NC - 2, -1);
        }
        if (h + 1 < H) get_noise<DBG, false>(a, b, h + 1, kc, z);     // the next step's randomness: this pipe is waiting for its turn anyway
        if (has_prev) {                                               // the other pipe's layer 2 has read the last of ITS hidden 1: A1 is mine
          ptx::mbar_wait2(evo + 8u * (uint32_t)(kDuDFull + NC + NC - 1), evo + 8u * (uint32_t)(kDuDFull + NC + NC - 2), prev_par, hc);
          ptx::tc_fence_after();
        }
        if (trs) stamp(t, sp, kDuTrA1Free);
        l1_epilogue(t, par, helper_available(t, sp) ? kOwnerSet : kAllSet, kDuA1Ready + NC - 1, trs ? sp : -1);
        if (trs) stamp(t, sp, kDuTrA1Done);
#pragma unroll 1
        for (uint32_t c = 0; c < (uint32_t)NC; ++c) {
          chunk(par, (uint32_t)NC + c, true, (int)(kDuDEmpty + NC + c), (uint32_t)kDuColAcc + 64u * (c & 1u), own, (int)(kDuA2Ready + c), c > 0 ? (int)(kDuA2Empty + c - 1) : -1,
                BIAS ? s_b2 + 64u * c : nullptr, trs ? sp : -1);
          if (trs && c == 0) stamp(t, sp, kDuTrL2Seen);
        }
        if (trs) stamp(t, sp, kDuTrL2Done);
        ptx::mbar_wait(evt + 8u * kDuD3Full, par, hc);
        ptx::tc_fence_after();
        if (trs) stamp(t, sp, kDuTrD3Seen);
        uint32_t v[32];
        ptx::tmem_ld32(lane_off + own + 32u, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        float delta[kD];
        if constexpr (BIAS) {
          const float* s_b3 = s_b2 + 64 * NC;
#pragma unroll
          for (int d = 0; d < kD; ++d) v[d] = __float_as_uint(__uint_as_float(v[d]) + s_b3[d]);
        }
#pragma unroll
        for (int d = 0; d < kD; ++d) delta[d] = fmaf(__uint_as_float(v[d]), a.std_delta[d], a.mean_delta[d]);   // dynamics.py:113
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
        if (trs) stamp(t, sp, kDuTrFinal);
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
      unsigned char* s_red = smem + L.off_misc + kDuOffRed + t * 128;
      double* r_cost = reinterpret_cast<double*>(s_red);
      int64_t* r_idx = reinterpret_cast<int64_t*>(s_red + 32);
      int* r_src = reinterpret_cast<int*>(s_red + 64);
      ptx::named_bar_sync(1 + t, 128);                              // (the previous unit's readers of the scratch are done)
      if (lane == 0) { r_cost[q] = c; r_idx[q] = idx; r_src[q] = src; }
      ptx::named_bar_sync(1 + t, 128);
      c = r_cost[0]; idx = r_idx[0]; src = r_src[0];
#pragma unroll
      for (int j = 1; j < 4; ++j)
        if (r_idx[j] != INT64_MAX && (idx == INT64_MAX || cand_better(r_cost[j], r_idx[j], c, idx))) { c = r_cost[j]; idx = r_idx[j]; src = r_src[j]; }
      if (row == src && p.cand) {
        Candidate& o = p.cand[(size_t)b * p.tiles_per_state + tile];
        o.cost = c;
        o.index = idx == INT64_MAX ? -1 : idx;
#pragma unroll
        for (int j = 0; j < kA; ++j) o.action[j] = (double)act0[j];
        for (int j = kA; j < kMaxAct; ++j) o.action[j] = 0.0;
      }
      if (p.out && (src >> 5) == q) fused_state_minloc(p, b, lane, src & 31);
    }      // work units
    if (kHelperSet != 0u && t == 0 && S0 >= 1 && S0 - 1 < S1) l1_epilogue(1, (uint32_t)(S0 - 1) & 1u, kHelperSet, kDuA1Ready + NC - 2, -1);   // pipe 1's last step comes after mine
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc_cg<CG>(0u, kTmemCols);
}

template <int NC, int VAR, bool DBG>
inline cudaError_t du_launch_t(const TmArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = rollout_duo_kernel<NC, VAR, DBG>;
  constexpr DuLayout L = DuLay<NC, VAR>::value;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_bytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
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
// hidden chunks 5..8 (hidden 257..510 with the biases in the GEMMs; below that rollout_narrow_kernel keeps two tiles' hidden 1
// in TMEM side by side); variants: 0 (ReLU, biases in the GEMMs), kVarBias | kVarTanh (plain tanh net)
inline bool du_supported(int nc, int var) { return nc >= 5 && nc <= 8 && (var == 0 || var == (kVarBias | kVarTanh)); }
cudaError_t du_launch_prod(int nc, int var, const TmArgs& a, dim3 grid, cudaStream_t st);
cudaError_t du_launch_dbg(int nc, int var, const TmArgs& a, dim3 grid, cudaStream_t st);
inline cudaError_t du_launch(int nc, int var, const TmArgs& a, dim3 grid, cudaStream_t st) {
  return tm_needs_debug_flavour(a.p) ? du_launch_dbg(nc, var, a, grid, st) : du_launch_prod(nc, var, a, grid, st);
}

}  // namespace tmx
}  // namespace bcmpc
