// BF16 fast-mode fused rollout kernel (BCMPC_PREC_BF16_FAST): tcgen05.mma + TMEM + bulk-async weight
// streaming.  One CTA owns 128 paths (UMMA M = 128) for all H steps; the 20-float state, the running
// cost and the first action of every path live in registers of the thread that owns the TMEM lane, so
// no intermediate state ever touches HBM.
//
// Per step and tile (hidden width S padded to a multiple of 128, NC = S/128 chunks, NKB = S/64 k-blocks):
//   L1  D[128 x 128] = A0[128 x 32]  x W1 chunk         (2 UMMA k-steps)       -> epilogue: +b1, ReLU, bf16 -> A1 (smem)
//   L2  D[128 x 128] = A1[128 x S ]  x W2 chunk         (NKB x 4 k-steps)      -> epilogue: +b2, ReLU, bf16 -> A2 ring (smem)
//   L3  D3[128 x 32] += A2 chunk     x W3 k-blocks      (accumulated as the ring fills; A2 never fully exists)
//   final epilogue: +b3, denormalise, cheetah cost on the pre-step state, s += delta, next action (Philox),
//                   normalise -> bf16 A0 for the next step.
// Roles: warp 0 lane 0 streams the weight "tape" (pre-swizzled UMMA smem images, L2 resident) through a
// ring of 16 KB stages with cp.async.bulk + mbarrier; warp 1 lane 0 issues every tcgen05.mma by walking a
// host-built op table; warp 2 owns the TMEM allocation; warps 4-11 are the epilogue (two warps per TMEM
// lane quadrant, 64 accumulator columns each).  Accumulators rotate through three 128-column TMEM buffers
// so the epilogue of chunk i overlaps the MMAs of chunk i+1.
//
// Replaces dynamics.py:106-119 + cost_functions.py:10-30,59-63 + controllers.py:53,63-85 for HalfCheetah
// shapes (D = 20, A = 6), two hidden ReLU layers, S <= 512, uniform shooting (P0).
#pragma once
#include <cuda_bf16.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace bcmpc {
namespace tc {

constexpr int kRows = 128;
constexpr int kBlkBytes = 128 * 128;   // one [128 rows x 64 bf16] swizzled block
constexpr int kThreads = 384;
constexpr int kOutN = 32;              // output layer padded to UMMA N = 32
constexpr int kD3Col = 384;
constexpr int kTmemCols = 512;
constexpr int kNumDBuf = 3;
constexpr int kD = 20, kA = 6;
constexpr int kMaxOps = 64;
constexpr int kSmemLimit = 232448;     // 227 KB

// barrier ids
constexpr int kBarWFull = 0, kBarWEmpty = 8, kBarA0Ready = 16, kBarD3Full = 17, kBarA2Ready = 18, kBarA2Empty = 20,
              kBarDFull = 22, kBarDEmpty = 25, kBarA1Ready = 32, kNumBars = 48;

enum : uint8_t { kOpZero = 1, kOpFirst = 2, kOpLast = 4, kOpL3 = 8, kOpD3Last = 16 };

struct TcOp {            // 16 bytes
  uint32_t w_off16;      // tape offset / 16
  uint16_t w_bytes16;    // bytes / 16
  uint16_t a_off16;      // A operand offset in dynamic smem / 16
  uint8_t n8;            // UMMA N / 8
  uint8_t ksteps;        // UMMA K steps (16 bf16 each)
  uint8_t flags;
  uint8_t chunk_id;      // hidden chunk number within the step (D buffer = (step*2NC + chunk_id) % 3)
  uint8_t wait_bar;      // barrier the A operand depends on, 0xFF = none
  uint8_t commit_bar;    // extra barrier released when this op's MMAs complete, 0xFF = none
  uint8_t pad[2];
};
static_assert(sizeof(TcOp) == 16, "TcOp layout");

struct TcArgs {
  PlanArgs p;
  const uint8_t* tape;
  const TcOp* ops;
  int n_ops, nkb, nc, nst;
  uint32_t off_a1, off_ring2, off_wring, off_misc;
  const float* b1;   // [S_pad]
  const float* b2;   // [S_pad]
  float mean_obs[kMaxObs], inv_den_obs[kMaxObs], mean_delta[kMaxObs], std_delta[kMaxObs], b3[kMaxObs];
  float mean_act[kMaxAct], inv_den_act[kMaxAct], low[kMaxAct], span[kMaxAct];
};

// ---------------------------------------------------------------------------------------------------------
// host: weight tape + op table
// ---------------------------------------------------------------------------------------------------------
inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return (uint16_t)((u >> 16) | 0x40);   // NaN
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// element (row n, k-column kk) of a K-major SWIZZLE_128B tile image (8-row groups of 1024 B)
inline size_t sw128_offset(int n, int kk) {
  return (size_t)(n / 8) * 1024 + (size_t)(n % 8) * 128 + (size_t)(((kk / 8) ^ (n % 8)) * 16) + (size_t)(kk % 8) * 2;
}

struct TcHostPlan {
  std::vector<uint16_t> tape;
  std::vector<TcOp> ops;
  std::vector<float> b1, b2;
  float b3[kMaxObs];
  int s_pad = 0, nkb = 0, nc = 0, nst = 0;
  uint32_t off_a1 = 0, off_ring2 = 0, off_wring = 0, off_misc = 0, smem_bytes = 0;
};

inline bool tc_supported(const bcmpc_config& c, std::string* why) {
  if (c.obs_dim != kD || c.act_dim != kA) { *why = "bf16 fast path is specialised for obs_dim=20, act_dim=6 (HalfCheetah)"; return false; }
  if (c.n_hidden != 2) { *why = "bf16 fast path supports n_hidden == 2 this round"; return false; }
  if (c.hidden_size < 1 || c.hidden_size > 512) { *why = "bf16 fast path supports hidden_size <= 512 this round"; return false; }
  if (c.activation != BCMPC_ACT_RELU) { *why = "bf16 fast path supports relu only"; return false; }
  if (c.proposal != BCMPC_PROP_UNIFORM) { *why = "bf16 fast path supports uniform shooting (P0) this round; use BCMPC_PREC_FP32_VERIFY for policy proposals"; return false; }
  return true;
}

// W[l] are TF kernels [in,out] row-major: W1 [26,S], W2 [S,S], W3 [S,20].  schedule 0 = layer by layer,
// 1 = L1 chunks interleaved with the first L2 chunk (hides the L1 epilogues under L2 MMAs).
inline void tc_build(const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                     const float* b3, int S, int schedule, TcHostPlan* out) {
  TcHostPlan& P = *out;
  const int in_dim = kD + kA;
  P.s_pad = (S + 127) / 128 * 128;
  P.nkb = P.s_pad / 64;
  P.nc = P.s_pad / 128;
  const int NC = P.nc, NKB = P.nkb;
  P.off_a1 = 0;
  P.off_ring2 = P.off_a1 + NKB * kBlkBytes;
  P.off_wring = P.off_ring2 + 2 * kBlkBytes;
  const uint32_t misc_bytes = 3072;
  P.nst = (int)((kSmemLimit - misc_bytes - P.off_wring) / kBlkBytes);
  if (P.nst > 8) P.nst = 8;
  P.off_misc = P.off_wring + P.nst * kBlkBytes;
  P.smem_bytes = P.off_misc + misc_bytes;
  if (schedule == 1 && NC > 2) {
    for (int j = 2; j < NC; ++j)
      if (j % kNumDBuf == NC % kNumDBuf) schedule = 0;   // reordering would swap two users of one TMEM buffer
  }
  P.b1.assign(P.s_pad, 0.f);
  P.b2.assign(P.s_pad, 0.f);
  for (int i = 0; i < S; ++i) { P.b1[i] = b1[i]; P.b2[i] = b2[i]; }
  for (int i = 0; i < kMaxObs; ++i) P.b3[i] = i < kD ? b3[i] : 0.f;

  P.tape.clear();
  P.ops.clear();
  auto new_tile = [&](int bytes) { size_t off = P.tape.size() * 2; P.tape.resize(P.tape.size() + bytes / 2, 0); return off; };
  const uint16_t a0_off16 = (uint16_t)(P.off_ring2 / 16);   // A0 aliases ring slot 0 (see kernel)

  auto emit_l1 = [&](int nc) {
    const size_t off = new_tile(kBlkBytes);
    for (int i = 0; i < 128; ++i) {
      const int n = nc * 128 + i;
      if (n >= S) continue;
      for (int k = 0; k < in_dim; ++k) P.tape[(off + sw128_offset(i, k)) / 2] = f32_to_bf16(W1[(size_t)k * S + n]);
    }
    TcOp op{};
    op.w_off16 = (uint32_t)(off / 16); op.w_bytes16 = kBlkBytes / 16; op.a_off16 = a0_off16;
    op.n8 = 128 / 8; op.ksteps = (uint8_t)((in_dim + 15) / 16);
    op.flags = kOpZero | kOpFirst | kOpLast; op.chunk_id = (uint8_t)nc;
    op.wait_bar = nc == 0 ? (uint8_t)kBarA0Ready : 0xFF; op.commit_bar = 0xFF;
    P.ops.push_back(op);
  };
  auto emit_l2 = [&](int nc, int kb) {
    const size_t off = new_tile(kBlkBytes);
    for (int i = 0; i < 128; ++i) {
      const int n = nc * 128 + i;
      if (n >= S) continue;
      for (int kk = 0; kk < 64; ++kk) {
        const int k = kb * 64 + kk;
        if (k < S) P.tape[(off + sw128_offset(i, kk)) / 2] = f32_to_bf16(W2[(size_t)k * S + n]);
      }
    }
    TcOp op{};
    op.w_off16 = (uint32_t)(off / 16); op.w_bytes16 = kBlkBytes / 16;
    op.a_off16 = (uint16_t)((P.off_a1 + kb * kBlkBytes) / 16);
    op.n8 = 128 / 8; op.ksteps = 4;
    op.flags = (uint8_t)((kb == 0 ? (kOpZero | kOpFirst) : 0) | (kb == NKB - 1 ? kOpLast : 0));
    op.chunk_id = (uint8_t)(NC + nc);
    op.wait_bar = nc == 0 ? (uint8_t)(kBarA1Ready + kb) : 0xFF; op.commit_bar = 0xFF;
    P.ops.push_back(op);
  };
  auto emit_l3 = [&](int nc) {
    for (int hh = 0; hh < 2; ++hh) {
      const int kb = 2 * nc + hh;
      const int bytes = kOutN * 128;
      const size_t off = new_tile(bytes);
      for (int n = 0; n < kD; ++n)
        for (int kk = 0; kk < 64; ++kk) {
          const int k = kb * 64 + kk;
          if (k < S) P.tape[(off + sw128_offset(n, kk)) / 2] = f32_to_bf16(W3[(size_t)k * kD + n]);
        }
      TcOp op{};
      op.w_off16 = (uint32_t)(off / 16); op.w_bytes16 = (uint16_t)(bytes / 16);
      op.a_off16 = (uint16_t)((P.off_ring2 + hh * kBlkBytes) / 16);
      op.n8 = kOutN / 8; op.ksteps = 4;
      op.flags = (uint8_t)(kOpL3 | (kb == 0 ? kOpZero : 0) | (kb == NKB - 1 ? kOpD3Last : 0));
      op.chunk_id = 0;
      op.wait_bar = (uint8_t)(kBarA2Ready + hh); op.commit_bar = (uint8_t)(kBarA2Empty + hh);
      P.ops.push_back(op);
    }
  };

  if (schedule == 0) {
    for (int nc = 0; nc < NC; ++nc) emit_l1(nc);
    for (int kb = 0; kb < NKB; ++kb) emit_l2(0, kb);
  } else {
    emit_l1(0);
    int kb = 0;
    for (int j = 1; j < NC; ++j) { emit_l1(j); emit_l2(0, kb++); emit_l2(0, kb++); }
    while (kb < NKB) emit_l2(0, kb++);
  }
  for (int nc = 1; nc < NC; ++nc) {
    for (int kb = 0; kb < NKB; ++kb) emit_l2(nc, kb);
    emit_l3(nc - 1);
  }
  emit_l3(NC - 1);
}

// ---------------------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// normalise (dynamics.py:109-110) -> bf16 -> the first four 16-byte chunks of this row of A0
__device__ __forceinline__ void build_input(const TcArgs& a, uint32_t a0_row_addr, int row, const float (&s)[kD],
                                            const float (&act)[kA]) {
  float x[32];
#pragma unroll
  for (int c = 0; c < kD; ++c) x[c] = (s[c] - a.mean_obs[c]) * a.inv_den_obs[c];
#pragma unroll
  for (int j = 0; j < kA; ++j) x[kD + j] = (act[j] - a.mean_act[j]) * a.inv_den_act[j];
#pragma unroll
  for (int c = kD + kA; c < 32; ++c) x[c] = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t addr = a0_row_addr + (uint32_t)(((c ^ (row & 7)) << 4));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16(x[8 * c + 0], x[8 * c + 1])),
                 "r"(pack_bf16(x[8 * c + 2], x[8 * c + 3])), "r"(pack_bf16(x[8 * c + 4], x[8 * c + 5])),
                 "r"(pack_bf16(x[8 * c + 6], x[8 * c + 7]))
                 : "memory");
  }
}

// the action of (path, step): injected tensor or Philox draw (controllers.py:53)
__device__ __forceinline__ void get_action(const TcArgs& a, int b, int h, int k, float (&act)[kA]) {
  const PlanArgs& p = a.p;
  if (p.inj) {
    const double* src = p.inj + (((size_t)b * p.H + h) * p.K + k) * kA;
#pragma unroll
    for (int j = 0; j < kA; ++j) act[j] = (float)src[j];
  } else {
    float v[kMaxAct];
    draw_injection(v, kA, kInjUniformBox, p.seed, (uint32_t)(p.path_offset + k), (uint32_t)b, (uint32_t)h, a.low, a.span);
#pragma unroll
    for (int j = 0; j < kA; ++j) act[j] = v[j];
  }
}

// One hidden-layer epilogue: this thread's row, 64 accumulator columns -> +bias, ReLU, bf16 -> one 128-byte
// swizzled row of the destination k-block.
__device__ __forceinline__ void hidden_epilogue(uint32_t bars, uint32_t tmem_base, int q, int hh, int row, uint32_t seq,
                                                const float* __restrict__ bias, uint32_t dst_block, uint32_t ready_bar,
                                                int empty_bar, uint32_t empty_parity) {
  const int buf = seq % kNumDBuf;
  const uint32_t par = (seq / kNumDBuf) & 1u;
  ptx::mbar_wait(bars + 8u * (kBarDFull + buf), par);
  ptx::tc_fence_after();
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + hh * 64);
  uint32_t v0[32], v1[32];
  ptx::tmem_ld32(taddr, v0);
  ptx::tmem_ld32(taddr + 32, v1);
  ptx::tmem_ld_wait();
  ptx::tc_fence_before();
  ptx::mbar_arrive(bars + 8u * (kBarDEmpty + buf));     // accumulator is in registers: release the TMEM buffer
  if (empty_bar >= 0) ptx::mbar_wait(bars + 8u * empty_bar, empty_parity);
  const uint32_t row_addr = dst_block + (uint32_t)row * 128u;
  const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 ba = __ldg(b4 + 2 * c), bb = __ldg(b4 + 2 * c + 1);
    const uint32_t* v = c < 4 ? &v0[8 * c] : &v1[8 * (c - 4)];
    const float e0 = fmaxf(__uint_as_float(v[0]) + ba.x, 0.f), e1 = fmaxf(__uint_as_float(v[1]) + ba.y, 0.f);
    const float e2 = fmaxf(__uint_as_float(v[2]) + ba.z, 0.f), e3 = fmaxf(__uint_as_float(v[3]) + ba.w, 0.f);
    const float e4 = fmaxf(__uint_as_float(v[4]) + bb.x, 0.f), e5 = fmaxf(__uint_as_float(v[5]) + bb.y, 0.f);
    const float e6 = fmaxf(__uint_as_float(v[6]) + bb.z, 0.f), e7 = fmaxf(__uint_as_float(v[7]) + bb.w, 0.f);
    const uint32_t addr = row_addr + (uint32_t)(((c ^ (row & 7)) << 4));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16(e0, e1)), "r"(pack_bf16(e2, e3)),
                 "r"(pack_bf16(e4, e5)), "r"(pack_bf16(e6, e7))
                 : "memory");
  }
  ptx::fence_proxy_async_smem();
  ptx::mbar_arrive(ready_bar);
}

__global__ void __launch_bounds__(kThreads, 1) rollout_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const PlanArgs& p = a.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, b = blockIdx.y;
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t bars = base + a.off_misc;                               // kNumBars x 8 B
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + a.off_misc + kNumBars * 8);
  TcOp* s_ops = reinterpret_cast<TcOp*>(smem + a.off_misc + kNumBars * 8 + 16);
  unsigned char* s_red = smem + a.off_misc + kNumBars * 8 + 16 + kMaxOps * 16;   // argmin scratch
  const int NC = a.nc, H = p.H;

  if (tid == 0) {
    if (base & 1023u) { printf("bcmpc: dynamic smem base not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < 8; ++i) { ptx::mbar_init(bars + 8u * (kBarWFull + i), 1); ptx::mbar_init(bars + 8u * (kBarWEmpty + i), 1); }
    ptx::mbar_init(bars + 8u * kBarA0Ready, 128);
    ptx::mbar_init(bars + 8u * kBarD3Full, 1);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bars + 8u * (kBarA2Ready + i), 128); ptx::mbar_init(bars + 8u * (kBarA2Empty + i), 1); }
    for (int i = 0; i < kNumDBuf; ++i) { ptx::mbar_init(bars + 8u * (kBarDFull + i), 1); ptx::mbar_init(bars + 8u * (kBarDEmpty + i), 256); }
    for (int i = 0; i < 16; ++i) ptx::mbar_init(bars + 8u * (kBarA1Ready + i), 128);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(s_tmem), kTmemCols);
  for (int i = tid; i < a.n_ops; i += kThreads) s_ops[i] = a.ops[i];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== weight-tape producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int h = 0; h < H; ++h) {
        for (int i = 0; i < a.n_ops; ++i) {
          const TcOp op = s_ops[i];
          ptx::mbar_wait(bars + 8u * (kBarWEmpty + stage), phase ^ 1u);
          const uint32_t bytes = (uint32_t)op.w_bytes16 * 16u;
          ptx::mbar_arrive_expect_tx(bars + 8u * (kBarWFull + stage), bytes);
          ptx::bulk_g2s(base + a.off_wring + stage * kBlkBytes, a.tape + (size_t)op.w_off16 * 16u, bytes,
                        bars + 8u * (kBarWFull + stage));
          if (++stage == (uint32_t)a.nst) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint64_t pbits = 0;   // parity to wait on next, per operand-ready barrier
      int buf = 0;
      for (int h = 0; h < H; ++h) {
        for (int i = 0; i < a.n_ops; ++i) {
          const TcOp op = s_ops[i];
          if (op.wait_bar != 0xFF) {
            ptx::mbar_wait(bars + 8u * op.wait_bar, (uint32_t)((pbits >> op.wait_bar) & 1ull));
            pbits ^= 1ull << op.wait_bar;
          }
          if (!(op.flags & kOpL3)) {
            const uint32_t seq = (uint32_t)(h * 2 * NC + op.chunk_id);
            buf = seq % kNumDBuf;     // ops of different chunks may interleave (schedule 1)
            if (op.flags & kOpFirst) ptx::mbar_wait(bars + 8u * (kBarDEmpty + buf), ((seq / kNumDBuf) & 1u) ^ 1u);
          }
          ptx::mbar_wait(bars + 8u * (kBarWFull + stage), phase);
          ptx::tc_fence_after();
          const uint32_t d = tmem_base + ((op.flags & kOpL3) ? (uint32_t)kD3Col : (uint32_t)(buf * 128));
          const uint64_t adesc = ptx::umma_desc_sw128(base + (uint32_t)op.a_off16 * 16u);
          const uint64_t bdesc = ptx::umma_desc_sw128(base + a.off_wring + stage * kBlkBytes);
          const uint32_t idesc = ptx::umma_idesc_bf16(kRows, op.n8 * 8);
          for (int ks = 0; ks < op.ksteps; ++ks)
            ptx::umma_bf16_ss(d, adesc + 2u * ks, bdesc + 2u * ks, idesc, (ks > 0 || !(op.flags & kOpZero)) ? 1u : 0u);
          ptx::umma_commit(bars + 8u * (kBarWEmpty + stage));
          if (op.flags & kOpLast) ptx::umma_commit(bars + 8u * (kBarDFull + buf));
          if (op.flags & kOpD3Last) ptx::umma_commit(bars + 8u * kBarD3Full);
          if (op.commit_bar != 0xFF) ptx::umma_commit(bars + 8u * op.commit_bar);
          if (++stage == (uint32_t)a.nst) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3, hh = (warp - 4) >> 2, row = q * 32 + lane;
    const bool is_final = hh == 0;
    const int k = tile * kRows + row;
    const bool valid = k < p.K;
    const int kc = valid ? k : p.K - 1;
    const uint32_t a0_row = base + a.off_ring2 + (uint32_t)row * 128u;
    float s[kD], act[kA], act0[kA], cost = 0.f;
    if (is_final) {
#pragma unroll
      for (int d = 0; d < kD; ++d) s[d] = (float)p.states[(size_t)(p.per_row_states ? kc : b) * kD + d];
      get_action(a, b, 0, kc, act);
#pragma unroll
      for (int j = 0; j < kA; ++j) act0[j] = act[j];
      if (p.states_out && valid)
        for (int d = 0; d < kD; ++d) p.states_out[(size_t)k * kD + d] = (double)s[d];
      if (p.actions_out && valid)
        for (int j = 0; j < kA; ++j) p.actions_out[(size_t)k * kA + j] = (double)act[j];
      build_input(a, a0_row, row, s, act);
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(bars + 8u * kBarA0Ready);
    }
    uint32_t ring_uses = 0;
    for (int h = 0; h < H; ++h) {
      const uint32_t seq0 = (uint32_t)(h * 2 * NC);
      for (int nc = 0; nc < NC; ++nc)
        hidden_epilogue(bars, tmem_base, q, hh, row, seq0 + nc, a.b1 + nc * 128 + hh * 64,
                        base + a.off_a1 + (uint32_t)(2 * nc + hh) * kBlkBytes, bars + 8u * (kBarA1Ready + 2 * nc + hh), -1, 0);
      for (int nc = 0; nc < NC; ++nc) {
        hidden_epilogue(bars, tmem_base, q, hh, row, seq0 + NC + nc, a.b2 + nc * 128 + hh * 64,
                        base + a.off_ring2 + (uint32_t)hh * kBlkBytes, bars + 8u * (kBarA2Ready + hh), kBarA2Empty + hh,
                        (ring_uses & 1u) ^ 1u);
        ++ring_uses;
      }
      if (is_final) {
        if (h + 1 < H) get_action(a, b, h + 1, kc, act);     // drawn while the last MMAs drain
        ptx::mbar_wait(bars + 8u * kBarD3Full, (uint32_t)(h & 1));
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kD3Col, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        float delta[kD];
#pragma unroll
        for (int d = 0; d < kD; ++d) delta[d] = fmaf(__uint_as_float(v[d]) + a.b3[d], a.std_delta[d], a.mean_delta[d]);   // dynamics.py:113
        if (s[5] >= 0.2f) cost += 10.f;                                      // cost_functions.py:16-26 (pre-step state)
        if (s[6] >= 0.f) cost += 10.f;
        if (s[7] >= 0.f) cost += 10.f;
        cost -= delta[17] / 0.01f;                                           // cost_functions.py:28
#pragma unroll
        for (int d = 0; d < kD; ++d) s[d] += delta[d];                       // dynamics.py:116
        if (p.states_out && valid)
          for (int d = 0; d < kD; ++d) p.states_out[((size_t)(h + 1) * p.K + k) * kD + d] = (double)s[d];
        if (h + 1 < H) {
          if (p.actions_out && valid)
            for (int j = 0; j < kA; ++j) p.actions_out[((size_t)(h + 1) * p.K + k) * kA + j] = (double)act[j];
          build_input(a, a0_row, row, s, act);
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(bars + 8u * kBarA0Ready);
        }
      }
    }
    if (is_final) {
      // ---- per-tile argmin over the 128 rows (np.argmin order) ----
      if (p.costs_out && valid) p.costs_out[(size_t)b * p.K + k] = (double)cost;
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
      double* r_cost = reinterpret_cast<double*>(s_red);
      int64_t* r_idx = reinterpret_cast<int64_t*>(s_red + 32);
      int* r_src = reinterpret_cast<int*>(s_red + 64);
      if (lane == 0) { r_cost[q] = c; r_idx[q] = idx; r_src[q] = src; }
      ptx::named_bar_sync(1, 128);
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
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace bcmpc
