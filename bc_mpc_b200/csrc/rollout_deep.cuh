// BF16 fast-mode rollout kernel for WIDE / DEEP dynamics nets: L >= 2 hidden layers of up to 1024 units (BASELINE.json
// configs[4]: 26 -> 1024 -> 1024 -> 1024 -> 20, K = 2^20 paths x H = 30), ReLU or tanh, with or without LayerNorm; uniform
// shooting at every width, policy-guided shooting (template flag POL) at 129..1024 units.
//
// Why a third kernel: the hidden activations of a 128-path tile at width 1024 are 256 KB in bf16 - all of TMEM - and a
// layer needs its whole input alive until its last output chunk, i.e. input + output = 512 KB per 128 paths, more than
// TMEM (256 KB) + shared memory (227 KB) of one SM.  So the HIDDEN DIMENSION IS SPLIT OVER A CLUSTER OF FOUR CTAs:
//
//   cluster rank r = 2 p + q     p = K-half (hidden units [512 p, 512 p + 512)), q = row half (128 of the unit's 256 paths)
//   CTAs (2p, 2p+1) form a cta_group::2 pair: M = 256 MMAs over the 256 paths, every weight tile fetched once per pair
//   CTA r keeps ITS K-half of every hidden activation of ITS 128 paths on chip and computes, for every output chunk of a
//   layer, the PARTIAL sum over its K-half.  Output chunks are owned by the pair whose K-half they belong to: the other
//   pair's fp32 partial travels through distributed shared memory (CTA r <-> r ^ 2: same paths, other K-half) and the
//   owner adds it to its own accumulator before bias / ReLU / bf16.  Layer 1 needs no exchange (K = 28 inputs, every
//   pair computes its own half of the units); the output layer exchanges its 20-column partial both ways, so both pairs
//   carry identical fp32 states (a + b == b + a) and build their own layer-1 input.
//
// Per CTA (NCH = 64-unit chunks per K-half; 8 at width 1024):
//   TMEM   [0, 32 NCH)            activation "X" (hidden 1, 3, ...): bf16 pairs, A operand from TMEM
//          [32 NCH, 32 NCH+32 NT) the first NT k-blocks of activation "Y" (hidden 2, 4, ...)
//          2 x 64 columns          fp32 accumulators (ping-pong); the output layer uses 32 columns of one
//   smem   A0 (input tile, 16 KB), the last NCH - NT k-blocks of Y as K-major SWIZZLE_128B tiles (A operand from smem),
//          2 x 32 KB receive slots for the partner's partials, a ring of 16 KB weight stages, the hidden biases
//   registers (epilogue group 0): state, cost, first action of the path that owns the TMEM lane
// Activations alternate between X and Y, so a layer's output never overwrites its input.  No activation, state or
// per-step cost touches HBM; weights stream from L2 (1 MB per pair and hidden layer at width 1024 - the same bytes per
// FLOP as rollout_tmem_kernel).
//
// Schedule of one step (every CTA; "use" = one accumulator buffer's life, buffers alternate):
//   layer 1      NCH uses: D = A0 x W1[:, own chunk c]                       -> ReLU -> X block c
//   layer l>=2   2 NCH uses in the order F0 O0 F1 O1 ...: Fi = partial of the partner's chunk i (sent), Oi = partial of own
//                chunk i (+ the partner's Fi + bias -> ReLU -> block i of the layer's output)
//   output       1 use: partial of the 20 deltas (exchanged both ways)     -> denormalise, cost, state update, next input
// Roles (384 threads): warp 0 weight producer; warp 1 MMA issuer (leader CTA of the pair) / "weights landed" forwarder
// (peer CTA); warp 2 TMEM allocation; warps 4-7 epilogue group 0 (layer-1 even chunks, F sends, output + state), warps
// 8-11 group 1 (layer-1 odd chunks, O finalisation).  Accumulator buffer g belongs to group g.
//
// Replaces dynamics.py:54-71,106-119 (any n_layers / size up to 1024) + cost_functions.py:10-30,59-63 +
// controllers.py:53,63-85 for HalfCheetah shapes (D = 20, A = 6).
#pragma once
#include "rollout_tmem.cuh"

namespace bcmpc {
namespace dpx {

using tmx::kA;
using tmx::kD;
using tmx::kRows;
using tmx::kThreads;

constexpr int kCluster = 4;
constexpr int kMaxNch = 8;                      // hidden <= 1024
constexpr int kMaxHidden = BCMPC_MAX_LAYERS - 1;  // hidden layers (the handle accepts n_hidden + 1 <= BCMPC_MAX_LAYERS)
constexpr uint32_t kStageBytes = 16384;         // one weight stage: four (64-unit chunk x 64-wide k-block) tiles of this CTA's row half
constexpr uint32_t kTileH = 4096, kTileO = 2048;   // hidden tile (32 rows x 128 B), output tile (16 rows x 128 B)
// one receive slot: HALF a chunk per message, 128 rows x 32 values (two whole-chunk slots cost the weight ring two stages); with
// 24-bit partials 12 KB - the two slots then leave room for a SEVENTH weight stage (the output layer's 10 KB message fits too)
// The hidden layers' partials cross the cluster as 24-bit floats (sign, exponent, 15 mantissa bits, round to nearest): four
// values in three words, 12 instead of 16 st.async per thread and use.  The exchange needs 16 B/clk per SM and direction to
// keep the tensor pipe busy and the SM-to-SM network delivers 17.5 (tools/micro/dsmem_bw.cu): with 25 % fewer bytes the steady
// state is the pipe's.  The rounding (2^-16 of the PARTNER's partial; the own partial stays fp32) is 2^8 times finer than the bf16
// rounding of the activation that follows.  The output layer's 20-column exchange stays fp32 (both K-halves must carry
// bit-identical states).  -DBCMPC_DEEP_FP32_EXCHANGE restores the fp32 partials (A/B, tools/deep_bench.py).
#ifdef BCMPC_DEEP_FP32_EXCHANGE
constexpr bool kPack24 = false;
#else
constexpr bool kPack24 = true;
#endif
constexpr uint32_t kMsgV4 = kPack24 ? 6 : 8;            // 16-byte stores per thread and half-chunk message (32 values)
constexpr uint32_t kMsgBytes = kMsgV4 * 2048u;         // bytes of one hidden-layer message (128 rows)
constexpr uint32_t kRecvBytes = kMsgBytes;
constexpr uint32_t kA0Bytes = 128 * 128;
constexpr uint32_t kBlockBytes = 128 * 128;     // one 64-unit k-block of an activation as an smem A tile

__host__ __device__ constexpr int stages_per_use(int nch) { return (nch + 3) / 4; }
constexpr int kNumAcc = 2;                      // accumulator buffers (64 columns each): buffer g belongs to epilogue group g
__host__ __device__ constexpr int tmem_y_blocks(int nch) { return nch < 12 - nch ? nch : 12 - nch; }   // NT: Y k-blocks that fit in TMEM next to X and 2 accumulators
// weight stages of one step: layer 1, 2 nch uses per hidden layer >= 2, the output layer - stages_per_use each
__host__ __device__ constexpr int stages_per_step(int nch, int L) { return stages_per_use(nch) * ((L - 1) * 2 * nch + 2); }

// misc smem: barriers then scratch
enum : uint32_t {
  kBarWFull = 0,        // [8] weight stage landed                  (producer / peer forwarder -> issuer)
  kBarWEmpty = 64,      // [8] weight stage consumed                (issuer commit -> producer)
  kBarA0Ready = 128,    // input rows of the step in smem           (group 0 of both CTAs -> issuer)
  kBarActReady = 136,   // [8] k-block kb of the current layer's output stored (both CTAs -> issuer)
  kBarAccFull = 200,    // [2] accumulator complete                 (issuer commit -> epilogue of both CTAs)
  kBarAccEmpty = 224,   // [2] accumulator drained                  (epilogue of both CTAs -> issuer)
  kBarRecvFull = 248,   // [2] the partner's partial is in my receive slot      (partner's st.async bytes + my arming arrive)
  kBarRecvEmpty = 264,  // [2] the partner has read ITS receive slot (I may overwrite it)   (partner's receiver warps -> my sender)
  //      280..295        (free)
  // LayerNorm variant: per-row statistics of a hidden layer are spread over the two K-halves (and, for hidden 1, over both epilogue groups)
  kBarStat = 296,       // [4] the PARTNER's partial sums of hidden layer l (index l - 1) are in my shared memory  (partner's st.async bytes + my arming arrive)
  kBarStatLoc = 328,    // [2] [0] group 0's hidden-1 partial is in shared memory (-> group 1);  [1] group 1's last-hidden partial (-> group 0)
  // proposal network (policy-guided shooting), each once per step
  kBarPolIn = 344,      // the policy input rows are in shared memory          (group 0 of both CTAs -> issuer)
  kBarPdFull = 352,     // [4] policy accumulators: layer-1 chunks 0, 1, layer-2 chunks 0, 1   (issuer commit -> epilogue of both CTAs)
  kBarAp1 = 384,        // [2] policy hidden-1 chunk j stored                  (group j of both CTAs -> issuer)
  kBarAp2 = 400,        // [2] policy hidden-2 chunk j stored
  kBarPolFull = 416,    // the action means are complete                       (issuer commit -> group 0 of both CTAs)
  kOffTmemPtr = 424,
  kOffRed = 432,        // argmin scratch (80 B)
  // [2][2] issue tokens of issuer warp w: the other issuer has issued its use (issuer -> issuer).  Token k of a warp goes to barrier
  // [w][k & 1]: a warp can be TWO tokens ahead of the other's waits (the first use of a layer is issued without a token and arrives
  // before the wait that consumes the previous one has run), and two arrivals on one single-count barrier would skip a phase - the
  // waiter then sits on a parity that never completes (seen as a hang under policy-guided shooting at 2 chunks per K-half).
  kBarSig = 512,
  kMiscBytes = 576
};
static_assert(kBarPolFull + 8u <= kOffTmemPtr && kOffTmemPtr + 8u <= kOffRed && kOffRed + 80u <= kBarSig && kBarSig + 4u * 8u <= kMiscBytes,
              "misc shared memory: barriers and scratch overlap");
constexpr uint32_t kPolImgBytes = 6u * 4096u + 2u * 2048u;      // resident proposal network per CTA: [Wp1: 2 chunks][Wp2: 2 x 2][Wp3: 2 k-blocks] (tm_pack_policy, CTA-pair geometry)
constexpr uint32_t kPolBytes = kPolImgBytes + tmx::kPolVecFloats * 4u;
// LayerNorm: (sum, sum of squares) per path.  Remote slots (the partner CTA's partials): [0], [1] hidden 1 of the partner's group 0 / 1,
// [2 + (l & 1)] hidden layer l >= 2; then the local slots [group] (this CTA's partials for its other group).  Hidden 1 has slots
// of its own because a group without a layer-1 chunk (NCH = 1: group 1) sends its (zero) partial of the NEXT step as soon as its
// last O use is done - before the other group has read the last layer's statistics of THIS step.
constexpr uint32_t kStatBytes = (4u + 2u) * 128u * 8u;
#ifndef BCMPC_DEEP_STAGES
#define BCMPC_DEEP_STAGES 7
#endif
constexpr int kMaxStages = BCMPC_DEEP_STAGES;      // (a seventh stage fits since the 24-bit exchange shrank the receive slots: A/B in tools/ab_deep.sh)
#ifndef BCMPC_DEEP_WEIGHTS_FIRST
#define BCMPC_DEEP_WEIGHTS_FIRST 1
#endif
constexpr bool kWeightsFirst = BCMPC_DEEP_WEIGHTS_FIRST != 0;   // issuer: wait for the weight stages before the accumulator's "drained" barrier

struct Layout {
  int nch = 0, L = 0, nt = 0, ns = 0, spu = 0, nst = 0;
  int nst_r[2] = {0, 0}, first_r[2] = {0, 0};     // the ring is split in two: one sub-ring per MMA issuer
  uint32_t off_a0 = 0, off_y = 0, off_recv = 0, off_pol = 0, off_vec = 0, vec_bytes = 0, off_stat = 0, off_ring = 0, off_misc = 0, smem_bytes = 0;
  uint32_t col_x = 0, col_y = 0, col_acc = 0;
};
constexpr Layout make_layout(int nch, int max_hidden, bool ln = false, bool pol = false) {
  Layout Y;
  Y.nch = nch; Y.L = max_hidden;
  Y.nt = tmem_y_blocks(nch); Y.ns = nch - Y.nt; Y.spu = stages_per_use(nch);
  Y.col_x = 0; Y.col_y = 32u * nch; Y.col_acc = 32u * nch + 32u * Y.nt;
  Y.off_a0 = 0;
  Y.off_y = kA0Bytes;
  Y.off_recv = Y.off_y + Y.ns * kBlockBytes;
  Y.off_pol = Y.off_recv + 2 * kRecvBytes;                      // (1024-byte aligned: UMMA tiles)
  Y.off_ring = Y.off_pol + (pol ? kPolBytes : 0u);
  // [hidden biases of layers 2..L: own K-half, 64 nch floats each][output bias 32] after the ring (sized for max_hidden layers)
  // (LayerNorm variant: correction vectors u, c instead of the bias - twice the floats - and the statistics slots)
  Y.vec_bytes = ln ? ((uint32_t)(max_hidden - 1) * 2u * 64u * nch + 64u) * 4u + kStatBytes : ((uint32_t)(max_hidden - 1) * 64u * nch + 32u) * 4u;
  const uint32_t fixed = Y.off_ring + ((Y.vec_bytes + 1023u) / 1024u * 1024u) + kMiscBytes;
  Y.nst = (int)((tmx::kSmemLimit - fixed) / kStageBytes);
  if (Y.nst > kMaxStages) Y.nst = kMaxStages;
#ifdef BCMPC_DEEP_ODD_TO_F
  Y.nst_r[0] = (Y.nst + 1) / 2; Y.nst_r[1] = Y.nst / 2;
#else
  Y.nst_r[0] = Y.nst / 2; Y.nst_r[1] = (Y.nst + 1) / 2;      // one sub-ring per issuer warp (an odd stage goes to the O issuer's)
#endif
  Y.first_r[0] = 0; Y.first_r[1] = Y.nst_r[0];
  Y.off_vec = Y.off_ring + Y.nst * kStageBytes;
  Y.off_stat = Y.off_vec + (ln ? ((uint32_t)(max_hidden - 1) * 2u * 64u * nch + 64u) * 4u : 0u);      // (8-byte aligned: a multiple of 256 bytes)
  Y.off_misc = Y.off_vec + (Y.vec_bytes + 1023u) / 1024u * 1024u;
  Y.smem_bytes = Y.off_misc + kMiscBytes;
  return Y;
}
template <int NCH, bool LN = false, bool POL = false>
struct Lay {
  static constexpr Layout value = make_layout(NCH, kMaxHidden, LN, POL);
};

struct DeepArgs {
  tmx::TmArgs t;          // plan arguments, normalisation, hint; t.stream = weight image [4 ranks][stages_per_step][16 KB]; t.lnvec = bias vector
  int n_hidden;           // L
  int s_half;             // 64 * NCH: units per K-half (padded)
};

// pipeline trace of CTA 0 (developer builds, -DBCMPC_TRACE; tools/deep_trace.py): [step (4)][role: issuer warp 1, group 0, group 1, issuer warp 3][use (64)][stamp (8)]
constexpr int kTrSteps = 4, kTrRoles = 4, kTrUses = 64, kTrStamps = 8, kTraceWords = kTrSteps * kTrRoles * kTrUses * kTrStamps;
#ifdef BCMPC_TRACE
#define BCMPC_DTRACE(tr, tstep, role, use, k)                                                                               \
  do {                                                                                                                      \
    if ((tr) != nullptr && blockIdx.x == 0 && (tstep) < (uint32_t)kTrSteps)                                                 \
      (tr)[(((tstep) * kTrRoles + (role)) * kTrUses + (use)) * kTrStamps + (k)] = (unsigned long long)clock64();           \
  } while (0)
#else
#define BCMPC_DTRACE(tr, tstep, role, use, k) do { } while (0)
#endif

inline bool deep_supported(const bcmpc_config& c, std::string* why) {
  if (c.obs_dim != kD || c.act_dim != kA) { *why = "bf16 fast path is specialised for obs_dim=20, act_dim=6 (HalfCheetah)"; return false; }
  if (c.n_hidden < 2 || c.n_hidden > kMaxHidden) { *why = "bf16 fast path (cluster-split kernel) supports 2..4 hidden layers"; return false; }
  if (c.hidden_size < 1 || c.hidden_size > 64 * 2 * kMaxNch) { *why = "bf16 fast path supports hidden_size <= 1024"; return false; }
  if (c.proposal != BCMPC_PROP_UNIFORM || c.first_proposal != 0) {
    if ((c.hidden_size + 127) / 128 < 2) {      // (the proposal network's activations take 128 TMEM columns of the dynamics net's)
      *why = "bf16 fast path: the cluster-split kernel (n_hidden > 2) runs policy-guided proposals for nets of 129..1024 hidden units; narrower ones run them in BCMPC_PREC_FP32_VERIFY";
      return false;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {      // arrive at `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// 16 bytes into another CTA's shared memory; the store itself reports its bytes to the destination's mbarrier
// (st.async, SASS STAS): the sender needs no fence and no separate "data has landed" arrive - a release.cluster arrive per
// thread behind 16 remote stores cost the sending group half of its time (MEMBAR + ERRBAR waiting for the store acks).
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t bar, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(x),
               "r"(y), "r"(z), "r"(w), "r"(bar)
               : "memory");
}
// four fp32 bit patterns -> three words of 24-bit floats (round to nearest; Inf / NaN survive), and back (low byte zero)
__device__ __forceinline__ void pack24(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
  a += 0x80u; b += 0x80u; c += 0x80u; d += 0x80u;
  w0 = __byte_perm(a, b, 0x3217);          // [a3 a2 a1 | b3]
  w1 = __byte_perm(b, c, 0x2176);          // [b2 b1 | c3 c2]
  w2 = __byte_perm(c, d, 0x1765);          // [c1 | d3 d2 d1]
}
__device__ __forceinline__ void unpack24(uint32_t w0, uint32_t w1, uint32_t w2, float& a, float& b, float& c, float& d) {
  a = __uint_as_float(w0 & 0xFFFFFF00u);
  b = __uint_as_float(__byte_perm(w0, w1, 0x0764) & 0xFFFFFF00u);      // [b3 = w0.0, b2 = w1.3, b1 = w1.2, -]
  c = __uint_as_float(__byte_perm(w1, w2, 0x1074) & 0xFFFFFF00u);      // [c3 = w1.1, c2 = w1.0, c1 = w2.3, -]
  d = __uint_as_float(w2 << 8);
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// TMEM -> registers: this warp's 32 lanes x 32 columns - reuse ptx::tmem_ld32

// TANH: hidden activation tanh (the reference builder's default, dynamics.py:60; tanh.approx.bf16x2 on the rounded pair, like
// rollout_tmem_kernel's kVarTanh) instead of ReLU - the biases of layers >= 2 are explicit here anyway, layer 1's ride in the GEMM.
// LN: LayerNorm after every hidden activation (dynamics.py:68-69, FLAGS.LAYER_NORM - the reference's default): like
// rollout_tmem_kernel the tensor cores multiply the UN-normalised ReLU outputs r by W' = diag(gamma) W and the owner's epilogue
// applies  LN(r) W + b = rho (r W') - rho mu u + c  (u = column sums of the bf16 W', c = beta^T W + b; rho, mu per path).  A path's
// hidden units are spread over the two K-halves, so every CTA keeps the partial (sum, sum of squares) of ITS units and the two
// partner CTAs exchange them (8 bytes per path and layer, st.async) while the next layer's MMAs run.
// POL: policy-guided shooting (controllers.py:189-237 / the legacy BC planners) on nets of 2..8 chunks per K-half (hidden 129..1024):
// the proposal phase of rollout_tmem_kernel at the top of every step -
// observation filter -> two tanh layers (<= 128 units) -> action means -> noise mixing - with the proposal network resident in
// every CTA (both K-half pairs compute the same actions from their bit-identical states), its accumulators in the two regular
// buffers and its hidden activations in the (then dead) activation columns [0, 128).
template <int NCH, bool DBG, bool TANH, bool LN, bool POL = false>
__global__ void __launch_bounds__(kThreads, 1) rollout_deep_kernel(const __grid_constant__ DeepArgs da) {
  static_assert(!POL || (NCH >= 2 && 32 * NCH + 32 * tmem_y_blocks(NCH) >= 128), "the proposal network needs 128 activation columns in TMEM");
  auto act2 = [](float x, float y) -> uint32_t {
    if constexpr (TANH) return ptx::tanh_bf16x2(tmx::pack_bf16(x, y));
    else return tmx::pack_relu_bf16(x, y);
  };
  constexpr Layout Y = Lay<NCH, LN, POL>::value;
  constexpr int NT = Y.nt, SPU = Y.spu, NST = Y.nst;
  constexpr uint32_t kIdescH = ptx::umma_idesc_bf16(2 * kRows, 64), kIdescO = ptx::umma_idesc_bf16(2 * kRows, 32);
  static_assert(Y.nst_r[0] >= SPU && Y.nst_r[1] >= SPU, "weight ring too small");
  static_assert(NCH >= 1 && NCH <= kMaxNch && NST >= 2 && NST <= kMaxStages && Y.col_acc + 64u * kNumAcc <= 512u, "unsupported width");
  const tmx::TmArgs& a = da.t;
  const PlanArgs& p = a.p;
  const int L = da.n_hidden;

  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t khalf = rank >> 1, qrow = rank & 1u;
  const bool leader = qrow == 0;
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t misc = base + Y.off_misc;
  const uint32_t misc_leader = ptx::mapa(misc, rank & ~1u);        // barriers the pair's issuer waits on
  const uint32_t base_partner = ptx::mapa(base, rank ^ 2u);        // same paths, other K-half
  const uint32_t misc_partner = base_partner + Y.off_misc;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Y.off_misc + kOffTmemPtr);
  // persistent clusters: cluster `first_unit` rolls the units first, first + stride, ...; a unit = 256 paths of one root state
  const int units_per_state = p.tiles_per_state / 2;
  const int total_units = p.B * units_per_state;
  const int first_unit = (int)blockIdx.x / kCluster, unit_stride = (int)gridDim.x / kCluster;
  const int n_it = first_unit < total_units ? (total_units - first_unit + unit_stride - 1) / unit_stride : 0;
  const int H = p.H, Htot = n_it * H;
  const uint32_t hc = a.hint_chain;
  const int stages_step = stages_per_step(NCH, L);
  const uint32_t xs_step = (uint32_t)(L - 1) * NCH * 2u + 1u;      // exchanged messages per step: two halves per hidden chunk + the output partial

  if (tid == 0) {
    if (base & 1023u) { printf("bcmpc: dynamic smem base not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < 8; ++i) {
      ptx::mbar_init(misc + kBarWFull + 8u * i, leader ? 2 : 1);   // leader: own producer + the peer's "my half has landed"
      ptx::mbar_init(misc + kBarWEmpty + 8u * i, 1);
    }
    ptx::mbar_init(misc + kBarA0Ready, 8);
    for (int i = 0; i < 8; ++i) ptx::mbar_init(misc + kBarActReady + 8u * i, 8);
    for (int i = 0; i < kNumAcc; ++i) {
      ptx::mbar_init(misc + kBarAccFull + 8u * i, 1);
      ptx::mbar_init(misc + kBarAccEmpty + 8u * i, 8);
    }
    for (int i = 0; i < 4; ++i) ptx::mbar_init(misc + kBarSig + 8u * i, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(misc + kBarRecvFull + 8u * i, 1);        // the receiver's arming arrive (+ the bytes of the partner's st.async)
      ptx::mbar_init(misc + kBarRecvEmpty + 8u * i, 4);       // one arrive per receiving warp
    }
    if constexpr (LN) {
      for (int i = 0; i < 4; ++i) ptx::mbar_init(misc + kBarStat + 8u * i, 1);
      for (int i = 0; i < 2; ++i) ptx::mbar_init(misc + kBarStatLoc + 8u * i, 4);
    }
    if constexpr (POL) {
      ptx::mbar_init(misc + kBarPolIn, 8);
      for (int i = 0; i < 4; ++i) ptx::mbar_init(misc + kBarPdFull + 8u * i, 1);
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(misc + kBarAp1 + 8u * i, 8); ptx::mbar_init(misc + kBarAp2 + 8u * i, 8); }
      ptx::mbar_init(misc + kBarPolFull, 1);
    }
    ptx::fence_mbar_init();
    if constexpr (LN) {
      // armed for the first step (nothing can arrive before the cluster barrier below): hidden 1 gets both groups' partials of the
      // partner (2 x 1 KB), every later layer one group's; each consumer re-arms its barrier right after it has waited on it
      if (n_it > 0) {
        ptx::mbar_arrive_expect_tx(misc + kBarStat, 2048u);
        for (int l = 2; l <= L; ++l) ptx::mbar_arrive_expect_tx(misc + kBarStat + 8u * (uint32_t)(l - 1), 1024u);
      }
    }
  }
  if constexpr (POL) {  // resident proposal network: this CTA's row half of every tile, then the three bias vectors
    const uint4* src = reinterpret_cast<const uint4*>(a.pol + (size_t)qrow * kPolImgBytes);
    uint4* dst = reinterpret_cast<uint4*>(smem + Y.off_pol);
    for (int i = tid; i < (int)(kPolImgBytes / 16u); i += kThreads) dst[i] = src[i];
    float* s_pv = reinterpret_cast<float*>(smem + Y.off_pol + kPolImgBytes);
    for (int i = tid; i < 2 * tmx::kPolMaxHidden + kMaxAct; i += kThreads) s_pv[i] = a.polvec[i];
    ptx::fence_proxy_async_smem();      // the tensor cores read the tiles through the async proxy (after the cluster barrier below)
  }
  if constexpr (LN) {   // per hidden layer >= 2: [u (own K-half)][c (own K-half)], then [u3 (32)][c3 (32)] of the output layer
    float* s_vec = reinterpret_cast<float*>(smem + Y.off_vec);
    const int sh = 64 * NCH;
    for (int i = tid; i < (L - 1) * 2 * sh; i += kThreads) {
      const int layer = i / (2 * sh), part = (i % (2 * sh)) / sh, jj = i % sh;
      s_vec[i] = a.lnvec[(size_t)layer * (4 * sh) + (size_t)part * (2 * sh) + khalf * sh + jj];
    }
    for (int i = tid; i < 64; i += kThreads) s_vec[(L - 1) * 2 * sh + i] = a.lnvec[(size_t)(L - 1) * 4 * sh + i];
  } else {   // biases: own K-half of every hidden layer >= 2, then the output bias
    float* s_vec = reinterpret_cast<float*>(smem + Y.off_vec);
    const int sh = 64 * NCH;
    for (int i = tid; i < (L - 1) * sh; i += kThreads) s_vec[i] = a.lnvec[(size_t)(i / sh) * (2 * sh) + khalf * sh + (i % sh)];
    for (int i = tid; i < 32; i += kThreads) s_vec[(L - 1) * sh + i] = a.lnvec[(size_t)(L - 1) * 2 * sh + i];
  }
  if (warp == 2) ptx::tmem_alloc_cg<2>(ptx::smem_u32(s_tmem), 512);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  if (*s_tmem != 0u) { if (tid == 0) printf("bcmpc: unexpected TMEM base 0x%x\n", *s_tmem); __trap(); }

  // Weight stages of one step, in consumption order: [layer 1: SPU stages of up to four chunk tiles] [layers 2..L: F0, O0, F1,
  // O1, ... SPU stages of four k-blocks each] [output: SPU stages].  TWO ISSUER WARPS: warp 1 issues layer 1, every F use and
  // the output layer (weight sub-ring 0), warp 3 every O use (sub-ring 1).  A UTCHMMA issue is paced by the tensor pipe (32
  // cycles per M=256,N=64 MMA), but every mbarrier wait / tcgen05.commit around it costs the issuing warp 100-250 cycles with
  // the pipe idle (measured with one issuer: 1560-1780 cycles per 32-MMA use against 1024 on the pipe).  With two warps
  // the waits of one (accumulator drained, weight stages landed, layer input stored) run while the other issues; a token
  // passed back and forth (mbarrier hand-off, given away just before the commits) keeps them taking turns instead of
  // interleaving, so consecutive uses finish one after the other and the three accumulator buffers cover the drains.
  // Every accumulator is written by one warp only: results do not depend on how the two streams interleave.
  if (warp == 0) {
    // ===================== weight producer: every CTA streams its own row halves =====================
    // (ONE producer walking the stages in consumption order.  A producer warp per sub-ring - no head-of-line blocking on the
    // other sub-ring's "stage empty" barrier - was measured and is slower, 51 550 vs 49 920 cycles per step: the stream is
    // bound by L2 -> SM bandwidth, weights 5.1 TB/s + exchange ~2 TB/s over 132 SMs, not by when the requests are issued.)
    uint32_t wsr[2] = {0, 0};
    auto fill = [&](int r, const uint8_t* src, uint32_t bytes) {
      const uint32_t n = wsr[r]++, stage = (uint32_t)Y.first_r[r] + n % (uint32_t)Y.nst_r[r], ph = (n / (uint32_t)Y.nst_r[r]) & 1u;
      ptx::mbar_wait(misc + kBarWEmpty + 8u * stage, ph ^ 1u);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(misc + kBarWFull + 8u * stage, bytes);
        ptx::bulk_g2s(base + Y.off_ring + stage * kStageBytes, src, bytes, misc + kBarWFull + 8u * stage);
      }
      __syncwarp();
    };
    for (int h = 0; h < Htot; ++h) {
      const uint8_t* src = a.stream + (size_t)rank * stages_step * kStageBytes;
      for (int sg = 0; sg < SPU; ++sg, src += kStageBytes) fill(0, src, (uint32_t)(NCH - 4 * sg < 4 ? NCH - 4 * sg : 4) * kTileH);
      for (int l = 2; l <= L; ++l)
        for (int i = 0; i < 2 * NCH; ++i)
          for (int sg = 0; sg < SPU; ++sg, src += kStageBytes) fill(i & 1, src, (uint32_t)(NCH - 4 * sg < 4 ? NCH - 4 * sg : 4) * kTileH);
      for (int sg = 0; sg < SPU; ++sg, src += kStageBytes) fill(0, src, (uint32_t)(NCH - 4 * sg < 4 ? NCH - 4 * sg : 4) * kTileO);
    }
  } else if (warp == 1 || warp == 3) {
    const uint32_t w = warp == 1 ? 0u : 1u;            // issuer 0: layer 1, F uses, output; issuer 1: O uses
    const uint32_t nst_g = (uint32_t)Y.nst_r[w], first_g = (uint32_t)Y.first_r[w];
    if (leader) {
      // ===================== MMA issuers of the pair =====================
      // Accumulator buffer g belongs to epilogue group g (layer-1 chunk c -> c & 1, F -> 0, O -> 1, output -> 0), so every phase
      // of a buffer's "full" barrier is seen by the same waiters in order (a waiter that skips phases can mistake "not yet
      // started" for "already complete": the parity bit has only two states).  Two buffers are enough because the two
      // issuers take turns: use n + 2 is issued ~2048 cycles after use n, whose life - issue, 1024 cycles on the pipe, the
      // drain with its two mbarrier hops (~650) - is over by then; the TMEM columns a third buffer would take keep two more
      // k-blocks of the Y activation out of shared memory, which buys the weight ring two more stages.
      const uint16_t pair_mask = (uint16_t)(3u << (2u * khalf));
      const uint64_t a0desc = ptx::umma_desc_sw128(base + Y.off_a0);
      uint32_t ws = 0, lo = 0, ntok = 0, nsent = 0;    // weight stages of this warp's sub-ring, layer outputs consumed, tokens received / given so far
      // uses of buffer 0 / 1 per step, and so far at the start of the step
      constexpr uint32_t kL1b[2] = {(NCH + 1) / 2, NCH / 2};
      const uint32_t cb[2] = {kL1b[0] + (uint32_t)(L - 1) * NCH + 1u, kL1b[1] + (uint32_t)(L - 1) * NCH};
      uint32_t nb[2] = {0, 0};
      auto tok_wait = [&]() {      // the next token for this warp
        ptx::mbar_wait(misc + kBarSig + 8u * (2u * w + (ntok & 1u)), (ntok >> 1) & 1u, hc);
        ++ntok;
      };
      auto tok_other = [&]() -> uint32_t { return misc + kBarSig + 8u * (2u * (w ^ 1u) + (nsent & 1u)); };      // where this warp's next token goes
      for (int h = 0; h < Htot; ++h, nb[0] += cb[0], nb[1] += cb[1]) {
        if constexpr (POL) {
          if (w == 0) {
            // ---- proposal network (issuer 0): policy input (smem) -> tanh layer -> tanh layer -> action means, as in rollout_tmem_kernel.
            // Accumulators: chunk j of both hidden layers in buffer j, the means in the first 32 columns of buffer 0; every one of them
            // has been drained before the barrier that lets its buffer's next user start completes.
            const uint64_t wpdesc = ptx::umma_desc_sw128(base + Y.off_pol);
            const uint32_t np1 = (uint32_t)a.np1, np2 = (uint32_t)a.np2, par = (uint32_t)h & 1u;
            ptx::mbar_wait(misc + kBarPolIn, par, hc);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              for (uint32_t j = 0; j < np1; ++j) {
#pragma unroll
                for (uint32_t ks = 0; ks < 2; ++ks)
                  ptx::umma_ss<2>(Y.col_acc + 64u * j, a0desc + 2u * ks, wpdesc + (uint64_t)(j * (kTileH >> 4) + 2u * ks), kIdescH, ks);
                umma_commit_pair(misc + kBarPdFull + 8u * j, pair_mask);
              }
            }
            __syncwarp();
            for (uint32_t j = 0; j < np1; ++j) ptx::mbar_wait(misc + kBarAp1 + 8u * j, par, hc);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              for (uint32_t j = 0; j < np2; ++j) {
                for (uint32_t kb = 0; kb < np1; ++kb)
#pragma unroll
                  for (uint32_t ks = 0; ks < 4; ++ks)
                    ptx::umma_ts<2>(Y.col_acc + 64u * j, kb * 32u + ks * 8u, wpdesc + (uint64_t)((2u + j * 2u + kb) * (kTileH >> 4) + 2u * ks), kIdescH,
                                    (kb | ks) ? 1u : 0u);
                umma_commit_pair(misc + kBarPdFull + 8u * (2u + j), pair_mask);
              }
            }
            __syncwarp();
            for (uint32_t j = 0; j < np2; ++j) ptx::mbar_wait(misc + kBarAp2 + 8u * j, par, hc);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              for (uint32_t kb = 0; kb < np2; ++kb)
#pragma unroll
                for (uint32_t ks = 0; ks < 4; ++ks)
                  ptx::umma_ts<2>(Y.col_acc, 64u + kb * 32u + ks * 8u, wpdesc + (uint64_t)(6u * (kTileH >> 4) + kb * (kTileO >> 4) + 2u * ks), kIdescO,
                                  (kb | ks) ? 1u : 0u);
              umma_commit_pair(misc + kBarPolFull, pair_mask);
            }
            __syncwarp();
          }
        }
        // ---- layer 1 (issuer 0): own chunks, K = 32 (2 k-steps), A from smem ----
        if (w == 0) {
          ptx::mbar_wait(misc + kBarA0Ready, (uint32_t)h & 1u, hc);
          for (int j = 0; j < SPU; ++j) {
            const uint32_t stage = first_g + ws % nst_g;
            ptx::mbar_wait(misc + kBarWFull + 8u * stage, (ws / nst_g) & 1u);
            const uint64_t bdesc = ptx::umma_desc_sw128(base + Y.off_ring + stage * kStageBytes);
            for (int t = 0; t < 4 && 4 * j + t < NCH; ++t) {
              const uint32_t c = (uint32_t)(4 * j + t), b = c & 1u, n = nb[b] + (c >> 1);      // the n-th use of buffer b
              if (lane == 0) BCMPC_DTRACE(a.trace, (uint32_t)h, 0, 4 * j + t, 0);
              ptx::mbar_wait(misc + kBarAccEmpty + 8u * b, (n & 1u) ^ 1u, hc);      // the buffer's previous use has been drained
              ptx::tc_fence_after();
              if (lane == 0) BCMPC_DTRACE(a.trace, (uint32_t)h, 0, 4 * j + t, 1);
              if (ptx::elect_one()) {
                const uint32_t d = Y.col_acc + 64u * b;
#pragma unroll
                for (uint32_t ks = 0; ks < 2; ++ks)
                  ptx::umma_ss<2>(d, a0desc + 2u * ks, bdesc + (uint64_t)((uint32_t)t * (kTileH >> 4) + 2u * ks), kIdescH, ks);
                umma_commit_pair(misc + kBarAccFull + 8u * b, pair_mask);
              }
              __syncwarp();
            }
            if (ptx::elect_one()) umma_commit_pair(misc + kBarWEmpty + 8u * stage, pair_mask);
            __syncwarp();
            ++ws;
          }
        }
        // ---- hidden layers 2..L and the output layer (l == L + 1): A = this CTA's K-half of the previous activation ----
        for (int l = 2; l <= L + 1; ++l, ++lo) {
          const bool in_y = ((l - 1) & 1) == 0;          // hidden 2, 4 live in the Y region (TMEM + smem), 1, 3 in X
          const bool outl = l == L + 1;
          const uint32_t par_in = lo & 1u;
          if (outl && w == 1) {
            // no output use for issuer 1, but it must still SEE this phase of every k-block's barrier
            for (uint32_t kb = 0; kb < (uint32_t)NCH; ++kb) ptx::mbar_wait(misc + kBarActReady + 8u * kb, par_in, hc);
            continue;
          }
          const uint32_t idesc = outl ? kIdescO : kIdescH, tile16 = (outl ? kTileO : kTileH) >> 4;
          for (int i = 0; i < (outl ? 1 : NCH); ++i) {
            [[maybe_unused]] const uint32_t j = (uint32_t)NCH + (uint32_t)(l - 2) * 2u * NCH + 2u * (uint32_t)i + (outl ? 0u : w);   // index of the use within the step
            const uint32_t b = outl ? 0u : w, d = Y.col_acc + 64u * b;
            const uint32_t n = nb[b] + (outl ? cb[0] - 1u : kL1b[b] + (uint32_t)(l - 2) * NCH + (uint32_t)i);      // the n-th use of buffer b
            const bool first = i == 0;                            // first use of the layer (the output layer's only one)
            const bool after_own = w == 0 && first && l == 2;     // issuer 0's F0 of layer 2 follows its own layer-1 uses: no token
            if (lane == 0) BCMPC_DTRACE(a.trace, (uint32_t)h, w ? 3 : 0, j, 0);
            // ---- what this use depends on, waited for while the other issuer feeds the pipe ----
            // (first use of a layer: the first pair of input k-blocks before the accumulator, so that this wait never runs
            // ahead of the OTHER issuer's last use of the buffer - a layer-1 chunk - see the note on phases above)
            // Layer 2's first uses wait for ALL of hidden 1: the previous use of buffer 1 is a layer-1 chunk issued by the OTHER
            // warp, and only "every layer-1 chunk is stored" says that it has been drained; later layers follow this warp's own
            // use of the buffer and start on the first pair of k-blocks.
            const int kb_first = first ? (l == 2 ? NCH : (NCH >= 2 ? 2 : 1)) : 0;      // input k-blocks waited for before the first MMA
            // The weight stages first: they were requested a whole use ago and have normally landed, but even a wait on a
            // completed phase costs ~165 cycles - taken here, they overlap the drain of the accumulator's previous use instead of
            // sitting behind it on the chain that paces the kernel (issue -> 1024 cycles of MMAs -> commit -> drain -> "empty"
            // -> issue of the buffer's next use: 2450 cycles per F/O pair before this reordering, the pipe needs 2048).
            uint32_t stage[SPU];
#pragma unroll
            for (int sg = 0; sg < SPU; ++sg) stage[sg] = first_g + (ws + (uint32_t)sg) % nst_g;
            auto wait_weights = [&]() {
#pragma unroll
              for (int sg = 0; sg < SPU; ++sg) ptx::mbar_wait(misc + kBarWFull + 8u * stage[sg], ((ws + (uint32_t)sg) / nst_g) & 1u);
              if (lane == 0) BCMPC_DTRACE(a.trace, (uint32_t)h, w ? 3 : 0, j, 4);
            };
            if constexpr (kWeightsFirst) wait_weights();
            for (int kb = 0; kb < kb_first; kb += 2) {
              if (kb + 1 < kb_first) ptx::mbar_wait2(misc + kBarActReady + 8u * (uint32_t)kb, misc + kBarActReady + 8u * (uint32_t)(kb + 1), par_in, hc);
              else ptx::mbar_wait(misc + kBarActReady + 8u * (uint32_t)kb, par_in, hc);
            }
            ptx::mbar_wait(misc + kBarAccEmpty + 8u * b, (n & 1u) ^ 1u, hc);        // the buffer's previous use has been drained
            if (lane == 0) BCMPC_DTRACE(a.trace, (uint32_t)h, w ? 3 : 0, j, 3);
            if constexpr (!kWeightsFirst) wait_weights();
            // the four k-steps of k-block kb: A from TMEM, or from smem for the tail of a Y activation; B = tile kb % 4 of stage kb / 4
            auto issue_kb = [&](int kb) {
              const uint64_t bt = ptx::umma_desc_sw128(base + Y.off_ring + stage[kb >> 2] * kStageBytes) + (uint64_t)((uint32_t)(kb & 3) * tile16);
              if (in_y && kb >= NT) {
                const uint64_t ad = ptx::umma_desc_sw128(base + Y.off_y + (uint32_t)(kb - NT) * kBlockBytes);
#pragma unroll
                for (uint32_t ks = 0; ks < 4; ++ks) ptx::umma_ss<2>(d, ad + 2u * ks, bt + 2u * ks, idesc, (kb | (int)ks) ? 1u : 0u);
              } else {
                const uint32_t ac = (in_y ? Y.col_y : Y.col_x) + 32u * (uint32_t)kb;
#pragma unroll
                for (uint32_t ks = 0; ks < 4; ++ks) ptx::umma_ts<2>(d, ac + 8u * ks, bt + 2u * ks, idesc, (kb | (int)ks) ? 1u : 0u);
              }
            };
            if (first) {
              // k-block pairs as the epilogues of the previous layer store them: both issuers run this concurrently (F0 and
              // O0, or the output layer alone), so most of the layer's first uses is done when its last input block lands
              if (w == 0 && !after_own) tok_wait();      // (issuer 1's last O use is issued)
              if (lane == 0) BCMPC_DTRACE(a.trace, (uint32_t)h, w ? 3 : 0, j, 1);
              for (int kb = 0; kb < NCH; kb += 2) {
                if (kb >= kb_first) {
                  if (kb + 1 < NCH) ptx::mbar_wait2(misc + kBarActReady + 8u * (uint32_t)kb, misc + kBarActReady + 8u * (uint32_t)(kb + 1), par_in, hc);
                  else ptx::mbar_wait(misc + kBarActReady + 8u * (uint32_t)kb, par_in, hc);
                }
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                  issue_kb(kb);
                  if (kb + 1 < NCH) issue_kb(kb + 1);
                }
                __syncwarp();
              }
              if (lane == 0) BCMPC_DTRACE(a.trace, (uint32_t)h, w ? 3 : 0, j, 5);
              if (ptx::elect_one()) {
                if (!outl) ptx::mbar_arrive(tok_other());
#pragma unroll
                for (int sg = 0; sg < SPU; ++sg) umma_commit_pair(misc + kBarWEmpty + 8u * stage[sg], pair_mask);
                umma_commit_pair(misc + kBarAccFull + 8u * b, pair_mask);
              }
              __syncwarp();
              if (!outl) ++nsent;
              if (w == 1) tok_wait();      // (consume F0's token: the counts stay in step)
            } else {
              // ---- the token: the other issuer has issued its use ----
              tok_wait();
              ptx::tc_fence_after();
              if (lane == 0) BCMPC_DTRACE(a.trace, (uint32_t)h, w ? 3 : 0, j, 1);
              if (ptx::elect_one()) {
#pragma unroll
                for (int sg = 0; sg < SPU; ++sg) {
#pragma unroll
                  for (int t = 0; t < 4; ++t)
                    if (4 * sg + t < NCH) issue_kb(4 * sg + t);
                  if (sg == SPU - 1) ptx::mbar_arrive(tok_other());       // all MMAs of the use are issued: the other issuer's turn
                  umma_commit_pair(misc + kBarWEmpty + 8u * stage[sg], pair_mask);
                }
                umma_commit_pair(misc + kBarAccFull + 8u * b, pair_mask);
              }
              __syncwarp();
              ++nsent;
            }
            ws += SPU;
            if (lane == 0) BCMPC_DTRACE(a.trace, (uint32_t)h, w ? 3 : 0, j, 2);
          }
        }
      }
    } else {
      // ===================== peer CTA: tell the leader when this CTA's weight halves (of this warp's sub-ring) have landed ====
      const int per_step = w == 0 ? SPU * (2 + (L - 1) * NCH) : SPU * (L - 1) * NCH;
      uint32_t ws = 0;
      for (int h = 0; h < Htot; ++h)
        for (int s = 0; s < per_step; ++s, ++ws) {
          const uint32_t stage = first_g + ws % nst_g;
          ptx::mbar_wait(misc + kBarWFull + 8u * stage, (ws / nst_g) & 1u);
          if (ptx::elect_one()) ptx::mbar_arrive_remote(misc_leader + kBarWFull + 8u * stage);
          __syncwarp();
        }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: group g drains accumulator buffer g =====================
    const int q = warp & 3, g = (warp - 4) >> 2, row = q * 32 + lane;
    const bool is_final = g == 0;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t a0_row = base + Y.off_a0 + (uint32_t)row * 128u;
    const float* s_vec = reinterpret_cast<const float*>(smem + Y.off_vec);
    unsigned char* s_red = smem + Y.off_misc + kOffRed;
    [[maybe_unused]] const bool tr = (warp == 4 || warp == 8) && lane == 0;
    uint32_t nuse = 0;               // uses of THIS group's accumulator buffer (buffer g) so far
    uint32_t step = 0;               // running step of this CTA
    auto arrive_leader = [&](uint32_t bar) { ptx::mbar_arrive_remote(misc_leader + bar); };
    // this group's next accumulator -> registers (64 or 32 columns), then release the buffer
    auto drain = [&](uint32_t (&v0)[32], uint32_t (&v1)[32], bool both) {
      const uint32_t ab = (uint32_t)g;
      ptx::mbar_wait(misc + kBarAccFull + 8u * ab, nuse & 1u, hc);
      ++nuse;
      ptx::tc_fence_after();
      const uint32_t col = Y.col_acc + 64u * ab;
      ptx::tmem_ld32(lane_off + col, v0);
      if (both) ptx::tmem_ld32(lane_off + col + 32u, v1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(kBarAccEmpty + 8u * ab);
    };
    // 64 activations of this row (bf16 pairs w0 | w1) -> k-block kb of the X or Y activation, then tell the issuer
    auto store_block = [&](bool to_y, uint32_t kb, const uint32_t (&w0)[16], const uint32_t (&w1)[16]) {
      if (to_y && kb >= (uint32_t)NT) {
        const uint32_t rb = base + Y.off_y + (kb - (uint32_t)NT) * kBlockBytes + (uint32_t)row * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t addr = rb + (uint32_t)(((c ^ (row & 7)) << 4));
          const uint32_t* w = c < 4 ? &w0[4 * c] : &w1[4 * (c - 4)];
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        }
        ptx::fence_proxy_async_smem();
      } else {
        const uint32_t dst = (to_y ? Y.col_y : Y.col_x) + 32u * kb;
        ptx::tmem_st16(lane_off + dst, w0);
        ptx::tmem_st16(lane_off + dst + 16u, w1);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) arrive_leader(kBarActReady + 8u * kb);
    };
    // ---- LayerNorm variant: per-row statistics ----
    [[maybe_unused]] float ln_s1 = 0.f, ln_s2 = 0.f;            // this thread's partial (sum, sum of squares) of the layer being produced
    [[maybe_unused]] float rho = 1.f, rmu = 0.f;                // 1 / sqrt(var + eps), rho * mean of the layer being consumed
    [[maybe_unused]] float2* s_rem = reinterpret_cast<float2*>(smem + Y.off_stat);               // [slot (4)][row]: the partner's partials
    [[maybe_unused]] float2* s_loc = reinterpret_cast<float2*>(smem + Y.off_stat) + 4 * 128;       // [group][row]: this CTA's partials
    // two neighbouring units: activation, statistics (ReLU: of the fp32 values; tanh: of the bf16-rounded values the tensor core
    // will read, as in rollout_tmem_kernel), bf16 pair
    [[maybe_unused]] auto relu_stat2 = [&](float x, float y) -> uint32_t {
      if constexpr (TANH) {
        const uint32_t t = ptx::tanh_bf16x2(tmx::pack_bf16(x, y));
        const float lo = __uint_as_float(t << 16), hi = __uint_as_float(t & 0xFFFF0000u);
        ln_s1 += lo + hi;
        ln_s2 = fmaf(lo, lo, fmaf(hi, hi, ln_s2));
        return t;
      } else {
        x = x < 0.f ? 0.f : x;
        y = y < 0.f ? 0.f : y;
        ln_s1 += x + y;
        ln_s2 = fmaf(x, x, fmaf(y, y, ln_s2));
        return tmx::pack_bf16(x, y);
      }
    };
    // this thread's partial of hidden layer l -> the partner CTA's slot (8 bytes, reported to its barrier of layer l)
    [[maybe_unused]] auto send_stats = [&](int l) {
      const uint32_t dst = base_partner + Y.off_stat + (uint32_t)(((l == 1 ? g : 2 + (l & 1)) * 128 + row) * 8);
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(dst), "r"(__float_as_uint(ln_s1)),
                   "r"(__float_as_uint(ln_s2)), "r"(misc_partner + kBarStat + 8u * (uint32_t)(l - 1))
                   : "memory");
    };
    // ... -> this CTA's slot [group], for the other epilogue group
    [[maybe_unused]] auto publish_local = [&](uint32_t which) {
      s_loc[g * 128 + row] = make_float2(ln_s1, ln_s2);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(misc + kBarStatLoc + 8u * which);      // (release: the slot writes of this warp are visible to the waiter)
    };
    [[maybe_unused]] auto set_row_stats = [&](float t1, float t2) {
      const float mean = t1 * a.ln_inv_n;
      const float var = fmaxf(t2 * a.ln_inv_n - mean * mean, 0.f);
      rho = rsqrtf(var + a.ln_eps);
      rmu = rho * mean;
    };
    for (int it = 0; it < n_it; ++it) {      // ---- one work unit: this CTA's 128 paths of one root state, H steps ----
      const int unit = first_unit + it * unit_stride;
      const int b = unit / units_per_state, tile = (unit % units_per_state) * 2 + (int)qrow;
      const int k = tile * kRows + row;
      const bool valid = k < p.K;
      const int kc = valid ? k : p.K - 1;
      const bool writer = khalf == 0;                              // both K-halves carry the state; one reports
      float s[kD], z[kA], act0[kA], cost = 0.f;
      if (is_final) {
#pragma unroll
        for (int d = 0; d < kD; ++d) s[d] = (float)root_state(p, (size_t)((DBG && p.per_row_states) ? kc : b), kD, d);
        tmx::get_noise<DBG, POL>(a, b, 0, kc, z);
        if constexpr (DBG) {
          if (writer && p.states_out && valid)
            for (int d = 0; d < kD; ++d) p.states_out[(size_t)k * kD + d] = (double)s[d];
        }
      }
      for (int h = 0; h < H; ++h, ++step) {
        const uint32_t xbase = step * xs_step;                    // exchange index of the step's first exchanged chunk
        [[maybe_unused]] float act[kA];
        if constexpr (POL) {
          // ---- the action of this step: proposal network on the tensor cores (ppo_bc_policy.py:60,70-72 / behavioral_cloning.py:21-26) ----
          if (is_final) {
            tmx::build_policy_input(a, a0_row, row, s);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) arrive_leader(kBarPolIn);
          }
          const float* s_pv = reinterpret_cast<const float*>(smem + Y.off_pol + kPolImgBytes);
          // chunk g of a tanh layer: accumulator + bias -> tanh -> bf16 -> TMEM operand of the next policy layer
          auto pol_chunk = [&](int layer) {
            const uint32_t j = (uint32_t)g;
            ptx::mbar_wait(misc + kBarPdFull + 8u * ((uint32_t)(layer - 1) * 2u + j), step & 1u, hc);
            ptx::tc_fence_after();
            uint32_t v0[32], v1[32], w0[16], w1[16];
            ptx::tmem_ld32(lane_off + Y.col_acc + 64u * j, v0);
            ptx::tmem_ld32(lane_off + Y.col_acc + 64u * j + 32u, v1);
            ptx::tmem_ld_wait();
            const float4* b4 = reinterpret_cast<const float4*>(s_pv + (layer == 1 ? 0 : tmx::kPolMaxHidden) + 64 * j);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = b4[i];
              w0[2 * i] = ptx::tanh_bf16x2(tmx::pack_bf16(__uint_as_float(v0[4 * i]) + bb.x, __uint_as_float(v0[4 * i + 1]) + bb.y));
              w0[2 * i + 1] = ptx::tanh_bf16x2(tmx::pack_bf16(__uint_as_float(v0[4 * i + 2]) + bb.z, __uint_as_float(v0[4 * i + 3]) + bb.w));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = b4[8 + i];
              w1[2 * i] = ptx::tanh_bf16x2(tmx::pack_bf16(__uint_as_float(v1[4 * i]) + bb.x, __uint_as_float(v1[4 * i + 1]) + bb.y));
              w1[2 * i + 1] = ptx::tanh_bf16x2(tmx::pack_bf16(__uint_as_float(v1[4 * i + 2]) + bb.z, __uint_as_float(v1[4 * i + 3]) + bb.w));
            }
            const uint32_t dst = (layer == 1 ? 0u : 64u) + 32u * j;
            ptx::tmem_st16(lane_off + dst, w0);
            ptx::tmem_st16(lane_off + dst + 16u, w1);
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader((layer == 1 ? kBarAp1 : kBarAp2) + 8u * j);
          };
          if (g < a.np1) pol_chunk(1);
          if (g < a.np2) pol_chunk(2);
          if (is_final) {
            ptx::mbar_wait(misc + kBarPolFull, step & 1u, hc);
            ptx::tc_fence_after();
            uint32_t v[32];
            ptx::tmem_ld32(lane_off + Y.col_acc, v);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
#pragma unroll
            for (int j = 0; j < kA; ++j) {
              const float mean = __uint_as_float(v[j]) + s_pv[2 * tmx::kPolMaxHidden + j];
              const int prop = h == 0 ? a.proposal0 : a.proposal;           // (the tree-search planner switches the mode after its first stage)
              if (prop == BCMPC_PROP_POLICY_GAUSSIAN) act[j] = fmaf(a.pol_std[j], z[j], mean);                        // distributions.py:80
              else if (prop == BCMPC_PROP_POLICY_MIX) act[j] = fmaf(a.one_minus_explore, mean, a.explore * z[j]);   // controllers.py:208
              else if (prop == BCMPC_PROP_POLICY_MEAN) act[j] = mean;                                                // controllers.py:429,433
              else if (prop == BCMPC_PROP_UNIFORM) act[j] = z[j];                                                    // controllers.py:411 (random first stage)
              else act[j] = mean + z[j];                                                                             // controllers.pyc (BC + U[-1,1))
            }
          }
        } else {
          if (is_final) {
#pragma unroll
            for (int j = 0; j < kA; ++j) act[j] = z[j];              // controllers.py:70: the sampled action itself
          }
        }
        if (is_final) {
          if (h == 0) {
#pragma unroll
            for (int j = 0; j < kA; ++j) act0[j] = act[j];
          }
          if constexpr (DBG) {
            if (writer && p.actions_out && valid)
              for (int j = 0; j < kA; ++j) p.actions_out[((size_t)h * p.K + k) * kA + j] = (double)act[j];
          }
          tmx::build_input(a, a0_row, row, s, act);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) arrive_leader(kBarA0Ready);
          if (tr) BCMPC_DTRACE(a.trace, step, 1, 0, 7);
        }
        // ---- layer 1: chunk c -> X block c (bias rides in the GEMM) ----
#pragma unroll 1
        for (uint32_t c = (uint32_t)g; c < (uint32_t)NCH; c += 2) {
          uint32_t v0[32], v1[32], w0[16], w1[16];
          if (tr) BCMPC_DTRACE(a.trace, step, 1 + g, c, 0);
          drain(v0, v1, true);
          if (tr) BCMPC_DTRACE(a.trace, step, 1 + g, c, 1);
          if constexpr (LN) {
#pragma unroll
            for (int i = 0; i < 16; ++i) w0[i] = relu_stat2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
#pragma unroll
            for (int i = 0; i < 16; ++i) w1[i] = relu_stat2(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) w0[i] = act2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
#pragma unroll
            for (int i = 0; i < 16; ++i) w1[i] = act2(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
          }
          store_block(false, c, w0, w1);
          if (tr) BCMPC_DTRACE(a.trace, step, 1 + g, c, 4);
        }
        [[maybe_unused]] float own1 = 0.f, own2 = 0.f;            // group 1: its own partial of the layer whose statistics it needs next
        if constexpr (LN) {
          // hidden 1 is converted by BOTH groups: each sends its partial to the partner CTA; group 0's also goes to group 1
          send_stats(1);
          if (g == 0) publish_local(0);
          else { own1 = ln_s1; own2 = ln_s2; }
          ln_s1 = 0.f; ln_s2 = 0.f;
        }
        // the next step's randomness, drawn where group 0 has slack
        if (is_final && h + 1 < H) tmx::get_noise<DBG, POL>(a, b, h + 1, kc, z);
        // ---- hidden layers 2..L ----
        for (int l = 2; l <= L; ++l) {
          const bool to_y = (l & 1) == 0;
          const float* bias_l = s_vec + (l - 2) * 64 * NCH;
#pragma unroll 1
          for (uint32_t i = 0; i < (uint32_t)NCH; ++i) {
            // two messages per chunk (columns 0..31, 32..63): message x uses receive slot x & 1, phase (x >> 1) & 1
            const uint32_t x0 = xbase + ((uint32_t)(l - 2) * NCH + i) * 2u;
            [[maybe_unused]] const uint32_t tuse = (uint32_t)NCH + (uint32_t)(l - 2) * 2u * NCH + 2u * i + (uint32_t)g;
            if (tr) BCMPC_DTRACE(a.trace, step, 1 + g, tuse, 0);
            if (g == 0) {
              // F: my partial of the PARTNER's chunk i -> the partner's receive slots
              uint32_t v0[32], v1[32];
              drain(v0, v1, true);
              if (tr) BCMPC_DTRACE(a.trace, step, 1, tuse, 1);
#pragma unroll
              for (uint32_t hf = 0; hf < 2; ++hf) {
                const uint32_t x = x0 + hf, slot = x & 1u, xph = (x >> 1) & 1u;
                ptx::mbar_wait(misc + kBarRecvEmpty + 8u * slot, xph ^ 1u, hc);        // the partner has read the slot's previous content
                if (tr && hf == 0) BCMPC_DTRACE(a.trace, step, 1, tuse, 2);
                const uint32_t dst = base_partner + Y.off_recv + slot * kRecvBytes + (uint32_t)row * 16u;
                const uint32_t rbar = misc_partner + kBarRecvFull + 8u * slot;
                const uint32_t(&v)[32] = hf ? v1 : v0;
                if constexpr (kPack24) {
                  uint32_t w[24];
#pragma unroll
                  for (int j = 0; j < 8; ++j) pack24(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], w[3 * j], w[3 * j + 1], w[3 * j + 2]);
#pragma unroll
                  for (int j = 0; j < 6; ++j) st_async_v4(dst + (uint32_t)j * 2048u, rbar, w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) st_async_v4(dst + (uint32_t)j * 2048u, rbar, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
              }
              if (tr) BCMPC_DTRACE(a.trace, step, 1, tuse, 3);
            } else {
              // O: own chunk i = my partial + the partner's + bias -> ReLU -> block i of this layer's output
              uint32_t v0[32], v1[32], w0[16], w1[16];
              drain(v0, v1, true);
              // arm both halves (16 KB each will arrive) only now: this accumulator exists, so the layer's input does, so the
              // previous step's output exchange - the last message the OTHER group armed on one of these slots - is complete
              if (warp == 8 && ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(misc + kBarRecvFull + 8u * (x0 & 1u), kMsgBytes);
                ptx::mbar_arrive_expect_tx(misc + kBarRecvFull + 8u * ((x0 + 1u) & 1u), kMsgBytes);
              }
              __syncwarp();
              if (tr) BCMPC_DTRACE(a.trace, step, 2, tuse, 1);
              [[maybe_unused]] const float4* b4 = reinterpret_cast<const float4*>(bias_l + 64u * i);
              if constexpr (LN) {
                if (i == 0) {
                  // statistics of hidden layer l - 1, whose LayerNorm output this layer consumes: own partial + (hidden 1: group 0's) +
                  // the partner CTA's.  The waits sit behind this use's drain, i.e. ~1000 cycles of MMAs after the partials were sent.
                  // ((mine) + (theirs): the partner CTA adds the same two numbers the other way round - identical statistics)
                  const int lp = l - 1;
                  float m1 = own1, m2 = own2;
                  if (lp == 1) {
                    ptx::mbar_wait(misc + kBarStatLoc, step & 1u, hc);
                    const float2 o = s_loc[row];
                    m1 += o.x; m2 += o.y;
                  }
                  ptx::mbar_wait(misc + kBarStat + 8u * (uint32_t)(lp - 1), step & 1u, hc);
                  if (warp == 8 && ptx::elect_one()) ptx::mbar_arrive_expect_tx(misc + kBarStat + 8u * (uint32_t)(lp - 1), lp == 1 ? 2048u : 1024u);   // next step's
                  __syncwarp();
                  float2 th = s_rem[(lp == 1 ? 1 : 2 + (lp & 1)) * 128 + row];
                  if (lp == 1) { const float2 r0 = s_rem[row]; th.x += r0.x; th.y += r0.y; }
                  set_row_stats(m1 + th.x, m2 + th.y);
                }
              }
              [[maybe_unused]] const float4* u4 = reinterpret_cast<const float4*>(s_vec + (l - 2) * 2 * 64 * NCH + 64u * i);
              [[maybe_unused]] const float4* c4 = reinterpret_cast<const float4*>(s_vec + (l - 2) * 2 * 64 * NCH + 64 * NCH + 64u * i);
#pragma unroll
              for (uint32_t hf = 0; hf < 2; ++hf) {
                const uint32_t x = x0 + hf, slot = x & 1u, xph = (x >> 1) & 1u;
                // (a CTA-scope acquire like every consumer of TMA-written shared memory uses: a cluster-scope acquire costs ~370
                // cycles per wait for an L1 invalidation that shared memory does not need)
                ptx::mbar_wait(misc + kBarRecvFull + 8u * slot, xph, hc);
                if (tr && hf == 0) BCMPC_DTRACE(a.trace, step, 2, tuse, 2);
                const uint32_t src = base + Y.off_recv + slot * kRecvBytes + (uint32_t)row * 16u;
                const uint32_t(&v)[32] = hf ? v1 : v0;
                uint32_t(&w)[16] = hf ? w1 : w0;
                float4 r[8];
                if constexpr (kPack24) {
                  float4 q[6];
#pragma unroll
                  for (int j = 0; j < 6; ++j) q[j] = ld_shared_v4(src + (uint32_t)j * 2048u);    // all six loads in flight before the first use
                  const uint32_t* qw = reinterpret_cast<const uint32_t*>(q);
#pragma unroll
                  for (int j = 0; j < 8; ++j) unpack24(qw[3 * j], qw[3 * j + 1], qw[3 * j + 2], r[j].x, r[j].y, r[j].z, r[j].w);
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) r[j] = ld_shared_v4(src + (uint32_t)j * 2048u);    // all eight loads in flight before the first use
                }
                if constexpr (LN) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 uu = u4[8 * hf + j], cc = c4[8 * hf + j];
                    auto pre = [&](uint32_t acc, float part, float u, float c) { return fmaf(rho, __uint_as_float(acc) + part, fmaf(-rmu, u, c)); };
                    w[2 * j] = relu_stat2(pre(v[4 * j], r[j].x, uu.x, cc.x), pre(v[4 * j + 1], r[j].y, uu.y, cc.y));
                    w[2 * j + 1] = relu_stat2(pre(v[4 * j + 2], r[j].z, uu.z, cc.z), pre(v[4 * j + 3], r[j].w, uu.w, cc.w));
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 bb = b4[8 * hf + j];
                    w[2 * j] = act2(__uint_as_float(v[4 * j]) + r[j].x + bb.x, __uint_as_float(v[4 * j + 1]) + r[j].y + bb.y);
                    w[2 * j + 1] = act2(__uint_as_float(v[4 * j + 2]) + r[j].z + bb.z, __uint_as_float(v[4 * j + 3]) + r[j].w + bb.w);
                  }
                }
                __syncwarp();                                                         // every lane's loads are in its registers (and used)
                if (lane == 0) ptx::mbar_arrive_remote(misc_partner + kBarRecvEmpty + 8u * slot);   // my slot may be overwritten
              }
              if (tr) BCMPC_DTRACE(a.trace, step, 2, tuse, 3);
              store_block(to_y, i, w0, w1);
              if (tr) BCMPC_DTRACE(a.trace, step, 2, tuse, 4);
              if constexpr (LN) {
                if (i == (uint32_t)NCH - 1) {      // this CTA's units of hidden layer l are complete
                  send_stats(l);
                  if (l == L) publish_local(1);    // the output layer is finalised by group 0
                  own1 = ln_s1; own2 = ln_s2;
                  ln_s1 = 0.f; ln_s2 = 0.f;
                }
              }
            }
          }
        }
        // ---- output layer: the two K-half partials of the 20 deltas are exchanged both ways ----
        if (is_final) {
          const uint32_t x = xbase + (uint32_t)(L - 1) * NCH * 2u, slot = x & 1u, xph = (x >> 1) & 1u;
          uint32_t v[32], dummy[32];
          [[maybe_unused]] const uint32_t tuse = (uint32_t)NCH + (uint32_t)(L - 1) * 2u * NCH;
          if (tr) BCMPC_DTRACE(a.trace, step, 1, tuse, 0);
          drain(v, dummy, false);
          if (tr) BCMPC_DTRACE(a.trace, step, 1, tuse, 1);
          if constexpr (LN) {
            // statistics of the last hidden layer: group 1's partial of this CTA + the partner CTA's.  Read BEFORE this CTA's output
            // partial is sent: the partner starts its next step only after that message, and its next statistics may land in the same slot.
            ptx::mbar_wait(misc + kBarStatLoc + 8u, step & 1u, hc);
            ptx::mbar_wait(misc + kBarStat + 8u * (uint32_t)(L - 1), step & 1u, hc);
            if (warp == 4 && ptx::elect_one()) ptx::mbar_arrive_expect_tx(misc + kBarStat + 8u * (uint32_t)(L - 1), 1024u);   // next step's
            __syncwarp();
            const float2 mine = s_loc[128 + row], th = s_rem[(2 + (L & 1)) * 128 + row];
            set_row_stats(mine.x + th.x, mine.y + th.y);
          }
          // arm the receive barrier only now: the output accumulator is complete, so every hidden exchange of the step -
          // in particular the slot's previous one, which the OTHER group waited for - has been consumed
          if (warp == 4 && ptx::elect_one()) ptx::mbar_arrive_expect_tx(misc + kBarRecvFull + 8u * slot, 5u * 2048u);   // 20 columns will arrive
          __syncwarp();
          ptx::mbar_wait(misc + kBarRecvEmpty + 8u * slot, xph ^ 1u, hc);
          const uint32_t dst = base_partner + Y.off_recv + slot * kRecvBytes + (uint32_t)row * 16u;
          const uint32_t rbar = misc_partner + kBarRecvFull + 8u * slot;
#pragma unroll
          for (int j = 0; j < kD / 4; ++j) st_async_v4(dst + (uint32_t)j * 2048u, rbar, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (tr) BCMPC_DTRACE(a.trace, step, 1, tuse, 3);
          ptx::mbar_wait(misc + kBarRecvFull + 8u * slot, xph, hc);
          if (tr) BCMPC_DTRACE(a.trace, step, 1, tuse, 4);
          const uint32_t src = base + Y.off_recv + slot * kRecvBytes + (uint32_t)row * 16u;
          const float* b3 = s_vec + (LN ? (L - 1) * 2 * 64 * NCH : (L - 1) * 64 * NCH);      // LN: [u3 (32)][c3 (32)]
          float y[kD];
#pragma unroll
          for (int j = 0; j < kD / 4; ++j) {
            const float4 r = ld_shared_v4(src + (uint32_t)j * 2048u);
            if constexpr (LN) {
              auto fin = [&](uint32_t acc, float part, int d) { return fmaf(rho, __uint_as_float(acc) + part, fmaf(-rmu, b3[d], b3[32 + d])); };
              y[4 * j] = fin(v[4 * j], r.x, 4 * j);
              y[4 * j + 1] = fin(v[4 * j + 1], r.y, 4 * j + 1);
              y[4 * j + 2] = fin(v[4 * j + 2], r.z, 4 * j + 2);
              y[4 * j + 3] = fin(v[4 * j + 3], r.w, 4 * j + 3);
            } else {
              y[4 * j] = (__uint_as_float(v[4 * j]) + r.x) + b3[4 * j];
              y[4 * j + 1] = (__uint_as_float(v[4 * j + 1]) + r.y) + b3[4 * j + 1];
              y[4 * j + 2] = (__uint_as_float(v[4 * j + 2]) + r.z) + b3[4 * j + 2];
              y[4 * j + 3] = (__uint_as_float(v[4 * j + 3]) + r.w) + b3[4 * j + 3];
            }
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote(misc_partner + kBarRecvEmpty + 8u * slot);
          float delta[kD];
#pragma unroll
          for (int d = 0; d < kD; ++d) delta[d] = fmaf(y[d], a.std_delta[d], a.mean_delta[d]);   // dynamics.py:113
          if (s[5] >= 0.2f) cost += 10.f;                                    // cost_functions.py:16-26 (pre-step state)
          if (s[6] >= 0.f) cost += 10.f;
          if (s[7] >= 0.f) cost += 10.f;
          cost -= delta[17] * 100.f;                                         // cost_functions.py:28 (/ 0.01)
#pragma unroll
          for (int d = 0; d < kD; ++d) s[d] += delta[d];                     // dynamics.py:116
          if constexpr (DBG) {
            if (writer && p.states_out && valid)
              for (int d = 0; d < kD; ++d) p.states_out[((size_t)(h + 1) * p.K + k) * kD + d] = (double)s[d];
          }
          if (tr) BCMPC_DTRACE(a.trace, step, 1, tuse, 5);
        }
      }
      if (is_final) {
        // ---- per-tile argmin over the 128 rows (np.argmin order); reported by the K-half-0 CTAs ----
        if constexpr (DBG) {
          if (writer && p.costs_out && valid) p.costs_out[(size_t)b * p.K + k] = (double)cost;
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
        double* r_cost = reinterpret_cast<double*>(s_red);
        int64_t* r_idx = reinterpret_cast<int64_t*>(s_red + 32);
        int* r_src = reinterpret_cast<int*>(s_red + 64);
        ptx::named_bar_sync(1, 128);                                         // the previous unit's readers are done with s_red
        if (lane == 0) { r_cost[q] = c; r_idx[q] = idx; r_src[q] = src; }
        ptx::named_bar_sync(1, 128);
        c = r_cost[0]; idx = r_idx[0]; src = r_src[0];
#pragma unroll
        for (int i = 1; i < 4; ++i)
          if (r_idx[i] != INT64_MAX && (idx == INT64_MAX || cand_better(r_cost[i], r_idx[i], c, idx))) { c = r_cost[i]; idx = r_idx[i]; src = r_src[i]; }
        if (writer) {
          if (row == src && p.cand) {
            Candidate& o = p.cand[(size_t)b * p.tiles_per_state + tile];
            o.cost = c;
            o.index = idx == INT64_MAX ? -1 : idx;
#pragma unroll
            for (int j = 0; j < kA; ++j) o.action[j] = (double)act0[j];
            for (int j = kA; j < kMaxAct; ++j) o.action[j] = 0.0;
          }
          if (p.out && (src >> 5) == q) fused_state_minloc(p, b, lane, src & 31);
        }
      }
    }      // work units
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc_cg<2>(0u, 512);
}

// ---------------------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------------------
template <int NCH, bool DBG, bool TANH, bool LN = false, bool POL = false>
inline cudaError_t deep_launch_t(const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out) {
  auto kern = rollout_deep_kernel<NCH, DBG, TANH, LN, POL>;
  constexpr Layout Y = Lay<NCH, LN, POL>::value;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y.smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Y.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = max_clusters;
  if (n <= 0) {          // clusters of four that can be co-resident (GPC granularity: fewer than SMs / 4 on most parts)
    cfg.gridDim = dim3(kCluster * 64);
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = 32; }
  }
  if (clusters_out) *clusters_out = n;
  const int64_t g = units < n ? units : n;
  cfg.gridDim = dim3((unsigned)(g * kCluster));
  return cudaLaunchKernelEx(&cfg, kern, a);
}
cudaError_t deep_launch_prod(int nch, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_dbg(int nch, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_tanh_prod(int nch, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_tanh_dbg(int nch, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_ln_prod(int nch, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_ln_dbg(int nch, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_ln_tanh_prod(int nch, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_ln_tanh_dbg(int nch, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
// policy-guided shooting: 2..8 chunks per K-half (hidden 129..1024), ReLU or tanh, with or without LayerNorm
constexpr int kPolMinNch = 2, kPolMaxNch = kMaxNch;
cudaError_t deep_launch_pol_prod(int nch, bool ln, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_pol_dbg(int nch, bool ln, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_pol_tanh_prod(int nch, bool ln, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
cudaError_t deep_launch_pol_tanh_dbg(int nch, bool ln, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out);
inline cudaError_t deep_launch(int nch, bool tanh_act, bool layer_norm, bool policy, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out) {
  if (policy) {
    if (nch < kPolMinNch || nch > kPolMaxNch) return cudaErrorInvalidValue;
    if (tanh_act)
      return tmx::tm_needs_debug_flavour(a.t.p) ? deep_launch_pol_tanh_dbg(nch, layer_norm, a, max_clusters, units, st, clusters_out)
                                                 : deep_launch_pol_tanh_prod(nch, layer_norm, a, max_clusters, units, st, clusters_out);
    return tmx::tm_needs_debug_flavour(a.t.p) ? deep_launch_pol_dbg(nch, layer_norm, a, max_clusters, units, st, clusters_out)
                                               : deep_launch_pol_prod(nch, layer_norm, a, max_clusters, units, st, clusters_out);
  }
  if (layer_norm && tanh_act)
    return tmx::tm_needs_debug_flavour(a.t.p) ? deep_launch_ln_tanh_dbg(nch, a, max_clusters, units, st, clusters_out)
                                               : deep_launch_ln_tanh_prod(nch, a, max_clusters, units, st, clusters_out);
  if (layer_norm) {
    return tmx::tm_needs_debug_flavour(a.t.p) ? deep_launch_ln_dbg(nch, a, max_clusters, units, st, clusters_out)
                                               : deep_launch_ln_prod(nch, a, max_clusters, units, st, clusters_out);
  }
  if (tanh_act)
    return tmx::tm_needs_debug_flavour(a.t.p) ? deep_launch_tanh_dbg(nch, a, max_clusters, units, st, clusters_out)
                                               : deep_launch_tanh_prod(nch, a, max_clusters, units, st, clusters_out);
  return tmx::tm_needs_debug_flavour(a.t.p) ? deep_launch_dbg(nch, a, max_clusters, units, st, clusters_out)
                                             : deep_launch_prod(nch, a, max_clusters, units, st, clusters_out);
}
#define BCMPC_DEEP_DEFINE_POL(NAME, DBG, TANH)                                                                                  \
  cudaError_t NAME(int nch, bool ln, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out) { \
    switch (nch) {                                                                                                              \
      case 2: return ln ? deep_launch_t<2, DBG, TANH, true, true>(a, max_clusters, units, st, clusters_out)                     \
                        : deep_launch_t<2, DBG, TANH, false, true>(a, max_clusters, units, st, clusters_out);                   \
      case 3: return ln ? deep_launch_t<3, DBG, TANH, true, true>(a, max_clusters, units, st, clusters_out)                     \
                        : deep_launch_t<3, DBG, TANH, false, true>(a, max_clusters, units, st, clusters_out);                   \
      case 4: return ln ? deep_launch_t<4, DBG, TANH, true, true>(a, max_clusters, units, st, clusters_out)                     \
                        : deep_launch_t<4, DBG, TANH, false, true>(a, max_clusters, units, st, clusters_out);                   \
      case 5: return ln ? deep_launch_t<5, DBG, TANH, true, true>(a, max_clusters, units, st, clusters_out)                     \
                        : deep_launch_t<5, DBG, TANH, false, true>(a, max_clusters, units, st, clusters_out);                   \
      case 6: return ln ? deep_launch_t<6, DBG, TANH, true, true>(a, max_clusters, units, st, clusters_out)                     \
                        : deep_launch_t<6, DBG, TANH, false, true>(a, max_clusters, units, st, clusters_out);                   \
      case 7: return ln ? deep_launch_t<7, DBG, TANH, true, true>(a, max_clusters, units, st, clusters_out)                     \
                        : deep_launch_t<7, DBG, TANH, false, true>(a, max_clusters, units, st, clusters_out);                   \
      case 8: return ln ? deep_launch_t<8, DBG, TANH, true, true>(a, max_clusters, units, st, clusters_out)                     \
                        : deep_launch_t<8, DBG, TANH, false, true>(a, max_clusters, units, st, clusters_out);                   \
    }                                                                                                                           \
    return cudaErrorInvalidValue;                                                                                               \
  }
#define BCMPC_DEEP_DEFINE(NAME, DBG, TANH) BCMPC_DEEP_DEFINE_LN(NAME, DBG, TANH, false)
#define BCMPC_DEEP_DEFINE_LN(NAME, DBG, TANH, LN)                                                                                        \
  cudaError_t NAME(int nch, const DeepArgs& a, int max_clusters, int64_t units, cudaStream_t st, int* clusters_out) {      \
    switch (nch) {                                                                                                          \
      case 1: return deep_launch_t<1, DBG, TANH, LN>(a, max_clusters, units, st, clusters_out);                                 \
      case 2: return deep_launch_t<2, DBG, TANH, LN>(a, max_clusters, units, st, clusters_out);                                 \
      case 3: return deep_launch_t<3, DBG, TANH, LN>(a, max_clusters, units, st, clusters_out);                                 \
      case 4: return deep_launch_t<4, DBG, TANH, LN>(a, max_clusters, units, st, clusters_out);                                 \
      case 5: return deep_launch_t<5, DBG, TANH, LN>(a, max_clusters, units, st, clusters_out);                                 \
      case 6: return deep_launch_t<6, DBG, TANH, LN>(a, max_clusters, units, st, clusters_out);                                 \
      case 7: return deep_launch_t<7, DBG, TANH, LN>(a, max_clusters, units, st, clusters_out);                                 \
      case 8: return deep_launch_t<8, DBG, TANH, LN>(a, max_clusters, units, st, clusters_out);                                 \
    }                                                                                                                       \
    return cudaErrorInvalidValue;                                                                                           \
  }

}  // namespace dpx
}  // namespace bcmpc
