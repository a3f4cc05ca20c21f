// BF16 fast-mode fused rollout kernel (BCMPC_PREC_BF16_FAST): tcgen05.mma with TMEM-resident activations.
//
// One CTA owns 128 paths (one TMEM lane per path) for all H steps, then takes the next 128 (persistent CTAs: the grid is
// one wave and every CTA group rolls its share of the (root state, tile) units back to back).  With CG = 2 two CTAs of a cluster form a
// cta_group::2 pair: 256 paths per UMMA (M = 256), every weight tile is fetched once per pair (each CTA loads
// half of its rows) and read by both tensor cores, which halves the L2 -> SM weight stream per path.
//
// Where the data lives (per CTA):
//   registers   state s[20] (fp32), running cost, first action of the path that owns the TMEM lane
//   TMEM (512 columns x 128 lanes)
//     [  0,256)  A1   hidden-1 activations, bf16 pairs, the A operand of layer 2  (tcgen05.mma A-from-TMEM)
//     [256,448)  D    3 x 64-column fp32 accumulators (layer 1 / layer 2 output chunks)
//     [448,480)  A2   one 64-unit chunk of hidden-2 activations (bf16 pairs), the A operand of layer 3
//     [480,512)  D3   output-layer accumulator (20 of 32 columns used)
//   shared memory
//     A0   [128 x 64] bf16 K-major SWIZZLE_128B: normalised (state, action, 1, 1) rows = A operand of layer 1
//     W1, W3 resident for the whole kernel (UMMA smem images, loaded once)
//     ring of W2 tiles streamed from L2 with cp.async.bulk + mbarriers
//   no activation or state ever touches HBM.
//
// Per step (hidden width S padded to a multiple of 64, NC = chunks of 64 hidden units):
//   L1  D[c]  = A0 x W1[c]                 (2 k-steps, A from smem)      -> epilogue ReLU, bf16 -> A1[:, 64c..]
//   L2  D[c]  = A1 x W2[:, c]              (S/16 k-steps, A from TMEM)   -> epilogue ReLU, bf16 -> A2
//   L3  D3   += A2 x W3[64c.., :]          (4 k-steps per chunk)
//   final epilogue: denormalise, cheetah cost on the pre-step state, s += delta, next action (Philox), normalise
//                   -> A0 of the next step.
// Biases ride in the GEMMs (two constant-1 input columns / hidden units x a hi+lo bf16 split of the bias), so the
// hidden epilogues are ReLU + convert only.
//
// Roles: warp 0 = weight producer; warps 1 and 3 = MMA issuers of the leader CTA (warp 1: the even layer-2 chunks,
// warp 3: proposal network, layer 1 and the odd layer-2 chunks - two issuers so that one can wait on barriers while the
// other keeps the tensor pipe fed; every accumulator is written by one issuer only, so results do not depend on the
// interleaving); warp 2 = TMEM allocation + layer-3 issuer; in the peer CTA warp 1 forwards "weights landed" to the
// leader; warps 4-11 = epilogue: two groups of four warps (one warp per TMEM lane quadrant) that take alternate
// accumulator chunks (layer 1: alternate pairs of chunks), 64 columns per warp and chunk.
// The step's schedule is fixed arithmetic on the chunk index (see below), so every TMEM column, barrier and descriptor
// offset is a uniform-register function of a loop counter.
//
// Template parameters: CG (CTAs per tcgen05 group), NC (hidden chunks of 64), VAR (bit flags: 1 = proposal network,
// 2 = LayerNorm dynamics, 4 = learned-reward model with its two heads as one double-width second layer), DBG (debug
// flavour: injected randomness, per-path outputs; the production flavour compiles those paths out).
//
// Replaces dynamics.py:106-119 (+ :121-238 for the reward model) + cost_functions.py:10-30,59-63 +
// controllers.py:53,63-85 (+ :189-237, :121-158, :319-363) for HalfCheetah shapes (D = 20, A = 6), two hidden layers,
// S <= 510.
#pragma once
#include <cuda_bf16.h>

#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace bcmpc {
namespace tmx {

constexpr int kRows = 128;
constexpr int kThreads = 384;
constexpr int kChunk = 64;             // hidden units per accumulator chunk (UMMA N)
constexpr int kOutN = 32;              // output layer padded to UMMA N = 32
constexpr int kColA1 = 0, kColD = 256, kColA2 = 448, kColD3 = 480, kTmemCols = 512;
constexpr int kNumDBuf = 3;            // accumulators of the layer-2 chunks
constexpr int kNumDBufL1 = 4;          // layer 1 also borrows [448,512): A2 and D3 are idle while layer 1 runs
__host__ __device__ constexpr int buf_col(int b) { return b < kNumDBuf ? kColD + kChunk * b : kColA2; }
// TMEM columns of hidden-2 chunk c (bf16 pairs, A operand of layer 3).  Layer 3 queues behind ~1000 cycles of layer-2
// MMAs in the in-order tensor pipe, so with ONE buffer the epilogue of chunk c+1 waits that long for layer 3 to consume
// chunk c.  Networks narrower than 512 leave part of the A1 region unused: those columns become extra buffers
// (2x256 net: 4 buffers).  With a single buffer (hidden = 512) at least the LAST chunk goes elsewhere - to the
// accumulator buffer of chunk nc-3, drained and without further user in the step - so the step's tail does not wait.
__host__ __device__ constexpr int a2_buffers(int nc) { return 1 + (256 - 32 * nc) / 32 > 4 ? 4 : 1 + (256 - 32 * nc) / 32; }
// (nc = hidden-1 chunks, which decide the free A1 columns; nc2 = hidden-2 chunks)
__host__ __device__ constexpr int a2_col(int c, int nc, int nc2) {
  const int nb = a2_buffers(nc), b = c % nb;
  if (nb == 1) return (nc2 >= 3 && c == nc2 - 1) ? kColD + kChunk * ((nc2 - 3) % kNumDBuf) : kColA2;
  return b == 0 ? kColA2 : kColA1 + 32 * nc + 32 * (b - 1);
}
constexpr int kD = 20, kA = 6;
constexpr int kMaxStages = 16;
constexpr int kSmemLimit = 232448;     // 227 KB
constexpr int kA0Bytes = 128 * 128;

// Event barriers.  Every one of them completes exactly ONCE per step, so every wait in the kernel is "parity = step & 1"
// and a waiter can never confuse two phases, however far the MMA issuers and the epilogue groups drift apart inside a
// step (they re-join at every step boundary through D3 / A0).  e = chunk index in epilogue order: layer-1 chunk c -> c,
// layer-2 chunk c -> NC + c.
enum : int {
  kEvA0Ready = 0,      // A0 rows of the step are in shared memory                     (epilogue -> L1 issuer)
  kEvD3Full = 1,       // output accumulator complete                                  (L3 issuer -> epilogue)
  kEvA1Ready = 2,      // [8]  hidden-1 k-block c is in TMEM                           (epilogue -> L2 issuers)
  kEvDFull = 10,       // [24] accumulator of chunk e complete                         (issuer -> epilogue)
  kEvDEmpty = 34,      // [24] accumulator of chunk e drained into registers           (epilogue -> issuer of the next user)
  kEvA2Ready = 58,     // [16] hidden-2 chunk c is in TMEM                             (epilogue -> L3 issuer)
  kEvA2Empty = 74,     // [16] layer 3 has consumed hidden-2 chunk c                   (L3 issuer -> epilogue)
  // proposal network (policy-guided shooting), also once per step each
  kEvAp0Ready = 90,    // policy input rows are in shared memory                        (epilogue -> issuer)
  kEvPdFull = 91,      // [8]  policy accumulators: layer-1 chunks 0..3, layer-2 chunks 0..3 (issuer -> epilogue)
  kEvAp1Ready = 99,    // [4]  policy hidden-1 chunk in TMEM                            (epilogue -> issuer)
  kEvAp2Ready = 103,   // [4]  policy hidden-2 chunk in TMEM
  kEvPolFull = 107,    // policy output accumulator complete                            (issuer -> epilogue)
  kNumEv = 108
};
constexpr int kMaxChunks1 = 8, kMaxChunks2 = 16;    // hidden-1 chunks; hidden-2 chunks (two heads of 8 in the learned-reward model)
// misc smem: [w_full 16][w_empty 16][res_full][events 108] x 8 B, tmem ptr, argmin scratch
constexpr int kOffWFull = 0, kOffWEmpty = 128, kOffResFull = 256, kOffEv = 264, kOffTmemPtr = 1136, kOffRed = 1152, kMiscBytes = 1280;
static_assert(kOffEv + 8 * kNumEv <= kOffTmemPtr && kOffRed + 80 <= kMiscBytes, "misc shared-memory map");
constexpr int kPolMaxHidden = 128;     // proposal network: two tanh layers of at most 128 units (MlpPolicy hid_size=128, BCnetwork 128/64) ...
constexpr int kPolVecFloats = 2 * 128 + 8 + 248;   // [bp1 (128)][bp2 (128)][bp3 (8)], padded to 2 KB
// ... or, in the kVarPolWide variants, of at most 256 units (MlpPolicy(hid_size=256), various_train/train_mpc_ppo_r.py:142):
// four 64-unit chunks per hidden layer, an 88 KB resident image per CTA, quarter-chunk weight stages for the dynamics net
constexpr int kPolWideHidden = 256;
constexpr int kPolWideVecFloats = 2 * 256 + 8 + 248;   // [bp1 (256)][bp2 (256)][bp3 (8)], padded to 3 KB

// ---------------------------------------------------------------------------------------------------------
// The per-step schedule is fixed arithmetic, the same for every step (all accumulators are drained at a step boundary):
//   layer 1 : chunk c accumulates in buffer c % 4, issued in pairs (c, c+1); pair n+2 waits for the drains of pair n
//   layer 2 : chunk c accumulates in buffer c % 3 (waits for the drain of chunk c-3), one MMA group per weight stage;
//             even chunks are issued by warp 1 from sub-ring 0, odd chunks by warp 3 from sub-ring 1
//   layer 3 : chunk c as soon as the epilogue has stored hidden-2 chunk c
// Chunk loops run at run time with compile-time inner unrolling: every tcgen05 operand is then a uniform-register
// value computed from the loop counter.  (A schedule TABLE walked at run time cost ~100 cycles per register-indexed
// constant load; a fully unrolled schedule made the kernel 160-200 KB of code and instruction-fetch bound.)
// ---------------------------------------------------------------------------------------------------------
// k-blocks per streamed weight stage: a whole chunk of W2 per stage for a CTA pair (32 KB per CTA at NC = 8), a
// quarter chunk for a single CTA (16 KB; the ring must hold several stages next to the resident W1 / W3)
// (a CTA pair that also holds a wide proposal network has ~70 KB left for the ring: quarter chunks there as well)
constexpr int stage_kblocks(int cg, int nc, bool pol_wide = false) { return (cg == 2 && !pol_wide) ? nc : (nc + 3) / 4; }

// shared-memory map, a function of (CTA-group size, hidden chunks) only: the kernel uses it as compile-time constants
struct Layout {
  int kbs = 0, nstg = 0, nst = 0;
  int nst_r[2] = {0, 0}, first_r[2] = {0, 0};     // the ring is split in two: even layer-2 chunks / odd layer-2 chunks, one per MMA issuer
  uint32_t blk_h = 0, blk_o = 0, w1_bytes = 0, res_bytes = 0, stage_bytes = 0;
  uint32_t pol_img_bytes = 0;
  uint32_t off_a0 = 0, off_res = 0, off_pol = 0, pol_bytes = 0, off_ln = 0, ln_bytes = 0, off_ring = 0, off_misc = 0, smem_bytes = 0;
};
// kernel variants (bit flags): policy-guided shooting, LayerNorm dynamics net, learned-reward model
// (NNDynamicsRewardModel, dynamics.py:121-238: shared layer -> state-delta head + reward head, objective = predicted reward)
// kVarBias: the biases of layers 2 / 3 are added in the epilogues instead of riding in the GEMMs through two constant-1
// hidden units - chosen when those two units would cost a whole extra 64-wide chunk (hidden sizes 63, 64, 127, 128, 255,
// 256: the reference's default size is 256, train_mpc_ppo.py:75).  The reward model always works that way (tanh).
// kVarTanh: a plain NNDynamicsModel with tanh hidden layers (the builder default of the reference, dynamics.py:60); always with
// kVarBias (a tanh does not preserve the constant-1 hidden units that carry the biases of layers 2 / 3 through a ReLU).
// kVarPolWide: the proposal network has up to 256 units per hidden layer (always with kVarPolicy; CTA pairs only).
enum : int { kVarUniform = 0, kVarPolicy = 1, kVarLayerNorm = 2, kVarReward = 4, kVarRewardPolicy = 5, kVarBias = 8, kVarTanh = 16, kVarPolWide = 32 };
// 64-unit chunks a hidden layer of the proposal network may have (= the image's chunk stride), its widest layer, its bias vector
__host__ __device__ constexpr int pol_chunks(int var) { return (var & kVarPolWide) ? 4 : 2; }
__host__ __device__ constexpr int pol_max_hidden(int var) { return (var & kVarPolWide) ? kPolWideHidden : kPolMaxHidden; }
__host__ __device__ constexpr int pol_vec_floats(int var) { return (var & kVarPolWide) ? kPolWideVecFloats : kPolVecFloats; }
__host__ __device__ constexpr int hidden2_chunks(int nc, int var) { return (var & kVarReward) ? 2 * nc : nc; }
constexpr int kRewardCol = kD;          // output column of the reward head in D3
constexpr uint32_t kLnStatsBytes = 3 * 2 * kRows * 8;      // [layer: hidden 1, hidden 2 (delta head), reward head][epilogue group][row] (sum, sum of squares)
constexpr Layout make_layout(int cg, int nc, int var) {
  const bool pol = (var & kVarPolicy) != 0, ln = (var & kVarLayerNorm) != 0, rew = (var & (kVarReward | kVarBias)) != 0;
  const int nc2 = hidden2_chunks(nc, var);
  Layout L;
  L.kbs = stage_kblocks(cg, nc, (var & kVarPolWide) != 0);
  L.nstg = (nc + L.kbs - 1) / L.kbs;
  L.blk_h = (kChunk / cg) * 128u;                  // bytes of one 64-wide k-block image of a hidden tile, per CTA
  L.blk_o = (kOutN / cg) * 128u;                   // ... of an output tile
  L.w1_bytes = nc * L.blk_h;
  L.res_bytes = L.w1_bytes + nc2 * L.blk_o;        // resident: W1 chunks then W3 k-blocks
  L.off_a0 = 0;
  L.off_res = kA0Bytes;
  L.off_pol = (L.off_res + L.res_bytes + 1023u) / 1024u * 1024u;
  // resident proposal network (sized for PN = 2 chunks = 128 hidden units per layer, or PN = 4 = 256 in the kVarPolWide
  // variants): [Wp1: PN chunks][Wp2: PN chunks x PN k-blocks][Wp3: PN k-blocks]
  const uint32_t pn = (uint32_t)pol_chunks(var);
  L.pol_img_bytes = pol ? (pn + pn * pn) * L.blk_h + pn * L.blk_o : 0u;
  L.pol_bytes = pol ? L.pol_img_bytes + pol_vec_floats(var) * 4u : 0u;    // + biases of the three layers (float32, from device memory)
  // LayerNorm variant: per-column correction vectors of layer 2 (u2, c2: 2 x 64*nc floats) + per-row statistics
  L.off_ln = (L.off_pol + L.pol_bytes + 1023u) / 1024u * 1024u;
  // LayerNorm: [u2 (64 nc2)][c2 (64 nc2)][u3 (32)][c3 (32)] correction vectors + per-row statistics
  L.ln_bytes = ln ? 2u * 64u * nc2 * 4u + 2u * kOutN * 4u + kLnStatsBytes : 0u;
  if (rew && !ln) L.ln_bytes = (64u * nc2 + kOutN) * 4u;   // learned-reward model: biases of both heads' hidden layers, then of the outputs
  L.off_ring = (L.off_ln + L.ln_bytes + 1023u) / 1024u * 1024u;
  L.stage_bytes = L.kbs * L.blk_h;
  L.nst = (int)((kSmemLimit - kMiscBytes - L.off_ring) / L.stage_bytes);
  if (L.nst > kMaxStages) L.nst = kMaxStages;
  L.nst_r[0] = (L.nst + 1) / 2; L.nst_r[1] = L.nst / 2;   // each sub-ring has ONE consumer, which sees every phase of its barriers
  L.first_r[0] = 0; L.first_r[1] = L.nst_r[0];
  L.off_misc = L.off_ring + L.nst * L.stage_bytes;
  L.smem_bytes = L.off_misc + kMiscBytes;
  return L;
}

template <int CG, int NC, int VAR>
struct Lay {
  static constexpr Layout value = make_layout(CG, NC, VAR);
};

struct TmArgs {
  PlanArgs p;
  const uint8_t* res;            // resident image   [CG][res_bytes]
  const uint8_t* stream;         // streamed image: per weight stage, in consumption order, [CG][stage_bytes]
  unsigned long long* trace;     // optional pipeline trace (tools/tc_trace.py), nullptr = off
  uint32_t hint_chain;           // try_wait suspend hint (ns) of the waits on the per-step dependency chain
  float mean_obs[kMaxObs], inv_den_obs[kMaxObs], mean_delta[kMaxObs], std_delta[kMaxObs];
  float mean_act[kMaxAct], inv_den_act[kMaxAct], low[kMaxAct], span[kMaxAct];
  // proposal network (policy-guided shooting); unused with uniform shooting
  const uint8_t* pol;            // resident image [CG][pol_bytes]
  int np1, np2;                  // 64-unit chunks of the two hidden layers
  int proposal, inj_kind, has_filter;
  int proposal0, inj_kind0;      // step 0 (the first stage of the tree-search planner, controllers.py:397-457; else = proposal / inj_kind)
  float clip, explore, one_minus_explore;
  float filt_mean[kMaxObs], filt_inv_std[kMaxObs], pol_std[kMaxAct];
  const float* polvec;           // [bp1 (128)][bp2 (128)][bp3 (8)] biases of the proposal network (device; staged into shared memory)
  // LayerNorm variant (dynamics.py:68-69): z = rho * acc - rho * mu * u + c per column, rho / mu per row
  const float* lnvec;            // [u2 (64*nc)][c2 (64*nc)] correction vectors of layer 2, [u3 (32)][c3 (32)] of the output layer
  float ln_inv_n, ln_eps;        // 1 / hidden_size, variance epsilon
  // learned-reward model: lnvec = [b2 of the delta head (64 * nc)][b2 of the reward head (64 * nc)][b3: 20 deltas, reward, 0...]
  // (its hidden activation is tanh, dynamics.py:162; other activations run in verification mode)
  float mean_reward, std_reward, gamma;
};
// pipeline trace of CTA (0,0), steps 0..3: [2 per L1 pair][2 per L2 chunk][2 per L3 chunk][4 per epilogue chunk][final 8][2 per stage]
constexpr int kTraceSteps = 4, kTrL1 = 0, kTrL2 = 8, kTrL3 = 24, kTrEpi = 40, kTrFinal = 104, kTrProd = 112, kTraceSlots = kTrProd + 2 * 64;

// ---------------------------------------------------------------------------------------------------------
// host: weight images
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return (uint16_t)((u >> 16) | 0x40);   // NaN
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
__host__ __device__ inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// byte offset of element (row n, k-column kk < 64) in a K-major SWIZZLE_128B image (8-row groups of 1024 B)
__host__ __device__ inline size_t sw128_offset(int n, int kk) {
  return (size_t)(n / 8) * 1024 + (size_t)(n % 8) * 128 + (size_t)(((kk / 8) ^ (n % 8)) * 16) + (size_t)(kk % 8) * 2;
}

struct TmHostPlan {
  size_t res_total = 0, stream_total = 0, lnvec_floats = 0;     // bytes / bytes / floats of the device images
  bool layer_norm = false, reward = false, explicit_bias = false;
  int cg = 1, s_pad = 0, nc = 0, nc2 = 0, hidden = 0;
  bool with_policy = false;
  int np1 = 0, np2 = 0;
  Layout lay;
};

inline bool tm_supported(const bcmpc_config& c, std::string* why) {
  if (c.obs_dim != kD || c.act_dim != kA) { *why = "bf16 fast path is specialised for obs_dim=20, act_dim=6 (HalfCheetah)"; return false; }
  if (c.n_hidden != 2) { *why = "rollout_tmem_kernel is built for n_hidden == 2"; return false; }
  if (c.hidden_size < 1 || c.hidden_size > 510) { *why = "rollout_tmem_kernel holds hidden_size <= 510 (TMEM keeps the 512-wide padded activations)"; return false; }
  return true;
}

// Proposal network for the fast path: sizes = {20, h1, h2, 6}, tanh hidden layers, h1, h2 <= 256 (more than 128 units in
// either layer selects the kVarPolWide kernel variants).  W[l] are [in,out].
inline bool tm_policy_supported(int n_layers, const int* sizes, std::string* why) {
  if (n_layers != 3) { *why = "bf16 fast path supports proposal networks with two hidden layers"; return false; }
  if (sizes[0] != kD || sizes[3] != kA) { *why = "bf16 fast path: proposal network must map 20 observations to 6 actions"; return false; }
  if (sizes[1] < 1 || sizes[1] > kPolWideHidden || sizes[2] < 1 || sizes[2] > kPolWideHidden) { *why = "bf16 fast path: proposal network hidden sizes must be <= 256"; return false; }
  return true;
}
inline bool tm_policy_is_wide(const int* sizes) { return sizes[1] > kPolMaxHidden || sizes[2] > kPolMaxHidden; }
// Dynamics weight images.  W[l] are TF kernels [in,out] (row stride ld): W1 [26,S], W2 [S,S], W3 [S,20].
// Bias-in-GEMM: input columns 26,27 of A0 are 1.0; hidden units S and S+1 of both hidden layers are wired to be
// exactly 1.0 after ReLU; row S / S+1 of the next layer's kernel hold bf16(b) and bf16(b - bf16(b)).
// ln_gamma / ln_beta: LayerNorm parameters of the two hidden layers, or nullptr.  With LayerNorm the kernel
// multiplies UN-normalised ReLU outputs r by W' = diag(gamma) W and corrects in the next epilogue:
//   LN(r) W + b = rho * (r W') - rho * mu * u + c,   u = gamma^T W (column sums of the bf16 W'),  c = beta^T W + b,
// rho = 1 / sqrt(var(r) + eps) and mu = mean(r) per path.  Biases of layers 2 / 3 live in c, so no "ones" units.
//
// The images are packed ON THE DEVICE from the float32 copies the handle keeps for the verification kernel, so a
// refresh after dyn_model.fit (train_mpc_ppo.py:279) is a copy plus two small kernels on the caller's stream whether
// the new weights come from the host (bcmpc_set_dynamics) or are already device-resident (bcmpc_set_dynamics_device).
inline void tm_geometry(int S, int cg, int var, TmHostPlan* out) {
  TmHostPlan& P = *out;
  const bool ln = (var & kVarLayerNorm) != 0, rew = (var & kVarReward) != 0, bias = (var & kVarBias) != 0;
  P.cg = cg;
  P.hidden = S;
  P.layer_norm = ln;
  P.reward = rew;
  P.explicit_bias = bias;
  P.s_pad = (((ln || rew || bias) ? S : S + 2) + 63) / 64 * 64;   // no hidden "ones" units with LayerNorm / the reward model / explicit biases
  P.nc = P.s_pad / 64;
  P.nc2 = hidden2_chunks(P.nc, var);
  P.lay = make_layout(cg, P.nc, var);
  P.res_total = (size_t)cg * P.lay.res_bytes;
  P.stream_total = (size_t)P.nc2 * P.lay.nstg * cg * P.lay.stage_bytes;
  P.lnvec_floats = ln ? (size_t)2 * 64 * P.nc2 + 2 * kOutN : 0;   // [u2][c2][u3][c3]
  if ((rew || bias) && !ln) P.lnvec_floats = (size_t)64 * P.nc2 + kOutN;    // [b2 (delta head)][b2 reward head][b3]
}

// ---------------------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------------------
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(f, std::make_integer_sequence<int, N>{});
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// ReLU + round-to-nearest bf16 of two accumulators in one instruction (negative -> +0, NaN stays NaN)
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// normalise (dynamics.py:109-110) -> bf16 -> the first four 16-byte chunks of this row of A0; columns 26, 27 = 1
__device__ __forceinline__ void build_input(const TmArgs& a, uint32_t a0_row_addr, int row, const float (&s)[kD],
                                            const float (&act)[kA]) {
  float x[32];
#pragma unroll
  for (int c = 0; c < kD; ++c) x[c] = (s[c] - a.mean_obs[c]) * a.inv_den_obs[c];   // subtract first: a zero-variance dim has inv_den = 1e10
#pragma unroll
  for (int j = 0; j < kA; ++j) x[kD + j] = (act[j] - a.mean_act[j]) * a.inv_den_act[j];
  x[kD + kA] = 1.f;
  x[kD + kA + 1] = 1.f;
#pragma unroll
  for (int c = kD + kA + 2; c < 32; ++c) x[c] = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t addr = a0_row_addr + (uint32_t)(((c ^ (row & 7)) << 4));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16(x[8 * c + 0], x[8 * c + 1])),
                 "r"(pack_bf16(x[8 * c + 2], x[8 * c + 3])), "r"(pack_bf16(x[8 * c + 4], x[8 * c + 5])),
                 "r"(pack_bf16(x[8 * c + 6], x[8 * c + 7]))
                 : "memory");
  }
}

// the injected randomness of (path, step): the given tensor or a Philox draw.  Uniform shooting: the action itself
// (controllers.py:53); policy-guided: N(0,1) / U[low,high) / U[-1,1) noise.
// DBG: the kernel flavour that accepts an injected tensor (and writes the debug outputs); POL: the noise kind depends on
// the proposal mode at run time - with uniform shooting it is always U[low,high) and the Box-Muller path is not compiled.
template <bool DBG, bool POL>
__device__ __forceinline__ void get_noise(const TmArgs& a, int b, int h, int k, float (&z)[kA]) {
  const PlanArgs& p = a.p;
  if (DBG && p.inj) {
    const double* src = p.inj + (((size_t)b * p.H + h) * p.K + k) * kA;
#pragma unroll
    for (int j = 0; j < kA; ++j) z[j] = (float)src[j];
  } else {
    float v[kMaxAct];
    draw_injection(v, kA, POL ? (h == 0 ? a.inj_kind0 : a.inj_kind) : (int)kInjUniformBox, plan_seed(p), (uint32_t)(p.path_offset + k), (uint32_t)b, (uint32_t)h, a.low, a.span);
#pragma unroll
    for (int j = 0; j < kA; ++j) z[j] = v[j];
  }
}

// observation filter of MlpPolicy (ppo_bc_policy.py:60) -> bf16 -> the first four 16-byte chunks of this row of the
// input tile (the proposal network reuses the A0 buffer; its biases are added in the epilogues)
__device__ __forceinline__ void build_policy_input(const TmArgs& a, uint32_t a0_row_addr, int row, const float (&s)[kD]) {
  float x[32];
#pragma unroll
  for (int c = 0; c < kD; ++c) {
    float v = s[c];
    if (a.has_filter) v = (v - a.filt_mean[c]) * a.filt_inv_std[c];
    if (a.clip > 0.f) v = fminf(fmaxf(v, -a.clip), a.clip);
    x[c] = v;
  }
#pragma unroll
  for (int c = kD; c < 32; ++c) x[c] = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t addr = a0_row_addr + (uint32_t)(((c ^ (row & 7)) << 4));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16(x[8 * c + 0], x[8 * c + 1])),
                 "r"(pack_bf16(x[8 * c + 2], x[8 * c + 3])), "r"(pack_bf16(x[8 * c + 4], x[8 * c + 5])),
                 "r"(pack_bf16(x[8 * c + 6], x[8 * c + 7]))
                 : "memory");
  }
}

// Barrier waits are executed by all 32 lanes of the role warp (measured: ~100 cycles when the phase is already
// complete; a single elected lane polling while the others park at the reconvergence barrier costs > 1000).
// DBG = true: the flavour used by tests / bcmpc_rollout_debug / bcmpc_predict / injected action tensors - it reads
// p.inj and writes states_out / actions_out / costs_out.  DBG = false: the production flavour (device-drawn
// randomness, records only); the same code with those paths compiled out, which is worth ~9 % of the step time
// (registers, instruction footprint).  BCMPC_TRACE (developer builds, tools/tc_trace.py) adds the pipeline trace.
__device__ __forceinline__ int p_tiles_per_state(const TmArgs& a) { return a.p.tiles_per_state; }

template <int CG, int NC, int VAR, bool DBG>
__global__ void __launch_bounds__(kThreads, 1) rollout_tmem_kernel(const __grid_constant__ TmArgs a) {
  constexpr bool POL = (VAR & kVarPolicy) != 0, LN = (VAR & kVarLayerNorm) != 0, REW = (VAR & kVarReward) != 0;
  constexpr bool BIAS = !LN && (REW || (VAR & kVarBias) != 0);   // biases of layers 2 / 3 added in the epilogues
  constexpr bool TANH = REW || (VAR & kVarTanh) != 0;            // hidden activation (the learned-reward model is tanh, dynamics.py:162)
  static_assert(!(VAR & kVarTanh) || ((BIAS || LN) && !REW), "the plain tanh variant uses explicit biases (or the LayerNorm correction vectors)");
  constexpr int NC2 = hidden2_chunks(NC, VAR);                // hidden-2 chunks (both heads of the learned-reward model)
  static_assert(NC <= kMaxChunks1 && NC2 <= kMaxChunks2, "unsupported kernel variant");
  static_assert(!(VAR & kVarPolWide) || (POL && CG == 2), "the wide proposal network exists for CTA pairs only");
  constexpr int PN = pol_chunks(VAR), PH = pol_max_hidden(VAR);   // proposal network: chunk slots per hidden layer, widest layer
  constexpr Layout L = Lay<CG, NC, VAR>::value;
  constexpr int kNstg = L.nstg;                               // weight stages per layer-2 chunk
  constexpr int kStagesPerStepR[2] = {((NC2 + 1) / 2) * kNstg, (NC2 / 2) * kNstg};   // weight stages per step and sub-ring (even / odd chunks)
  constexpr uint32_t kIdescH = ptx::umma_idesc_bf16(kRows * CG, kChunk), kIdescO = ptx::umma_idesc_bf16(kRows * CG, kOutN);
  static_assert(L.nst_r[0] >= 1 && L.nst_r[1] >= 1, "weight ring too small");

  extern __shared__ __align__(1024) uint8_t smem[];
  const PlanArgs& p = a.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  // Persistent CTAs: the grid is at most one wave (one CTA group per SM or SM pair); group `first` rolls the work units
  // first, first + stride, ... (a unit = the CG tiles of one root state that share an MMA) back to back.  TMEM, the
  // barriers, the resident weight images and the weight ring live across units - a unit boundary is just another step
  // boundary - so the ~7 us of per-CTA start-up are paid once per launch instead of once per tile.
  const int units_per_state = p_tiles_per_state(a) / CG;
  const int total_units = a.p.B * units_per_state;
  const int first_unit = (int)blockIdx.x / CG, unit_stride = (int)gridDim.x / CG;
  const int n_it = first_unit < total_units ? (total_units - first_unit + unit_stride - 1) / unit_stride : 0;
  const bool leader = rank == 0;
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t misc = base + L.off_misc;
  const uint32_t ev = misc + kOffEv;                                  // event barriers of THIS CTA
  const uint32_t ev_leader = CG == 2 ? ptx::mapa(ev, 0) : ev;        // ... of the leader CTA (shared::cluster address)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.off_misc + kOffTmemPtr);
  unsigned char* s_red = smem + L.off_misc + kOffRed;
  const int H = p.H;
  const int Htot = n_it * H;                 // steps this CTA group runs (all of its work units)

  if (tid == 0) {
    if (base & 1023u) { printf("bcmpc: dynamic smem base not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < kMaxStages; ++i) {
      // leader: its own producer's expect_tx arrive + the peer's "my half has landed" arrive
      ptx::mbar_init(misc + kOffWFull + 8u * i, (CG == 2 && leader) ? 2 : 1);
      ptx::mbar_init(misc + kOffWEmpty + 8u * i, 1);
    }
    ptx::mbar_init(misc + kOffResFull, (CG == 2 && leader) ? 2 : 1);
    ptx::mbar_init(ev + 8u * kEvA0Ready, 4 * CG);
    ptx::mbar_init(ev + 8u * kEvD3Full, 1);
    for (int i = 0; i < kMaxChunks1; ++i) ptx::mbar_init(ev + 8u * (kEvA1Ready + i), 4 * CG);
    for (int i = 0; i < kMaxChunks2; ++i) {
      ptx::mbar_init(ev + 8u * (kEvA2Ready + i), 4 * CG);
      ptx::mbar_init(ev + 8u * (kEvA2Empty + i), 1);
    }
    for (int i = 0; i < kMaxChunks1 + kMaxChunks2; ++i) { ptx::mbar_init(ev + 8u * (kEvDFull + i), 1); ptx::mbar_init(ev + 8u * (kEvDEmpty + i), 4 * CG); }
    ptx::mbar_init(ev + 8u * kEvAp0Ready, 4 * CG);
    for (int i = 0; i < 8; ++i) ptx::mbar_init(ev + 8u * (kEvPdFull + i), 1);
    for (int i = 0; i < 4; ++i) { ptx::mbar_init(ev + 8u * (kEvAp1Ready + i), 4 * CG); ptx::mbar_init(ev + 8u * (kEvAp2Ready + i), 4 * CG); }
    ptx::mbar_init(ev + 8u * kEvPolFull, 1);
    ptx::fence_mbar_init();
  }
  if constexpr (POL) {
    float* s_pv = reinterpret_cast<float*>(smem + L.off_pol + L.pol_img_bytes);
    for (int i = tid; i < 2 * PH + kMaxAct; i += kThreads) s_pv[i] = a.polvec[i];
  }
  if constexpr (BIAS) {            // [b2 (64 * NC2: both heads of the reward model)][b3 (32)]
    float* s_vec = reinterpret_cast<float*>(smem + L.off_ln);
    for (int i = tid; i < 64 * NC2 + kOutN; i += kThreads) s_vec[i] = a.lnvec[i];
  }
  if constexpr (LN) {
    float* s_vec = reinterpret_cast<float*>(smem + L.off_ln);
    for (int i = tid; i < 2 * 64 * NC2 + 2 * kOutN; i += kThreads) s_vec[i] = a.lnvec[i];
    // per-row statistics start at zero: a group without a chunk in some layer never writes its partial
    for (int i = tid; i < (int)(kLnStatsBytes / 4); i += kThreads) s_vec[2 * 64 * NC2 + 2 * kOutN + i] = 0.f;
  }
  if (warp == 2) ptx::tmem_alloc_cg<CG>(ptx::smem_u32(s_tmem), kTmemCols);
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  // The allocation is the whole TMEM (512 columns), so it starts at lane 0 / column 0; using the literal keeps every
  // TMEM address a compile-time value.
  if (*s_tmem != 0u) { if (tid == 0) printf("bcmpc: unexpected TMEM base 0x%x\n", *s_tmem); __trap(); }
#ifdef BCMPC_TRACE
  const bool tracing = !REW && a.trace != nullptr && blockIdx.x == 0;   // (the trace slots are laid out for NC2 == NC; first work unit only)
#else
  constexpr bool tracing = false;
#endif
  const uint32_t hc = a.hint_chain;

  // Warps 0-3 run their loops warp-uniformly and elect one lane only for the instructions themselves: every
  // descriptor / address is then a uniform-register value (an `if (lane == 0)` role costs a divergent-value waterfall
  // loop around every UTCHMMA).
  if (warp == 0) {
    // ===================== weight producer (every CTA loads its own half) =====================
    if (ptx::elect_one()) {
      const uint8_t* src = a.res + (size_t)rank * L.res_bytes;
      ptx::mbar_arrive_expect_tx(misc + kOffResFull, L.res_bytes + L.pol_img_bytes);
      for (uint32_t o = 0; o < L.res_bytes; o += 16384u)
        ptx::bulk_g2s(base + L.off_res + o, src + o, min(16384u, L.res_bytes - o), misc + kOffResFull);
      if constexpr (POL) {
        const uint8_t* psrc = a.pol + (size_t)rank * L.pol_img_bytes;
        for (uint32_t o = 0; o < L.pol_img_bytes; o += 16384u)
          ptx::bulk_g2s(base + L.off_pol + o, psrc + o, min(16384u, L.pol_img_bytes - o), misc + kOffResFull);
      }
    }
    __syncwarp();
    uint32_t stage[2] = {0, 0}, phase[2] = {0, 0};
    for (int h = 0; h < Htot; ++h) {
      const uint8_t* src = a.stream + (size_t)rank * L.stage_bytes;
      const bool trp = tracing && h < kTraceSteps;
      for (int c = 0; c < NC2; ++c) {
        const int r = c & 1;                     // sub-ring: even chunks feed issuer warp 1, odd chunks issuer warp 3
        for (int j = 0; j < kNstg; ++j) {
          const int i = c * kNstg + j;
          const uint32_t st = (uint32_t)L.first_r[r] + stage[r];
          if (trp && lane == 0) a.trace[h * kTraceSlots + kTrProd + 2 * i] = clock64();
          ptx::mbar_wait(misc + kOffWEmpty + 8u * st, phase[r] ^ 1u);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(misc + kOffWFull + 8u * st, L.stage_bytes);
#pragma unroll
            for (uint32_t o = 0; o < L.stage_bytes; o += 16384u)
              ptx::bulk_g2s(base + L.off_ring + st * L.stage_bytes + o, src + o, min(16384u, L.stage_bytes - o), misc + kOffWFull + 8u * st);
            if (trp) a.trace[h * kTraceSlots + kTrProd + 2 * i + 1] = clock64();
          }
          __syncwarp();
          src += (size_t)CG * L.stage_bytes;
          if (++stage[r] == (uint32_t)L.nst_r[r]) { stage[r] = 0; phase[r] ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    if (leader) {
      // ===================== MMA issuers =====================
      // warp 1: the even layer-2 chunks (sub-ring 0); warp 3: proposal network, layer 1, the odd layer-2 chunks (sub-ring 1)
      const int mine = warp == 1 ? 0 : 1;
      ptx::mbar_wait(misc + kOffResFull, 0);
      const uint64_t a0desc = ptx::umma_desc_sw128(base + L.off_a0);
      const uint64_t w1desc = ptx::umma_desc_sw128(base + L.off_res);
      const uint32_t nst_mine = (uint32_t)L.nst_r[mine], first_mine = (uint32_t)L.first_r[mine];
      const uint32_t stages_per_step = (uint32_t)kStagesPerStepR[mine];
      for (int h = 0; h < Htot; ++h) {
        const uint32_t uh = (uint32_t)h, par = uh & 1u;
        const bool trw = tracing && h < kTraceSteps;
        if (mine == 1) {
          if constexpr (POL) {
            // ---- proposal network: policy input (smem) -> tanh layer -> tanh layer -> action means (D3 columns 0..5) ----
            // Accumulators: layer-1 chunks in buffers 0,1 (2,3: wide network); layer-2 chunks in buffers 2,0 (1,3) (every policy
            // accumulator has been drained before the barrier that lets its buffer's next user start completes; buffer 3 =
            // [448,512) overlaps the output accumulator D3, which is written only after all of hidden 2 has been stored).
            const uint64_t wpdesc = ptx::umma_desc_sw128(base + L.off_pol);
            const uint32_t np1 = (uint32_t)a.np1, np2 = (uint32_t)a.np2;
            ptx::mbar_wait(ev + 8u * kEvAp0Ready, par, hc);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              for (uint32_t j = 0; j < np1; ++j) {
                const uint32_t d = PN == 2 ? (j ? (uint32_t)buf_col(1) : (uint32_t)buf_col(0)) : (j < 3u ? (uint32_t)kColD + (uint32_t)kChunk * j : (uint32_t)kColA2);
#pragma unroll
                for (uint32_t ks = 0; ks < 2; ++ks)
                  ptx::umma_ss<CG>(d, a0desc + 2u * ks, wpdesc + (uint64_t)(j * (L.blk_h >> 4) + 2u * ks), kIdescH, ks);
                ptx::umma_commit_cg<CG>(ev + 8u * (kEvPdFull + j));
              }
            }
            __syncwarp();
            if (np1 > 1) ptx::mbar_wait2(ev + 8u * kEvAp1Ready, ev + 8u * (kEvAp1Ready + 1), par, hc);
            else ptx::mbar_wait(ev + 8u * kEvAp1Ready, par, hc);
            if constexpr (PN > 2) {
              if (np1 > 3) ptx::mbar_wait2(ev + 8u * (kEvAp1Ready + 2), ev + 8u * (kEvAp1Ready + 3), par, hc);
              else if (np1 > 2) ptx::mbar_wait(ev + 8u * (kEvAp1Ready + 2), par, hc);
            }
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              for (uint32_t j = 0; j < np2; ++j) {
                // buffers 2, 0 (, 1, 3)
                const uint32_t d = PN == 2 ? (j ? (uint32_t)buf_col(0) : (uint32_t)buf_col(2)) : (j < 3u ? (uint32_t)kColD + (uint32_t)kChunk * ((j + 2u) % 3u) : (uint32_t)kColA2);
                for (uint32_t kb = 0; kb < np1; ++kb)
#pragma unroll
                  for (uint32_t ks = 0; ks < 4; ++ks)
                    ptx::umma_ts<CG>(d, kColA1 + kb * 32u + ks * 8u, wpdesc + (uint64_t)(((uint32_t)PN + j * (uint32_t)PN + kb) * (L.blk_h >> 4) + 2u * ks), kIdescH,
                                     (kb | ks) ? 1u : 0u);
                ptx::umma_commit_cg<CG>(ev + 8u * (kEvPdFull + PN + j));
              }
            }
            __syncwarp();
            if (np2 > 1) ptx::mbar_wait2(ev + 8u * kEvAp2Ready, ev + 8u * (kEvAp2Ready + 1), par, hc);
            else ptx::mbar_wait(ev + 8u * kEvAp2Ready, par, hc);
            if constexpr (PN > 2) {
              if (np2 > 3) ptx::mbar_wait2(ev + 8u * (kEvAp2Ready + 2), ev + 8u * (kEvAp2Ready + 3), par, hc);
              else if (np2 > 2) ptx::mbar_wait(ev + 8u * (kEvAp2Ready + 2), par, hc);
            }
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              for (uint32_t kb = 0; kb < np2; ++kb)
#pragma unroll
                for (uint32_t ks = 0; ks < 4; ++ks)
                  ptx::umma_ts<CG>(kColD3, kColA1 + 32u * (uint32_t)PN + kb * 32u + ks * 8u,
                                   wpdesc + (uint64_t)((uint32_t)(PN + PN * PN) * (L.blk_h >> 4) + kb * (L.blk_o >> 4) + 2u * ks), kIdescO, (kb | ks) ? 1u : 0u);
              ptx::umma_commit_cg<CG>(ev + 8u * kEvPolFull);
            }
            __syncwarp();
          }
        }
        // ---- layer 1: pairs of chunks over four accumulators.  Warp 3 issues pairs 0, 1 and the later even pairs,
        // warp 1 the later odd pairs: the second wave starts as soon as its own accumulators are drained. ----
        if (mine == 1) ptx::mbar_wait(ev + 8u * kEvA0Ready, par, hc);
        for (uint32_t c0 = 0; c0 < (uint32_t)NC; c0 += 2) {
          const uint32_t pr = c0 >> 1;
          const int owner = (pr >= 2 && (pr & 1u)) ? 0 : 1;
          if (owner != mine) continue;
          const bool two = c0 + 1 < (uint32_t)NC;
          if (c0 >= (uint32_t)kNumDBufL1) {      // the previous users of this pair of accumulators have been drained
            if (two) ptx::mbar_wait2(ev + 8u * (kEvDEmpty + c0 - kNumDBufL1), ev + 8u * (kEvDEmpty + c0 - kNumDBufL1 + 1), par, hc);
            else ptx::mbar_wait(ev + 8u * (kEvDEmpty + c0 - kNumDBufL1), par, hc);
          }
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            if (trw) a.trace[h * kTraceSlots + kTrL1 + c0] = clock64();
            constexpr int nks = (kD + kA + 2 + 15) / 16;
#pragma unroll
            for (uint32_t t = 0; t < 2; ++t) {
              if (t == 0 || two) {
                const uint32_t c = c0 + t, bsel = c % (uint32_t)kNumDBufL1;
                const uint32_t d = bsel < (uint32_t)kNumDBuf ? (uint32_t)kColD + (uint32_t)kChunk * bsel : (uint32_t)kColA2;
#pragma unroll
                for (int ks = 0; ks < nks; ++ks)
                  ptx::umma_ss<CG>(d, a0desc + 2u * ks, w1desc + (uint64_t)(c * (L.blk_h >> 4) + 2u * ks), kIdescH, ks > 0 ? 1u : 0u);
                ptx::umma_commit_cg<CG>(ev + 8u * (kEvDFull + c));
              }
            }
            if (trw) a.trace[h * kTraceSlots + kTrL1 + c0 + 1] = clock64();
          }
          __syncwarp();
        }
        // ---- layer 2: this issuer's chunks, one MMA group per weight stage ----
        for (uint32_t c = (uint32_t)mine; c < (uint32_t)NC2; c += 2) {
          const uint32_t gs0 = uh * stages_per_step + (c >> 1) * (uint32_t)kNstg;
          // the chunk's first weight stage landed long ago: take that wait before the ones on the dependency chain
          ptx::mbar_wait(misc + kOffWFull + 8u * (first_mine + gs0 % nst_mine), (gs0 / nst_mine) & 1u);
          if (c < 2) {
            // first layer-2 chunk of this issuer in the step: all of A1 must be in TMEM (which also means that every
            // layer-1 accumulator has been drained).  Each epilogue group signals its chunks in order, so the last two
            // barriers imply the rest.
            if constexpr (NC >= 2) ptx::mbar_wait2(ev + 8u * (kEvA1Ready + NC - 1), ev + 8u * (kEvA1Ready + NC - 2), par, hc);
            else ptx::mbar_wait(ev + 8u * (kEvA1Ready + NC - 1), par, hc);
          }
          if (c >= (uint32_t)kNumDBuf) ptx::mbar_wait(ev + 8u * (kEvDEmpty + NC + c - kNumDBuf), par, hc);   // previous user of the accumulator drained
          const uint32_t d = (uint32_t)kColD + (uint32_t)kChunk * (c % (uint32_t)kNumDBuf);
#pragma unroll
          for (int j = 0; j < kNstg; ++j) {
            const uint32_t gs = gs0 + (uint32_t)j, stage = first_mine + gs % nst_mine, phase = (gs / nst_mine) & 1u;
            if (j > 0) ptx::mbar_wait(misc + kOffWFull + 8u * stage, phase);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              if (trw && j == 0) a.trace[h * kTraceSlots + kTrL2 + 2 * c] = clock64();
              const uint64_t bdesc = ptx::umma_desc_sw128(base + L.off_ring + stage * L.stage_bytes);
              const int nblk = (NC - j * L.kbs < L.kbs) ? NC - j * L.kbs : L.kbs;
#pragma unroll
              for (int t = 0; t < L.kbs; ++t) {
                if (t < nblk) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    ptx::umma_ts<CG>(d, kColA1 + (j * L.kbs + t) * 32 + ks * 8, bdesc + (uint64_t)(t * (L.blk_h >> 4) + 2u * ks), kIdescH,
                                     (j == 0 && t == 0 && ks == 0) ? 0u : 1u);
                }
              }
              ptx::umma_commit_cg<CG>(misc + kOffWEmpty + 8u * stage);
              if (j == kNstg - 1) {
                ptx::umma_commit_cg<CG>(ev + 8u * (kEvDFull + NC + c));
                if (trw) a.trace[h * kTraceSlots + kTrL2 + 2 * c + 1] = clock64();
              }
            }
            __syncwarp();
          }
        }
      }
    } else if (warp == 1) {
      // ===================== peer CTA: tell the leader when this CTA's weight halves have landed =====================
      const uint32_t misc_leader = ptx::mapa(misc, 0);
      ptx::mbar_wait(misc + kOffResFull, 0);
      if (ptx::elect_one()) ptx::mbar_arrive_remote(misc_leader + kOffResFull);
      __syncwarp();
      uint32_t stage[2] = {0, 0}, phase[2] = {0, 0};
      for (int h = 0; h < Htot; ++h) {
        for (int c = 0; c < NC2; ++c) {
          const int r = c & 1;
          for (int j = 0; j < kNstg; ++j) {
            const uint32_t st = (uint32_t)L.first_r[r] + stage[r];
            ptx::mbar_wait(misc + kOffWFull + 8u * st, phase[r]);
            if (ptx::elect_one()) ptx::mbar_arrive_remote(misc_leader + kOffWFull + 8u * st);
            __syncwarp();
            if (++stage[r] == (uint32_t)L.nst_r[r]) { stage[r] = 0; phase[r] ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    if (leader) {
      // ===================== layer-3 MMA issuer: D3 (+)= A2 x W3[chunk] as the hidden-2 chunks appear =====================
      ptx::mbar_wait(misc + kOffResFull, 0);
      const uint64_t w3desc = ptx::umma_desc_sw128(base + L.off_res + L.w1_bytes);
      for (int h = 0; h < Htot; ++h) {
        const uint32_t uh = (uint32_t)h;
        const bool trw = tracing && h < kTraceSteps;
        for (uint32_t c = 0; c < (uint32_t)NC2; ++c) {
          ptx::mbar_wait(ev + 8u * (kEvA2Ready + c), uh & 1u, hc);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            if (trw) a.trace[h * kTraceSlots + kTrL3 + 2 * c] = clock64();
            uint32_t acol;
            if constexpr (a2_buffers(NC) == 1) {
              acol = c == (uint32_t)NC2 - 1 ? (uint32_t)a2_col(NC2 - 1, NC, NC2) : (uint32_t)kColA2;
            } else {
              const uint32_t bsel = c % (uint32_t)a2_buffers(NC);
              acol = bsel == 0 ? (uint32_t)kColA2 : (uint32_t)(kColA1 + 32 * NC) + 32u * (bsel - 1u);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_ts<CG>(kColD3, acol + ks * 8, w3desc + (uint64_t)(c * (L.blk_o >> 4) + 2u * ks), kIdescO, (c == 0 && ks == 0) ? 0u : 1u);
            if (c == (uint32_t)NC2 - 1) ptx::umma_commit_cg<CG>(ev + 8u * kEvD3Full);
            else ptx::umma_commit_cg<CG>(ev + 8u * (kEvA2Empty + c));
            if (trw) a.trace[h * kTraceSlots + kTrL3 + 2 * c + 1] = clock64();
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: group g (4 warps, one per lane quadrant) takes the chunks with (index & 1) == g =====================
    const int q = warp & 3, g = (warp - 4) >> 2, row = q * 32 + lane;
    const bool is_final = g == 0;
    const bool tr = tracing && q == 0 && lane == 0;
    for (int it = 0; it < n_it; ++it) {      // ---- one work unit: this CTA's tile of one root state, H steps ----
    const int unit = first_unit + it * unit_stride;
    const int b = unit / units_per_state, tile = (unit % units_per_state) * CG + (int)rank;
    const int k = tile * kRows + row;
    const bool valid = k < p.K;
    const int kc = valid ? k : p.K - 1;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t a0_row = base + L.off_a0 + (uint32_t)row * 128u;
    // hand-offs that concern TMEM contents / tensor-core reads: default (CTA-scope release) arrive, like CUTLASS' 2-SM kernels
    auto arrive_leader = [&](uint32_t id) {
      if constexpr (CG == 1) ptx::mbar_arrive(ev + 8u * id);
      else ptx::mbar_arrive_remote(ev_leader + 8u * id);
    };
    // A0 rows are generic-proxy shared-memory writes read by THIS CTA's tensor core (A is split by rows across the
    // pair); the proxy fence makes them visible to the async proxy, the arrive tells the leader's issuer to go
    auto publish = [&](uint32_t id) {
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive_leader(id);
    };
    float s[kD], z[kA], act[kA], act0[kA], cost = 0.f;
    [[maybe_unused]] float disc = 1.f;          // gamma ** h of the learned-reward objective
    if (is_final) {
#pragma unroll
      for (int d = 0; d < kD; ++d) s[d] = (float)root_state(p, (size_t)((DBG && p.per_row_states) ? kc : b), kD, d);
      get_noise<DBG, POL>(a, b, 0, kc, z);
      if constexpr (DBG) {
        if (p.states_out && valid)
          for (int d = 0; d < kD; ++d) p.states_out[(size_t)k * kD + d] = (double)s[d];
      }
    }
    for (int h = 0; h < H; ++h) {
      const uint32_t uh = (uint32_t)(it * H + h);           // running step of this CTA: the barriers' phase
      const bool trh = tr && it == 0 && h < kTraceSteps;
      // ---- the action of this step ----
      if constexpr (POL) {
        // proposal network (ppo_bc_policy.py:60,70-72 / behavioral_cloning.py:21-26) on the tensor cores
        if (is_final) {
          build_policy_input(a, a0_row, row, s);
          publish(kEvAp0Ready);
          if (trh) a.trace[h * kTraceSlots + kTrFinal + 2] = clock64();
        }
        // one 64-unit chunk of a tanh layer: accumulator + bias -> tanh -> bf16 -> TMEM operand of the next policy layer
        auto pol_chunk = [&](auto lc, auto jc) {
          constexpr int layer = decltype(lc)::value, j = decltype(jc)::value;
          constexpr uint32_t col = layer == 1 ? (uint32_t)buf_col(j) : (j < 3 ? (uint32_t)buf_col((j + 2) % 3) : (uint32_t)buf_col(3));
          constexpr uint32_t dst = kColA1 + (layer == 1 ? 0u : 32u * PN) + 32u * j;
          ptx::mbar_wait(ev + 8u * (kEvPdFull + (layer - 1) * PN + j), uh & 1u, hc);
          ptx::tc_fence_after();
          if (trh) a.trace[h * kTraceSlots + kTrFinal + 1 + 2 * layer] = clock64();
          uint32_t v0[32], v1[32], w0[16], w1[16];
          ptx::tmem_ld32(lane_off + col, v0);
          ptx::tmem_ld32(lane_off + col + 32u, v1);
          ptx::tmem_ld_wait();
          // (the bias vector starts 1024-byte aligned and 64 j floats are 256 bytes: 16-byte loads, 16 instead of 64 per chunk)
          const float4* b4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(smem + L.off_pol + L.pol_img_bytes) + (layer == 1 ? 0 : PH) + 64 * j);
          // bias, round to bf16, then one MUFU.TANH per two units (tanh.approx.bf16x2): the layer is MUFU-bound otherwise
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            w0[2 * i] = ptx::tanh_bf16x2(pack_bf16(__uint_as_float(v0[4 * i]) + bb.x, __uint_as_float(v0[4 * i + 1]) + bb.y));
            w0[2 * i + 1] = ptx::tanh_bf16x2(pack_bf16(__uint_as_float(v0[4 * i + 2]) + bb.z, __uint_as_float(v0[4 * i + 3]) + bb.w));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[8 + i];
            w1[2 * i] = ptx::tanh_bf16x2(pack_bf16(__uint_as_float(v1[4 * i]) + bb.x, __uint_as_float(v1[4 * i + 1]) + bb.y));
            w1[2 * i + 1] = ptx::tanh_bf16x2(pack_bf16(__uint_as_float(v1[4 * i + 2]) + bb.z, __uint_as_float(v1[4 * i + 3]) + bb.w));
          }
          ptx::tmem_st16(lane_off + dst, w0);
          ptx::tmem_st16(lane_off + dst + 16u, w1);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader((layer == 1 ? kEvAp1Ready : kEvAp2Ready) + j);
          if (trh) a.trace[h * kTraceSlots + kTrFinal + 2 + 2 * layer] = clock64();
        };
        // group g takes chunk g of both hidden layers
        if (g == 0) pol_chunk(std::integral_constant<int, 1>{}, std::integral_constant<int, 0>{});
        else if (a.np1 > 1) pol_chunk(std::integral_constant<int, 1>{}, std::integral_constant<int, 1>{});
        if constexpr (PN > 2) {      // wide network: group g also takes chunk 2 + g
          if (g == 0) { if (a.np1 > 2) pol_chunk(std::integral_constant<int, 1>{}, std::integral_constant<int, 2>{}); }
          else if (a.np1 > 3) pol_chunk(std::integral_constant<int, 1>{}, std::integral_constant<int, 3>{});
        }
        if (g == 0) pol_chunk(std::integral_constant<int, 2>{}, std::integral_constant<int, 0>{});
        else if (a.np2 > 1) pol_chunk(std::integral_constant<int, 2>{}, std::integral_constant<int, 1>{});
        if constexpr (PN > 2) {
          if (g == 0) { if (a.np2 > 2) pol_chunk(std::integral_constant<int, 2>{}, std::integral_constant<int, 2>{}); }
          else if (a.np2 > 3) pol_chunk(std::integral_constant<int, 2>{}, std::integral_constant<int, 3>{});
        }
        if (is_final) {
          ptx::mbar_wait(ev + 8u * kEvPolFull, uh & 1u, hc);
          ptx::tc_fence_after();
          if (trh) a.trace[h * kTraceSlots + kTrFinal + 7] = clock64();
          uint32_t v[32];
          ptx::tmem_ld32(lane_off + (uint32_t)kColD3, v);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
#pragma unroll
          for (int j = 0; j < kA; ++j) {
            const float mean = __uint_as_float(v[j]) + reinterpret_cast<const float*>(smem + L.off_pol + L.pol_img_bytes)[2 * PH + j];
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
          for (int j = 0; j < kA; ++j) act[j] = z[j];                        // controllers.py:70: the sampled action itself
        }
      }
      if (is_final) {
        if (h == 0) {
#pragma unroll
          for (int j = 0; j < kA; ++j) act0[j] = act[j];
        }
        if constexpr (DBG) {
          if (p.actions_out && valid)
            for (int j = 0; j < kA; ++j) p.actions_out[((size_t)h * p.K + k) * kA + j] = (double)act[j];
        }
        build_input(a, a0_row, row, s, act);
        publish(kEvA0Ready);
        if (trh) a.trace[h * kTraceSlots + kTrFinal + 1] = clock64();
      }
      // ---- LayerNorm variant: statistics of this row's ReLU outputs, correction of the previous layer's statistics ----
      // (each row's chunks are split between the two epilogue groups: partial sums meet in shared memory)
      float ln_sum = 0.f, ln_sq = 0.f, rho = 1.f, rmu = 0.f;
      [[maybe_unused]] float2* s_stats = reinterpret_cast<float2*>(smem + L.off_ln + (2u * 64u * NC2 + 2u * kOutN) * 4u);   // [layer][group][row]
      [[maybe_unused]] const float* s_u2 = reinterpret_cast<const float*>(smem + L.off_ln);
      [[maybe_unused]] const float* s_c2 = s_u2 + 64 * NC2;
      [[maybe_unused]] const float* s_u3 = s_u2 + 2 * 64 * NC2;
      [[maybe_unused]] const float* s_c3 = s_u3 + kOutN;
      // rho = 1 / sqrt(var + eps), rmu = rho * mean of the layer whose partial sums are in s_stats[layer]
      auto ln_row_stats = [&](int layer) {
        const float2 p0 = s_stats[(layer * 2 + 0) * kRows + row], p1 = s_stats[(layer * 2 + 1) * kRows + row];
        const float mean = (p0.x + p1.x) * a.ln_inv_n;
        const float var = fmaxf((p0.y + p1.y) * a.ln_inv_n - mean * mean, 0.f);
        rho = rsqrtf(var + a.ln_eps);
        rmu = rho * mean;
      };
      // one chunk with LayerNorm.  corr: apply the correction of the PREVIOUS layer's statistics (hidden 2; hidden 1 is
      // the plain activation of the accumulator); slot: whose statistics this chunk adds to (0 = hidden 1, 1 = hidden 2 /
      // delta head, 2 = reward head).  Activation: ReLU, or tanh for the learned-reward model (statistics of the
      // bf16-rounded values the tensor core will read).
      auto chunk_ln = [&](uint32_t e, uint32_t col, uint32_t dst, uint32_t ready_id, int wait_a2, bool corr, int slot, uint32_t c) {
        uint32_t v0[32], v1[32], w0[16], w1[16];
        ptx::tmem_ld32(lane_off + col, v0);
        ptx::tmem_ld32(lane_off + col + 32u, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(kEvDEmpty + e);
        const float4* u4 = reinterpret_cast<const float4*>(s_u2 + 64u * c);
        const float4* c4 = reinterpret_cast<const float4*>(s_c2 + 64u * c);
        auto pre = [&](float acc, float u, float cc) { return corr ? fmaf(rho, acc, fmaf(-rmu, u, cc)) : acc; };
        auto relu_stat = [&](float zv) {
          zv = zv < 0.f ? 0.f : zv;
          ln_sum += zv;
          ln_sq = fmaf(zv, zv, ln_sq);
          return zv;
        };
        // two neighbouring units -> activation -> bf16 pair (+ statistics)
        auto act_pair = [&](float x, float y) -> uint32_t {
          if constexpr (TANH) {
            const uint32_t t = ptx::tanh_bf16x2(pack_bf16(x, y));
            const float lo = __uint_as_float(t << 16), hi = __uint_as_float(t & 0xFFFF0000u);
            ln_sum += lo + hi;
            ln_sq = fmaf(lo, lo, fmaf(hi, hi, ln_sq));
            return t;
          } else {
            const float rx = relu_stat(x), ry = relu_stat(y);
            return pack_bf16(rx, ry);
          }
        };
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 uu = make_float4(0.f, 0.f, 0.f, 0.f), cc = uu;
          if (corr) { uu = u4[i]; cc = c4[i]; }
          w0[2 * i] = act_pair(pre(__uint_as_float(v0[4 * i]), uu.x, cc.x), pre(__uint_as_float(v0[4 * i + 1]), uu.y, cc.y));
          w0[2 * i + 1] = act_pair(pre(__uint_as_float(v0[4 * i + 2]), uu.z, cc.z), pre(__uint_as_float(v0[4 * i + 3]), uu.w, cc.w));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 uu = make_float4(0.f, 0.f, 0.f, 0.f), cc = uu;
          if (corr) { uu = u4[8 + i]; cc = c4[8 + i]; }
          w1[2 * i] = act_pair(pre(__uint_as_float(v1[4 * i]), uu.x, cc.x), pre(__uint_as_float(v1[4 * i + 1]), uu.y, cc.y));
          w1[2 * i + 1] = act_pair(pre(__uint_as_float(v1[4 * i + 2]), uu.z, cc.z), pre(__uint_as_float(v1[4 * i + 3]), uu.w, cc.w));
        }
        s_stats[(slot * 2 + g) * kRows + row] = make_float2(ln_sum, ln_sq);     // running partial of this group
        __threadfence_block();
        if (wait_a2 >= 0) {
          ptx::mbar_wait(ev + 8u * (kEvA2Empty + wait_a2), uh & 1u, hc);
          ptx::tc_fence_after();
        }
        ptx::tmem_st16(lane_off + dst, w0);
        ptx::tmem_st16(lane_off + dst + 16u, w1);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(ready_id);
      };
      // one accumulator chunk: TMEM -> registers, ReLU + bf16, -> TMEM operand of the next layer
      auto chunk = [&](uint32_t e, uint32_t col, uint32_t dst, uint32_t ready_id, int wait_a2) {
        if (trh) a.trace[h * kTraceSlots + kTrEpi + 4 * e] = clock64();
        uint32_t v0[32], v1[32], w0[16], w1[16];
        ptx::tmem_ld32(lane_off + col, v0);
        ptx::tmem_ld32(lane_off + col + 32u, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(kEvDEmpty + e);             // accumulator is in registers: release the TMEM buffer
        if (trh) a.trace[h * kTraceSlots + kTrEpi + 4 * e + 1] = clock64();
#pragma unroll
        for (int i = 0; i < 16; ++i) w0[i] = pack_relu_bf16(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
#pragma unroll
        for (int i = 0; i < 16; ++i) w1[i] = pack_relu_bf16(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
        if (wait_a2 >= 0) {
          ptx::mbar_wait(ev + 8u * (kEvA2Empty + wait_a2), uh & 1u, hc);   // layer 3 has consumed the buffer's previous chunk
          ptx::tc_fence_after();
        }
        if (trh) a.trace[h * kTraceSlots + kTrEpi + 4 * e + 2] = clock64();
        ptx::tmem_st16(lane_off + dst, w0);
        ptx::tmem_st16(lane_off + dst + 16u, w1);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(ready_id);
        if (trh) a.trace[h * kTraceSlots + kTrEpi + 4 * e + 3] = clock64();
      };
      // learned-reward model: the hidden "ones" units that carry the biases of layers 2 / 3 through a ReLU do not
      // survive a tanh, so those biases are added here (layer 1 keeps its bias in the GEMM: its ones are INPUT columns);
      // activation = tanh (tanh.approx.bf16x2 on the rounded pair, as in the proposal network)
      [[maybe_unused]] const float* s_b2 = reinterpret_cast<const float*>(smem + L.off_ln);
      [[maybe_unused]] auto chunk_bias = [&](uint32_t e, uint32_t col, uint32_t dst, uint32_t ready_id, int wait_a2, const float* bias) {
        uint32_t v0[32], v1[32], w0[16], w1[16];
        ptx::tmem_ld32(lane_off + col, v0);
        ptx::tmem_ld32(lane_off + col + 32u, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(kEvDEmpty + e);
        const float4* b4 = reinterpret_cast<const float4*>(bias);
        auto act2 = [&](float x, float y) {
          if constexpr (TANH) return ptx::tanh_bf16x2(pack_bf16(x, y));
          else return pack_relu_bf16(x, y);
        };
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) bb = b4[i];
          w0[2 * i] = act2(__uint_as_float(v0[4 * i]) + bb.x, __uint_as_float(v0[4 * i + 1]) + bb.y);
          w0[2 * i + 1] = act2(__uint_as_float(v0[4 * i + 2]) + bb.z, __uint_as_float(v0[4 * i + 3]) + bb.w);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) bb = b4[8 + i];
          w1[2 * i] = act2(__uint_as_float(v1[4 * i]) + bb.x, __uint_as_float(v1[4 * i + 1]) + bb.y);
          w1[2 * i + 1] = act2(__uint_as_float(v1[4 * i + 2]) + bb.z, __uint_as_float(v1[4 * i + 3]) + bb.w);
        }
        if (wait_a2 >= 0) {
          ptx::mbar_wait(ev + 8u * (kEvA2Empty + wait_a2), uh & 1u, hc);
          ptx::tc_fence_after();
        }
        ptx::tmem_st16(lane_off + dst, w0);
        ptx::tmem_st16(lane_off + dst + 16u, w1);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(ready_id);
      };
      // layer 1: the groups take alternate PAIRS of chunks (the pairs the issuer commits together): one combined wait
      for (uint32_t c0 = 2u * (uint32_t)g; c0 < (uint32_t)NC; c0 += 4) {
        const bool two = c0 + 1 < (uint32_t)NC;
        if (two) ptx::mbar_wait2(ev + 8u * (kEvDFull + c0), ev + 8u * (kEvDFull + c0 + 1), uh & 1u, hc);
        else ptx::mbar_wait(ev + 8u * (kEvDFull + c0), uh & 1u, hc);
        ptx::tc_fence_after();
#pragma unroll 1
        for (uint32_t c = c0; c < c0 + (two ? 2u : 1u); ++c) {
          const uint32_t bsel = c % (uint32_t)kNumDBufL1;
          const uint32_t col = bsel < (uint32_t)kNumDBuf ? (uint32_t)kColD + (uint32_t)kChunk * bsel : (uint32_t)kColA2;
          if constexpr (LN) chunk_ln(c, col, (uint32_t)kColA1 + 32u * c, (uint32_t)kEvA1Ready + c, -1, false, 0, c);
          else if constexpr (BIAS) chunk_bias(c, col, (uint32_t)kColA1 + 32u * c, (uint32_t)kEvA1Ready + c, -1, nullptr);
          else chunk(c, col, (uint32_t)kColA1 + 32u * c, (uint32_t)kEvA1Ready + c, -1);
        }
      }
      // the next step's randomness is drawn between the layer-1 and layer-2 chunks, where these warps have slack (at the
      // top of the step it delays the first layer-1 epilogue, at the bottom the state update)
      if (is_final && h + 1 < H) get_noise<DBG, POL>(a, b, h + 1, kc, z);
      // layer 2: alternate chunks
      bool first_l2 = true;
      [[maybe_unused]] bool in_reward_head = false;
#pragma unroll 1
      for (uint32_t c = (uint32_t)g; c < (uint32_t)NC2; c += 2) {
        ptx::mbar_wait(ev + 8u * (kEvDFull + NC + c), uh & 1u, hc);
        ptx::tc_fence_after();
        if constexpr (LN) {
          if (first_l2) {      // all hidden-1 chunks of both groups are stored (layer 2 has run): their statistics are complete
            ln_row_stats(0);
            ln_sum = 0.f; ln_sq = 0.f;
            first_l2 = false;
          }
        }
        uint32_t dst;
        int wait_a2;
        if constexpr (a2_buffers(NC) == 1) {
          const bool last = NC2 >= 3 && c == (uint32_t)NC2 - 1;
          dst = last ? (uint32_t)a2_col(NC2 - 1, NC, NC2) : (uint32_t)kColA2;
          wait_a2 = (c == 0 || last) ? -1 : (int)c - 1;
        } else {
          const uint32_t bsel = c % (uint32_t)a2_buffers(NC);
          dst = bsel == 0 ? (uint32_t)kColA2 : (uint32_t)(kColA1 + 32 * NC) + 32u * (bsel - 1u);
          wait_a2 = c >= (uint32_t)a2_buffers(NC) ? (int)c - a2_buffers(NC) : -1;
        }
        if constexpr (LN && REW) {
          if (c >= (uint32_t)NC && !in_reward_head) { ln_sum = 0.f; ln_sq = 0.f; in_reward_head = true; }   // the reward head keeps its own statistics
        }
        if constexpr (LN) chunk_ln((uint32_t)NC + c, (uint32_t)kColD + (uint32_t)kChunk * (c % (uint32_t)kNumDBuf), dst, (uint32_t)kEvA2Ready + c, wait_a2, true,
                                   (REW && c >= (uint32_t)NC) ? 2 : 1, c);
        else if constexpr (BIAS) chunk_bias((uint32_t)NC + c, (uint32_t)kColD + (uint32_t)kChunk * (c % (uint32_t)kNumDBuf), dst, (uint32_t)kEvA2Ready + c, wait_a2, s_b2 + 64u * c);
        else chunk((uint32_t)NC + c, (uint32_t)kColD + (uint32_t)kChunk * (c % (uint32_t)kNumDBuf), dst, (uint32_t)kEvA2Ready + c, wait_a2);
      }

      if (is_final) {
        ptx::mbar_wait(ev + 8u * kEvD3Full, uh & 1u, hc);
        ptx::tc_fence_after();
        if (trh) a.trace[h * kTraceSlots + kTrFinal] = clock64();
        uint32_t v[32];
        ptx::tmem_ld32(lane_off + (uint32_t)kColD3, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        float delta[kD];
        if constexpr (LN) {
          ln_row_stats(1);
#pragma unroll
          for (int d = 0; d < kD; ++d) v[d] = __float_as_uint(fmaf(rho, __uint_as_float(v[d]), fmaf(-rmu, s_u3[d], s_c3[d])));
          if constexpr (REW) {
            ln_row_stats(2);
            v[kRewardCol] = __float_as_uint(fmaf(rho, __uint_as_float(v[kRewardCol]), fmaf(-rmu, s_u3[kRewardCol], s_c3[kRewardCol])));
          }
          ln_sum = 0.f; ln_sq = 0.f;
        }
        if constexpr (BIAS) {
          const float* s_b3 = s_b2 + 64 * NC2;
#pragma unroll
          for (int d = 0; d <= kRewardCol; ++d) v[d] = __float_as_uint(__uint_as_float(v[d]) + s_b3[d]);
        }
#pragma unroll
        for (int d = 0; d < kD; ++d) delta[d] = fmaf(__uint_as_float(v[d]), a.std_delta[d], a.mean_delta[d]);   // dynamics.py:113
        if constexpr (REW) {
          // controllers.py:137-139 / :340-345: reward = y * std_reward + mean_reward (dynamics.py:236), times gamma ** h;
          // the planner minimises, so the running "cost" is minus the discounted reward sum
          cost -= fmaf(__uint_as_float(v[kRewardCol]), a.std_reward, a.mean_reward) * disc;
          disc *= a.gamma;
        } else {
          if (s[5] >= 0.2f) cost += 10.f;                                    // cost_functions.py:16-26 (pre-step state)
          if (s[6] >= 0.f) cost += 10.f;
          if (s[7] >= 0.f) cost += 10.f;
          cost -= delta[17] * 100.f;                                         // cost_functions.py:28 (/ 0.01)
        }
#pragma unroll
        for (int d = 0; d < kD; ++d) s[d] += delta[d];                       // dynamics.py:116
        if constexpr (DBG) {
          if (p.states_out && valid)
            for (int d = 0; d < kD; ++d) p.states_out[((size_t)(h + 1) * p.K + k) * kD + d] = (double)s[d];
        }
      }
    }
    if (is_final) {
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
      if (p.out && (src >> 5) == q) fused_state_minloc(p, b, lane, src & 31);   // (src is the same in all 128 threads)
    }
    }      // work units
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc_cg<CG>(0u, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------
// launch: picks the instantiation for (CTA-group size, hidden chunks, variant, flavour)
// ---------------------------------------------------------------------------------------------------------
template <int CG, int NC, int VAR, bool DBG>
inline cudaError_t tm_launch_t(const TmArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = rollout_tmem_kernel<CG, NC, VAR, DBG>;
  constexpr Layout L = Lay<CG, NC, VAR>::value;
  static bool attr_done[64] = {};   // per instantiation and device (a process may own planners on several GPUs)
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
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a);
}
// CTA pairs (the default): every variant at every width (1..8 hidden chunks), except the explicit-bias ReLU variants, which
// exist only where they save a chunk (hidden 63/64, 127/128, 191/192, 255/256).  Single CTAs (BCMPC_TC_CG=1, a developer
// mode): variants 0..2 everywhere, the combinations at 4 and 8 chunks, no plain-tanh variant.
inline bool tm_width_supported(int cg, int nc, int var) {
  if (nc < 1 || nc > 8) return false;
  if (var & kVarPolWide) {      // wide proposal network: CTA pairs; ReLU (biases in the GEMMs), LayerNorm, the reward model, plain tanh
    const int base = var & ~kVarPolWide;
    return cg == 2 && (base == 1 || base == 3 || base == 5 || base == 7 || base == (kVarPolicy | kVarBias | kVarTanh));
  }
  if (var & kVarTanh) return cg == 2 && (var & ~(kVarPolicy | kVarLayerNorm)) == (kVarBias | kVarTanh);      // plain tanh net, with or without LayerNorm
  if (var & kVarBias) return var <= (kVarBias | kVarPolicy) && nc <= 4;
  if (cg == 2) return var >= 0 && var < 8;
  if (var == (kVarLayerNorm | kVarPolicy)) return nc <= 4 || nc == 8;
  return var < 3 || nc == 4 || nc == 8;
}
template <int CG, int NC, int VAR, bool DBG>
inline cudaError_t tm_launch_fit(const TmArgs& a, dim3 grid, cudaStream_t st) {
  // (a CTA that cannot hold the resident images next to a two-stage weight ring has no instantiation: tm_prepare reports it)
  if constexpr (Lay<CG, NC, VAR>::value.nst >= 2) return tm_launch_t<CG, NC, VAR, DBG>(a, grid, st);
  return cudaErrorInvalidValue;
}
template <int CG, int VAR, bool DBG>
inline cudaError_t tm_launch_cg(int nc, const TmArgs& a, dim3 grid, cudaStream_t st) {
  if (!tm_width_supported(CG, nc, VAR)) return cudaErrorInvalidValue;
  constexpr bool all = CG == 2 || VAR < 3;              // every width is instantiated
  constexpr bool narrow = (VAR & kVarBias) && !(VAR & kVarTanh);   // only 1..4 chunks
  if (nc == 1) { if constexpr (all || narrow || VAR == 3) return tm_launch_fit<CG, 1, VAR, DBG>(a, grid, st); }
  if (nc == 2) { if constexpr (all || narrow || VAR == 3) return tm_launch_fit<CG, 2, VAR, DBG>(a, grid, st); }
  if (nc == 3) { if constexpr (all || narrow || VAR == 3) return tm_launch_fit<CG, 3, VAR, DBG>(a, grid, st); }
  if (nc == 4) return tm_launch_fit<CG, 4, VAR, DBG>(a, grid, st);
  if constexpr (!narrow) {
    if (nc == 5) { if constexpr (all) return tm_launch_fit<CG, 5, VAR, DBG>(a, grid, st); }
    if (nc == 6) { if constexpr (all) return tm_launch_fit<CG, 6, VAR, DBG>(a, grid, st); }
    if (nc == 7) { if constexpr (all) return tm_launch_fit<CG, 7, VAR, DBG>(a, grid, st); }
    if (nc == 8) return tm_launch_fit<CG, 8, VAR, DBG>(a, grid, st);
  }
  return cudaErrorInvalidValue;
}

// One translation unit per (variant, flavour) - csrc/tm_var*.cu - so that the ~150 kernel instantiations compile in parallel.
template <int VAR, bool DBG>
cudaError_t tm_launch_var(int cg, int nc, const TmArgs& a, dim3 grid, cudaStream_t st);
#define BCMPC_TM_DEFINE_VARIANT(VAR, DBG)                                                                        \
  template <>                                                                                                   \
  cudaError_t tm_launch_var<VAR, DBG>(int cg, int nc, const TmArgs& a, dim3 grid, cudaStream_t st) {            \
    if (cg == 2) return tm_launch_cg<2, VAR, DBG>(nc, a, grid, st);                                             \
    if constexpr (((VAR) & (kVarTanh | kVarPolWide)) == 0) return tm_launch_cg<1, VAR, DBG>(nc, a, grid, st);   \
    return cudaErrorInvalidValue;                                                                               \
  }
#define BCMPC_TM_DECLARE_VARIANT(VAR)                                                                            \
  template <> cudaError_t tm_launch_var<VAR, false>(int cg, int nc, const TmArgs& a, dim3 grid, cudaStream_t st); \
  template <> cudaError_t tm_launch_var<VAR, true>(int cg, int nc, const TmArgs& a, dim3 grid, cudaStream_t st);
BCMPC_TM_DECLARE_VARIANT(0) BCMPC_TM_DECLARE_VARIANT(1) BCMPC_TM_DECLARE_VARIANT(2) BCMPC_TM_DECLARE_VARIANT(3)
BCMPC_TM_DECLARE_VARIANT(4) BCMPC_TM_DECLARE_VARIANT(5) BCMPC_TM_DECLARE_VARIANT(6) BCMPC_TM_DECLARE_VARIANT(7)
BCMPC_TM_DECLARE_VARIANT(8) BCMPC_TM_DECLARE_VARIANT(9) BCMPC_TM_DECLARE_VARIANT(24) BCMPC_TM_DECLARE_VARIANT(25)
BCMPC_TM_DECLARE_VARIANT(26) BCMPC_TM_DECLARE_VARIANT(27)
BCMPC_TM_DECLARE_VARIANT(33) BCMPC_TM_DECLARE_VARIANT(35) BCMPC_TM_DECLARE_VARIANT(37) BCMPC_TM_DECLARE_VARIANT(39) BCMPC_TM_DECLARE_VARIANT(57)
template <bool DBG>
inline cudaError_t tm_launch_flavour(int cg, int nc, int var, const TmArgs& a, dim3 grid, cudaStream_t st) {
  switch (var) {
    case 0: return tm_launch_var<0, DBG>(cg, nc, a, grid, st);   // uniform shooting
    case 1: return tm_launch_var<1, DBG>(cg, nc, a, grid, st);   // proposal network
    case 2: return tm_launch_var<2, DBG>(cg, nc, a, grid, st);   // LayerNorm dynamics
    case 3: return tm_launch_var<3, DBG>(cg, nc, a, grid, st);   // LayerNorm + proposal network
    case 4: return tm_launch_var<4, DBG>(cg, nc, a, grid, st);   // learned-reward model
    case 5: return tm_launch_var<5, DBG>(cg, nc, a, grid, st);   // ... + proposal network
    case 6: return tm_launch_var<6, DBG>(cg, nc, a, grid, st);   // learned-reward model with LayerNorm
    case 7: return tm_launch_var<7, DBG>(cg, nc, a, grid, st);   // ... + proposal network (what run.sh launches)
    case 8: return tm_launch_var<8, DBG>(cg, nc, a, grid, st);   // uniform shooting, explicit biases (hidden 64 / 128 / 256)
    case 9: return tm_launch_var<9, DBG>(cg, nc, a, grid, st);   // ... + proposal network
    case 24: return tm_launch_var<24, DBG>(cg, nc, a, grid, st); // plain tanh dynamics (explicit biases)
    case 25: return tm_launch_var<25, DBG>(cg, nc, a, grid, st); // ... + proposal network
    case 26: return tm_launch_var<26, DBG>(cg, nc, a, grid, st); // plain tanh dynamics with LayerNorm (the builder default activation + FLAGS.LAYER_NORM)
    case 27: return tm_launch_var<27, DBG>(cg, nc, a, grid, st); // ... + proposal network
    // proposal networks of up to 256 units per hidden layer (kVarPolWide)
    case 33: return tm_launch_var<33, DBG>(cg, nc, a, grid, st); // ReLU dynamics
    case 35: return tm_launch_var<35, DBG>(cg, nc, a, grid, st); // LayerNorm dynamics
    case 37: return tm_launch_var<37, DBG>(cg, nc, a, grid, st); // learned-reward model (train_mpc_ppo_r.py:142 runs this pair)
    case 39: return tm_launch_var<39, DBG>(cg, nc, a, grid, st); // learned-reward model with LayerNorm
    case 57: return tm_launch_var<57, DBG>(cg, nc, a, grid, st); // plain tanh dynamics
  }
  return cudaErrorInvalidValue;
}
// debug flavour iff the call injects randomness, starts every row from its own state or asks for any per-path output
inline bool tm_needs_debug_flavour(const PlanArgs& p) {
  return p.inj != nullptr || p.states_out != nullptr || p.actions_out != nullptr || p.costs_out != nullptr || p.per_row_states != 0;
}
inline cudaError_t tm_launch(int cg, int nc, int var, const TmArgs& a, dim3 grid, cudaStream_t st) {
  return tm_needs_debug_flavour(a.p) ? tm_launch_flavour<true>(cg, nc, var, a, grid, st) : tm_launch_flavour<false>(cg, nc, var, a, grid, st);
}

}  // namespace tmx
}  // namespace bcmpc
