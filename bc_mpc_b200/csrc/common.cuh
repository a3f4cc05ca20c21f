// Shared device-side definitions of the planner kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/bcmpc.h"

namespace bcmpc {

constexpr int kMaxLayers = BCMPC_MAX_LAYERS;
constexpr int kMaxObs = BCMPC_MAX_OBS;
constexpr int kMaxAct = BCMPC_MAX_ACT;

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter = (global path index, root-state index, step,
// block), key = seed: the sampled action set is identical for any sharding of the paths.
// Replaces numpy's MT19937 draw of controllers.py:53 (bit parity with MT19937 is not a goal;
// parity runs inject the tensor, sampling is validated statistically).
// ---------------------------------------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c0), hi1 = __umulhi(M1, c2);
#else
    uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c0) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c2) >> 32);
#endif
    uint32_t lo0 = M0 * c0, lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }   // [0,1)
__device__ __forceinline__ float u01_open(uint32_t x) { return ((float)(x >> 8) + 1.0f) * 5.9604644775390625e-8f; }  // (0,1]

enum InjKind { kInjUniformBox = 0, kInjNormal = 1, kInjUniformPm1 = 2 };

// The injected randomness of one (path, step): `n_act` float32 draws (n_act <= 8).
// kInjUniformBox: U[low,high)   (P0 actions, P2 exploration)   controllers.py:53
// kInjNormal    : N(0,1)        (P1, tf.random_normal)         distributions.py:80
// kInjUniformPm1: U[-1,1)       (P3, np.random.rand*2-1)       controllers.pyc src 120-163
__device__ __forceinline__ void draw_injection(float* out, int n_act, int kind, uint64_t seed, uint32_t gpath,
                                               uint32_t root, uint32_t step, const float* low, const float* span) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t r[8];
  Philox4 p = philox4x32_10(gpath, root, step, 0u, k0, k1);
  r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = p.w;
  if (n_act > 4) {
    p = philox4x32_10(gpath, root, step, 1u, k0, k1);
    r[4] = p.x; r[5] = p.y; r[6] = p.z; r[7] = p.w;
  } else {
    r[4] = r[5] = r[6] = r[7] = 0u;
  }
  if (kind == kInjNormal) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // Box-Muller pairs
      float rad = sqrtf(-2.0f * logf(u01_open(r[2 * j])));
      float s, c;
      sincospif(2.0f * u01(r[2 * j + 1]), &s, &c);
      if (2 * j < n_act) out[2 * j] = rad * c;
      if (2 * j + 1 < n_act) out[2 * j + 1] = rad * s;
    }
  } else if (kind == kInjUniformBox) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < n_act) out[j] = fmaf(u01(r[j]), span[j], low[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < n_act) out[j] = fmaf(u01(r[j]), 2.0f, -1.0f);
  }
}

// ---------------------------------------------------------------------------------------------
// np.argmin order (controllers.py:82): lowest index on ties; the first NaN wins.
// ---------------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ __forceinline__ bool cand_better(T ca, int64_t ia, T cb, int64_t ib) {
  const bool na = ca != ca, nb = cb != cb;
  if (na || nb) {
    if (na && nb) return ia < ib;
    return na;
  }
  if (ca < cb) return true;
  if (ca > cb) return false;
  return ia < ib;
}

// Per-tile candidate written by the rollout kernels, reduced per root state afterwards.
using Candidate = bcmpc_record;   // {double cost; int64 index (-1 = empty); double action[8]}

// Everything a rollout kernel needs that is not kernel-specific.
struct PlanArgs {
  int D, A, K, H, B;
  int proposal, inj_kind;
  int proposal0, inj_kind0;   // step 0 (differs from proposal / inj_kind in the tree-search planner only)
  int64_t path_offset;
  uint64_t seed;
  const unsigned long long* seed_ptr;   // device-resident seed (bcmpc_plan_device_seeded: a captured CUDA graph replays with new draws), or nullptr = `seed`
  const double* states;       // [B, D] root states  (or [n, D] per-row states in predict mode)
  const double* inj;          // [B, H, K, A] or nullptr
  Candidate* cand;            // [B, tiles_per_state]
  int tiles_per_state;
  // debug / verification outputs (single root state): may be nullptr
  double* states_out;         // [H+1, K, D]
  double* actions_out;        // [H, K, A]
  double* costs_out;          // [B, K]
  int per_row_states;         // predict mode: row k starts from states[k] instead of states[b]
  // fused per-state min-loc (fast kernels, few tiles per state): the last tile of a root state to finish reduces the
  // state's candidates itself instead of a second kernel launch
  bcmpc_record* out;          // [B] winning records, or nullptr = a separate reduce_candidates_kernel follows
  unsigned int* ticket;       // [B] finished-tile counters, zero between plans
  // single host-buffer plans (the reference's env loop, utils.py:200): the root state rides in the launch arguments - no
  // host-to-device copy ahead of the kernel (one API call and one DMA round trip less per plan; `states` is not read then)
  int inline_state;
  double state0[kMaxObs];
};

// entry d of the root state of row `row` ([B, D] states, or the state in the launch arguments)
__device__ __forceinline__ double root_state(const PlanArgs& p, size_t row, int D, int d) {
  return p.inline_state ? p.state0[d] : p.states[row * (size_t)D + d];
}

// the plan's Philox key: the launch argument, or the device scalar a captured graph reads at replay time
__device__ __forceinline__ uint64_t plan_seed(const PlanArgs& p) { return p.seed_ptr ? (uint64_t)__ldg(p.seed_ptr) : p.seed; }

// Called by every lane of the warp that holds a tile's winning row, after that row's lane (writer_lane) has written the
// tile's candidate.  np.argmin order, same as reduce_candidates_kernel.
__device__ __forceinline__ void fused_state_minloc(const PlanArgs& p, int b, int lane, int writer_lane) {
  unsigned int prev = 0;
  if (lane == writer_lane) {
    __threadfence();
    prev = atomicAdd(&p.ticket[b], 1u);
  }
  prev = __shfl_sync(0xffffffffu, prev, writer_lane);
  if (prev != (unsigned int)p.tiles_per_state - 1u) return;
  __threadfence();
  const Candidate* cd = p.cand + (size_t)b * p.tiles_per_state;
  double c = 0.0;
  long long idx = -1;
  int src = -1;
  for (int t = lane; t < p.tiles_per_state; t += 32) {
    const double cc = __ldcg(&cd[t].cost);
    const long long ii = __ldcg(reinterpret_cast<const long long*>(&cd[t].index));
    if (ii >= 0 && (idx < 0 || cand_better(cc, (int64_t)ii, c, (int64_t)idx))) { c = cc; idx = ii; src = t; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double oc = __shfl_xor_sync(0xffffffffu, c, off);
    const long long oi = __shfl_xor_sync(0xffffffffu, idx, off);
    const int os = __shfl_xor_sync(0xffffffffu, src, off);
    if (oi >= 0 && (idx < 0 || cand_better(oc, (int64_t)oi, c, (int64_t)idx))) { c = oc; idx = oi; src = os; }
  }
  if (lane == 0) {
    bcmpc_record r;
    r.cost = idx < 0 ? 0.0 : c;
    r.index = idx;
    for (int j = 0; j < kMaxAct; ++j) r.action[j] = src >= 0 ? __ldcg(&cd[src].action[j]) : 0.0;
    p.out[b] = r;
    p.ticket[b] = 0u;
  }
}

struct NormArgs {
  double mean_obs[kMaxObs], den_obs[kMaxObs];      // den = std + 1e-10  (dynamics.py:109)
  double mean_act[kMaxAct], den_act[kMaxAct];      // (dynamics.py:110)
  double mean_delta[kMaxObs], std_delta[kMaxObs];  // (dynamics.py:113, no epsilon)
  float low[kMaxAct], span[kMaxAct];               // action box (controllers.py:53)
};

}  // namespace bcmpc
