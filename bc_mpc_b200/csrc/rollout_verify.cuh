// FP32 verification-mode rollout kernel (BCMPC_PREC_FP32_VERIFY).
//
// Follows the reference precision map exactly (dynamics.py:106-119): float64 normalise, float32
// MLP (FFMA, k-ascending accumulation, bias added after the sum like MatMul+BiasAdd), float64
// denormalise / state carry / cost (cost_functions.py:10-30,59-63).  One CTA owns M paths for
// all H steps; activations ping-pong between two shared-memory buffers, weights stream from L2.
// This is the mode every parity test runs against the oracle; the tcgen05 kernel in
// rollout_tc.cuh is the throughput path.
#pragma once
#include "common.cuh"

namespace bcmpc {

struct MlpDesc {
  int n_layers;
  int in_dim[kMaxLayers], out_dim[kMaxLayers];   // logical sizes
  int in_pad[kMaxLayers], out_pad[kMaxLayers];   // padded to multiples of 4 (zero filled)
  const float* W[kMaxLayers];                    // [in_pad][out_pad] row-major (TF kernel layout)
  const float* b[kMaxLayers];                    // [out_pad]
  const float* ln_gamma[kMaxLayers];             // LayerNorm after the activation of hidden layer l (nullptr = none)
  const float* ln_beta[kMaxLayers];
  float ln_eps;
};

struct VerifyArgs {
  PlanArgs p;
  NormArgs n;
  MlpDesc dyn;
  MlpDesc pol;
  int dyn_act;                 // BCMPC_ACT_*
  int ld;                      // row stride (floats) of the activation buffers
  int has_filter;
  float filt_mean[kMaxObs], filt_std[kMaxObs];
  float clip;                  // <= 0: none
  float pol_std[kMaxAct];      // exp(logstd), float32 (distributions.py:65)
  float one_minus_explore;     // float32(1 - e): python scalar * float32 array (controllers.py:208)
  double explore;
  // learned-reward planners (NNDynamicsRewardModel, dynamics.py:121-238; controllers.py:90-158,289-363): the reward head
  // branches off the first hidden layer of `dyn`; the objective is minus the (discounted) sum of predicted rewards
  MlpDesc rew;                 // 2 layers: hidden (dyn activation) and linear [hidden -> 1]; n_layers == 0: off
  double mean_reward, std_reward, gamma;
};

constexpr int kVerifyThreads = 256;

// out[M][out_pad] = act(in[M][in_pad] @ W + b).  Each warp owns 8 rows x 128 columns per pass;
// a lane owns 4 consecutive columns.  `in` reads are warp-broadcast LDS.128, W reads are
// coalesced LDG.128 that the M/8 row-group warps share through L1.
template <int M>
__device__ __forceinline__ void dense_layer(const float* __restrict__ W, const float* __restrict__ bias, int in_pad,
                                            int out_pad, const float* in_s, float* out_s, int ld, int act, int ld_out = 0) {
  if (ld_out == 0) ld_out = ld;
  constexpr int MG = M / 8;
  constexpr int NS = (kVerifyThreads / 32) / MG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mg = warp % MG, ns = warp / MG;
  const float* in_rows = in_s + (mg * 8) * ld;
  for (int n0 = ns * 128; n0 < out_pad; n0 += NS * 128) {
    const int n = n0 + lane * 4;
    if (n >= out_pad) continue;
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    const float* wp = W + n;
#pragma unroll 2
    for (int k = 0; k < in_pad; k += 4) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + 0) * out_pad));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + 1) * out_pad));
      const float4 w2 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + 2) * out_pad));
      const float4 w3 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + 3) * out_pad));
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(in_rows + r * ld + k);
        acc[r][0] = fmaf(a.x, w0.x, acc[r][0]); acc[r][1] = fmaf(a.x, w0.y, acc[r][1]);
        acc[r][2] = fmaf(a.x, w0.z, acc[r][2]); acc[r][3] = fmaf(a.x, w0.w, acc[r][3]);
        acc[r][0] = fmaf(a.y, w1.x, acc[r][0]); acc[r][1] = fmaf(a.y, w1.y, acc[r][1]);
        acc[r][2] = fmaf(a.y, w1.z, acc[r][2]); acc[r][3] = fmaf(a.y, w1.w, acc[r][3]);
        acc[r][0] = fmaf(a.z, w2.x, acc[r][0]); acc[r][1] = fmaf(a.z, w2.y, acc[r][1]);
        acc[r][2] = fmaf(a.z, w2.z, acc[r][2]); acc[r][3] = fmaf(a.z, w2.w, acc[r][3]);
        acc[r][0] = fmaf(a.w, w3.x, acc[r][0]); acc[r][1] = fmaf(a.w, w3.y, acc[r][1]);
        acc[r][2] = fmaf(a.w, w3.z, acc[r][2]); acc[r][3] = fmaf(a.w, w3.w, acc[r][3]);
      }
    }
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float4 o;
      o.x = acc[r][0] + bv.x; o.y = acc[r][1] + bv.y; o.z = acc[r][2] + bv.z; o.w = acc[r][3] + bv.w;
      if (act == BCMPC_ACT_RELU) {   // (x < 0) ? 0 : x propagates NaN like Eigen's / numpy's max (fmaxf would drop it)
        o.x = o.x < 0.f ? 0.f : o.x; o.y = o.y < 0.f ? 0.f : o.y; o.z = o.z < 0.f ? 0.f : o.z; o.w = o.w < 0.f ? 0.f : o.w;
      } else if (act == BCMPC_ACT_TANH) {
        o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w);
      }
      *reinterpret_cast<float4*>(out_s + (mg * 8 + r) * ld_out + n) = o;
    }
  }
}

// tf.contrib.layers.layer_norm over the last axis (dynamics.py:68-69), float32: one warp per row,
// mean, then mean of squared deviations, then (x - mean) * rsqrt(var + eps) * gamma + beta.  Padding columns stay 0.
template <int M>
__device__ __forceinline__ void layer_norm_rows(float* buf, int n, int n_pad, int ld, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < M; r += kVerifyThreads / 32) {
    float* x = buf + r * ld;
    float sum = 0.f;
    for (int i = lane; i < n; i += 32) sum += x[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float mean = sum / (float)n;
    float sq = 0.f;
    for (int i = lane; i < n; i += 32) { const float d = x[i] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    const float inv = 1.0f / sqrtf(sq / (float)n + eps);
    for (int i = lane; i < n_pad; i += 32) x[i] = i < n ? (x[i] - mean) * inv * gamma[i] + beta[i] : 0.f;
  }
}

// Runs an MLP whose input already sits in bufA (zero padded to in_pad[0]); hidden layers use
// `hidden_act`, the last layer is linear.  Returns the buffer holding the output.
template <int M>
__device__ __forceinline__ float* run_mlp(const MlpDesc& m, int hidden_act, float* bufA, float* bufB, int ld) {
  float* in = bufA;
  float* out = bufB;
  for (int l = 0; l < m.n_layers; ++l) {
    const int act = (l + 1 < m.n_layers) ? hidden_act : -1;
    dense_layer<M>(m.W[l], m.b[l], m.in_pad[l], m.out_pad[l], in, out, ld, act);
    __syncthreads();
    if (l + 1 < m.n_layers && m.ln_gamma[l] != nullptr) {
      layer_norm_rows<M>(out, m.out_dim[l], m.out_pad[l], ld, m.ln_gamma[l], m.ln_beta[l], m.ln_eps);
      __syncthreads();
    }
    float* t = in; in = out; out = t;
  }
  return in;
}

// Observation filter of MlpPolicy (ppo_bc_policy.py:60), float32 like the TF graph.
__device__ __forceinline__ float policy_input(const VerifyArgs& a, double s, int d) {
  float x = (float)s;   // feed cast to the tf.float32 placeholder
  if (a.has_filter) x = __fdiv_rn(__fsub_rn(x, a.filt_mean[d]), a.filt_std[d]);
  if (a.clip > 0.f) x = fminf(fmaxf(x, -a.clip), a.clip);
  return x;
}

template <int M>
__global__ void __launch_bounds__(kVerifyThreads) rollout_verify_kernel(const __grid_constant__ VerifyArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_state = reinterpret_cast<double*>(smem_raw);        // [M][kMaxObs]
  double* s_action = s_state + M * kMaxObs;                     // [M][kMaxAct] action of this step
  double* s_act0 = s_action + M * kMaxAct;                      // [M][kMaxAct] action of step 0
  double* s_inj = s_act0 + M * kMaxAct;                         // [M][kMaxAct]
  double* s_cost = s_inj + M * kMaxAct;                         // [M]
  float* s_rew = reinterpret_cast<float*>(s_cost + M);          // [M][4] reward-head output (column 0)
  float* bufA = s_rew + M * 4;
  float* bufB = bufA + M * a.ld;

  const PlanArgs& p = a.p;
  const int D = p.D, A = p.A, K = p.K, H = p.H, ld = a.ld;
  const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int k_base = tile * M;

  for (int i = tid; i < M * D; i += kVerifyThreads) {
    const int r = i / D, d = i % D;
    const int k = min(k_base + r, K - 1);
    s_state[r * kMaxObs + d] = root_state(p, (size_t)(p.per_row_states ? k : b), D, d);
    if (p.states_out && k_base + r < K) p.states_out[(size_t)(k_base + r) * D + d] = s_state[r * kMaxObs + d];
  }
  if (tid < M) s_cost[tid] = 0.0;
  __syncthreads();

  for (int h = 0; h < H; ++h) {
    const int prop = h == 0 ? p.proposal0 : p.proposal;
    // ---- injected randomness of this step (controllers.py:53 / distributions.py:80) ----
    if (p.inj) {
      for (int i = tid; i < M * A; i += kVerifyThreads) {
        const int r = i / A, j = i % A;
        const int k = min(k_base + r, K - 1);
        s_inj[r * kMaxAct + j] = p.inj[(((size_t)b * H + h) * K + k) * A + j];
      }
    } else if (tid < M) {
      float v[kMaxAct];
      draw_injection(v, A, h == 0 ? p.inj_kind0 : p.inj_kind, plan_seed(p), (uint32_t)(p.path_offset + k_base + tid), (uint32_t)b, (uint32_t)h,
                     a.n.low, a.n.span);
      for (int j = 0; j < A; ++j) s_inj[tid * kMaxAct + j] = (double)v[j];
    }
    // ---- proposal network (ppo_bc_policy.py:60,70-72 / behavioral_cloning.py:21-26) ----
    const float* mean = nullptr;
    if (prop != BCMPC_PROP_UNIFORM) {
      const int ip = a.pol.in_pad[0];
      for (int i = tid; i < M * ip; i += kVerifyThreads) {
        const int r = i / ip, d = i % ip;
        bufA[r * ld + d] = d < D ? policy_input(a, s_state[r * kMaxObs + d], d) : 0.f;
      }
      __syncthreads();
      mean = run_mlp<M>(a.pol, BCMPC_ACT_TANH, bufA, bufB, ld);
    } else {
      __syncthreads();
    }
    // ---- action of this step ----
    for (int i = tid; i < M * A; i += kVerifyThreads) {
      const int r = i / A, j = i % A;
      const double z = s_inj[r * kMaxAct + j];
      double act;
      if (prop == BCMPC_PROP_UNIFORM) {
        act = z;                                                                   // controllers.py:70
      } else if (prop == BCMPC_PROP_POLICY_GAUSSIAN) {
        act = (double)__fadd_rn(mean[r * ld + j], __fmul_rn(a.pol_std[j], (float)z));   // distributions.py:80
      } else if (prop == BCMPC_PROP_POLICY_MIX) {
        act = (double)__fmul_rn(a.one_minus_explore, mean[r * ld + j]) + a.explore * z;  // controllers.py:208
      } else if (prop == BCMPC_PROP_POLICY_MEAN) {
        act = (double)mean[r * ld + j];                                            // controllers.py:429
      } else {
        act = (double)(float)((double)mean[r * ld + j] + z);   // float32 array += float64 (controllers.pyc)
      }
      s_action[r * kMaxAct + j] = act;
      if (h == 0) s_act0[r * kMaxAct + j] = act;
      if (p.actions_out && k_base + r < K) p.actions_out[((size_t)h * K + k_base + r) * A + j] = act;
    }
    __syncthreads();
    // ---- normalise -> float32 network input (dynamics.py:109-110,26) ----
    {
      const int ip = a.dyn.in_pad[0];
      for (int i = tid; i < M * ip; i += kVerifyThreads) {
        const int r = i / ip, c = i % ip;
        float x = 0.f;
        if (c < D) x = (float)((s_state[r * kMaxObs + c] - a.n.mean_obs[c]) / a.n.den_obs[c]);
        else if (c < D + A) x = (float)((s_action[r * kMaxAct + c - D] - a.n.mean_act[c - D]) / a.n.den_act[c - D]);
        bufA[r * ld + c] = x;
      }
    }
    __syncthreads();
    const float* y;
    if (a.rew.n_layers == 0) {
      y = run_mlp<M>(a.dyn, a.dyn_act, bufA, bufB, ld);   // dynamics.py:112
    } else {
      // NNDynamicsRewardModel.build_network (dynamics.py:162-193): shared layer, then the reward head (bufB must survive
      // it), then the state-delta head
      dense_layer<M>(a.dyn.W[0], a.dyn.b[0], a.dyn.in_pad[0], a.dyn.out_pad[0], bufA, bufB, ld, a.dyn_act);
      __syncthreads();
      if (a.dyn.ln_gamma[0]) { layer_norm_rows<M>(bufB, a.dyn.out_dim[0], a.dyn.out_pad[0], ld, a.dyn.ln_gamma[0], a.dyn.ln_beta[0], a.dyn.ln_eps); __syncthreads(); }
      dense_layer<M>(a.rew.W[0], a.rew.b[0], a.rew.in_pad[0], a.rew.out_pad[0], bufB, bufA, ld, a.dyn_act);
      __syncthreads();
      if (a.rew.ln_gamma[0]) { layer_norm_rows<M>(bufA, a.rew.out_dim[0], a.rew.out_pad[0], ld, a.rew.ln_gamma[0], a.rew.ln_beta[0], a.rew.ln_eps); __syncthreads(); }
      dense_layer<M>(a.rew.W[1], a.rew.b[1], a.rew.in_pad[1], a.rew.out_pad[1], bufA, s_rew, ld, -1, 4);
      __syncthreads();
      dense_layer<M>(a.dyn.W[1], a.dyn.b[1], a.dyn.in_pad[1], a.dyn.out_pad[1], bufB, bufA, ld, a.dyn_act);
      __syncthreads();
      if (a.dyn.ln_gamma[1]) { layer_norm_rows<M>(bufA, a.dyn.out_dim[1], a.dyn.out_pad[1], ld, a.dyn.ln_gamma[1], a.dyn.ln_beta[1], a.dyn.ln_eps); __syncthreads(); }
      dense_layer<M>(a.dyn.W[2], a.dyn.b[2], a.dyn.in_pad[2], a.dyn.out_pad[2], bufA, bufB, ld, -1);
      __syncthreads();
      y = bufB;
    }
    // ---- cost on the pre-step state (cost_functions.py:16-28), then s' = s + delta ----
    if (tid < M) {
      const double* s = s_state + tid * kMaxObs;
      double c = 0.0;
      if (a.rew.n_layers != 0) {
        // controllers.py:137-139: rewards_all.append(reward * gamma ** i), reward = y * std_reward + mean_reward (dynamics.py:236)
        c = -(((double)s_rew[tid * 4] * a.std_reward + a.mean_reward) * pow(a.gamma, (double)h));
      } else if (D > 17) {
        if (s[5] >= 0.2) c += 10.0;
        if (s[6] >= 0.0) c += 10.0;
        if (s[7] >= 0.0) c += 10.0;
        const double nx = s[17] + ((double)y[tid * ld + 17] * a.n.std_delta[17] + a.n.mean_delta[17]);
        c -= (nx - s[17]) / 0.01;
      }
      s_cost[tid] += c;                                              // cost_functions.py:61-62
    }
    __syncthreads();
    for (int i = tid; i < M * D; i += kVerifyThreads) {
      const int r = i / D, d = i % D;
      const double delta = (double)y[r * ld + d] * a.n.std_delta[d] + a.n.mean_delta[d];   // dynamics.py:113
      const double nx = s_state[r * kMaxObs + d] + delta;                                  // dynamics.py:116
      s_state[r * kMaxObs + d] = nx;
      if (p.states_out && k_base + r < K) p.states_out[((size_t)(h + 1) * K + k_base + r) * D + d] = nx;
    }
    __syncthreads();
  }

  // ---- per-tile argmin (np.argmin order), one warp ----
  if (tid < 32) {
    double c = 0.0;
    int64_t idx = INT64_MAX;
    int src = tid;
    if (tid < M && k_base + tid < K) {
      c = s_cost[tid];
      idx = p.path_offset + k_base + tid;
      if (p.costs_out) p.costs_out[(size_t)b * K + k_base + tid] = c;
    }
    bool have = idx != INT64_MAX;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double oc = __shfl_xor_sync(0xffffffffu, c, off);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, idx, off);
      const int os = __shfl_xor_sync(0xffffffffu, src, off);
      const bool oh = __shfl_xor_sync(0xffffffffu, (int)have, off);
      if (oh && (!have || cand_better(oc, oi, c, idx))) { c = oc; idx = oi; src = os; have = true; }
    }
    if (tid == 0 && p.cand) {
      Candidate& o = p.cand[(size_t)b * p.tiles_per_state + tile];
      o.cost = c;
      o.index = have ? idx : -1;
      for (int j = 0; j < kMaxAct; ++j) o.action[j] = (have && j < A) ? s_act0[src * kMaxAct + j] : 0.0;
    }
  }
}

inline size_t verify_smem_bytes(int M, int ld) {
  return (size_t)M * (kMaxObs + 3 * kMaxAct + 1) * sizeof(double) + (size_t)M * 4 * sizeof(float) + 2 * (size_t)M * ld * sizeof(float);
}

// Stand-alone forward of the proposal network: policy_net.act / BCnetwork.predict.
template <int M>
__global__ void __launch_bounds__(kVerifyThreads) policy_act_kernel(const __grid_constant__ VerifyArgs a,
                                                                   const double* __restrict__ obs,
                                                                   const float* __restrict__ noise, int64_t n,
                                                                   float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* bufA = reinterpret_cast<float*>(smem_raw);
  float* bufB = bufA + M * a.ld;
  const int D = a.p.D, A = a.p.A, ld = a.ld, tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * M;
  const int ip = a.pol.in_pad[0];
  for (int i = tid; i < M * ip; i += kVerifyThreads) {
    const int r = i / ip, d = i % ip;
    const int64_t row = min(base + r, n - 1);
    bufA[r * ld + d] = d < D ? policy_input(a, obs[row * D + d], d) : 0.f;
  }
  __syncthreads();
  const float* mean = run_mlp<M>(a.pol, BCMPC_ACT_TANH, bufA, bufB, ld);
  for (int i = tid; i < M * A; i += kVerifyThreads) {
    const int r = i / A, j = i % A;
    if (base + r >= n) continue;
    float v = mean[r * ld + j];
    if (noise) v = __fadd_rn(v, __fmul_rn(a.pol_std[j], noise[(base + r) * A + j]));
    out[(base + r) * A + j] = v;
  }
}

}  // namespace bcmpc
