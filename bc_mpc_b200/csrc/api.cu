// C ABI of the planner (include/bcmpc.h): handle, weight upload, launch logic.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bcmpc.h"
#include "aux_kernels.cuh"
#include "common.cuh"
#include "rollout_tmem.cuh"
#include "weight_pack.cuh"
#include "rollout_narrow.cuh"
#include "rollout_duo.cuh"
#include "rollout_deep.cuh"
#include "deep_pack.cuh"
#include "rollout_verify.cuh"

using namespace bcmpc;

namespace {
thread_local std::string g_err;

struct DeviceMlp {
  MlpDesc desc{};
  std::vector<float*> bufs;
  bool set = false;
  void release() {
    for (float* p : bufs) cudaFree(p);
    bufs.clear();
    set = false;
  }
};
}  // namespace

struct bcmpc_handle {
  bcmpc_config cfg{};
  std::string err;
  cudaStream_t stream = nullptr;
  DeviceMlp dyn, pol, rew;       // rew: reward head of NNDynamicsRewardModel (bcmpc_set_reward_head)
  double mean_reward = 0.0, std_reward = 1.0, gamma = 1.0;
  std::vector<float*> ln_bufs;   // LayerNorm gamma/beta of the dynamics net (bcmpc_set_layer_norm)
  float* rew_ln_buf = nullptr;   // [gamma (hidden)][beta (hidden)] of the reward head's hidden layer (bcmpc_set_reward_layer_norm)
  bool rew_ln_on = false;
  NormArgs norm{};
  bool have_norm = false, have_bounds = false;
  // proposal network meta
  bool has_filter = false, has_logstd = false;
  float filt_mean[kMaxObs]{}, filt_std[kMaxObs]{}, pol_std[kMaxAct]{}, clip = 0.f;
  // verify-kernel geometry
  int vM = 32, vld = 0;
  // scratch
  unsigned int* d_ticket = nullptr;   // [max_states] finished-tile counters of the fused min-loc
  Candidate* d_cand = nullptr;
  size_t cand_cap = 0;
  double* d_states = nullptr;
  bcmpc_record* d_records = nullptr;
  double* h_states = nullptr;        // pinned
  bcmpc_record* h_records = nullptr; // pinned, mapped into the device address space
  bcmpc_record* h_records_dev = nullptr;   // its device-side address: small plans write their records straight to the host
  double* d_inj = nullptr;
  size_t inj_cap = 0;
  int64_t launches = 0;
  bool timing = false;
  bool inline_state_ok = true;      // BCMPC_INLINE_STATE=0: single plans copy their root state like batched ones
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  // tcgen05 path (rollout_tmem.cuh)
  tmx::TmHostPlan tmp;
  uint8_t *d_res = nullptr, *d_stream = nullptr, *d_pol = nullptr;
  size_t pol_cap = 0;
  bool pol_wide = false;          // the proposal network has more than 128 units in a hidden layer: kVarPolWide kernel variants
  float* d_polvec = nullptr;      // biases of the proposal network for the fast kernel
  float* d_lnvec = nullptr;
  size_t res_cap = 0, stream_cap = 0, lnvec_cap = 0;
  bool ln_on = false;     // LayerNorm parameters uploaded (ln_bufs / dyn.desc.ln_*), both precision modes
  float ln_eps = 1e-12f;
  cudaEvent_t ev_weights = nullptr;   // recorded after an asynchronous weight hand-off (bcmpc_set_dynamics_device)
  bool weights_pending = false;       // plans on other streams wait on ev_weights until it has completed
  unsigned long long* d_trace = nullptr;
  bool tm_ready = false;
  bool tm_narrow = false; // the weight images suit rollout_narrow_kernel (hidden <= 256, CTA pairs, no reward head)
  bool tm_wide_ok = true; // ... and rollout_tmem_kernel has an instantiation that reads the same images
  bool tm_use_narrow = false;   // the kernel chosen for the launch in flight (run_plan)
  bool tm_use_duo = false;      // ... rollout_duo_kernel (two tiles per CTA taking turns on one hidden-1 buffer: hidden 257..510, more than one wave)
  int duo_mode = 0;             // BCMPC_DUO: 0 = never (default: measured 3-5 % slower than rollout_tmem_kernel), 1 = when the launch has more CTA-pair units than the device has SM pairs, 2 = always where it applies
  int sm_pairs = 74;      // CTA-pair slots of the device (SMs / 2): one wave of rollout_tmem_kernel
  int narrow_threshold = 74;   // launches with more CTA-pair units than this go to rollout_narrow_kernel (where it applies)
  int tm_narrow_mode = 1; // BCMPC_NARROW=0 keeps hidden <= 256 on rollout_tmem_kernel (rollout_narrow_kernel is the default there)
  PeerExchangeArgs peer{};     // bcmpc_set_peer_exchange: world == 0 means off
  bool deep = false;      // the cluster-split kernel (rollout_deep.cuh): n_hidden != 2 or hidden_size > 510
  int deep_nch = 0;       // 64-unit chunks per K-half
  int deep_clusters = 0;  // co-resident clusters of four (cudaOccupancyMaxActiveClusters, cached after the first launch)
  int tm_cg = 2;          // BCMPC_TC_CG: CTAs per tcgen05 group (2 = CTA pairs, the default; 1 = single CTAs)
  uint32_t tm_hint_chain = 0x989680u;   // BCMPC_TC_HINT_CHAIN: mbarrier try_wait suspend hint (ns) on the per-step dependency chain
  std::string trace_path; // BCMPC_TC_TRACE: dump the pipeline trace of CTA (0,0) after every launch
};

namespace {

int fail(bcmpc_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  if (h) h->err = buf;
  return code;
}

#define CK(h, expr)                                                                                   \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess) return fail(h, BCMPC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)


// Every entry point runs on the handle's device and restores the caller's current device on return (a drop-in must not
// re-point a torch process's current device: allocations and current streams would silently move to another GPU).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    else if (err == cudaSuccess) prev = -1;          // already current: nothing to restore
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ON_DEVICE(h)                                                                                        \
  DeviceGuard dev_guard__((h)->cfg.device);                                                                 \
  if (dev_guard__.err != cudaSuccess) return fail(h, BCMPC_ERR_CUDA, "cudaSetDevice(%d): %s", (h)->cfg.device, cudaGetErrorString(dev_guard__.err))

int round4(int x) { return (x + 3) / 4 * 4; }

// Weight refresh into persistent zero-padded float32 device buffers: a strided copy - from the host or from device
// memory - on `st`; the buffers are (re)allocated only when the layer shapes change (the dynamics shape is fixed by
// bcmpc_config, a proposal network keeps its shape between PPO / BC updates).
// kind: cudaMemcpyHostToDevice ([in,out] host arrays), cudaMemcpyDeviceToDevice ([in,out] device arrays), or -1 for
// device arrays in torch.nn.Linear layout [out,in] (transposed by a small kernel).
int refresh_mlp(bcmpc_handle* h, DeviceMlp& m, int n_layers, const int* sizes, const float* const* W, const float* const* b, int kind,
                cudaStream_t st) {
  bool same = m.set && m.desc.n_layers == n_layers && (int)m.bufs.size() == 2 * n_layers;
  for (int l = 0; l < n_layers && same; ++l) same = m.desc.in_dim[l] == sizes[l] && m.desc.out_dim[l] == sizes[l + 1];
  if (!same) {
    CK(h, cudaDeviceSynchronize());     // nothing may still read the old buffers
    m.release();
    m.desc = MlpDesc{};
    m.desc.n_layers = n_layers;
    for (int l = 0; l < n_layers; ++l) {
      const int ip = round4(sizes[l]), op = round4(sizes[l + 1]);
      m.desc.in_dim[l] = sizes[l]; m.desc.out_dim[l] = sizes[l + 1]; m.desc.in_pad[l] = ip; m.desc.out_pad[l] = op;
      float *dw = nullptr, *db = nullptr;
      CK(h, cudaMalloc(&dw, (size_t)ip * op * sizeof(float)));
      m.bufs.push_back(dw);
      CK(h, cudaMalloc(&db, (size_t)op * sizeof(float)));
      m.bufs.push_back(db);
      CK(h, cudaMemsetAsync(dw, 0, (size_t)ip * op * sizeof(float), st));
      CK(h, cudaMemsetAsync(db, 0, (size_t)op * sizeof(float), st));
      m.desc.W[l] = dw;
      m.desc.b[l] = db;
    }
  }
  for (int l = 0; l < n_layers; ++l) {
    const size_t row = sizeof(float) * sizes[l + 1];
    if (kind < 0) {
      const int n = sizes[l] * sizes[l + 1];
      transpose_in_kernel<<<(n + 255) / 256, 256, 0, st>>>(const_cast<float*>(m.desc.W[l]), m.desc.out_pad[l], W[l], sizes[l], sizes[l + 1]);
      CK(h, cudaGetLastError());
      h->launches++;
      CK(h, cudaMemcpyAsync(const_cast<float*>(m.desc.b[l]), b[l], row, cudaMemcpyDeviceToDevice, st));
    } else {
      CK(h, cudaMemcpy2DAsync(const_cast<float*>(m.desc.W[l]), sizeof(float) * m.desc.out_pad[l], W[l], row, row, sizes[l], (cudaMemcpyKind)kind, st));
      CK(h, cudaMemcpyAsync(const_cast<float*>(m.desc.b[l]), b[l], row, (cudaMemcpyKind)kind, st));
    }
  }
  for (int l = 0; l < kMaxLayers; ++l) m.desc.ln_gamma[l] = m.desc.ln_beta[l] = nullptr;   // LayerNorm parameters belong to one checkpoint
  m.set = true;
  return BCMPC_OK;
}

int refresh_dynamics(bcmpc_handle* h, const float* const* W, const float* const* b, int kind, cudaStream_t st) {
  const int n_layers = h->cfg.n_hidden + 1;
  int sizes[kMaxLayers + 1];
  sizes[0] = h->cfg.obs_dim + h->cfg.act_dim;
  for (int l = 1; l <= h->cfg.n_hidden; ++l) sizes[l] = h->cfg.hidden_size;
  sizes[n_layers] = h->cfg.obs_dim;
  int rc = refresh_mlp(h, h->dyn, n_layers, sizes, W, b, kind, st);
  if (rc) return rc;
  h->ln_on = false;     // new weights reset LayerNorm: call bcmpc_set_layer_norm again
  return BCMPC_OK;
}

void update_verify_geometry(bcmpc_handle* h) {
  int ld = 4;
  auto scan = [&](const DeviceMlp& m) {
    if (!m.set) return;
    for (int l = 0; l < m.desc.n_layers; ++l) { ld = std::max(ld, m.desc.in_pad[l]); ld = std::max(ld, m.desc.out_pad[l]); }
  };
  scan(h->dyn);
  scan(h->pol);
  scan(h->rew);
  h->vld = ld;
  h->vM = ld <= 512 ? 32 : 16;
}

int inj_kind_of(int proposal) {
  switch (proposal) {
    case BCMPC_PROP_POLICY_GAUSSIAN: return kInjNormal;
    case BCMPC_PROP_NET_PLUS_UNIFORM: return kInjUniformPm1;
    default: return kInjUniformBox;
  }
}

int first_proposal_of(const bcmpc_config& c) { return c.first_proposal > 0 ? c.first_proposal - 1 : c.proposal; }
bool needs_policy(const bcmpc_config& c) { return c.proposal != BCMPC_PROP_UNIFORM || first_proposal_of(c) != BCMPC_PROP_UNIFORM; }

int check_ready(bcmpc_handle* h, bool need_policy) {
  if (!h->dyn.set) return fail(h, BCMPC_ERR_STATE, "dynamics weights not set (bcmpc_set_dynamics)");
  if (!h->have_norm) return fail(h, BCMPC_ERR_STATE, "normalisation not set (bcmpc_set_normalization)");
  if (need_policy) {
    if (!h->pol.set) return fail(h, BCMPC_ERR_STATE, "proposal network not set (bcmpc_set_policy)");
    if ((h->cfg.proposal == BCMPC_PROP_POLICY_GAUSSIAN || first_proposal_of(h->cfg) == BCMPC_PROP_POLICY_GAUSSIAN) && !h->has_logstd)
      return fail(h, BCMPC_ERR_STATE, "POLICY_GAUSSIAN needs logstd in bcmpc_set_policy");
  }
  return BCMPC_OK;
}

void fill_verify_args(bcmpc_handle* h, VerifyArgs* a) {
  memset(a, 0, sizeof *a);
  a->n = h->norm;
  a->dyn = h->dyn.desc;
  if (h->pol.set) a->pol = h->pol.desc;
  if (h->rew.set) { a->rew = h->rew.desc; a->mean_reward = h->mean_reward; a->std_reward = h->std_reward; a->gamma = h->gamma; }
  a->dyn_act = h->cfg.activation;
  a->ld = h->vld;
  a->has_filter = h->has_filter;
  memcpy(a->filt_mean, h->filt_mean, sizeof a->filt_mean);
  memcpy(a->filt_std, h->filt_std, sizeof a->filt_std);
  a->clip = h->clip;
  memcpy(a->pol_std, h->pol_std, sizeof a->pol_std);
  a->one_minus_explore = (float)(1.0 - h->cfg.explore);
  a->explore = h->cfg.explore;
}

template <int M>
int launch_verify_t(bcmpc_handle* h, const VerifyArgs& a, dim3 grid, cudaStream_t st) {
  const size_t smem = verify_smem_bytes(M, a.ld);
  CK(h, cudaFuncSetAttribute(rollout_verify_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rollout_verify_kernel<M><<<grid, kVerifyThreads, smem, st>>>(a);
  CK(h, cudaGetLastError());
  h->launches++;
  return BCMPC_OK;
}

int ensure_cand(bcmpc_handle* h, size_t n) {
  if (n <= h->cand_cap) return BCMPC_OK;
  if (h->d_cand) cudaFree(h->d_cand);
  h->d_cand = nullptr;
  h->cand_cap = 0;
  CK(h, cudaMalloc(&h->d_cand, n * sizeof(Candidate)));
  h->cand_cap = n;
  return BCMPC_OK;
}

// the narrow-net kernel: plain ReLU dynamics of at most 256 hidden units, uniform shooting, CTA pairs
bool narrow_applies(const bcmpc_handle* h) {
  if (h->pol_wide && needs_policy(h->cfg)) return false;        // wide proposal networks run on rollout_tmem_kernel only
  return h->tm_narrow_mode != 0 && h->tm_cg == 2 && !h->rew.set && h->cfg.activation == BCMPC_ACT_RELU &&
         h->cfg.hidden_size <= 64 * tmx::kNwMaxChunks;
}

// the two-tiles-per-CTA kernel for wide nets: plain dynamics net (ReLU with the biases in the GEMMs, or tanh), uniform shooting, CTA pairs
bool duo_applies(const bcmpc_handle* h, const PlanArgs& p) {
  const tmx::TmHostPlan& P = h->tmp;
  if (h->duo_mode == 0 || h->tm_cg != 2 || !h->tm_ready || P.layer_norm || P.reward) return false;
  if (p.proposal != BCMPC_PROP_UNIFORM || p.proposal0 != BCMPC_PROP_UNIFORM) return false;
  const bool plain_tanh = h->cfg.activation == BCMPC_ACT_TANH;
  const int var = (P.explicit_bias ? tmx::kVarBias : 0) | (plain_tanh ? tmx::kVarTanh : 0);
  return tmx::du_supported(P.nc, var);
}

int tm_variant(const bcmpc_handle* h) {
  const int pol = needs_policy(h->cfg) ? tmx::kVarPolicy : 0;
  const int wide = (pol && h->pol_wide) ? tmx::kVarPolWide : 0;
  int var = pol | wide | (h->ln_on ? tmx::kVarLayerNorm : 0) | (h->rew.set ? tmx::kVarReward : 0);
  if (narrow_applies(h)) return pol | (h->ln_on ? tmx::kVarLayerNorm : tmx::kVarBias);     // rollout_narrow_kernel: explicit biases (or LayerNorm vectors)
  // a plain tanh NNDynamicsModel (the reference builder's default activation, dynamics.py:60): biases in the epilogues
  if (h->cfg.activation == BCMPC_ACT_TANH && !h->rew.set) return pol | wide | tmx::kVarBias | tmx::kVarTanh | (h->ln_on ? tmx::kVarLayerNorm : 0);
  // plain ReLU dynamics whose two constant-1 hidden units would need an extra chunk (hidden 63/64, 127/128, 255/256):
  // biases in the epilogues instead
  const int S = h->cfg.hidden_size;
  if (!wide && (var & ~tmx::kVarPolicy) == 0 && (S + 2 + 63) / 64 > (S + 63) / 64 && tmx::tm_width_supported(h->tm_cg, (S + 63) / 64, var | tmx::kVarBias))
    var |= tmx::kVarBias;
  return var;
}

// (re)packs the fast-mode weight images on the device, on stream `st`, from the float32 copies in h->dyn / h->ln_bufs
// cluster-split kernel: (re)pack its weight image [4 ranks][stages per step][16 KB] and bias vector on the device
int deep_prepare(bcmpc_handle* h, cudaStream_t st) {
  const int S = h->cfg.hidden_size, L = h->cfg.n_hidden;
  const int nch = (S + 127) / 128;
  h->deep_nch = nch;
  const size_t img = (size_t)dpx::kCluster * dpx::stages_per_step(nch, L) * dpx::kStageBytes;
  const size_t vec = (h->ln_on ? (size_t)(L - 1) * 2 * 128 * nch + 64 : (size_t)(L - 1) * 128 * nch + 32) * sizeof(float);   // LayerNorm: [u][c] per layer
  auto ensure = [&](void** ptr, size_t* cap, size_t need) -> cudaError_t {
    if (need <= *cap) return cudaSuccess;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(ptr, need);
    if (e == cudaSuccess) *cap = need;
    return e;
  };
  CK(h, ensure((void**)&h->d_stream, &h->stream_cap, img));
  CK(h, ensure((void**)&h->d_lnvec, &h->lnvec_cap, vec));
  dpx::PackArgs a{};
  const MlpDesc& d = h->dyn.desc;
  for (int l = 0; l <= L; ++l) { a.W[l] = d.W[l]; a.b[l] = d.b[l]; a.ld[l] = d.out_pad[l]; }
  if (h->ln_on)
    for (int l = 0; l < L; ++l) { a.g[l] = d.ln_gamma[l]; a.be[l] = d.ln_beta[l]; }
  a.L = L; a.S = S; a.nch = nch;
  a.stream = reinterpret_cast<uint16_t*>(h->d_stream); a.vec = h->d_lnvec;
  CK(h, dpx::deep_pack(a, st));
  h->launches += h->ln_on ? 2 : 1;
  if (!h->trace_path.empty() && !h->d_trace) CK(h, cudaMalloc(&h->d_trace, sizeof(unsigned long long) * dpx::kTraceWords));
  h->tm_ready = true;
  return BCMPC_OK;
}

// proposal-network arguments of the fast kernels (policy-guided shooting)
int fill_policy_args(bcmpc_handle* h, tmx::TmArgs& a) {
  const tmx::TmHostPlan& P = h->tmp;
  if (!P.with_policy || !h->d_pol) return fail(h, BCMPC_ERR_STATE, "proposal network not set (bcmpc_set_policy)");
  a.pol = h->d_pol; a.np1 = P.np1; a.np2 = P.np2;
  a.has_filter = h->has_filter ? 1 : 0; a.clip = h->clip;
  a.explore = (float)h->cfg.explore; a.one_minus_explore = (float)(1.0 - h->cfg.explore);
  for (int d = 0; d < h->cfg.obs_dim; ++d) { a.filt_mean[d] = h->filt_mean[d]; a.filt_inv_std[d] = 1.0f / h->filt_std[d]; }
  for (int j = 0; j < h->cfg.act_dim; ++j) a.pol_std[j] = h->pol_std[j];
  a.polvec = h->d_polvec;
  return BCMPC_OK;
}

int launch_deep(bcmpc_handle* h, const PlanArgs& p, cudaStream_t st) {
  if (!h->tm_ready) return fail(h, BCMPC_ERR_STATE, "bf16 fast path not prepared: set the dynamics weights first");
  dpx::DeepArgs a;
  memset(&a, 0, sizeof a);
  a.t.p = p;
  a.t.stream = h->d_stream; a.t.lnvec = h->d_lnvec;
  a.t.hint_chain = h->tm_hint_chain;
  for (int d = 0; d < h->cfg.obs_dim; ++d) {
    a.t.mean_obs[d] = (float)h->norm.mean_obs[d];
    a.t.inv_den_obs[d] = (float)(1.0 / h->norm.den_obs[d]);
    a.t.mean_delta[d] = (float)h->norm.mean_delta[d];
    a.t.std_delta[d] = (float)h->norm.std_delta[d];
  }
  for (int j = 0; j < h->cfg.act_dim; ++j) {
    a.t.mean_act[j] = (float)h->norm.mean_act[j];
    a.t.inv_den_act[j] = (float)(1.0 / h->norm.den_act[j]);
    a.t.low[j] = h->norm.low[j];
    a.t.span[j] = h->norm.span[j];
  }
  a.t.proposal = p.proposal; a.t.inj_kind = p.inj_kind;
  a.t.proposal0 = p.proposal0; a.t.inj_kind0 = p.inj_kind0;
  const bool pol = p.proposal != BCMPC_PROP_UNIFORM || p.proposal0 != BCMPC_PROP_UNIFORM;
  if (pol) {
    if (h->pol_wide) return fail(h, BCMPC_ERR_UNSUPPORTED, "bf16 fast path: the cluster-split kernel holds proposal networks of at most 128 units per hidden layer");
    int rc = fill_policy_args(h, a.t);
    if (rc) return rc;
  }
  a.n_hidden = h->cfg.n_hidden; a.s_half = 64 * h->deep_nch;
  if (h->ln_on) { a.t.ln_inv_n = 1.0f / (float)h->cfg.hidden_size; a.t.ln_eps = h->ln_eps; }
  const int64_t units = (int64_t)p.B * (p.tiles_per_state / 2);
  a.t.trace = h->d_trace;
  if (h->d_trace) CK(h, cudaMemsetAsync(h->d_trace, 0, sizeof(unsigned long long) * dpx::kTraceWords, st));
  CK(h, dpx::deep_launch(h->deep_nch, h->cfg.activation == BCMPC_ACT_TANH, h->ln_on, pol, a, h->deep_clusters, units, st, &h->deep_clusters));
  CK(h, cudaGetLastError());
  h->launches++;
  if (h->d_trace) {
    CK(h, cudaStreamSynchronize(st));
    std::vector<unsigned long long> t((size_t)dpx::kTraceWords);
    CK(h, cudaMemcpy(t.data(), h->d_trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(h->trace_path.c_str(), "wb")) {
      const int hdr[6] = {dpx::kTrSteps, dpx::kTrRoles, dpx::kTrUses, dpx::kTrStamps, h->deep_nch, h->cfg.n_hidden};
      fwrite(hdr, sizeof hdr, 1, f);
      fwrite(t.data(), sizeof(unsigned long long), t.size(), f);
      fclose(f);
    }
  }
  return BCMPC_OK;
}

int tm_prepare(bcmpc_handle* h, cudaStream_t st) {
  if (h->deep) return deep_prepare(h, st);
  const int S = h->cfg.hidden_size;
  const tmx::TmHostPlan old = h->tmp;
  tmx::tm_geometry(S, h->tm_cg, tm_variant(h), &h->tmp);
  if (old.with_policy) {   // the proposal network (its own device images) survives a dynamics refresh
    h->tmp.with_policy = true; h->tmp.np1 = old.np1; h->tmp.np2 = old.np2;
  }
  const tmx::TmHostPlan& P = h->tmp;
  if (P.nc < 1 || P.nc > 8) return fail(h, BCMPC_ERR_UNSUPPORTED, "hidden width %d does not fit the TMEM-resident kernel", S);
  h->tm_narrow = narrow_applies(h);
  h->tm_wide_ok = tmx::tm_width_supported(h->tm_cg, P.nc, tm_variant(h));
  if (h->cfg.activation == BCMPC_ACT_TANH && !h->rew.set && h->ln_on && !h->tm_wide_ok) {
    // plain tanh + LayerNorm has no instantiation in this mode (single-CTA developer mode, or a 256-unit proposal network), but the
    // learned-reward model might: its head may still follow (bcmpc_set_reward_head comes after bcmpc_set_layer_norm); a plan
    // before that fails in launch_tm
    h->tm_ready = false;
    return BCMPC_OK;
  }
  if (!h->tm_narrow && !h->tm_wide_ok) {
    return fail(h, BCMPC_ERR_UNSUPPORTED, "bf16 fast path: no tensor-core variant for this combination at hidden size %d with BCMPC_TC_CG=%d (single-CTA developer mode: the learned-reward model and LayerNorm + proposal network are built for hidden 193..256 and 449..512, no plain tanh); use BCMPC_PREC_FP32_VERIFY", S, h->tm_cg);
  }
  if (P.reward && h->ln_on != h->rew_ln_on) {
    // between bcmpc_set_layer_norm and bcmpc_set_reward_layer_norm: the images are packed once all three LayerNorms agree
    h->tm_ready = false;
    return BCMPC_OK;
  }
  if (P.lay.nst < 2) return fail(h, BCMPC_ERR_UNSUPPORTED, "weight ring does not fit in shared memory");
  auto ensure = [&](void** ptr, size_t* cap, size_t need) -> cudaError_t {
    if (need <= *cap) return cudaSuccess;
    if (*ptr) cudaFree(*ptr);           // (implicitly waits for every stream of the device)
    *ptr = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(ptr, need);
    if (e != cudaSuccess) return e;
    *cap = need;
    return cudaMemsetAsync(*ptr, 0, need, st);
  };
  CK(h, ensure((void**)&h->d_res, &h->res_cap, P.res_total));
  CK(h, ensure((void**)&h->d_stream, &h->stream_cap, P.stream_total));
  if (P.lnvec_floats) CK(h, ensure((void**)&h->d_lnvec, &h->lnvec_cap, P.lnvec_floats * sizeof(float)));
  tmx::PackArgs a{};
  const MlpDesc& d = h->dyn.desc;
  a.W1 = d.W[0]; a.b1 = d.b[0]; a.W2 = d.W[1]; a.b2 = d.b[1]; a.W3 = d.W[2]; a.b3 = d.b[2];
  a.ld1 = d.out_pad[0]; a.ld2 = d.out_pad[1]; a.ld3 = d.out_pad[2];
  if (P.layer_norm) { a.g1 = d.ln_gamma[0]; a.be1 = d.ln_beta[0]; a.g2 = d.ln_gamma[1]; a.be2 = d.ln_beta[1]; }
  if (P.reward) {
    const MlpDesc& r = h->rew.desc;
    a.Wr1 = r.W[0]; a.br1 = r.b[0]; a.Wr2 = r.W[1]; a.br2 = r.b[1];
    a.ldr1 = r.out_pad[0]; a.ldr2 = r.out_pad[1];
    if (P.layer_norm) { a.gr = r.ln_gamma[0]; a.ber = r.ln_beta[0]; }
  }
  a.explicit_bias = P.explicit_bias ? 1 : 0;
  a.S = S; a.s_pad = P.s_pad; a.nc = P.nc; a.cg = P.cg; a.kbs = P.lay.kbs; a.nstg = P.lay.nstg;
  a.blk_h = P.lay.blk_h; a.blk_o = P.lay.blk_o; a.w1_bytes = P.lay.w1_bytes; a.res_bytes = P.lay.res_bytes; a.stage_bytes = P.lay.stage_bytes;
  a.res = reinterpret_cast<uint16_t*>(h->d_res); a.stream = reinterpret_cast<uint16_t*>(h->d_stream); a.lnvec = h->d_lnvec;
  CK(h, tmx::tm_pack(a, st));
  h->launches += (P.layer_norm || P.reward || P.explicit_bias) ? 2 : 1;
  if (!h->trace_path.empty() && !h->d_trace)
    CK(h, cudaMalloc(&h->d_trace, sizeof(unsigned long long) * tmx::kTraceSteps * tmx::kTraceSlots));
  h->tm_ready = true;
  return BCMPC_OK;
}

int launch_tm(bcmpc_handle* h, const PlanArgs& p, cudaStream_t st) {
  if (h->deep) return launch_deep(h, p, st);
  if (!h->tm_ready) {
    if (h->cfg.activation == BCMPC_ACT_TANH && !h->rew.set && h->ln_on)
      return fail(h, BCMPC_ERR_UNSUPPORTED, "bf16 fast path: a plain tanh NNDynamicsModel with LayerNorm has no tensor-core variant in this mode (BCMPC_TC_CG=1, or "
                                            "together with a proposal network of more than 128 units); use BCMPC_PREC_FP32_VERIFY");
    return fail(h, BCMPC_ERR_STATE, "bf16 fast path not prepared: set the dynamics weights, and with a reward head either all three LayerNorms "
                                    "(bcmpc_set_layer_norm + bcmpc_set_reward_layer_norm) or none");
  }
  const tmx::TmHostPlan& P = h->tmp;
  tmx::TmArgs a;
  memset(&a, 0, sizeof a);
  a.p = p;
  a.res = h->d_res; a.stream = h->d_stream; a.trace = h->d_trace;
  a.hint_chain = h->tm_hint_chain;
  for (int d = 0; d < h->cfg.obs_dim; ++d) {
    a.mean_obs[d] = (float)h->norm.mean_obs[d];
    a.inv_den_obs[d] = (float)(1.0 / h->norm.den_obs[d]);
    a.mean_delta[d] = (float)h->norm.mean_delta[d];
    a.std_delta[d] = (float)h->norm.std_delta[d];
  }
  for (int j = 0; j < h->cfg.act_dim; ++j) {
    a.mean_act[j] = (float)h->norm.mean_act[j];
    a.inv_den_act[j] = (float)(1.0 / h->norm.den_act[j]);
    a.low[j] = h->norm.low[j];
    a.span[j] = h->norm.span[j];
  }
  const bool pol = p.proposal != BCMPC_PROP_UNIFORM || p.proposal0 != BCMPC_PROP_UNIFORM;
  const bool plain_tanh = h->cfg.activation == BCMPC_ACT_TANH && !P.reward;
  const int var = (P.reward ? tmx::kVarReward : 0) | (pol ? tmx::kVarPolicy : 0) | (P.layer_norm ? tmx::kVarLayerNorm : 0) |
                  (P.explicit_bias ? tmx::kVarBias : 0) | (plain_tanh ? tmx::kVarTanh : 0) | ((pol && h->pol_wide) ? tmx::kVarPolWide : 0);
  // (the streamed dynamics image is packed for the variant's stage size: set_policy re-packs it when the network's class changes)
  if (tmx::make_layout(h->tm_cg, P.nc, var).kbs != P.lay.kbs)
    return fail(h, BCMPC_ERR_STATE, "bf16 fast path: the weight images were packed for another kernel variant (set the proposal network before planning)");
  a.proposal = p.proposal; a.inj_kind = p.inj_kind;
  a.proposal0 = p.proposal0; a.inj_kind0 = p.inj_kind0;
  if (P.explicit_bias) a.lnvec = h->d_lnvec;
  if (P.reward) {
    a.lnvec = h->d_lnvec;
    if (P.layer_norm && fabsf(h->rew.desc.ln_eps - h->ln_eps) > 0.f)
      return fail(h, BCMPC_ERR_INVALID, "bf16 fast path: the three LayerNorms of the reward model must share one epsilon");
    a.mean_reward = (float)h->mean_reward; a.std_reward = (float)h->std_reward; a.gamma = (float)h->gamma;
  }
  if (P.layer_norm) {
    a.lnvec = h->d_lnvec;
    a.ln_inv_n = 1.0f / (float)h->cfg.hidden_size; a.ln_eps = h->ln_eps;
  }
  if (pol) { int rc = fill_policy_args(h, a); if (rc) return rc; }
  if (h->d_trace) CK(h, cudaMemsetAsync(h->d_trace, 0, sizeof(unsigned long long) * tmx::kTraceSteps * tmx::kTraceSlots, st));
  // tiles_per_state is even when tm_cg == 2 (run_plan pads it)
  if (h->tm_use_narrow) {
    const int64_t units = (int64_t)p.B * (p.tiles_per_state / 4);          // persistent CTA pairs, four tiles per unit
    const unsigned groups = (unsigned)std::min<int64_t>(units, h->sm_pairs);
    CK(h, tmx::nw_launch(P.nc, P.layer_norm, pol, a, dim3(groups * 2, 1), st));
  }
  if (h->tm_use_duo) {
    const int64_t units = (int64_t)p.B * (p.tiles_per_state / 2);          // persistent CTA pairs; each pair's units alternate between its two pipes
    const unsigned groups = (unsigned)std::min<int64_t>(units, h->sm_pairs);
    CK(h, tmx::du_launch(P.nc, var, a, dim3(groups * 2, 1), st));
  } else if (!h->tm_use_narrow) {
    // persistent CTAs: at most one wave of CTA groups, each rolling its share of the (root state, tile group) units
    const int64_t units = (int64_t)p.B * (p.tiles_per_state / h->tm_cg);
    const int64_t slots = h->tm_cg == 2 ? h->sm_pairs : 2 * h->sm_pairs;
    const unsigned groups = (unsigned)std::min<int64_t>(units, slots);
    CK(h, tmx::tm_launch(h->tm_cg, P.nc, var, a, dim3(groups * h->tm_cg, 1), st));
  }
  CK(h, cudaGetLastError());
  h->launches++;
  if (h->d_trace) {
    CK(h, cudaStreamSynchronize(st));
    std::vector<unsigned long long> t((size_t)tmx::kTraceSteps * tmx::kTraceSlots);
    CK(h, cudaMemcpy(t.data(), h->d_trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (h->tm_use_duo) {
      if (FILE* f = fopen(h->trace_path.c_str(), "wb")) {
        const int hdr[10] = {-2, tmx::kDuTrSteps, tmx::kDuTrSlots, P.nc, 0, 0, 0, 0, 0, 0};
        fwrite(hdr, sizeof hdr, 1, f);
        fwrite(t.data(), sizeof(unsigned long long), t.size(), f);
        fclose(f);
      }
    } else if (FILE* f = fopen(h->trace_path.c_str(), "wb")) {
      const int hdr[10] = {tmx::kTraceSteps, tmx::kTraceSlots, P.nc, P.lay.nstg, tmx::kTrL1, tmx::kTrL2, tmx::kTrL3, tmx::kTrEpi,
                           tmx::kTrFinal, tmx::kTrProd};
      fwrite(hdr, sizeof hdr, 1, f);
      fwrite(t.data(), sizeof(unsigned long long), t.size(), f);
      fclose(f);
    }
  }
  return BCMPC_OK;
}

// rollout (either precision) + per-state reduction of the tile candidates into d_out[B]
int run_plan(bcmpc_handle* h, PlanArgs p, bcmpc_record* d_out, cudaStream_t st, int force_verify) {
  const bool fast = h->cfg.precision == BCMPC_PREC_BF16_FAST && !force_verify;
  const int M = fast ? tmx::kRows : h->vM;
  p.tiles_per_state = (p.K + M - 1) / M;
  if (fast && (h->tm_cg == 2 || h->deep)) p.tiles_per_state = (p.tiles_per_state + 1) / 2 * 2;   // CTA pairs; a padding tile has no valid row
  // Hidden <= 256: two kernels read the same weight images.  rollout_narrow_kernel runs two tiles per CTA - ~1.45x the step
  // time for twice the paths - so it wins as soon as rollout_tmem_kernel would need more than one wave of CTA pairs; a
  // launch that fits in one wave (the single-state plans of the reference's env loop) is 25-30 % quicker on the wide kernel.
  h->tm_use_narrow = fast && !h->deep && h->tm_narrow && (!h->tm_wide_ok || (int64_t)p.B * (p.tiles_per_state / 2) > h->narrow_threshold);
  if (h->tm_use_narrow) p.tiles_per_state = (p.tiles_per_state + 3) / 4 * 4;        // two tiles per CTA of a pair
  // Hidden 257..510, uniform shooting: with more than one wave of CTA-pair units every pair rolls two tiles that take turns on
  // the hidden-1 buffer (rollout_duo_kernel: the idle stretches of one tile's step are the other's layer 2); same weight images
  h->tm_use_duo = fast && !h->deep && !h->tm_use_narrow && duo_applies(h, p) &&
                  (h->duo_mode == 2 || (int64_t)p.B * (p.tiles_per_state / 2) > h->sm_pairs);
  if (d_out) {
    int rc = ensure_cand(h, (size_t)p.B * p.tiles_per_state);
    if (rc) return rc;
    p.cand = h->d_cand;
  } else {
    p.cand = nullptr;
  }
  if (h->weights_pending) {     // an asynchronous weight hand-off on another stream: order this plan after it
    if (cudaEventQuery(h->ev_weights) == cudaSuccess) h->weights_pending = false;
    else CK(h, cudaStreamWaitEvent(st, h->ev_weights, 0));
  }
  // few tiles per root state: the rollout kernel's last tile of a state does the min-loc itself (no second launch)
  // (a latency measure for single plans - the reference plans one state per env step; large batches keep the separate kernel)
  const bool exchange = h->peer.world > 0 && d_out != nullptr;     // collective plan: winners over all ranks through peer memory
  const bool fused = fast && d_out != nullptr && p.tiles_per_state <= 64 && p.B <= 8 && !exchange;
  p.out = fused ? d_out : nullptr;
  p.ticket = h->d_ticket;
  if (h->timing) { CK(h, cudaEventRecord(h->ev0, st)); }
  int rc;
  if (fast) {
    rc = launch_tm(h, p, st);
  } else {
    VerifyArgs a;
    fill_verify_args(h, &a);
    a.p = p;
    dim3 grid(p.tiles_per_state, p.B);
    rc = h->vM == 32 ? launch_verify_t<32>(h, a, grid, st) : launch_verify_t<16>(h, a, grid, st);
  }
  if (rc) return rc;
  if (h->timing) { CK(h, cudaEventRecord(h->ev1, st)); h->ev_valid = true; }
  if (exchange) {
    // the exchange epoch is a launch argument that must advance with every plan: a captured graph would replay a stale one
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
      return fail(h, BCMPC_ERR_UNSUPPORTED, "a plan with the NVLink peer-store exchange cannot be captured into a CUDA graph (use BCMPC_EXCHANGE=nccl or capture per-rank plans)");
    h->peer.epoch++;
    peer_exchange_kernel<<<h->peer.world, 256, 0, st>>>(h->peer, h->d_cand, p.tiles_per_state, p.B, d_out);
    CK(h, cudaGetLastError());
    h->launches++;
  } else if (d_out && !fused) {
    reduce_candidates_kernel<<<p.B, 256, 0, st>>>(h->d_cand, p.tiles_per_state, p.B, 0, d_out);
    CK(h, cudaGetLastError());
    h->launches++;
  }
  return BCMPC_OK;
}

// Asynchronous (device-pointer) weight hand-offs are serialised through ONE event: a hand-off on stream B first waits for
// the previous hand-off's event (which may have been recorded on another stream A), so the latest record implies every
// earlier hand-off; plans on any stream wait on it (run_plan), and the handle's own stream - where bcmpc_policy_act,
// bcmpc_predict and the host uploads run - waits on it right away.
int begin_handoff(bcmpc_handle* h, bool device, cudaStream_t st) {
  if (device && h->weights_pending && h->ev_weights) CK(h, cudaStreamWaitEvent(st, h->ev_weights, 0));
  return BCMPC_OK;
}
// host (kind = cudaMemcpyHostToDevice, synchronous) and device (cudaMemcpyDeviceToDevice, enqueued on st) variants share these
int after_handoff(bcmpc_handle* h, bool device, cudaStream_t st) {
  if (!device) { CK(h, cudaStreamSynchronize(st)); return BCMPC_OK; }
  if (!h->ev_weights) CK(h, cudaEventCreateWithFlags(&h->ev_weights, cudaEventDisableTiming));
  CK(h, cudaEventRecord(h->ev_weights, st));
  h->weights_pending = true;
  if (st != h->stream) CK(h, cudaStreamWaitEvent(h->stream, h->ev_weights, 0));
  return BCMPC_OK;
}

PlanArgs base_plan_args(bcmpc_handle* h, int B, uint64_t seed) {
  PlanArgs p;
  memset(&p, 0, sizeof p);
  p.D = h->cfg.obs_dim; p.A = h->cfg.act_dim; p.K = h->cfg.num_paths; p.H = h->cfg.horizon; p.B = B;
  p.proposal = h->cfg.proposal;
  p.inj_kind = inj_kind_of(h->cfg.proposal);
  p.proposal0 = first_proposal_of(h->cfg);
  p.inj_kind0 = inj_kind_of(p.proposal0);
  p.path_offset = h->cfg.path_offset;
  p.seed = seed;
  return p;
}

}  // namespace

// =====================================================================================================
extern "C" {

int bcmpc_abi_version(void) { return BCMPC_ABI_VERSION; }

const char* bcmpc_last_error(const bcmpc_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int bcmpc_create(const bcmpc_config* cfg, bcmpc_handle** out) {
  if (!cfg || !out) return fail(nullptr, BCMPC_ERR_INVALID, "null argument");
  *out = nullptr;
  const bcmpc_config& c = *cfg;
  if (c.obs_dim < 1 || c.obs_dim > kMaxObs) return fail(nullptr, BCMPC_ERR_INVALID, "obs_dim %d out of [1,%d]", c.obs_dim, kMaxObs);
  if (c.act_dim < 1 || c.act_dim > kMaxAct) return fail(nullptr, BCMPC_ERR_INVALID, "act_dim %d out of [1,%d]", c.act_dim, kMaxAct);
  if (c.n_hidden < 1 || c.n_hidden + 1 > kMaxLayers) return fail(nullptr, BCMPC_ERR_INVALID, "n_hidden %d out of [1,%d]", c.n_hidden, kMaxLayers - 1);
  if (c.hidden_size < 1 || c.hidden_size > 1536) return fail(nullptr, BCMPC_ERR_INVALID, "hidden_size %d out of [1,1536]", c.hidden_size);
  if (c.num_paths < 1) return fail(nullptr, BCMPC_ERR_INVALID, "num_paths must be >= 1");
  if (c.horizon < 1) return fail(nullptr, BCMPC_ERR_INVALID, "horizon must be >= 1");
  if (c.max_states < 1) return fail(nullptr, BCMPC_ERR_INVALID, "max_states must be >= 1");
  if (c.activation != BCMPC_ACT_RELU && c.activation != BCMPC_ACT_TANH) return fail(nullptr, BCMPC_ERR_INVALID, "unknown activation %d", c.activation);
  if (c.precision != BCMPC_PREC_FP32_VERIFY && c.precision != BCMPC_PREC_BF16_FAST) return fail(nullptr, BCMPC_ERR_INVALID, "unknown precision %d", c.precision);
  if (c.proposal < BCMPC_PROP_UNIFORM || c.proposal > BCMPC_PROP_POLICY_MEAN) return fail(nullptr, BCMPC_ERR_INVALID, "unknown proposal %d", c.proposal);
  if (c.first_proposal < 0 || c.first_proposal > 1 + BCMPC_PROP_POLICY_MEAN) return fail(nullptr, BCMPC_ERR_INVALID, "unknown first_proposal %d", c.first_proposal);
  if (c.path_offset < 0 || c.path_offset + c.num_paths > 0xFFFFFFFFLL) return fail(nullptr, BCMPC_ERR_INVALID, "path_offset out of range");
  // bf16 fast mode: two hidden layers of up to 510 units run on rollout_tmem_kernel / rollout_narrow_kernel, everything
  // wider or deeper (up to 4 x 1024) on the cluster-split rollout_deep_kernel
  // (BCMPC_DEEP=1, a developer switch: also run two hidden layers of <= 510 units on the cluster-split kernel where it applies -
  // half the layer-2 work per SM and step, the latency experiment of DESIGN.md section 7)
  bool force_deep = false;
  if (const char* s = getenv("BCMPC_DEEP")) {
    std::string why;
    force_deep = atoi(s) == 1 && c.precision == BCMPC_PREC_BF16_FAST && dpx::deep_supported(c, &why);
  }
  const bool want_deep = c.precision == BCMPC_PREC_BF16_FAST && (c.n_hidden != 2 || c.hidden_size > 510 || force_deep);
  if (c.precision == BCMPC_PREC_BF16_FAST) {
    std::string why;
    if (want_deep ? !dpx::deep_supported(c, &why) : !tmx::tm_supported(c, &why)) return fail(nullptr, BCMPC_ERR_UNSUPPORTED, "%s", why.c_str());
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(nullptr, BCMPC_ERR_CUDA, "no CUDA device: the planner has no CPU fallback");
  if (c.device < 0 || c.device >= ndev) return fail(nullptr, BCMPC_ERR_INVALID, "device %d out of [0,%d)", c.device, ndev);
  bcmpc_handle* h = new bcmpc_handle();
  h->cfg = c;
  h->deep = want_deep;
  if (h->cfg.total_paths <= 0) h->cfg.total_paths = c.path_offset + c.num_paths;
  if (const char* s = getenv("BCMPC_TC_CG")) h->tm_cg = atoi(s) == 2 ? 2 : 1;
  if (const char* s = getenv("BCMPC_INLINE_STATE")) h->inline_state_ok = atoi(s) != 0;
  if (h->deep) h->tm_cg = 2;          // the cluster-split kernel always works in CTA pairs (its proposal-network image is the pair geometry)
  if (const char* s = getenv("BCMPC_TC_TRACE")) {
#ifdef BCMPC_TRACE
    h->trace_path = s;
#else
    delete h;
    return fail(nullptr, BCMPC_ERR_UNSUPPORTED, "BCMPC_TC_TRACE is set but this library was built without the pipeline trace: rebuild with "
                                                "BCMPC_NVCC_EXTRA=-DBCMPC_TRACE python -m bc_mpc_b200._build (developer builds only, ~6%% slower)");
#endif
  }
  if (const char* s = getenv("BCMPC_NARROW")) h->tm_narrow_mode = atoi(s);
  if (const char* s = getenv("BCMPC_DUO")) h->duo_mode = atoi(s);
  {
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, c.device) == cudaSuccess && prop.multiProcessorCount >= 2) h->sm_pairs = prop.multiProcessorCount / 2;
    h->narrow_threshold = h->tm_narrow_mode == 2 ? 0 : h->sm_pairs;      // BCMPC_NARROW=2: always the narrow kernel where it applies (tests)
  }
  if (const char* s = getenv("BCMPC_TC_HINT_CHAIN")) h->tm_hint_chain = (uint32_t)strtoul(s, nullptr, 0);
  DeviceGuard guard(c.device);        // the caller's current device is restored on return
  cudaError_t e = guard.err;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_states, sizeof(double) * c.max_states * c.obs_dim);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_records, sizeof(bcmpc_record) * c.max_states);
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_states, sizeof(double) * c.max_states * c.obs_dim);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_ticket, sizeof(unsigned int) * c.max_states);
  if (e == cudaSuccess) e = cudaMemset(h->d_ticket, 0, sizeof(unsigned int) * c.max_states);
  if (e == cudaSuccess) e = cudaHostAlloc(&h->h_records, sizeof(bcmpc_record) * c.max_states, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&h->h_records_dev, h->h_records, 0);
  if (e != cudaSuccess) {
    int rc = fail(nullptr, BCMPC_ERR_CUDA, "bcmpc_create: %s", cudaGetErrorString(e));
    bcmpc_destroy(h);
    return rc;
  }
  // default action box [-1, 1) (HalfCheetah, controllers.py:53); std_act etc. must be set explicitly
  for (int j = 0; j < kMaxAct; ++j) { h->norm.low[j] = -1.f; h->norm.span[j] = 2.f; }
  *out = h;
  return BCMPC_OK;
}

int bcmpc_destroy(bcmpc_handle* h) {
  if (!h) return BCMPC_OK;
  DeviceGuard guard(h->cfg.device);
  cudaDeviceSynchronize();
  h->dyn.release();
  h->pol.release();
  h->rew.release();
  for (float* p : h->ln_bufs) cudaFree(p);
  cudaFree(h->rew_ln_buf);
  cudaFree(h->d_ticket);
  cudaFree(h->d_cand); cudaFree(h->d_states); cudaFree(h->d_records); cudaFree(h->d_inj);
  cudaFree(h->d_res); cudaFree(h->d_stream); cudaFree(h->d_pol); cudaFree(h->d_polvec); cudaFree(h->d_lnvec); cudaFree(h->d_trace);
  if (h->h_states) cudaFreeHost(h->h_states);
  if (h->h_records) cudaFreeHost(h->h_records);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_weights) cudaEventDestroy(h->ev_weights);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return BCMPC_OK;
}

namespace {
int set_dynamics_common(bcmpc_handle* h, const float* const* W, const float* const* b, int32_t n_layers, int kind, cudaStream_t st) {
  if (!h || !W || !b) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n_layers != h->cfg.n_hidden + 1) return fail(h, BCMPC_ERR_INVALID, "n_layers %d != n_hidden+1 = %d", n_layers, h->cfg.n_hidden + 1);
  for (int l = 0; l < n_layers; ++l)
    if (!W[l] || !b[l]) return fail(h, BCMPC_ERR_INVALID, "null weight pointer for layer %d", l);
  ON_DEVICE(h);
  int rc = refresh_dynamics(h, W, b, kind, st);
  if (rc) return rc;
  update_verify_geometry(h);
  if (h->cfg.precision == BCMPC_PREC_BF16_FAST) rc = tm_prepare(h, st);
  return rc;
}
}  // namespace

int bcmpc_set_dynamics(bcmpc_handle* h, const float* const* W, const float* const* b, int32_t n_layers) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  ON_DEVICE(h);
  CK(h, cudaStreamSynchronize(h->stream));
  int rc = set_dynamics_common(h, W, b, n_layers, cudaMemcpyHostToDevice, h->stream);
  if (rc) return rc;
  CK(h, cudaStreamSynchronize(h->stream));     // the caller may reuse its host arrays; plans on any stream see the new weights
  return BCMPC_OK;
}

int bcmpc_set_dynamics_device(bcmpc_handle* h, const float* const* dW, const float* const* db, int32_t n_layers, int32_t layout,
                              void* stream) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (layout != BCMPC_LAYOUT_IN_OUT && layout != BCMPC_LAYOUT_OUT_IN) return fail(h, BCMPC_ERR_INVALID, "unknown layout %d", layout);
  ON_DEVICE(h);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  int rc = begin_handoff(h, true, st);
  if (rc) return rc;
  rc = set_dynamics_common(h, dW, db, n_layers, layout == BCMPC_LAYOUT_OUT_IN ? -1 : (int)cudaMemcpyDeviceToDevice, st);
  if (rc) return rc;
  return after_handoff(h, true, st);
}

namespace {
int set_layer_norm_common(bcmpc_handle* h, const float* const* gamma, const float* const* beta, int32_t n_hidden, float eps, bool device,
                          cudaStream_t st) {
  if (!h->dyn.set) return fail(h, BCMPC_ERR_STATE, "dynamics weights not set (bcmpc_set_dynamics) - set them before the LayerNorm parameters");
  const bool fast = h->cfg.precision == BCMPC_PREC_BF16_FAST;
  if (gamma && !beta) return fail(h, BCMPC_ERR_INVALID, "gamma and beta must be given together");
  if (gamma && n_hidden != h->cfg.n_hidden) return fail(h, BCMPC_ERR_INVALID, "n_hidden %d != %d", n_hidden, h->cfg.n_hidden);
  for (int l = 0; l < n_hidden && gamma; ++l)
    if (!gamma[l] || !beta[l]) return fail(h, BCMPC_ERR_INVALID, "null LayerNorm pointer for layer %d", l);
  ON_DEVICE(h);
  if (!device) CK(h, cudaDeviceSynchronize());      // plans and hand-offs on any stream have finished with the old parameters
  { int rc = begin_handoff(h, device, st); if (rc) return rc; }
  const size_t n = (size_t)h->cfg.hidden_size;
  if (gamma && h->ln_bufs.empty())                   // persistent: [gamma, beta] per hidden layer
    for (int l = 0; l < 2 * h->cfg.n_hidden; ++l) {
      float* p = nullptr;
      CK(h, cudaMalloc(&p, sizeof(float) * n));
      h->ln_bufs.push_back(p);
    }
  for (int l = 0; l < kMaxLayers; ++l) h->dyn.desc.ln_gamma[l] = h->dyn.desc.ln_beta[l] = nullptr;
  h->ln_on = gamma != nullptr;
  h->ln_eps = eps;
  const cudaMemcpyKind kind = device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (int l = 0; l < n_hidden && gamma; ++l) {
    CK(h, cudaMemcpyAsync(h->ln_bufs[2 * l], gamma[l], sizeof(float) * n, kind, st));
    CK(h, cudaMemcpyAsync(h->ln_bufs[2 * l + 1], beta[l], sizeof(float) * n, kind, st));
    h->dyn.desc.ln_gamma[l] = h->ln_bufs[2 * l];
    h->dyn.desc.ln_beta[l] = h->ln_bufs[2 * l + 1];
  }
  h->dyn.desc.ln_eps = eps;
  if (fast) {
    int rc = tm_prepare(h, st);
    if (rc) return rc;
  }
  return after_handoff(h, device, st);
}

int set_reward_head_common(bcmpc_handle* h, const float* const* W, const float* const* b, double mean_reward, double std_reward,
                           double gamma, int kind, bool device, cudaStream_t st) {
  const bool fast = h->cfg.precision == BCMPC_PREC_BF16_FAST;
  if (fast && !h->dyn.set) return fail(h, BCMPC_ERR_STATE, "dynamics weights not set (bcmpc_set_dynamics) - set them before the reward head");
  ON_DEVICE(h);
  if (!device) CK(h, cudaDeviceSynchronize());
  { int rc = begin_handoff(h, device, st); if (rc) return rc; }
  if (!W) {
    CK(h, cudaDeviceSynchronize());
    h->rew.release();
    h->rew_ln_on = false;
    update_verify_geometry(h);
    if (fast) { int rc = tm_prepare(h, st); if (rc) return rc; }
    return after_handoff(h, device, st);
  }
  if (!b || !W[0] || !W[1] || !b[0] || !b[1]) return fail(h, BCMPC_ERR_INVALID, "null reward-head pointer");
  if (fast && h->deep) return fail(h, BCMPC_ERR_UNSUPPORTED, "bf16 fast path: the learned-reward model runs on rollout_tmem_kernel (n_hidden = 2, hidden_size <= 510)");
  if (fast && h->cfg.activation != BCMPC_ACT_TANH)
    return fail(h, BCMPC_ERR_UNSUPPORTED, "bf16 fast path: the learned-reward model is built with tanh layers (dynamics.py:162); other activations run in BCMPC_PREC_FP32_VERIFY");
  if (h->cfg.n_hidden != 2) return fail(h, BCMPC_ERR_INVALID, "NNDynamicsRewardModel has a shared layer and one hidden layer per head: n_hidden must be 2");
  const int sizes[3] = {h->cfg.hidden_size, h->cfg.hidden_size, 1};
  int rc = refresh_mlp(h, h->rew, 2, sizes, W, b, kind, st);
  if (rc) return rc;
  h->rew_ln_on = false;        // (a new head resets its LayerNorm, like new dynamics weights do)
  h->mean_reward = mean_reward; h->std_reward = std_reward; h->gamma = gamma;
  update_verify_geometry(h);
  if (fast) {
    rc = tm_prepare(h, st);
    if (rc) { cudaDeviceSynchronize(); h->rew.release(); update_verify_geometry(h); return rc; }
  }
  return after_handoff(h, device, st);
}

int set_reward_layer_norm_common(bcmpc_handle* h, const float* gamma, const float* beta, float eps, bool device, cudaStream_t st) {
  if (!h->rew.set) return fail(h, BCMPC_ERR_STATE, "reward head not set (bcmpc_set_reward_head)");
  if (gamma && !beta) return fail(h, BCMPC_ERR_INVALID, "gamma and beta must be given together");
  const bool fast = h->cfg.precision == BCMPC_PREC_BF16_FAST;
  ON_DEVICE(h);
  if (!device) CK(h, cudaDeviceSynchronize());
  { int rc = begin_handoff(h, device, st); if (rc) return rc; }
  const size_t n = (size_t)h->cfg.hidden_size;
  h->rew.desc.ln_gamma[0] = h->rew.desc.ln_beta[0] = nullptr;
  h->rew_ln_on = gamma != nullptr;
  if (gamma) {
    const cudaMemcpyKind kind = device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (!h->rew_ln_buf) CK(h, cudaMalloc(&h->rew_ln_buf, 2 * n * sizeof(float)));
    CK(h, cudaMemcpyAsync(h->rew_ln_buf, gamma, n * sizeof(float), kind, st));
    CK(h, cudaMemcpyAsync(h->rew_ln_buf + n, beta, n * sizeof(float), kind, st));
    h->rew.desc.ln_gamma[0] = h->rew_ln_buf;
    h->rew.desc.ln_beta[0] = h->rew_ln_buf + n;
    h->rew.desc.ln_eps = eps;
  }
  if (fast) {
    int rc = tm_prepare(h, st);
    if (rc) return rc;
  }
  return after_handoff(h, device, st);
}
cudaStream_t pick_stream(bcmpc_handle* h, void* stream) { return stream ? (cudaStream_t)stream : h->stream; }
}  // namespace

int bcmpc_set_layer_norm(bcmpc_handle* h, const float* const* gamma, const float* const* beta, int32_t n_hidden, float eps) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  return set_layer_norm_common(h, gamma, beta, n_hidden, eps, false, h->stream);
}
int bcmpc_set_layer_norm_device(bcmpc_handle* h, const float* const* d_gamma, const float* const* d_beta, int32_t n_hidden, float eps,
                                void* stream) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  return set_layer_norm_common(h, d_gamma, d_beta, n_hidden, eps, true, pick_stream(h, stream));
}

int bcmpc_set_reward_head(bcmpc_handle* h, const float* const* W, const float* const* b, double mean_reward,
                          double std_reward, double gamma) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  return set_reward_head_common(h, W, b, mean_reward, std_reward, gamma, cudaMemcpyHostToDevice, false, h->stream);
}
int bcmpc_set_reward_head_device(bcmpc_handle* h, const float* const* dW, const float* const* db, int32_t layout, double mean_reward,
                                 double std_reward, double gamma, void* stream) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (layout != BCMPC_LAYOUT_IN_OUT && layout != BCMPC_LAYOUT_OUT_IN) return fail(h, BCMPC_ERR_INVALID, "unknown layout %d", layout);
  return set_reward_head_common(h, dW, db, mean_reward, std_reward, gamma, layout == BCMPC_LAYOUT_OUT_IN ? -1 : (int)cudaMemcpyDeviceToDevice,
                                true, pick_stream(h, stream));
}

int bcmpc_set_reward_layer_norm(bcmpc_handle* h, const float* gamma, const float* beta, float eps) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  return set_reward_layer_norm_common(h, gamma, beta, eps, false, h->stream);
}
int bcmpc_set_reward_layer_norm_device(bcmpc_handle* h, const float* d_gamma, const float* d_beta, float eps, void* stream) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  return set_reward_layer_norm_common(h, d_gamma, d_beta, eps, true, pick_stream(h, stream));
}

int bcmpc_set_normalization(bcmpc_handle* h, const double* mean_obs, const double* std_obs, const double* mean_act,
                            const double* std_act, const double* mean_delta, const double* std_delta) {
  if (!h || !mean_obs || !std_obs || !mean_act || !std_act || !mean_delta || !std_delta)
    return fail(h, BCMPC_ERR_INVALID, "null argument");
  for (int d = 0; d < h->cfg.obs_dim; ++d) {
    h->norm.mean_obs[d] = mean_obs[d];
    h->norm.den_obs[d] = std_obs[d] + 1e-10;     // dynamics.py:109
    h->norm.mean_delta[d] = mean_delta[d];
    h->norm.std_delta[d] = std_delta[d];         // dynamics.py:113 (no epsilon)
  }
  for (int j = 0; j < h->cfg.act_dim; ++j) {
    h->norm.mean_act[j] = mean_act[j];
    h->norm.den_act[j] = std_act[j] + 1e-10;     // dynamics.py:110
  }
  h->have_norm = true;
  return BCMPC_OK;
}

int bcmpc_set_action_bounds(bcmpc_handle* h, const double* low, const double* high) {
  if (!h || !low || !high) return fail(h, BCMPC_ERR_INVALID, "null argument");
  for (int j = 0; j < h->cfg.act_dim; ++j) {
    if (!(high[j] >= low[j])) return fail(h, BCMPC_ERR_INVALID, "action bound %d: high < low", j);
    h->norm.low[j] = (float)low[j];
    h->norm.span[j] = (float)(high[j] - low[j]);
  }
  h->have_bounds = true;
  return BCMPC_OK;
}

namespace {
int set_policy_common(bcmpc_handle* h, int32_t n_layers, const int32_t* sizes, const float* const* W, const float* const* b,
                      const float* logstd, const float* filt_mean, const float* filt_std, float clip, int kind, cudaStream_t st) {
  if (!h || !sizes || !W || !b) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n_layers < 1 || n_layers > kMaxLayers) return fail(h, BCMPC_ERR_INVALID, "n_layers %d out of [1,%d]", n_layers, kMaxLayers);
  if (sizes[0] != h->cfg.obs_dim || sizes[n_layers] != h->cfg.act_dim)
    return fail(h, BCMPC_ERR_INVALID, "policy sizes must start with obs_dim and end with act_dim");
  for (int l = 0; l <= n_layers; ++l)
    if (sizes[l] < 1 || sizes[l] > 1536) return fail(h, BCMPC_ERR_INVALID, "policy layer size %d out of range", sizes[l]);
  for (int l = 0; l < n_layers; ++l)
    if (!W[l] || !b[l]) return fail(h, BCMPC_ERR_INVALID, "null weight pointer for policy layer %d", l);
  if ((filt_mean == nullptr) != (filt_std == nullptr)) return fail(h, BCMPC_ERR_INVALID, "filt_mean and filt_std must be given together");
  const bool fast = h->cfg.precision == BCMPC_PREC_BF16_FAST && needs_policy(h->cfg);
  if (fast) {
    std::string why;
    if (!tmx::tm_policy_supported(n_layers, sizes, &why)) return fail(h, BCMPC_ERR_UNSUPPORTED, "%s", why.c_str());
  }
  ON_DEVICE(h);
  int rc = refresh_mlp(h, h->pol, n_layers, sizes, W, b, kind, st);
  if (rc) return rc;
  h->has_filter = filt_mean != nullptr;
  for (int d = 0; d < h->cfg.obs_dim; ++d) {
    h->filt_mean[d] = filt_mean ? filt_mean[d] : 0.f;
    h->filt_std[d] = filt_std ? filt_std[d] : 1.f;
  }
  h->clip = clip;
  h->has_logstd = logstd != nullptr;
  for (int j = 0; j < h->cfg.act_dim; ++j) h->pol_std[j] = logstd ? expf(logstd[j]) : 0.f;   // distributions.py:65
  update_verify_geometry(h);
  if (fast) {
    // the image layout depends only on the CTA-group size: it can be packed before the dynamics weights arrive
    // (BCMPC_POL_WIDE=1, a developer switch: run a <= 128-unit network on the wide variants - their cost is the smaller weight ring)
    const char* force_wide = getenv("BCMPC_POL_WIDE");
    const bool wide = tmx::tm_policy_is_wide(sizes) || (force_wide && atoi(force_wide) == 1 && h->tm_cg == 2);
    if (wide && h->tm_cg != 2) return fail(h, BCMPC_ERR_UNSUPPORTED, "bf16 fast path: proposal networks wider than 128 units need CTA pairs (BCMPC_TC_CG=2, the default)");
    if (wide && h->deep) return fail(h, BCMPC_ERR_UNSUPPORTED, "bf16 fast path: the cluster-split kernel (n_hidden > 2 or hidden_size > 510) holds proposal networks of at most 128 units per hidden layer");
    const int pvar = tmx::kVarPolicy | (wide ? tmx::kVarPolWide : 0);
    const tmx::Layout lay = tmx::make_layout(h->tm_cg, 1, pvar);
    const size_t img_total = (size_t)h->tm_cg * lay.pol_img_bytes;
    if (img_total > h->pol_cap) {
      if (h->d_pol) { CK(h, cudaDeviceSynchronize()); CK(h, cudaFree(h->d_pol)); h->d_pol = nullptr; h->pol_cap = 0; }
      CK(h, cudaMalloc(&h->d_pol, img_total));
      h->pol_cap = img_total;
    }
    if (!h->d_polvec) CK(h, cudaMalloc(&h->d_polvec, tmx::kPolWideVecFloats * sizeof(float)));
    tmx::PolicyPackArgs a{};
    const MlpDesc& d = h->pol.desc;
    for (int l = 0; l < 3; ++l) { a.W[l] = d.W[l]; a.b[l] = d.b[l]; a.ld[l] = d.out_pad[l]; }
    a.h1 = sizes[1]; a.h2 = sizes[2]; a.cg = h->tm_cg; a.pn = tmx::pol_chunks(pvar);
    a.blk_h = lay.blk_h; a.blk_o = lay.blk_o; a.img_bytes = lay.pol_img_bytes;
    a.img = reinterpret_cast<uint16_t*>(h->d_pol); a.vec = h->d_polvec;
    CK(h, tmx::tm_pack_policy(a, st));
    h->launches++;
    h->tmp.with_policy = true;
    h->tmp.np1 = (sizes[1] + 63) / 64; h->tmp.np2 = (sizes[2] + 63) / 64;
    if (wide != h->pol_wide) {
      // the network changed class: another kernel variant, whose weight ring has another stage size - re-pack the dynamics images
      h->pol_wide = wide;
      if (h->dyn.set) { rc = tm_prepare(h, st); if (rc) return rc; }
    }
  }
  return BCMPC_OK;
}
}  // namespace

int bcmpc_set_policy(bcmpc_handle* h, int32_t n_layers, const int32_t* sizes, const float* const* W,
                     const float* const* b, const float* logstd, const float* filt_mean, const float* filt_std,
                     float clip) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  ON_DEVICE(h);
  CK(h, cudaStreamSynchronize(h->stream));
  int rc = set_policy_common(h, n_layers, sizes, W, b, logstd, filt_mean, filt_std, clip, cudaMemcpyHostToDevice, h->stream);
  if (rc) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return BCMPC_OK;
}

int bcmpc_set_policy_device(bcmpc_handle* h, int32_t n_layers, const int32_t* sizes, const float* const* dW, const float* const* db,
                            int32_t layout, const float* logstd, const float* filt_mean, const float* filt_std, float clip,
                            void* stream) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (layout != BCMPC_LAYOUT_IN_OUT && layout != BCMPC_LAYOUT_OUT_IN) return fail(h, BCMPC_ERR_INVALID, "unknown layout %d", layout);
  ON_DEVICE(h);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  int rc = begin_handoff(h, true, st);
  if (rc) return rc;
  rc = set_policy_common(h, n_layers, sizes, dW, db, logstd, filt_mean, filt_std, clip,
                         layout == BCMPC_LAYOUT_OUT_IN ? -1 : (int)cudaMemcpyDeviceToDevice, st);
  if (rc) return rc;
  return after_handoff(h, true, st);
}

namespace {
int plan_device_common(bcmpc_handle* h, const double* d_states, int32_t n_states, const double* d_inj, uint64_t seed,
                       const uint64_t* d_seed, bcmpc_record* d_out, void* stream, const double* host_state = nullptr) {
  if (!h || !d_states || !d_out) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n_states < 1 || n_states > h->cfg.max_states) return fail(h, BCMPC_ERR_INVALID, "n_states %d out of [1,%d]", n_states, h->cfg.max_states);
  int rc = check_ready(h, needs_policy(h->cfg));
  if (rc) return rc;
  if (!h->rew.set && h->cfg.obs_dim < 18) return fail(h, BCMPC_ERR_INVALID, "cheetah_cost_fn indexes obs[17]; obs_dim = %d (set a reward head to plan on a learned reward)", h->cfg.obs_dim);
  ON_DEVICE(h);
  PlanArgs p = base_plan_args(h, n_states, seed);
  p.seed_ptr = reinterpret_cast<const unsigned long long*>(d_seed);
  p.states = d_states;
  p.inj = d_inj;
  if (host_state) {      // one root state in the launch arguments (bcmpc_plan with n_states == 1)
    p.inline_state = 1;
    for (int d = 0; d < h->cfg.obs_dim; ++d) p.state0[d] = host_state[d];
  }
  return run_plan(h, p, d_out, stream ? (cudaStream_t)stream : h->stream, 0);
}
}  // namespace

int bcmpc_plan_device(bcmpc_handle* h, const double* d_states, int32_t n_states, const double* d_inj, uint64_t seed,
                      bcmpc_record* d_out, void* stream) {
  return plan_device_common(h, d_states, n_states, d_inj, seed, nullptr, d_out, stream);
}

int bcmpc_plan_device_seeded(bcmpc_handle* h, const double* d_states, int32_t n_states, const double* d_inj,
                             const uint64_t* d_seed, bcmpc_record* d_out, void* stream) {
  if (!d_seed) return fail(h, BCMPC_ERR_INVALID, "null argument");
  return plan_device_common(h, d_states, n_states, d_inj, 0, d_seed, d_out, stream);
}

int bcmpc_plan(bcmpc_handle* h, const double* states, int32_t n_states, const double* inj, uint64_t seed,
               bcmpc_record* out) {
  if (!h || !states || !out) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n_states < 1 || n_states > h->cfg.max_states) return fail(h, BCMPC_ERR_INVALID, "n_states %d out of [1,%d]", n_states, h->cfg.max_states);
  ON_DEVICE(h);
  // a single plan carries its root state in the launch arguments (BCMPC_INLINE_STATE=0 restores the copy: A/B in tools/latency_breakdown.py)
  const bool inline_state = n_states == 1 && h->inline_state_ok;
  if (!inline_state) {
    const size_t sbytes = sizeof(double) * n_states * h->cfg.obs_dim;
    memcpy(h->h_states, states, sbytes);
    CK(h, cudaMemcpyAsync(h->d_states, h->h_states, sbytes, cudaMemcpyHostToDevice, h->stream));
  }
  const double* d_inj = nullptr;
  if (inj) {
    const size_t n = (size_t)n_states * h->cfg.horizon * h->cfg.num_paths * h->cfg.act_dim;
    if (n > h->inj_cap) {
      if (h->d_inj) cudaFree(h->d_inj);
      h->d_inj = nullptr; h->inj_cap = 0;
      CK(h, cudaMalloc(&h->d_inj, n * sizeof(double)));
      h->inj_cap = n;
    }
    CK(h, cudaMemcpyAsync(h->d_inj, inj, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    d_inj = h->d_inj;
  }
  // up to 64 records (5 KB) are written by the min-loc kernel straight into mapped host memory: one copy call and one DMA
  // round trip less on the single-plan latency path (the reference plans one state per env step)
  const bool direct = n_states <= 64 && h->h_records_dev != nullptr;
  int rc = plan_device_common(h, h->d_states, n_states, d_inj, seed, nullptr, direct ? h->h_records_dev : h->d_records, h->stream,
                              inline_state ? states : nullptr);
  if (rc) return rc;
  if (!direct) CK(h, cudaMemcpyAsync(h->h_records, h->d_records, sizeof(bcmpc_record) * n_states, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  memcpy(out, h->h_records, sizeof(bcmpc_record) * n_states);
  return BCMPC_OK;
}

int bcmpc_rollout_debug(bcmpc_handle* h, const double* state, const double* inj, uint64_t seed, double* states_out,
                        double* actions_out, double* costs_out) {
  if (!h || !state) return fail(h, BCMPC_ERR_INVALID, "null argument");
  int rc = check_ready(h, needs_policy(h->cfg));
  if (rc) return rc;
  ON_DEVICE(h);
  const int D = h->cfg.obs_dim, A = h->cfg.act_dim, K = h->cfg.num_paths, H = h->cfg.horizon;
  double *d_state = nullptr, *d_inj = nullptr, *d_so = nullptr, *d_ao = nullptr, *d_co = nullptr;
  auto cleanup = [&]() { cudaFree(d_state); cudaFree(d_inj); cudaFree(d_so); cudaFree(d_ao); cudaFree(d_co); };
#define CKC(expr)                                                                                      \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) { cleanup(); return fail(h, BCMPC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); } \
  } while (0)
  CKC(cudaMalloc(&d_state, sizeof(double) * D));
  CKC(cudaMemcpy(d_state, state, sizeof(double) * D, cudaMemcpyHostToDevice));
  if (inj) {
    CKC(cudaMalloc(&d_inj, sizeof(double) * H * K * A));
    CKC(cudaMemcpy(d_inj, inj, sizeof(double) * H * K * A, cudaMemcpyHostToDevice));
  }
  if (states_out) CKC(cudaMalloc(&d_so, sizeof(double) * (H + 1) * K * D));
  if (actions_out) CKC(cudaMalloc(&d_ao, sizeof(double) * H * K * A));
  if (costs_out) CKC(cudaMalloc(&d_co, sizeof(double) * K));
  PlanArgs p = base_plan_args(h, 1, seed);
  p.states = d_state; p.inj = d_inj; p.states_out = d_so; p.actions_out = d_ao; p.costs_out = d_co;
  rc = run_plan(h, p, nullptr, h->stream, 0);
  if (rc) { cleanup(); return rc; }
  CKC(cudaStreamSynchronize(h->stream));
  if (states_out) CKC(cudaMemcpy(states_out, d_so, sizeof(double) * (H + 1) * K * D, cudaMemcpyDeviceToHost));
  if (actions_out) CKC(cudaMemcpy(actions_out, d_ao, sizeof(double) * H * K * A, cudaMemcpyDeviceToHost));
  if (costs_out) CKC(cudaMemcpy(costs_out, d_co, sizeof(double) * K, cudaMemcpyDeviceToHost));
  cleanup();
  return BCMPC_OK;
}

int bcmpc_predict(bcmpc_handle* h, const double* states, const double* actions, int64_t n, double* next_states) {
  if (!h || !states || !actions || !next_states) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n < 1 || n > 0x7FFFFFFF) return fail(h, BCMPC_ERR_INVALID, "row count out of range");
  int rc = check_ready(h, false);
  if (rc) return rc;
  ON_DEVICE(h);
  const int D = h->cfg.obs_dim, A = h->cfg.act_dim;
  double *d_s = nullptr, *d_a = nullptr, *d_o = nullptr;
  auto cleanup = [&]() { cudaFree(d_s); cudaFree(d_a); cudaFree(d_o); };
  CKC(cudaMalloc(&d_s, sizeof(double) * n * D));
  CKC(cudaMalloc(&d_a, sizeof(double) * n * A));
  CKC(cudaMalloc(&d_o, sizeof(double) * 2 * n * D));
  CKC(cudaMemcpy(d_s, states, sizeof(double) * n * D, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(d_a, actions, sizeof(double) * n * A, cudaMemcpyHostToDevice));
  // one step, every row from its own state with its own action: the predict of dynamics.py:106-119
  PlanArgs p = base_plan_args(h, 1, 0);
  p.K = (int)n; p.H = 1; p.proposal = p.proposal0 = BCMPC_PROP_UNIFORM; p.inj_kind = p.inj_kind0 = kInjUniformBox;
  p.states = d_s; p.inj = d_a; p.states_out = d_o; p.per_row_states = 1;
  rc = run_plan(h, p, nullptr, h->stream, /*force_verify=*/h->cfg.precision != BCMPC_PREC_BF16_FAST);
  if (rc) { cleanup(); return rc; }
  CKC(cudaStreamSynchronize(h->stream));
  CKC(cudaMemcpy(next_states, d_o + n * D, sizeof(double) * n * D, cudaMemcpyDeviceToHost));
  cleanup();
  return BCMPC_OK;
}

int bcmpc_predict_reward(bcmpc_handle* h, const double* states, const double* actions, int64_t n, double* next_states,
                         double* rewards) {
  if (!h || !states || !actions || !next_states || !rewards) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n < 1 || n > 0x7FFFFFFF) return fail(h, BCMPC_ERR_INVALID, "row count out of range");
  if (!h->rew.set) return fail(h, BCMPC_ERR_STATE, "reward head not set (bcmpc_set_reward_head)");
  int rc = check_ready(h, false);
  if (rc) return rc;
  ON_DEVICE(h);
  const int D = h->cfg.obs_dim, A = h->cfg.act_dim;
  double *d_s = nullptr, *d_a = nullptr, *d_o = nullptr, *d_c = nullptr;
  auto cleanup = [&]() { cudaFree(d_s); cudaFree(d_a); cudaFree(d_o); cudaFree(d_c); };
  CKC(cudaMalloc(&d_s, sizeof(double) * n * D));
  CKC(cudaMalloc(&d_a, sizeof(double) * n * A));
  CKC(cudaMalloc(&d_o, sizeof(double) * 2 * n * D));
  CKC(cudaMalloc(&d_c, sizeof(double) * n));
  CKC(cudaMemcpy(d_s, states, sizeof(double) * n * D, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(d_a, actions, sizeof(double) * n * A, cudaMemcpyHostToDevice));
  // one step, every row from its own state with its own action; the step objective is -reward * gamma^0
  PlanArgs p = base_plan_args(h, 1, 0);
  p.K = (int)n; p.H = 1; p.proposal = p.proposal0 = BCMPC_PROP_UNIFORM; p.inj_kind = p.inj_kind0 = kInjUniformBox;
  p.states = d_s; p.inj = d_a; p.states_out = d_o; p.costs_out = d_c; p.per_row_states = 1;
  rc = run_plan(h, p, nullptr, h->stream, /*force_verify=*/1);
  if (rc) { cleanup(); return rc; }
  CKC(cudaStreamSynchronize(h->stream));
  CKC(cudaMemcpy(next_states, d_o + n * D, sizeof(double) * n * D, cudaMemcpyDeviceToHost));
  CKC(cudaMemcpy(rewards, d_c, sizeof(double) * n, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) rewards[i] = -rewards[i];
  cleanup();
  return BCMPC_OK;
}

int bcmpc_policy_act(bcmpc_handle* h, const double* obs, const float* noise, int64_t n, float* out) {
  if (!h || !obs || !out) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n < 1 || n > 0x7FFFFFFF) return fail(h, BCMPC_ERR_INVALID, "row count out of range");
  if (!h->pol.set) return fail(h, BCMPC_ERR_STATE, "proposal network not set (bcmpc_set_policy)");
  if (noise && !h->has_logstd) return fail(h, BCMPC_ERR_STATE, "stochastic act needs logstd in bcmpc_set_policy");
  ON_DEVICE(h);
  const int D = h->cfg.obs_dim, A = h->cfg.act_dim;
  double* d_obs = nullptr;
  float *d_noise = nullptr, *d_out = nullptr;
  auto cleanup = [&]() { cudaFree(d_obs); cudaFree(d_noise); cudaFree(d_out); };
  CKC(cudaMalloc(&d_obs, sizeof(double) * n * D));
  CKC(cudaMalloc(&d_out, sizeof(float) * n * A));
  CKC(cudaMemcpy(d_obs, obs, sizeof(double) * n * D, cudaMemcpyHostToDevice));
  if (noise) {
    CKC(cudaMalloc(&d_noise, sizeof(float) * n * A));
    CKC(cudaMemcpy(d_noise, noise, sizeof(float) * n * A, cudaMemcpyHostToDevice));
  }
  VerifyArgs a;
  fill_verify_args(h, &a);
  a.p = base_plan_args(h, 1, 0);
  const int M = h->vM;
  const size_t smem = 2 * (size_t)M * a.ld * sizeof(float);
  const unsigned grid = (unsigned)((n + M - 1) / M);
  if (M == 32) {
    CKC(cudaFuncSetAttribute(policy_act_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    policy_act_kernel<32><<<grid, kVerifyThreads, smem, h->stream>>>(a, d_obs, d_noise, n, d_out);
  } else {
    CKC(cudaFuncSetAttribute(policy_act_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    policy_act_kernel<16><<<grid, kVerifyThreads, smem, h->stream>>>(a, d_obs, d_noise, n, d_out);
  }
  CKC(cudaGetLastError());
  h->launches++;
  CKC(cudaStreamSynchronize(h->stream));
  CKC(cudaMemcpy(out, d_out, sizeof(float) * n * A, cudaMemcpyDeviceToHost));
  cleanup();
  return BCMPC_OK;
}

int bcmpc_cheetah_cost(bcmpc_handle* h, const double* states, const double* next_states, int64_t n, double* out) {
  if (!h || !states || !next_states || !out) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n < 1) return fail(h, BCMPC_ERR_INVALID, "row count out of range");
  if (h->cfg.obs_dim < 18) return fail(h, BCMPC_ERR_INVALID, "cheetah_cost_fn indexes obs[17]; obs_dim = %d", h->cfg.obs_dim);
  ON_DEVICE(h);
  const int D = h->cfg.obs_dim;
  double *d_s = nullptr, *d_n = nullptr, *d_o = nullptr;
  auto cleanup = [&]() { cudaFree(d_s); cudaFree(d_n); cudaFree(d_o); };
  CKC(cudaMalloc(&d_s, sizeof(double) * n * D));
  CKC(cudaMalloc(&d_n, sizeof(double) * n * D));
  CKC(cudaMalloc(&d_o, sizeof(double) * n));
  CKC(cudaMemcpy(d_s, states, sizeof(double) * n * D, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(d_n, next_states, sizeof(double) * n * D, cudaMemcpyHostToDevice));
  if (h->timing) { CKC(cudaEventRecord(h->ev0, h->stream)); }
  cheetah_cost_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_s, d_n, n, D, d_o);
  CKC(cudaGetLastError());
  if (h->timing) { CKC(cudaEventRecord(h->ev1, h->stream)); h->ev_valid = true; }
  h->launches++;
  CKC(cudaStreamSynchronize(h->stream));
  CKC(cudaMemcpy(out, d_o, sizeof(double) * n, cudaMemcpyDeviceToHost));
  cleanup();
  return BCMPC_OK;
}

int bcmpc_sample_injection(bcmpc_handle* h, uint64_t seed, int32_t n_states, double* out) {
  if (!h || !out) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n_states < 1) return fail(h, BCMPC_ERR_INVALID, "n_states must be >= 1");
  ON_DEVICE(h);
  const int A = h->cfg.act_dim, K = h->cfg.num_paths, H = h->cfg.horizon;
  const int64_t rows = (int64_t)n_states * H * K;
  double* d_o = nullptr;
  auto cleanup = [&]() { cudaFree(d_o); };
  CKC(cudaMalloc(&d_o, sizeof(double) * rows * A));
  if (h->timing) { CKC(cudaEventRecord(h->ev0, h->stream)); }
  sample_injection_kernel<<<dim3((unsigned)((K + kSampleRows - 1) / kSampleRows), (unsigned)H, (unsigned)(n_states < 65535 ? n_states : 65535)), 256, 0, h->stream>>>(
      h->norm, A, K, H, n_states, inj_kind_of(first_proposal_of(h->cfg)), inj_kind_of(h->cfg.proposal), seed, h->cfg.path_offset, d_o);
  CKC(cudaGetLastError());
  if (h->timing) { CKC(cudaEventRecord(h->ev1, h->stream)); h->ev_valid = true; }
  h->launches++;
  CKC(cudaStreamSynchronize(h->stream));
  CKC(cudaMemcpy(out, d_o, sizeof(double) * rows * A, cudaMemcpyDeviceToHost));
  cleanup();
  return BCMPC_OK;
}
#undef CKC

int bcmpc_reduce_records(bcmpc_handle* h, const bcmpc_record* d_gathered, int32_t n_ranks, int32_t n_states,
                         bcmpc_record* d_out, void* stream) {
  if (!h || !d_gathered || !d_out) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (n_ranks < 1 || n_states < 1) return fail(h, BCMPC_ERR_INVALID, "n_ranks and n_states must be >= 1");
  ON_DEVICE(h);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  reduce_candidates_kernel<<<n_states, 256, 0, st>>>(d_gathered, n_ranks, n_states, 1, d_out);
  CK(h, cudaGetLastError());
  h->launches++;
  return BCMPC_OK;
}

int bcmpc_current_device(int32_t* device) {
  if (!device) return fail(nullptr, BCMPC_ERR_INVALID, "null argument");
  int d = 0;
  cudaError_t e = cudaGetDevice(&d);
  if (e != cudaSuccess) return fail(nullptr, BCMPC_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  *device = d;
  return BCMPC_OK;
}

int64_t bcmpc_peer_exchange_bytes(int32_t world, int32_t max_states) {
  if (world < 1 || world > kMaxPeers || max_states < 1) return -1;
  return (int64_t)peer_exchange_bytes(world, max_states);
}

int bcmpc_set_peer_exchange(bcmpc_handle* h, int32_t world, int32_t rank, void* const* peer_bufs) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (!peer_bufs) { h->peer = PeerExchangeArgs{}; return BCMPC_OK; }
  if (world < 1 || world > kMaxPeers) return fail(h, BCMPC_ERR_INVALID, "world %d out of [1,%d]", world, kMaxPeers);
  if (rank < 0 || rank >= world) return fail(h, BCMPC_ERR_INVALID, "rank %d out of [0,%d)", rank, world);
  PeerExchangeArgs x{};
  for (int r = 0; r < world; ++r) {
    if (!peer_bufs[r]) return fail(h, BCMPC_ERR_INVALID, "null peer buffer for rank %d", r);
    x.buf[r] = static_cast<unsigned char*>(peer_bufs[r]);
  }
  x.world = world; x.rank = rank; x.max_states = h->cfg.max_states; x.epoch = 0;
  h->peer = x;
  return BCMPC_OK;
}

int64_t bcmpc_kernel_launches(const bcmpc_handle* h) { return h ? h->launches : 0; }

int bcmpc_enable_timing(bcmpc_handle* h, int32_t on) {
  if (!h) return fail(h, BCMPC_ERR_INVALID, "null argument");
  h->timing = on != 0;
  h->ev_valid = false;
  return BCMPC_OK;
}

int bcmpc_last_kernel_ms(bcmpc_handle* h, float* ms) {
  if (!h || !ms) return fail(h, BCMPC_ERR_INVALID, "null argument");
  if (!h->ev_valid) return fail(h, BCMPC_ERR_STATE, "no timed rollout kernel yet (bcmpc_enable_timing)");
  CK(h, cudaEventSynchronize(h->ev1));
  CK(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return BCMPC_OK;
}

}  // extern "C"
