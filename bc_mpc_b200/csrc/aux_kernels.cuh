// Small HBM-bound helper kernels: candidate reduction (argmin), stand-alone cost, injection dump.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace bcmpc {

// Reduces rows of candidates to one record each: in [n_out][n_in] -> out[n_out]
// (np.argmin order, controllers.py:82).  One CTA per output record.  With `strided` the input is
// [n_in][n_out] (gathered per-rank records, rank-major) instead.
__global__ void __launch_bounds__(256) reduce_candidates_kernel(const Candidate* __restrict__ in, int n_in, int n_out,
                                                                int strided, bcmpc_record* __restrict__ out) {
  const int o = blockIdx.x;
  __shared__ double s_cost[256];
  __shared__ int64_t s_idx[256];
  __shared__ int s_src[256];
  double c = 0.0;
  int64_t idx = -1;
  int src = -1;
  for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
    const Candidate& cd = strided ? in[(size_t)i * n_out + o] : in[(size_t)o * n_in + i];
    if (cd.index < 0) continue;
    if (idx < 0 || cand_better(cd.cost, cd.index, c, idx)) { c = cd.cost; idx = cd.index; src = i; }
  }
  s_cost[threadIdx.x] = c; s_idx[threadIdx.x] = idx; s_src[threadIdx.x] = src;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      const int t = threadIdx.x + off;
      if (s_idx[t] >= 0 && (s_idx[threadIdx.x] < 0 || cand_better(s_cost[t], s_idx[t], s_cost[threadIdx.x], s_idx[threadIdx.x]))) {
        s_cost[threadIdx.x] = s_cost[t]; s_idx[threadIdx.x] = s_idx[t]; s_src[threadIdx.x] = s_src[t];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    bcmpc_record r;
    r.cost = s_cost[0];
    r.index = s_idx[0];
    for (int j = 0; j < kMaxAct; ++j) r.action[j] = 0.0;
    if (s_src[0] >= 0) {
      const Candidate& cd = strided ? in[(size_t)s_src[0] * n_out + o] : in[(size_t)o * n_in + s_src[0]];
      for (int j = 0; j < kMaxAct; ++j) r.action[j] = cd.action[j];
    }
    out[o] = r;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Cross-GPU min-loc over NVLink peer memory: the one exchange step of a sharded plan as ONE kernel
// per rank instead of reduce kernel + ncclAllGather + reduce kernel.
//   every rank owns a peer-mapped buffer  [2 slots][world][max_states] records  + flags [2 slots][world] (uint32)
//   block p of rank r: per-state min-loc of rank r's tile candidates (recomputed by every block: a few tiles per state) ->
//                      plain stores into PEER p's buffer[slot][r][*] over NVLink -> __threadfence_system ->
//                      release store of the epoch number into peer p's flag[slot][r]
//   block 0 of rank r: acquire-polls its OWN flags [slot][0..world) for the epoch, then reduces its own buffer[slot] over
//                      the ranks (lowest GLOBAL index on ties: np.argmin order, identical on every rank) -> out
// slot = epoch & 1: a rank can be at most one exchange ahead of a peer (it needs every peer's records of exchange e to
// finish e), so two slots are enough and nothing is ever reset.  Every rank of the group must launch it (a collective).
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerExchangeArgs {
  unsigned char* buf[kMaxPeers];     // base of every rank's exchange buffer (peer-mapped device pointers; [rank] is local)
  int world, rank, max_states;
  uint32_t epoch;                    // 1, 2, 3, ... (flags start at 0)
};
__host__ __device__ inline size_t peer_exchange_bytes(int world, int max_states) {
  return (size_t)2 * world * max_states * sizeof(bcmpc_record) + 2 * kMaxPeers * sizeof(uint32_t);
}
__global__ void __launch_bounds__(256) peer_exchange_kernel(PeerExchangeArgs x, const Candidate* __restrict__ cand, int tiles_per_state,
                                                            int n_states, bcmpc_record* __restrict__ out) {
  const uint32_t slot = x.epoch & 1u;
  const size_t rec_bytes = (size_t)2 * x.world * x.max_states * sizeof(bcmpc_record);
  const int peer = blockIdx.x;
  // ---- my records -> peer `peer` ----
  bcmpc_record* dst = reinterpret_cast<bcmpc_record*>(x.buf[peer]) + ((size_t)slot * x.world + x.rank) * x.max_states;
  for (int b = threadIdx.x; b < n_states; b += blockDim.x) {
    const Candidate* cd = cand + (size_t)b * tiles_per_state;
    double c = 0.0;
    long long idx = -1;
    int src = -1;
    for (int t = 0; t < tiles_per_state; ++t) {
      const long long ii = cd[t].index;
      if (ii >= 0 && (idx < 0 || cand_better(cd[t].cost, (int64_t)ii, c, (int64_t)idx))) { c = cd[t].cost; idx = ii; src = t; }
    }
    bcmpc_record r;
    r.cost = idx < 0 ? 0.0 : c;
    r.index = idx;
    for (int j = 0; j < kMaxAct; ++j) r.action[j] = src >= 0 ? cd[src].action[j] : 0.0;
    dst[b] = r;                                          // st.global on a peer-mapped pointer: travels over NVLink
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* flag = reinterpret_cast<uint32_t*>(x.buf[peer] + rec_bytes) + slot * kMaxPeers + x.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(x.epoch) : "memory");
  }
  if (blockIdx.x != 0) return;
  // ---- block 0: wait for every rank's records of this exchange, then the G-way min-loc ----
  if ((int)threadIdx.x < x.world) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(x.buf[x.rank] + rec_bytes) + slot * kMaxPeers + threadIdx.x;
    const long long t0 = clock64();
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v != x.epoch && clock64() - t0 > 20000000000LL) {
        printf("bcmpc: peer exchange timed out (rank %d waiting for rank %d, epoch %u, flag %u)\n", x.rank, (int)threadIdx.x, x.epoch, v);
        __trap();
      }
    } while (v != x.epoch);
  }
  __syncthreads();
  const bcmpc_record* mine = reinterpret_cast<const bcmpc_record*>(x.buf[x.rank]) + (size_t)slot * x.world * x.max_states;
  for (int b = threadIdx.x; b < n_states; b += blockDim.x) {
    double c = 0.0;
    long long idx = -1;
    int src = -1;
    for (int g = 0; g < x.world; ++g) {
      const volatile bcmpc_record* r = mine + (size_t)g * x.max_states + b;
      const long long ii = r->index;
      const double cc = r->cost;
      if (ii >= 0 && (idx < 0 || cand_better(cc, (int64_t)ii, c, (int64_t)idx))) { c = cc; idx = ii; src = g; }
    }
    bcmpc_record r;
    r.cost = idx < 0 ? 0.0 : c;
    r.index = idx;
    for (int j = 0; j < kMaxAct; ++j) r.action[j] = src >= 0 ? ((const volatile bcmpc_record*)(mine + (size_t)src * x.max_states + b))->action[j] : 0.0;
    out[b] = r;
  }
}

// cheetah_cost_fn batched branch (cost_functions.py:10-30), float64.
__global__ void cheetah_cost_kernel(const double* __restrict__ s, const double* __restrict__ ns, int64_t n, int D,
                                    double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* a = s + i * D;
  double c = 0.0;
  if (a[5] >= 0.2) c += 10.0;
  if (a[6] >= 0.0) c += 10.0;
  if (a[7] >= 0.0) c += 10.0;
  c -= (ns[i * D + 17] - a[17]) / 0.01;
  out[i] = c;
}

// dst[i * ld + o] = src[o * in + i]: a torch.nn.Linear weight [out,in] into the padded [in,out] layout of the handle
__global__ void transpose_in_kernel(float* __restrict__ dst, int ld, const float* __restrict__ src, int in, int out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= in * out) return;
  const int i = e / out, o = e % out;
  dst[(size_t)i * ld + o] = src[(size_t)o * in + i];
}

// Writes the injection tensor [B,H,K,A] exactly as the rollout kernels draw it.
// grid = (ceil(K / kSampleRows), H, min(B, 65535)).  HalfCheetah rows are 48 bytes: a per-row store touches twelve lines per warp
// instruction and the L1 store path becomes the limit (ncu: l1tex 85 %, DRAM 42 %), so a block stages its rows in shared
// memory (16-byte lanes at a 48-byte stride are conflict-free) and one thread hands the contiguous span to the TMA engine.
constexpr int kSampleRows = 512;   // rows per block: two per thread (independent Philox chains in flight)
__global__ void __launch_bounds__(256) sample_injection_kernel(NormArgs n, int A, int K, int H, int B, int kind0, int kind,
                                                               uint64_t seed, int64_t path_offset,
                                                               double* __restrict__ out) {
  __shared__ __align__(128) double2 stage[kSampleRows * 3];
  const int k0 = blockIdx.x * kSampleRows;
  const int h = blockIdx.y;
  const int knd = h == 0 ? kind0 : kind;
  const int nrows = K - k0 < kSampleRows ? K - k0 : kSampleRows;
  for (int b = blockIdx.z; b < B; b += gridDim.z) {
    const int64_t row0 = ((int64_t)b * H + h) * K + k0;
    float v[2][kMaxAct];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = k0 + r * 256 + threadIdx.x;
      if (k < K) {
        if (A == 6)   // literal width: the draw's per-element predicates fold away
          draw_injection(v[r], 6, knd, seed, (uint32_t)(path_offset + k), (uint32_t)b, (uint32_t)h, n.low, n.span);
        else
          draw_injection(v[r], A, knd, seed, (uint32_t)(path_offset + k), (uint32_t)b, (uint32_t)h, n.low, n.span);
      }
    }
    if (A == 6) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int t = r * 256 + threadIdx.x;
        if (t < nrows) {
          stage[t * 3 + 0] = make_double2((double)v[r][0], (double)v[r][1]);
          stage[t * 3 + 1] = make_double2((double)v[r][2], (double)v[r][3]);
          stage[t * 3 + 2] = make_double2((double)v[r][4], (double)v[r][5]);
        }
      }
      ptx::fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        ptx::bulk_s2g(out + row0 * 6, ptx::smem_u32(stage), (uint32_t)nrows * 48u);
        ptx::bulk_wait_read();
      }
      __syncthreads();
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int t = r * 256 + threadIdx.x;
        if (t < nrows)
          for (int j = 0; j < A; ++j) out[(row0 + t) * A + j] = (double)v[r][j];
      }
    }
  }
}

}  // namespace bcmpc
