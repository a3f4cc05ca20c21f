// Micro-benchmark: how fast can one SM push a 32 KB fp32 partial into a cluster partner's shared memory?
//   mode 0: st.async.v4 from registers (16 per thread, 128 threads) - what rollout_deep_kernel's sending group does
//   mode 1: the same data staged in local shared memory, then ONE thread hands it to the bulk-copy engine
//           (cp.async.bulk.shared::cluster.shared::cta, 2 x 16 KB)
//   mode 2: mode 1 without the staging stores (copy engine alone)
// Both CTAs of every pair (r <-> r ^ partner_xor) send and receive at the same time, all clusters of the grid at once.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bw dsmem_bw.cu      Run: ./dsmem_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t bar, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(x), "r"(y),
               "r"(z), "r"(w), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

constexpr int kMsg = 32768, kPerRound = 8;

template <int MODE>
__global__ void __launch_bounds__(128, 1) bw_kernel(int rounds, int partner_xor, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* recv = smem;                  // 32 KB receive buffer
  uint8_t* stage = smem + kMsg;          // 32 KB staging buffer (modes 1, 2)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kMsg);
  const uint32_t rank = ctarank(), partner = rank ^ (uint32_t)partner_xor;
  const int tid = threadIdx.x;
  const uint32_t bar_l = smem_u32(bar), bar_p = mapa(bar_l, partner);
  const uint32_t recv_p = mapa(smem_u32(recv), partner), stage_l = smem_u32(stage);
  if (tid == 0) { mbar_init(bar_l, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  cluster_sync();
  uint32_t v[16];
  for (int i = 0; i < 16; ++i) v[i] = tid * 16 + i;
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    if (tid == 0) mbar_expect_tx(bar_l, kMsg * kPerRound);
    for (int m = 0; m < kPerRound; ++m) {
      if (MODE == 0) {
        // 128 threads x 16 stores x 16 B = 32 KB; thread t writes [j][t] (row t of 16-byte column j), like the kernel
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_async_v4(recv_p + (uint32_t)(hf * 16384 + j * 2048 + tid * 16), bar_p, v[0] + m, v[1] + j, v[2] + hf, v[3]);
      } else {
        if (MODE == 1) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_l + (uint32_t)(hf * 16384 + j * 2048 + tid * 16)), "r"(v[0] + m),
                           "r"(v[1] + j), "r"(v[2] + hf), "r"(v[3])
                           : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncthreads();
        }
        if (tid == 0) {
          bulk_s2c(recv_p, stage_l, 16384, bar_p);
          bulk_s2c(recv_p + 16384, stage_l + 16384, 16384, bar_p);
        }
      }
    }
    mbar_wait(bar_l, (uint32_t)r & 1u);        // my partner's 8 messages have landed
    cluster_sync();                            // (nobody runs a round ahead)
  }
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char* name, int cluster, int partner_xor, int grid) {
  auto kern = bw_kernel<MODE>;
  const int smem = 2 * kMsg + 64;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned long long* d;
  cudaMalloc(&d, sizeof(unsigned long long) * grid);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const int rounds = 200;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, rounds, partner_xor, d);
    if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  unsigned long long h[1024];
  cudaMemcpy(h, d, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
  double mx = 0, sum = 0;
  for (int i = 0; i < grid; ++i) { sum += (double)h[i]; if ((double)h[i] > mx) mx = (double)h[i]; }
  const double per_msg = sum / grid / (rounds * kPerRound);
  printf("%-58s cluster %d, %3d CTAs: %7.0f cycles per 32 KB message (incl. 1/8 of a round's wait + cluster sync) = %5.1f B/clk per SM and direction (slowest CTA %.0f)\n",
         name, cluster, grid, per_msg, kMsg / per_msg, mx / (rounds * kPerRound));
  cudaFree(d);
}

int main() {
  run<0>("st.async from registers (16 x 16 B per thread)", 2, 1, 2);
  run<1>("smem staging + cp.async.bulk smem -> partner smem", 2, 1, 2);
  run<2>("cp.async.bulk smem -> partner smem alone", 2, 1, 2);
  run<0>("st.async from registers (16 x 16 B per thread)", 4, 2, 132);
  run<1>("smem staging + cp.async.bulk smem -> partner smem", 4, 2, 132);
  run<2>("cp.async.bulk smem -> partner smem alone", 4, 2, 132);
  run<0>("st.async from registers (16 x 16 B per thread)", 2, 1, 148);
  run<2>("cp.async.bulk smem -> partner smem alone", 2, 1, 148);
  return 0;
}
