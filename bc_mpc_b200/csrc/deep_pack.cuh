// Device-side packing of the cluster-split kernel's weight image (rollout_deep.cuh).  Included by api.cu only.
#pragma once
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {

// ---------------------------------------------------------------------------------------------------------
// weight image, packed on the device from the float32 copies the handle keeps (dynamics.py:54-71 layer stack):
//   per cluster rank r = 2 p + q, per step the stages in consumption order (see the schedule above); a tile is the UMMA B
//   operand of one (output chunk, k-block): this CTA's half of the chunk's rows (q) x 64 k, K-major SWIZZLE_128B.
//   Layer 1 carries its bias in the GEMM (input columns 26 / 27 are 1: rows 26 / 27 hold bf16(b), bf16(b - bf16(b))).
// bias vector: [layer 2 (128 nch)] ... [layer L (128 nch)][output (32)]
// LayerNorm (g / be = gamma / beta of hidden layer j at index j - 1, dynamics.py:68-69): the kernels of layers >= 2 and of the
// output layer are packed as W' = diag(gamma) W, and the vector holds, per layer, the corrections of
//   LN(r) W + b = rho (r W') - rho mu u + c :   [u (128 nch)][c (128 nch)] ... [u3 (32)][c3 (32)],
// u = column sums of the bf16-rounded W' (what the tensor core multiplies), c = beta^T W + b.
// ---------------------------------------------------------------------------------------------------------
struct PackArgs {
  const float* W[BCMPC_MAX_LAYERS];
  const float* b[BCMPC_MAX_LAYERS];
  const float* g[BCMPC_MAX_LAYERS];     // LayerNorm gamma / beta of hidden layer j at [j - 1], or all nullptr
  const float* be[BCMPC_MAX_LAYERS];
  int ld[BCMPC_MAX_LAYERS];
  int L, S, nch;
  uint16_t* stream;
  float* vec;
};

__global__ void pack_deep_kernel(PackArgs a) {
  const int nch = a.nch, spu = stages_per_use(nch), L = a.L, S = a.S;
  const int nstg = stages_per_step(nch, L);
  const size_t per_rank = (size_t)nstg * (kStageBytes / 2);
  const size_t total = (size_t)kCluster * per_rank;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / per_rank), p = r >> 1, q = r & 1;
    const size_t er = e % per_rank;
    const int s = (int)(er / (kStageBytes / 2));
    const uint32_t off = (uint32_t)(er % (kStageBytes / 2)) * 2u;        // byte offset in the stage
    uint16_t v = 0;
    auto untile = [](uint32_t o, int& n_local, int& kk) {                // inverse of sw128_offset within a tile
      const int g = (int)(o / 1024u), rr = (int)((o % 1024u) / 128u), c16 = (int)((o % 128u) / 16u) ^ rr, e2 = (int)((o % 16u) / 2u);
      n_local = g * 8 + rr;
      kk = c16 * 8 + e2;
    };
    const int l1s = spu;
    if (s < l1s) {                                                        // ---- layer 1: tiles of own chunks 4 s + t ----
      const int t = (int)(off / kTileH), c = 4 * s + t;
      int nl, kk;
      untile(off % kTileH, nl, kk);
      const int n = (p * nch + c) * 64 + q * 32 + nl;
      if (c < nch && n < S) {
        constexpr int in_dim = kD + kA;
        if (kk < in_dim) v = tmx::f32_to_bf16(a.W[0][(size_t)kk * a.ld[0] + n]);
        else if (kk == in_dim) v = tmx::f32_to_bf16(a.b[0][n]);
        else if (kk == in_dim + 1) v = tmx::f32_to_bf16(__fsub_rn(a.b[0][n], tmx::bf16_to_f32(tmx::f32_to_bf16(a.b[0][n]))));
      }
    } else if (s < nstg - spu) {                                          // ---- hidden layers 2..L ----
      const int s2 = s - l1s, per_layer = 2 * nch * spu;
      const int l = 1 + s2 / per_layer;                                   // weight index (W[l]: hidden l -> hidden l+1)
      const int rem = s2 % per_layer, use = rem / spu, sg = rem % spu;
      const int i = use >> 1, own = use & 1;
      const int gc = (own ? p : 1 - p) * nch + i;                         // global output chunk
      const int t = (int)(off / kTileH), kb = 4 * sg + t;
      int nl, kk;
      untile(off % kTileH, nl, kk);
      const int k = (p * nch + kb) * 64 + kk, n = gc * 64 + q * 32 + nl;
      if (kb < nch && k < S && n < S) v = tmx::f32_to_bf16(a.g[0] ? a.g[l - 1][k] * a.W[l][(size_t)k * a.ld[l] + n] : a.W[l][(size_t)k * a.ld[l] + n]);
    } else {                                                              // ---- output layer: stage sg = k-blocks 4 sg .. 4 sg + 3 ----
      const int sg = s - (nstg - spu), t = (int)(off / kTileO), kb = 4 * sg + t;
      int nl, kk;
      untile(off % kTileO, nl, kk);
      const int k = (p * nch + kb) * 64 + kk, n = q * 16 + nl;
      if (t < 4 && kb < nch && k < S && n < kD) v = tmx::f32_to_bf16(a.g[0] ? a.g[L - 1][k] * a.W[L][(size_t)k * a.ld[L] + n] : a.W[L][(size_t)k * a.ld[L] + n]);
    }
    a.stream[e] = v;
  }
  if (a.g[0]) return;                                   // LayerNorm: pack_deep_ln_kernel writes the vector
  const int sp2 = 2 * nch * 64, nvec = (L - 1) * sp2 + 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < (L - 1) * sp2) { const int l = 1 + i / sp2, n = i % sp2; v = n < S ? a.b[l][n] : 0.f; }
    else { const int d = i - (L - 1) * sp2; v = d < kD ? a.b[L][d] : 0.f; }
    a.vec[i] = v;
  }
}
// LayerNorm correction vectors: one warp per (weight index l = 1..L, output column n)
__global__ void pack_deep_ln_kernel(PackArgs a) {
  const int L = a.L, S = a.S, sp2 = 2 * a.nch * 64;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = (L - 1) * sp2 + 32;
  if (warp >= total) return;
  const bool outl = warp >= (L - 1) * sp2;
  const int l = outl ? L : 1 + warp / sp2, n = outl ? warp - (L - 1) * sp2 : warp % sp2;
  const bool live = outl ? n < kD : n < S;
  double u = 0.0, c = 0.0;
  if (live)
    for (int k = lane; k < S; k += 32) {
      const float w = a.W[l][(size_t)k * a.ld[l] + n];
      u += (double)tmx::bf16_to_f32(tmx::f32_to_bf16(a.g[l - 1][k] * w));
      c += (double)a.be[l - 1][k] * (double)w;
    }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { u += __shfl_xor_sync(0xffffffffu, u, off); c += __shfl_xor_sync(0xffffffffu, c, off); }
  if (lane == 0) {
    if (live) c += (double)a.b[l][n];
    float* pu = outl ? a.vec + (size_t)(L - 1) * 2 * sp2 : a.vec + (size_t)(l - 1) * 2 * sp2;
    float* pc = outl ? pu + 32 : pu + sp2;
    pu[n] = live ? (float)u : 0.f;
    pc[n] = live ? (float)c : 0.f;
  }
}
inline cudaError_t deep_pack(const PackArgs& a, cudaStream_t st) {
  pack_deep_kernel<<<592, 256, 0, st>>>(a);
  if (a.g[0]) {
    const int warps = (a.L - 1) * 2 * a.nch * 64 + 32;
    pack_deep_ln_kernel<<<(warps + 7) / 8, 256, 0, st>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace dpx
}  // namespace bcmpc
