// Device-side packing of the fast-mode weight images (see tm_geometry in rollout_tmem.cuh for the layouts).
// Included by api.cu only.
#pragma once
#include "rollout_tmem.cuh"

namespace bcmpc {
namespace tmx {

struct PackArgs {
  const float *W1, *b1, *W2, *b2, *W3, *b3;     // device, row strides ld1..ld3
  const float *g1, *be1, *g2, *be2;             // LayerNorm gamma / beta of hidden layers 1, 2 (nullptr = off)
  const float *Wr1, *br1, *Wr2, *br2;           // reward head [S,S], [S,1] (nullptr = off; learned-reward model)
  const float *gr, *ber;                        // LayerNorm gamma / beta of the reward head's hidden layer
  int explicit_bias;                            // biases of layers 2 / 3 go to lnvec instead of into the GEMMs (no hidden "ones" units)
  int ld1, ld2, ld3, ldr1, ldr2;
  int S, s_pad, nc, cg, kbs, nstg;
  uint32_t blk_h, blk_o, w1_bytes, res_bytes, stage_bytes;
  uint16_t *res, *stream;
  float* lnvec;
};

__device__ __forceinline__ uint16_t pk_hi(float v) { return f32_to_bf16(v); }
__device__ __forceinline__ uint16_t pk_lo(float v) { return f32_to_bf16(__fsub_rn(v, bf16_to_f32(f32_to_bf16(v)))); }
// effective (bias-augmented, zero-padded) kernels
__device__ __forceinline__ uint16_t pk_w1(const PackArgs& a, int k, int n) {        // k < 64 (input column), n < s_pad (hidden unit)
  constexpr int in_dim = kD + kA;
  const bool ln = a.g1 != nullptr || a.Wr1 != nullptr || a.explicit_bias;      // variants without hidden "ones" units
  if (n < a.S) return k < in_dim ? f32_to_bf16(a.W1[(size_t)k * a.ld1 + n]) : k == in_dim ? pk_hi(a.b1[n]) : k == in_dim + 1 ? pk_lo(a.b1[n]) : 0;
  if (!ln && (n == a.S || n == a.S + 1)) return k == in_dim ? (uint16_t)0x3F80 : 0;   // the hidden "ones" units: relu(1 * 1) = 1
  return 0;
}
// Hidden layer 2 and the output layer as seen by the kernel.  Learned-reward model: layer 2 has 2 * s_pad columns
// (delta head, then reward head) and the output layer 2 * s_pad rows, block-structured: delta-head units feed columns
// 0..19, reward-head units column 20.  `raw` = the float32 weight, `gamma`/`beta` = the LayerNorm of the INPUT of that
// weight row, `bias` = the bias of that output column.
__device__ __forceinline__ bool w2_live(const PackArgs& a, int k, int n) { return k < a.S && (a.Wr1 ? n % a.s_pad : n) < a.S && n < (a.Wr1 ? 2 : 1) * a.s_pad; }
__device__ __forceinline__ float w2_raw(const PackArgs& a, int k, int n) {
  if (a.Wr1 && n >= a.s_pad) return a.Wr1[(size_t)k * a.ldr1 + (n - a.s_pad)];
  return a.W2[(size_t)k * a.ld2 + n];
}
__device__ __forceinline__ float w2_bias(const PackArgs& a, int n) { return (a.Wr1 && n >= a.s_pad) ? a.br1[n - a.s_pad] : a.b2[n]; }
__device__ __forceinline__ bool w3_live(const PackArgs& a, int k, int n) {
  if (!a.Wr1) return k < a.S && n < kD;
  if (k % a.s_pad >= a.S || k >= 2 * a.s_pad) return false;
  return k < a.s_pad ? n < kD : n == kRewardCol;
}
__device__ __forceinline__ float w3_raw(const PackArgs& a, int k, int n) {
  if (a.Wr1 && k >= a.s_pad) return a.Wr2[(size_t)(k - a.s_pad) * a.ldr2];
  return a.W3[(size_t)k * a.ld3 + n];
}
__device__ __forceinline__ float w3_gamma(const PackArgs& a, int k) { return (a.Wr1 && k >= a.s_pad) ? a.gr[k - a.s_pad] : a.g2[k]; }
__device__ __forceinline__ float w3_beta(const PackArgs& a, int k) { return (a.Wr1 && k >= a.s_pad) ? a.ber[k - a.s_pad] : a.be2[k]; }
__device__ __forceinline__ float w3_bias(const PackArgs& a, int n) { return (a.Wr1 && n == kRewardCol) ? a.br2[0] : a.b3[n]; }

__device__ __forceinline__ uint16_t pk_w2(const PackArgs& a, int k, int n) {
  const int S = a.S;
  if (a.g1 || a.Wr1 || a.explicit_bias) {         // no bias rows: LayerNorm folds the bias into c, the others add it in the epilogue
    if (!w2_live(a, k, n)) return 0;
    const float w = w2_raw(a, k, n);
    return f32_to_bf16(a.g1 ? __fmul_rn(a.g1[k], w) : w);
  }
  if (n < S) return k < S ? f32_to_bf16(a.W2[(size_t)k * a.ld2 + n]) : k == S ? pk_hi(a.b2[n]) : k == S + 1 ? pk_lo(a.b2[n]) : 0;
  if (n == S || n == S + 1) return k == S ? (uint16_t)0x3F80 : 0;
  return 0;
}
__device__ __forceinline__ uint16_t pk_w3(const PackArgs& a, int k, int n) {        // n < 32 (output column)
  const int S = a.S;
  if (a.g1 || a.Wr1 || a.explicit_bias) {
    if (!w3_live(a, k, n)) return 0;
    const float w = w3_raw(a, k, n);
    return f32_to_bf16(a.g1 ? __fmul_rn(w3_gamma(a, k), w) : w);
  }
  if (n >= kD) return 0;
  return k < S ? f32_to_bf16(a.W3[(size_t)k * a.ld3 + n]) : k == S ? pk_hi(a.b3[n]) : k == S + 1 ? pk_lo(a.b3[n]) : 0;
}

// One thread per element of the three padded kernels (n fastest: coalesced reads of the float32 rows).
//   resident image, per CTA rank r: [W1 chunk 0..NC-1][W3 k-block 0..NC-1]
//   streamed image: one slot [CG][stage_bytes] per (chunk, stage), chunk-major = consumption order
__global__ void pack_images_kernel(PackArgs a) {
  const int sp = a.s_pad, rows_h = kChunk / a.cg, rows_o = kOutN / a.cg;
  const int sp2 = a.Wr1 ? 2 * sp : sp;            // width of hidden layer 2 (both heads)
  const int n1 = 64 * sp, n3 = sp2 * kOutN, n2 = sp * sp2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n3 + n2; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int k = i / sp, n = i % sp, c = n / kChunk, j = n % kChunk, r = j / rows_h, ii = j % rows_h;
      a.res[((size_t)r * a.res_bytes + (size_t)c * a.blk_h + sw128_offset(ii, k)) / 2] = pk_w1(a, k, n);
    } else if (i < n1 + n3) {
      const int e = i - n1, k = e / kOutN, n = e % kOutN, r = n / rows_o, ii = n % rows_o;
      a.res[((size_t)r * a.res_bytes + a.w1_bytes + (size_t)(k / 64) * a.blk_o + sw128_offset(ii, k % 64)) / 2] = pk_w3(a, k, n);
    } else {
      const int e = i - n1 - n3, k = e / sp2, n = e % sp2, c = n / kChunk, j = n % kChunk, r = j / rows_h, ii = j % rows_h;
      const int kb = k / 64, stg = kb / a.kbs, t = kb % a.kbs;
      a.stream[((size_t)(c * a.nstg + stg) * a.cg * a.stage_bytes + (size_t)r * a.stage_bytes + (size_t)t * a.blk_h +
                sw128_offset(ii, k % 64)) / 2] = pk_w2(a, k, n);
    }
  }
}

// LayerNorm correction vectors, one warp per column: lnvec = [u2 (sp2)][c2 (sp2)][u3 (32)][c3 (32)], sp2 = width of
// hidden layer 2:  u = column sum of the bf16 image (gamma folded in), c = beta^T W + bias
__global__ void pack_ln_kernel(PackArgs a) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  const int sp2 = a.Wr1 ? 2 * a.s_pad : a.s_pad, S = a.S;
  if (col >= sp2 + kOutN) return;
  const bool l3 = col >= sp2;
  const int n = l3 ? col - sp2 : col;
  double u = 0.0, c = 0.0;
  bool live = false;
  if (l3) {
    // the rows that feed output column n: the delta head's units for n < 20, the reward head's for n == 20
    const int k0 = (a.Wr1 && n == kRewardCol) ? a.s_pad : 0;
    live = w3_live(a, k0, n);
    if (live)
      for (int k = k0 + lane; k < k0 + S; k += 32) {
        u += (double)bf16_to_f32(pk_w3(a, k, n));
        c += (double)w3_beta(a, k) * (double)w3_raw(a, k, n);
      }
  } else {
    live = w2_live(a, 0, n);
    if (live)
      for (int k = lane; k < S; k += 32) {
        u += (double)bf16_to_f32(pk_w2(a, k, n));
        c += (double)a.be1[k] * (double)w2_raw(a, k, n);
      }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { u += __shfl_xor_sync(0xffffffffu, u, off); c += __shfl_xor_sync(0xffffffffu, c, off); }
  if (lane == 0) {
    if (live) c += (double)(l3 ? w3_bias(a, n) : w2_bias(a, n));
    float* pu = l3 ? a.lnvec + 2 * sp2 : a.lnvec;
    float* pc = l3 ? a.lnvec + 2 * sp2 + kOutN : a.lnvec + sp2;
    pu[n] = live ? (float)u : 0.f;
    pc[n] = live ? (float)c : 0.f;
  }
}

// explicit biases: lnvec = [b2 (s_pad)]{[br1 (s_pad)] with the reward model}[b3 (20), {br2,} 0 ...]
__global__ void pack_reward_bias_kernel(PackArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, sp = a.s_pad, sp2 = a.Wr1 ? 2 * sp : sp;
  if (i >= sp2 + kOutN) return;
  float v = 0.f;
  if (i < sp) v = i < a.S ? a.b2[i] : 0.f;
  else if (i < sp2) v = i - sp < a.S ? a.br1[i - sp] : 0.f;
  else { const int d = i - sp2; v = d < kD ? a.b3[d] : (a.Wr1 && d == kRewardCol) ? a.br2[0] : 0.f; }
  a.lnvec[i] = v;
}

// Proposal network for the fast path: sizes = {20, h1, h2, 6}, tanh hidden layers, h1, h2 <= 64 pn (pn = 2 chunk slots per
// layer, or 4 for the kVarPolWide variants).  W[l] are [in,out] with row stride ld[l] (device).  Image per CTA rank r
// (zero-filled by the caller first):
//   [Wp1 chunk 0..pn-1: 64 units x 64 k (20 used)][Wp2 chunk j x k-block kb (pn x pn)][Wp3 k-block 0..pn-1: 32 outputs (6 used) x 64 k]
// and the bias vector polvec = [bp1 (64 pn)][bp2 (64 pn)][bp3 (8)].
struct PolicyPackArgs {
  const float *W[3], *b[3];
  int ld[3];
  int h1, h2, cg, pn;
  uint32_t blk_h, blk_o, img_bytes;
  uint16_t* img;
  float* vec;
};
__global__ void pack_policy_kernel(PolicyPackArgs a) {
  const int rows_h = kChunk / a.cg, rows_o = kOutN / a.cg;
  const int n1 = kD * a.h1, n2 = a.h1 * a.h2, n3 = a.h2 * kA;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int k = i / a.h1, n = i % a.h1, j = n / kChunk, r = (n % kChunk) / rows_h, ii = (n % kChunk) % rows_h;
      a.img[((size_t)r * a.img_bytes + (size_t)j * a.blk_h + sw128_offset(ii, k)) / 2] = f32_to_bf16(a.W[0][(size_t)k * a.ld[0] + n]);
    } else if (i < n1 + n2) {
      const int e = i - n1, k = e / a.h2, n = e % a.h2, j = n / kChunk, r = (n % kChunk) / rows_h, ii = (n % kChunk) % rows_h;
      a.img[((size_t)r * a.img_bytes + (size_t)(a.pn + j * a.pn + k / 64) * a.blk_h + sw128_offset(ii, k % 64)) / 2] =
          f32_to_bf16(a.W[1][(size_t)k * a.ld[1] + n]);
    } else {
      const int e = i - n1 - n2, k = e / kA, n = e % kA, r = n / rows_o, ii = n % rows_o;
      a.img[((size_t)r * a.img_bytes + (size_t)(a.pn + a.pn * a.pn) * a.blk_h + (size_t)(k / 64) * a.blk_o + sw128_offset(ii, k % 64)) / 2] =
          f32_to_bf16(a.W[2][(size_t)k * a.ld[2] + n]);
    }
  }
  const int t = blockIdx.x * blockDim.x + threadIdx.x, ph = 64 * a.pn;
  if (t < 2 * ph + kMaxAct) {
    float v = 0.f;
    if (t < ph) v = t < a.h1 ? a.b[0][t] : 0.f;
    else if (t < 2 * ph) v = t - ph < a.h2 ? a.b[1][t - ph] : 0.f;
    else v = t - 2 * ph < kA ? a.b[2][t - 2 * ph] : 0.f;
    a.vec[t] = v;
  }
}
inline cudaError_t tm_pack_policy(const PolicyPackArgs& a, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(a.img, 0, (size_t)a.cg * a.img_bytes, st);
  if (e != cudaSuccess) return e;
  const int total = kD * a.h1 + a.h1 * a.h2 + a.h2 * kA;
  const int nvec = 2 * 64 * a.pn + kMaxAct;
  const int threads = total > nvec ? total : nvec;   // the first threads also write the bias vector
  pack_policy_kernel<<<(threads + 255) / 256, 256, 0, st>>>(a);
  return cudaGetLastError();
}

inline cudaError_t tm_pack(const PackArgs& a, cudaStream_t st) {
  const int sp2 = a.Wr1 ? 2 * a.s_pad : a.s_pad;
  const int total = 64 * a.s_pad + sp2 * kOutN + a.s_pad * sp2;
  pack_images_kernel<<<(total + 255) / 256, 256, 0, st>>>(a);
  if (a.g1) pack_ln_kernel<<<(sp2 + kOutN + 7) / 8, 256, 0, st>>>(a);
  else if (a.Wr1 || a.explicit_bias) pack_reward_bias_kernel<<<(sp2 + kOutN + 255) / 256, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace tmx
}  // namespace bcmpc
