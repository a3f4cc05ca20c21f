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
  const bool ln = a.g1 != nullptr || a.Wr1 != nullptr;      // variants without hidden "ones" units
  if (n < a.S) return k < in_dim ? f32_to_bf16(a.W1[(size_t)k * a.ld1 + n]) : k == in_dim ? pk_hi(a.b1[n]) : k == in_dim + 1 ? pk_lo(a.b1[n]) : 0;
  if (!ln && (n == a.S || n == a.S + 1)) return k == in_dim ? (uint16_t)0x3F80 : 0;   // the hidden "ones" units: relu(1 * 1) = 1
  return 0;
}
__device__ __forceinline__ uint16_t pk_w2(const PackArgs& a, int k, int n) {
  const int S = a.S;
  if (a.Wr1) {        // columns [0, s_pad): state-delta head, [s_pad, 2 s_pad): reward head; biases are added in the epilogue
    const int nn = n % a.s_pad;
    if (nn >= S || k >= S) return 0;
    return f32_to_bf16(n < a.s_pad ? a.W2[(size_t)k * a.ld2 + nn] : a.Wr1[(size_t)k * a.ldr1 + nn]);
  }
  if (a.g1) return (n < S && k < S) ? f32_to_bf16(__fmul_rn(a.g1[k], a.W2[(size_t)k * a.ld2 + n])) : 0;
  if (n < S) return k < S ? f32_to_bf16(a.W2[(size_t)k * a.ld2 + n]) : k == S ? pk_hi(a.b2[n]) : k == S + 1 ? pk_lo(a.b2[n]) : 0;
  if (n == S || n == S + 1) return k == S ? (uint16_t)0x3F80 : 0;
  return 0;
}
__device__ __forceinline__ uint16_t pk_w3(const PackArgs& a, int k, int n) {        // n < 32 (output column)
  const int S = a.S;
  if (a.Wr1) {        // rows [0, s_pad): hidden units of the delta head -> columns 0..19; [s_pad, 2 s_pad): reward head -> column 20
    const int kk = k % a.s_pad;
    if (kk >= S) return 0;
    if (k < a.s_pad) return n < kD ? f32_to_bf16(a.W3[(size_t)kk * a.ld3 + n]) : 0;
    return n == kRewardCol ? f32_to_bf16(a.Wr2[(size_t)kk * a.ldr2]) : 0;
  }
  if (n >= kD) return 0;
  if (a.g1) return k < S ? f32_to_bf16(__fmul_rn(a.g2[k], a.W3[(size_t)k * a.ld3 + n])) : 0;
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

// LayerNorm correction vectors, one warp per column: lnvec = [u2 (s_pad)][c2 (s_pad)][u3 (32)][c3 (32)]
__global__ void pack_ln_kernel(PackArgs a) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  const int sp = a.s_pad, S = a.S;
  if (col >= sp + kOutN) return;
  const bool l3 = col >= sp;
  const int n = l3 ? col - sp : col;
  const bool live = l3 ? n < kD : n < S;
  double u = 0.0, c = 0.0;
  if (live)
    for (int k = lane; k < S; k += 32) {
      if (l3) { u += (double)bf16_to_f32(pk_w3(a, k, n)); c += (double)a.be2[k] * (double)a.W3[(size_t)k * a.ld3 + n]; }
      else    { u += (double)bf16_to_f32(pk_w2(a, k, n)); c += (double)a.be1[k] * (double)a.W2[(size_t)k * a.ld2 + n]; }
    }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { u += __shfl_xor_sync(0xffffffffu, u, off); c += __shfl_xor_sync(0xffffffffu, c, off); }
  if (lane == 0) {
    if (live) c += (double)(l3 ? a.b3[n] : a.b2[n]);
    float* pu = l3 ? a.lnvec + 2 * sp : a.lnvec;
    float* pc = l3 ? a.lnvec + 2 * sp + kOutN : a.lnvec + sp;
    pu[n] = live ? (float)u : 0.f;
    pc[n] = live ? (float)c : 0.f;
  }
}

// learned-reward model: lnvec = [b2 (s_pad)][br1 (s_pad)][b3 (20), br2, 0 ...]
__global__ void pack_reward_bias_kernel(PackArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, sp = a.s_pad;
  if (i >= 2 * sp + kOutN) return;
  float v = 0.f;
  if (i < sp) v = i < a.S ? a.b2[i] : 0.f;
  else if (i < 2 * sp) v = i - sp < a.S ? a.br1[i - sp] : 0.f;
  else { const int d = i - 2 * sp; v = d < kD ? a.b3[d] : d == kRewardCol ? a.br2[0] : 0.f; }
  a.lnvec[i] = v;
}

inline cudaError_t tm_pack(const PackArgs& a, cudaStream_t st) {
  const int sp2 = a.Wr1 ? 2 * a.s_pad : a.s_pad;
  const int total = 64 * a.s_pad + sp2 * kOutN + a.s_pad * sp2;
  pack_images_kernel<<<(total + 255) / 256, 256, 0, st>>>(a);
  if (a.g1) pack_ln_kernel<<<(a.s_pad + kOutN + 7) / 8, 256, 0, st>>>(a);
  if (a.Wr1) pack_reward_bias_kernel<<<(2 * a.s_pad + kOutN + 255) / 256, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace tmx
}  // namespace bcmpc
