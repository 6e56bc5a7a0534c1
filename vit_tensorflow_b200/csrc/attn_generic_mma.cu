// Second-tier attention path for the bf16 engine: shapes / variants the tcgen05 kernel does not cover yet
// (DeepViT re-attention deepvit.py:83-84, CaiT talking heads cait.py:121-127, dim_head != 64, 1-row class-attention
// queries).  The score tensor is still materialised (fp32, like the reference does), but
//   * QK^T and PV run on the tensor cores through mma.sync.m16n8k16 (legacy HMMA path: simple, any dh % 16 == 0),
//   * pre-mix -> softmax -> post-mix / LayerNorm-over-heads is ONE kernel (one read + one write of the scores
//     instead of three full passes).
// Fusing these variants into the tcgen05 kernel (head mixing as in-kernel epilogues) is the next step.
#include "attention.cuh"
#include "kernels.cuh"

#include <cmath>

namespace vb {
namespace {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int MAXDH = 128;
constexpr int PITCH = MAXDH + 8;   // bf16 elements; +8 keeps the fragment loads bank-conflict free

// S[bh, i, j] = scale * sum_d q[b,i,h,d] k[b,j,h,d];  block: 64 x 64 tile, 4 warps x (16 rows x 64 cols)
__global__ void __launch_bounds__(128)
scores_mma_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk, float* __restrict__ S,
                  int heads, int nq, int nk, int dh, float scale) {
  __shared__ __align__(16) __nv_bfloat16 Qs[64][PITCH];
  __shared__ __align__(16) __nv_bfloat16 Ks[64][PITCH];
  const int bh = blockIdx.z, b = bh / heads, h = bh % heads;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vec = dh >> 3;                                        // 16-byte vectors per row
  for (int e = threadIdx.x; e < 64 * vec; e += 128) {
    const int r = e / vec, c = (e % vec) * 8;
    uint4 vq = make_uint4(0, 0, 0, 0), vk = make_uint4(0, 0, 0, 0);
    if (i0 + r < nq) vq = *reinterpret_cast<const uint4*>(q + (static_cast<size_t>(b) * nq + i0 + r) * ldq + h * dh + c);
    if (j0 + r < nk) vk = *reinterpret_cast<const uint4*>(k + (static_cast<size_t>(b) * nk + j0 + r) * ldk + h * dh + c);
    *reinterpret_cast<uint4*>(&Qs[r][c]) = vq;
    *reinterpret_cast<uint4*>(&Ks[r][c]) = vk;
  }
  __syncthreads();
  float acc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
  const int ar = warp * 16 + (lane >> 2), ac = 2 * (lane & 3);
  for (int kk = 0; kk < dh; kk += 16) {
    uint32_t a[4];
    a[0] = *reinterpret_cast<const uint32_t*>(&Qs[ar][kk + ac]);
    a[1] = *reinterpret_cast<const uint32_t*>(&Qs[ar + 8][kk + ac]);
    a[2] = *reinterpret_cast<const uint32_t*>(&Qs[ar][kk + ac + 8]);
    a[3] = *reinterpret_cast<const uint32_t*>(&Qs[ar + 8][kk + ac + 8]);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      uint32_t bb[2];
      bb[0] = *reinterpret_cast<const uint32_t*>(&Ks[n * 8 + (lane >> 2)][kk + ac]);
      bb[1] = *reinterpret_cast<const uint32_t*>(&Ks[n * 8 + (lane >> 2)][kk + ac + 8]);
      mma_bf16_16816(acc[n], a, bb);
    }
  }
  const int r0 = i0 + ar;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c0 = j0 + n * 8 + ac;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = r0 + half * 8;
      if (r < nq) {
        float* dst = S + (static_cast<size_t>(bh) * nq + r) * nk + c0;
        if (c0 < nk) dst[0] = acc[n][2 * half] * scale;
        if (c0 + 1 < nk) dst[1] = acc[n][2 * half + 1] * scale;
      }
    }
  }
}

// out[b,i,h,d] = sum_j P[bh,i,j] v[b,j,h,d];  block: 64 query rows x dh, keys in chunks of 64.
// SPLIT: P is fed as bf16 hi + bf16 lo (two MMAs, ~2^-17 relative): DeepViT's re-attention weights are LayerNorm
// outputs of magnitude up to 1/sqrt(eps) with mixed signs, and a single bf16 rounding of them costs ~1e-1 absolute
// on the output, outside the bf16 tolerance of the parity tests.
template <bool SPLIT>
__global__ void __launch_bounds__(128)
pv_mma_kernel(const float* __restrict__ P, const __nv_bfloat16* __restrict__ v, int ldv, __nv_bfloat16* __restrict__ out, int ldo,
              int heads, int nq, int nk, int dh) {
  __shared__ __align__(16) __nv_bfloat16 Ps[64][64 + 8];
  __shared__ __align__(16) __nv_bfloat16 Pl[SPLIT ? 64 : 1][64 + 8];
  __shared__ __align__(16) __nv_bfloat16 Vt[MAXDH][64 + 8];        // transposed: Vt[d][j]
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int i0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = dh >> 3;
  float acc[MAXDH / 8][4];
#pragma unroll
  for (int n = 0; n < MAXDH / 8; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
  const int ar = warp * 16 + (lane >> 2), ac = 2 * (lane & 3);
  for (int j0 = 0; j0 < nk; j0 += 64) {
    for (int e = threadIdx.x; e < 64 * 64; e += 128) {
      const int r = e >> 6, c = e & 63;
      float p = 0.f;
      if (i0 + r < nq && j0 + c < nk) p = P[(static_cast<size_t>(bh) * nq + i0 + r) * nk + j0 + c];
      const __nv_bfloat16 hi = __float2bfloat16_rn(p);
      Ps[r][c] = hi;
      if (SPLIT) Pl[r][c] = __float2bfloat16_rn(p - __bfloat162float(hi));
    }
    for (int e = threadIdx.x; e < 64 * dh; e += 128) {
      const int r = e / dh, c = e % dh;
      __nv_bfloat16 x = __float2bfloat16_rn(0.f);
      if (j0 + r < nk) x = v[(static_cast<size_t>(b) * nk + j0 + r) * ldv + h * dh + c];
      Vt[c][r] = x;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 64; kk += 16) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(&Ps[ar][kk + ac]);
      a[1] = *reinterpret_cast<const uint32_t*>(&Ps[ar + 8][kk + ac]);
      a[2] = *reinterpret_cast<const uint32_t*>(&Ps[ar][kk + ac + 8]);
      a[3] = *reinterpret_cast<const uint32_t*>(&Ps[ar + 8][kk + ac + 8]);
      uint32_t al[4] = {0, 0, 0, 0};
      if (SPLIT) {
        al[0] = *reinterpret_cast<const uint32_t*>(&Pl[ar][kk + ac]);
        al[1] = *reinterpret_cast<const uint32_t*>(&Pl[ar + 8][kk + ac]);
        al[2] = *reinterpret_cast<const uint32_t*>(&Pl[ar][kk + ac + 8]);
        al[3] = *reinterpret_cast<const uint32_t*>(&Pl[ar + 8][kk + ac + 8]);
      }
#pragma unroll
      for (int n = 0; n < MAXDH / 8; ++n) {
        if (n < ntiles) {
          uint32_t bb[2];
          bb[0] = *reinterpret_cast<const uint32_t*>(&Vt[n * 8 + (lane >> 2)][kk + ac]);
          bb[1] = *reinterpret_cast<const uint32_t*>(&Vt[n * 8 + (lane >> 2)][kk + ac + 8]);
          mma_bf16_16816(acc[n], a, bb);
          if (SPLIT) mma_bf16_16816(acc[n], al, bb);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int n = 0; n < MAXDH / 8; ++n) {
    if (n < ntiles) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = i0 + ar + half * 8;
        if (r < nq) {
          __nv_bfloat162 o = __floats2bfloat162_rn(acc[n][2 * half], acc[n][2 * half + 1]);
          *reinterpret_cast<__nv_bfloat162*>(out + (static_cast<size_t>(b) * nq + r) * ldo + h * dh + n * 8 + ac) = o;
        }
      }
    }
  }
}

constexpr int MIX_MAX_HEADS = 32;
// One block per (query row i, batch b): all heads' score rows in shared memory.
//   variant 2: pre-softmax head mix -> softmax -> post-softmax head mix      (cait.py:123-125)
//   variant 1: softmax -> head mix -> LayerNorm across heads (eps 1e-3)       (deepvit.py:80-84)
//   variant 0: softmax
__global__ void __launch_bounds__(256)
mid_fused_kernel(float* __restrict__ S, const float* __restrict__ mix_a, const float* __restrict__ mix_b, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int heads, int nq, int nk, int variant) {
  extern __shared__ float buf[];                    // [heads][nk] then 2 x [heads*heads] mix matrices
  float* Wa = buf + heads * nk;
  float* Wb = Wa + heads * heads;
  const int i = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const size_t plane = static_cast<size_t>(nq) * nk;
  float* base = S + (static_cast<size_t>(b) * heads * nq + i) * nk;        // head h row at base + h * plane
  for (int e = threadIdx.x; e < heads * heads; e += blockDim.x) {
    Wa[e] = mix_a ? mix_a[e] : 0.f;
    Wb[e] = mix_b ? mix_b[e] : 0.f;
  }
  for (int h = 0; h < heads; ++h)
    for (int j = threadIdx.x; j < nk; j += blockDim.x) buf[h * nk + j] = base[h * plane + j];
  __syncthreads();
  auto mix = [&](const float* W, bool ln) {
    for (int j = threadIdx.x; j < nk; j += blockDim.x) {
      float x[MIX_MAX_HEADS], y[MIX_MAX_HEADS];
      for (int h = 0; h < heads; ++h) x[h] = buf[h * nk + j];
      for (int g = 0; g < heads; ++g) {
        float a = 0.f;
        for (int h = 0; h < heads; ++h) a = fmaf(x[h], W[h * heads + g], a);
        y[g] = a;
      }
      if (ln) {
        float mean = 0.f;
        for (int g = 0; g < heads; ++g) mean += y[g];
        mean /= heads;
        float var = 0.f;
        for (int g = 0; g < heads; ++g) { const float d = y[g] - mean; var += d * d; }
        const float rstd = rsqrtf(var / heads + 1e-3f);
        for (int g = 0; g < heads; ++g) y[g] = (y[g] - mean) * rstd * gamma[g] + beta[g];
      }
      for (int g = 0; g < heads; ++g) buf[g * nk + j] = y[g];
    }
  };
  if (variant == 2) { mix(Wa, false); __syncthreads(); }
  for (int h = warp; h < heads; h += nwarps) {      // softmax over keys, one warp per head row
    float* r = buf + h * nk;
    float mx = -INFINITY;
    for (int j = lane; j < nk; j += 32) mx = fmaxf(mx, r[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < nk; j += 32) { const float e = expf(r[j] - mx); r[j] = e; sum += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int j = lane; j < nk; j += 32) r[j] *= inv;
  }
  __syncthreads();
  if (variant == 1) { mix(Wa, true); __syncthreads(); }
  if (variant == 2) { mix(Wb, false); __syncthreads(); }
  for (int h = 0; h < heads; ++h)
    for (int j = threadIdx.x; j < nk; j += blockDim.x) base[h * plane + j] = buf[h * nk + j];
}

}  // namespace

bool attention_generic_mma(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                           __nv_bfloat16* out, int ldo, float* S, int B, int nq, int nk, int heads, int dh, int variant,
                           const float* mix_a, const float* mix_b, const float* ln_gamma, const float* ln_beta, cudaStream_t s) {
  if (dh % 16 != 0 || dh > MAXDH || heads > MIX_MAX_HEADS) return false;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 2)) return false;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k)) % 16) return false;
  const size_t smem = (static_cast<size_t>(heads) * nk + 2 * heads * heads) * sizeof(float);
  if (smem > 200 * 1024) return false;
  static bool configured = false;
  if (!configured) {
    VB_CUDA(cudaFuncSetAttribute(mid_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  scores_mma_kernel<<<dim3((nk + 63) / 64, (nq + 63) / 64, B * heads), 128, 0, s>>>(q, ldq, k, ldk, S, heads, nq, nk, dh, scale);
  VB_CUDA(cudaGetLastError());
  mid_fused_kernel<<<dim3(nq, B), 256, smem, s>>>(S, mix_a, mix_b, ln_gamma, ln_beta, heads, nq, nk, variant);
  VB_CUDA(cudaGetLastError());
  if (variant == 1) pv_mma_kernel<true><<<dim3((nq + 63) / 64, B * heads), 128, 0, s>>>(S, v, ldv, out, ldo, heads, nq, nk, dh);
  else pv_mma_kernel<false><<<dim3((nq + 63) / 64, B * heads), 128, 0, s>>>(S, v, ldv, out, ldo, heads, nq, nk, dh);
  VB_CUDA(cudaGetLastError());
  count_launch(3);
  return true;
}

}  // namespace vb
