// Second-tier attention path for the bf16 engine: shapes / variants the tcgen05 kernel does not cover yet
// (DeepViT re-attention deepvit.py:83-84, CaiT talking heads cait.py:121-127, dim_head != 64, 1-row class-attention
// queries).  The score tensor is still materialised (fp32, like the reference does), but
//   * QK^T and PV run on the tensor cores through mma.sync.m16n8k16 (legacy HMMA path: simple, any dh % 16 == 0),
//   * pre-mix -> softmax -> post-mix / LayerNorm-over-heads is ONE kernel (one read + one write of the scores
//     instead of three full passes).
// Fusing these variants into the tcgen05 kernel (head mixing as in-kernel epilogues) is the next step.
#include "attention.cuh"
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

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


// ------------------------------------------------------------------------------------------ fused variant kernel
// One block = 16 query rows of one image, ALL heads: the [heads, 16, nk] score block lives in shared memory, so the
// cross-head steps (CaiT talking heads cait.py:123-125, DeepViT re-attention + LayerNorm over heads deepvit.py:83-84)
// are local, and nothing of the [b,h,n,n] tensor ever reaches HBM.
//   phase 1  for each head: K_h -> smem, S_h = scale * Q_h K_h^T by mma.sync (warps split the key tiles)
//   phase 2  pre-mix (v2) -> softmax -> post-mix (v2) / mix + LN over heads (v1); probabilities rewritten in place as
//            bf16 hi|lo pairs (fp32 word -> two bf16 halves) so that phase 3 can feed P at ~2^-17 precision
//   phase 3  for each head: V_g^T -> smem, O_g = P_g V_g by mma.sync (one 8-wide d tile per warp)
constexpr int FV_ROWS = 16;
constexpr int FV_THREADS = 256;

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

template <int H>   // H = heads when known at compile time (registers instead of local arrays), 0 = run-time
__global__ void __launch_bounds__(FV_THREADS)
attn_variant_fused_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                          const __nv_bfloat16* __restrict__ v, int ldv, __nv_bfloat16* __restrict__ out, int ldo,
                          const float* __restrict__ mix_a, const float* __restrict__ mix_b, const float* __restrict__ gamma,
                          const float* __restrict__ beta, int heads, int nq, int nk, int dh, int variant, int spitch, float scale) {
  extern __shared__ __align__(16) uint8_t fv_smem[];
  float* Sb = reinterpret_cast<float*>(fv_smem);                                  // [heads][16][spitch]
  const int nk16 = (nk + 15) & ~15;
  const int kvpitch = dh + 8;                                                     // K staging: [nk][dh+8] bf16
  const int vtpitch = nk16 + 8;                                                   // V staging: [dh][nk16+8] bf16 (transposed)
  __nv_bfloat16* KV = reinterpret_cast<__nv_bfloat16*>(Sb + static_cast<size_t>(heads) * FV_ROWS * spitch);
  float* Wa = reinterpret_cast<float*>(KV + max(nk16 * kvpitch, dh * vtpitch));
  float* Wb = Wa + heads * heads;
  const int b = blockIdx.y, i0 = blockIdx.x * FV_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = FV_THREADS / 32;
  const int fr = lane >> 2, fc = 2 * (lane & 3);                                  // fragment row / column pair

  for (int e = threadIdx.x; e < heads * heads; e += FV_THREADS) {
    Wa[e] = mix_a ? mix_a[e] : 0.f;
    Wb[e] = mix_b ? mix_b[e] : 0.f;
  }
  // ---------------------------------------------------------------- phase 1: scores
  const int ntiles = (nk + 7) >> 3;
  for (int h = 0; h < heads; ++h) {
    __syncthreads();
    const int vec = dh >> 3;
    for (int e = threadIdx.x; e < nk16 * vec; e += FV_THREADS) {
      const int r = e / vec, c = (e % vec) * 8;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (r < nk) val = *reinterpret_cast<const uint4*>(k + (static_cast<size_t>(b) * nk + r) * ldk + h * dh + c);
      *reinterpret_cast<uint4*>(KV + r * kvpitch + c) = val;
    }
    __syncthreads();
    uint32_t qa[MAXDH / 16][4];
    {
      const int r0 = i0 + fr, r1 = r0 + 8;
      const __nv_bfloat16* q0 = q + (static_cast<size_t>(b) * nq + r0) * ldq + h * dh;
      const __nv_bfloat16* q1 = q + (static_cast<size_t>(b) * nq + r1) * ldq + h * dh;
#pragma unroll
      for (int ks = 0; ks < MAXDH / 16; ++ks) {
        if (ks * 16 < dh) {
          qa[ks][0] = r0 < nq ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + fc) : 0u;
          qa[ks][1] = r1 < nq ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + fc) : 0u;
          qa[ks][2] = r0 < nq ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + fc + 8) : 0u;
          qa[ks][3] = r1 < nq ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + fc + 8) : 0u;
        }
      }
    }
    float* Sh = Sb + static_cast<size_t>(h) * FV_ROWS * spitch;
    for (int nt = warp; nt < ntiles; nt += nwarps) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < MAXDH / 16; ++ks) {
        if (ks * 16 < dh) {
          uint32_t bb[2];
          const __nv_bfloat16* kr = KV + (nt * 8 + fr) * kvpitch + ks * 16 + fc;
          bb[0] = *reinterpret_cast<const uint32_t*>(kr);
          bb[1] = *reinterpret_cast<const uint32_t*>(kr + 8);
          mma_bf16_16816(acc, qa[ks], bb);
        }
      }
      const int c0 = nt * 8 + fc;
      *reinterpret_cast<float2*>(Sh + fr * spitch + c0) = make_float2(acc[0] * scale, acc[1] * scale);
      *reinterpret_cast<float2*>(Sh + (fr + 8) * spitch + c0) = make_float2(acc[2] * scale, acc[3] * scale);
    }
  }
  __syncthreads();
  // ---------------------------------------------------------------- phase 2: mix / softmax / mix (+LN)
  auto mix = [&](const float* W, bool ln) {
    constexpr int HM = H > 0 ? H : MIX_MAX_HEADS;
    const int hn = H > 0 ? H : heads;
    for (int e = threadIdx.x; e < FV_ROWS * nk; e += FV_THREADS) {
      const int r = e / nk, j = e % nk;
      float* col = Sb + r * spitch + j;
      float x[HM], y[HM];
#pragma unroll
      for (int h = 0; h < HM; ++h) if (h < hn) x[h] = col[static_cast<size_t>(h) * FV_ROWS * spitch];
#pragma unroll
      for (int g = 0; g < HM; ++g) y[g] = 0.f;
#pragma unroll
      for (int h = 0; h < HM; ++h) {
        if (h < hn) {
#pragma unroll
          for (int g = 0; g < HM; ++g) if (g < hn) y[g] = fmaf(x[h], W[h * hn + g], y[g]);
        }
      }
      if (ln) {
        float mean = 0.f;
#pragma unroll
        for (int g = 0; g < HM; ++g) if (g < hn) mean += y[g];
        mean /= hn;
        float var = 0.f;
#pragma unroll
        for (int g = 0; g < HM; ++g) if (g < hn) { const float d = y[g] - mean; var += d * d; }
        const float rstd = rsqrtf(var / hn + 1e-3f);
#pragma unroll
        for (int g = 0; g < HM; ++g) if (g < hn) y[g] = (y[g] - mean) * rstd * gamma[g] + beta[g];
      }
#pragma unroll
      for (int g = 0; g < HM; ++g) if (g < hn) col[static_cast<size_t>(g) * FV_ROWS * spitch] = y[g];
    }
  };
  if (variant == 2) { mix(Wa, false); __syncthreads(); }
  for (int row = warp; row < heads * FV_ROWS; row += nwarps) {                    // softmax, one warp per (head, query row)
    float* r = Sb + static_cast<size_t>(row) * spitch;
    float mx = -INFINITY;
    for (int j = lane; j < nk; j += 32) mx = fmaxf(mx, r[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < nk; j += 32) { const float e = __expf(r[j] - mx); r[j] = e; sum += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int j = lane; j < nk; j += 32) r[j] *= inv;
  }
  __syncthreads();
  if (variant == 1) { mix(Wa, true); __syncthreads(); }
  if (variant == 2) { mix(Wb, false); __syncthreads(); }
  // probabilities -> (bf16 hi | bf16 lo << 16) in place; columns >= nk of the 16-key padding become zero
  for (int e = threadIdx.x; e < heads * FV_ROWS * nk16; e += FV_THREADS) {
    const int j = e % nk16, row = e / nk16;
    float* w = Sb + static_cast<size_t>(row) * spitch + j;
    const float p = j < nk ? *w : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(p);
    const __nv_bfloat16 lo = __float2bfloat16_rn(p - __bfloat162float(hi));
    *reinterpret_cast<uint32_t*>(w) = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo)) << 16);
  }
  // ---------------------------------------------------------------- phase 3: O_g = P_g V_g
  const int dtiles = dh >> 3;
  const bool split = variant == 1;
  for (int g = 0; g < heads; ++g) {
    __syncthreads();
    for (int e = threadIdx.x; e < nk16 * dh; e += FV_THREADS) {                   // V_g transposed: Vt[d][j]
      const int j = e / dh, d = e % dh;
      __nv_bfloat16 x = __float2bfloat16_rn(0.f);
      if (j < nk) x = v[(static_cast<size_t>(b) * nk + j) * ldv + g * dh + d];
      KV[d * vtpitch + j] = x;
    }
    __syncthreads();
    if (warp < dtiles) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t* P0 = reinterpret_cast<const uint32_t*>(Sb + (static_cast<size_t>(g) * FV_ROWS + fr) * spitch);
      const uint32_t* P1 = P0 + 8 * spitch;
      for (int kk = 0; kk < nk16; kk += 16) {
        const uint32_t w00 = P0[kk + fc], w01 = P0[kk + fc + 1], w10 = P1[kk + fc], w11 = P1[kk + fc + 1];
        const uint32_t w02 = P0[kk + fc + 8], w03 = P0[kk + fc + 9], w12 = P1[kk + fc + 8], w13 = P1[kk + fc + 9];
        uint32_t a[4] = {prmt(w00, w01, 0x5410), prmt(w10, w11, 0x5410), prmt(w02, w03, 0x5410), prmt(w12, w13, 0x5410)};
        uint32_t bb[2];
        const __nv_bfloat16* vr = KV + (warp * 8 + fr) * vtpitch + kk + fc;
        bb[0] = *reinterpret_cast<const uint32_t*>(vr);
        bb[1] = *reinterpret_cast<const uint32_t*>(vr + 8);
        mma_bf16_16816(acc, a, bb);
        if (split) {
          uint32_t al[4] = {prmt(w00, w01, 0x7632), prmt(w10, w11, 0x7632), prmt(w02, w03, 0x7632), prmt(w12, w13, 0x7632)};
          mma_bf16_16816(acc, al, bb);
        }
      }
      const int r0 = i0 + fr, r1 = r0 + 8, c0 = g * dh + warp * 8 + fc;
      if (r0 < nq) *reinterpret_cast<__nv_bfloat162*>(out + (static_cast<size_t>(b) * nq + r0) * ldo + c0) = __floats2bfloat162_rn(acc[0], acc[1]);
      if (r1 < nq) *reinterpret_cast<__nv_bfloat162*>(out + (static_cast<size_t>(b) * nq + r1) * ldo + c0) = __floats2bfloat162_rn(acc[2], acc[3]);
    }
  }
}

// false when the score block of 16 query rows does not fit in shared memory
bool attention_variant_fused(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                             __nv_bfloat16* out, int ldo, int B, int nq, int nk, int heads, int dh, int variant,
                             const float* mix_a, const float* mix_b, const float* ln_gamma, const float* ln_beta, cudaStream_t s) {
  const int nk16 = (nk + 15) & ~15;
  int spitch = nk16 + 4;                                       // even, and 16 consecutive rows spread over the banks
  if ((spitch & 31) == 0) spitch += 4;
  const size_t kvbytes = static_cast<size_t>(std::max(nk16 * (dh + 8), dh * (nk16 + 8))) * 2;
  const size_t smem = static_cast<size_t>(heads) * FV_ROWS * spitch * 4 + kvbytes + 2 * heads * heads * 4;
  if (smem > 225 * 1024) return false;
  static bool configured = false;
  if (!configured) {
    VB_CUDA(cudaFuncSetAttribute(attn_variant_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    VB_CUDA(cudaFuncSetAttribute(attn_variant_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    VB_CUDA(cudaFuncSetAttribute(attn_variant_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    configured = true;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  auto kern = heads == 8 ? attn_variant_fused_kernel<8> : heads == 16 ? attn_variant_fused_kernel<16> : attn_variant_fused_kernel<0>;
  kern<<<dim3((nq + FV_ROWS - 1) / FV_ROWS, B), FV_THREADS, smem, s>>>(
      q, ldq, k, ldk, v, ldv, out, ldo, mix_a, mix_b, ln_gamma, ln_beta, heads, nq, nk, dh, variant, spitch, scale);
  VB_CUDA(cudaGetLastError());
  count_launch();
  return true;
}

}  // namespace

bool attention_generic_mma(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                           __nv_bfloat16* out, int ldo, float* S, int B, int nq, int nk, int heads, int dh, int variant,
                           const float* mix_a, const float* mix_b, const float* ln_gamma, const float* ln_beta, cudaStream_t s) {
  if (dh % 16 != 0 || dh > MAXDH || heads > MIX_MAX_HEADS) return false;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 2)) return false;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k)) % 16) return false;
  if (dh <= 64 && (ldo % 2) == 0 && getenv("VB_FUSED_VARIANT") != nullptr &&
      attention_variant_fused(q, ldq, k, ldk, v, ldv, out, ldo, B, nq, nk, heads, dh, variant, mix_a, mix_b, ln_gamma, ln_beta, s))
    return true;
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
