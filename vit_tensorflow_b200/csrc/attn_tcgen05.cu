// tcgen05 multi-head attention for sm_100a (bf16 operands, fp32 accumulation, dim_head = 64).
//
//   out[b, i, h, :] = softmax_j( scale * q[b,i,h,:] . k[b,j,h,:] ) @ v[b,j,h,:]        (vit.py:77-82)
//
// The [b,h,n,n] score tensor the reference materialises twice in memory (dots, attn) lives only in tensor memory:
// one CTA owns a 128-row query tile of one (b, h).
//   warp 4      TMA producer: Q tile, then K/V blocks of 128 keys into a 2-stage ring.  Q, K and V are read
//               straight out of the row-major projection output ([B, n, 3*h*dh] for the fused to_qkv) through
//               3-D tensor maps (box 64 x 128 x 1): the head split 'b n (h d) -> b h n d' (vit.py:74) costs
//               nothing, and rows past n are zero-filled by TMA instead of bleeding into the next image.
//   warp 5      MMA issuer: S = Q K^T (128 x 128 x 64, K-major operands) into TMEM, then PV = P V
//               (128 x 64 x 128, P K-major from shared memory, V MN-major exactly as TMA delivered it).
//   warps 0-3   softmax: thread i owns query row i (TMEM lane i): tcgen05.ld of S, running max / sum in the
//               exp2 domain, P written as bf16 into 128B-swizzled shared memory, running output row in registers
//               (O = O * alpha + PV), final 1/l normalisation and a 128-byte row store in 'b n (h d)' order
//               (the reference's merge-heads rearrange, vit.py:82).
// 2 CTAs per SM (192 TMEM columns each, ~113 KB shared memory) hide the serial S -> softmax -> PV chain.
#include "attention.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

#include <map>
#include <tuple>

namespace vb {
namespace {

constexpr int DH = 64;
constexpr int BQ = 128;          // query rows per CTA
constexpr int BKV = 128;         // keys per block
constexpr int KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * 128;        // 128 rows x 64 bf16
constexpr int P_BYTES = 2 * TILE_BYTES;      // 128 rows x 128 keys bf16 as two 64-column swizzled blocks
constexpr int ATT_THREADS = 192;
constexpr int SMEM_DATA = TILE_BYTES /*Q*/ + KV_STAGES * 2 * TILE_BYTES /*K,V*/ + P_BYTES;
constexpr int ATT_SMEM = SMEM_DATA + 1024;   // 1 KB covers the 1024-byte alignment slack and the barriers
constexpr int TMEM_COLS_ATT = 256;           // S: 128 columns, PV: 64 columns
constexpr int TM_S = 0, TM_PV = 128;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int ldo, int heads, int nq, int nk,
                float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t slack_front = base - raw;
  // barriers (128 bytes) go wherever the alignment slack leaves room
  const uint32_t bar_base = (1024u - slack_front >= 128u) ? base + SMEM_DATA : raw;
  const uint32_t sQ = base;
  const uint32_t sK0 = base + TILE_BYTES;
  const uint32_t sV0 = sK0 + KV_STAGES * TILE_BYTES;
  const uint32_t sP = sV0 + KV_STAGES * TILE_BYTES;
  const uint32_t q_full = bar_base, kv_full0 = bar_base + 8, kv_empty0 = bar_base + 24, s_full = bar_base + 40,
                 s_empty = bar_base + 48, p_full = bar_base + 56, pv_full = bar_base + 64, tmem_slot = bar_base + 72;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (nq + BQ - 1) / BQ;
  const int qt = blockIdx.x % nqt;
  const int bh = blockIdx.x / nqt;
  const int h = bh % heads, b = bh / heads;
  const int q0 = qt * BQ;
  const int nblk = (nk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full0 + 8 * s, 1); mbar_init(kv_empty0 + 8 * s, 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<TMEM_COLS_ATT>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 4) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &tmap_q, q_full, h * DH, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = static_cast<uint32_t>(j / KV_STAGES) & 1u;
        mbar_wait(kv_empty0 + 8 * st, ph ^ 1u);
        mbar_arrive_expect_tx(kv_full0 + 8 * st, 2 * TILE_BYTES);
        tma_load_3d(sK0 + st * TILE_BYTES, &tmap_k, kv_full0 + 8 * st, h * DH, j * BKV, b);
        tma_load_3d(sV0 + st * TILE_BYTES, &tmap_v, kv_full0 + 8 * st, h * DH, j * BKV, b);
      }
    }
  } else if (warp == 5) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      auto issue_s = [&](int j) {
        const int st = j % KV_STAGES;
        mbar_wait(kv_full0 + 8 * st, static_cast<uint32_t>(j / KV_STAGES) & 1u);
        if (j > 0) mbar_wait(s_empty, static_cast<uint32_t>(j - 1) & 1u);   // softmax has read S_{j-1}
        tcgen05_fence_after();
        const int valid = min(BKV, nk - j * BKV);
        const int n_mma = (valid + 15) & ~15;
        const uint32_t idesc = make_idesc_bf16(BQ, n_mma, 0, 0);
        const uint64_t dq = make_smem_desc(sQ, 16, 1024, 2);
        const uint64_t dk = make_smem_desc(sK0 + st * TILE_BYTES, 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tmem_base + TM_S, dq + 2u * k, dk + 2u * k, idesc, k != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        const int st = j % KV_STAGES;
        mbar_wait(p_full, static_cast<uint32_t>(j) & 1u);                    // P_j is in shared memory
        tcgen05_fence_after();
        const int valid = min(BKV, nk - j * BKV);
        const int ksteps = (valid + 15) >> 4;
        constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 0, 1);          // B (= V) is MN-major
        for (int s = 0; s < ksteps; ++s) {
          // A = P: 16 keys = 32 bytes inside a 64-key swizzled block; blocks are 16 KB apart
          const uint64_t dp = make_smem_desc(sP + (s >> 2) * TILE_BYTES + (s & 3) * 32, 16, 1024, 2);
          // B = V: rows are keys (the MMA K dimension); 16 keys = 2 swizzle atoms of 8 rows x 128 bytes
          const uint64_t dv = make_smem_desc(sV0 + st * TILE_BYTES + s * 2048, 8192, 1024, 2);
          umma_f16_ss(tmem_base + TM_PV, dp, dv, idesc_pv, s != 0);
        }
        umma_commit(pv_full);
        umma_commit(kv_empty0 + 8 * st);
      }
    }
  } else {
    // ===================================================================== softmax / output (warps 0-3)
    const int row_local = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int valid = min(BKV, nk - j * BKV);
      const int nchunk = (valid + 31) >> 5;
      mbar_wait(s_full, static_cast<uint32_t>(j) & 1u);
      tcgen05_fence_after();
      // pass 1: row maximum of the raw scores
      float mx = -INFINITY;
      for (int c = 0; c < nchunk; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(lane_addr + TM_S + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx * scale_log2);
      const float alpha = exp2f(m_run - m_new);   // exp2(-inf) = 0 on the first block
      // fold the previous block's PV into the running output before its TMEM / P buffers are reused
      if (j > 0) {
        mbar_wait(pv_full, static_cast<uint32_t>(j - 1) & 1u);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(lane_addr + TM_PV + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
        }
      }
      // pass 2: probabilities -> bf16 -> swizzled shared memory (A operand of the PV product)
      float rsum = 0.f;
      for (int c = 0; c < nchunk; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(lane_addr + TM_S + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = exp2f(fmaf(__uint_as_float(v[2 * i]), scale_log2, -m_new));
          float p1 = exp2f(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2, -m_new));
          if (c * 32 + 2 * i >= valid) p0 = 0.f;
          if (c * 32 + 2 * i + 1 >= valid) p1 = 0.f;
          rsum += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        const uint32_t rowp = sP + (c >> 1) * TILE_BYTES + row_local * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t slot = static_cast<uint32_t>(((c & 1) * 4 + k) ^ (row_local & 7));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + slot * 16), "r"(pk[4 * k]), "r"(pk[4 * k + 1]),
                       "r"(pk[4 * k + 2]), "r"(pk[4 * k + 3]) : "memory");
        }
      }
      l_run = fmaf(l_run, alpha, rsum);
      m_run = m_new;
      alpha_prev = alpha;
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(s_empty); mbar_arrive(p_full); }
    }
    mbar_wait(pv_full, static_cast<uint32_t>(nblk - 1) & 1u);
    tcgen05_fence_after();
    const float inv_l = 1.0f / l_run;
    const int row = q0 + row_local;
    __nv_bfloat16* orow = out + (static_cast<size_t>(b) * nq + row) * ldo + h * DH;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(lane_addr + TM_PV + c * 32, v);
      tmem_ld_wait();
      if (row < nq) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int d = c * 32 + k * 8 + 2 * i;
            const float a0 = fmaf(o[d], alpha_prev, __uint_as_float(v[k * 8 + 2 * i])) * inv_l;
            const float a1 = fmaf(o[d + 1], alpha_prev, __uint_as_float(v[k * 8 + 2 * i + 1])) * inv_l;
            pk[i] = pack_bf16x2(a0, a1);
          }
          *reinterpret_cast<uint4*>(orow + c * 32 + k * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS_ATT>(tmem_base);
  }
}

struct AttnPlan { CUtensorMap q, k, v; };
using AttnKey = std::tuple<const void*, int, const void*, int, const void*, int, int, int, int, int>;
std::map<AttnKey, AttnPlan>& plan_cache() {
  static std::map<AttnKey, AttnPlan> c;
  return c;
}

}  // namespace

template <>
bool attention_fast<__nv_bfloat16>(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                                   __nv_bfloat16* out, int ldo, int B, int nq, int nk, int heads, int dh, int variant,
                                   const float*, const float*, const float*, const float*, cudaStream_t s) {
  if (variant != 0 || dh != DH || nq < 16) return false;   // head-mixing variants and 1-row queries: generic path
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) % 16) return false;
  static bool configured = false;
  if (!configured) {
    VB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    configured = true;
  }
  AttnKey key{q, ldq, k, ldk, v, ldv, B, nq, nk, heads};
  auto& cache = plan_cache();
  auto it = cache.find(key);
  if (it == cache.end()) {
    if (cache.size() > 4096) cache.clear();
    AttnPlan p;
    const uint64_t inner = static_cast<uint64_t>(heads) * DH;
    p.q = make_tmap_3d(q, inner, nq, B, static_cast<uint64_t>(ldq) * 2, static_cast<uint64_t>(nq) * ldq * 2, DH, BQ, 1);
    p.k = make_tmap_3d(k, inner, nk, B, static_cast<uint64_t>(ldk) * 2, static_cast<uint64_t>(nk) * ldk * 2, DH, BKV, 1);
    p.v = make_tmap_3d(v, inner, nk, B, static_cast<uint64_t>(ldv) * 2, static_cast<uint64_t>(nk) * ldv * 2, DH, BKV, 1);
    it = cache.emplace(key, p).first;
  }
  const int nqt = (nq + BQ - 1) / BQ;
  const float scale_log2 = (1.0f / sqrtf(static_cast<float>(dh))) * 1.4426950408889634f;
  attn_fwd_kernel<<<B * heads * nqt, ATT_THREADS, ATT_SMEM, s>>>(it->second.q, it->second.k, it->second.v, out, ldo, heads, nq, nk,
                                                               scale_log2);
  VB_CUDA(cudaGetLastError());
  count_launch();
  return true;
}

}  // namespace vb
