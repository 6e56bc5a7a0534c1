// tcgen05 multi-head attention for sm_100a (bf16 operands, fp32 accumulation, dim_head = 64).
//
//   out[b, i, h, :] = softmax_j( scale * q[b,i,h,:] . k[b,j,h,:] ) @ v[b,j,h,:]        (vit.py:77-82)
//
// The [b,h,n,n] score tensor the reference materialises twice in memory (dots, attn) lives only in tensor memory.
// Persistent kernel, one CTA per SM, work item = (b, h, pair of 128-row query tiles); the two tiles of an item
// share every K/V block and ping-pong between the tensor core and two softmax warpgroups:
//   warp 0       TMA producer: Q tiles, then K/V blocks of 128 keys into a 3-stage ring, running ahead across
//                items.  Q, K and V are read straight out of the row-major projection output ([B, n, 3*h*dh] for
//                the fused to_qkv) through 3-D tensor maps (box 64 x 128 x 1): the head split
//                'b n (h d) -> b h n d' (vit.py:74) costs nothing, and rows past n are zero-filled by TMA instead
//                of bleeding into the next image.
//   warp 1       MMA issuer: S_t = Q_t K^T (128 x 128 x 64, K-major operands) into TMEM, PV_t = P_t V
//                (128 x 64 x 128; P K-major from shared memory, V MN-major exactly as TMA delivered it).  While
//                warpgroup t runs its softmax on S_t, the tensor core works on tile 1-t.
//   warps 4-11 / 12-19  softmax group of tile 0 / 1: 8 warps per tile; query row i (TMEM lane i) is shared by two
//                threads, one per 64-key half of every block (more warps per scheduler to hide the
//                tcgen05.ld -> FFMA -> MUFU.EX2 dependency chains; ncu showed the 1-thread-per-row form
//                latency-bound at IPC 0.5).  tcgen05.ld of S, running max / sum in the exp2 domain (row max
//                exchanged through shared memory), P written as bf16 into 128B-swizzled shared memory, running
//                output half-row in registers (O = O * alpha + PV), final 1/l normalisation and a 64-byte row
//                store in 'b n (h d)' order (the merge-heads rearrange, vit.py:82).
#include "attention.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

#include <map>
#include <tuple>

namespace vb {
namespace {

constexpr int DH = 64;
constexpr int BQ = 128;          // query rows per tile (two tiles per work item)
constexpr int BKV = 128;         // keys per block
constexpr int KV_ST = 3;
constexpr int TILE_BYTES = 128 * 128;        // 128 rows x 64 bf16
constexpr int P_BYTES = 2 * TILE_BYTES;      // 128 rows x 128 keys bf16 as two 64-column swizzled blocks
constexpr int ATT_THREADS = 640;
constexpr int SMEM_DATA = 2 * TILE_BYTES /*Q*/ + KV_ST * 2 * TILE_BYTES /*K,V*/ + 2 * P_BYTES;
constexpr int XCHG_BYTES = 2 * 2 * 2 * 128 * 4;  // [parity][tile][half][row] floats exchanged between the two threads of a row
constexpr int ATT_SMEM = SMEM_DATA + XCHG_BYTES + 256 + 1024;
constexpr int TMEM_COLS_ATT = 512;
constexpr int TM_S = 0, TM_PV = 256;         // S_t at TM_S + 128 t, PV_t at TM_PV + 64 t

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int ldo, int heads, int nq, int nk,
                int num_items, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + 2 * TILE_BYTES;
  const uint32_t sV = sK + KV_ST * TILE_BYTES;
  const uint32_t sP = sV + KV_ST * TILE_BYTES;
  const uint32_t sX = sP + 2 * P_BYTES;
  const uint32_t bars = sX + XCHG_BYTES;
  auto q_full = [&](int t) { return bars + 8u * t; };
  auto q_empty = [&](int t) { return bars + 16u + 8u * t; };
  auto kv_full = [&](int s) { return bars + 32u + 8u * s; };
  auto kv_empty = [&](int s) { return bars + 56u + 8u * s; };
  auto s_full = [&](int t) { return bars + 80u + 8u * t; };
  auto s_empty = [&](int t) { return bars + 96u + 8u * t; };
  auto p_full = [&](int t) { return bars + 112u + 8u * t; };
  auto pv_full = [&](int t) { return bars + 128u + 8u * t; };
  const uint32_t tmem_slot = bars + 144u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pairs = (nq + 2 * BQ - 1) / (2 * BQ);
  const int nblk = (nk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(q_full(t), 1); mbar_init(q_empty(t), 1);
      mbar_init(s_full(t), 1); mbar_init(s_empty(t), 8);
      mbar_init(p_full(t), 8); mbar_init(pv_full(t), 1);
    }
    for (int s = 0; s < KV_ST; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS_ATT>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
      uint32_t kv_cnt = 0, qcnt[2] = {0, 0};
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        const int pair = it % pairs, bh = it / pairs;
        const int h = bh % heads, b = bh / heads;
        const int row0 = pair * 2 * BQ;
        const int ntiles = (nq - row0 > BQ) ? 2 : 1;
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(q_empty(t), (qcnt[t] & 1u) ^ 1u);
          mbar_arrive_expect_tx(q_full(t), TILE_BYTES);
          tma_load_3d(sQ + t * TILE_BYTES, &tmap_q, q_full(t), h * DH, row0 + t * BQ, b);
          ++qcnt[t];
        }
        for (int j = 0; j < nblk; ++j) {
          const int st = kv_cnt % KV_ST;
          mbar_wait(kv_empty(st), ((kv_cnt / KV_ST) & 1u) ^ 1u);
          mbar_arrive_expect_tx(kv_full(st), 2 * TILE_BYTES);
          tma_load_3d(sK + st * TILE_BYTES, &tmap_k, kv_full(st), h * DH, j * BKV, b);
          tma_load_3d(sV + st * TILE_BYTES, &tmap_v, kv_full(st), h * DH, j * BKV, b);
          ++kv_cnt;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      uint32_t kv_cnt = 0;               // K/V blocks consumed before this item
      uint32_t qn[2] = {0, 0};           // Q tiles consumed per tile slot
      uint32_t sn[2] = {0, 0};           // S products issued per tile slot
      uint32_t pn[2] = {0, 0};           // PV products issued per tile slot
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        const int pair = it % pairs;
        const int row0 = pair * 2 * BQ;
        const int ntiles = (nq - row0 > BQ) ? 2 : 1;
        auto issue_s = [&](int t, int j) {
          const uint32_t c = kv_cnt + j;
          const int st = c % KV_ST;
          mbar_wait(kv_full(st), (c / KV_ST) & 1u);
          if (sn[t] > 0) mbar_wait(s_empty(t), (sn[t] - 1) & 1u);     // softmax has read the previous S_t
          tcgen05_fence_after();
          const int valid = min(BKV, nk - j * BKV);
          const uint32_t idesc = make_idesc_bf16(BQ, (valid + 15) & ~15, 0, 0);
          const uint64_t dq = make_smem_desc(sQ + t * TILE_BYTES, 16, 1024, 2);
          const uint64_t dk = make_smem_desc(sK + st * TILE_BYTES, 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tmem_base + TM_S + t * 128, dq + 2u * k, dk + 2u * k, idesc, k != 0);
          umma_commit(s_full(t));
          ++sn[t];
          if (j == nblk - 1) umma_commit(q_empty(t));                  // last use of Q_t for this item
        };
        for (int t = 0; t < ntiles; ++t) { mbar_wait(q_full(t), qn[t] & 1u); ++qn[t]; }
        for (int t = 0; t < ntiles; ++t) issue_s(t, 0);
        for (int j = 0; j < nblk; ++j) {
          const uint32_t c = kv_cnt + j;
          const int st = c % KV_ST;
          const int valid = min(BKV, nk - j * BKV);
          const int ksteps = (valid + 15) >> 4;
          constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 0, 1);    // B (= V) is MN-major
          for (int t = 0; t < ntiles; ++t) {
            mbar_wait(p_full(t), pn[t] & 1u);                              // P_t(j) is in shared memory
            tcgen05_fence_after();
            for (int s = 0; s < ksteps; ++s) {
              // A = P: 16 keys = 32 bytes inside a 64-key swizzled block; blocks are 16 KB apart
              const uint64_t dp = make_smem_desc(sP + t * P_BYTES + (s >> 2) * TILE_BYTES + (s & 3) * 32, 16, 1024, 2);
              // B = V: rows are keys (the MMA K dimension); 16 keys = 2 swizzle atoms of 8 rows x 128 bytes
              const uint64_t dv = make_smem_desc(sV + st * TILE_BYTES + s * 2048, 8192, 1024, 2);
              umma_f16_ss(tmem_base + TM_PV + t * 64, dp, dv, idesc_pv, s != 0);
            }
            umma_commit(pv_full(t));
            ++pn[t];
            if (j + 1 < nblk) issue_s(t, j + 1);
          }
          umma_commit(kv_empty(st));                                       // both tiles are done with K_j / V_j
        }
        kv_cnt += nblk;
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== softmax groups (8 warps per tile)
    const int wl = warp - 4;
    const int t = wl >> 3;                  // tile slot
    const int hf = (wl >> 2) & 1;           // which 64-key half of each block / which 32 output dims
    const int wq = warp & 3;                // TMEM lane quarter
    const int row_local = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t tS = lane_addr + TM_S + t * 128 + hf * 64, tPV = lane_addr + TM_PV + t * 64 + hf * 32;
    const uint32_t sPt = sP + t * P_BYTES + hf * TILE_BYTES + row_local * 128;
    const uint32_t pair_bar = 1 + t * 4 + wq;                              // named barrier of the two warps sharing rows
    auto xchg = [&](uint32_t par, int half) { return sX + (((par * 2 + t) * 2 + half) * 128 + row_local) * 4; };
    auto exchange = [&](uint32_t par, float mine) {                        // returns the partner thread's value
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xchg(par, hf)), "f"(mine) : "memory");
      named_bar_sync(pair_bar, 64);
      float other;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xchg(par, hf ^ 1)) : "memory");
      return other;
    };
    uint32_t sc = 0, pvc = 0, xpar = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int pair = it % pairs, bh = it / pairs;
      const int h = bh % heads, b = bh / heads;
      const int q0 = pair * 2 * BQ + t * BQ;
      if (q0 >= nq) continue;                                             // this item has a single tile
      f32x2 o[16];                                                        // running output, 32 dims as fp32 pairs
#pragma unroll
      for (int d = 0; d < 16; ++d) o[d] = 0ull;
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
      for (int j = 0; j < nblk; ++j) {
        const int valid = min(BKV, nk - j * BKV);
        const int nval = max(0, min(64, valid - hf * 64));                // valid keys among this thread's 64 columns
        const int nfull = nval >> 5;                                      // full 32-key chunks
        const int ntail = nval & 31;
        mbar_wait(s_full(t), sc & 1u);
        ++sc;
        tcgen05_fence_after();
        // pass 1: maximum of the raw scores over this thread's columns, then over the row
        float mx = -INFINITY;
        for (int c = 0; c < nfull; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        if (ntail) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS + nfull * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = (i < ntail) ? fmaxf(mx, __uint_as_float(v[i])) : mx;
        }
        mx = fmaxf(mx, exchange(xpar, mx));
        xpar ^= 1u;
        const float m_new = fmaxf(m_run, mx * scale_log2);
        const float alpha = ex2_approx(m_run - m_new);                    // ex2(-inf) = 0 on the first block
        // fold the previous block's PV into the running output before its TMEM / P buffers are reused
        if (j > 0) {
          mbar_wait(pv_full(t), pvc & 1u);
          ++pvc;
          tcgen05_fence_after();
          uint32_t v[32];
          tmem_ld_32x32b_x32(tPV, v);
          tmem_ld_wait();
          const f32x2 a2 = splat2(alpha_prev);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = fma2(o[i], a2, pack2u(v[2 * i], v[2 * i + 1]));
        }
        // pass 2: probabilities -> bf16 -> swizzled shared memory (A operand of the PV product)
        f32x2 rsum2 = 0ull;
        const f32x2 sc2 = splat2(scale_log2), nm2 = splat2(-m_new);
        auto store_chunk = [&](int c, const uint32_t (&pk)[16]) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t slot = static_cast<uint32_t>((c * 4 + k) ^ (row_local & 7));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPt + slot * 16), "r"(pk[4 * k]), "r"(pk[4 * k + 1]),
                         "r"(pk[4 * k + 2]), "r"(pk[4 * k + 3]) : "memory");
          }
        };
        for (int c = 0; c < nfull; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS + c * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0, x1;
            unpack2(fma2(pack2u(v[2 * i], v[2 * i + 1]), sc2, nm2), x0, x1);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            rsum2 = add2(rsum2, pack2(p0, p1));
            pk[i] = pack_bf16x2(p0, p1);
          }
          store_chunk(c, pk);
        }
        if (ntail) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS + nfull * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0, x1;
            unpack2(fma2(pack2u(v[2 * i], v[2 * i + 1]), sc2, nm2), x0, x1);
            float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            p0 = (2 * i < ntail) ? p0 : 0.f;
            p1 = (2 * i + 1 < ntail) ? p1 : 0.f;
            rsum2 = add2(rsum2, pack2(p0, p1));
            pk[i] = pack_bf16x2(p0, p1);
          }
          store_chunk(nfull, pk);
        }
        float rs0, rs1;
        unpack2(rsum2, rs0, rs1);
        l_run = fmaf(l_run, alpha, rs0 + rs1);
        m_run = m_new;
        alpha_prev = alpha;
        tcgen05_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(s_empty(t)); mbar_arrive(p_full(t)); }
      }
      mbar_wait(pv_full(t), pvc & 1u);
      ++pvc;
      tcgen05_fence_after();
      const float l_tot = l_run + exchange(xpar, l_run);                  // both halves of the row
      xpar ^= 1u;
      const float inv_l = 1.0f / l_tot;
      const int row = q0 + row_local;
      __nv_bfloat16* orow = out + (static_cast<size_t>(b) * nq + row) * ldo + h * DH + hf * 32;
      uint32_t v[32];
      tmem_ld_32x32b_x32(tPV, v);
      tmem_ld_wait();
      if (row < nq) {
        const f32x2 a2 = splat2(alpha_prev), il2 = splat2(inv_l);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int d = k * 4 + i;                                          // pair index: dims 2d, 2d+1
            pk[i] = pack_bf16x2_from(mul2(fma2(o[d], a2, pack2u(v[2 * d], v[2 * d + 1])), il2));
          }
          *reinterpret_cast<uint4*>(orow + k * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tcgen05_fence_before();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
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
  const int pairs = (nq + 2 * BQ - 1) / (2 * BQ);
  const int num_items = B * heads * pairs;
  const int grid = num_items < sm_count() ? num_items : sm_count();
  const float scale_log2 = (1.0f / sqrtf(static_cast<float>(dh))) * 1.4426950408889634f;
  attn_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM, s>>>(it->second.q, it->second.k, it->second.v, out, ldo, heads, nq, nk, num_items,
                                                     scale_log2);
  VB_CUDA(cudaGetLastError());
  count_launch();
  return true;
}

}  // namespace vb
