// tcgen05 multi-head attention for sm_100a (bf16 operands, fp32 accumulation, dim_head = 64).
//
//   out[b, i, h, :] = softmax_j( scale * q[b,i,h,:] . k[b,j,h,:] ) @ v[b,j,h,:]        (vit.py:77-82)
//
// The [b,h,n,n] score tensor the reference materialises twice in memory (dots, attn) lives only in tensor memory.
// Persistent kernel, one CTA per SM, work item = (b, h, pair of 128-row query tiles); the two tiles of an item
// share every K/V block and ping-pong between the tensor core and two softmax warpgroups:
//   warp 8       TMA producer: Q tiles, then K/V blocks of 128 keys into a 3-stage ring, running ahead across
//                items.  Q, K and V are read straight out of the row-major projection output ([B, n, 3*h*dh] for
//                the fused to_qkv) through 3-D tensor maps (box 64 x 128 x 1): the head split
//                'b n (h d) -> b h n d' (vit.py:74) costs nothing, and rows past n are zero-filled by TMA instead
//                of bleeding into the next image.
//   warps 10, 11  MMA issuers, one per query tile (independent pipelines): S_t = Q_t K^T (128 x 128 x 64,
//                K-major operands) into TMEM, O_t += P_t V (128 x 64 x 128; P K-major from shared memory, V MN-major
//                exactly as TMA delivered it).  S(k+1) is issued BEFORE PV(k): the softmax group keeps S in
//                registers, so the S buffer is free again long before P(k) is ready -- a clock64 trace of the
//                single-issuer form (profiles/r01_attn_trace_v5.txt) showed the softmax groups waiting 1.6-2 K
//                cycles per block for S and the issuer thread spending ~100 cycles per MMA on descriptor math.
//   warps 0-3 / 4-7   softmax group of tile 0 / 1: thread i owns query row i of its tile (TMEM lane i).
//                S is read from TMEM exactly ONCE per block (tcgen05.ld -> 128 registers: TMEM->register bandwidth,
//                not MUFU, bounded the earlier two-pass forms -- see profiles/), the S buffer is released to the
//                tensor core immediately, max / exp2 / sum run on registers with packed fp32x2 math, P goes as bf16
//                into 128B-swizzled shared memory.  O accumulates in TMEM across key blocks (MMA accumulate); the
//                softmax reference point only moves -- and O is only rescaled in TMEM -- when a block's row max
//                exceeds it by more than 2^8 (exact; FlashAttention-4-style lazy rescale).  Final 1/l normalisation
//                and a 128-byte row store in 'b n (h d)' order (the merge-heads rearrange, vit.py:82).
#include "attention.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <map>
#include <tuple>

namespace vb {
namespace {

constexpr int DH = 64;
constexpr int BQ = 128;          // query rows per tile (two tiles per work item)
constexpr int BKV = 128;         // keys per block
#ifndef VB_ATTN_KV_ST
#define VB_ATTN_KV_ST 3
#endif
constexpr int KV_ST = VB_ATTN_KV_ST;   // K/V ring depth (3: 193 KB of shared memory, 4: 225 KB)
constexpr int TILE_BYTES = 128 * 128;        // 128 rows x 64 bf16
constexpr int P_BYTES = 2 * TILE_BYTES;      // 128 rows x 128 keys bf16 as two 64-column swizzled blocks
// KS = false: 8 softmax warps (warps 0-3 / 4-7 = tile 0 / 1, thread = query row, all 128 keys of a block).
// KS = true ("key split"): 16 softmax warps -- two per (tile, TMEM lane quarter), each taking two of the block's four 32-key
// chunks (chunks hf and hf + 2), i.e. FOUR warps per sub-partition whose TMEM-load, MUFU and FP32 phases overlap instead of two
// running them one after the other; the two threads of a row exchange their half row max (and, once per item, their half row
// sum) through shared memory and a 64-thread named barrier.  Then: the TMA producer warp, a spare, the two MMA issuers.
template <bool KS>
struct AttCfg {
  static constexpr int NSOFT = KS ? 16 : 8;
  static constexpr int PRODUCER_WARP = NSOFT, MMA_WARP0 = NSOFT + 2;   // issue arbiter favours high warp ids: issuers on top
  static constexpr int THREADS = (NSOFT + 4) * 32;
};
constexpr int SMEM_DATA = 2 * TILE_BYTES /*Q*/ + KV_ST * 2 * TILE_BYTES /*K,V*/ + 2 * P_BYTES;
constexpr int XCH_BYTES = 3 * 2 * 2 * 128 * 4;     // KS: exchange slots [row max (2 block parities) | row sum][tile][half][row]
constexpr int ATT_SMEM = SMEM_DATA + XCH_BYTES + 256 + 1024;
constexpr int TMEM_COLS_ATT = 512;
constexpr int TM_S = 0, TM_PV = 256;         // S_t at TM_S + 128 t, PV_t at TM_PV + 64 t

// 2^t on the FMA / ALU pipes instead of MUFU (the FlashAttention-4 split): t = n + f with n = round(t) through the 1.5 * 2^23
// magic-number add, 2^f on [-1/2, 1/2] as a degree-3 minimax polynomial (max relative error 7.5e-5, far below the 2^-9 rounding of
// P to bf16 that follows), and the 2^n scaling as an integer add on the exponent field ((bits(r) << 23) carries n modulo 2^9 into
// bits 23-31, which wraps correctly for negative n).  t <= 8 by the lazy-rescale invariant; t is clamped at -126 (result ~1e-38).
// Per PAIR of elements: 2 FMNMX + 3 FADD2 + 3 FFMA2 + 2 SHL + 2 IADD, against 2 MUFU.EX2 (16 issue cycles of the 4-lane
// MUFU unit per 64 elements): the two paths run on different pipes, so splitting the 8-key groups between them shortens the
// exp phase (profiles/r01_attn_item_timeline.md: 4.8 K of 10.4 K cycles at 1.3x the MUFU floor).  MEASURED (round 2,
// profiles/r02_ab_attn_poly.txt, ViT-B/16 B = 256 shapes, one box): mask 0x0 117.2 us, 0x8 (25 %) 116.4 us, 0xA (50 %) 121.2 us,
// 0xE (75 %) 131.6 us -- the kernel is not MUFU-bound (nor issue-bound: ~1400 warp instructions per sub-partition and item in
// 10.4 K cycles); the serial S -> softmax -> P -> PV round trips of the two warps per sub-partition are.  Default: off.
#ifndef VB_ATTN_KEYSPLIT
#define VB_ATTN_KEYSPLIT 0                 // default form of the kernel (env VB_ATTN_KEYSPLIT / VB_ATTN_NO_KEYSPLIT flips it)
#endif
#ifndef VB_ATTN_COND_LD
#define VB_ATTN_COND_LD 0            // 1: skip TMEM loads of 32-key chunks past `valid` (measured slower: 36 more spilled registers)
#endif
#ifndef VB_ATTN_PINGPONG
#define VB_ATTN_PINGPONG 1
#endif
#ifndef VB_ATTN_POLY_MASK
#define VB_ATTN_POLY_MASK 0x0            // bit k: the k-th 8-key group of every 32-key chunk takes the polynomial path
#endif
__device__ __forceinline__ f32x2 ex2_poly2(f32x2 t) {
  float t0, t1;
  unpack2(t, t0, t1);
  t = pack2(fmaxf(t0, -126.0f), fmaxf(t1, -126.0f));
  const f32x2 r = add2(t, splat2(12582912.0f));
  const f32x2 fl = add2(r, splat2(-12582912.0f));
  const f32x2 f = fma2(fl, splat2(-1.0f), t);
  f32x2 q = fma2(splat2(0.055171649903059006f), f, splat2(0.2426111251115799f));
  q = fma2(q, f, splat2(0.6932609677314758f));
  q = fma2(q, f, splat2(0.9999280571937561f));
  const uint32_t lo = static_cast<uint32_t>(q) + (static_cast<uint32_t>(r) << 23);
  const uint32_t hi = static_cast<uint32_t>(q >> 32) + (static_cast<uint32_t>(r >> 32) << 23);
  return pack2u(lo, hi);
}

template <bool KS>
__global__ void __launch_bounds__(AttCfg<KS>::THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o, int heads, int nq, int nk,
                int num_items, float scale_log2, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + 2 * TILE_BYTES;
  const uint32_t sV = sK + KV_ST * TILE_BYTES;
  const uint32_t sP = sV + KV_ST * TILE_BYTES;
  constexpr int ATT_PRODUCER_WARP = AttCfg<KS>::PRODUCER_WARP, ATT_MMA_WARP0 = AttCfg<KS>::MMA_WARP0;
  const uint32_t sX = sP + 2 * P_BYTES;                          // KS exchange slots
  const uint32_t bars = sX + XCH_BYTES;
  auto q_full = [&](int t) { return bars + 8u * t; };
  auto q_empty = [&](int t) { return bars + 16u + 8u * t; };
  auto s_full = [&](int t) { return bars + 32u + 8u * t; };
  auto s_empty = [&](int t) { return bars + 48u + 8u * t; };
  auto p_full = [&](int t) { return bars + 64u + 8u * t; };
  auto pv_full = [&](int t) { return bars + 80u + 8u * t; };
  const uint32_t tmem_slot = bars + 96u;
  auto kv_full = [&](int s) { return bars + 104u + 8u * s; };
  auto kv_empty = [&](int s) { return bars + 104u + 8u * KV_ST + 8u * s; };   // <= 104 + 64 = 168 < 256 reserved bytes

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pairs = (nq + 2 * BQ - 1) / (2 * BQ);
  const int nblk = (nk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(q_full(t), 1); mbar_init(q_empty(t), 1);
      mbar_init(s_full(t), 1); mbar_init(s_empty(t), AttCfg<KS>::NSOFT / 2);
      mbar_init(p_full(t), AttCfg<KS>::NSOFT / 2); mbar_init(pv_full(t), 1);
    }
    for (int s = 0; s < KV_ST; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2); }   // one release per tile issuer
    fence_mbar_init();
  }
  if (warp == ATT_MMA_WARP0) tmem_alloc<TMEM_COLS_ATT>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                 // prologue above overlapped the producer GEMM's tail; Q/K/V are complete from here on
  pdl_launch_dependents();
  // optional timeline trace (VB_ATTN_TRACE=1 through vb_op_attention): CTA 0 records (tag, clock) pairs per role
  int dbg_n = 0;
  auto trace = [&](int role, int tag) {
    if (dbg != nullptr && blockIdx.x == 0 && dbg_n < 250) {
      dbg[role * 512 + 2 * dbg_n] = tag;
      dbg[role * 512 + 2 * dbg_n + 1] = clock64();
      ++dbg_n;
    }
  };

  if (warp == ATT_PRODUCER_WARP) {
    // ===================================================================== TMA producer (warp-uniform loop, elected lane issues)
    {
      if (lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); }
      uint32_t kv_cnt = 0, qcnt[2] = {0, 0};
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        const int pair = it % pairs, bh = it / pairs;
        const int h = bh % heads, b = bh / heads;
        const int row0 = pair * 2 * BQ;
        const int ntiles = (nq - row0 > BQ) ? 2 : 1;
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(q_empty(t), (qcnt[t] & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(q_full(t), TILE_BYTES);
            tma_load_3d(sQ + t * TILE_BYTES, &tmap_q, q_full(t), h * DH, row0 + t * BQ, b);
          }
          __syncwarp();
          ++qcnt[t];
        }
        for (int j = 0; j < nblk; ++j) {
          const int st = kv_cnt % KV_ST;
          mbar_wait(kv_empty(st), ((kv_cnt / KV_ST) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(kv_full(st), 2 * TILE_BYTES);
            tma_load_3d(sK + st * TILE_BYTES, &tmap_k, kv_full(st), h * DH, j * BKV, b);
            tma_load_3d(sV + st * TILE_BYTES, &tmap_v, kv_full(st), h * DH, j * BKV, b);
          }
          __syncwarp();
          ++kv_cnt;
        }
      }
    }
  } else if (warp >= ATT_MMA_WARP0) {
    // ===================================================================== MMA issuer of tile t (warp-uniform loop)
    {
      const int t = warp - ATT_MMA_WARP0;
      const uint32_t tS_d = tmem_base + TM_S + t * 128, tO_d = tmem_base + TM_PV + t * 64;
      const uint64_t dq = make_smem_desc(sQ + t * TILE_BYTES, 16, 1024, 2);
      const uint64_t dk0 = make_smem_desc(sK, 16, 1024, 2);
      const uint64_t dp = make_smem_desc(sP + t * P_BYTES, 16, 1024, 2);
      const uint64_t dv0 = make_smem_desc(sV, 8192, 1024, 2);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 0, 1);       // B (= V) is MN-major
      constexpr uint32_t ST16 = TILE_BYTES >> 4;                           // descriptor address units per stage / P block
      uint32_t sn = 0, pn = 0, qn = 0;
      // flat (item, block) cursor for S products; `kv` = global K/V block counter of the cursor's block
      struct Cur { int it; int j; uint32_t kv; bool ok; };
      auto has_tile = [&](int it) { return (nq - (it % pairs) * 2 * BQ) > t * BQ; };
      auto first = [&]() {
        Cur c{static_cast<int>(blockIdx.x), 0, 0u, false};
        while (c.it < num_items && !has_tile(c.it)) { c.it += gridDim.x; c.kv += nblk; }
        c.ok = c.it < num_items;
        return c;
      };
      auto advance = [&](Cur c) {
        ++c.kv;
        if (++c.j == nblk) {
          c.j = 0;
          c.it += gridDim.x;
          while (c.it < num_items && !has_tile(c.it)) { c.it += gridDim.x; c.kv += nblk; }
        }
        c.ok = c.it < num_items;
        return c;
      };
      // Issue S for cursor c.  blocking = false: only if Q / K are already resident (the run-ahead before PV(k) must
      // never wait on loads that themselves wait for PV(k)'s stage release -- that deadlocked at n = 577).
      auto issue_s = [&](const Cur& c, bool blocking) -> bool {
        const uint32_t st = c.kv % KV_ST;
        if (!blocking) {
          bool ready = mbar_try_wait(kv_full(st), (c.kv / KV_ST) & 1u);
          if (c.j == 0) ready = ready && mbar_try_wait(q_full(t), qn & 1u);
          if (!__all_sync(0xffffffffu, ready)) return false;
        }
        if (c.j == 0) { mbar_wait(q_full(t), qn & 1u); ++qn; }
        mbar_wait(kv_full(st), (c.kv / KV_ST) & 1u);
        if (sn > 0) mbar_wait(s_empty(t), (sn - 1) & 1u);                  // softmax holds the previous S_t in registers
        tcgen05_fence_after();
        const int valid = min(BKV, nk - c.j * BKV);
        const uint32_t idesc = make_idesc_bf16(BQ, (valid + 15) & ~15, 0, 0);
        const uint64_t dk = dk0 + st * ST16;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_f16_ss(tS_d, dq + 2u * k, dk + 2u * k, idesc, k != 0);
          umma_commit(s_full(t));
          if (c.j == nblk - 1) umma_commit(q_empty(t));                    // last use of Q_t for this item
          trace(t == 0 ? 0 : 3, 100 + c.j);
        }
        __syncwarp();
        ++sn;
        return true;
      };
      Cur cur = first();
      if (cur.ok) issue_s(cur, true);
      while (cur.ok) {
        const Cur nxt = advance(cur);
        const bool early = nxt.ok && issue_s(nxt, false);                  // S(k+1) before PV(k) when its inputs are resident
        const uint32_t st = cur.kv % KV_ST;
        const int valid = min(BKV, nk - cur.j * BKV);
        const int ksteps = (valid + 15) >> 4;
        mbar_wait(p_full(t), pn & 1u);                                     // P_t(k) is in shared memory
        ++pn;
        tcgen05_fence_after();
        const uint64_t dv = dv0 + st * ST16;
        const bool solo = (t == 0) && (nq - (cur.it % pairs) * 2 * BQ) <= BQ;   // item without a second tile
        if (elect_one()) {
#pragma unroll
          for (int s8 = 0; s8 < BKV / 16; ++s8) {
            // A = P: 16 keys = 32 bytes inside a 64-key swizzled block (blocks 16 KB apart);
            // B = V: 16 keys = 2 swizzle atoms of 8 rows x 128 bytes
            if (s8 < ksteps)
              umma_f16_ss(tO_d, dp + ((s8 >> 2) * ST16 + (s8 & 3) * 2), dv + s8 * 128u, idesc_pv, (cur.j | s8) != 0);
          }
          umma_commit(pv_full(t));
          umma_commit(kv_empty(st));
          if (solo) umma_commit(kv_empty(st));
          trace(t == 0 ? 0 : 3, 400 + cur.j);
        }
        __syncwarp();
        if (nxt.ok && !early) issue_s(nxt, true);
        cur = nxt;
      }
    }
  } else if (KS) {
    // ===================================================================== softmax, key-split form (2 warps per tile and lane quarter)
    if (warp >= 16) goto done;              // spare warp
    const int t = warp >> 3, hf = (warp >> 2) & 1, wq = warp & 3;
    const int row_local = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t tS = lane_addr + TM_S + t * 128, tPV = lane_addr + TM_PV + t * 64 + hf * 32;
    const uint32_t sPt = sP + t * P_BYTES + row_local * 128;
    const uint32_t pair_bar = 1 + t * 4 + wq;                 // named barrier of the two warps that share these 32 rows
    auto xslot = [&](int buf, int half) { return sX + static_cast<uint32_t>(((buf * 2 + t) * 2 + half) * 128 + row_local) * 4u; };
    uint32_t sc = 0, pvc = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int pair = it % pairs, bh = it / pairs;
      const int h = bh % heads, b = bh / heads;
      const int q0 = pair * 2 * BQ + t * BQ;
      if (q0 >= nq) continue;                                             // this item has a single tile
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nblk; ++j) {
        const int valid = min(BKV, nk - j * BKV);
        mbar_wait(s_full(t), sc & 1u);
        ++sc;
        tcgen05_fence_after();
        uint32_t va[32], vb[32];                                          // chunks hf and hf + 2 of this row's 128 scores
        tmem_ld_32x32b_x32(tS + hf * 32, va);
        tmem_ld_32x32b_x32(tS + (hf + 2) * 32, vb);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(t));
        auto chunk_max = [&](int c, const uint32_t (&v)[32]) -> float {
          const int nv = valid - c * 32;
          float a = -INFINITY, b2 = -INFINITY;
          if (nv >= 32) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              a = fmaxf(a, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
              b2 = fmaxf(b2, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
            }
          } else if (nv > 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              a = (i < nv) ? fmaxf(a, __uint_as_float(v[i])) : a;
              b2 = (i + 1 < nv) ? fmaxf(b2, __uint_as_float(v[i + 1])) : b2;
            }
          }
          return fmaxf(a, b2);
        };
        float mx = fmaxf(chunk_max(hf, va), chunk_max(hf + 2, vb));
        if (j == 0 && hf == 0) {                                          // previous item's output slab (same bytes as P) drained
          if (lane == 0) bulk_wait_group_read<0>();
          __syncwarp();
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(xslot(j & 1, hf)), "f"(mx) : "memory");
        named_bar_sync(pair_bar, 64);
        {
          float mo;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mo) : "r"(xslot(j & 1, hf ^ 1)) : "memory");
          mx = fmaxf(mx, mo);
        }
        const float m_blk = mx * scale_log2;
        float alpha = 1.0f;
        const bool need = m_blk > m_ref + 8.0f;                           // identical in both threads of the row
        if (need) { alpha = ex2_approx(m_ref - m_blk); m_ref = m_blk; l_run *= alpha; }
        if (j > 0) {
          mbar_wait(pv_full(t), pvc & 1u);
          ++pvc;
          if (__any_sync(0xffffffffu, need)) {                            // each half rescales its 32 output columns
            tcgen05_fence_after();
            const f32x2 a2 = splat2(alpha);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint32_t ov[16];
              tmem_ld_32x32b_x16(tPV + g * 16, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float x0, x1;
                unpack2(mul2(pack2u(ov[2 * i], ov[2 * i + 1]), a2), x0, x1);
                ov[2 * i] = __float_as_uint(x0);
                ov[2 * i + 1] = __float_as_uint(x1);
              }
              tmem_st_32x32b_x16(tPV + g * 16, ov);
            }
            tmem_st_wait();
          }
        }
        f32x2 rsum2 = 0ull;
        const f32x2 sc2 = splat2(scale_log2), nm2 = splat2(-m_ref);
        auto emit = [&](int c, const uint32_t (&v)[32]) {                 // chunk c = keys [32c, 32c+32) of the block
          const int nv = valid - c * 32;
          if (nv <= 0) return;
          const uint32_t rowp = sPt + (c >> 1) * TILE_BYTES;
          if (nv >= 32) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint32_t pk[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int e = k * 8 + 2 * i;
                float x0, x1;
                unpack2(fma2(pack2u(v[e], v[e + 1]), sc2, nm2), x0, x1);
                const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                rsum2 = add2(rsum2, pack2(p0, p1));
                pk[i] = pack_bf16x2(p0, p1);
              }
              const uint32_t slot = static_cast<uint32_t>(((c & 1) * 4 + k) ^ (row_local & 7));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + slot * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                           "r"(pk[3]) : "memory");
            }
            return;
          }
          const int nv16 = (nv + 15) & ~15;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k * 8 >= nv16) break;
            uint32_t pk[4] = {0u, 0u, 0u, 0u};
            if (k * 8 < nv) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int e = k * 8 + 2 * i;
                float x0, x1;
                unpack2(fma2(pack2u(v[e], v[e + 1]), sc2, nm2), x0, x1);
                const float p0 = (e < nv) ? ex2_approx(x0) : 0.f;
                const float p1 = (e + 1 < nv) ? ex2_approx(x1) : 0.f;
                rsum2 = add2(rsum2, pack2(p0, p1));
                pk[i] = pack_bf16x2(p0, p1);
              }
            }
            const uint32_t slot = static_cast<uint32_t>(((c & 1) * 4 + k) ^ (row_local & 7));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + slot * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                         "r"(pk[3]) : "memory");
          }
        };
        emit(hf, va);
        emit(hf + 2, vb);
        float rs0, rs1;
        unpack2(rsum2, rs0, rs1);
        l_run += rs0 + rs1;
        tcgen05_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(t));
      }
      mbar_wait(pv_full(t), pvc & 1u);
      ++pvc;
      tcgen05_fence_after();
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xslot(2, hf)), "f"(l_run) : "memory");
      named_bar_sync(pair_bar, 64);
      float lo;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lo) : "r"(xslot(2, hf ^ 1)) : "memory");
      const f32x2 il2 = splat2(1.0f / (l_run + lo));
      {
        uint32_t oa[32];                                                   // this half's 32 output columns
        tmem_ld_32x32b_x32(tPV, oa);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int d = k * 4 + i;
            pk[i] = pack_bf16x2_from(mul2(pack2u(oa[2 * d], oa[2 * d + 1]), il2));
          }
          const uint32_t slot = static_cast<uint32_t>((hf * 4 + k) ^ (row_local & 7));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPt + slot * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                       "r"(pk[3]) : "memory");
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(pair_bar, 64);                                       // both halves of the 32 rows are in the slab
      if (hf == 0 && lane == 0) {
        tma_store_3d(&tmap_o, sP + t * P_BYTES + wq * 4096, h * DH, q0 + wq * 32, b);
        bulk_commit_group();
      }
      __syncwarp();
    }
    if (hf == 0 && lane == 0) bulk_wait_group_read<0>();
  } else {
    // ===================================================================== softmax groups (4 warps per tile)
    if (warp >= 8) goto done;               // spare warp 9
    const int t = warp >> 2;                // tile slot
    const int wq = warp & 3;                // TMEM lane quarter this warp may access
    const int row_local = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t tS = lane_addr + TM_S + t * 128, tPV = lane_addr + TM_PV + t * 64;
    const uint32_t sPt = sP + t * P_BYTES + row_local * 128;
    uint32_t sc = 0, pvc = 0;
    if (VB_ATTN_PINGPONG && t == 1) named_bar_arrive(2, 256);              // tile 0 opens every item
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int pair = it % pairs, bh = it / pairs;
      const int h = bh % heads, b = bh / heads;
      const int q0 = pair * 2 * BQ + t * BQ;
      if (q0 >= nq) continue;                                             // this item has a single tile
      // both tiles of the item exist, and at most two key blocks: measured (profiles/r02_ab_attn_variants.txt, one box) 114.3 vs
      // 117.5 us at n = 197 (2 blocks) but 132.1 vs 122.1 us at n = 577 (5 blocks, where the token serialises more than it de-phases)
      const bool pingpong = VB_ATTN_PINGPONG && (pair * 2 * BQ + BQ < nq) && nblk <= 2;
      float m_ref = -INFINITY, l_run = 0.f;
      // (Letting a warp whose 32 rows all lie past nq -- rows 96-127 of every second tile at n = 197 -- skip its TMEM reads and
      //  exponentials and only keep the barrier protocol measured SLOWER: 118.6 vs 113.5 us, profiles/r02_ab_attn_deadwarp_skip.txt.)
      for (int j = 0; j < nblk; ++j) {
        const int valid = min(BKV, nk - j * BKV);
        if (wq == 0 && lane == 0) trace(1 + t, 10 + j);
        mbar_wait(s_full(t), sc & 1u);
        ++sc;
        tcgen05_fence_after();
        if (wq == 0 && lane == 0) trace(1 + t, 20 + j);
        // the only TMEM read of S: this row's 128 scores
        uint32_t v0[32], v1[32], v2[32], v3[32];
        tmem_ld_32x32b_x32(tS, v0);                                       // 32-key chunks past `valid` are never read: the TMEM ->
        if (!VB_ATTN_COND_LD || valid > 32) tmem_ld_32x32b_x32(tS + 32, v1);   // register path (64 B/clk/SM) is as scarce as the MUFU
        if (!VB_ATTN_COND_LD || valid > 64) tmem_ld_32x32b_x32(tS + 64, v2);
        if (!VB_ATTN_COND_LD || valid > 96) tmem_ld_32x32b_x32(tS + 96, v3);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(t));                           // the tensor core may overwrite S_t now
        // row max: two independent chains per 32-key chunk (eight in flight) instead of one 64-deep dependent chain
        auto chunk_max = [&](int c, const uint32_t (&v)[32]) -> float {    // full / partial / empty 32-key chunk
          const int nv = valid - c * 32;
          float a = -INFINITY, b2 = -INFINITY;
          if (nv >= 32) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              a = fmaxf(a, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
              b2 = fmaxf(b2, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
            }
          } else if (nv > 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              a = (i < nv) ? fmaxf(a, __uint_as_float(v[i])) : a;
              b2 = (i + 1 < nv) ? fmaxf(b2, __uint_as_float(v[i + 1])) : b2;
            }
          }
          return fmaxf(a, b2);
        };
        const float mx = fmaxf(fmaxf(chunk_max(0, v0), chunk_max(1, v1)), fmaxf(chunk_max(2, v2), chunk_max(3, v3)));
        if (wq == 0 && lane == 0) trace(1 + t, 30 + j);
        const float m_blk = mx * scale_log2;
        // lazy reference update: exact as long as every exponent stays <= 2^8 above the reference
        float alpha = 1.0f;
        const bool need = m_blk > m_ref + 8.0f;
        if (need) { alpha = ex2_approx(m_ref - m_blk); m_ref = m_blk; l_run *= alpha; }   // first block: alpha = 0, l = 0
        if (j > 0) {
          mbar_wait(pv_full(t), pvc & 1u);                                // PV(j-1) retired: P buffer free, O_t complete
          ++pvc;
          if (__any_sync(0xffffffffu, need)) {
            tcgen05_fence_after();
            const f32x2 a2 = splat2(alpha);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t ov[16];
              tmem_ld_32x32b_x16(tPV + g * 16, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float x0, x1;
                unpack2(mul2(pack2u(ov[2 * i], ov[2 * i + 1]), a2), x0, x1);
                ov[2 * i] = __float_as_uint(x0);
                ov[2 * i + 1] = __float_as_uint(x1);
              }
              tmem_st_32x32b_x16(tPV + g * 16, ov);
            }
            tmem_st_wait();
          }
        }
        if (wq == 0 && lane == 0) trace(1 + t, 40 + j);
        // probabilities -> bf16 -> swizzled shared memory (A operand of the PV product)
        f32x2 rsum2 = 0ull;
        const f32x2 sc2 = splat2(scale_log2), nm2 = splat2(-m_ref);
        auto emit = [&](int c, const uint32_t (&v)[32]) {                 // chunk c = keys [32c, 32c+32) of the block
          const int nv = valid - c * 32;
          if (nv <= 0) return;                                            // beyond what the PV product reads
          const uint32_t rowp = sPt + (c >> 1) * TILE_BYTES;
          if (nv >= 32) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint32_t pk[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int e = k * 8 + 2 * i;
                const f32x2 t2 = fma2(pack2u(v[e], v[e + 1]), sc2, nm2);
                if ((VB_ATTN_POLY_MASK >> k) & 1) {                           // FMA-pipe exponential
                  const f32x2 p2 = ex2_poly2(t2);
                  rsum2 = add2(rsum2, p2);
                  pk[i] = pack_bf16x2_from(p2);
                } else {                                                      // MUFU.EX2
                  float x0, x1;
                  unpack2(t2, x0, x1);
                  const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                  rsum2 = add2(rsum2, pack2(p0, p1));
                  pk[i] = pack_bf16x2(p0, p1);
                }
              }
              const uint32_t slot = static_cast<uint32_t>(((c & 1) * 4 + k) ^ (row_local & 7));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + slot * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                           "r"(pk[3]) : "memory");
            }
          } else {
            // partial chunk: the PV product reads whole 16-key steps, so only the 8-key groups below round_up(nv, 16) are
            // written, and only those below nv cost MUFU work (n = 197: 8 exponentials for the 5 keys of the last chunk, not 32)
            const int nv16 = (nv + 15) & ~15;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k * 8 >= nv16) break;
              uint32_t pk[4] = {0u, 0u, 0u, 0u};
              if (k * 8 < nv) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int e = k * 8 + 2 * i;
                  float x0, x1;
                  unpack2(fma2(pack2u(v[e], v[e + 1]), sc2, nm2), x0, x1);
                  const float p0 = (e < nv) ? ex2_approx(x0) : 0.f;
                  const float p1 = (e + 1 < nv) ? ex2_approx(x1) : 0.f;
                  rsum2 = add2(rsum2, pack2(p0, p1));
                  pk[i] = pack_bf16x2(p0, p1);
                }
              }
              const uint32_t slot = static_cast<uint32_t>(((c & 1) * 4 + k) ^ (row_local & 7));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + slot * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                           "r"(pk[3]) : "memory");
            }
          }
        };
        if (j == 0) {                                                     // previous item's output slab (same bytes) must be drained
          if (lane == 0) bulk_wait_group_read<0>();
          __syncwarp();
        }
        // Ping-pong of the two tiles' exponential phases (VB_ATTN_PINGPONG): a named-barrier token makes the groups take turns, so
        // that one tile's MUFU / FP32 phase runs against the other's TMEM-load + row-max phase instead of both contending for the
        // same unit in lock step (the shared K/V ring re-aligned the tiles at every block; a start-up offset alone measured nothing).
        if (pingpong) named_bar_sync(2 + t, 256);
        emit(0, v0);
        emit(1, v1);
        emit(2, v2);
        emit(3, v3);
        if (pingpong) named_bar_arrive(2 + (t ^ 1), 256);
        float rs0, rs1;
        unpack2(rsum2, rs0, rs1);
        l_run += rs0 + rs1;
        tcgen05_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(t));
        if (wq == 0 && lane == 0) trace(1 + t, 50 + j);
      }
      mbar_wait(pv_full(t), pvc & 1u);
      if (wq == 0 && lane == 0) trace(1 + t, 60);
      ++pvc;
      tcgen05_fence_after();
      const f32x2 il2 = splat2(1.0f / l_run);
      // normalised row -> this warp's 32-row slab of the (now idle) P buffer -> TMA store in 'b n (h d)' order;
      // rows past nq are clipped by the tensor map
      {
        uint32_t oa[32], ob[32];                                                // both halves in flight, one wait
        tmem_ld_32x32b_x32(tPV, oa);
        tmem_ld_32x32b_x32(tPV + 32, ob);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int d = k * 4 + i;                                          // pair index within this half
              pk[i] = pack_bf16x2_from(mul2(c == 0 ? pack2u(oa[2 * d], oa[2 * d + 1]) : pack2u(ob[2 * d], ob[2 * d + 1]), il2));
            }
            const uint32_t slot = static_cast<uint32_t>((c * 4 + k) ^ (row_local & 7));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPt + slot * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                         "r"(pk[3]) : "memory");
          }
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmap_o, sP + t * P_BYTES + wq * 4096, h * DH, q0 + wq * 32, b);
        bulk_commit_group();
      }
      __syncwarp();
      if (wq == 0 && lane == 0) trace(1 + t, 70);
    }
    if (lane == 0) bulk_wait_group_read<0>();   // the slab must outlive the store's reads; the writes drain before the grid completes
  }
done:
  tcgen05_fence_before();
  __syncthreads();
  if (warp == ATT_MMA_WARP0) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS_ATT>(tmem_base);
  }
}

}  // namespace
long long*& attn_trace_buffer() {   // device buffer [4 roles][256 (tag, clock) pairs]; null = tracing off
  static long long* p = nullptr;
  return p;
}
namespace {
struct AttnPlan { CUtensorMap q, k, v, o; };
using AttnKey = std::tuple<const void*, int, const void*, int, const void*, int, const void*, int, int, int, int, int>;
std::map<AttnKey, AttnPlan>& plan_cache() {
  static std::map<AttnKey, AttnPlan> c;
  return c;
}

}  // namespace

template <>
bool attention_fast<__nv_bfloat16>(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                                   __nv_bfloat16* out, int ldo, int B, int nq, int nk, int heads, int dh, int variant,
                                   const float* mix_a, const float* mix_b, const float* ln_g, const float* ln_b, cudaStream_t s,
                                   float scale) {
  if (nq == 1 && attention_cls(q, ldq, k, ldk, v, ldv, out, ldo, B, nk, heads, dh, variant, mix_a, mix_b, ln_g, ln_b, s, scale)) return true;
  if (variant != 0) return attention_mix(q, ldq, k, ldk, v, ldv, out, ldo, B, nq, nk, heads, dh, variant, mix_a, mix_b, ln_g, ln_b, s, scale);
  if (dh != DH || nq < 2) return false;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) % 16) return false;
  static unsigned long long seen[4] = {0, 0, 0, 0};
  // Form of the kernel.  Measured on one box (profiles/r02_ab_attn_keysplit.txt): n = 197 (two key blocks) 113.5 us with 8 softmax
  // warps + ping-pong against 113.0 us key-split; n = 577 (five blocks) 125.5 against 117.2 us.  Neither four warps per
  // sub-partition nor de-phasing moves the n = 197 case: there the sub-partition alternates between TMEM -> register loads
  // (S in fp32: 64 KB per sub-partition and item at 16 B/clk) and the exp / pack instruction stream, and both are needed in full.
  // Key-split is used where it wins: more than two key blocks.  VB_ATTN_KEYSPLIT=0/1 forces a form.
  static const char* ks_env = getenv("VB_ATTN_KEYSPLIT");
  const bool key_split = ks_env != nullptr ? (ks_env[0] != '0') : (VB_ATTN_KEYSPLIT || nk > 2 * BKV);
  if (first_use_on_this_device(seen)) {
    VB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    VB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
  }
  AttnKey key{q, ldq, k, ldk, v, ldv, out, ldo, B, nq, nk, heads};
  AttnPlan plan;                                   // copied out under the lock: another thread may clear the cache meanwhile
  {
  std::lock_guard<std::mutex> lock(global_cache_mutex());
  auto& cache = plan_cache();
  auto it = cache.find(key);
  if (it == cache.end()) {
    if (cache.size() > 4096) cache.clear();
    AttnPlan p;
    const uint64_t inner = static_cast<uint64_t>(heads) * DH;
    p.q = make_tmap_3d(q, inner, nq, B, static_cast<uint64_t>(ldq) * 2, static_cast<uint64_t>(nq) * ldq * 2, DH, BQ, 1);
    p.k = make_tmap_3d(k, inner, nk, B, static_cast<uint64_t>(ldk) * 2, static_cast<uint64_t>(nk) * ldk * 2, DH, BKV, 1);
    p.v = make_tmap_3d(v, inner, nk, B, static_cast<uint64_t>(ldv) * 2, static_cast<uint64_t>(nk) * ldv * 2, DH, BKV, 1);
    p.o = make_tmap_3d(out, inner, nq, B, static_cast<uint64_t>(ldo) * 2, static_cast<uint64_t>(nq) * ldo * 2, DH, 32, 1);
    it = cache.emplace(key, p).first;
  }
  plan = it->second;
  }
  const int pairs = (nq + 2 * BQ - 1) / (2 * BQ);
  const int num_items = B * heads * pairs;
  const int grid = num_items < sm_count() ? num_items : sm_count();
  const float scale_log2 = (scale > 0.f ? scale : 1.0f / sqrtf(static_cast<float>(dh))) * 1.4426950408889634f;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(key_split ? AttCfg<true>::THREADS : AttCfg<false>::THREADS);
  cfg.dynamicSmemBytes = ATT_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (key_split) VB_CUDA(cudaLaunchKernelEx(&cfg, attn_fwd_kernel<true>, plan.q, plan.k, plan.v, plan.o, heads, nq, nk, num_items, scale_log2,
                                            attn_trace_buffer()));
  else VB_CUDA(cudaLaunchKernelEx(&cfg, attn_fwd_kernel<false>, plan.q, plan.k, plan.v, plan.o, heads, nq, nk, num_items, scale_log2,
                                  attn_trace_buffer()));
  VB_CUDA(cudaGetLastError());
  count_launch();
  return true;
}

}  // namespace vb
