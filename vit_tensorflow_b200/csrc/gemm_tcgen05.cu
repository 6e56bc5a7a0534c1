// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   out[M,N] (bf16) = epi( A[M,K] (bf16, K-major) x Wt[N,K]^T (bf16, K-major) ),  fp32 accumulation in TMEM
//   epi(v) = (+bias[n]) -> exact-erf GELU -> (*scale[n]) -> (+res[m,n])
//
// Replaces, on the reference's hot path, every nn.Dense: patch embedding (vit.py:143), to_qkv (vit.py:59,72),
// to_out + residual (vit.py:62-69,101), MLP fc1+GELU / fc2 + residual (vit.py:38-44,102), CaiT to_q/to_kv
// (cait.py:94-95) with LayerScale folded in (cait.py:48).
//
// Structure (one CTA per SM, 448 threads, static round-robin tile scheduler).  CG = 2 pairs two SMs
// (cta_group::2, a 2-CTA cluster) on one 256 x BN tile: each CTA stages its own 128 A rows and HALF of the B rows,
// the leader CTA issues M = 256 MMAs that read B from both CTAs' shared memory.  Per SM this cuts the TMA-write +
// UMMA-read shared-memory traffic from 192 to 128 B/clk -- the 1-CTA form measured 65 % tensor-pipe activity
// (profiles/r01_ncu_gemm_qkv_1cta.txt), exactly the 128/192 shared-memory-bandwidth bound.
//   warp 12     TMA producer: A tile 128x64 + B tile (BN/CG)x64 per k-block into a STAGES-deep 128B-swizzled ring
//   warp 13     MMA issuer (leader CTA): one thread issues tcgen05.mma (128*CG x BN x 16) into a double-buffered
//               TMEM accumulator (each CTA holds its 128 rows)
//   warps 0-11  epilogue: tcgen05.ld -> bias/GELU/LayerScale/residual in registers -> bf16 -> per-warp swizzled smem
//               slab -> per-warp TMA store.  Three warps per TMEM lane quarter (two for 128-wide tiles) take the
//               64-column chunks of the tile stream round-robin, so a warp owns 4/3 chunks per 256-wide tile and
//               the chunk time (about 3 K cycles with GELU) stays under the 5.3 K-cycle MMA time of a K = 768 tile.
#include "common.h"
#include "kernels.cuh"
#include "ptx.cuh"

#include <cstdlib>

namespace vb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;             // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
#ifndef VB_GEMM_DIRECT_STORE
#define VB_GEMM_DIRECT_STORE 0            // 0: registers -> smem slab -> TMA store; 1 (experiment): registers -> global (STG.256)
#endif
#ifndef VB_GEMM_EW_NORES
#define VB_GEMM_EW_NORES 3               // epilogue warps per TMEM lane quarter, 256-wide tiles WITHOUT a residual operand (2 or 3)
#endif
#ifndef VB_GEMM_BN192
#define VB_GEMM_BN192 1                   // 192-column tiles for N % 192 == 0 && N % 256 != 0: 1 = N >= 1024, 2 = every such N
#endif
constexpr int STAGING_BYTES = 4096;       // per-warp slab: 32 rows x 64 bf16 columns (128-byte swizzled rows)
constexpr int CONST_BYTES = 2048;         // folded LayerNorm: (mean, rstd) of the tile's 128 rows, double-buffered

// RES decides the epilogue organisation.  With a residual operand (to_out, fc2, patch embedding) a warp needs two 4 KB slabs
// (residual-in of the next chunk / output of this one) and 2 warps per TMEM lane quarter keep the 5-stage operand ring.  Without
// one (to_qkv, fc1 + GELU -- the LayerNorm-folded, epilogue-heavy GEMMs whose MMA warp measured 4-6 K cycles of tmem_empty
// wait every other tile with 8 epilogue warps) one slab per warp is enough, so 12 warps (3 per sub-partition) fit beside the
// same ring; their chunk body runs on 32-column halves to stay inside the 136 registers a 480-thread CTA allows.
// OF32: fp32 output (the T2T soft-split attention scores, softmaxed by a separate kernel): a 64-column chunk is two 128-byte-row
// slabs of 32 columns, each with its own TMA store through an fp32 tensor map.
template <int BN, int CG, bool RES, bool OF32 = false>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;               // each CTA of a pair stages half of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int CPT = BN / 64;                               // 64-column chunks per tile
  static constexpr int EW = (CPT >= 3 && !RES) ? VB_GEMM_EW_NORES : 2;  // epilogue warps per TMEM lane quarter (<= CPT)
  static constexpr int EPI_WARPS = 4 * EW;                          // TMEM lane quarter = warp % 4, chunk lane = warp / 4
  static constexpr int PRODUCER_WARP = EPI_WARPS;                   // the issue arbiter favours high warp ids: keep the latency-critical
  static constexpr int MMA_WARP = EPI_WARPS + 1;                    // single-instruction-stream roles above the math-heavy epilogue warps
  static constexpr int STATS_WARP = EPI_WARPS + 2;                  // folded LayerNorm: (sum, sumsq) partials -> (mean, rstd), a tile ahead
  static constexpr int NUM_THREADS = (EPI_WARPS + 3) * 32;
  // 4 KB slabs per epilogue warp: residual-in (double-buffered, prefetched one chunk ahead) and, with the TMA-store
  // epilogue, the output staging (shared with the residual slab)
  static constexpr int SLABS = (RES || OF32) ? 2 : (VB_GEMM_DIRECT_STORE ? 0 : 1);
  static constexpr bool HALVES = EW >= 3;                           // chunk body on 32-column halves (register diet)
  static constexpr int NUM_STAGING = SLABS * EPI_WARPS;
  static constexpr int STAGES_FIT = (227 * 1024 - NUM_STAGING * STAGING_BYTES - CONST_BYTES - 512 - 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;  // double-buffered accumulator, allocation rounded to a power of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NUM_STAGING * STAGING_BYTES + CONST_BYTES + 512 /*barriers*/ + 1024 /*align*/;
  static_assert(8 * (2 * STAGES + 9 + 2 * EPI_WARPS) + 8 <= 512, "barrier area overflow");
};

// Exact-erf GELU (vit.py:34), written as  gelu(x) = x/2 - |x| * (E(|x|) - 1/2),  E(a) = erfc(a/sqrt2) / 2  (for x > 0
// this is x - x E, for x < 0 it is x E), with E evaluated as 2^q(a): q is a degree-5 polynomial (weighted minimax fit
// of log2(erfc(a/sqrt2)/2) on a in [0, 6], constant term exactly -1; tools/gelu_error.py refits and checks it).  Its
// leading coefficient is negative and q is monotone beyond the fit range, so no clamp is needed: for |x| > 6 the tail
// term |x| 2^q is below 6 * 2^-29 and shrinks.  Max abs error of the GELU 5.7e-7 = 0.003 bf16 ulp of the result for
// every finite x; gelu(+inf) = +inf and gelu(-inf) = NaN as in the reference's x * Phi(x).
// Cost on fp32 PAIRS: 8 FFMA2/FADD2/FMUL2 + 2 LOP3 + 2 MUFU.EX2 per two elements, all on register pairs in place.
// The epilogue is issue-bound (ubench: FFMA2 1.7, MUFU 8, FMNMX/F2FP 2 cycles per warp instruction per sub-partition;
// ncu: 770 warp instructions per 64-column chunk with the Abramowitz-Stegun rcp + ex2 form), so the form with the
// fewest instructions, not the fewest flops, wins.
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x) {
  const f32x2 na = x | 0x8000000080000000ull;                                       // -|x|
  f32x2 q = fma2(splat2(4.881368368e-04f), na, splat2(7.198925130e-03f));          // odd coefficients negated: q(-na)
  q = fma2(q, na, splat2(5.214704946e-02f));
  q = fma2(q, na, splat2(-4.595955014e-01f));
  q = fma2(q, na, splat2(1.151000619e+00f));
  q = fma2(q, na, splat2(-1.0f));
  float q0, q1;
  unpack2(q, q0, q1);
  const f32x2 e = pack2(ex2_approx(q0), ex2_approx(q1));                           // erfc(|x|/sqrt2) / 2
  return fma2(na, add2(e, splat2(-0.5f)), mul2(x, splat2(0.5f)));
}

// EPI: 0 = no per-column addend, 1 = + bias[n], 2 = folded LayerNorm (c1 = ln_c1, c2 = bias)
template <int BN, bool GELU, bool RES, int CG, int EPI, bool OF32>
__global__ void __launch_bounds__((Cfg<BN, CG, RES, OF32>::NUM_THREADS), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r, int M, int N, int K,
                 __nv_bfloat16* __restrict__ out, int ldc, const float* __restrict__ bias, const float* __restrict__ scale,
                 const __nv_bfloat16* res, int ldr, const float* __restrict__ ln_c1, const float2* ln_stats, int ln_parts,
                 float ln_inv_d, float2* __restrict__ stats_out, int stats_parts, long long* __restrict__ dbg) {
  using C = Cfg<BN, CG, RES, OF32>;
  static_assert(!OF32 || (!GELU && !RES && !VB_GEMM_DIRECT_STORE), "fp32 output: plain / bias epilogue only");
  constexpr int PRODUCER_WARP = C::PRODUCER_WARP, MMA_WARP = C::MMA_WARP, STATS_WARP = C::STATS_WARP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_stage0 = smem_base;
  const uint32_t smem_staging = smem_base + C::STAGES * C::STAGE_BYTES;
  const uint32_t smem_consts = smem_staging + C::NUM_STAGING * STAGING_BYTES;
  const uint32_t bar_base = smem_consts + CONST_BYTES;
  // barrier layout (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * C::STAGES + 4);
  auto res_bar = [&](int w, int p) { return bar_base + 8u * (2 * C::STAGES + 5 + w * 2 + p); };   // per epilogue warp x slab
  auto lnfull_bar = [&](int b) { return bar_base + 8u * (2 * C::STAGES + 5 + 2 * C::EPI_WARPS + b); };
  auto lnempty_bar = [&](int b) { return bar_base + 8u * (2 * C::STAGES + 7 + 2 * C::EPI_WARPS + b); };
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to aligned base

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // trace role 3 = kernel phases seen by thread 0 of CTA 0: 90 entry, 91 set-up done, 92 predecessor grid complete, 99 exit
  auto phase_mark = [&](int slot, int tag) {
    if (dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { dbg[3 * 512 + 2 * slot] = tag; dbg[3 * 512 + 2 * slot + 1] = clock64(); }
  };
  phase_mark(0, 90);
  constexpr int TM = BM * CG;                                      // rows per tile (per CTA pair when CG == 2)
  const int tiles_m = (M + TM - 1) / TM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  const int cta_rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = cta_rank == 0;
  const int tile0 = blockIdx.x / CG;                               // persistent schedule over clusters
  const int tile_step = gridDim.x / CG;

  if (warp == PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    if (RES) tma_prefetch_desc(&tmap_r);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), CG * C::EPI_WARPS);         // the leader counts both CTAs' epilogue warps
    }
    for (int w = 0; w < C::EPI_WARPS; ++w) { mbar_init(res_bar(w, 0), 1); mbar_init(res_bar(w, 1), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(lnfull_bar(b), 1); mbar_init(lnempty_bar(b), C::EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    if (CG == 2) tmem_alloc_cg2<C::TMEM_COLS>(tmem_ptr_smem); else tmem_alloc<C::TMEM_COLS>(tmem_ptr_smem);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                                 // peer barriers are initialised before any remote signal
  tcgen05_fence_after();
  // Everything above overlapped the previous kernel's tail (PDL); from here on we touch its outputs.
  phase_mark(1, 91);
  pdl_wait();
  pdl_launch_dependents();
  phase_mark(2, 92);
  // optional timeline trace (VB_GEMM_TRACE=path through vb_op_linear): CTA 0 records (tag, clock) pairs per role
  int dbg_n = 0;
  auto trace = [&](int role, int tag) {
    if (dbg != nullptr && blockIdx.x == 0 && dbg_n < 250) {
      dbg[role * 512 + 2 * dbg_n] = tag;
      dbg[role * 512 + 2 * dbg_n + 1] = clock64();
      ++dbg_n;
    }
  };
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  if (warp == PRODUCER_WARP) {
    // ===================================================================== TMA producer (warp-uniform loop, one elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = tile0; t < num_tiles; t += tile_step) {
        const int m0 = (t / tiles_n) * TM + cta_rank * BM;          // this CTA's 128 A rows
        const int n0 = (t % tiles_n) * BN + cta_rank * (BN / CG);   // this CTA's share of the B rows
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_stage0 + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          if (elect_one()) {
            if (CG == 2) {
              // both CTAs' bytes are counted on the leader's barrier, which expects the pair's total
              const uint32_t lbar = full_bar(stage) & kPeerBitMask;
              if (leader) mbar_arrive_expect_tx(full_bar(stage), CG * C::STAGE_BYTES);
              tma_load_2d_cg2(sa, &tmap_a, lbar, kb * BK, m0);
              tma_load_2d_cg2(sb, &tmap_b, lbar, kb * BK, n0);
            } else {
              mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
              tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BK, m0);
              tma_load_2d(sb, &tmap_b, full_bar(stage), kb * BK, n0);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================================================================== MMA issuer (leader CTA; warp-uniform loop)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(TM, BN, 0, 0);
      const uint64_t da0 = make_smem_desc(smem_stage0, 16, 1024, 2);
      const uint64_t db0 = make_smem_desc(smem_stage0 + C::A_BYTES, 16, 1024, 2);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = tile0; t < num_tiles; t += tile_step) {
        if (lane == 0) trace(0, 1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        if (lane == 0) trace(0, 2);
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint64_t da = da0 + static_cast<uint32_t>(stage * (C::STAGE_BYTES >> 4));
          const uint64_t db = db0 + static_cast<uint32_t>(stage * (C::STAGE_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the 16-byte address field
              if (CG == 2) umma_f16_ss_cg2(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_f16_ss(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // frees the smem stage (in both CTAs) when these MMAs retire
            if (CG == 2) umma_commit_cg2(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
            // accumulator complete -> epilogue warps of both CTAs
            if (kb == num_kb - 1) { if (CG == 2) umma_commit_cg2(tfull_bar(acc), 3); else umma_commit(tfull_bar(acc)); }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (lane == 0) trace(0, 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == STATS_WARP) {
    // ===================================================================== LayerNorm statistics (EPI == 2 only)
    // The rows' statistics arrive as `ln_parts` (sum, sumsq) partials, one per 64-column chunk of the row, in the layout
    // [part][M] that the producing residual GEMM's epilogue (or row_stats_bf16) wrote; this warp reduces them in a fixed
    // order to (mean, rstd) for the 128 rows of every tile of this CTA, one or two tiles ahead of the epilogue warps, into
    // a double-buffered 1 KB table.  It replaces the former row_stats_finalize_kernel (24 launches per ViT-B/16 forward)
    // without putting global loads on the epilogue's critical path.  ln_stats may be a buffer that a LATER kernel of the
    // stream overwrites, never this one (a folded GEMM has no stats_out).
    if (EPI == 2) {
      uint32_t it = 0;
      for (int t = tile0; t < num_tiles; t += tile_step, ++it) {
        const int buf = it & 1;
        mbar_wait(lnempty_bar(buf), ((it >> 1) & 1u) ^ 1u);
        const int r0 = (t / tiles_n) * TM + cta_rank * BM;
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < ln_parts; ++i) {
          const float2* p = ln_stats + static_cast<size_t>(i) * M + r0 + lane;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (r0 + lane + 32 * j < M) { const float2 v = __ldg(p + 32 * j); s1[j] += v.x; s2[j] += v.y; }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float mu = s1[j] * ln_inv_d;
          const float2 mr = make_float2(mu, rsqrtf(fmaxf(s2[j] * ln_inv_d - mu * mu, 0.f) + 1e-3f));
          *reinterpret_cast<float2*>(smem_gen + (smem_consts - smem_base) + (buf * 128 + lane + 32 * j) * 8) = mr;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(lnfull_bar(buf));
      }
    }
  } else if (warp < C::EPI_WARPS) {
    // ===================================================================== epilogue (4 x EW warps)
    // Warp (q, j) owns accumulator rows [32q, 32q+32) (its TMEM lane quarter) and every EW-th 64-column chunk (one
    // 128-byte swizzled row) of this CTA's tile stream, starting at chunk j, with two private 4 KB staging slabs and its
    // own TMA loads / stores: no cross-warp barrier anywhere in the epilogue.  EW <= chunks per tile, so every warp
    // visits every tile and releases its accumulator buffer exactly once.
    constexpr int CPT = C::CPT, EW = C::EW;
    const int e = warp;
    const int q = e & 3;
    const int j = e >> 2;
    const int row_local = q * 32 + lane;
    const uint32_t slab0 = smem_staging + e * C::SLABS * 4096;     // slab(s) of 32 rows x 128 bytes, 1024-aligned
    const bool st256 = (ldc % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 32 == 0);   // 32-byte aligned 16-column groups
    const int my_tiles = tile0 < num_tiles ? (num_tiles - tile0 + tile_step - 1) / tile_step : 0;
    const uint32_t total_chunks = static_cast<uint32_t>(my_tiles) * CPT;
    uint32_t cnt = 0;                                              // chunks processed by this warp (slab = cnt & 1)
    uint32_t rph = 0;                                              // bit p = mbarrier phase of slab p's residual barrier
    // folded LayerNorm of the A operand: y = rstd * acc + (-rstd * mu) * c1[n] + c2[n]   (c2 arrives through `bias`);
    // (mu, rstd) of this thread's row come from the statistics warp's table
    // Residual tiles come in through TMA (32 rows x 64 columns into the slab that will also stage the output), one
    // chunk ahead of the math.  Per-thread row reads (32 distinct 128-byte lines per LDG) saturated the LSU: clock64
    // traces showed ~4 K cycles per chunk in those loads, 2.4x the tile's MMA time at K = 768.
    auto res_issue = [&](uint32_t k) {                              // k-th chunk of this warp in schedule order; lane 0 only
      const uint32_t kk = static_cast<uint32_t>(j) + k * EW;
      if (kk >= total_chunks) return;
      const int tt = tile0 + static_cast<int>(kk / CPT) * tile_step;
      const int cx = (tt % tiles_n) * BN + static_cast<int>(kk % CPT) * 64;
      const int cy = (tt / tiles_n) * TM + cta_rank * BM + q * 32;
      if (cx < N) {
        mbar_arrive_expect_tx(res_bar(e, k & 1), 4096);
        tma_load_2d(slab0 + (k & 1) * 4096, &tmap_r, res_bar(e, k & 1), cx, cy);
      }
    };
    if (RES && lane == 0) res_issue(0);
    int cur_it = -1, m0 = 0, n0 = 0, row = 0;
    f32x2 ln_rstd2 = 0ull, ln_nmr2 = 0ull;
#pragma unroll 1
    for (uint32_t kk = static_cast<uint32_t>(j); kk < total_chunks; kk += EW, ++cnt) {
      const int it = static_cast<int>(kk / CPT);
      const int cc = static_cast<int>(kk % CPT);                    // 64-column chunk of the tile
      const int acc = it & 1;
      if (it != cur_it) {                                           // first chunk of mine in this tile
        cur_it = it;
        const int t = tile0 + it * tile_step;
        m0 = (t / tiles_n) * TM + cta_rank * BM;
        n0 = (t % tiles_n) * BN;
        row = m0 + row_local;
        if (EPI == 2) {
          const int buf = it & 1;
          mbar_wait(lnfull_bar(buf), static_cast<uint32_t>(it >> 1) & 1u);
          const float2 mr = *reinterpret_cast<const float2*>(smem_gen + (smem_consts - smem_base) + (buf * 128 + row_local) * 8);
          __syncwarp();
          if (lane == 0) mbar_arrive(lnempty_bar(buf));
          ln_rstd2 = splat2(mr.y);
          ln_nmr2 = splat2(-mr.x * mr.y);
        }
        if (lane == 0 && q == 0) trace(1 + j, 10);
        mbar_wait(tfull_bar(acc), static_cast<uint32_t>(it >> 1) & 1u);
        tcgen05_fence_after();
        if (lane == 0 && q == 0) trace(1 + j, 11);
      }
      {
        const int c = cc;
        const int ncol0 = n0 + cc * 64;
        const bool col_ok = ncol0 < N;                              // N % 64 == 0: a chunk is entirely in or out
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + cc * 64;
        const uint32_t slab = slab0 + (RES ? (cnt & 1u) * 4096 : 0u);    // OF32: slab0 = columns 0-31, slab0 + 4096 = columns 32-63
        const uint32_t srow = slab + lane * 128;
        const bool last_of_tile = (kk + EW >= total_chunks) || (static_cast<int>((kk + EW) / CPT) != it);
        // All of this warp's TMEM reads of the accumulator buffer are complete once the last chunk's values sit in registers:
        // hand the buffer back to the MMA warp THEN, not after the chunk's math and store (a chunk body of ~3 K cycles that the
        // round-1 kernel kept on the tensor core's critical path whenever the epilogue paced the tile).
        auto release_acc = [&]() {
          if (last_of_tile) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(tempty_bar(acc) & kPeerBitMask);   // the leader's MMA warp owns this barrier
              else mbar_arrive(tempty_bar(acc));
            }
          }
        };
        f32x2 st1 = 0ull, st2 = 0ull;                               // (sum, sum of squares) of the stored bf16 outputs
        // 16 accumulator columns [ncol0 + 16 g, +16) of this thread's row: epilogue math, bf16, into the slab
        auto process = [&](int g, const uint32_t (&vv)[16]) {
          const int ncol = ncol0 + g * 16;
          f32x2 f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = pack2u(vv[2 * j], vv[2 * j + 1]);
          if (col_ok) {
            if (EPI == 2) {
              const ulonglong2* cp = reinterpret_cast<const ulonglong2*>(ln_c1 + ncol);
              const ulonglong2* bp = reinterpret_cast<const ulonglong2*>(bias + ncol);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const ulonglong2 c4 = __ldg(cp + k), b4 = __ldg(bp + k);
                f[2 * k + 0] = fma2(f[2 * k + 0], ln_rstd2, fma2(c4.x, ln_nmr2, b4.x));
                f[2 * k + 1] = fma2(f[2 * k + 1], ln_rstd2, fma2(c4.y, ln_nmr2, b4.y));
              }
            } else if (EPI == 1) {
              const ulonglong2* bp = reinterpret_cast<const ulonglong2*>(bias + ncol);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const ulonglong2 b4 = __ldg(bp + k);
                f[2 * k + 0] = add2(f[2 * k + 0], b4.x);
                f[2 * k + 1] = add2(f[2 * k + 1], b4.y);
              }
            }
            if (GELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = gelu_erf2(f[j]);
            }
            if (scale != nullptr) {
              const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(scale + ncol);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const ulonglong2 s4 = __ldg(sp + k);
                f[2 * k + 0] = mul2(f[2 * k + 0], s4.x);
                f[2 * k + 1] = mul2(f[2 * k + 1], s4.y);
              }
            }
            if (RES) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint32_t slot = static_cast<uint32_t>((g * 2 + k) ^ (lane & 7));
                uint32_t w0, w1, w2, w3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(srow + slot * 16));
                f[4 * k + 0] = add2(f[4 * k + 0], bf16x2_to_f32x2(w0));
                f[4 * k + 1] = add2(f[4 * k + 1], bf16x2_to_f32x2(w1));
                f[4 * k + 2] = add2(f[4 * k + 2], bf16x2_to_f32x2(w2));
                f[4 * k + 3] = add2(f[4 * k + 3], bf16x2_to_f32x2(w3));
              }
            }
          }
          if (OF32) {                                                 // 16 floats = 64 bytes of this row's 128-byte slab row
            const uint32_t srow32 = slab0 + (g >> 1) * 4096 + lane * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t slot = static_cast<uint32_t>(((g & 1) * 4 + k) ^ (lane & 7));
              asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(srow32 + slot * 16), "l"(f[2 * k]), "l"(f[2 * k + 1]) : "memory");
            }
            return;
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2_from(f[j]);
          if (stats_out != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const f32x2 ab = bf16x2_to_f32x2(pk[j]);
              st1 = add2(st1, ab);
              st2 = fma2(ab, ab, st2);
            }
          }
          if (VB_GEMM_DIRECT_STORE) {
            // EXPERIMENT (off): registers -> global, this thread's 16 columns = 32 contiguous bytes of its output row = one full
            // sector per lane and store (STG.256), to take the epilogue's smem round trip (64 KB written + 64 KB read per
            // 128 x 256 tile) off the shared-memory port the mainloop saturates.  MEASURED SLOWER (profiles/r02_ab_gemm_direct_store.txt,
            // one box): to_qkv 1066 vs 1161 TF/s, fc1 + GELU 1051 vs 1129, dim-384 GEMMs 519 vs 599 -- 32 distinct lines per store
            // instruction cost more in the LSU / L1 than the slab + one TMA store per chunk.
            if (col_ok && row < M) {
              __nv_bfloat16* orow = out + static_cast<size_t>(row) * ldc + ncol;
              if (st256) {
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                             "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
              } else {
                asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(orow), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(orow + 8), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint32_t slot = static_cast<uint32_t>((g * 2 + k) ^ (lane & 7));   // 128B swizzle: 16-byte slot ^ (row % 8)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + slot * 16), "r"(pk[4 * k]), "r"(pk[4 * k + 1]),
                           "r"(pk[4 * k + 2]), "r"(pk[4 * k + 3]) : "memory");
            }
          }
        };
        if (!C::HALVES) {
          uint32_t v[4][16];
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld_32x32b_x16(tcol + g * 16, v[g]);
          // the OTHER slab was last read by the TMA store of the previous chunk: once that has drained, prefetch the next
          // chunk's residual into it (or, without a residual, make the single slab writable again)
          if (lane == 0) {
            if (!VB_GEMM_DIRECT_STORE) bulk_wait_group_read<0>();
            else if (RES) fence_proxy_async_smem();                     // this warp's ld.shared of that slab precede the TMA write
            if (RES) res_issue(cnt + 1);
          }
          tmem_ld_wait();
          release_acc();
          if (lane == 0 && q == 0) trace(1 + j, 30 + c);
          if (RES && col_ok) {                                        // this chunk's residual has landed
            mbar_wait(res_bar(e, cnt & 1u), (rph >> (cnt & 1u)) & 1u);
            rph ^= 1u << (cnt & 1u);
          }
          __syncwarp();
          if (lane == 0 && q == 0) trace(1 + j, 40 + c);
#pragma unroll
          for (int g = 0; g < 4; ++g) process(g, v[g]);
        } else {
          // 32-column halves: half the live accumulator registers (12 epilogue warps share the register file)
          uint32_t v[2][16];
          tmem_ld_32x32b_x16(tcol, v[0]);
          tmem_ld_32x32b_x16(tcol + 16, v[1]);
          if (!VB_GEMM_DIRECT_STORE && lane == 0) bulk_wait_group_read<0>();   // the single slab: the previous chunk's store has read it
          tmem_ld_wait();
          __syncwarp();
          if (lane == 0 && q == 0) trace(1 + j, 30 + c);
          process(0, v[0]);
          process(1, v[1]);
          tmem_ld_32x32b_x16(tcol + 32, v[0]);
          tmem_ld_32x32b_x16(tcol + 48, v[1]);
          tmem_ld_wait();
          release_acc();
          if (lane == 0 && q == 0) trace(1 + j, 40 + c);
          process(2, v[0]);
          process(3, v[1]);
        }
        if (stats_out != nullptr && col_ok && row < M) {
          float a0, a1, b0, b1;
          unpack2(st1, a0, a1);
          unpack2(st2, b0, b1);
          stats_out[static_cast<size_t>(ncol0 >> 6) * M + row] = make_float2(a0 + a1, b0 + b1);   // [part][M]: coalesced over the rows
        }
        if (lane == 0 && q == 0) trace(1 + j, 50 + c);
        if (!VB_GEMM_DIRECT_STORE) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (col_ok) {
              tma_store_2d(&tmap_c, slab, ncol0, m0 + q * 32);
              if (OF32) tma_store_2d(&tmap_c, slab + 4096, ncol0 + 32, m0 + q * 32);
            }
            bulk_commit_group();
          }
        }
        if (lane == 0 && q == 0) trace(1 + j, 20 + c);
      }
    }
    if (!VB_GEMM_DIRECT_STORE && lane == 0) bulk_wait_group_read<0>();   // the slab must outlive the store's reads; the writes drain before the grid completes
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();     // no CTA exits (or frees TMEM) while its peer can still signal / read it
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc_cg2<C::TMEM_COLS>(tmem_base); else tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
  phase_mark(3, 99);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static const EncodeTiledFn fn = [] {             // C++11 magic static: initialised once, thread-safe
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    VB_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace
long long*& gemm_trace_buffer() {   // debugging aid: device trace buffer [4 roles][256 (tag, clock)]; null = off
  static long long* p = nullptr;
  return p;
}
namespace {

template <int BN, bool GELU, bool RES, int CG, int EPI, bool OF32 = false>
void launch(const GemmBf16& g, cudaStream_t stream) {
  auto kern = gemm_bf16_kernel<BN, GELU, RES, CG, EPI, OF32>;
  static unsigned long long seen[4] = {0, 0, 0, 0};
  if (first_use_on_this_device(seen)) VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, CG, RES, OF32>::SMEM_BYTES));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.grid);
  cfg.blockDim = dim3(Cfg<BN, CG, RES, OF32>::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg<BN, CG, RES, OF32>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  VB_CUDA(cudaLaunchKernelEx(&cfg, kern, g.tmap_a, g.tmap_b, g.tmap_c, g.tmap_r, g.M, g.N, g.K, g.out, g.ldc, g.bias, g.scale, g.res, g.ldr, g.ln_c1,
                             reinterpret_cast<const float2*>(g.ln_stats), g.ln_parts, g.ln_inv_d, reinterpret_cast<float2*>(g.stats_out),
                             g.stats_parts, gemm_trace_buffer()));
  count_launch();
}

template <int BN, int CG>
void launch_epi(const GemmBf16& g, cudaStream_t stream) {
  const bool res = g.res != nullptr;
  const int epi = g.ln_c1 != nullptr ? 2 : g.bias != nullptr ? 1 : 0;
  VB_CHECK(epi != 2 || (g.bias != nullptr && g.ln_stats != nullptr && g.ln_parts > 0), "folded LayerNorm needs c1, c2 and the row statistics");
  if (g.out_f32) {
    VB_CHECK(!g.gelu && !res && epi != 2 && g.scale == nullptr && g.stats_out == nullptr, "fp32-output GEMM: plain or bias epilogue only");
    if (epi == 0) return launch<BN, false, false, CG, 0, true>(g, stream);
    return launch<BN, false, false, CG, 1, true>(g, stream);
  }
#define VB_GEMM_CASE(G, R, E) if (g.gelu == G && res == R && epi == E) return launch<BN, G, R, CG, E>(g, stream)
  VB_GEMM_CASE(false, false, 0); VB_GEMM_CASE(false, false, 1); VB_GEMM_CASE(false, false, 2);
  VB_GEMM_CASE(true, false, 0);  VB_GEMM_CASE(true, false, 1);  VB_GEMM_CASE(true, false, 2);
  VB_GEMM_CASE(false, true, 0);  VB_GEMM_CASE(false, true, 1);  VB_GEMM_CASE(false, true, 2);
  VB_GEMM_CASE(true, true, 0);   VB_GEMM_CASE(true, true, 1);   VB_GEMM_CASE(true, true, 2);
#undef VB_GEMM_CASE
}

}  // namespace

int sm_count() {                                  // of the CURRENT device (one process may hold handles on several GPUs)
  static int n[64] = {0};
  int dev = 0;
  VB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(global_cache_mutex());
  int& v = n[dev & 63];
  if (v == 0) VB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
  return v;
}

CUtensorMap make_tmap_2d(const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes, uint32_t box_inner,
                         uint32_t box_outer, bool swizzle128, bool f32) {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed with CUresult " + std::to_string(static_cast<int>(r)));
  return m;
}

CUtensorMap make_tmap_3d(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                         uint32_t box0, uint32_t box1, uint32_t box2, bool swizzle128) {
  CUtensorMap m;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult " + std::to_string(static_cast<int>(r)));
  return m;
}

bool gemm_bf16_supported(int M, int N, int K, int lda, int ldw, int ldc) {
  return M > 0 && N > 0 && K > 0 && (N % 64 == 0) && (K % 8 == 0) && (lda % 8 == 0) && (ldw % 8 == 0) && (ldc % 8 == 0);
}

GemmBf16 gemm_bf16_plan(const __nv_bfloat16* A, int lda, const __nv_bfloat16* Wt, int ldw, __nv_bfloat16* out, int ldc, int M,
                        int N, int K, const float* bias, const float* scale, const __nv_bfloat16* res, int ldr, bool gelu,
                        bool out_f32, int b_rows) {
  VB_CHECK(gemm_bf16_supported(M, N, K, lda, ldw, ldc), "gemm_bf16: unsupported shape (need N%64==0, K%8==0, ld%8==0)");
  VB_CHECK(res == nullptr || ldr % 8 == 0, "gemm_bf16: residual leading dimension must be a multiple of 8");
  VB_CHECK((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Wt) | reinterpret_cast<uintptr_t>(out) |
            reinterpret_cast<uintptr_t>(res)) % 16 == 0, "gemm_bf16: operands must be 16-byte aligned");
  GemmBf16 g;
  g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.scale = scale; g.res = res; g.ldr = ldr; g.gelu = gelu;
  g.out = out; g.ldc = ldc; g.out_f32 = out_f32;
  // 256-wide tiles when they divide N; for a wide N that they do not divide, 192-wide ones when those do (CaiT to_qkv,
  // N = 1152: 6 exact tile columns instead of 5 with the last half empty, +6 % measured, profiles/r02_ab_gemm_bn192.txt);
  // 128-wide ones for a narrow N.  192-wide tiles for N = 384 measured no better than 128-wide ones (those GEMMs are bound by
  // their epilogue, not by operand ingest), so a narrow N keeps the 128-wide form.  VB_GEMM_BN192: 0 = never, 1 = N >= 1024
  // (default), 2 = every N % 192 == 0.
  static const char* bn192_env = getenv("VB_GEMM_BN192");
  const int bn192 = bn192_env != nullptr ? atoi(bn192_env) : VB_GEMM_BN192;
  if (N % 256 == 0) g.block_n = 256;
  else if (N % 192 == 0 && bn192 > 0 && (N >= 1024 || bn192 > 1)) g.block_n = 192;
  else g.block_n = N >= 1024 ? 256 : 128;
  g.cta_group = (M > BM) ? 2 : 1;      // pair two SMs on 256-row tiles unless the whole problem is one 128-row tile
  g.tmap_a = make_tmap_2d(A, K, M, static_cast<uint64_t>(lda) * 2, BK, BM);
  // b_rows: rows of Wt that exist (< N when the output is column-padded: TMA zero-fills the rest instead of reading on)
  g.tmap_b = make_tmap_2d(Wt, K, b_rows > 0 ? b_rows : N, static_cast<uint64_t>(ldw) * 2, BK, g.block_n / g.cta_group);
  // fp32 output: `out` really is a float*, ldc in floats, 32-column (128-byte) boxes
  g.tmap_c = out_f32 ? make_tmap_2d(out, N, M, static_cast<uint64_t>(ldc) * 4, 32, 32, true, true)
                     : make_tmap_2d(out, N, M, static_cast<uint64_t>(ldc) * 2, 64, 32);
  g.tmap_r = res ? make_tmap_2d(res, N, M, static_cast<uint64_t>(ldr) * 2, 64, 32) : g.tmap_c;
  const int tm = BM * g.cta_group;
  const int tiles = ((M + tm - 1) / tm) * ((N + g.block_n - 1) / g.block_n);
  const int clusters = sm_count() / g.cta_group;
  g.grid = (tiles < clusters ? tiles : clusters) * g.cta_group;
  return g;
}

void gemm_bf16_run(const GemmBf16& g, cudaStream_t stream) {
  if (g.cta_group == 2) {
    if (g.block_n == 256) launch_epi<256, 2>(g, stream); else if (g.block_n == 192) launch_epi<192, 2>(g, stream); else launch_epi<128, 2>(g, stream);
  } else {
    if (g.block_n == 256) launch_epi<256, 1>(g, stream); else if (g.block_n == 192) launch_epi<192, 1>(g, stream); else launch_epi<128, 1>(g, stream);
  }
}

}  // namespace vb
