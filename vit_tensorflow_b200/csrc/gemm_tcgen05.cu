// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   out[M,N] (bf16) = epi( A[M,K] (bf16, K-major) x Wt[N,K]^T (bf16, K-major) ),  fp32 accumulation in TMEM
//   epi(v) = (+bias[n]) -> exact-erf GELU -> (*scale[n]) -> (+res[m,n])
//
// Replaces, on the reference's hot path, every nn.Dense: patch embedding (vit.py:143), to_qkv (vit.py:59,72),
// to_out + residual (vit.py:62-69,101), MLP fc1+GELU / fc2 + residual (vit.py:38-44,102), CaiT to_q/to_kv
// (cait.py:94-95) with LayerScale folded in (cait.py:48).
//
// Structure (one CTA per SM, 384 threads, static round-robin tile scheduler):
//   warp 0      TMA producer: A tile 128x64 + B tile BNx64 per k-block into a STAGES-deep 128B-swizzled ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma (128 x BN x 16) into a double-buffered TMEM accumulator
//   warps 4-11  epilogue: tcgen05.ld -> bias/GELU/LayerScale/residual in registers -> bf16 -> swizzled smem
//               staging (2 x 16 KB) -> TMA store.  Runs concurrently with the next tile's main loop.
#include "common.h"
#include "kernels.cuh"
#include "ptx.cuh"

namespace vb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;             // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_THREADS = 256;
constexpr int STAGING_BYTES = BM * 128;   // one 128-row x 64-col bf16 chunk
constexpr int NUM_STAGING = 2;

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered accumulator (power of two: 256 or 512)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NUM_STAGING * STAGING_BYTES + 256 /*barriers*/ + 1024 /*align*/;
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int BN, bool GELU, bool RES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, int M, int N, int K,
                 const float* __restrict__ bias, const float* __restrict__ scale,
                 const __nv_bfloat16* res, int ldr) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_stage0 = smem_base;
  const uint32_t smem_staging = smem_base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = smem_staging + NUM_STAGING * STAGING_BYTES;
  // barrier layout (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * C::STAGES + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to aligned base

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), NUM_EPI_THREADS / 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * BM;
        const int n0 = (t % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_stage0 + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BK, m0);
          tma_load_2d(sb, &tmap_b, full_bar(stage), kb * BK, n0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_stage0 + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da = make_smem_desc(sa, 16, 1024, 2);
          const uint64_t db = make_smem_desc(sb, 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the 16-byte address field
            umma_f16_ss(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));             // frees the smem stage when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));                 // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================================================================== epilogue (8 warps)
    const int e = warp - EPI_WARP0;
    const int q = e & 3;          // TMEM lane quarter this warp may touch (warp_id % 4)
    const int ch = e >> 2;        // which 32-column half of each 64-column chunk
    const int row_local = q * 32 + lane;
    const bool is_store_thread = (threadIdx.x == EPI_WARP0 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_counter = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * BM;
      const int n0 = (t % tiles_n) * BN;
      const int row = m0 + row_local;
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int ncol = n0 + c * 64 + ch * 32;
        const bool col_ok = (n0 + c * 64) < N;       // N % 64 == 0: a chunk is entirely in or out
        uint4 rres[4];
        if (RES) {
          if (col_ok && row < M) {
            const uint4* rp = reinterpret_cast<const uint4*>(res + static_cast<size_t>(row) * ldr + ncol);
#pragma unroll
            for (int k = 0; k < 4; ++k) rres[k] = rp[k];
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) rres[k] = make_uint4(0, 0, 0, 0);
          }
        }
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 64 + ch * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (col_ok) {
          if (bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(bias + ncol);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 b4 = __ldg(bp + k);
              f[4 * k + 0] += b4.x; f[4 * k + 1] += b4.y; f[4 * k + 2] += b4.z; f[4 * k + 3] += b4.w;
            }
          }
          if (GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          }
          if (scale != nullptr) {
            const float4* sp = reinterpret_cast<const float4*>(scale + ncol);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 s4 = __ldg(sp + k);
              f[4 * k + 0] *= s4.x; f[4 * k + 1] *= s4.y; f[4 * k + 2] *= s4.z; f[4 * k + 3] *= s4.w;
            }
          }
          if (RES) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t w4[4] = {rres[k].x, rres[k].y, rres[k].z, rres[k].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                f[8 * k + 2 * i + 0] += bf16_lo(w4[i]);
                f[8 * k + 2 * i + 1] += bf16_hi(w4[i]);
              }
            }
          }
        }
        // staging buffer must have been fully read by the TMA store issued two chunks ago
        const uint32_t buf = chunk_counter & 1u;
        if (is_store_thread) bulk_wait_group_read<NUM_STAGING - 1>();
        named_bar_sync(1, NUM_EPI_THREADS);
        const uint32_t srow = smem_staging + buf * STAGING_BYTES + row_local * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t slot = static_cast<uint32_t>((ch * 4 + k) ^ (row_local & 7));
          const uint32_t p0 = pack_bf16x2(f[8 * k + 0], f[8 * k + 1]);
          const uint32_t p1 = pack_bf16x2(f[8 * k + 2], f[8 * k + 3]);
          const uint32_t p2 = pack_bf16x2(f[8 * k + 4], f[8 * k + 5]);
          const uint32_t p3 = pack_bf16x2(f[8 * k + 6], f[8 * k + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + slot * 16), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
        }
        fence_proxy_async_smem();
        named_bar_sync(2, NUM_EPI_THREADS);
        if (is_store_thread) {
          if (col_ok) tma_store_2d(&tmap_c, smem_staging + buf * STAGING_BYTES, n0 + c * 64, m0);
          bulk_commit_group();
        }
        ++chunk_counter;
      }
      // all TMEM reads of this accumulator buffer are complete (tcgen05.wait::ld above)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (is_store_thread) bulk_wait_group<0>();   // all output bytes are globally visible before exit
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    VB_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int BN, bool GELU, bool RES>
void launch(const GemmBf16& g, cudaStream_t stream) {
  auto kern = gemm_bf16_kernel<BN, GELU, RES>;
  static bool configured = false;
  if (!configured) {
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
    configured = true;
  }
  kern<<<g.grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, stream>>>(g.tmap_a, g.tmap_b, g.tmap_c, g.M, g.N, g.K, g.bias, g.scale,
                                                          g.res, g.ldr);
  VB_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

CUtensorMap make_tmap_2d(const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes, uint32_t box_inner,
                         uint32_t box_outer, bool swizzle128) {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
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
                        int N, int K, const float* bias, const float* scale, const __nv_bfloat16* res, int ldr, bool gelu) {
  VB_CHECK(gemm_bf16_supported(M, N, K, lda, ldw, ldc), "gemm_bf16: unsupported shape (need N%64==0, K%8==0, ld%8==0)");
  VB_CHECK(res == nullptr || ldr % 8 == 0, "gemm_bf16: residual leading dimension must be a multiple of 8");
  VB_CHECK((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Wt) | reinterpret_cast<uintptr_t>(out) |
            reinterpret_cast<uintptr_t>(res)) % 16 == 0, "gemm_bf16: operands must be 16-byte aligned");
  GemmBf16 g;
  g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.scale = scale; g.res = res; g.ldr = ldr; g.gelu = gelu;
  // 256-wide tiles unless N only fills 128-wide ones well (e.g. CaiT dim 384) or the problem is tiny.
  g.block_n = (N % 256 == 0 || N >= 1024) ? 256 : 128;
  g.tmap_a = make_tmap_2d(A, K, M, static_cast<uint64_t>(lda) * 2, BK, BM);
  g.tmap_b = make_tmap_2d(Wt, K, N, static_cast<uint64_t>(ldw) * 2, BK, g.block_n);
  g.tmap_c = make_tmap_2d(out, N, M, static_cast<uint64_t>(ldc) * 2, 64, BM);
  const int tiles = ((M + BM - 1) / BM) * ((N + g.block_n - 1) / g.block_n);
  g.grid = tiles < sm_count() ? tiles : sm_count();
  return g;
}

void gemm_bf16_run(const GemmBf16& g, cudaStream_t stream) {
  const bool res = g.res != nullptr;
  if (g.block_n == 256) {
    if (g.gelu) { if (res) launch<256, true, true>(g, stream); else launch<256, true, false>(g, stream); }
    else        { if (res) launch<256, false, true>(g, stream); else launch<256, false, false>(g, stream); }
  } else {
    if (g.gelu) { if (res) launch<128, true, true>(g, stream); else launch<128, true, false>(g, stream); }
    else        { if (res) launch<128, false, true>(g, stream); else launch<128, false, false>(g, stream); }
  }
}

}  // namespace vb
