// Fused head-mixing attention for sm_100a: DeepViT re-attention (deepvit.py:79-87) and CaiT talking heads (cait.py:121-127)
// in ONE kernel, tcgen05 for both products, no score tensor in HBM.
//
//   variant 1 (DeepViT):  P_h = softmax_j(scale q_h k_h^T);  A_g = LN_over_g( sum_h W[h,g] P_h ) * gamma_g + beta_g;  O_g = A_g V_g
//   variant 2 (CaiT):     S'_g = sum_h Wpre[h,g] (scale q_h k_h^T);  P_g = softmax_j(S'_g);  A_g = sum_g' Wpost[g',g] P_g';  O_g = A_g V_g
//
// Both couple ALL heads of one (image, query row, key), so the per-head FlashAttention tiling does not apply: the mix needs
// the scores of every head for a key before any head's probabilities exist, and (DeepViT) the LayerNorm over heads is not
// linear, so it cannot be pushed through the PV product either.  What bounds the op is the CUDA-core work between the two
// products (H fused multiply-adds per score for each mix + the exponentials: 26-33 instructions per score), not the tensor
// core (the same 4 B h n^2 dh flops as plain attention) and not HBM (q, k, v, out once).  Structure, per work item =
// (image b, tile of 64 query rows), persistent CTAs, 10 warps:
//
//   warp 8   TMA producer.  Phase A: Q_h tiles (64 rows x 64 columns per head, 128B swizzle) once, then the key blocks of 16 keys
//            (all heads) through a ring, twice (two softmax passes).  Phase B: per head, the mixed attention weights A_g back
//            from this CTA's scratch slot (one contiguous bulk copy: the layout below IS the canonical no-swizzle K-major UMMA
//            operand) and V_g (128-key boxes, MN-major B operand exactly as TMA delivers it).
//   warp 9   MMA issuer.  Phase A: S_h = Q_h K_h^T as UMMA M = 64, N = 16 per (head, key block) into tensor memory.  An M = 64
//            accumulator occupies lanes (r % 16) + 32 (r / 16): two of them interleave in the same columns (lane offsets 0 and
//            16), so key blocks 2t and 2t+1 share one set of H x 16 columns and every mixer thread owns one (row, block) pair.
//            Phase B: O_g = A_g V_g, M = 64, N = dh, accumulating over the key steps; with M = 64 all H <= 16 heads' outputs
//            (H x 64 columns over the two lane halves) fit the 512 columns at once.
//   warps 0-7   mixers, two groups of four (group G takes key super-blocks G, G+2, ...; its own S columns).  Thread = (query
//            row, key block of the pair): reads 4 keys x H heads from tensor memory, mixes in registers against weights that
//            are kernel parameters (constant-bank operands of scalar FFMA), LayerNorm over heads in registers.
//            Pass 1: softmax statistics (per head for DeepViT, per mixed head for CaiT), merged across the four threads
//            of a row through a shuffle and 8 KB of shared memory.  Pass 2: scores again (the tensor core recomputes them:
//            cheaper than parking H x n x 64 fp32 anywhere), probabilities, mix (+ LayerNorm), bf16 (hi, and a lo plane for
//            DeepViT whose LayerNorm output reaches +-30 with mixed signs), 16-byte stores into the scratch slot
//            [head][8-key group][64 rows][8 keys] -- 512 contiguous bytes per warp and store.
//            Phase B epilogue: O from tensor memory -> bf16 -> 'b n (h d)' rows of the output.
//
// The scratch slot (H x round_up(nk, 16) x 128 bytes per plane, one per CTA: 426 KB for DeepViT-24 hi, 63 MB over 148 CTAs) is
// written and read back by the same CTA within microseconds and overwritten by its next item: it lives in the 126 MB L2, not
// in HBM -- the "score tensor" of the reference (2 x B h n^2 fp32 through memory) never exists.
// Algorithmic bytes per launch: 2 B (nq + 2 nk) h dh + 2 B nq h dh (q, k, v in; out).
#include <map>
#include <utility>
#include "attention.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

#include <cstdlib>

namespace vb {
namespace {

constexpr int MX_THREADS = 320;            // the register file is allocated per 4 warps: 168 registers per thread
constexpr int MX_PRODUCER = 8, MX_MMA = 9;
constexpr int MX_ROWS = 64;                 // query rows per item
constexpr int MX_KSLOTS_BYTES = 64 * 1024;  // key ring at H = 16 (Q takes 128 KB); H = 8 gets 128 KB
constexpr int MX_MAXNK = 256;               // keys per image the phase-B stage is sized for
constexpr int MX_QT = 64 * 128;             // Q tile of one head: 64 rows x 128 bytes
// Keys per block = N of the S products.  A tcgen05.mma costs the issuer ~100 cycles however small it is (M = 64 reads its A
// operand from shared memory per instruction): at N = 16 the 48 (H = 8) / 128 (H = 16) products of a super-block paced the whole
// kernel (ncu: mixers waiting on s_full 15-19 %).  H = 8 has the tensor-memory columns for N = 32 (2 x 8 x 32 = 512).
template <int H> struct MxKB { static constexpr int value = H <= 8 ? 32 : 16; };
constexpr int MX_VBOX = 128 * 128;          // V box: 128 keys x 128 bytes

template <int H>
struct MixW { float wa[H * H], wb[H * H], gamma[H], beta[H]; };

template <int H, bool SPLIT>
struct MxCfg {
  // Key ring: whole 16-key blocks (all heads, H x 2 KB), as deep as shared memory allows beside Q: 8 blocks at H = 8, 2 at
  // H = 16 (Q alone is 128 KB there).  With 2 blocks at H = 8 the mixers waited for S 37 % of the time (ncu source page of the
  // first build, profiles/r02_ncu_mix.md: one block per ~1 K cycles of TMA round trip against ~0.6 K of mixing); a ring of
  // single (block, head) tiles was worse still -- 16 / 32 barrier round trips per block on the issuer's critical path
  // (mix_cait 210 -> 298 us, profiles/r02_ab_mix.txt).
  static constexpr int KB = MxKB<H>::value;
  static constexpr int KT = KB * 128;                                                       // key block of one head: KB keys x 128 bytes
  static constexpr int KSLOTS = (H <= 8 ? 2 : 1) * MX_KSLOTS_BYTES / (H * KT);
  static constexpr int A_BYTES = H * MX_QT + KSLOTS * H * KT;                               // phase A: Q + key-block ring
  static constexpr int PB_STAGE = (SPLIT ? 2 : 1) * (MX_MAXNK / 8) * 1024 + 2 * MX_VBOX;    // phase B stage: A_g plane(s) + V_g
  static constexpr int B_BYTES = 2 * PB_STAGE;
  static constexpr int DATA = A_BYTES > B_BYTES ? A_BYTES : B_BYTES;
  static constexpr int STAT = 2 * H * 2 * MX_ROWS * 4;                                      // final + exchange tables
  static constexpr int BARS = 160 + 16 * KSLOTS;
  static constexpr int SMEM = DATA + STAT + BARS + 1024;
  static_assert(SMEM <= 227 * 1024, "attn_mix: shared memory budget");
};

__device__ __forceinline__ void tmem_ld_32x32b_x4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_global_v2(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

template <int H, int VARIANT, bool SPLIT>
__global__ void __launch_bounds__(MX_THREADS, 1)
attn_mix_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ MixW<H> W, __nv_bfloat16* __restrict__ out, int ldo,
                uint8_t* __restrict__ scratch, long long slot_bytes, int B, int nq, int nk, int dh, int num_items, float scale_log2) {
  // Item order: image-major -- the tiles of one image are adjacent items, i.e. run on neighbouring CTAs at the same time, so
  // the image's K and V are read from HBM once and served to the other tiles (and to both softmax passes) by the L2; the tile
  // index is rotated by the image index so that a CTA's static stride over the items does not always hit the same (e.g. the
  // short last) tile.  (Tile-major order measured 3.6x the algorithmic DRAM traffic at DeepViT-24: ncu, profiles/r02_ncu_mix.md.)
  const int tiles = (nq + MX_ROWS - 1) / MX_ROWS;
  using C = MxCfg<H, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base;
  const uint32_t sK = sQ + H * MX_QT;
  const uint32_t sPB = base;                                     // phase B stages overlay the phase A buffers
  float* stat = reinterpret_cast<float*>(gen + C::DATA);         // [H][2][64]: (m, 1/l) of the softmax rows
  float* part = stat + H * 2 * MX_ROWS;                          // exchange between the two mixer groups
  const uint32_t bars = base + C::DATA + C::STAT;
  const uint32_t q_full = bars, pa_done = bars + 8, o_full = bars + 16, o_empty = bars + 24;
  auto s_full = [&](int g) { return bars + 32u + 8u * g; };
  auto s_empty = [&](int g) { return bars + 48u + 8u * g; };
  auto pb_full = [&](int s) { return bars + 64u + 8u * s; };
  auto pb_empty = [&](int s) { return bars + 80u + 8u * s; };
  const uint32_t tmem_slot = bars + 96u;
  auto k_full = [&](uint32_t s) { return bars + 160u + 8u * s; };
  auto k_empty = [&](uint32_t s) { return bars + 160u + 8u * C::KSLOTS + 8u * s; };
  constexpr uint32_t KSLOTS = C::KSLOTS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KB = C::KB, MX_KT = C::KT, QPB = KB / 4;         // keys per block, bytes per (block, head) tile, 4-key quarters per block
  const int nblk = (nk + KB - 1) / KB;                           // key blocks
  const int nsb = (nblk + 1) >> 1;                               // super-blocks (pairs of key blocks sharing S columns)
  const int nk16 = (nk + 15) >> 4;                               // 16-key steps of the PV products
  const int nkg = 2 * nk16;                                      // 8-key groups in the scratch layout
  const int ksteps = dh >> 4;
  const uint32_t plane_bytes = static_cast<uint32_t>(H) * nkg * 1024u;
  uint8_t* slot = scratch + static_cast<long long>(blockIdx.x) * slot_bytes;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1); mbar_init(pa_done, 8); mbar_init(o_full, 1); mbar_init(o_empty, 8);
    for (uint32_t s = 0; s < KSLOTS; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(s_full(g), 1); mbar_init(s_empty(g), 8); mbar_init(pb_full(g), 1); mbar_init(pb_empty(g), 1); }
    fence_mbar_init();
  }
  if (warp == MX_MMA) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();
  pdl_launch_dependents();

  if (warp == MX_PRODUCER) {
    // ===================================================================== TMA producer
    if (lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); }
    uint32_t kcnt = 0, pcnt = 0, n = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      const int b = it / tiles, q0 = ((it + b) % tiles) * MX_ROWS;   // see item order below
      if (n > 0) mbar_wait(o_full, (n - 1) & 1u);                  // every PV product of the previous item has read its stage
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, H * MX_QT);
        for (int h = 0; h < H; ++h) tma_load_3d(sQ + h * MX_QT, &tmap_q, q_full, h * dh, q0, b);
      }
      __syncwarp();
      for (int pass = 0; pass < 2; ++pass) {
        for (int blk = 0; blk < nblk; ++blk, ++kcnt) {
          const uint32_t st = kcnt % KSLOTS;
          mbar_wait(k_empty(st), ((kcnt / KSLOTS) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(k_full(st), H * MX_KT);
            for (int h = 0; h < H; ++h) tma_load_3d(sK + (st * H + h) * MX_KT, &tmap_k, k_full(st), h * dh, blk * KB, b);
          }
          __syncwarp();
        }
      }
      // ---- phase B: A_g (scratch) + V_g per head
      mbar_wait(pa_done, n & 1u);                                  // the mixers' scratch writes are visible to the async proxy
      const int vboxes = nk > 128 ? 2 : 1;
      const uint32_t stage_bytes = (SPLIT ? 2u : 1u) * nkg * 1024u + vboxes * MX_VBOX;
      for (int g = 0; g < H; ++g, ++pcnt) {
        const int st = pcnt & 1;
        mbar_wait(pb_empty(st), ((pcnt >> 1) & 1u) ^ 1u);
        if (elect_one()) {
          const uint32_t sA = sPB + st * C::PB_STAGE;
          const uint32_t sV = sA + (SPLIT ? 2 : 1) * (MX_MAXNK / 8) * 1024;
          mbar_arrive_expect_tx(pb_full(st), stage_bytes);
          bulk_load_1d(sA, slot + static_cast<size_t>(g) * nkg * 1024, nkg * 1024u, pb_full(st));
          if (SPLIT) bulk_load_1d(sA + (MX_MAXNK / 8) * 1024, slot + plane_bytes + static_cast<size_t>(g) * nkg * 1024, nkg * 1024u, pb_full(st));
          for (int vb = 0; vb < vboxes; ++vb) tma_load_3d(sV + vb * MX_VBOX, &tmap_v, pb_full(st), g * dh, vb * 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == MX_MMA) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc_s = make_idesc_bf16(MX_ROWS, KB, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(MX_ROWS, dh, 0, 1);        // B (= V) is MN-major
    uint32_t kcnt = 0, pcnt = 0, n = 0, sbc = 0;                   // sbc: super-blocks issued so far (S buffer = sbc & 1)
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      if (n > 0) mbar_wait(o_empty, (n - 1) & 1u);                 // the previous item's outputs have left tensor memory
      mbar_wait(q_full, n & 1u);
      tcgen05_fence_after();
      for (int pass = 0; pass < 2; ++pass) {
        for (int blk = 0; blk < nblk; ++blk, ++kcnt) {
          // Both mixer groups work on the SAME super-block (group 0: key quarters 0, 1 of each block; group 1: quarters 2, 3),
          // so the two S column sets double-buffer: super-block sb + 1 is multiplied while sb is being mixed.
          const int G = static_cast<int>(sbc & 1u), p = blk & 1;
          const uint32_t st = kcnt % KSLOTS;
          mbar_wait(k_full(st), (kcnt / KSLOTS) & 1u);
          if (p == 0 && sbc >= 2) mbar_wait(s_empty(G), ((sbc >> 1) - 1) & 1u);   // all eight mixer warps have read buffer G
          tcgen05_fence_after();
          const uint32_t d0 = tmem_base + (static_cast<uint32_t>(p * 16) << 16) + G * (H * KB);
          const bool last_of_sb = (p == 1 || blk == nblk - 1);
          const uint64_t dq0 = make_smem_desc(sQ, 16, 1024, 2);
          const uint64_t dk0 = make_smem_desc(sK + st * H * MX_KT, 16, 1024, 2);
          if (elect_one()) {
#pragma unroll 4
            for (int h = 0; h < H; ++h) {
              const uint64_t dq = dq0 + static_cast<uint32_t>(h * (MX_QT >> 4)), dk = dk0 + static_cast<uint32_t>(h * (MX_KT >> 4));
              for (int ks = 0; ks < ksteps; ++ks) umma_f16_ss(d0 + h * KB, dq + 2u * ks, dk + 2u * ks, idesc_s, ks != 0);
            }
            umma_commit(k_empty(st));
            if (last_of_sb) umma_commit(s_full(G));
          }
          __syncwarp();
          if (last_of_sb) ++sbc;
        }
      }
      // ---- phase B: O_g = A_g V_g
      for (int g = 0; g < H; ++g, ++pcnt) {
        const int st = pcnt & 1;
        mbar_wait(pb_full(st), (pcnt >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t sA = sPB + st * C::PB_STAGE;
        const uint32_t sV = sA + (SPLIT ? 2 : 1) * (MX_MAXNK / 8) * 1024;
        const uint32_t dO = tmem_base + (static_cast<uint32_t>((g / (H / 2)) * 16) << 16) + (g % (H / 2)) * 64;
        // A: no-swizzle K-major, core matrix = 8 rows x 16 bytes contiguous; 8-row groups 128 bytes apart (SBO), the two 8-key
        // halves of a 16-key step 1024 bytes apart (LBO); one step = 2048 bytes.  B: V as delivered, MN-major 128B swizzle.
        const uint64_t da = make_smem_desc(sA, 1024, 128, 0);
        const uint64_t dl = make_smem_desc(sA + (MX_MAXNK / 8) * 1024, 1024, 128, 0);
        const uint64_t dv = make_smem_desc(sV, 8192, 1024, 2);
        if (elect_one()) {
          for (int ks = 0; ks < nk16; ++ks) {
            umma_f16_ss(dO, da + ks * 128u, dv + ks * 128u, idesc_pv, ks != 0);
            if (SPLIT) umma_f16_ss(dO, dl + ks * 128u, dv + ks * 128u, idesc_pv, 1u);
          }
          umma_commit(pb_empty(st));
          if (g == H - 1) umma_commit(o_full);
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    // ===================================================================== mixers
    const int G = warp >> 2, q = warp & 3;
    const int row = q * 16 + (lane & 15), par = lane >> 4;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);   // this warp's lane quarter (thread -> lane q*32 + lane)
    uint32_t sbc = 0, n = 0;                                         // super-blocks consumed so far (S buffer = sbc & 1)
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      const int b = it / tiles, q0 = ((it + b) % tiles) * MX_ROWS;   // see item order below
      const bool active = q * 16 < nq - q0;                        // a warp whose 16 rows all lie past nq only keeps the protocol
      float m_run[H], l_run[H];
#pragma unroll
      for (int h = 0; h < H; ++h) { m_run[h] = -INFINITY; l_run[h] = 0.f; }
      for (int pass = 0; pass < 2; ++pass) {
        for (int sb = 0; sb < nsb; ++sb, ++sbc) {
          const int buf = static_cast<int>(sbc & 1u);
          const uint32_t tS = t_lane + buf * (H * KB);
          mbar_wait(s_full(buf), (sbc >> 1) & 1u);
          tcgen05_fence_after();
          const int blk = 2 * sb + par;
          const bool blk_ok = blk < nblk;
#pragma unroll 1
          for (int r = G * (QPB / 2); r < (G + 1) * (QPB / 2); ++r) {   // this group's half of the block's 4-key quarters
            float x[H][4];
            if (active) {
#pragma unroll
              for (int h = 0; h < H; ++h) {
                uint32_t v[4];
                tmem_ld_32x32b_x4(tS + h * KB + r * 4, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) x[h][j] = __uint_as_float(v[j]);
              }
              tmem_ld_wait();
            }
            if (r == (G + 1) * (QPB / 2) - 1) {                    // this warp's last read of the super-block's S columns
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(s_empty(buf));
            }
            const int key0 = blk * KB + r * 4;
            if (!active || !blk_ok || key0 >= nk16 * 16) continue;   // past the last 16-key step: neither scored nor multiplied
            const int nvalid = nk - key0;                          // keys [key0, key0 + 4) below nk
#pragma unroll
            for (int h = 0; h < H; ++h)
#pragma unroll
              for (int j = 0; j < 4; ++j) x[h][j] *= scale_log2;
            if (VARIANT == 2) {                                    // cait.py:123: dots <- einsum(dots, mix_heads_pre_attn) (linear: commutes with the scale)
              // packed over PAIRS OF OUTPUT HEADS: FFMA2 takes the weight pair (W[h][2g], W[h][2g+1]) from a uniform-register
              // pair (LDCU from the kernel parameters) and the score as a broadcast scalar operand -- half the instructions of
              // the scalar form, no packing moves
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                f32x2 y2[H / 2];
#pragma unroll
                for (int gp = 0; gp < H / 2; ++gp) y2[gp] = 0ull;
#pragma unroll
                for (int h = 0; h < H; ++h) {
                  const f32x2 xs = splat2(x[h][j]);
#pragma unroll
                  for (int gp = 0; gp < H / 2; ++gp) y2[gp] = fma2(pack2(W.wa[h * H + 2 * gp], W.wa[h * H + 2 * gp + 1]), xs, y2[gp]);
                }
#pragma unroll
                for (int gp = 0; gp < H / 2; ++gp) unpack2(y2[gp], x[2 * gp][j], x[2 * gp + 1][j]);
              }
            }
            if (nvalid < 4) {
#pragma unroll
              for (int h = 0; h < H; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j >= nvalid) x[h][j] = -INFINITY;
            }
            if (pass == 0) {
              // ---- softmax statistics (online over this thread's keys)
#pragma unroll
              for (int h = 0; h < H; ++h) {
                const float mx = fmaxf(fmaxf(x[h][0], x[h][1]), fmaxf(x[h][2], x[h][3]));
                const float mn = fmaxf(m_run[h], mx);
                const float ms = (mn == -INFINITY) ? 0.f : mn;
                float s = ex2_approx(x[h][0] - ms) + ex2_approx(x[h][1] - ms) + ex2_approx(x[h][2] - ms) + ex2_approx(x[h][3] - ms);
                l_run[h] = fmaf(l_run[h], ex2_approx(m_run[h] - ms), s);
                m_run[h] = mn;
              }
            } else {
              // ---- probabilities, head mix (+ LayerNorm over heads), bf16, scratch
              const float* st_m = stat + row;
#pragma unroll
              for (int h = 0; h < H; ++h) {
                const float mh = st_m[(h * 2) * MX_ROWS], il = st_m[(h * 2 + 1) * MX_ROWS];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[h][j] = ex2_approx(x[h][j] - mh) * il;
              }
              uint32_t nh[H][2], nl[SPLIT ? H : 1][2];
#pragma unroll
              for (int jp = 0; jp < 2; ++jp) {                     // key pairs (2 jp, 2 jp + 1)
                // the mix of keys (2 jp, 2 jp + 1), packed over pairs of output heads (see the pre-softmax mix above)
                f32x2 b0[H / 2], b1[H / 2];
#pragma unroll
                for (int gp = 0; gp < H / 2; ++gp) { b0[gp] = 0ull; b1[gp] = 0ull; }
#pragma unroll
                for (int h = 0; h < H; ++h) {
                  const f32x2 x0 = splat2(x[h][2 * jp]), x1 = splat2(x[h][2 * jp + 1]);
#pragma unroll
                  for (int gp = 0; gp < H / 2; ++gp) {
                    const f32x2 w2 = (VARIANT == 2) ? pack2(W.wb[h * H + 2 * gp], W.wb[h * H + 2 * gp + 1])
                                                    : pack2(W.wa[h * H + 2 * gp], W.wa[h * H + 2 * gp + 1]);
                    b0[gp] = fma2(w2, x0, b0[gp]);
                    b1[gp] = fma2(w2, x1, b1[gp]);
                  }
                }
                if (VARIANT == 1) {                                // deepvit.py:84: LayerNorm over the head axis, eps 1e-3
                  f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
                  for (int gp = 0; gp < H / 2; ++gp) { s0 = add2(s0, b0[gp]); s1 = add2(s1, b1[gp]); }
                  float t0, t1, t2, t3;
                  unpack2(s0, t0, t1); unpack2(s1, t2, t3);
                  const float mu0 = (t0 + t1) * (1.0f / H), mu1 = (t2 + t3) * (1.0f / H);
                  const f32x2 nm0 = splat2(-mu0), nm1 = splat2(-mu1);
                  f32x2 v0 = 0ull, v1 = 0ull;
#pragma unroll
                  for (int gp = 0; gp < H / 2; ++gp) {
                    b0[gp] = add2(b0[gp], nm0); b1[gp] = add2(b1[gp], nm1);
                    v0 = fma2(b0[gp], b0[gp], v0); v1 = fma2(b1[gp], b1[gp], v1);
                  }
                  unpack2(v0, t0, t1); unpack2(v1, t2, t3);
                  const f32x2 r0 = splat2(rsqrtf((t0 + t1) * (1.0f / H) + 1e-3f)), r1 = splat2(rsqrtf((t2 + t3) * (1.0f / H) + 1e-3f));
#pragma unroll
                  for (int gp = 0; gp < H / 2; ++gp) {
                    const f32x2 g2 = pack2(W.gamma[2 * gp], W.gamma[2 * gp + 1]), be2 = pack2(W.beta[2 * gp], W.beta[2 * gp + 1]);
                    b0[gp] = fma2(mul2(b0[gp], r0), g2, be2);
                    b1[gp] = fma2(mul2(b1[gp], r1), g2, be2);
                  }
                }
                float a0[H], a1[H];
#pragma unroll
                for (int gp = 0; gp < H / 2; ++gp) { unpack2(b0[gp], a0[2 * gp], a0[2 * gp + 1]); unpack2(b1[gp], a1[2 * gp], a1[2 * gp + 1]); }
#pragma unroll
                for (int g = 0; g < H; ++g) {
                  const uint32_t hb = pack_bf16x2(a0[g], a1[g]);
                  nh[g][jp] = hb;
                  if (SPLIT) nl[g][jp] = pack_bf16x2(a0[g] - bf16_lo(hb), a1[g] - bf16_hi(hb));
                }
              }
              const int kg = blk * (KB / 8) + (r >> 1);
              uint8_t* dst = slot + (static_cast<size_t>(kg) * MX_ROWS + row) * 16;
              // 8 bytes per head (and plane) and quarter: holding a quarter back for 16-byte stores costs 2 x H registers
              // that the 168-register budget (320 threads) does not have at H = 16; the slot is L2-resident either way
#pragma unroll
              for (int g = 0; g < H; ++g) {
                st_global_v2(dst + (r & 1) * 8 + static_cast<size_t>(g) * nkg * 1024, nh[g][0], nh[g][1]);
                if (SPLIT) st_global_v2(dst + (r & 1) * 8 + plane_bytes + static_cast<size_t>(g) * nkg * 1024, nl[g][0], nl[g][1]);
              }
            }
          }
        }
        if (pass == 0) {
          // ---- merge the statistics of the four threads of a row: the other key-block parity (lane ^ 16), then the other group
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const float mo = __shfl_xor_sync(0xffffffffu, m_run[h], 16), lo_ = __shfl_xor_sync(0xffffffffu, l_run[h], 16);
            const float mn = fmaxf(m_run[h], mo);
            const float ms = (mn == -INFINITY) ? 0.f : mn;
            l_run[h] = l_run[h] * ex2_approx(m_run[h] - ms) + lo_ * ex2_approx(mo - ms);
            m_run[h] = mn;
          }
          if (G == 1 && par == 0) {
#pragma unroll
            for (int h = 0; h < H; ++h) { part[(h * 2) * MX_ROWS + row] = m_run[h]; part[(h * 2 + 1) * MX_ROWS + row] = l_run[h]; }
          }
          named_bar_sync(1, 256);
          if (G == 0 && par == 0) {
#pragma unroll
            for (int h = 0; h < H; ++h) {
              const float mo = part[(h * 2) * MX_ROWS + row], lo_ = part[(h * 2 + 1) * MX_ROWS + row];
              const float mn = fmaxf(m_run[h], mo);
              const float ms = (mn == -INFINITY) ? 0.f : mn;
              const float l = l_run[h] * ex2_approx(m_run[h] - ms) + lo_ * ex2_approx(mo - ms);
              stat[(h * 2) * MX_ROWS + row] = ms;
              stat[(h * 2 + 1) * MX_ROWS + row] = 1.0f / l;
            }
          }
          named_bar_sync(1, 256);
        }
      }
      // ---- the scratch slot is complete: publish it to the async proxy (bulk copies of phase B), then wait for the products
      fence_proxy_async_all();
      __syncwarp();
      if (lane == 0) mbar_arrive(pa_done);
      mbar_wait(o_full, n & 1u);
      tcgen05_fence_after();
      if (active) {
        const int r_glob = q0 + row;
        const int half = par;                                       // lanes 16-31 of a quarter hold the second half of the heads
#pragma unroll 1
        for (int sl = G * (H / 4); sl < (G + 1) * (H / 4); ++sl) {
          const int g = half * (H / 2) + sl;
          __nv_bfloat16* orow = out + (static_cast<size_t>(b) * nq + r_glob) * ldo + g * dh;
#pragma unroll 1
          for (int c = 0; c < dh; c += 16) {
            uint32_t ov[16];
            tmem_ld_32x32b_x16(t_lane + sl * 64 + c, ov);
            tmem_ld_wait();
            if (r_glob < nq) {
              st_global_v4(orow + c, pack_bf16x2(__uint_as_float(ov[0]), __uint_as_float(ov[1])), pack_bf16x2(__uint_as_float(ov[2]), __uint_as_float(ov[3])),
                           pack_bf16x2(__uint_as_float(ov[4]), __uint_as_float(ov[5])), pack_bf16x2(__uint_as_float(ov[6]), __uint_as_float(ov[7])));
              st_global_v4(orow + c + 8, pack_bf16x2(__uint_as_float(ov[8]), __uint_as_float(ov[9])), pack_bf16x2(__uint_as_float(ov[10]), __uint_as_float(ov[11])),
                           pack_bf16x2(__uint_as_float(ov[12]), __uint_as_float(ov[13])), pack_bf16x2(__uint_as_float(ov[14]), __uint_as_float(ov[15])));
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == MX_MMA) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// scratch for the mixed attention weights, one per (device, stream): kernels on different streams of one device (two handles
// driven by two threads, the two half-batches of a split forward) may run concurrently and each CTA owns slot blockIdx.x of ITS
// launch's scratch.  Grows, never shrinks; allocated outside any stream capture (the first call of a shape is always eager).
struct MixScratch { void* p = nullptr; size_t bytes = 0; };
MixScratch& mix_scratch(int dev, cudaStream_t s) {
  static std::map<std::pair<int, cudaStream_t>, MixScratch> m;
  return m[std::make_pair(dev, s)];
}

template <int H, int VARIANT, bool SPLIT>
void launch_mix(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const MixParams& P, __nv_bfloat16* out, int ldo,
                int B, int nq, int nk, int dh, float scale_log2, cudaStream_t s) {
  using C = MxCfg<H, SPLIT>;
  auto kern = attn_mix_kernel<H, VARIANT, SPLIT>;
  static unsigned long long seen[4] = {0, 0, 0, 0};
  if (first_use_on_this_device(seen)) VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  MixW<H> W;
  for (int i = 0; i < H * H; ++i) { W.wa[i] = P.wa[i]; W.wb[i] = P.wb[i]; }
  for (int i = 0; i < H; ++i) { W.gamma[i] = P.gamma[i]; W.beta[i] = P.beta[i]; }
  const int tiles = (nq + MX_ROWS - 1) / MX_ROWS;
  const int num_items = B * tiles;
  const int nsm = sm_count();
  const int grid = num_items < nsm ? num_items : nsm;
  const int nk16 = (nk + 15) / 16;
  const long long slot_bytes = static_cast<long long>(SPLIT ? 2 : 1) * H * (2 * nk16) * 1024;
  int dev = 0;
  VB_CUDA(cudaGetDevice(&dev));
  void* scratch = nullptr;
  {
    std::lock_guard<std::mutex> lock(global_cache_mutex());
    MixScratch& ms = mix_scratch(dev, s);
    const size_t need = static_cast<size_t>(nsm) * slot_bytes;
    if (need > ms.bytes) {
      if (ms.p) { cudaDeviceSynchronize(); cudaFree(ms.p); ms.p = nullptr; ms.bytes = 0; }
      VB_CUDA(cudaMalloc(&ms.p, need));
      ms.bytes = need;
    }
    scratch = ms.p;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(MX_THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VB_CUDA(cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, W, out, ldo, static_cast<uint8_t*>(scratch), slot_bytes, B, nq, nk, dh, num_items,
                             scale_log2));
  count_launch();
}

}  // namespace

// false when the shape is not covered (the caller falls back to the three-kernel row path)
bool attention_mix(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                   __nv_bfloat16* out, int ldo, int B, int nq, int nk, int heads, int dh, int variant, const float* mix_a,
                   const float* mix_b, const float* ln_g, const float* ln_b, cudaStream_t s, float scale) {
  if (variant != 1 && variant != 2) return false;
  if (heads != 8 && heads != 16) return false;
  if (dh % 16 != 0 || dh > 64 || dh < 16 || nk > MX_MAXNK || nk < 1 || nq < 1) return false;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8) || ((heads * dh) % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) % 16) return false;
  if (getenv("VB_NO_ATTN_MIX") != nullptr) return false;
  MixParams P;
  if (!attention_mix_params(mix_a, mix_b, ln_g, ln_b, heads, s, &P)) return false;
  const uint64_t inner = static_cast<uint64_t>(heads) * dh;
  // 3-D maps over (columns, rows of one image, image): rows past n are zero-filled instead of bleeding into the next image;
  // boxes are 64 columns wide whatever dh is (the columns past a head's dh are never multiplied)
  const CUtensorMap tq = make_tmap_3d(q, inner, nq, B, static_cast<uint64_t>(ldq) * 2, static_cast<uint64_t>(nq) * ldq * 2, 64, MX_ROWS, 1);
  const CUtensorMap tk = make_tmap_3d(k, inner, nk, B, static_cast<uint64_t>(ldk) * 2, static_cast<uint64_t>(nk) * ldk * 2, 64,
                                      heads == 8 ? MxKB<8>::value : MxKB<16>::value, 1);
  const CUtensorMap tv = make_tmap_3d(v, inner, nk, B, static_cast<uint64_t>(ldv) * 2, static_cast<uint64_t>(nk) * ldv * 2, 64, 128, 1);
  const float scale_log2 = (scale > 0.f ? scale : 1.0f / sqrtf(static_cast<float>(dh))) * 1.4426950408889634f;
  // DeepViT's post-LayerNorm attention weights are O(1) with mixed signs (not probabilities).  One bf16 plane (2^-9 relative, like
  // every other activation of the engine) measured the same end-to-end error as hi + lo planes at DeepViT-24 (max |err| 0.037 /
  // 0.039 vs 0.038 / 0.044 on logits of std 1, profiles/r02_config_size_parity.json) at half the scratch traffic: the second
  // plane is opt-in (VB_ATTN_MIX_SPLIT=1).
  const bool split = variant == 1 && getenv("VB_ATTN_MIX_SPLIT") != nullptr;
  if (heads == 8) {
    if (variant == 2) launch_mix<8, 2, false>(tq, tk, tv, P, out, ldo, B, nq, nk, dh, scale_log2, s);
    else if (split) launch_mix<8, 1, true>(tq, tk, tv, P, out, ldo, B, nq, nk, dh, scale_log2, s);
    else launch_mix<8, 1, false>(tq, tk, tv, P, out, ldo, B, nq, nk, dh, scale_log2, s);
  } else {
    if (variant == 2) launch_mix<16, 2, false>(tq, tk, tv, P, out, ldo, B, nq, nk, dh, scale_log2, s);
    else if (split) launch_mix<16, 1, true>(tq, tk, tv, P, out, ldo, B, nq, nk, dh, scale_log2, s);
    else launch_mix<16, 1, false>(tq, tk, tv, P, out, ldo, B, nq, nk, dh, scale_log2, s);
  }
  return true;
}

}  // namespace vb
