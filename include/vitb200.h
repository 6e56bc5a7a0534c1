/* libvitb200 -- C-ABI of the B200-native ViT-family forward engine.
 *
 * The reference (taki0112/vit-tensorflow) has NO plugin / FFI interface: its hot path sits behind plain
 * Python classes.  The boundary preserved is the constructor + call surface
 *     ViT(...)(img) -> logits        vit_tensorflow/vit.py:107-108,159-177
 *     DeepViT(...)(img)              vit_tensorflow/deepvit.py:113-114,139-157
 *     CaiT(...)(img)                 vit_tensorflow/cait.py:156-157,180-194
 *     CrossViT(...)(img)             vit_tensorflow/cross_vit.py:233-253,290-303
 *     parallel_vit.ViT(...)(img)     vit_tensorflow/parallel_vit.py:120-133,167-185   
 *     T2TViT(...)(img)               vit_tensorflow/t2t.py:50-54,96-116               
 *     vit_with_patch_merger.ViT(...)(img)   vit_tensorflow/vit_with_patch_merger.py:134-146,174-185   (8f, f4)
 *     efficient.ViT(...)(img)        vit_tensorflow/efficient.py:13-14,39-55 = vb_forward_embed -> caller's transformer -> vb_forward_head
 * and this header is what the Python host classes (vit_tensorflow_b200/models.py, _lib.py) bind with ctypes.
 * Plain pointers and sizes only; no torch / C++ types cross the boundary.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is vb_last_error(handle)
 *     (pass NULL for errors of vb_create / the vb_op_* helpers).  Nothing throws or aborts across the ABI.
 *   - a handle is bound to one CUDA device and is not thread-safe; distinct handles are independent.
 *   - weights: caller keeps ownership of host arrays, the engine copies/packs them to the device.
 *   - images are NHWC float32 (vit.py:159, usage vit.py:193), logits float32 [batch, num_classes].
 */
#ifndef VITB200_H_
#define VITB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_ABI_VERSION 4
#if defined(__GNUC__)
#define VB_API __attribute__((visibility("default")))
#else
#define VB_API
#endif

typedef struct vb_handle vb_handle;

enum { VB_KIND_VIT = 0, VB_KIND_DEEPVIT = 1, VB_KIND_CAIT = 2, VB_KIND_CROSSVIT = 3, VB_KIND_PARALLEL_VIT = 4,
       VB_KIND_PATCH_MERGER_VIT = 5, VB_KIND_T2T_VIT = 6 };
enum { VB_PRECISION_FP32 = 0, VB_PRECISION_BF16 = 1 };
enum { VB_POOL_CLS = 0, VB_POOL_MEAN = 1 };
enum { VB_MEM_HOST = 0, VB_MEM_DEVICE = 1 };

/* Constructor kwargs of the reference classes, flattened.  Unused fields are ignored per kind. */
typedef struct vb_config {
  int32_t struct_size;            /* sizeof(vb_config), ABI guard */
  int32_t kind;                   /* VB_KIND_* */
  int32_t precision;              /* VB_PRECISION_*: FP32 = exact-fp32 SIMT path (numerics gate); BF16 = tcgen05 path */
  int32_t image_h, image_w;       /* vit.py:133 pair(image_size) */
  int32_t patch_h, patch_w;       /* vit.py:134 pair(patch_size) */
  int32_t channels;               /* 3 */
  int32_t num_classes;
  int32_t dim, depth, heads, dim_head, mlp_dim;
  int32_t pool;                   /* VB_POOL_* (vit.py:139,170-173) */
  int32_t cls_depth;              /* CaiT (cait.py:156) */
  int32_t max_batch;              /* workspace is sized for this batch */
  /* CrossViT (cross_vit.py:233-253) */
  int32_t sm_dim, lg_dim;
  int32_t sm_patch_size, sm_enc_depth, sm_enc_heads, sm_enc_mlp_dim, sm_enc_dim_head;
  int32_t lg_patch_size, lg_enc_depth, lg_enc_heads, lg_enc_mlp_dim, lg_enc_dim_head;
  int32_t cross_attn_depth, cross_attn_heads, cross_attn_dim_head;
  int32_t cross_depth;            /* CrossViT `depth`: number of multi-scale blocks */
  int32_t parallel_branches;      /* parallel ViT `num_parallel_branches` (parallel_vit.py:130); ignored by the other kinds */
  /* vit_with_patch_merger.ViT (vit_with_patch_merger.py:104-109,141-142) */
  int32_t patch_merge_layer_index;   /* default(patch_merge_layer, depth // 2) - 1: merge AFTER this layer; outside [0, depth) = never */
  int32_t patch_merge_num_tokens;
  /* T2TViT `t2t_layers` (t2t.py:54): up to 4 (kernel_size, stride) soft-split layers; image_h == image_w, patch_* unused */
  int32_t t2t_num_layers;
  int32_t t2t_k0, t2t_s0, t2t_k1, t2t_s1, t2t_k2, t2t_s2, t2t_k3, t2t_s3;
} vb_config;

VB_API int vb_abi_version(void);

/* Replaces <Model>.__init__ (vit.py:107-157 etc.): validates the config and allocates device state. */
VB_API int vb_create(const vb_config* cfg, int device, vb_handle** out);

/* Replaces Keras variable assignment: one call per weight, names/shapes/layouts per SURVEY.md App. B
 * (Dense kernel [in,out], float32).  shape/ndim are checked against the config. */
VB_API int vb_set_weight(vb_handle* h, const char* name, const float* host_data, const int64_t* shape, int32_t ndim);

/* Number of weights the config expects, and the name/shape of the i-th (for the host side to enumerate). */
VB_API int vb_num_weights(vb_handle* h);
VB_API int vb_weight_info(vb_handle* h, int32_t index, const char** name, int64_t* shape4, int32_t* ndim);

/* Packs weights for the device (bf16 K-major copies, folded constants).  Fails if a weight is missing. */
VB_API int vb_finalize(vb_handle* h);

/* Replaces <Model>.call(img) (vit.py:159-177, deepvit.py:139-157, cait.py:180-194, cross_vit.py:290-303)
 * with inference semantics (dropout = identity).  img: NHWC float32 [batch,h,w,channels] in host or device
 * memory (img_mem); logits: float32 [batch,num_classes] written to host or device memory (logits_mem).
 * stream: a cudaStream_t (may be NULL = default stream).  With device buffers the call is asynchronous on
 * `stream`; with a host logits buffer it returns after the copy completes.
 * h, w may be smaller than the configured image (pos_embedding[:, :n+1] truncation, vit.py:165). */
VB_API int vb_forward(vb_handle* h, const float* img, int32_t img_mem, int32_t batch, int32_t img_h, int32_t img_w,
               float* logits, int32_t logits_mem, void* stream);

/* Replaces model.transformer(tokens) (vit.py:99-104) for ViT / DeepViT with an arbitrary token count n,
 * the entry the reference's wrappers use (mae.py:69, simmim.py:116, mpp.py:212).
 * tokens/out: float32 [batch, n, dim]. */
VB_API int vb_forward_tokens(vb_handle* h, const float* tokens, int32_t tokens_mem, int32_t batch, int32_t n,
                      float* out, int32_t out_mem, void* stream);

/* DistillMixin.call with a distillation token (reference: distill.py:16-45, DistillableViT distill.py:47-58; ViT only):
 * patch embedding + cls + positions, the token appended as the LAST row, the transformer over n + 2 rows, then
 * logits = mlp_head(pool(x[:, :-1])) and distill_out = x[:, -1].
 * distill_token: HOST float32 [dim] (the caller's trainable variable, distill.py:133).  logits [batch, num_classes] and
 * distill_out [batch, dim] live in `out_mem` memory.  img as in vb_forward. */
VB_API int vb_forward_distill(vb_handle* h, const float* img, int32_t img_mem, int32_t batch, int32_t img_h, int32_t img_w,
                       const float* distill_token, float* logits, float* distill_out, int32_t out_mem, void* stream);

/* ---- the stages of <Model>.call on their own: the attribute surface the reference's wrappers and the
 * injected-transformer shell use.  ViT / DeepViT / parallel ViT / CaiT / patch-merger ViT / T2TViT; not CrossViT. ----------- */

/* Number of token rows vb_forward_embed produces for an img_h x img_w image (patches + cls where the model has one);
 * negative on error. */
VB_API int vb_embed_rows(vb_handle* h, int32_t img_h, int32_t img_w);

/* Everything `call` does before `self.transformer` (vit.py:160-166, cait.py:181-184, t2t.py:97-103, efficient.py:40-45,
 * vit_with_patch_merger.py:175-179): patch embedding (T2T: the whole tokens-to-token module), cls token, positions.
 * tokens: float32 [batch, vb_embed_rows, dim]. */
VB_API int vb_forward_embed(vb_handle* h, const float* img, int32_t img_mem, int32_t batch, int32_t img_h, int32_t img_w,
                     float* tokens, int32_t tokens_mem, void* stream);

/* Everything `call` does after `self.transformer` (vit.py:170-175, efficient.py:48-55): pooling (cls row / mean over
 * the n rows, per the config) + mlp_head (LayerNorm, Dense).  tokens float32 [batch, n, dim] -> logits [batch, num_classes].
 * With n == 1 this is `model.mlp_head(x)`. */
VB_API int vb_forward_head(vb_handle* h, const float* tokens, int32_t tokens_mem, int32_t batch, int32_t n, float* logits,
                    int32_t logits_mem, void* stream);

/* `patch_embedding.layers[0]` (the einops Rearrange, vit.py:142; mae.py:37 `to_patch`): img -> float32
 * [batch, num_patches, patch_h*patch_w*channels].  Not for T2TViT. */
VB_API int vb_to_patch(vb_handle* h, const float* img, int32_t img_mem, int32_t batch, int32_t img_h, int32_t img_w,
                float* patches, int32_t patches_mem, void* stream);

/* `patch_embedding.layers[1]` / `.layers[-1]` (the Dense, vit.py:143; mae.py:37 `patch_to_emb`, mpp.py:200):
 * patches float32 [rows, patch_dim] -> float32 [rows, dim] (patch_dim = the Dense's input width; T2T: the last soft split's). */
VB_API int vb_patch_to_emb(vb_handle* h, const float* patches, int32_t patches_mem, int32_t rows, float* out, int32_t out_mem,
                    void* stream);

/* ---- data parallel: one handle per GPU (one process or thread each), the batch sharded contiguously, the
 * forward free of communication, ONE in-place NCCL all-gather of the fp32 logits at the end.  The reference has no
 * distributed path; these entries are what its maintainer would call from a multi-process launcher.  NCCL is loaded at
 * vb_dp_init time with dlopen ($VB_NCCL_LIB, else libnccl.so.2): the library itself has no load-time NCCL dependency. ------ */

/* Rank 0: fill id128 (128 bytes, an ncclUniqueId) and hand it to every rank by any transport (file, socket, MPI, gloo). */
VB_API int vb_dp_unique_id(void* id128);

/* Every rank (collective; blocks until all `world` ranks arrive): joins the handle's device to the communicator. */
VB_API int vb_dp_init(vb_handle* h, const void* id128, int32_t rank, int32_t world);

/* vb_forward on this rank's shard of `local_batch` images, its logits written straight into rows
 * [rank*local_batch, (rank+1)*local_batch) of `gathered` (DEVICE float32 [world*local_batch, num_classes]), then
 * ncclAllGather in place on `stream`.  Asynchronous on `stream`. */
VB_API int vb_forward_allgather(vb_handle* h, const float* img, int32_t img_mem, int32_t local_batch, int32_t img_h, int32_t img_w,
                         float* gathered, void* stream);

/* Kernels launched by this handle's most recent forward call. */
VB_API int64_t vb_last_launch_count(vb_handle* h);

/* Per-kernel-class device timing (CUDA events recorded on the launch stream around every launch of the class)
 * for the roofline report.  Classes: 0 tcgen05 GEMM (plain / LayerNorm-folded epilogue: to_qkv, to_q, to_kv), 1 attention,
 * 2 LayerNorm / row statistics, 3 im2col, 4 other (SIMT fallbacks), 5 tcgen05 GEMM with GELU epilogue (fc1),
 * 6 tcgen05 GEMM with residual epilogue (patch embed, to_out, fc2).
 * vb_profile_read synchronises the device and returns accumulated milliseconds, algorithmic FLOPs, algorithmic
 * bytes and launch counts per class (arrays of VB_PROF_NUM); reset != 0 clears the accumulators. */
#define VB_PROF_NUM 7
VB_API int vb_profile_enable(vb_handle* h, int32_t on);
VB_API int vb_profile_read(vb_handle* h, double* ms, double* flops, double* bytes, int64_t* calls, int32_t reset);

VB_API const char* vb_last_error(vb_handle* h);
VB_API void vb_destroy(vb_handle* h);

/* ---- single-operator entry points (used by the parity tests and the per-kernel benchmarks) -------------
 * All buffers are HOST float32; bf16 variants round operands to bf16 (RNE) on the way in.
 * *elapsed_ms (may be NULL) receives the average device time per launch over `iters` launches (CUDA events). */

/* out[M,N] = epi(a[M,K] x w[K,N]): epi = (+bias[N]) -> exact-erf GELU (gelu!=0) -> (*scale[N]) -> (+res[M,N]).
 * Any of bias/scale/res may be NULL.  precision BF16 runs the tcgen05 kernel, FP32 the SIMT kernel. */
VB_API int vb_op_linear(int32_t precision, const float* a, const float* w, const float* bias, const float* scale,
                 const float* res, int32_t gelu, float* out, int32_t M, int32_t N, int32_t K, int32_t iters,
                 float* elapsed_ms);

/* Multi-head attention core (vit.py:77-82): q [B,nq,h*dh], k,v [B,nk,h*dh] -> out [B,nq,h*dh];
 * variant 0 = plain, 1 = DeepViT re-attention (mix [h,h] + LayerNorm over heads, gamma/beta [h]; deepvit.py:83-84),
 * 2 = CaiT talking heads (mix_pre, mix_post [h,h]; cait.py:123-125). */
VB_API int vb_op_attention(int32_t precision, int32_t variant, const float* q, const float* k, const float* v,
                    const float* mix_a, const float* mix_b, const float* ln_gamma, const float* ln_beta,
                    float* out, int32_t B, int32_t nq, int32_t nk, int32_t heads, int32_t dim_head, int32_t iters,
                    float* elapsed_ms);

/* LayerNorm over the last axis, eps 1e-3 (vit.py:18): x [M,D] -> out [M,D]. */
VB_API int vb_op_layernorm(int32_t precision, const float* x, const float* gamma, const float* beta, float* out,
                    int32_t M, int32_t D, int32_t iters, float* elapsed_ms);

/* PatchMerger.call (vit_with_patch_merger.py:49-55): x [B,n,D] -> LayerNorm -> softmax(queries [nt,D] . x^T * D^-0.5) . x
 * -> out [B,nt,D]. */
VB_API int vb_op_patch_merger(int32_t precision, const float* x, const float* gamma, const float* beta, const float* queries,
                       float* out, int32_t B, int32_t n, int32_t D, int32_t nt, int32_t iters, float* elapsed_ms);

/* PreNorm + Dense as the bf16 engine runs it (vit.py:18-22 followed by vit.py:39 / :59): LayerNorm over the last axis
 * (eps 1e-3) FOLDED into the following Dense -- gamma scaled into the packed weights, the per-row (mean, rstd) reduced in
 * the GEMM epilogue from (sum, sumsq) partials of the bf16 rows.  x [M,K], w [K,N] (Keras layout), bias [N] or NULL,
 * out [M,N] = act(LN(x) w + bias), act = exact-erf GELU when gelu != 0.  bf16 engine only; N % 64 == 0, K % 64 == 0. */
VB_API int vb_op_ln_linear(const float* x, const float* gamma, const float* beta, const float* w, const float* bias,
                    int32_t gelu, float* out, int32_t M, int32_t N, int32_t K, int32_t iters, float* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* VITB200_H_ */
