// Attention entry points: the generic materialised-scores path (any shape / variant, both precisions) and the
// tcgen05 fast path (bf16; returns false when the shape is not covered so the caller falls back to generic).
#pragma once
#include "common.h"

namespace vb {

// variant 0: softmax(QK^T*scale)V                    (vit.py:77-82, cross_vit.py:87-91)
// variant 1: DeepViT re-attention: softmax -> head mix (mix_a [h,h]) -> LayerNorm over heads (deepvit.py:79-87)
// variant 2: CaiT talking heads: mix_a before softmax, mix_b after (cait.py:121-127)
template <typename T>
void attention_generic(const T* q, int ldq, const T* k, int ldk, const T* v, int ldv, T* out, int ldo, float* S, int B, int nq,
                       int nk, int heads, int dh, int variant, const float* mix_a, const float* mix_b, const float* ln_gamma,
                       const float* ln_beta, cudaStream_t s);

// scale: softmax scale; <= 0 means dh^-0.5 (vit.py:57).  A layer whose heads were zero-padded to the kernels' head width
// (engine.cu: dh 48 -> 64) passes its true dim_head^-0.5 here.
template <typename T>
bool attention_fast(const T* q, int ldq, const T* k, int ldk, const T* v, int ldv, T* out, int ldo, int B, int nq, int nk,
                    int heads, int dh, int variant, const float* mix_a, const float* mix_b, const float* ln_gamma,
                    const float* ln_beta, cudaStream_t s, float scale = 0.f);

// The talking-heads / re-attention path keeps host copies of the head-mix weights keyed by their device pointers; whoever
// frees or rewrites such weights (vb_finalize, vb_destroy, the op-level test entry) must drop them.
void attention_mix_cache_clear();

// Tensor-core (mma.sync) + fused-middle version of attention_generic for the bf16 engine; false if the shape is not covered.
bool attention_generic_mma(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                           __nv_bfloat16* out, int ldo, float* S, int B, int nq, int nk, int heads, int dh, int variant,
                           const float* mix_a, const float* mix_b, const float* ln_gamma, const float* ln_beta, cudaStream_t s);

// Head-mix weights as kernel parameters (heads <= 16): wa = pre-softmax mix (CaiT) / re-attention weights (DeepViT), wb = CaiT's
// post-softmax mix, both [h][g] with row pitch `heads`; gamma / beta = DeepViT's LayerNorm over heads.
struct MixParams {
  float wa[256];
  float wb[256];
  float gamma[16], beta[16];
};
bool attention_mix_params(const float* mix_a, const float* mix_b, const float* ln_gamma, const float* ln_beta, int heads,
                          cudaStream_t s, MixParams* out);

// Fused tcgen05 head-mixing attention (attn_mix_tcgen05.cu): variants 1 / 2, heads 8 / 16, dim_head <= 64, nk <= 256.
bool attention_mix(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                   __nv_bfloat16* out, int ldo, int B, int nq, int nk, int heads, int dh, int variant, const float* mix_a,
                   const float* mix_b, const float* ln_gamma, const float* ln_beta, cudaStream_t s, float scale = 0.f);

// One-kernel attention for a single query row per image (CaiT class attention, CrossViT cross attention; any variant):
// attn_cls.cu.  false if the shape is not covered.
bool attention_cls(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                   __nv_bfloat16* out, int ldo, int B, int nk, int heads, int dh, int variant, const float* mix_a,
                   const float* mix_b, const float* ln_gamma, const float* ln_beta, cudaStream_t s, float scale = 0.f);

long long*& attn_trace_buffer();   // debugging aid: device trace buffer of the tcgen05 attention kernel (null = off)

}  // namespace vb
