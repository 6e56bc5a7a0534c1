// libvitb200 engine: weight registry, workspace arena, forward orchestration of ViT / DeepViT / CaiT / CrossViT
// and the extern "C" boundary declared in include/vitb200.h.
//
// The op sequences restate the reference's call graphs:
//   ViT / DeepViT  vit.py:159-177, deepvit.py:139-157      CaiT  cait.py:180-194      CrossViT  cross_vit.py:290-303
// with inference semantics (dropout = identity).  T = float runs the exact-fp32 SIMT kernels (numerics gate),
// T = __nv_bfloat16 runs the tcgen05 GEMM / attention kernels with fp32 accumulation and statistics.
#include "../../include/vitb200.h"
#include "attention.cuh"
#include "common.h"
#include "kernels.cuh"

#include <dlfcn.h>

#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#ifndef VB_FWD_STREAMS
#define VB_FWD_STREAMS 1                 // 2 = split a large ViT batch into two half-batches on two streams (forward_impl); measured no gain
#endif
#ifndef VB_FWD_SPLIT_MIN_HALF
#define VB_FWD_SPLIT_MIN_HALF 64         // smallest half-batch worth a stream of its own
#endif
namespace vb {
namespace {

thread_local std::string g_last_error;

// ------------------------------------------------------------------------------------------ memory
struct DevMem {
  void* p = nullptr;
  size_t bytes = 0;
  DevMem() = default;
  DevMem(const DevMem&) = delete;
  DevMem& operator=(const DevMem&) = delete;
  ~DevMem() { if (p) cudaFree(p); }
  void ensure(size_t n) {
    if (n <= bytes) return;
    if (p) { cudaFree(p); p = nullptr; bytes = 0; }
    VB_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
};

// Bump allocator over a few large device blocks; reset() per forward keeps pointers stable across calls with
// the same shapes (so cached TMA descriptors stay valid).
class Arena {
 public:
  void reset() { for (auto& b : blocks_) b.off = 0; cur_ = 0; }
  void* alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~static_cast<size_t>(1023);
    for (; cur_ < blocks_.size(); ++cur_) {
      Block& b = blocks_[cur_];
      if (b.off + bytes <= b.mem->bytes) { void* p = static_cast<char*>(b.mem->p) + b.off; b.off += bytes; return p; }
    }
    Block nb;
    nb.mem.reset(new DevMem());
    nb.mem->ensure(bytes > kBlock ? bytes : kBlock);
    nb.off = bytes;
    blocks_.push_back(std::move(nb));
    cur_ = blocks_.size() - 1;
    return blocks_.back().mem->p;
  }
  template <typename T> T* get(size_t count) { return static_cast<T*>(alloc(count * sizeof(T))); }

 private:
  static constexpr size_t kBlock = static_cast<size_t>(256) << 20;
  struct Block { std::unique_ptr<DevMem> mem; size_t off = 0; };
  std::vector<Block> blocks_;
  size_t cur_ = 0;
};

// ------------------------------------------------------------------------------------------ weights
struct Weight {
  std::string name;
  std::vector<int64_t> shape;
  size_t count = 0;
  float* dev = nullptr;   // fp32 copy in the Keras layout
  bool set = false;
};

struct Linear {
  const float* W = nullptr;      // [K, N] fp32
  const float* bias = nullptr;   // [N] or null
  __nv_bfloat16* Wt = nullptr;   // [N, ldw] bf16, K-major, zero padded
  int K = 0, N = 0, ldw = 0;
  // LayerNorm folded into this Dense (bf16 engine): Wt holds gamma-scaled weights, ln_c2 replaces the bias
  const float* ln_c1 = nullptr;
  const float* ln_c2 = nullptr;
};
struct Norm { const float* gamma = nullptr; const float* beta = nullptr; int D = 0; };

struct Epi {
  const float* bias = nullptr;
  const float* scale = nullptr;
  const void* res = nullptr;
  int ldr = 0;
  bool gelu = false;
  const float* ln_stats = nullptr;   // (sum, sumsq) partials per 64-column chunk of the A operand's rows for a LayerNorm-folded Dense, [M, K/64, 2]
  float* stats_out = nullptr;        // emit (sum, sumsq) partials of every 64-column chunk of the output rows, [M, N/64, 2]
};

struct LayerW {                    // one pre-norm transformer layer in any of the four dialects
  Norm attn_norm, ff_norm;
  Linear to_qkv, to_q, to_kv, to_out, fc1, fc2;
  bool fused_qkv = false, project_out = true;
  const float* mix_a = nullptr;    // DeepViT reattn_weights / CaiT mix_pre
  const float* mix_b = nullptr;    // CaiT mix_post
  Norm reattn_norm;                // DeepViT
  const float* attn_scale = nullptr;  // CaiT LayerScale
  const float* ff_scale = nullptr;
  int heads = 0, dim_head = 0, variant = 0;   // dim_head: head width of the q/k/v ACTIVATIONS (the padded width when dh_model < it)
  int dh_model = 0;                  // the model's dim_head: softmax scale dh_model^-0.5 (vit.py:57)
  int t2t_D = 0, t2t_Dp = 0;         // tensor-core T2T soft-split layer: true width D (LayerNorm, softmax scale), padded width Dp
  bool folded = false;               // attn_norm / ff_norm folded into to_qkv (to_q, to_kv) / fc1
};

struct EmbedW { Linear patch; const float* pos = nullptr; const float* cls = nullptr; int dim = 0, n_pos = 0; };
struct CrossW { bool proj = false; Linear project_in, project_out, to_q, to_kv, to_out; Norm norm; };

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace
}  // namespace vb

using namespace vb;

// ------------------------------------------------------------------------------------------ NCCL (loaded on demand)
// The five entry points of the NCCL 2.x C API the data-parallel path needs, declared here so that neither nccl.h nor a
// link-time libnccl is required: ncclUniqueId is 128 opaque bytes passed by value, ncclComm_t an opaque pointer,
// ncclFloat32 == 7, ncclSuccess == 0.
namespace vb {
namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  if (api.ok) return api;
  void* lib = nullptr;
  if (const char* p = getenv("VB_NCCL_LIB")) lib = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  VB_CHECK(lib != nullptr, std::string("NCCL not found (set VB_NCCL_LIB to libnccl.so.2): ") + (dlerror() ? dlerror() : ""));
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(lib, "ncclAllGather"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  VB_CHECK(api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString,
           "the NCCL library lacks one of ncclGetUniqueId / ncclCommInitRank / ncclAllGather / ncclCommDestroy / ncclGetErrorString");
  api.ok = true;
  return api;
}
void nccl_check(int rc, const char* what) {
  if (rc != 0) throw vb::Error(5, std::string(what) + " failed: " + nccl().GetErrorString(rc));
}
}  // namespace
}  // namespace vb

struct vb_handle {
  vb_config cfg;
  int device = 0;
  void* dp_comm = nullptr;          // ncclComm_t of the data-parallel group (vb_dp_init)
  int dp_rank = 0, dp_world = 1;
  bool finalized = false;
  std::string error;
  std::vector<Weight> weights;
  std::map<std::string, int> windex;
  std::vector<std::unique_ptr<DevMem>> owned;
  Arena arena;
  DevMem img_dev, logits_dev, tok_dev, tokens_in, tokens_out;
  long long last_launches = 0;

  // optional per-kernel-class timing: CUDA events recorded on the launch stream around each call
  enum { PROF_GEMM = 0, PROF_ATTN = 1, PROF_LN = 2, PROF_EMBED = 3, PROF_OTHER = 4, PROF_GEMM_GELU = 5, PROF_GEMM_RES = 6, PROF_NUM = 7 };
  struct ProfRec { int cls; cudaEvent_t a, b; double flops; double bytes; };
  bool profiling = false;
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> event_pool;
  double prof_ms[PROF_NUM] = {0}, prof_flops[PROF_NUM] = {0}, prof_bytes[PROF_NUM] = {0};
  long long prof_calls[PROF_NUM] = {0};
  // side streams of the per-image T2T soft-split attention (layer_t2t): forked from / joined into the forward's stream with
  // timing-less events, so that the branch is part of a captured graph like everything else
  static constexpr int T2T_MAX_STREAMS = 4;
  cudaStream_t side_streams[T2T_MAX_STREAMS - 1] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork_event = nullptr, join_events[T2T_MAX_STREAMS - 1] = {nullptr, nullptr, nullptr};
  cudaStream_t half_stream = nullptr;                    // second half-batch of a split forward (forward_impl)
  cudaEvent_t fwd_fork_event = nullptr, fwd_join_event = nullptr;
  void ensure_side_streams() {
    if (fork_event != nullptr) return;
    VB_CUDA(cudaStreamCreateWithFlags(&half_stream, cudaStreamNonBlocking));
    VB_CUDA(cudaEventCreateWithFlags(&fwd_fork_event, cudaEventDisableTiming));
    VB_CUDA(cudaEventCreateWithFlags(&fwd_join_event, cudaEventDisableTiming));
    VB_CUDA(cudaEventCreateWithFlags(&fork_event, cudaEventDisableTiming));
    for (int i = 0; i < T2T_MAX_STREAMS - 1; ++i) {
      VB_CUDA(cudaStreamCreateWithFlags(&side_streams[i], cudaStreamNonBlocking));
      VB_CUDA(cudaEventCreateWithFlags(&join_events[i], cudaEventDisableTiming));
    }
  }
  void destroy_side_streams() {
    if (fork_event == nullptr) return;
    cudaEventDestroy(fork_event);
    cudaEventDestroy(fwd_fork_event); cudaEventDestroy(fwd_join_event); cudaStreamDestroy(half_stream);
    for (int i = 0; i < T2T_MAX_STREAMS - 1; ++i) { cudaEventDestroy(join_events[i]); cudaStreamDestroy(side_streams[i]); }
    fork_event = nullptr;
  }
  cudaEvent_t get_event() {
    if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
    cudaEvent_t e;
    VB_CUDA(cudaEventCreate(&e));
    return e;
  }
  struct ProfScope {
    vb_handle* h; cudaStream_t s; int idx = -1;
    ProfScope(vb_handle* h_, int cls, double flops, double bytes, cudaStream_t s_) : h(h_), s(s_) {
      if (!h->profiling) return;
      ProfRec r{cls, h->get_event(), h->get_event(), flops, bytes};
      cudaEventRecord(r.a, s);
      h->prof_recs.push_back(r);
      idx = static_cast<int>(h->prof_recs.size()) - 1;
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(h->prof_recs[idx].b, s); }
  };
  void prof_collect() {
    if (prof_recs.empty()) return;
    cudaDeviceSynchronize();
    for (auto& r : prof_recs) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
        prof_ms[r.cls] += ms; prof_flops[r.cls] += r.flops; prof_bytes[r.cls] += r.bytes; prof_calls[r.cls] += 1;
      }
      event_pool.push_back(r.a); event_pool.push_back(r.b);
    }
    prof_recs.clear();
  }

  // model structure (filled by finalize)
  EmbedW embed, sm_embed, lg_embed;
  std::vector<LayerW> layers, cls_layers, t2t_layers;   // t2t_layers: the one-layer transformers between the soft splits (t2t.py:35)
  Norm merger_norm;                                     // PatchMerger (vit_with_patch_merger.py:46-47)
  const float* merger_queries = nullptr;
  struct T2TStage { int k, stride, dim; };              // dim = channels * prod(k^2) up to and including this stage (t2t.py:63)
  std::vector<T2TStage> t2t_stages() const {
    const int ks[4] = {cfg.t2t_k0, cfg.t2t_k1, cfg.t2t_k2, cfg.t2t_k3}, ss[4] = {cfg.t2t_s0, cfg.t2t_s1, cfg.t2t_s2, cfg.t2t_s3};
    std::vector<T2TStage> st;
    int d = cfg.channels;
    for (int i = 0; i < cfg.t2t_num_layers; ++i) { d *= ks[i] * ks[i]; st.push_back(T2TStage{ks[i], ss[i], d}); }
    return st;
  }
  static int conv_output_size(int size, int k, int stride) {   // t2t.py:14-15 with padding = stride // 2
    return static_cast<int>((static_cast<double>(size - k + 2 * (stride / 2)) / stride) + 1);
  }
  struct XBlock { std::vector<LayerW> sm_layers, lg_layers; Norm sm_final, lg_final; std::vector<CrossW> sm_attend_lg, lg_attend_sm; };
  std::vector<XBlock> xblocks;
  Norm head_norm, sm_head_norm, lg_head_norm;
  Linear head, sm_head, lg_head;

  // cached patch-embedding residual terms, keyed by (embed ptr, batch, rows)
  struct ResKey { const void* e; int B, rows; bool operator<(const ResKey& o) const { return std::tie(e, B, rows) < std::tie(o.e, o.B, o.rows); } };
  std::map<ResKey, std::unique_ptr<DevMem>> embed_res;

  // cached tcgen05 GEMM plans (TMA descriptors)
  using PlanKey = std::array<uintptr_t, 16>;
  std::map<PlanKey, GemmBf16> plans;

  // Whole-forward CUDA graphs (vb_forward): the launch sequence of one (image pointer, logits pointer, batch, h, w, stream)
  // combination is captured on its SECOND call (the first runs eagerly and does every lazy host-side step: arena growth,
  // kernel attributes, descriptor / residual caches) and replayed from the third on.  Workspace pointers are stable (the
  // arena is reset, never freed, per forward), so a replay touches exactly the memory the eager run would.  Programmatic
  // dependent launches are captured as programmatic edges.  VB_NO_GRAPH=1 disables it; profiling and the legacy default
  // stream (which cannot be captured) always run eagerly.
  struct GraphKey {
    const void* img; void* out; int B, H, W; cudaStream_t s;
    bool operator<(const GraphKey& o) const { return std::tie(img, out, B, H, W, s) < std::tie(o.img, o.out, o.B, o.H, o.W, o.s); }
  };
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int calls = 0; long long launches = 0; bool failed = false; };
  std::map<GraphKey, GraphEntry> graphs;
  void drop_graphs() {
    for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    graphs.clear();
  }

  bool bf16() const { return cfg.precision == VB_PRECISION_BF16; }

  // ---------------------------------------------------------------- weight registry
  void expect(const std::string& name, std::vector<int64_t> shape) {
    Weight w;
    w.name = name;
    w.shape = std::move(shape);
    w.count = 1;
    for (auto d : w.shape) w.count *= static_cast<size_t>(d);
    windex[name] = static_cast<int>(weights.size());
    weights.push_back(std::move(w));
  }
  void expect_dense(const std::string& n, int din, int dout, bool bias = true) {
    expect(n + ".kernel", {din, dout});
    if (bias) expect(n + ".bias", {dout});
  }
  void expect_ln(const std::string& n, int d) { expect(n + ".gamma", {d}); expect(n + ".beta", {d}); }
  void expect_layer(const std::string& pre, int dim, int heads, int dh, int mlp, int kind) {
    const int inner = heads * dh;
    expect_ln(pre + "attn_norm", dim);
    if (kind == VB_KIND_VIT || kind == VB_KIND_DEEPVIT) expect_dense(pre + "to_qkv", dim, 3 * inner, false);
    else { expect_dense(pre + "to_q", dim, inner, false); expect_dense(pre + "to_kv", dim, 2 * inner, false); }
    if (kind == VB_KIND_DEEPVIT) { expect(pre + "reattn_weights", {heads, heads}); expect_ln(pre + "reattn_norm", heads); }
    if (kind == VB_KIND_CAIT) { expect(pre + "mix_pre", {heads, heads}); expect(pre + "mix_post", {heads, heads}); }
    const bool po = !(kind == VB_KIND_VIT && heads == 1 && dh == dim);   // vit.py:53
    if (po) expect_dense(pre + "to_out", inner, dim);
    expect_ln(pre + "ff_norm", dim);
    expect_dense(pre + "fc1", dim, mlp);
    expect_dense(pre + "fc2", mlp, dim);
  }
  void build_expected() {
    const vb_config& c = cfg;
    const int C = c.channels;
    if (c.kind == VB_KIND_VIT || c.kind == VB_KIND_DEEPVIT || c.kind == VB_KIND_PARALLEL_VIT) {
      const int np = (c.image_h / c.patch_h) * (c.image_w / c.patch_w);
      expect("pos_embedding", {1, np + 1, c.dim});
      expect("cls_token", {1, 1, c.dim});
      expect_dense("patch", c.patch_h * c.patch_w * C, c.dim);
      for (int L = 0; L < c.depth; ++L) {
        if (c.kind == VB_KIND_PARALLEL_VIT) {   // branch i = (attention fn i, feed-forward fn i) of layer L (parallel_vit.py:109-112)
          for (int i = 0; i < c.parallel_branches; ++i)
            expect_layer("layers." + std::to_string(L) + ".branch" + std::to_string(i) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, VB_KIND_VIT);
        } else {
          expect_layer("layers." + std::to_string(L) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, c.kind);
        }
      }
      expect_ln("head_norm", c.dim);
      expect_dense("head", c.dim, c.num_classes);
    } else if (c.kind == VB_KIND_PATCH_MERGER_VIT) {
      const int np = (c.image_h / c.patch_h) * (c.image_w / c.patch_w);
      expect("pos_embedding", {1, np + 1, c.dim});                       // vit_with_patch_merger.py:163
      expect_dense("patch", c.patch_h * c.patch_w * C, c.dim);
      for (int L = 0; L < c.depth; ++L) expect_layer("layers." + std::to_string(L) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, VB_KIND_VIT);
      expect_ln("patch_merger.norm", c.dim);
      expect("patch_merger.queries", {c.patch_merge_num_tokens, c.dim});
      expect_ln("head_norm", c.dim);
      expect_dense("head", c.dim, c.num_classes);
    } else if (c.kind == VB_KIND_T2T_VIT) {
      const auto st = t2t_stages();
      int out = c.image_h;
      for (size_t i = 0; i < st.size(); ++i) {
        out = conv_output_size(out, st[i].k, st[i].stride);             // t2t.py:66
        if (i + 1 < st.size())                                           // Transformer(dim=d, heads=1, depth=1, dim_head=d, mlp_dim=d) t2t.py:69-70
          expect_layer("t2t." + std::to_string(i) + ".layers.0.", st[i].dim, 1, st[i].dim, st[i].dim, VB_KIND_VIT);
      }
      expect_dense("patch", st.back().dim, c.dim);                        // t2t.py:73
      expect("pos_embedding", {1, out * out + 1, c.dim});                 // :76
      expect("cls_token", {1, 1, c.dim});
      for (int L = 0; L < c.depth; ++L) expect_layer("layers." + std::to_string(L) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, VB_KIND_VIT);
      expect_ln("head_norm", c.dim);
      expect_dense("head", c.dim, c.num_classes);
    } else if (c.kind == VB_KIND_CAIT) {
      const int np = (c.image_h / c.patch_h) * (c.image_w / c.patch_w);
      expect("pos_embedding", {1, np, c.dim});
      expect("cls_token", {1, 1, c.dim});
      expect_dense("patch", c.patch_h * c.patch_w * C, c.dim);
      for (int st = 0; st < 2; ++st) {
        const std::string stack = st == 0 ? "patch_transformer" : "cls_transformer";
        const int depth = st == 0 ? c.depth : c.cls_depth;
        for (int L = 0; L < depth; ++L) {
          const std::string pre = stack + ".layers." + std::to_string(L) + ".";
          expect(pre + "attn_scale", {1, 1, c.dim});
          expect(pre + "ff_scale", {1, 1, c.dim});
          expect_layer(pre, c.dim, c.heads, c.dim_head, c.mlp_dim, VB_KIND_CAIT);
        }
      }
      expect_ln("head_norm", c.dim);
      expect_dense("head", c.dim, c.num_classes);
    } else {
      for (int b = 0; b < 2; ++b) {
        const std::string br = b == 0 ? "sm" : "lg";
        const int dim = b == 0 ? c.sm_dim : c.lg_dim, p = b == 0 ? c.sm_patch_size : c.lg_patch_size;
        const int np = (c.image_h / p) * (c.image_w / p);
        expect_dense(br + "_embed.patch", p * p * C, dim);
        expect(br + "_embed.pos_embedding", {1, np + 1, dim});
        expect(br + "_embed.cls_token", {1, 1, dim});
      }
      const int xinner = c.cross_attn_heads * c.cross_attn_dim_head;
      for (int D = 0; D < c.cross_depth; ++D) {
        const std::string bp = "blocks." + std::to_string(D) + ".";
        for (int b = 0; b < 2; ++b) {
          const std::string br = b == 0 ? "sm" : "lg";
          const int dim = b == 0 ? c.sm_dim : c.lg_dim;
          const int depth = b == 0 ? c.sm_enc_depth : c.lg_enc_depth, heads = b == 0 ? c.sm_enc_heads : c.lg_enc_heads;
          const int dh = b == 0 ? c.sm_enc_dim_head : c.lg_enc_dim_head, mlp = b == 0 ? c.sm_enc_mlp_dim : c.lg_enc_mlp_dim;
          for (int L = 0; L < depth; ++L) expect_layer(bp + br + "_enc.layers." + std::to_string(L) + ".", dim, heads, dh, mlp, VB_KIND_CROSSVIT);
          expect_ln(bp + br + "_enc.final_norm", dim);
        }
        for (int R = 0; R < c.cross_attn_depth; ++R) {
          for (int b = 0; b < 2; ++b) {
            const std::string n = bp + "cross." + std::to_string(R) + (b == 0 ? ".sm_attend_lg." : ".lg_attend_sm.");
            const int din = b == 0 ? c.sm_dim : c.lg_dim, dout = b == 0 ? c.lg_dim : c.sm_dim;
            if (din != dout) { expect_dense(n + "project_in", din, dout); expect_dense(n + "project_out", dout, din); }
            expect_ln(n + "norm", dout);
            expect_dense(n + "to_q", dout, xinner, false);
            expect_dense(n + "to_kv", dout, 2 * xinner, false);
            expect_dense(n + "to_out", xinner, dout);
          }
        }
      }
      expect_ln("sm_head_norm", c.sm_dim); expect_dense("sm_head", c.sm_dim, c.num_classes);
      expect_ln("lg_head_norm", c.lg_dim); expect_dense("lg_head", c.lg_dim, c.num_classes);
    }
  }

  // finalize-time replacements of registered weights by derived device copies (head-padded kernels)
  std::map<std::string, const float*> woverride;
  const float* W(const std::string& name) const {
    auto ov = woverride.find(name);
    if (ov != woverride.end()) return ov->second;
    auto it = windex.find(name);
    VB_CHECK(it != windex.end(), "internal: unknown weight " + name);
    return weights[it->second].dev;
  }
  bool has(const std::string& name) const { return windex.count(name) != 0; }

  Linear make_linear(const std::string& n, int K, int N, bool bias = true, const Norm* fold = nullptr, int ldw_min = 0) {
    Linear L;
    L.W = W(n + ".kernel");
    L.bias = bias ? W(n + ".bias") : nullptr;
    L.K = K; L.N = N; L.ldw = round_up(K, 8) > ldw_min ? round_up(K, 8) : ldw_min;
    if (bf16()) {
      owned.emplace_back(new DevMem());
      owned.back()->ensure(static_cast<size_t>(N) * L.ldw * sizeof(__nv_bfloat16));
      L.Wt = static_cast<__nv_bfloat16*>(owned.back()->p);
      pack_weight_bf16(L.W, L.Wt, K, N, L.ldw, 0, fold ? fold->gamma : nullptr);
      if (fold) {
        owned.emplace_back(new DevMem());
        owned.back()->ensure(static_cast<size_t>(N) * 2 * sizeof(float));
        float* c = static_cast<float*>(owned.back()->p);
        ln_fold_consts(L.W, L.Wt, L.ldw, fold->beta, L.bias, c, c + N, K, N, 0);
        L.ln_c1 = c; L.ln_c2 = c + N;
      }
    }
    return L;
  }
  // Two bias-free Dense layers applied to the same input (CaiT to_q / to_kv on x, cait.py:114-119) as ONE tcgen05 GEMM:
  // the packed K-major weights and the LayerNorm-fold constants of `a` and `b` are laid out back to back, so the output
  // columns are [a | b] = [q | k | v], the layout the fused to_qkv of vit.py produces.  bf16 engine only.
  Linear make_linear_pair(const std::string& na, int NA, const std::string& nb, int NB, int K, const Norm* fold) {
    Linear L;
    L.W = nullptr; L.bias = nullptr;
    L.K = K; L.N = NA + NB; L.ldw = round_up(K, 8);
    owned.emplace_back(new DevMem());
    owned.back()->ensure(static_cast<size_t>(L.N) * L.ldw * sizeof(__nv_bfloat16));
    L.Wt = static_cast<__nv_bfloat16*>(owned.back()->p);
    float* c = nullptr;
    if (fold) {
      owned.emplace_back(new DevMem());
      owned.back()->ensure(static_cast<size_t>(L.N) * 2 * sizeof(float));
      c = static_cast<float*>(owned.back()->p);
      L.ln_c1 = c; L.ln_c2 = c + L.N;
    }
    const std::string names[2] = {na, nb};
    const int widths[2] = {NA, NB};
    int n0 = 0;
    for (int i = 0; i < 2; ++i) {
      const float* Wi = W(names[i] + ".kernel");
      __nv_bfloat16* Wti = L.Wt + static_cast<size_t>(n0) * L.ldw;
      pack_weight_bf16(Wi, Wti, K, widths[i], L.ldw, 0, fold ? fold->gamma : nullptr);
      if (fold) ln_fold_consts(Wi, Wti, L.ldw, fold->beta, nullptr, c + n0, c + L.N + n0, K, widths[i], 0);
      n0 += widths[i];
    }
    return L;
  }
  Norm make_norm(const std::string& n, int D) { return Norm{W(n + ".gamma"), W(n + ".beta"), D}; }
  // fold_ok: the layer is only ever used as a self-attention layer (its LayerNorms feed nothing but GEMMs)
  LayerW make_layer(const std::string& pre, int dim, int heads, int dh, int mlp, int kind, bool fold_ok = true) {
    LayerW l;
    l.dh_model = dh;
    // Plain softmax attention with dim_head < 64 (bf16 engine): widen every head to the tcgen05 attention kernel's 64 columns
    // with zero weights -- extra to_qkv / to_q / to_kv output columns (q, k, v pad columns are exactly 0, so QK^T and the
    // real PV columns are unchanged) and zero to_out input rows.  Costs (64 / dh - 1) more flops in those two GEMMs and buys
    // the tensor-core attention path for e.g. dim_head 48 / 32 (the head-mixing variants have their own kernel, attn_mix).
    const bool to_out_present = has(pre + "to_out.kernel");
    if (bf16() && dh < 64 && dh % 8 == 0 && to_out_present && dim % 8 == 0 && getenv("VB_NO_HEAD_PAD") == nullptr &&
        (kind == VB_KIND_VIT || kind == VB_KIND_CROSSVIT)) {
      const int dhp = 64;
      auto padded = [&](const std::string& n, int other, int groups, int pad_rows) {
        owned.emplace_back(new DevMem());
        owned.back()->ensure(static_cast<size_t>(other) * groups * heads * dhp * sizeof(float));
        float* wp = static_cast<float*>(owned.back()->p);
        pad_heads_f32(W(n), wp, other, groups, heads, dh, dhp, pad_rows, 0);
        woverride[n] = wp;
      };
      if (kind == VB_KIND_VIT) padded(pre + "to_qkv.kernel", dim, 3, 0);
      else { padded(pre + "to_q.kernel", dim, 1, 0); padded(pre + "to_kv.kernel", dim, 2, 0); }
      padded(pre + "to_out.kernel", dim, 1, 1);
      dh = dhp;
    }
    const int inner = heads * dh;
    l.heads = heads; l.dim_head = dh;
    l.attn_norm = make_norm(pre + "attn_norm", dim);
    l.ff_norm = make_norm(pre + "ff_norm", dim);
    // LayerNorm folding needs every GEMM of the layer on the tcgen05 path (all widths multiples of 64)
    l.folded = bf16() && fold_ok && dim % 64 == 0 && inner % 64 == 0 && mlp % 64 == 0 && getenv("VB_NO_LN_FOLD") == nullptr;
    const Norm* fa = l.folded ? &l.attn_norm : nullptr;
    const Norm* ff = l.folded ? &l.ff_norm : nullptr;
    if (kind == VB_KIND_VIT || kind == VB_KIND_DEEPVIT) { l.fused_qkv = true; l.to_qkv = make_linear(pre + "to_qkv", dim, 3 * inner, false, fa); }
    else if (bf16() && fold_ok && dim % 8 == 0) {   // self-attention only: q and kv read the same rows -> one GEMM
      l.fused_qkv = true;
      l.to_qkv = make_linear_pair(pre + "to_q", inner, pre + "to_kv", 2 * inner, dim, fa);
    } else { l.to_q = make_linear(pre + "to_q", dim, inner, false, fa); l.to_kv = make_linear(pre + "to_kv", dim, 2 * inner, false, fa); }
    if (kind == VB_KIND_DEEPVIT) { l.variant = 1; l.mix_a = W(pre + "reattn_weights"); l.reattn_norm = make_norm(pre + "reattn_norm", heads); }
    if (kind == VB_KIND_CAIT) {
      l.variant = 2; l.mix_a = W(pre + "mix_pre"); l.mix_b = W(pre + "mix_post");
      l.attn_scale = W(pre + "attn_scale"); l.ff_scale = W(pre + "ff_scale");
    }
    l.project_out = has(pre + "to_out.kernel");
    if (l.project_out) l.to_out = make_linear(pre + "to_out", inner, dim);
    l.fc1 = make_linear(pre + "fc1", dim, mlp, true, ff);
    l.fc2 = make_linear(pre + "fc2", mlp, dim);
    return l;
  }
  // One-layer transformer between two T2T soft splits (t2t.py:35,45-46: heads = 1, dim_head = mlp_dim = dim = D = channels * prod(k^2),
  // 147 and 1323 at the default t2t_layers; no out-projection, vit.py:53) on the tensor cores: every width is zero-padded to
  // Dp = round_up(D, 64) -- token rows [n, Dp] with zero pad columns, to_qkv columns [q | k | v] each Dp wide, fc1 / fc2 Dp x Dp --
  // so that all four GEMMs and the per-image QK^T / PV products run on the tcgen05 GEMM kernel; LayerNorm and the softmax scale keep
  // the true D.  Zero weights and biases keep the pad columns exactly zero through the layer (GELU(0) = 0).
  static bool t2t_tensor_path_enabled() { static const bool off = getenv("VB_NO_T2T_TC") != nullptr; return !off; }
  LayerW make_t2t_layer(const std::string& pre, int D) {
    const int Dp = round_up(D, 64);
    auto padded = [&](const std::string& n, int other, int groups, int pad_rows, const float* src, int dh_src) {
      owned.emplace_back(new DevMem());
      owned.back()->ensure(static_cast<size_t>(other) * groups * Dp * sizeof(float));
      float* wp = static_cast<float*>(owned.back()->p);
      pad_heads_f32(src, wp, other, groups, 1, dh_src, Dp, pad_rows, 0);
      woverride[n] = wp;
      return wp;
    };
    padded(pre + "to_qkv.kernel", D, 3, 0, W(pre + "to_qkv.kernel"), D);                    // [D, 3 D] -> [D, 3 Dp]
    padded(pre + "fc1.kernel", D, 1, 0, W(pre + "fc1.kernel"), D);                          // [D, D] -> [D, Dp]
    padded(pre + "fc1.bias", 1, 1, 0, W(pre + "fc1.bias"), D);
    const float* rows_padded = padded(pre + "fc2.kernel", D, 1, 1, W(pre + "fc2.kernel"), D);   // [D, D] -> [Dp, D] (zero input rows)
    padded(pre + "fc2.kernel", Dp, 1, 0, rows_padded, D);                                   // -> [Dp, Dp]
    padded(pre + "fc2.bias", 1, 1, 0, W(pre + "fc2.bias"), D);
    LayerW l;
    l.heads = 1; l.dim_head = Dp; l.dh_model = D; l.t2t_D = D; l.t2t_Dp = Dp;
    l.attn_norm = make_norm(pre + "attn_norm", D);
    l.ff_norm = make_norm(pre + "ff_norm", D);
    l.fused_qkv = true; l.project_out = false;
    l.to_qkv = make_linear(pre + "to_qkv", D, 3 * Dp, false, nullptr, Dp);
    l.fc1 = make_linear(pre + "fc1", D, Dp, true, nullptr, Dp);
    l.fc2 = make_linear(pre + "fc2", Dp, Dp, true, nullptr, Dp);
    l.to_qkv.K = Dp; l.fc1.K = Dp;                                                        // the A operands are Dp wide (zero pad columns)
    return l;
  }
  EmbedW make_embed(const std::string& pre, int p_h, int p_w, int dim, int n_pos, bool with_cls = true, int K = 0) {
    EmbedW e;
    e.patch = make_linear(pre + "patch", K > 0 ? K : p_h * p_w * cfg.channels, dim);
    e.pos = W(pre + "pos_embedding");
    e.cls = with_cls ? W(pre + "cls_token") : nullptr;   // CaiT adds its cls token after the patch stage (cait.py:189)
    e.dim = dim; e.n_pos = n_pos;
    return e;
  }

  void finalize() {
    for (auto& w : weights) VB_CHECK(w.set, "vb_finalize: weight '" + w.name + "' was never set");
    VB_CUDA(cudaSetDevice(device));
    owned.clear(); layers.clear(); cls_layers.clear(); t2t_layers.clear(); xblocks.clear(); plans.clear(); embed_res.clear();
    woverride.clear();
    drop_graphs();
    const vb_config& c = cfg;
    if (c.kind == VB_KIND_VIT || c.kind == VB_KIND_DEEPVIT || c.kind == VB_KIND_PARALLEL_VIT) {
      const int np = (c.image_h / c.patch_h) * (c.image_w / c.patch_w);
      embed = make_embed("", c.patch_h, c.patch_w, c.dim, np + 1);
      for (int L = 0; L < c.depth; ++L) {
        if (c.kind == VB_KIND_PARALLEL_VIT) {   // `layers` holds depth x parallel_branches entries, branches of a layer adjacent
          for (int i = 0; i < c.parallel_branches; ++i)
            layers.push_back(make_layer("layers." + std::to_string(L) + ".branch" + std::to_string(i) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, VB_KIND_VIT));
        } else {
          layers.push_back(make_layer("layers." + std::to_string(L) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, c.kind));
        }
      }
      head_norm = make_norm("head_norm", c.dim);
      head = make_linear_f32("head", c.dim, c.num_classes);
    } else if (c.kind == VB_KIND_PATCH_MERGER_VIT) {
      const int np = (c.image_h / c.patch_h) * (c.image_w / c.patch_w);
      embed = make_embed("", c.patch_h, c.patch_w, c.dim, np + 1, false);
      for (int L = 0; L < c.depth; ++L) layers.push_back(make_layer("layers." + std::to_string(L) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, VB_KIND_VIT));
      merger_norm = make_norm("patch_merger.norm", c.dim);
      merger_queries = W("patch_merger.queries");
      head_norm = make_norm("head_norm", c.dim);
      head = make_linear_f32("head", c.dim, c.num_classes);
    } else if (c.kind == VB_KIND_T2T_VIT) {
      const auto st = t2t_stages();
      int out = c.image_h;
      for (size_t i = 0; i < st.size(); ++i) {
        out = conv_output_size(out, st[i].k, st[i].stride);
        if (i + 1 < st.size()) {
          const std::string pre = "t2t." + std::to_string(i) + ".layers.0.";
          if (bf16() && t2t_tensor_path_enabled()) t2t_layers.push_back(make_t2t_layer(pre, st[i].dim));
          else t2t_layers.push_back(make_layer(pre, st[i].dim, 1, st[i].dim, st[i].dim, VB_KIND_VIT, false));
        }
      }
      embed = make_embed("", 0, 0, c.dim, out * out + 1, true, st.back().dim);
      for (int L = 0; L < c.depth; ++L) layers.push_back(make_layer("layers." + std::to_string(L) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, VB_KIND_VIT));
      head_norm = make_norm("head_norm", c.dim);
      head = make_linear_f32("head", c.dim, c.num_classes);
    } else if (c.kind == VB_KIND_CAIT) {
      const int np = (c.image_h / c.patch_h) * (c.image_w / c.patch_w);
      embed = make_embed("", c.patch_h, c.patch_w, c.dim, np, false);
      for (int L = 0; L < c.depth; ++L) layers.push_back(make_layer("patch_transformer.layers." + std::to_string(L) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, VB_KIND_CAIT));
      for (int L = 0; L < c.cls_depth; ++L) cls_layers.push_back(make_layer("cls_transformer.layers." + std::to_string(L) + ".", c.dim, c.heads, c.dim_head, c.mlp_dim, VB_KIND_CAIT, false));
      head_norm = make_norm("head_norm", c.dim);
      head = make_linear_f32("head", c.dim, c.num_classes);
    } else {
      const int nps = (c.image_h / c.sm_patch_size) * (c.image_w / c.sm_patch_size);
      const int npl = (c.image_h / c.lg_patch_size) * (c.image_w / c.lg_patch_size);
      sm_embed = make_embed("sm_embed.", c.sm_patch_size, c.sm_patch_size, c.sm_dim, nps + 1);
      lg_embed = make_embed("lg_embed.", c.lg_patch_size, c.lg_patch_size, c.lg_dim, npl + 1);
      const int xinner = c.cross_attn_heads * c.cross_attn_dim_head;
      for (int D = 0; D < c.cross_depth; ++D) {
        XBlock xb;
        const std::string bp = "blocks." + std::to_string(D) + ".";
        for (int L = 0; L < c.sm_enc_depth; ++L) xb.sm_layers.push_back(make_layer(bp + "sm_enc.layers." + std::to_string(L) + ".", c.sm_dim, c.sm_enc_heads, c.sm_enc_dim_head, c.sm_enc_mlp_dim, VB_KIND_CROSSVIT));
        for (int L = 0; L < c.lg_enc_depth; ++L) xb.lg_layers.push_back(make_layer(bp + "lg_enc.layers." + std::to_string(L) + ".", c.lg_dim, c.lg_enc_heads, c.lg_enc_dim_head, c.lg_enc_mlp_dim, VB_KIND_CROSSVIT));
        xb.sm_final = make_norm(bp + "sm_enc.final_norm", c.sm_dim);
        xb.lg_final = make_norm(bp + "lg_enc.final_norm", c.lg_dim);
        for (int R = 0; R < c.cross_attn_depth; ++R) {
          for (int b = 0; b < 2; ++b) {
            const std::string n = bp + "cross." + std::to_string(R) + (b == 0 ? ".sm_attend_lg." : ".lg_attend_sm.");
            const int din = b == 0 ? c.sm_dim : c.lg_dim, dout = b == 0 ? c.lg_dim : c.sm_dim;
            CrossW x;
            x.proj = din != dout;
            if (x.proj) { x.project_in = make_linear(n + "project_in", din, dout); x.project_out = make_linear(n + "project_out", dout, din); }
            x.norm = make_norm(n + "norm", dout);
            x.to_q = make_linear(n + "to_q", dout, xinner, false);
            x.to_kv = make_linear(n + "to_kv", dout, 2 * xinner, false);
            x.to_out = make_linear(n + "to_out", xinner, dout);
            (b == 0 ? xb.sm_attend_lg : xb.lg_attend_sm).push_back(x);
          }
        }
        xblocks.push_back(std::move(xb));
      }
      sm_head_norm = make_norm("sm_head_norm", c.sm_dim); sm_head = make_linear_f32("sm_head", c.sm_dim, c.num_classes);
      lg_head_norm = make_norm("lg_head_norm", c.lg_dim); lg_head = make_linear_f32("lg_head", c.lg_dim, c.num_classes);
    }
    VB_CUDA(cudaDeviceSynchronize());
    finalized = true;
  }
  Linear make_linear_f32(const std::string& n, int K, int N) {  // classifier head always runs in fp32
    Linear L;
    L.W = W(n + ".kernel"); L.bias = W(n + ".bias"); L.K = K; L.N = N; L.ldw = K;
    return L;
  }

  // ---------------------------------------------------------------- ops
  template <typename T>
  void linear(const T* A, int lda, int M, const Linear& L, T* out, int ldc, const Epi& e, cudaStream_t s);

  template <typename T>
  void attention(const T* q, int ldq, const T* k, int ldk, const T* v, int ldv, T* out, int ldo, int B, int nq, int nk,
                 const LayerW& l, cudaStream_t s) {
    const float scale = l.dh_model != l.dim_head ? 1.0f / sqrtf(static_cast<float>(l.dh_model)) : 0.f;
    attention_dispatch<T>(q, ldq, k, ldk, v, ldv, out, ldo, B, nq, nk, l.heads, l.dim_head, l.variant, l.mix_a, l.mix_b,
                          l.reattn_norm.gamma, l.reattn_norm.beta, s, scale);
  }
  template <typename T>
  void attention_dispatch(const T* q, int ldq, const T* k, int ldk, const T* v, int ldv, T* out, int ldo, int B, int nq, int nk,
                          int heads, int dh, int variant, const float* mix_a, const float* mix_b, const float* g, const float* b,
                          cudaStream_t s, float scale = 0.f);

  template <typename T>
  const T* embed_residual(const EmbedW& e, int B, int rows, cudaStream_t s) {
    ResKey key{&e, B, rows};
    auto it = embed_res.find(key);
    if (it == embed_res.end()) {
      // one entry per (embedding, batch, rows) actually in use; a server that sweeps batch / image sizes must not grow
      // without bound (77 MB per ViT-B/16 B=256 entry): beyond a handful of shapes start over
      if (embed_res.size() >= 6) { VB_CUDA(cudaStreamSynchronize(s)); embed_res.clear(); drop_graphs(); }
      std::unique_ptr<DevMem> m(new DevMem());
      m->ensure(static_cast<size_t>(B) * rows * e.dim * sizeof(T));
      build_embed_residual<T>(static_cast<T*>(m->p), e.pos, e.cls, e.patch.bias, B, rows, e.dim, e.cls != nullptr, s);
      it = embed_res.emplace(key, std::move(m)).first;
    }
    return static_cast<const T*>(it->second->p);
  }

  // patch embedding + cls token + pos embedding -> X [B*rows, dim]   (vit.py:160-165 / cait.py:181-184)
  template <typename T>
  T* embed_tokens(const EmbedW& e, const float* img, int B, int H, int Wd, int ph, int pw, int* rows_out, cudaStream_t s,
                  float** stats_out = nullptr) {
    VB_CHECK(H % ph == 0 && Wd % pw == 0, "Image dimensions must be divisible by the patch size.");
    const int np = (H / ph) * (Wd / pw);
    const int has_cls = e.cls != nullptr ? 1 : 0;
    const int rows = np + has_cls;
    VB_CHECK(rows <= e.n_pos, "image has more patches than pos_embedding rows");
    const int Kp = bf16() ? e.patch.ldw : e.patch.K;
    const int M = B * rows;
    T* col = arena.get<T>(static_cast<size_t>(M) * Kp);
    {
      ProfScope ps(this, PROF_EMBED, 0.0, 4.0 * B * H * Wd * cfg.channels + static_cast<double>(sizeof(T)) * M * Kp, s);
      im2col<T>(img, col, B, H, Wd, cfg.channels, ph, pw, has_cls, Kp, s);
    }
    *rows_out = rows;
    return embed_from_cols<T>(e, col, Kp, B, rows, s, stats_out);
  }
  // X = cols . W + bias + (pos (+ cls on row 0)): the Dense of the patch embedding with cls concat and positions folded
  // into its residual operand.  cols [B*rows, Kp]: patch vectors, zero in the cls rows and in the pad columns.
  template <typename T>
  T* embed_from_cols(const EmbedW& e, const T* col, int Kp, int B, int rows, cudaStream_t s, float** stats_out) {
    const int M = B * rows;
    const T* R = embed_residual<T>(e, B, rows, s);
    T* X = arena.get<T>(static_cast<size_t>(M) * e.dim);
    Epi ep; ep.bias = e.patch.bias; ep.res = R; ep.ldr = e.dim;
    if (stats_out != nullptr && bf16() && e.dim % 64 == 0 && getenv("VB_NO_LN_FOLD") == nullptr) {
      *stats_out = arena.get<float>(static_cast<size_t>(M) * (e.dim / 64) * 2);
      ep.stats_out = *stats_out;
    }
    Linear L = e.patch;
    L.K = Kp;  // im2col zero-pads the patch vector to the packed weight pitch
    linear<T>(col, Kp, M, L, X, e.dim, ep, s);
    return X;
  }

  // T2TViT.patch_embedding + cls + positions (t2t.py:58-74,97-103): soft split i = unfold_same over the previous token map
  // (the image for i = 0), every split but the last followed by a one-layer transformer of width channels * prod(k^2);
  // the last split writes the im2col operand of the Dense(dim) directly (cls rows / pad columns zero).
  template <typename T>
  T* embed_t2t(const float* img, int B, int H, int Wd, int* rows_out, cudaStream_t s, float** stats_out) {
    const auto st = t2t_stages();
    const T* map = nullptr;
    int mh = H, mw = Wd, mc = cfg.channels, map_ld = 0;
    for (size_t i = 0; i < st.size(); ++i) {
      const int oh = (mh + st[i].stride - 1) / st[i].stride, ow = (mw + st[i].stride - 1) / st[i].stride;
      const int D = st[i].dim, n = oh * ow;
      VB_CHECK(i == 0 || mh == mw, "T2TViT: token maps after the first soft split must be square (t2t.py:41)");
      const bool last = i + 1 == st.size();
      const bool tc = !last && t2t_layers[i].t2t_Dp > 0;               // tensor-core layer: token rows padded to Dp columns
      const int ld = last ? (bf16() ? embed.patch.ldw : embed.patch.K) : (tc ? t2t_layers[i].t2t_Dp : D);
      const int cls_row = last ? 1 : 0;
      T* out = arena.get<T>(static_cast<size_t>(B) * (n + cls_row) * ld);
      {
        ProfScope ps(this, PROF_EMBED, 0.0, static_cast<double>(i == 0 ? 4 : sizeof(T)) * B * mh * mw * mc + static_cast<double>(sizeof(T)) * B * (n + cls_row) * ld, s);
        if (i == 0) unfold_same<float, T>(img, out, B, mh, mw, mc, st[i].k, st[i].stride, cls_row, ld, s);
        else unfold_same<T, T>(map, out, B, mh, mw, mc, st[i].k, st[i].stride, cls_row, ld, s, map_ld);
      }
      if (last) {
        VB_CHECK(n + 1 <= embed.n_pos, "image has more patches than pos_embedding rows");
        *rows_out = n + 1;
        return embed_from_cols<T>(embed, out, ld, B, n + 1, s, stats_out);
      }
      if (tc) layer_t2t<T>(out, B, n, t2t_layers[i], s);
      else layer_self<T>(out, B, n, D, t2t_layers[i], s);
      map = out; mh = oh; mw = ow; mc = D; map_ld = ld;
    }
    VB_CHECK(false, "T2TViT needs at least one t2t layer");
    return nullptr;
  }

  // `call` up to the transformer for every kind with one token stream
  template <typename T>
  T* embed_any(const float* img, int B, int H, int Wd, int* rows_out, cudaStream_t s, float** stats_out) {
    VB_CHECK(cfg.kind != VB_KIND_CROSSVIT, "CrossViT has two token streams: no single embedding stage");
    if (cfg.kind == VB_KIND_T2T_VIT) return embed_t2t<T>(img, B, H, Wd, rows_out, s, stats_out);
    return embed_tokens<T>(embed, img, B, H, Wd, cfg.patch_h, cfg.patch_w, rows_out, s, stats_out);
  }

  // PatchMerger.call (vit_with_patch_merger.py:49-55) = single-head attention of nt learned queries over LN(x) with
  // keys = values = LN(x) and scale dim^-0.5 (:45) -- exactly attention with heads = 1, dim_head = dim.
  template <typename T>
  T* patch_merge(const T* X, int B, int rows, int dim, const Norm& norm, const float* queries, int nt, cudaStream_t s) {
    T* Y = arena.get<T>(static_cast<size_t>(B) * rows * dim);
    ln<T>(X, norm, Y, B * rows, dim, s);
    T* Q = arena.get<T>(static_cast<size_t>(B) * nt * dim);
    broadcast_rows<T>(queries, Q, B, nt, dim, s);
    T* O = arena.get<T>(static_cast<size_t>(B) * nt * dim);
    attention_dispatch<T>(Q, dim, Y, dim, Y, dim, O, dim, B, nt, rows, 1, dim, 0, nullptr, nullptr, nullptr, nullptr, s);
    return O;
  }

  // tcgen05 GEMM on raw operands through the plan cache (the T2T soft-split attention products)
  void gemm_cached(const __nv_bfloat16* A, int lda, const __nv_bfloat16* Wt, int ldw, int b_rows, void* out, int ldc, int M, int N, int K,
                   const __nv_bfloat16* res, int ldr, bool out_f32, int cls, cudaStream_t s) {
    ProfScope ps(this, cls, 2.0 * M * N * K, 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K) + (out_f32 ? 4.0 : 2.0) * M * N, s);
    PlanKey key{};
    const uintptr_t parts[16] = {reinterpret_cast<uintptr_t>(A), static_cast<uintptr_t>(lda), reinterpret_cast<uintptr_t>(Wt),
                                 reinterpret_cast<uintptr_t>(out), static_cast<uintptr_t>(ldc), static_cast<uintptr_t>(M),
                                 static_cast<uintptr_t>(N), static_cast<uintptr_t>(K), static_cast<uintptr_t>(ldw),
                                 static_cast<uintptr_t>(b_rows), reinterpret_cast<uintptr_t>(res), static_cast<uintptr_t>(ldr),
                                 static_cast<uintptr_t>(out_f32), 0x7247u, 0, 0};
    for (int i = 0; i < 16; ++i) key[i] = parts[i];
    auto it = plans.find(key);
    if (it == plans.end()) {
      if (plans.size() > 8192) plans.clear();
      it = plans.emplace(key, gemm_bf16_plan(A, lda, Wt, ldw, static_cast<__nv_bfloat16*>(out), ldc, M, N, K, nullptr, nullptr, res, ldr,
                                             false, out_f32, b_rows)).first;
    }
    gemm_bf16_run(it->second, s);
  }

  // One T2T soft-split transformer layer on the tensor cores (see make_t2t_layer).  X [B*n, Dp] bf16, zero pad columns, updated
  // in place.  Attention with ONE head of width D = 147 / 1323 over n = 3136 / 784 tokens (t2t.py:35) does not fit the fused
  // attention kernels' head widths: per image, S = Q K^T (tcgen05 GEMM, fp32 out), softmax rows -> bf16 P, O = P V (tcgen05 GEMM
  // against V^T, residual X added in its epilogue).  The score / probability buffers of ONE image (39 + 20 MB at n = 3136) are
  // reused for every image, so they stay in the 126 MB L2 instead of streaming B x n x n floats through HBM.
  template <typename T>
  void layer_t2t(T* X, int B, int n, const LayerW& l, cudaStream_t s);

  // one pre-norm layer, self-attention over all rows (vit.py:101-102, cait.py:150-151, cross_vit.py:109-111).
  // `stats` (bf16 engine, folded layers): per-row (sum, sumsq) partials of X, valid on entry iff *stats_valid; the
  // residual GEMMs keep them up to date, so no LayerNorm kernel runs at all.
  template <typename T>
  void layer_self(T* X, int B, int rows, int dim, const LayerW& l, cudaStream_t s, float* stats = nullptr,
                  bool* stats_valid = nullptr) {
    const int M = B * rows, inner = l.heads * l.dim_head;
    const bool fold = l.folded && stats != nullptr;
    if (fold && !*stats_valid) { ensure_stats(X, dim, stats, M, s); *stats_valid = true; }
    T* Y = arena.get<T>(static_cast<size_t>(M) * dim);
    T* O = arena.get<T>(static_cast<size_t>(M) * inner);
    const T* A = X;
    Epi eq;
    if (fold) eq.ln_stats = stats;
    else { VB_CHECK(!l.folded, "internal: folded layer without statistics"); ln<T>(X, l.attn_norm, Y, M, dim, s); A = Y; }
    if (l.fused_qkv) {
      T* QKV = arena.get<T>(static_cast<size_t>(M) * 3 * inner);
      Epi e = eq; e.bias = l.to_qkv.ln_c2;
      linear<T>(A, dim, M, l.to_qkv, QKV, 3 * inner, e, s);
      attention<T>(QKV, 3 * inner, QKV + inner, 3 * inner, QKV + 2 * inner, 3 * inner, O, inner, B, rows, rows, l, s);
    } else {
      T* Q = arena.get<T>(static_cast<size_t>(M) * inner);
      T* KV = arena.get<T>(static_cast<size_t>(M) * 2 * inner);
      Epi e1 = eq; e1.bias = l.to_q.ln_c2;
      Epi e2 = eq; e2.bias = l.to_kv.ln_c2;
      linear<T>(A, dim, M, l.to_q, Q, inner, e1, s);
      linear<T>(A, dim, M, l.to_kv, KV, 2 * inner, e2, s);
      attention<T>(Q, inner, KV, 2 * inner, KV + inner, 2 * inner, O, inner, B, rows, rows, l, s);
    }
    if (l.project_out) {
      Epi e; e.bias = l.to_out.bias; e.scale = l.attn_scale; e.res = X; e.ldr = dim;
      if (fold) e.stats_out = stats;
      linear<T>(O, inner, M, l.to_out, X, dim, e, s);
    } else {
      add_tokens<T>(X, O, static_cast<long long>(M) * dim, s);   // vit.py:53: identity out-projection
      if (fold) ensure_stats(X, dim, stats, M, s);
    }
    feed_forward<T>(X, M, dim, l, Y, s, fold ? stats : nullptr);
  }
  // One parallel_vit layer (parallel_vit.py:114-117): x = sum_i attn_i(LN_i(x)) + x ; x = sum_i ff_i(LN'_i(x)) + x.
  // Every branch reads the SAME x, so the sums accumulate in a second buffer through the residual epilogue of the
  // branch's last GEMM (branch 0: res = x, out = acc; branch i > 0: res = out = acc); the attention half goes X -> Acc, the
  // feed-forward half Acc -> X.  The branches' LayerNorms share the row statistics of x, so folding still applies.
  template <typename T>
  void layer_parallel(T* X, T* Acc, int B, int rows, int dim, const LayerW* br, int nbr, cudaStream_t s, float* stats, bool* stats_valid) {
    const int M = B * rows, inner = br[0].heads * br[0].dim_head;
    const bool fold = br[0].folded && stats != nullptr;
    if (fold && !*stats_valid) { ensure_stats(X, dim, stats, M, s); *stats_valid = true; }
    T* Y = arena.get<T>(static_cast<size_t>(M) * dim);
    T* O = arena.get<T>(static_cast<size_t>(M) * inner);
    T* QKV = arena.get<T>(static_cast<size_t>(M) * 3 * inner);
    T* Hb = arena.get<T>(static_cast<size_t>(M) * br[0].fc1.N);
    // ---- attention branches: X -> Acc
    for (int i = 0; i < nbr; ++i) {
      const LayerW& l = br[i];
      const T* A = X;
      Epi eq;
      if (fold) { eq.ln_stats = stats; eq.bias = l.to_qkv.ln_c2; }
      else { VB_CHECK(!l.folded, "internal: folded layer without statistics"); ln<T>(X, l.attn_norm, Y, M, dim, s); A = Y; }
      linear<T>(A, dim, M, l.to_qkv, QKV, 3 * inner, eq, s);
      attention<T>(QKV, 3 * inner, QKV + inner, 3 * inner, QKV + 2 * inner, 3 * inner, O, inner, B, rows, rows, l, s);
      const bool last = i == nbr - 1;
      if (l.project_out) {
        Epi e; e.bias = l.to_out.bias; e.res = i == 0 ? X : Acc; e.ldr = dim;
        if (fold && last) e.stats_out = stats;
        linear<T>(O, inner, M, l.to_out, Acc, dim, e, s);
      } else {                                               // parallel_vit.py:74: Identity out-projection (inner == dim)
        if (i == 0) VB_CUDA(cudaMemcpyAsync(Acc, X, static_cast<size_t>(M) * dim * sizeof(T), cudaMemcpyDeviceToDevice, s));
        add_tokens<T>(Acc, O, static_cast<long long>(M) * dim, s);
        if (fold && last) ensure_stats(Acc, dim, stats, M, s);
      }
    }
    // ---- feed-forward branches: Acc -> X
    for (int i = 0; i < nbr; ++i) {
      const LayerW& l = br[i];
      const T* A = Acc;
      Epi e1; e1.gelu = true;
      if (fold) { e1.ln_stats = stats; e1.bias = l.fc1.ln_c2; }
      else { ln<T>(Acc, l.ff_norm, Y, M, dim, s); A = Y; e1.bias = l.fc1.bias; }
      linear<T>(A, dim, M, l.fc1, Hb, l.fc1.N, e1, s);
      Epi e2; e2.bias = l.fc2.bias; e2.res = i == 0 ? Acc : X; e2.ldr = dim;
      if (fold && i == nbr - 1) e2.stats_out = stats;
      linear<T>(Hb, l.fc1.N, M, l.fc2, X, dim, e2, s);
    }
  }
  template <typename T>
  void ensure_stats(const T* X, int dim, float* stats, int M, cudaStream_t s);
  template <typename T>
  void ln(const T* x, const Norm& n, T* y, int M, int dim, cudaStream_t s) {
    ProfScope ps(this, PROF_LN, 0.0, 2.0 * sizeof(T) * M * dim, s);
    layernorm<T>(x, dim, n.gamma, n.beta, y, dim, M, dim, s);
  }
  template <typename T>
  void feed_forward(T* X, int M, int dim, const LayerW& l, T* Y, cudaStream_t s, float* stats = nullptr) {
    T* Hb = arena.get<T>(static_cast<size_t>(M) * l.fc1.N);
    Epi e1; e1.gelu = true;
    const T* A = X;
    if (stats != nullptr) { e1.ln_stats = stats; e1.bias = l.fc1.ln_c2; }
    else { VB_CHECK(!l.folded, "internal: folded layer without statistics"); ln<T>(X, l.ff_norm, Y, M, dim, s); A = Y; e1.bias = l.fc1.bias; }
    linear<T>(A, dim, M, l.fc1, Hb, l.fc1.N, e1, s);
    Epi e2; e2.bias = l.fc2.bias; e2.scale = l.ff_scale; e2.res = X; e2.ldr = dim;
    e2.stats_out = stats;
    linear<T>(Hb, l.fc1.N, M, l.fc2, X, dim, e2, s);
  }
  template <typename T>
  void add_tokens(T* X, const T* O, long long count, cudaStream_t s);

  // CaiT class-attention layer: x [B,1,dim] attends over [LN(x) ; patches] (cait.py:57-58,109-112,150-151)
  template <typename T>
  void layer_cls(T* Cx, T* ctx, int B, int nctx, int dim, const LayerW& l, cudaStream_t s) {
    const int inner = l.heads * l.dim_head;
    T* Yc = arena.get<T>(static_cast<size_t>(B) * dim);
    layernorm<T>(Cx, dim, l.attn_norm.gamma, l.attn_norm.beta, Yc, dim, B, dim, s);
    copy_tokens<T>(Yc, 1, 0, ctx, nctx, 0, 1, B, dim, s);
    T* Q = arena.get<T>(static_cast<size_t>(B) * inner);
    T* KV = arena.get<T>(static_cast<size_t>(B) * nctx * 2 * inner);
    T* O = arena.get<T>(static_cast<size_t>(B) * inner);
    linear<T>(Yc, dim, B, l.to_q, Q, inner, Epi(), s);
    linear<T>(ctx, dim, B * nctx, l.to_kv, KV, 2 * inner, Epi(), s);
    attention<T>(Q, inner, KV, 2 * inner, KV + inner, 2 * inner, O, inner, B, 1, nctx, l, s);
    Epi e; e.bias = l.to_out.bias; e.scale = l.attn_scale; e.res = Cx; e.ldr = dim;
    linear<T>(O, inner, B, l.to_out, Cx, dim, e, s);
    feed_forward<T>(Cx, B, dim, l, Yc, s);
  }

  // CrossViT: cls [B,dcls] attends over [LN(Pin(cls)) ; other-branch patches]  (cross_vit.py:128-138,69-93,159-160)
  template <typename T>
  void cross_attend(T* cls, int dcls, T* ctx, int nctx, int dctx, const CrossW& x, int B, cudaStream_t s) {
    const int heads = cfg.cross_attn_heads, dh = cfg.cross_attn_dim_head, inner = heads * dh;
    T* xp = cls;
    if (x.proj) {
      xp = arena.get<T>(static_cast<size_t>(B) * dctx);
      Epi e; e.bias = x.project_in.bias;
      linear<T>(cls, dcls, B, x.project_in, xp, dctx, e, s);
    }
    T* y = arena.get<T>(static_cast<size_t>(B) * dctx);
    layernorm<T>(xp, dctx, x.norm.gamma, x.norm.beta, y, dctx, B, dctx, s);
    copy_tokens<T>(y, 1, 0, ctx, nctx, 0, 1, B, dctx, s);
    T* Q = arena.get<T>(static_cast<size_t>(B) * inner);
    T* KV = arena.get<T>(static_cast<size_t>(B) * nctx * 2 * inner);
    T* O = arena.get<T>(static_cast<size_t>(B) * inner);
    linear<T>(y, dctx, B, x.to_q, Q, inner, Epi(), s);
    linear<T>(ctx, dctx, B * nctx, x.to_kv, KV, 2 * inner, Epi(), s);
    attention_dispatch<T>(Q, inner, KV, 2 * inner, KV + inner, 2 * inner, O, inner, B, 1, nctx, heads, dh, 0, nullptr, nullptr,
                          nullptr, nullptr, s);
    if (x.proj) {
      T* a = arena.get<T>(static_cast<size_t>(B) * dctx);
      Epi e1; e1.bias = x.to_out.bias;
      linear<T>(O, inner, B, x.to_out, a, dctx, e1, s);
      Epi e2; e2.bias = x.project_out.bias; e2.res = cls; e2.ldr = dcls;
      linear<T>(a, dctx, B, x.project_out, cls, dcls, e2, s);
    } else {
      Epi e; e.bias = x.to_out.bias; e.res = cls; e.ldr = dcls;
      linear<T>(O, inner, B, x.to_out, cls, dcls, e, s);
    }
  }

  template <typename T>
  void classify(const T* X, int rows, int dim, const Norm& hn, const Linear& hd, int B, int mean_pool, float* logits,
                bool accumulate, cudaStream_t s) {
    float* z = arena.get<float>(static_cast<size_t>(B) * dim);
    pool_layernorm<T>(X, rows, dim, hn.gamma, hn.beta, z, B, dim, mean_pool, s);
    gemm_simt<float, float, float>(z, dim, hd.W, hd.N, 1, logits, hd.N, B, hd.N, dim, hd.bias, nullptr,
                                   accumulate ? logits : nullptr, hd.N, 0, s);
  }

  // Images are independent, so a large batch can run as two half-batches on two streams (the forward's own and `half_stream`,
  // forked / joined with timing-less events and therefore part of a captured graph): while one half's kernel drains -- last
  // epilogues, CTAs finishing at different times, the dependent launch waiting for the whole grid -- the other half's next
  // kernel already has CTAs on the freed SMs.  Results are bit-identical to the unsplit forward (every kernel's tile
  // arithmetic is independent of the batch, tests: test_batch_independence_and_determinism).  ViT, DeepViT and CaiT only (the
  // head-mixing attention kernel's scratch is per stream; T2T and CrossViT fork streams of their own).
  // MEASURED (profiles/r02_ab_fwd_streams.txt): correct (253 GPU tests, identical logits) but no faster -- ViT-B/16 B = 256
  // 9.02 / 9.14 ms unsplit vs 9.42 / 9.11 ms split, ViT-L/16-384 47.06 vs 47.30 ms: every kernel here is a persistent grid of
  // one CTA per SM, so the second stream's CTAs only get SMs as the first kernel's CTAs exit, and what the overlap of the tails
  // gains the doubled per-launch cost (half the tiles per kernel) loses.  Off by default (VB_FWD_STREAMS=2 enables it).
  template <typename T>
  void forward_impl(const float* img, int B, int H, int Wd, float* logits, cudaStream_t s) {
    arena.reset();
    static const char* fs_env = getenv("VB_FWD_STREAMS");
    const int want = fs_env != nullptr ? atoi(fs_env) : VB_FWD_STREAMS;
    const bool split = want >= 2 && !profiling && bf16() && B >= 2 * VB_FWD_SPLIT_MIN_HALF &&
                       (cfg.kind == VB_KIND_VIT || cfg.kind == VB_KIND_DEEPVIT || cfg.kind == VB_KIND_CAIT);
    if (!split) { forward_body<T>(img, B, H, Wd, logits, s); return; }
    ensure_side_streams();
    const int B0 = (B + 1) / 2;
    VB_CUDA(cudaEventRecord(fwd_fork_event, s));
    VB_CUDA(cudaStreamWaitEvent(half_stream, fwd_fork_event, 0));
    forward_body<T>(img, B0, H, Wd, logits, s);
    forward_body<T>(img + static_cast<size_t>(B0) * H * Wd * cfg.channels, B - B0, H, Wd, logits + static_cast<size_t>(B0) * cfg.num_classes, half_stream);
    VB_CUDA(cudaEventRecord(fwd_join_event, half_stream));
    VB_CUDA(cudaStreamWaitEvent(s, fwd_join_event, 0));
  }

  template <typename T>
  void forward_body(const float* img, int B, int H, int Wd, float* logits, cudaStream_t s) {
    const vb_config& c = cfg;
    if (c.kind == VB_KIND_VIT || c.kind == VB_KIND_DEEPVIT || c.kind == VB_KIND_T2T_VIT) {
      int rows = 0;
      float* stats = nullptr;
      T* X = embed_any<T>(img, B, H, Wd, &rows, s, &stats);
      bool sv = stats != nullptr;
      for (const auto& l : layers) layer_self<T>(X, B, rows, c.dim, l, s, stats, &sv);
      classify<T>(X, rows, c.dim, head_norm, head, B, c.pool == VB_POOL_MEAN, logits, false, s);
    } else if (c.kind == VB_KIND_PATCH_MERGER_VIT) {   // vit_with_patch_merger.py:174-185, Transformer.call :118-126
      int rows = 0;
      float* stats = nullptr;
      T* X = embed_tokens<T>(embed, img, B, H, Wd, c.patch_h, c.patch_w, &rows, s, &stats);
      bool sv = stats != nullptr;
      for (int L = 0; L < c.depth; ++L) {
        layer_self<T>(X, B, rows, c.dim, layers[L], s, stats, &sv);
        if (L == c.patch_merge_layer_index) {
          X = patch_merge<T>(X, B, rows, c.dim, merger_norm, merger_queries, c.patch_merge_num_tokens, s);
          rows = c.patch_merge_num_tokens;
          sv = false;                                    // new token rows: the row statistics are recomputed by the next layer
        }
      }
      classify<T>(X, rows, c.dim, head_norm, head, B, 1, logits, false, s);   // Reduce('b n d -> b d', 'mean') :169
    } else if (c.kind == VB_KIND_PARALLEL_VIT) {
      int rows = 0;
      float* stats = nullptr;
      T* X = embed_tokens<T>(embed, img, B, H, Wd, c.patch_h, c.patch_w, &rows, s, &stats);
      bool sv = stats != nullptr;
      T* Acc = arena.get<T>(static_cast<size_t>(B) * rows * c.dim);
      for (int L = 0; L < c.depth; ++L)
        layer_parallel<T>(X, Acc, B, rows, c.dim, &layers[static_cast<size_t>(L) * c.parallel_branches], c.parallel_branches, s, stats, &sv);
      classify<T>(X, rows, c.dim, head_norm, head, B, c.pool == VB_POOL_MEAN, logits, false, s);
    } else if (c.kind == VB_KIND_CAIT) {
      int rows = 0;
      float* stats = nullptr;
      T* X = embed_tokens<T>(embed, img, B, H, Wd, c.patch_h, c.patch_w, &rows, s, &stats);
      bool sv = stats != nullptr;
      for (const auto& l : layers) layer_self<T>(X, B, rows, c.dim, l, s, stats, &sv);
      T* Cx = arena.get<T>(static_cast<size_t>(B) * c.dim);
      broadcast_row<T>(W("cls_token"), Cx, 1, B, c.dim, s);
      T* ctx = arena.get<T>(static_cast<size_t>(B) * (rows + 1) * c.dim);
      copy_tokens<T>(X, rows, 0, ctx, rows + 1, 1, rows, B, c.dim, s);
      for (const auto& l : cls_layers) layer_cls<T>(Cx, ctx, B, rows + 1, c.dim, l, s);
      classify<T>(Cx, 1, c.dim, head_norm, head, B, 0, logits, false, s);
    } else {
      int ns = 0, nl = 0;
      float *st_s = nullptr, *st_g = nullptr;
      T* S = embed_tokens<T>(sm_embed, img, B, H, Wd, c.sm_patch_size, c.sm_patch_size, &ns, s, &st_s);
      T* G = embed_tokens<T>(lg_embed, img, B, H, Wd, c.lg_patch_size, c.lg_patch_size, &nl, s, &st_g);
      bool sv_s = st_s != nullptr, sv_g = st_g != nullptr;
      T* sm_cls = arena.get<T>(static_cast<size_t>(B) * c.sm_dim);
      T* lg_cls = arena.get<T>(static_cast<size_t>(B) * c.lg_dim);
      T* ctx_lg = arena.get<T>(static_cast<size_t>(B) * nl * c.lg_dim);   // [LN(Pin(sm_cls)) ; lg patches]
      T* ctx_sm = arena.get<T>(static_cast<size_t>(B) * ns * c.sm_dim);   // [LN(Pin(lg_cls)) ; sm patches]
      // The two towers of a multi-scale block are independent until the cross-attention (cross_vit.py:185-190): the large-patch
      // tower (n = 17 at the README configuration: GEMMs of a few tiles) runs on a side stream beside the small-patch one, forked
      // and joined with timing-less events (part of the captured graph).  VB_CROSSVIT_STREAMS=1 keeps one stream.
      static const char* xs_env = getenv("VB_CROSSVIT_STREAMS");
      const bool two = (xs_env != nullptr ? atoi(xs_env) : 2) >= 2 && !profiling && bf16();
      cudaStream_t sg = s;
      if (two) { ensure_side_streams(); sg = side_streams[0]; }
      for (const auto& xb : xblocks) {
        if (two) { VB_CUDA(cudaEventRecord(fork_event, s)); VB_CUDA(cudaStreamWaitEvent(sg, fork_event, 0)); }
        for (const auto& l : xb.sm_layers) layer_self<T>(S, B, ns, c.sm_dim, l, s, st_s, &sv_s);
        layernorm<T>(S, c.sm_dim, xb.sm_final.gamma, xb.sm_final.beta, S, c.sm_dim, B * ns, c.sm_dim, s);
        for (const auto& l : xb.lg_layers) layer_self<T>(G, B, nl, c.lg_dim, l, sg, st_g, &sv_g);
        layernorm<T>(G, c.lg_dim, xb.lg_final.gamma, xb.lg_final.beta, G, c.lg_dim, B * nl, c.lg_dim, sg);
        sv_s = sv_g = false;   // the trailing LayerNorm and the cls write-back below change the token rows
        copy_tokens<T>(G, nl, 0, lg_cls, 1, 0, 1, B, c.lg_dim, sg);
        copy_tokens<T>(G, nl, 1, ctx_lg, nl, 1, nl - 1, B, c.lg_dim, sg);
        if (two) { VB_CUDA(cudaEventRecord(join_events[0], sg)); VB_CUDA(cudaStreamWaitEvent(s, join_events[0], 0)); }
        copy_tokens<T>(S, ns, 0, sm_cls, 1, 0, 1, B, c.sm_dim, s);
        copy_tokens<T>(S, ns, 1, ctx_sm, ns, 1, ns - 1, B, c.sm_dim, s);
        for (size_t R = 0; R < xb.sm_attend_lg.size(); ++R) {
          cross_attend<T>(sm_cls, c.sm_dim, ctx_lg, nl, c.lg_dim, xb.sm_attend_lg[R], B, s);
          cross_attend<T>(lg_cls, c.lg_dim, ctx_sm, ns, c.sm_dim, xb.lg_attend_sm[R], B, s);
        }
        copy_tokens<T>(sm_cls, 1, 0, S, ns, 0, 1, B, c.sm_dim, s);
        copy_tokens<T>(lg_cls, 1, 0, G, nl, 0, 1, B, c.lg_dim, s);
      }
      classify<T>(S, ns, c.sm_dim, sm_head_norm, sm_head, B, 0, logits, false, s);
      classify<T>(G, nl, c.lg_dim, lg_head_norm, lg_head, B, 0, logits, true, s);
    }
  }

  // DistillMixin.call (distill.py:16-45) on top of a ViT: embed -> append the distillation token as the LAST row ->
  // transformer over n + 2 rows -> head on the first n + 1 rows, and the last row returned as is.
  template <typename T>
  void distill_impl(const float* img, int B, int H, int Wd, const float* distill_token, float* logits, float* distill_out, cudaStream_t s) {
    const vb_config& c = cfg;
    VB_CHECK(c.kind == VB_KIND_VIT, "vb_forward_distill supports ViT (distill.py:47 DistillableViT)");
    arena.reset();
    int rows = 0;
    T* E = embed_tokens<T>(embed, img, B, H, Wd, c.patch_h, c.patch_w, &rows, s, nullptr);
    const int rd = rows + 1;
    T* X = arena.get<T>(static_cast<size_t>(B) * rd * c.dim);
    copy_tokens<T>(E, rows, 0, X, rd, 0, rows, B, c.dim, s);
    T* last = X + static_cast<size_t>(rows) * c.dim;                 // row `rows` of image 0; image pitch rd rows
    broadcast_row<T>(distill_token, last, rd, B, c.dim, s);
    float* stats = (bf16() && c.dim % 64 == 0 && getenv("VB_NO_LN_FOLD") == nullptr) ? arena.get<float>(static_cast<size_t>(B) * rd * (c.dim / 64) * 2) : nullptr;
    bool sv = false;
    for (const auto& l : layers) layer_self<T>(X, B, rd, c.dim, l, s, stats, &sv);
    T* Xh = arena.get<T>(static_cast<size_t>(B) * rows * c.dim);     // x[:, :-1]
    copy_tokens<T>(X, rd, 0, Xh, rows, 0, rows, B, c.dim, s);
    classify<T>(Xh, rows, c.dim, head_norm, head, B, c.pool == VB_POOL_MEAN, logits, false, s);
    T* D = arena.get<T>(static_cast<size_t>(B) * c.dim);             // x[:, -1]
    copy_tokens<T>(X, rd, rows, D, 1, 0, 1, B, c.dim, s);
    const long long count = static_cast<long long>(B) * c.dim;
    if (sizeof(T) == 4) VB_CUDA(cudaMemcpyAsync(distill_out, D, count * 4, cudaMemcpyDeviceToDevice, s));
    else convert<__nv_bfloat16, float>(reinterpret_cast<const __nv_bfloat16*>(D), distill_out, count, s);
  }

  template <typename T>
  void tokens_impl(const float* tok, int B, int n, float* out, cudaStream_t s) {
    VB_CHECK(cfg.kind == VB_KIND_VIT || cfg.kind == VB_KIND_DEEPVIT || cfg.kind == VB_KIND_T2T_VIT,
             "vb_forward_tokens supports ViT / DeepViT / T2TViT");
    arena.reset();
    const long long count = static_cast<long long>(B) * n * cfg.dim;
    T* X;
    if (sizeof(T) == 4) {
      X = reinterpret_cast<T*>(arena.get<float>(count));
      VB_CUDA(cudaMemcpyAsync(X, tok, count * 4, cudaMemcpyDeviceToDevice, s));
    } else {
      X = arena.get<T>(count);
      convert<float, __nv_bfloat16>(tok, reinterpret_cast<__nv_bfloat16*>(X), count, s);
    }
    float* stats = (bf16() && cfg.dim % 64 == 0 && getenv("VB_NO_LN_FOLD") == nullptr) ? arena.get<float>(static_cast<size_t>(B) * n * (cfg.dim / 64) * 2) : nullptr;
    bool sv = false;
    for (const auto& l : layers) layer_self<T>(X, B, n, cfg.dim, l, s, stats, &sv);
    if (sizeof(T) == 4) VB_CUDA(cudaMemcpyAsync(out, X, count * 4, cudaMemcpyDeviceToDevice, s));
    else convert<__nv_bfloat16, float>(reinterpret_cast<const __nv_bfloat16*>(X), out, count, s);
  }

  // ---------------------------------------------------------------- stage entries (vb_forward_embed / _head / vb_patch_to_emb)
  int embed_rows(int H, int Wd) const {
    VB_CHECK(cfg.kind != VB_KIND_CROSSVIT, "CrossViT has two token streams: no single embedding stage");
    if (cfg.kind == VB_KIND_T2T_VIT) {
      int h = H, w = Wd;
      for (const auto& st : t2t_stages()) { h = (h + st.stride - 1) / st.stride; w = (w + st.stride - 1) / st.stride; }
      return h * w + 1;
    }
    VB_CHECK(H % cfg.patch_h == 0 && Wd % cfg.patch_w == 0, "Image dimensions must be divisible by the patch size.");
    const bool has_cls = cfg.kind != VB_KIND_CAIT && cfg.kind != VB_KIND_PATCH_MERGER_VIT;
    return (H / cfg.patch_h) * (Wd / cfg.patch_w) + (has_cls ? 1 : 0);
  }
  template <typename T>
  void to_f32(const T* X, float* out, long long count, cudaStream_t s) {
    if (sizeof(T) == 4) VB_CUDA(cudaMemcpyAsync(out, X, count * 4, cudaMemcpyDeviceToDevice, s));
    else convert<__nv_bfloat16, float>(reinterpret_cast<const __nv_bfloat16*>(X), out, count, s);
  }
  template <typename T>
  void embed_impl(const float* img, int B, int H, int Wd, float* tokens, cudaStream_t s) {
    arena.reset();
    int rows = 0;
    T* X = embed_any<T>(img, B, H, Wd, &rows, s, nullptr);
    to_f32<T>(X, tokens, static_cast<long long>(B) * rows * cfg.dim, s);
  }
  template <typename T>
  void head_impl(const float* tok, int B, int n, float* logits, cudaStream_t s) {
    VB_CHECK(cfg.kind != VB_KIND_CROSSVIT, "CrossViT sums two heads: no single mlp_head stage");
    arena.reset();
    const long long count = static_cast<long long>(B) * n * cfg.dim;
    T* X = arena.get<T>(count);
    convert_rows<float, T>(tok, cfg.dim, X, cfg.dim, static_cast<long long>(B) * n, cfg.dim, s);
    const bool mean = cfg.kind == VB_KIND_PATCH_MERGER_VIT || (cfg.kind != VB_KIND_CAIT && cfg.pool == VB_POOL_MEAN);
    classify<T>(X, n, cfg.dim, head_norm, head, B, mean ? 1 : 0, logits, false, s);
  }
  template <typename T>
  void patch_to_emb_impl(const float* patches, int rows, float* out, cudaStream_t s) {
    VB_CHECK(cfg.kind != VB_KIND_CROSSVIT, "CrossViT has two patch embeddings");
    arena.reset();
    const int K = embed.patch.K, Kp = bf16() ? embed.patch.ldw : K;
    T* col = arena.get<T>(static_cast<size_t>(rows) * Kp);
    convert_rows<float, T>(patches, K, col, Kp, rows, K, s);
    T* Y = arena.get<T>(static_cast<size_t>(rows) * cfg.dim);
    Epi ep; ep.bias = embed.patch.bias;
    Linear L = embed.patch;
    L.K = Kp;
    linear<T>(col, Kp, rows, L, Y, cfg.dim, ep, s);
    to_f32<T>(Y, out, static_cast<long long>(rows) * cfg.dim, s);
  }
};

// ------------------------------------------------------------------------------------------ op dispatch
template <>
void vb_handle::linear<float>(const float* A, int lda, int M, const Linear& L, float* out, int ldc, const Epi& e, cudaStream_t s) {
  gemm_simt<float, float, float>(A, lda, L.W, L.N, 1, out, ldc, M, L.N, L.K, e.bias, e.scale,
                                 static_cast<const float*>(e.res), e.ldr, e.gelu ? 1 : 0, s);
}
template <>
void vb_handle::linear<__nv_bfloat16>(const __nv_bfloat16* A, int lda, int M, const Linear& L, __nv_bfloat16* out, int ldc,
                                      const Epi& e, cudaStream_t s) {
  const int K = L.K;
  const __nv_bfloat16* res = static_cast<const __nv_bfloat16*>(e.res);
  const bool fast = gemm_bf16_supported(M, L.N, K, lda, L.ldw, ldc) && (res == nullptr || e.ldr % 8 == 0);
  const bool folded = L.ln_c1 != nullptr;
  VB_CHECK(!folded || (fast && e.ln_stats != nullptr && K % 64 == 0), "internal: LayerNorm-folded Dense needs the tcgen05 path and row statistics");
  VB_CHECK(e.stats_out == nullptr || fast, "internal: row statistics requested from a non-tcgen05 GEMM");
  ProfScope ps(this, !fast ? PROF_OTHER : e.gelu ? PROF_GEMM_GELU : res ? PROF_GEMM_RES : PROF_GEMM, 2.0 * M * L.N * K,
               2.0 * (static_cast<double>(M) * K + static_cast<double>(L.N) * K + static_cast<double>(M) * L.N * (res ? 2 : 1)), s);
  if (fast) {
    PlanKey key{};
    const uintptr_t parts[16] = {reinterpret_cast<uintptr_t>(A), static_cast<uintptr_t>(lda), reinterpret_cast<uintptr_t>(L.Wt),
                                 reinterpret_cast<uintptr_t>(out), static_cast<uintptr_t>(ldc), static_cast<uintptr_t>(M),
                                 static_cast<uintptr_t>(L.N), static_cast<uintptr_t>(K), reinterpret_cast<uintptr_t>(e.bias),
                                 reinterpret_cast<uintptr_t>(e.scale), reinterpret_cast<uintptr_t>(res),
                                 static_cast<uintptr_t>(e.ldr), static_cast<uintptr_t>(e.gelu),
                                 reinterpret_cast<uintptr_t>(e.ln_stats), reinterpret_cast<uintptr_t>(L.ln_c1),
                                 reinterpret_cast<uintptr_t>(e.stats_out)};
    for (int i = 0; i < 16; ++i) key[i] = parts[i];
    auto it = plans.find(key);
    if (it == plans.end()) {
      if (plans.size() > 8192) plans.clear();      // shape sweeps: bounded host memory (a plan is four 128-byte tensor maps)
      GemmBf16 g = gemm_bf16_plan(A, lda, L.Wt, L.ldw, out, ldc, M, L.N, K, e.bias, e.scale, res, e.ldr, e.gelu);
      if (folded) { g.ln_c1 = L.ln_c1; g.ln_stats = e.ln_stats; g.ln_parts = K / 64; g.ln_inv_d = 1.0f / static_cast<float>(K); }
      if (e.stats_out) { g.stats_out = e.stats_out; g.stats_parts = L.N / 64; }
      it = plans.emplace(key, g).first;
    }
    gemm_bf16_run(it->second, s);
  } else {
    gemm_simt<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(A, lda, L.Wt, 1, L.ldw, out, ldc, M, L.N, K, e.bias, e.scale, res,
                                                           e.ldr, e.gelu ? 1 : 0, s);
  }
}

template <>
void vb_handle::ensure_stats<float>(const float*, int, float*, int, cudaStream_t) {}
template <>
void vb_handle::ensure_stats<__nv_bfloat16>(const __nv_bfloat16* X, int dim, float* stats, int M, cudaStream_t s) {
  ProfScope ps(this, PROF_LN, 0.0, 2.0 * M * dim, s);
  row_stats_bf16(X, dim, stats, M, dim, s);
}

template <>
void vb_handle::layer_t2t<float>(float*, int, int, const LayerW&, cudaStream_t) { VB_CHECK(false, "internal: tensor-core T2T layer in the fp32 engine"); }
template <>
void vb_handle::layer_t2t<__nv_bfloat16>(__nv_bfloat16* X, int B, int n, const LayerW& l, cudaStream_t s) {
  using bf = __nv_bfloat16;
  const int D = l.t2t_D, Dp = l.t2t_Dp, M = B * n, npad = round_up(n, 64);
  bf* Y = arena.get<bf>(static_cast<size_t>(M) * Dp);
  bf* QKV = arena.get<bf>(static_cast<size_t>(M) * 3 * Dp);
  bf* Vt = arena.get<bf>(static_cast<size_t>(B) * Dp * npad);
  // Images are independent: ns of them are in flight on ns streams (the forward's own + side streams), each with its own score /
  // probability buffers.  The per-image kernels are small (16 pair tiles for Q K^T at n = 784) and strictly dependent, so one
  // stream leaves most SMs idle and pays every launch gap (389 launches, 6.1 of the 6.9 ms T2T step at batch 64).  Two streams
  // when one image's buffers are large (n = 3136: 59 MB, two of them still fit the 126 MB L2), four otherwise.
  const size_t s_elems = static_cast<size_t>(n) * npad;
  static const char* ns_env = getenv("VB_T2T_STREAMS");
  int ns = ns_env != nullptr ? atoi(ns_env) : (s_elems * 6 > (48u << 20) ? 2 : 4);
  ns = std::max(1, std::min(ns, std::min(B, static_cast<int>(T2T_MAX_STREAMS))));
  if (profiling) ns = 1;                                         // per-class event timing wants one stream
  float* S[T2T_MAX_STREAMS];
  bf* P[T2T_MAX_STREAMS];
  cudaStream_t st[T2T_MAX_STREAMS];
  for (int i = 0; i < ns; ++i) { S[i] = arena.get<float>(s_elems); P[i] = arena.get<bf>(s_elems); }
  st[0] = s;
  if (ns > 1) {
    ensure_side_streams();
    for (int i = 1; i < ns; ++i) st[i] = side_streams[i - 1];
  }
  { ProfScope ps(this, PROF_LN, 0.0, 4.0 * M * D, s); layernorm<bf>(X, Dp, l.attn_norm.gamma, l.attn_norm.beta, Y, Dp, M, D, s, Dp); }
  linear<bf>(Y, Dp, M, l.to_qkv, QKV, 3 * Dp, Epi(), s);
  {
    ProfScope ps(this, PROF_OTHER, 0.0, 4.0 * M * Dp, s);
    transpose_rows_bf16(QKV + 2 * Dp, 3 * Dp, static_cast<long long>(n) * 3 * Dp, Vt, npad, static_cast<long long>(Dp) * npad, B, n, npad, Dp, s);
  }
  const float scale_log2 = (1.0f / sqrtf(static_cast<float>(D))) * 1.4426950408889634f;
  if (ns > 1) {
    VB_CUDA(cudaEventRecord(fork_event, s));
    for (int i = 1; i < ns; ++i) VB_CUDA(cudaStreamWaitEvent(st[i], fork_event, 0));
  }
  for (int b = 0; b < B; ++b) {
    const int i = b % ns;
    const bf* Qb = QKV + static_cast<size_t>(b) * n * 3 * Dp;
    bf* Xb = X + static_cast<size_t>(b) * n * Dp;
    gemm_cached(Qb, 3 * Dp, Qb + Dp, 3 * Dp, n, S[i], npad, n, npad, Dp, nullptr, 0, true, PROF_ATTN, st[i]);       // S = Q K^T
    { ProfScope ps(this, PROF_ATTN, 0.0, 6.0 * n * npad, st[i]); softmax_rows_bf16(S[i], npad, P[i], npad, n, n, npad, scale_log2, st[i]); }
    gemm_cached(P[i], npad, Vt + static_cast<size_t>(b) * Dp * npad, npad, 0, Xb, Dp, n, Dp, npad, Xb, Dp, false, PROF_ATTN, st[i]);   // X += P V
  }
  for (int i = 1; i < ns; ++i) {
    VB_CUDA(cudaEventRecord(join_events[i - 1], st[i]));
    VB_CUDA(cudaStreamWaitEvent(s, join_events[i - 1], 0));
  }
  { ProfScope ps(this, PROF_LN, 0.0, 4.0 * M * D, s); layernorm<bf>(X, Dp, l.ff_norm.gamma, l.ff_norm.beta, Y, Dp, M, D, s, Dp); }
  bf* Hb = arena.get<bf>(static_cast<size_t>(M) * Dp);
  Epi e1; e1.gelu = true; e1.bias = l.fc1.bias;
  linear<bf>(Y, Dp, M, l.fc1, Hb, Dp, e1, s);
  Epi e2; e2.bias = l.fc2.bias; e2.res = X; e2.ldr = Dp;
  linear<bf>(Hb, Dp, M, l.fc2, X, Dp, e2, s);
}

template <typename T>
void vb_handle::attention_dispatch(const T* q, int ldq, const T* k, int ldk, const T* v, int ldv, T* out, int ldo, int B, int nq,
                                   int nk, int heads, int dh, int variant, const float* mix_a, const float* mix_b, const float* g,
                                   const float* b, cudaStream_t s, float scale) {
  ProfScope ps(this, PROF_ATTN, 4.0 * B * heads * nq * nk * dh + (variant == 1 ? 2.0 : variant == 2 ? 4.0 : 0.0) * B * nq * nk * heads * heads,
               static_cast<double>(sizeof(T)) * B * heads * dh * (2.0 * nq + 2.0 * nk), s);
  if (attention_fast<T>(q, ldq, k, ldk, v, ldv, out, ldo, B, nq, nk, heads, dh, variant, mix_a, mix_b, g, b, s, scale)) return;
  VB_CHECK(scale <= 0.f, "internal: a head-padded layer must run on the tcgen05 attention kernel");
  float* S = arena.get<float>(static_cast<size_t>(B) * heads * nq * ((nk + 15) & ~15));   // row pitch padded for the bf16-P path
  attention_generic<T>(q, ldq, k, ldk, v, ldv, out, ldo, S, B, nq, nk, heads, dh, variant, mix_a, mix_b, g, b, s);
}

template <>
void vb_handle::add_tokens<float>(float* X, const float* O, long long count, cudaStream_t s) { add_inplace_f32(X, O, count, s); }
template <>
void vb_handle::add_tokens<__nv_bfloat16>(__nv_bfloat16* X, const __nv_bfloat16* O, long long count, cudaStream_t s) {
  // rare path (heads == 1 and dim_head == dim): go through fp32 scratch
  float* a = arena.get<float>(count);
  float* b = arena.get<float>(count);
  convert<__nv_bfloat16, float>(X, a, count, s);
  convert<__nv_bfloat16, float>(O, b, count, s);
  add_inplace_f32(a, b, count, s);
  convert<float, __nv_bfloat16>(a, X, count, s);
}

// ------------------------------------------------------------------------------------------ C ABI
namespace {

template <typename F>
int guarded(vb_handle* h, F&& f) {
  try {
    f();
    return 0;
  } catch (const vb::Error& e) {
    (h ? h->error : g_last_error) = e.what();
    return e.code;
  } catch (const std::exception& e) {
    (h ? h->error : g_last_error) = e.what();
    return 3;
  } catch (...) {
    (h ? h->error : g_last_error) = "unknown error";
    return 4;
  }
}

void validate(const vb_config& c) {
  VB_CHECK(c.struct_size == static_cast<int32_t>(sizeof(vb_config)), "vb_config.struct_size mismatch (ABI)");
  VB_CHECK(c.kind >= VB_KIND_VIT && c.kind <= VB_KIND_T2T_VIT, "unknown model kind");
  VB_CHECK(c.kind != VB_KIND_PARALLEL_VIT || (c.parallel_branches >= 1 && c.parallel_branches <= 8), "num_parallel_branches must be in [1, 8]");
  VB_CHECK(c.precision == VB_PRECISION_FP32 || c.precision == VB_PRECISION_BF16, "unknown precision");
  VB_CHECK(c.channels > 0 && c.num_classes > 0 && c.image_h > 0 && c.image_w > 0, "bad image / class configuration");
  if (c.kind == VB_KIND_CROSSVIT) {
    VB_CHECK(c.image_h == c.image_w, "CrossViT takes a square integer image_size");
    VB_CHECK(c.sm_patch_size > 0 && c.lg_patch_size > 0 && c.image_h % c.sm_patch_size == 0 && c.image_h % c.lg_patch_size == 0,
             "Image dimensions must be divisible by the patch size.");
    VB_CHECK(c.sm_dim > 0 && c.lg_dim > 0 && c.cross_depth > 0 && c.cross_attn_depth >= 0, "bad CrossViT dimensions");
    VB_CHECK(c.sm_enc_heads <= 32 && c.lg_enc_heads <= 32 && c.cross_attn_heads <= 32, "at most 32 heads");
  } else {
    if (c.kind == VB_KIND_T2T_VIT) {
      VB_CHECK(c.image_h == c.image_w, "T2TViT takes a square integer image_size");
      VB_CHECK(c.t2t_num_layers >= 1 && c.t2t_num_layers <= 4, "t2t_layers: between 1 and 4 (kernel_size, stride) pairs");
      const int ks[4] = {c.t2t_k0, c.t2t_k1, c.t2t_k2, c.t2t_k3}, ss[4] = {c.t2t_s0, c.t2t_s1, c.t2t_s2, c.t2t_s3};
      long long d = c.channels;
      for (int i = 0; i < c.t2t_num_layers; ++i) {
        VB_CHECK(ks[i] > 0 && ss[i] > 0, "t2t_layers: kernel sizes and strides must be positive");
        d *= static_cast<long long>(ks[i]) * ks[i];
        VB_CHECK(d <= (1 << 20), "t2t_layers: token width channels * prod(kernel_size^2) is too large");
      }
    } else {
      VB_CHECK(c.patch_h > 0 && c.patch_w > 0 && c.image_h % c.patch_h == 0 && c.image_w % c.patch_w == 0,
               "Image dimensions must be divisible by the patch size.");
    }
    VB_CHECK(c.kind != VB_KIND_PATCH_MERGER_VIT || c.patch_merge_num_tokens > 0, "patch_merge_num_tokens must be positive");
    VB_CHECK(c.dim > 0 && c.depth >= 0 && c.heads > 0 && c.dim_head > 0 && c.mlp_dim > 0, "bad transformer dimensions");
    VB_CHECK(c.heads <= 32, "at most 32 heads");
    VB_CHECK(c.pool == VB_POOL_CLS || c.pool == VB_POOL_MEAN, "pool type must be either cls (cls token) or mean (mean pooling)");
  }
}

}  // namespace

// ---------------------------------------------------------------- single-operator entry points
namespace {
struct Timer {
  cudaEvent_t a, b;
  Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};
template <typename F>
void timed(int iters, float* elapsed_ms, F&& f) {
  f();  // first run produces the result (and warms up)
  VB_CUDA(cudaDeviceSynchronize());
  if (elapsed_ms && iters > 0) {
    Timer t;
    VB_CUDA(cudaEventRecord(t.a, 0));
    for (int i = 0; i < iters; ++i) f();
    VB_CUDA(cudaEventRecord(t.b, 0));
    VB_CUDA(cudaEventSynchronize(t.b));
    float ms = 0.f;
    VB_CUDA(cudaEventElapsedTime(&ms, t.a, t.b));
    *elapsed_ms = ms / iters;
  }
}
void require_gpu() {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  VB_CHECK(e == cudaSuccess && ndev > 0, "no CUDA device available -- libvitb200 has no CPU fallback");
}
template <typename T>
T* upload(DevMem& m, const float* host, size_t count) {
  m.ensure(count * sizeof(T) + 16);
  if (sizeof(T) == 4) {
    VB_CUDA(cudaMemcpy(m.p, host, count * 4, cudaMemcpyHostToDevice));
  } else {
    DevMem tmp;
    tmp.ensure(count * 4);
    VB_CUDA(cudaMemcpy(tmp.p, host, count * 4, cudaMemcpyHostToDevice));
    convert<float, __nv_bfloat16>(static_cast<const float*>(tmp.p), static_cast<__nv_bfloat16*>(m.p), static_cast<long long>(count), 0);
    VB_CUDA(cudaDeviceSynchronize());
  }
  return static_cast<T*>(m.p);
}
template <typename T>
void download(const T* dev, float* host, size_t count) {
  if (sizeof(T) == 4) {
    VB_CUDA(cudaMemcpy(host, dev, count * 4, cudaMemcpyDeviceToHost));
  } else {
    DevMem tmp;
    tmp.ensure(count * 4);
    convert<__nv_bfloat16, float>(reinterpret_cast<const __nv_bfloat16*>(dev), static_cast<float*>(tmp.p), static_cast<long long>(count), 0);
    VB_CUDA(cudaMemcpy(host, tmp.p, count * 4, cudaMemcpyDeviceToHost));
  }
}
}  // namespace

namespace {
// Shared plumbing of the stage entries: optional host->device staging of the input, device staging of a host output,
// launch accounting, and the final copy + synchronise when the output is a host buffer.
template <typename F>
void staged_call(vb_handle* h, const float* in, int32_t in_mem, size_t in_bytes, float* out, int32_t out_mem, size_t out_bytes,
                 cudaStream_t s, F&& body) {
  VB_CUDA(cudaSetDevice(h->device));
  const long long before = launch_counter();
  const float* in_d = in;
  if (in_mem == VB_MEM_HOST) {
    h->tokens_in.ensure(in_bytes);
    VB_CUDA(cudaMemcpyAsync(h->tokens_in.p, in, in_bytes, cudaMemcpyHostToDevice, s));
    in_d = static_cast<const float*>(h->tokens_in.p);
  }
  float* out_d = out;
  if (out_mem == VB_MEM_HOST) {
    h->tokens_out.ensure(out_bytes);
    out_d = static_cast<float*>(h->tokens_out.p);
  }
  body(in_d, out_d);
  h->last_launches = launch_counter() - before;
  if (out_mem == VB_MEM_HOST) {
    VB_CUDA(cudaMemcpyAsync(out, out_d, out_bytes, cudaMemcpyDeviceToHost, s));
    VB_CUDA(cudaStreamSynchronize(s));
  }
}
}  // namespace

extern "C" {

int vb_abi_version(void) { return VB_ABI_VERSION; }

int vb_create(const vb_config* cfg, int device, vb_handle** out) {
  return guarded(nullptr, [&] {
    VB_CHECK(cfg != nullptr && out != nullptr, "vb_create: null argument");
    validate(*cfg);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    VB_CHECK(e == cudaSuccess && ndev > 0, "vb_create: no CUDA device available -- libvitb200 has no CPU fallback");
    VB_CHECK(device >= 0 && device < ndev, "vb_create: bad device index");
    VB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VB_CUDA(cudaGetDeviceProperties(&prop, device));
    VB_CHECK(prop.major == 10, "vb_create: libvitb200 is built for sm_100a (Blackwell B200) only");
    std::unique_ptr<vb_handle> h(new vb_handle());
    h->cfg = *cfg;
    h->device = device;
    h->build_expected();
    *out = h.release();
  });
}

int vb_num_weights(vb_handle* h) { return h ? static_cast<int>(h->weights.size()) : -1; }

int vb_weight_info(vb_handle* h, int32_t index, const char** name, int64_t* shape4, int32_t* ndim) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr, "null handle");
    VB_CHECK(index >= 0 && index < static_cast<int>(h->weights.size()), "weight index out of range");
    const Weight& w = h->weights[index];
    if (name) *name = w.name.c_str();
    if (ndim) *ndim = static_cast<int32_t>(w.shape.size());
    if (shape4) for (size_t i = 0; i < w.shape.size() && i < 4; ++i) shape4[i] = w.shape[i];
  });
}

int vb_set_weight(vb_handle* h, const char* name, const float* host_data, const int64_t* shape, int32_t ndim) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && name != nullptr && host_data != nullptr && shape != nullptr, "vb_set_weight: null argument");
    auto it = h->windex.find(name);
    VB_CHECK(it != h->windex.end(), std::string("vb_set_weight: this model has no weight named '") + name + "'");
    Weight& w = h->weights[it->second];
    bool ok = static_cast<size_t>(ndim) == w.shape.size();
    for (int i = 0; ok && i < ndim; ++i) ok = shape[i] == w.shape[i];
    VB_CHECK(ok, std::string("vb_set_weight: shape mismatch for '") + name + "'");
    VB_CUDA(cudaSetDevice(h->device));
    if (w.dev == nullptr) VB_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.dev), w.count * sizeof(float)));
    VB_CUDA(cudaMemcpy(w.dev, host_data, w.count * sizeof(float), cudaMemcpyHostToDevice));
    w.set = true;
    h->finalized = false;
  });
}

int vb_finalize(vb_handle* h) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr, "null handle");
    attention_mix_cache_clear();
    h->finalize();
  });
}

int vb_forward(vb_handle* h, const float* img, int32_t img_mem, int32_t batch, int32_t img_h, int32_t img_w, float* logits,
               int32_t logits_mem, void* stream) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && img != nullptr && logits != nullptr, "vb_forward: null argument");
    VB_CHECK(h->finalized, "vb_forward: call vb_finalize after setting the weights");
    VB_CHECK(batch > 0 && img_h > 0 && img_w > 0, "vb_forward: bad batch / image size");
    VB_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long before = launch_counter();
    const size_t img_bytes = static_cast<size_t>(batch) * img_h * img_w * h->cfg.channels * sizeof(float);
    const size_t out_bytes = static_cast<size_t>(batch) * h->cfg.num_classes * sizeof(float);
    const float* img_d = img;
    if (img_mem == VB_MEM_HOST) {
      h->img_dev.ensure(img_bytes);
      VB_CUDA(cudaMemcpyAsync(h->img_dev.p, img, img_bytes, cudaMemcpyHostToDevice, s));
      img_d = static_cast<const float*>(h->img_dev.p);
    }
    float* out_d = logits;
    if (logits_mem == VB_MEM_HOST) {
      h->logits_dev.ensure(out_bytes);
      out_d = static_cast<float*>(h->logits_dev.p);
    }
    auto run_eager = [&] {
      if (h->bf16()) h->forward_impl<__nv_bfloat16>(img_d, batch, img_h, img_w, out_d, s);
      else h->forward_impl<float>(img_d, batch, img_h, img_w, out_d, s);
    };
    static const bool graphs_off = getenv("VB_NO_GRAPH") != nullptr;
    const bool graphable = !graphs_off && !h->profiling && s != nullptr && s != cudaStreamLegacy;
    bool done = false;
    if (graphable) {
      if (h->graphs.size() > 64) h->drop_graphs();                         // shape / pointer sweeps: bounded
      vb_handle::GraphEntry& ge = h->graphs[vb_handle::GraphKey{img_d, out_d, batch, img_h, img_w, s}];
      ++ge.calls;
      if (ge.exec != nullptr) {
        VB_CUDA(cudaGraphLaunch(ge.exec, s));
        count_launch(static_cast<int>(ge.launches));
        done = true;
      } else if (ge.calls == 2 && !ge.failed) {
        cudaGraph_t graph = nullptr;
        const long long l0 = launch_counter();
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          bool ok = true;
          std::string why;
          try { run_eager(); } catch (const std::exception& e) { ok = false; why = e.what(); }
          const cudaError_t ec = cudaStreamEndCapture(s, &graph);
          if (ok && ec == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&ge.exec, graph, 0) == cudaSuccess) {
            ge.launches = launch_counter() - l0;
            VB_CUDA(cudaGraphLaunch(ge.exec, s));
            done = true;
          } else {
            ge.failed = true;                                               // stay eager for this key
            ge.exec = nullptr;
            cudaGetLastError();
          }
          if (graph != nullptr) cudaGraphDestroy(graph);
        } else {
          ge.failed = true;
          cudaGetLastError();
        }
      }
    }
    if (!done) run_eager();
    h->last_launches = launch_counter() - before;
    if (logits_mem == VB_MEM_HOST) {
      VB_CUDA(cudaMemcpyAsync(logits, out_d, out_bytes, cudaMemcpyDeviceToHost, s));
      VB_CUDA(cudaStreamSynchronize(s));
    }
  });
}

int vb_forward_distill(vb_handle* h, const float* img, int32_t img_mem, int32_t batch, int32_t img_h, int32_t img_w,
                       const float* distill_token, float* logits, float* distill_out, int32_t out_mem, void* stream) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && img != nullptr && distill_token != nullptr && logits != nullptr && distill_out != nullptr,
             "vb_forward_distill: null argument");
    VB_CHECK(h->finalized, "vb_forward_distill: call vb_finalize after setting the weights");
    VB_CHECK(batch > 0 && img_h > 0 && img_w > 0, "vb_forward_distill: bad batch / image size");
    VB_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long before = launch_counter();
    const int dim = h->cfg.dim;
    const size_t img_bytes = static_cast<size_t>(batch) * img_h * img_w * h->cfg.channels * sizeof(float);
    const size_t log_bytes = static_cast<size_t>(batch) * h->cfg.num_classes * sizeof(float);
    const size_t dis_bytes = static_cast<size_t>(batch) * dim * sizeof(float);
    const float* img_d = img;
    if (img_mem == VB_MEM_HOST) {
      h->img_dev.ensure(img_bytes);
      VB_CUDA(cudaMemcpyAsync(h->img_dev.p, img, img_bytes, cudaMemcpyHostToDevice, s));
      img_d = static_cast<const float*>(h->img_dev.p);
    }
    // the distillation token is always a host vector of `dim` floats (a trainable variable of the caller, distill.py:133)
    h->tokens_in.ensure(static_cast<size_t>(dim) * sizeof(float));
    VB_CUDA(cudaMemcpyAsync(h->tokens_in.p, distill_token, static_cast<size_t>(dim) * sizeof(float), cudaMemcpyHostToDevice, s));
    float* log_d = logits;
    float* dis_d = distill_out;
    if (out_mem == VB_MEM_HOST) {
      h->logits_dev.ensure(log_bytes);
      h->tokens_out.ensure(dis_bytes);
      log_d = static_cast<float*>(h->logits_dev.p);
      dis_d = static_cast<float*>(h->tokens_out.p);
    }
    const float* tok_d = static_cast<const float*>(h->tokens_in.p);
    if (h->bf16()) h->distill_impl<__nv_bfloat16>(img_d, batch, img_h, img_w, tok_d, log_d, dis_d, s);
    else h->distill_impl<float>(img_d, batch, img_h, img_w, tok_d, log_d, dis_d, s);
    h->last_launches = launch_counter() - before;
    if (out_mem == VB_MEM_HOST) {
      VB_CUDA(cudaMemcpyAsync(logits, log_d, log_bytes, cudaMemcpyDeviceToHost, s));
      VB_CUDA(cudaMemcpyAsync(distill_out, dis_d, dis_bytes, cudaMemcpyDeviceToHost, s));
      VB_CUDA(cudaStreamSynchronize(s));
    }
  });
}

int vb_forward_tokens(vb_handle* h, const float* tokens, int32_t tokens_mem, int32_t batch, int32_t n, float* out,
                      int32_t out_mem, void* stream) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && tokens != nullptr && out != nullptr, "vb_forward_tokens: null argument");
    VB_CHECK(h->finalized, "vb_forward_tokens: call vb_finalize after setting the weights");
    VB_CHECK(batch > 0 && n > 0, "vb_forward_tokens: bad shape");
    VB_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long before = launch_counter();
    const size_t bytes = static_cast<size_t>(batch) * n * h->cfg.dim * sizeof(float);
    const float* in_d = tokens;
    float* out_d = out;
    if (tokens_mem == VB_MEM_HOST || out_mem == VB_MEM_HOST) h->tok_dev.ensure(2 * bytes);
    if (tokens_mem == VB_MEM_HOST) {
      VB_CUDA(cudaMemcpyAsync(h->tok_dev.p, tokens, bytes, cudaMemcpyHostToDevice, s));
      in_d = static_cast<const float*>(h->tok_dev.p);
    }
    if (out_mem == VB_MEM_HOST) out_d = reinterpret_cast<float*>(static_cast<char*>(h->tok_dev.p) + bytes);
    if (h->bf16()) h->tokens_impl<__nv_bfloat16>(in_d, batch, n, out_d, s);
    else h->tokens_impl<float>(in_d, batch, n, out_d, s);
    h->last_launches = launch_counter() - before;
    if (out_mem == VB_MEM_HOST) {
      VB_CUDA(cudaMemcpyAsync(out, out_d, bytes, cudaMemcpyDeviceToHost, s));
      VB_CUDA(cudaStreamSynchronize(s));
    }
  });
}

int vb_embed_rows(vb_handle* h, int32_t img_h, int32_t img_w) {
  int rows = -1;
  const int rc = guarded(h, [&] {
    VB_CHECK(h != nullptr && img_h > 0 && img_w > 0, "vb_embed_rows: bad arguments");
    rows = h->embed_rows(img_h, img_w);
  });
  return rc == 0 ? rows : -rc;
}

int vb_forward_embed(vb_handle* h, const float* img, int32_t img_mem, int32_t batch, int32_t img_h, int32_t img_w, float* tokens,
                     int32_t tokens_mem, void* stream) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && img != nullptr && tokens != nullptr, "vb_forward_embed: null argument");
    VB_CHECK(h->finalized, "vb_forward_embed: call vb_finalize after setting the weights");
    VB_CHECK(batch > 0 && img_h > 0 && img_w > 0, "vb_forward_embed: bad batch / image size");
    const int rows = h->embed_rows(img_h, img_w);
    const size_t in_bytes = static_cast<size_t>(batch) * img_h * img_w * h->cfg.channels * sizeof(float);
    const size_t out_bytes = static_cast<size_t>(batch) * rows * h->cfg.dim * sizeof(float);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    staged_call(h, img, img_mem, in_bytes, tokens, tokens_mem, out_bytes, s, [&](const float* in_d, float* out_d) {
      if (h->bf16()) h->embed_impl<__nv_bfloat16>(in_d, batch, img_h, img_w, out_d, s);
      else h->embed_impl<float>(in_d, batch, img_h, img_w, out_d, s);
    });
  });
}

int vb_forward_head(vb_handle* h, const float* tokens, int32_t tokens_mem, int32_t batch, int32_t n, float* logits,
                    int32_t logits_mem, void* stream) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && tokens != nullptr && logits != nullptr, "vb_forward_head: null argument");
    VB_CHECK(h->finalized, "vb_forward_head: call vb_finalize after setting the weights");
    VB_CHECK(batch > 0 && n > 0, "vb_forward_head: bad shape");
    const size_t in_bytes = static_cast<size_t>(batch) * n * h->cfg.dim * sizeof(float);
    const size_t out_bytes = static_cast<size_t>(batch) * h->cfg.num_classes * sizeof(float);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    staged_call(h, tokens, tokens_mem, in_bytes, logits, logits_mem, out_bytes, s, [&](const float* in_d, float* out_d) {
      if (h->bf16()) h->head_impl<__nv_bfloat16>(in_d, batch, n, out_d, s);
      else h->head_impl<float>(in_d, batch, n, out_d, s);
    });
  });
}

int vb_to_patch(vb_handle* h, const float* img, int32_t img_mem, int32_t batch, int32_t img_h, int32_t img_w, float* patches,
                int32_t patches_mem, void* stream) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && img != nullptr && patches != nullptr, "vb_to_patch: null argument");
    VB_CHECK(h->cfg.kind != VB_KIND_CROSSVIT && h->cfg.kind != VB_KIND_T2T_VIT, "vb_to_patch: the model has no single Rearrange patch layer");
    VB_CHECK(batch > 0 && img_h > 0 && img_w > 0, "vb_to_patch: bad batch / image size");
    const vb_config& c = h->cfg;
    VB_CHECK(img_h % c.patch_h == 0 && img_w % c.patch_w == 0, "Image dimensions must be divisible by the patch size.");
    const int np = (img_h / c.patch_h) * (img_w / c.patch_w), pd = c.patch_h * c.patch_w * c.channels;
    const size_t in_bytes = static_cast<size_t>(batch) * img_h * img_w * c.channels * sizeof(float);
    const size_t out_bytes = static_cast<size_t>(batch) * np * pd * sizeof(float);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    staged_call(h, img, img_mem, in_bytes, patches, patches_mem, out_bytes, s, [&](const float* in_d, float* out_d) {
      im2col<float>(in_d, out_d, batch, img_h, img_w, c.channels, c.patch_h, c.patch_w, 0, pd, s);
    });
  });
}

int vb_patch_to_emb(vb_handle* h, const float* patches, int32_t patches_mem, int32_t rows, float* out, int32_t out_mem, void* stream) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && patches != nullptr && out != nullptr, "vb_patch_to_emb: null argument");
    VB_CHECK(h->finalized, "vb_patch_to_emb: call vb_finalize after setting the weights");
    VB_CHECK(h->cfg.kind != VB_KIND_CROSSVIT, "vb_patch_to_emb: CrossViT has two patch embeddings");
    VB_CHECK(rows > 0, "vb_patch_to_emb: bad shape");
    const size_t in_bytes = static_cast<size_t>(rows) * h->embed.patch.K * sizeof(float);
    const size_t out_bytes = static_cast<size_t>(rows) * h->cfg.dim * sizeof(float);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    staged_call(h, patches, patches_mem, in_bytes, out, out_mem, out_bytes, s, [&](const float* in_d, float* out_d) {
      if (h->bf16()) h->patch_to_emb_impl<__nv_bfloat16>(in_d, rows, out_d, s);
      else h->patch_to_emb_impl<float>(in_d, rows, out_d, s);
    });
  });
}

int vb_dp_unique_id(void* id128) {
  return guarded(nullptr, [&] {
    VB_CHECK(id128 != nullptr, "vb_dp_unique_id: null argument");
    NcclId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id128, id.internal, sizeof id.internal);
  });
}

int vb_dp_init(vb_handle* h, const void* id128, int32_t rank, int32_t world) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && id128 != nullptr, "vb_dp_init: null argument");
    VB_CHECK(world >= 1 && rank >= 0 && rank < world, "vb_dp_init: need 0 <= rank < world");
    VB_CHECK(h->dp_comm == nullptr, "vb_dp_init: the handle already belongs to a data-parallel group");
    VB_CUDA(cudaSetDevice(h->device));
    NcclId id;
    memcpy(id.internal, id128, sizeof id.internal);
    void* comm = nullptr;
    nccl_check(nccl().CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
    h->dp_comm = comm; h->dp_rank = rank; h->dp_world = world;
  });
}

int vb_forward_allgather(vb_handle* h, const float* img, int32_t img_mem, int32_t local_batch, int32_t img_h, int32_t img_w,
                         float* gathered, void* stream) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr && img != nullptr && gathered != nullptr, "vb_forward_allgather: null argument");
    VB_CHECK(h->dp_comm != nullptr, "vb_forward_allgather: call vb_dp_init first");
    const size_t count = static_cast<size_t>(local_batch) * h->cfg.num_classes;
    float* mine = gathered + static_cast<size_t>(h->dp_rank) * count;        // this rank's slice of the gather buffer
    const int rc = vb_forward(h, img, img_mem, local_batch, img_h, img_w, mine, VB_MEM_DEVICE, stream);
    if (rc != 0) throw vb::Error(rc, h->error);
    // in place (sendbuff = recvbuff + rank * count), same stream: the head kernel's logits feed the collective directly
    nccl_check(nccl().AllGather(mine, gathered, count, /*ncclFloat32*/ 7, h->dp_comm, static_cast<cudaStream_t>(stream)),
               "ncclAllGather");
  });
}

int64_t vb_last_launch_count(vb_handle* h) { return h ? h->last_launches : -1; }

int vb_profile_enable(vb_handle* h, int32_t on) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr, "null handle");
    VB_CUDA(cudaSetDevice(h->device));
    h->prof_collect();
    h->profiling = on != 0;
  });
}

int vb_profile_read(vb_handle* h, double* ms, double* flops, double* bytes, int64_t* calls, int32_t reset) {
  return guarded(h, [&] {
    VB_CHECK(h != nullptr, "null handle");
    VB_CUDA(cudaSetDevice(h->device));
    h->prof_collect();
    for (int i = 0; i < vb_handle::PROF_NUM; ++i) {
      if (ms) ms[i] = h->prof_ms[i];
      if (flops) flops[i] = h->prof_flops[i];
      if (bytes) bytes[i] = h->prof_bytes[i];
      if (calls) calls[i] = h->prof_calls[i];
      if (reset) { h->prof_ms[i] = h->prof_flops[i] = h->prof_bytes[i] = 0; h->prof_calls[i] = 0; }
    }
  });
}

const char* vb_last_error(vb_handle* h) { return h ? h->error.c_str() : g_last_error.c_str(); }

void vb_destroy(vb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->dp_comm != nullptr) { nccl().CommDestroy(h->dp_comm); h->dp_comm = nullptr; }
  attention_mix_cache_clear();
  h->drop_graphs();
  for (auto& w : h->weights) if (w.dev) cudaFree(w.dev);
  h->prof_collect();
  for (auto e : h->event_pool) cudaEventDestroy(e);
  h->destroy_side_streams();
  delete h;
}


int vb_op_linear(int32_t precision, const float* a, const float* w, const float* bias, const float* scale, const float* res,
                 int32_t gelu, float* out, int32_t M, int32_t N, int32_t K, int32_t iters, float* elapsed_ms) {
  return guarded(nullptr, [&] {
    require_gpu();
    VB_CHECK(a && w && out && M > 0 && N > 0 && K > 0, "vb_op_linear: bad arguments");
    DevMem dA, dW, dWt, dB, dS, dR, dO;
    const float* db = bias ? upload<float>(dB, bias, N) : nullptr;
    const float* ds = scale ? upload<float>(dS, scale, N) : nullptr;
    const float* dw = upload<float>(dW, w, static_cast<size_t>(K) * N);
    if (precision == VB_PRECISION_FP32) {
      const float* da = upload<float>(dA, a, static_cast<size_t>(M) * K);
      const float* dr = res ? upload<float>(dR, res, static_cast<size_t>(M) * N) : nullptr;
      dO.ensure(static_cast<size_t>(M) * N * 4);
      float* dout = static_cast<float*>(dO.p);
      timed(iters, elapsed_ms, [&] { gemm_simt<float, float, float>(da, K, dw, N, 1, dout, N, M, N, K, db, ds, dr, N, gelu, 0); });
      download<float>(dout, out, static_cast<size_t>(M) * N);
    } else {
      VB_CHECK(gemm_bf16_supported(M, N, K, K, K, N), "vb_op_linear(bf16): need N % 64 == 0 and K % 8 == 0");
      const __nv_bfloat16* da = upload<__nv_bfloat16>(dA, a, static_cast<size_t>(M) * K);
      const __nv_bfloat16* dr = res ? upload<__nv_bfloat16>(dR, res, static_cast<size_t>(M) * N) : nullptr;
      dWt.ensure(static_cast<size_t>(N) * K * 2);
      pack_weight_bf16(dw, static_cast<__nv_bfloat16*>(dWt.p), K, N, K, 0);
      dO.ensure(static_cast<size_t>(M) * N * 2);
      __nv_bfloat16* dout = static_cast<__nv_bfloat16*>(dO.p);
      GemmBf16 g = gemm_bf16_plan(da, K, static_cast<const __nv_bfloat16*>(dWt.p), K, dout, N, M, N, K, db, ds, dr, N, gelu != 0);
      if (const char* trace_path = getenv("VB_GEMM_TRACE")) {   // one traced launch, dumped as text: role tag clock
        DevMem dTrace;
        dTrace.ensure(4 * 512 * 8);
        VB_CUDA(cudaMemset(dTrace.p, 0, 4 * 512 * 8));
        gemm_trace_buffer() = static_cast<long long*>(dTrace.p);
        gemm_bf16_run(g, 0);
        VB_CUDA(cudaDeviceSynchronize());
        gemm_trace_buffer() = nullptr;
        std::vector<long long> h(4 * 512);
        VB_CUDA(cudaMemcpy(h.data(), dTrace.p, h.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(trace_path, "w")) {
          for (int r = 0; r < 4; ++r)
            for (int i = 0; i < 250 && h[r * 512 + 2 * i + 1] != 0; ++i) fprintf(f, "%d %lld %lld\n", r, h[r * 512 + 2 * i], h[r * 512 + 2 * i + 1]);
          fclose(f);
        }
      }
      timed(iters, elapsed_ms, [&] { gemm_bf16_run(g, 0); });
      download<__nv_bfloat16>(dout, out, static_cast<size_t>(M) * N);
    }
  });
}

int vb_op_attention(int32_t precision, int32_t variant, const float* q, const float* k, const float* v, const float* mix_a,
                    const float* mix_b, const float* ln_gamma, const float* ln_beta, float* out, int32_t B, int32_t nq, int32_t nk,
                    int32_t heads, int32_t dim_head, int32_t iters, float* elapsed_ms) {
  return guarded(nullptr, [&] {
    require_gpu();
    VB_CHECK(q && k && v && out && B > 0 && nq > 0 && nk > 0 && heads > 0 && dim_head > 0, "vb_op_attention: bad arguments");
    VB_CHECK(variant >= 0 && variant <= 2, "vb_op_attention: variant must be 0, 1 or 2");
    attention_mix_cache_clear();
    const int inner = heads * dim_head;
    DevMem dQ, dK, dV, dO, dS, dMa, dMb, dG, dBt;
    const float* ma = mix_a ? upload<float>(dMa, mix_a, heads * heads) : nullptr;
    const float* mb = mix_b ? upload<float>(dMb, mix_b, heads * heads) : nullptr;
    const float* g = ln_gamma ? upload<float>(dG, ln_gamma, heads) : nullptr;
    const float* bt = ln_beta ? upload<float>(dBt, ln_beta, heads) : nullptr;
    const size_t cq = static_cast<size_t>(B) * nq * inner, ck = static_cast<size_t>(B) * nk * inner;
    auto run = [&](auto tag) {
      using T = decltype(tag);
      const T* q_d = upload<T>(dQ, q, cq);
      const T* k_d = upload<T>(dK, k, ck);
      const T* v_d = upload<T>(dV, v, ck);
      dO.ensure(cq * sizeof(T));
      T* o_d = static_cast<T*>(dO.p);
      dS.ensure(static_cast<size_t>(B) * heads * nq * ((nk + 15) & ~15) * 4);
      DevMem dTrace;
      const char* trace_path = getenv("VB_ATTN_TRACE");
      if (trace_path) { dTrace.ensure(4 * 512 * 8); VB_CUDA(cudaMemset(dTrace.p, 0, 4 * 512 * 8)); attn_trace_buffer() = static_cast<long long*>(dTrace.p); }
      struct TraceOff { ~TraceOff() { attn_trace_buffer() = nullptr; } } trace_off;
      if (trace_path) {   // one traced launch, dumped as text: role tag clock
        attention_fast<T>(q_d, inner, k_d, inner, v_d, inner, o_d, inner, B, nq, nk, heads, dim_head, variant, ma, mb, g, bt, 0);
        VB_CUDA(cudaDeviceSynchronize());
        std::vector<long long> h(4 * 512);
        VB_CUDA(cudaMemcpy(h.data(), dTrace.p, h.size() * 8, cudaMemcpyDeviceToHost));
        attn_trace_buffer() = nullptr;
        if (FILE* f = fopen(trace_path, "w")) {
          for (int r = 0; r < 4; ++r)
            for (int i = 0; i < 250 && h[r * 512 + 2 * i + 1] != 0; ++i) fprintf(f, "%d %lld %lld\n", r, h[r * 512 + 2 * i], h[r * 512 + 2 * i + 1]);
          fclose(f);
        }
      }
      timed(iters, elapsed_ms, [&] {
        if (!attention_fast<T>(q_d, inner, k_d, inner, v_d, inner, o_d, inner, B, nq, nk, heads, dim_head, variant, ma, mb, g, bt, 0))
          attention_generic<T>(q_d, inner, k_d, inner, v_d, inner, o_d, inner, static_cast<float*>(dS.p), B, nq, nk, heads,
                               dim_head, variant, ma, mb, g, bt, 0);
      });
      download<T>(o_d, out, cq);
    };
    if (precision == VB_PRECISION_FP32) run(float());
    else run(__nv_bfloat16());
  });
}

int vb_op_patch_merger(int32_t precision, const float* x, const float* gamma, const float* beta, const float* queries, float* out,
                       int32_t B, int32_t n, int32_t D, int32_t nt, int32_t iters, float* elapsed_ms) {
  return guarded(nullptr, [&] {
    require_gpu();
    VB_CHECK(x && gamma && beta && queries && out && B > 0 && n > 0 && D > 0 && nt > 0, "vb_op_patch_merger: bad arguments");
    DevMem dX, dG, dB, dQf, dY, dQ, dO, dS;
    const float* g = upload<float>(dG, gamma, D);
    const float* b = upload<float>(dB, beta, D);
    const float* qf = upload<float>(dQf, queries, static_cast<size_t>(nt) * D);
    const size_t cx = static_cast<size_t>(B) * n * D, co = static_cast<size_t>(B) * nt * D;
    auto run = [&](auto tag) {
      using T = decltype(tag);
      const T* x_d = upload<T>(dX, x, cx);
      dY.ensure(cx * sizeof(T) + 16); dQ.ensure(co * sizeof(T) + 16); dO.ensure(co * sizeof(T) + 16);
      dS.ensure(static_cast<size_t>(B) * nt * ((n + 15) & ~15) * 4);
      T* y_d = static_cast<T*>(dY.p);
      T* q_d = static_cast<T*>(dQ.p);
      T* o_d = static_cast<T*>(dO.p);
      timed(iters, elapsed_ms, [&] {
        layernorm<T>(x_d, D, g, b, y_d, D, B * n, D, 0);
        broadcast_rows<T>(qf, q_d, B, nt, D, 0);
        if (!attention_fast<T>(q_d, D, y_d, D, y_d, D, o_d, D, B, nt, n, 1, D, 0, nullptr, nullptr, nullptr, nullptr, 0))
          attention_generic<T>(q_d, D, y_d, D, y_d, D, o_d, D, static_cast<float*>(dS.p), B, nt, n, 1, D, 0, nullptr, nullptr, nullptr,
                               nullptr, 0);
      });
      download<T>(o_d, out, co);
    };
    if (precision == VB_PRECISION_FP32) run(float());
    else run(__nv_bfloat16());
  });
}

int vb_op_layernorm(int32_t precision, const float* x, const float* gamma, const float* beta, float* out, int32_t M, int32_t D,
                    int32_t iters, float* elapsed_ms) {
  return guarded(nullptr, [&] {
    require_gpu();
    VB_CHECK(x && gamma && beta && out && M > 0 && D > 0, "vb_op_layernorm: bad arguments");
    DevMem dX, dG, dB, dO;
    const float* g = upload<float>(dG, gamma, D);
    const float* b = upload<float>(dB, beta, D);
    const size_t cnt = static_cast<size_t>(M) * D;
    auto run = [&](auto tag) {
      using T = decltype(tag);
      const T* x_d = upload<T>(dX, x, cnt);
      dO.ensure(cnt * sizeof(T));
      T* o_d = static_cast<T*>(dO.p);
      timed(iters, elapsed_ms, [&] { layernorm<T>(x_d, D, g, b, o_d, D, M, D, 0); });
      download<T>(o_d, out, cnt);
    };
    if (precision == VB_PRECISION_FP32) run(float());
    else run(__nv_bfloat16());
  });
}

int vb_op_ln_linear(const float* x, const float* gamma, const float* beta, const float* w, const float* bias, int32_t gelu,
                    float* out, int32_t M, int32_t N, int32_t K, int32_t iters, float* elapsed_ms) {
  return guarded(nullptr, [&] {
    require_gpu();
    VB_CHECK(x && gamma && beta && w && out && M > 0 && N > 0 && K > 0, "vb_op_ln_linear: bad arguments");
    VB_CHECK(N % 64 == 0 && K % 64 == 0, "vb_op_ln_linear: the folded form needs N % 64 == 0 and K % 64 == 0");
    DevMem dX, dG, dBt, dW, dWt, dB, dC, dSt, dO;
    const __nv_bfloat16* dx = upload<__nv_bfloat16>(dX, x, static_cast<size_t>(M) * K);
    const float* g = upload<float>(dG, gamma, K);
    const float* bt = upload<float>(dBt, beta, K);
    const float* dw = upload<float>(dW, w, static_cast<size_t>(K) * N);
    const float* db = bias ? upload<float>(dB, bias, N) : nullptr;
    dWt.ensure(static_cast<size_t>(N) * K * 2);
    __nv_bfloat16* wt = static_cast<__nv_bfloat16*>(dWt.p);
    pack_weight_bf16(dw, wt, K, N, K, 0, g);                       // gamma folded into the packed weight rows
    dC.ensure(static_cast<size_t>(N) * 2 * sizeof(float));
    float* c = static_cast<float*>(dC.p);
    ln_fold_consts(dw, wt, K, bt, db, c, c + N, K, N, 0);
    dSt.ensure(static_cast<size_t>(M) * (K / 64) * 2 * sizeof(float));
    float* st = static_cast<float*>(dSt.p);
    dO.ensure(static_cast<size_t>(M) * N * 2);
    __nv_bfloat16* dout = static_cast<__nv_bfloat16*>(dO.p);
    GemmBf16 gm = gemm_bf16_plan(dx, K, wt, K, dout, N, M, N, K, c + N, nullptr, nullptr, 0, gelu != 0);
    gm.ln_c1 = c; gm.ln_stats = st; gm.ln_parts = K / 64; gm.ln_inv_d = 1.0f / static_cast<float>(K);
    timed(iters, elapsed_ms, [&] { row_stats_bf16(dx, K, st, M, K, 0); gemm_bf16_run(gm, 0); });
    download<__nv_bfloat16>(dout, out, static_cast<size_t>(M) * N);
  });
}

}  // extern "C"
