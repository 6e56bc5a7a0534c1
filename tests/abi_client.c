/* A plain-C99 client of include/vitb200.h: proves that the boundary is a C ABI (this file is compiled by gcc, not g++ or
 * nvcc, and includes nothing but the header and libc) and exercises it without any Python host code.
 *
 *   abi_client <out.txt>      builds a small ViT (vit.py:107-157 kwargs flattened into vb_config), fills every weight the
 *                             engine asks for with a closed-form pattern, runs vb_forward on a closed-form image with
 *                             HOST buffers and writes the logits as text.  tests/test_abi_c_client.py regenerates the same
 *                             weights / image in numpy and checks the logits against the oracle.
 * Exit code: 0 ok, 3 "no CUDA device" (the no-CPU-fallback contract on a machine without a GPU), 1 anything else. */
#include "vitb200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* uniform [-1, 1) from a 32-bit integer hash (murmur3 finaliser): bit-identical in C and in numpy */
static double hash_unit(unsigned seed, unsigned long i) {
  unsigned x = (unsigned)(i * 2654435761ul) + seed;
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return (double)x / 4294967296.0 * 2.0 - 1.0;
}

/* value of element i of the weight called `name`: FNV-1a of the name seeds the hash, the leaf name picks the distribution
 * (the reference's: glorot-uniform kernels; perturbed LayerNorm / bias values so that affine wiring bugs cannot hide) */
static float pattern(const char* name, long i, const int64_t* shape, int nd) {
  unsigned h = 2166136261u;
  const char* p;
  const char* leaf = strrchr(name, '.');
  double v;
  for (p = name; *p; ++p) h = (h ^ (unsigned char)*p) * 16777619u;
  v = hash_unit(h, (unsigned long)i);
  leaf = leaf ? leaf + 1 : name;
  if (strcmp(leaf, "kernel") == 0 && nd == 2) return (float)(v * sqrt(6.0 / (double)(shape[0] + shape[1])));
  if (strcmp(leaf, "gamma") == 0) return (float)(1.0 + 0.2 * v);
  if (strcmp(leaf, "bias") == 0 || strcmp(leaf, "beta") == 0) return (float)(0.2 * v);
  return (float)(1.7 * v); /* pos_embedding, cls_token: about unit variance */
}

int main(int argc, char** argv) {
  vb_config cfg;
  vb_handle* h = NULL;
  int rc, i, n, b = 3;
  const int H = 32, W = 48, C = 3, classes = 7;
  float *img, *logits;
  long k;
  FILE* f;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = (int32_t)sizeof cfg;
  cfg.kind = VB_KIND_VIT;
  cfg.precision = (argc > 2 && strcmp(argv[2], "bf16") == 0) ? VB_PRECISION_BF16 : VB_PRECISION_FP32;
  cfg.image_h = H; cfg.image_w = W; cfg.patch_h = 8; cfg.patch_w = 16; cfg.channels = C;
  cfg.num_classes = classes; cfg.dim = 64; cfg.depth = 2; cfg.heads = 2; cfg.dim_head = 32; cfg.mlp_dim = 96;
  cfg.pool = VB_POOL_CLS;
  if (vb_abi_version() != VB_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
  rc = vb_create(&cfg, 0, &h);
  if (rc != 0) {
    const char* msg = vb_last_error(NULL);
    fprintf(stderr, "vb_create: %s\n", msg);
    return strstr(msg, "no CUDA device") ? 3 : 1;
  }
  n = vb_num_weights(h);
  for (i = 0; i < n; ++i) {
    const char* name; int64_t shape[4]; int32_t nd; long count = 1; float* w; int d;
    if (vb_weight_info(h, i, &name, shape, &nd) != 0) { fprintf(stderr, "%s\n", vb_last_error(h)); return 1; }
    for (d = 0; d < nd; ++d) count *= (long)shape[d];
    w = (float*)malloc(sizeof(float) * (size_t)count);
    for (k = 0; k < count; ++k) w[k] = pattern(name, k, shape, nd);
    if (vb_set_weight(h, name, w, shape, nd) != 0) { fprintf(stderr, "%s\n", vb_last_error(h)); return 1; }
    free(w);
  }
  if (vb_finalize(h) != 0) { fprintf(stderr, "%s\n", vb_last_error(h)); return 1; }
  img = (float*)malloc(sizeof(float) * (size_t)(b * H * W * C));
  logits = (float*)malloc(sizeof(float) * (size_t)(b * classes));
  for (k = 0; k < (long)b * H * W * C; ++k) img[k] = (float)(1.7 * hash_unit(12345u, (unsigned long)k));
  if (vb_forward(h, img, VB_MEM_HOST, b, H, W, logits, VB_MEM_HOST, NULL) != 0) { fprintf(stderr, "%s\n", vb_last_error(h)); return 1; }
  f = fopen(argc > 1 ? argv[1] : "abi_client_logits.txt", "w");
  if (!f) return 1;
  for (i = 0; i < b * classes; ++i) fprintf(f, "%.9g\n", (double)logits[i]);
  fclose(f);
  printf("abi_client: %d weights, %lld kernel launches\n", n, (long long)vb_last_launch_count(h));
  vb_destroy(h);
  free(img); free(logits);
  return 0;
}
