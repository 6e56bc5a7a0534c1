// Issue-rate microbenchmark for the instruction mixes of the GEMM / attention epilogues (sm_100a).
// Each warp runs 8 independent dependency chains of one instruction kind; cycles per warp-instruction per SM
// sub-partition are reported for 1, 2 and 4 warps per sub-partition.   nvcc -arch=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int KIND>
__global__ void bench(float* out, long long* cycles) {
  float a[8];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = (static_cast<unsigned long long>(__float_as_uint(a[i])) << 32) | __float_as_uint(a[i] + 0.5f); }
  float c = out[0], d = out[1];
  unsigned long long cc = (static_cast<unsigned long long>(__float_as_uint(c)) << 32) | __float_as_uint(c);
  unsigned long long dd = (static_cast<unsigned long long>(__float_as_uint(d)) << 32) | __float_as_uint(d);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));                 // 3-register FFMA
      if (KIND == 1) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(a[i]) : "f"(c));                   // immediate addend
      if (KIND == 2) asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, %1;" : "+f"(a[i]) : "f"(d));                   // immediate multiplier
      if (KIND == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(cc), "l"(dd));               // FFMA2
      if (KIND == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));                            // FMUL2
      if (KIND == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(dd));                            // FADD2
      if (KIND == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));                                     // MUFU.EX2
      if (KIND == 7) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));                                   // FMNMX
      if (KIND == 8) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));                                // FMUL
      if (KIND == 9) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(c)); a[i] = __uint_as_float(r); }   // F2FP
      if (KIND == 10) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));                                    // MUFU.RCP
      if (KIND == 11) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(d));                                  // FADD
      if (KIND == 12) { unsigned r = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }   // bf16x2: ptxas emits TWO MUFU.EX2.BF16 (lo, .H1) + PRMT -- no packed SFU op on sm_100a
      if (KIND == 13) { unsigned r = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }        // MUFU.EX2 f16x2
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(static_cast<unsigned>(p[i])) + __uint_as_float(static_cast<unsigned>(p[i] >> 32));
  if (s == 123.456f) out[2] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int KIND>
void run(const char* name, float* dout, long long* dcyc) {
  printf("%-28s", name);
  for (int wps : {1, 2, 4}) {
    bench<KIND><<<1, 128 * wps>>>(dout, dcyc);
    cudaDeviceSynchronize();
    bench<KIND><<<1, 128 * wps>>>(dout, dcyc);
    long long c = 0;
    cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
    printf("  %d warp/SMSP: %.2f cyc/instr/SMSP", wps, static_cast<double>(c) / (ITER * 8.0 * wps));
  }
  printf("\n");
}

int main() {
  float* dout; long long* dcyc;
  cudaMalloc(&dout, 64); cudaMalloc(&dcyc, 64);
  float h[4] = {0.999f, 0.001f, 0.f, 0.f};
  cudaMemcpy(dout, h, 16, cudaMemcpyHostToDevice);
  run<0>("FFMA r,r,r", dout, dcyc);
  run<1>("FFMA r,r,imm (addend)", dout, dcyc);
  run<2>("FFMA r,imm,r (multiplier)", dout, dcyc);
  run<3>("FFMA2", dout, dcyc);
  run<4>("FMUL2", dout, dcyc);
  run<5>("FADD2", dout, dcyc);
  run<8>("FMUL", dout, dcyc);
  run<11>("FADD", dout, dcyc);
  run<6>("MUFU.EX2", dout, dcyc);
  run<10>("MUFU.RCP", dout, dcyc);
  run<12>("MUFU.EX2 bf16x2 (2 results)", dout, dcyc);
  run<13>("MUFU.EX2 f16x2 (2 results)", dout, dcyc);
  run<7>("FMNMX", dout, dcyc);
  run<9>("F2FP.BF16 pack", dout, dcyc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
