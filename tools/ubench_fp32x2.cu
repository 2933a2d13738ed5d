// Microbenchmark: issue throughput of FFMA vs FFMA2 (packed f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32x2 ubench_fp32x2.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int MODE>
__global__ void kern(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float m = 1.0000001f, c = 1e-7f;
  unsigned long long p0, p1, p2, p3, pm, pc;
  { float2 t = make_float2(a0, a1); p0 = *(unsigned long long*)&t; t = make_float2(a2, a3); p1 = *(unsigned long long*)&t;
    t = make_float2(a4, a5); p2 = *(unsigned long long*)&t; t = make_float2(a6, a7); p3 = *(unsigned long long*)&t;
    t = make_float2(m, m); pm = *(unsigned long long*)&t; t = make_float2(c, c); pc = *(unsigned long long*)&t; }
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // 8 scalar FFMA = 8 flop-lanes
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
        a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
      }
    } else if (MODE == 1) {  // 4 FFMA2 = same 8 lanes of work
#pragma unroll
      for (int u = 0; u < 4; ++u) { p0 = ffma2(p0, pm, pc); p1 = ffma2(p1, pm, pc); p2 = ffma2(p2, pm, pc); p3 = ffma2(p3, pm, pc); }
    } else if (MODE == 2) {  // FADD2
#pragma unroll
      for (int u = 0; u < 4; ++u) { p0 = fadd2(p0, pc); p1 = fadd2(p1, pc); p2 = fadd2(p2, pc); p3 = fadd2(p3, pc); }
    } else {  // scalar FADD
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a0 = __fadd_rn(a0, c); a1 = __fadd_rn(a1, c); a2 = __fadd_rn(a2, c); a3 = __fadd_rn(a3, c);
        a4 = __fadd_rn(a4, c); a5 = __fadd_rn(a5, c); a6 = __fadd_rn(a6, c); a7 = __fadd_rn(a7, c);
      }
    }
  }
  float2 r0 = *(float2*)&p0, r1 = *(float2*)&p1, r2 = *(float2*)&p2, r3 = *(float2*)&p3;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + r0.x + r0.y + r1.x + r1.y + r2.x + r2.y + r3.x + r3.y;
}

template <int MODE>
void run(const char* name, float* out) {
  const int iters = 20000, blocks = 148 * 4, threads = 512;
  kern<MODE><<<blocks, threads>>>(out, 100, 1.0f);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<MODE><<<blocks, threads>>>(out, iters, 1.0f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double lanes = (double)blocks * threads * iters * 32.0;  // fp32 lane-ops (8 per inner x4)
  printf("%-12s %8.3f ms  %8.2f Tlane-op/s  (%.1f lane-ops/clk/SM @1.9GHz)\n", name, ms, lanes / ms / 1e9,
         lanes / (ms * 1e-3) / 148 / 1.9e9);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  run<0>("FFMA", out); run<1>("FFMA2", out); run<3>("FADD", out); run<2>("FADD2", out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
