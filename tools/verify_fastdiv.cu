// Exhaustive check: for a constant divisor d, is
//     q0 = rn(x*r); e = fma(-q0, d, x); q = fma(e, r, q0)      (r = rn(1/d))
// bit-identical to the IEEE quotient __fdiv_rn(x, d) for EVERY float x?
// Reports mismatches per magnitude band so the kernel can guard only what fails.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>

__global__ void verify(float d, float r, unsigned long long* bad, unsigned long long* bad_band) {
  const unsigned long long total = 1ull << 32;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned bits = (unsigned)t;
    const float x = __uint_as_float(bits);
    if ((bits & 0x7f800000u) == 0x7f800000u) continue;  // inf / nan
    const float want = __fdiv_rn(x, d);
    const float q0 = __fmul_rn(x, r);
    const float e = __fmaf_rn(-q0, d, x);
    const float q = __fmaf_rn(e, r, q0);
    if (__float_as_uint(q) != __float_as_uint(want)) {
      atomicAdd(bad, 1ull);
      atomicAdd(&bad_band[(bits >> 23) & 0xff], 1ull);
    }
  }
}

int main(int argc, char** argv) {
  unsigned long long *bad, *band;
  cudaMalloc(&bad, 8); cudaMalloc(&band, 256 * 8);
  int sizes[] = {2, 3, 4, 5, 8, 16, 17, 32, 33, 48, 64, 65, 96, 100, 128, 129, 160, 176, 192, 200, 224, 240, 256, 257, 320, 384, 448, 512, 513, 640, 768, 1024};
  for (int s : sizes) {
    float d = (float)(s - 1) * 0.5f;           // hd = max(size-1,1)/2
    float r = (float)(1.0 / (double)d);        // rn(1/d): double quotient rounded once (exact enough: checked below)
    cudaMemset(bad, 0, 8); cudaMemset(band, 0, 256 * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    verify<<<148 * 16, 256>>>(d, r, bad, band);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    unsigned long long hb, hband[256];
    cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hband, band, 256 * 8, cudaMemcpyDeviceToHost);
    int lo = 256, hi = -1;
    for (int e = 0; e < 256; ++e) if (hband[e]) { if (e < lo) lo = e; if (e > hi) hi = e; }
    printf("size %4d  d=%8.2f  mismatches=%llu  biased-exponent range of failing x: [%d, %d]  (%.2f ms)\n",
           s, d, hb, lo, hi, ms);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
