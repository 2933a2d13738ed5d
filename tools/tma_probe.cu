// Standalone probe of the 4-D TMA box load used by resample_tile.cu.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BOX>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, int c3, float* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* box = (float*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  unsigned long long* bar = (unsigned long long*)(box + BOX * BOX * BOX);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"((uint32_t)(BOX * BOX * BOX * 4)) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(box)),
        "l"((unsigned long long)&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
  }
  uint32_t done; int spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
  } while (!done && ++spins < 2000000);
  if (threadIdx.x == 0) out[BOX * BOX * BOX] = (float)spins;
  for (int t = threadIdx.x; t < BOX * BOX * BOX; t += blockDim.x) out[t] = box[t];
}

int main(int argc, char** argv) {
  const int I = 64, J = 64, K = 64, BC = 2;
  std::vector<float> h((size_t)BC * I * J * K);
  for (size_t t = 0; t < h.size(); ++t) h[t] = (float)(t % 1000003);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn encode = (EncodeTiledFn)p;
  printf("entry %p q=%d\n", p, (int)q);
  constexpr int BOX = 24;
  CUtensorMap tm;
  const cuuint64_t gdim[4] = {K, J, I, BC};
  const cuuint64_t gstride[3] = {K * 4ull, (cuuint64_t)J * K * 4, (cuuint64_t)I * J * K * 4};
  const cuuint32_t bdim[4] = {BOX, BOX, BOX, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gdim, gstride, bdim, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)rc);
  cudaMalloc(&o, (BOX * BOX * BOX + 1) * 4);
  size_t smem = 128 + BOX * BOX * BOX * 4 + 64;
  cudaFuncSetAttribute(probe<BOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int c0 = argc > 1 ? atoi(argv[1]) : -2, c1 = argc > 2 ? atoi(argv[2]) : 5, c2 = argc > 3 ? atoi(argv[3]) : 50, c3 = argc > 4 ? atoi(argv[4]) : 1;
  printf("coords %d %d %d %d\n", c0, c1, c2, c3);
  probe<BOX><<<1, 256, smem>>>(tm, c0, c1, c2, c3, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  std::vector<float> r(BOX * BOX * BOX + 1);
  cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost);
  printf("spins %.0f\n", r[BOX * BOX * BOX]);
  long bad = 0;
  for (int a = 0; a < BOX; ++a) for (int b = 0; b < BOX; ++b) for (int c = 0; c < BOX; ++c) {
    int i = c2 + a, j = c1 + b, k = c0 + c;
    float want = (i >= 0 && i < I && j >= 0 && j < J && k >= 0 && k < K) ? h[(((size_t)c3 * I + i) * J + j) * K + k] : 0.f;
    if (r[(a * BOX + b) * BOX + c] != want) ++bad;
  }
  printf("mismatches %ld\n", bad);
  return 0;
}
