// patches.cu — patch extraction for the Queue path.
//
// The reference cuts patches one at a time as tensor views
// (`image[:, si, sj, sk]`, data/sampler.py:54-67) and later copies each of them
// again in `torch.stack` when a batch is collated (loader.py:15-24,
// data/batch.py:52-58).  On the device one launch gathers all `n` patches of a
// volume into a dense (n, C, pi, pj, pk) block, reading every source byte once.
// Pure data movement: bound by HBM, algorithmic bytes = 2 x patch bytes.
#include "common.cuh"

namespace tio {

template <typename T>
__global__ void __launch_bounds__(256)
crop_patches_kernel(const T* __restrict__ src, T* __restrict__ dst, int C, int I, int J, int K, int n,
                    const int32_t* __restrict__ corners, int pi, int pj, int pk) {
  // one CTA row-group: blockIdx.x -> (patch, channel, i, group of rows j), threads run along k
  const int rows_per_block = blockDim.y;
  const long long row = (long long)blockIdx.x * rows_per_block + threadIdx.y;  // over n*C*pi*pj
  const long long total_rows = (long long)n * C * pi * pj;
  if (row >= total_rows) return;
  const int j = (int)(row % pj);
  const int i = (int)((row / pj) % pi);
  const int c = (int)((row / ((long long)pj * pi)) % C);
  const int p = (int)(row / ((long long)pj * pi * C));
  const int ci = corners[3 * p + 0], cj = corners[3 * p + 1], ck = corners[3 * p + 2];
  const T* s = src + (((long long)c * I + (ci + i)) * J + (cj + j)) * K + ck;
  T* d = dst + row * pk;
  for (int k = threadIdx.x; k < pk; k += blockDim.x) d[k] = s[k];
}

template <typename T>
static void launch_crop(const void* src, void* dst, int C, int I, int J, int K, int n,
                        const int32_t* corners, int pi, int pj, int pk, cudaStream_t st) {
  const long long rows = (long long)n * C * pi * pj;
  const int tx = pk >= 128 ? 128 : (pk >= 64 ? 64 : 32);
  dim3 block(tx, 256 / tx);
  const unsigned blocks = (unsigned)((rows + block.y - 1) / block.y);
  crop_patches_kernel<T><<<blocks, block, 0, st>>>((const T*)src, (T*)dst, C, I, J, K, n, corners, pi, pj, pk);
}

}  // namespace tio

using namespace tio;

extern "C" int tio_crop_patches(const void* src, void* dst, int elem_bytes, int C, int I, int J,
                                int K, int n, const int32_t* corners, int pi, int pj, int pk,
                                void* stream) {
  TIO_CHECK_ARG(src && dst && corners, "tio_crop_patches: null pointer");
  TIO_CHECK_ARG(C > 0 && I > 0 && J > 0 && K > 0 && n > 0, "tio_crop_patches: bad shape");
  TIO_CHECK_ARG(pi > 0 && pj > 0 && pk > 0 && pi <= I && pj <= J && pk <= K,
                "tio_crop_patches: patch (%d,%d,%d) does not fit the volume (%d,%d,%d)", pi, pj, pk, I, J, K);
  TIO_CHECK_ARG((long long)n * C * pi * pj / 2 < (1ll << 31), "tio_crop_patches: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  switch (elem_bytes) {
    case 1: launch_crop<uint8_t>(src, dst, C, I, J, K, n, corners, pi, pj, pk, st); break;
    case 2: launch_crop<uint16_t>(src, dst, C, I, J, K, n, corners, pi, pj, pk, st); break;
    case 4: launch_crop<uint32_t>(src, dst, C, I, J, K, n, corners, pi, pj, pk, st); break;
    case 8: launch_crop<uint64_t>(src, dst, C, I, J, K, n, corners, pi, pj, pk, st); break;
    default: TIO_CHECK_ARG(false, "tio_crop_patches: element size %d not in {1,2,4,8}", elem_bytes);
  }
  TIO_CHECK_LAUNCH();
  return 0;
}
