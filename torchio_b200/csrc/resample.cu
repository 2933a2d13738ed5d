// resample.cu — K1: fused affine + elastic displacement + gather + fill.
//
// One pass replaces the reference's grid construction and sampling
// (transforms/spatial/spatial.py:1504-1579, 1651-1731, 1764-1857 of TorchIO
// 2.0.0a2): arange/meshgrid/cat/sgemm, F.interpolate(trilinear), two
// F.grid_sample passes and torch.where.  Coordinate arithmetic reproduces the
// CPU op order of those ATen calls rounding-for-rounding (see
// oracle/c/tio_oracle.c, which is bit-exact against the reference), so
// nearest-neighbour label maps and fill decisions are bit-identical.
//
// Work decomposition: a CTA owns a (TJ x TK) window of output columns and
// walks TI output planes along I.  A thread owns one (j,k) column, so
//   - stores (and, for small rotations, gathers) are coalesced along K;
//   - the displacement field d = lerp_I(lerp_J(lerp_K(cp))) keeps its two
//     inner levels in registers across the I walk (recomputed only when the
//     walk enters a new control-grid cell), i.e. ~6 FMA/voxel instead of 42.
#include "common.cuh"

namespace tio {

template <typename T>
struct ElemTraits;
template <>
struct ElemTraits<float> {
  static __device__ __forceinline__ float to_f32(float v) { return v; }
  static __device__ __forceinline__ float from_f32(float v) { return v; }
};
#define TIO_INT_TRAITS(T)                                                    \
  template <>                                                                \
  struct ElemTraits<T> {                                                     \
    static __device__ __forceinline__ float to_f32(T v) { return (float)v; } \
    static __device__ __forceinline__ T from_f32(float v) {                  \
      return (T)(long long)v; /* Tensor.to(int): truncation */               \
    }                                                                        \
  };
TIO_INT_TRAITS(uint8_t)
TIO_INT_TRAITS(int8_t)
TIO_INT_TRAITS(int16_t)
TIO_INT_TRAITS(int32_t)
TIO_INT_TRAITS(int64_t)

struct ResampleArgs {
  const void* src;
  void* dst;
  const float* mat;      // [B][12]
  const float* cp;       // [B][ni][nj][nk][3] or null
  const uint8_t* flags;  // [B] or null
  const float* fill;     // [C] or null
  int B, C, I, J, K, OI, OJ, OK;
  int ni, nj, nk;
  float sc_i, sc_j, sc_k;        // (n-1)/(O-1) upsample scales (fp32)
  float sp_in[3], sp_out[3];     // spacings
  float nm1[3];                  // max(size-1, 1) (normalise)
  float sm1[3];                  // size-1       (ATen un-normalise)
  int affine_first;
  int cp_in_smem;
};

constexpr int TK = 64;   // lanes along K
constexpr int TJ = 4;    // rows along J  -> 256 threads
constexpr int TI = 16;   // planes walked per CTA

// [p,1] @ M^T exactly as the reference's CPU sgemm rounds it: sequential FMA
// chain from the rounded first product (spatial.py:1621-1624).
__device__ __forceinline__ float affine_row(const float* m, float pi, float pj, float pk) {
  float acc = __fmul_rn(pi, m[0]);
  acc = __fmaf_rn(pj, m[1], acc);
  acc = __fmaf_rn(pk, m[2], acc);
  acc = __fmaf_rn(1.0f, m[3], acc);
  return acc;
}

// 2.0*q/nm1 - 1.0 (spatial.py:1646) then ATen's ((g+1)/2)*(size-1).
__device__ __forceinline__ float renormalise(float q, float nm1, float sm1) {
  float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, q), nm1), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), sm1);
}

template <typename T, int MODE, bool HAS_CP, bool HAS_FILL>
__global__ void __launch_bounds__(TK* TJ)
resample_kernel(const ResampleArgs a) {
  extern __shared__ float smem_cp[];
  const int tiles_i = (a.OI + TI - 1) / TI;
  const int b = blockIdx.z / tiles_i;
  const int oi0 = (blockIdx.z % tiles_i) * TI;
  const int ok = blockIdx.x * TK + threadIdx.x;
  const int oj = blockIdx.y * TJ + threadIdx.y;
  const int64_t n_in = (int64_t)a.I * a.J * a.K;
  const int64_t n_out = (int64_t)a.OI * a.OJ * a.OK;
  const uint8_t fl = a.flags ? a.flags[b] : 0;
  const T* __restrict__ src = (const T*)a.src + (int64_t)b * a.C * n_in;
  T* __restrict__ dst = (T*)a.dst + (int64_t)b * a.C * n_out;
  const int oi_end = min(oi0 + TI, a.OI);

  if (fl & TIO_FLAG_PASSTHROUGH) {  // exact copy (spatial.py:1101-1106)
    if (ok < a.OK && oj < a.OJ)
      for (int c = 0; c < a.C; ++c)
        for (int oi = oi0; oi < oi_end; ++oi) {
          int64_t o = ((int64_t)oi * a.OJ + oj) * a.OK + ok;
          dst[c * n_out + o] = src[c * n_in + o];
        }
    return;
  }

  const bool elastic = HAS_CP && (fl & TIO_FLAG_ELASTIC);
  const float* g = nullptr;
  if (HAS_CP && elastic) {
    const int ncp = a.ni * a.nj * a.nk * 3;
    const float* gsrc = a.cp + (int64_t)b * ncp;
    if (a.cp_in_smem) {
      for (int t = threadIdx.y * TK + threadIdx.x; t < ncp; t += TK * TJ) smem_cp[t] = gsrc[t];
      __syncthreads();
      g = smem_cp;
    } else {
      g = gsrc;
    }
  }
  if (ok >= a.OK || oj >= a.OJ) return;

  float m[12];
#pragma unroll
  for (int t = 0; t < 12; ++t) m[t] = a.mat[b * 12 + t];

  // per-thread J/K lerp setup for the displacement field
  LerpAxis lj, lk;
  int64_t o00 = 0, o01 = 0, o10 = 0, o11 = 0;  // (j0|j1, k0|k1) offsets in cp, x3
  if (HAS_CP && elastic) {
    lj = lerp_axis(a.sc_j, a.nj, oj);
    lk = lerp_axis(a.sc_k, a.nk, ok);
    o00 = ((int64_t)lj.i0 * a.nk + lk.i0) * 3;
    o01 = ((int64_t)lj.i0 * a.nk + lk.i1) * 3;
    o10 = ((int64_t)lj.i1 * a.nk + lk.i0) * 3;
    o11 = ((int64_t)lj.i1 * a.nk + lk.i1) * 3;
  }
  const int plane = a.nj * a.nk * 3;
  int cur_i0 = -1, cur_i1 = -1;
  float r_lo[3] = {0.f, 0.f, 0.f}, r_hi[3] = {0.f, 0.f, 0.f};

  const float pj = (float)oj, pk = (float)ok;
  for (int oi = oi0; oi < oi_end; ++oi) {
    const float pi = (float)oi;
    float d[3] = {0.f, 0.f, 0.f};
    if (HAS_CP && elastic) {
      LerpAxis li = lerp_axis(a.sc_i, a.ni, oi);  // warp-uniform
      if (li.i0 != cur_i0 || li.i1 != cur_i1) {
        const float* p0 = g + (int64_t)li.i0 * plane;
        const float* p1 = g + (int64_t)li.i1 * plane;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          float a00 = lerp2(lk.l0, p0[o00 + ax], lk.l1, p0[o01 + ax]);
          float a01 = lerp2(lk.l0, p0[o10 + ax], lk.l1, p0[o11 + ax]);
          r_lo[ax] = lerp2(lj.l0, a00, lj.l1, a01);
          float b00 = lerp2(lk.l0, p1[o00 + ax], lk.l1, p1[o01 + ax]);
          float b01 = lerp2(lk.l0, p1[o10 + ax], lk.l1, p1[o11 + ax]);
          r_hi[ax] = lerp2(lj.l0, b00, lj.l1, b01);
        }
        cur_i0 = li.i0;
        cur_i1 = li.i1;
      }
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) d[ax] = lerp2(li.l0, r_lo[ax], li.l1, r_hi[ax]);
    }

    float q[3];
    if (!(HAS_CP && elastic)) {
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) q[ax] = affine_row(m + 4 * ax, pi, pj, pk);
    } else if (a.affine_first) {  // spatial.py:1570-1573
#pragma unroll
      for (int ax = 0; ax < 3; ++ax)
        q[ax] = __fadd_rn(affine_row(m + 4 * ax, pi, pj, pk), __fdiv_rn(d[ax], a.sp_in[ax]));
    } else {  // spatial.py:1574-1577
      float e0 = __fadd_rn(pi, __fdiv_rn(d[0], a.sp_out[0]));
      float e1 = __fadd_rn(pj, __fdiv_rn(d[1], a.sp_out[1]));
      float e2 = __fadd_rn(pk, __fdiv_rn(d[2], a.sp_out[2]));
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) q[ax] = affine_row(m + 4 * ax, e0, e1, e2);
    }
    float u[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) u[ax] = renormalise(q[ax], a.nm1[ax], a.sm1[ax]);

    // trilinear corner weights: needed for MODE==linear and for the mask
    float f0 = floorf(u[0]), f1 = floorf(u[1]), f2 = floorf(u[2]);
    // clamp before the int conversion so wild coordinates stay out of bounds
    int c0 = (int)fminf(fmaxf(f0, -2.0f), (float)a.I);
    int c1 = (int)fminf(fmaxf(f1, -2.0f), (float)a.J);
    int c2 = (int)fminf(fmaxf(f2, -2.0f), (float)a.K);
    const bool interior = (c0 >= 0) & (c0 + 1 < a.I) & (c1 >= 0) & (c1 + 1 < a.J) &
                          (c2 >= 0) & (c2 + 1 < a.K);

    float w[8];
    bool inb[8];
    bool need_w = (MODE == TIO_LINEAR) || (HAS_FILL && !interior);
    if (need_w) {
      // ATen: weight_lo = (c+1) - u, weight_hi = u - c  (exact ints as floats)
      float lo0 = __fsub_rn(__fadd_rn(f0, 1.0f), u[0]), hi0 = __fsub_rn(u[0], f0);
      float lo1 = __fsub_rn(__fadd_rn(f1, 1.0f), u[1]), hi1 = __fsub_rn(u[1], f1);
      float lo2 = __fsub_rn(__fadd_rn(f2, 1.0f), u[2]), hi2 = __fsub_rn(u[2], f2);
      float w00 = __fmul_rn(lo0, lo1), w10 = __fmul_rn(hi0, lo1);
      float w01 = __fmul_rn(lo0, hi1), w11 = __fmul_rn(hi0, hi1);
      // order: i fastest, then j, then k  (tnw, tne, tsw, tse, bnw, ...)
      w[0] = __fmul_rn(w00, lo2); w[1] = __fmul_rn(w10, lo2);
      w[2] = __fmul_rn(w01, lo2); w[3] = __fmul_rn(w11, lo2);
      w[4] = __fmul_rn(w00, hi2); w[5] = __fmul_rn(w10, hi2);
      w[6] = __fmul_rn(w01, hi2); w[7] = __fmul_rn(w11, hi2);
    }
    if (!interior) {
      const bool i_lo = (c0 >= 0) & (c0 < a.I), i_hi = (c0 + 1 >= 0) & (c0 + 1 < a.I);
      const bool j_lo = (c1 >= 0) & (c1 < a.J), j_hi = (c1 + 1 >= 0) & (c1 + 1 < a.J);
      const bool k_lo = (c2 >= 0) & (c2 < a.K), k_hi = (c2 + 1 >= 0) & (c2 + 1 < a.K);
      inb[0] = i_lo & j_lo & k_lo; inb[1] = i_hi & j_lo & k_lo;
      inb[2] = i_lo & j_hi & k_lo; inb[3] = i_hi & j_hi & k_lo;
      inb[4] = i_lo & j_lo & k_hi; inb[5] = i_hi & j_lo & k_hi;
      inb[6] = i_lo & j_hi & k_hi; inb[7] = i_hi & j_hi & k_hi;
    }
    bool use_fill = false;
    if (HAS_FILL && !interior) {
      float msum = 0.0f;
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (inb[t]) msum = __fadd_rn(msum, w[t]);
      use_fill = !(msum > 0.5f);
    }

    const int64_t o_off = ((int64_t)oi * a.OJ + oj) * a.OK + ok;
    if (MODE == TIO_NEAREST) {
      // round-half-to-even like std::nearbyint; clamp keeps the cvt in range
      int r0 = __float2int_rn(fminf(fmaxf(u[0], -2.0f), (float)a.I + 1.0f));
      int r1 = __float2int_rn(fminf(fmaxf(u[1], -2.0f), (float)a.J + 1.0f));
      int r2 = __float2int_rn(fminf(fmaxf(u[2], -2.0f), (float)a.K + 1.0f));
      const bool ok_in = (r0 >= 0) & (r0 < a.I) & (r1 >= 0) & (r1 < a.J) & (r2 >= 0) & (r2 < a.K);
      const int64_t off = ((int64_t)r0 * a.J + r1) * a.K + r2;
      for (int c = 0; c < a.C; ++c) {
        T v;
        if (HAS_FILL && use_fill) v = ElemTraits<T>::from_f32(a.fill[c]);
        else v = ok_in ? __ldg(src + c * n_in + off) : (T)0;
        dst[c * n_out + o_off] = v;
      }
    } else {
      const int64_t base = ((int64_t)c0 * a.J + c1) * a.K + c2;
      const int64_t sI = (int64_t)a.J * a.K, sJ = a.K;
      for (int c = 0; c < a.C; ++c) {
        const T* s = src + c * n_in;
        float v = 0.0f;
        if (HAS_FILL && use_fill) {
          v = a.fill[c];
        } else if (interior) {
          const T* p = s + base;
          v = __fadd_rn(v, __fmul_rn(ElemTraits<T>::to_f32(__ldg(p)), w[0]));
          v = __fadd_rn(v, __fmul_rn(ElemTraits<T>::to_f32(__ldg(p + sI)), w[1]));
          v = __fadd_rn(v, __fmul_rn(ElemTraits<T>::to_f32(__ldg(p + sJ)), w[2]));
          v = __fadd_rn(v, __fmul_rn(ElemTraits<T>::to_f32(__ldg(p + sI + sJ)), w[3]));
          v = __fadd_rn(v, __fmul_rn(ElemTraits<T>::to_f32(__ldg(p + 1)), w[4]));
          v = __fadd_rn(v, __fmul_rn(ElemTraits<T>::to_f32(__ldg(p + sI + 1)), w[5]));
          v = __fadd_rn(v, __fmul_rn(ElemTraits<T>::to_f32(__ldg(p + sJ + 1)), w[6]));
          v = __fadd_rn(v, __fmul_rn(ElemTraits<T>::to_f32(__ldg(p + sI + sJ + 1)), w[7]));
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (inb[t]) {
              int64_t off = base + (t & 1) * sI + ((t >> 1) & 1) * sJ + ((t >> 2) & 1);
              v = __fadd_rn(v, __fmul_rn(ElemTraits<T>::to_f32(__ldg(s + off)), w[t]));
            }
        }
        dst[c * n_out + o_off] = ElemTraits<T>::from_f32(v);
      }
    }
  }
}

template <typename T, int MODE, bool HAS_CP>
static int launch_fill(const ResampleArgs& a, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  if (a.fill)
    resample_kernel<T, MODE, HAS_CP, true><<<grid, block, smem, st>>>(a);
  else
    resample_kernel<T, MODE, HAS_CP, false><<<grid, block, smem, st>>>(a);
  return 0;
}

template <typename T>
static int launch_typed(const ResampleArgs& a, int mode, cudaStream_t st) {
  dim3 block(TK, TJ, 1);
  const int tiles_i = (a.OI + TI - 1) / TI;
  dim3 grid((a.OK + TK - 1) / TK, (a.OJ + TJ - 1) / TJ, (unsigned)(a.B * tiles_i));
  size_t smem = (a.cp && a.cp_in_smem) ? (size_t)a.ni * a.nj * a.nk * 3 * sizeof(float) : 0;
  const bool has_cp = a.cp != nullptr;
#define TIO_SET_SMEM(KERN)                                                              \
  if (smem > 48 * 1024)                                                                 \
    cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (has_cp) {
    TIO_SET_SMEM((resample_kernel<T, TIO_NEAREST, true, true>))
    TIO_SET_SMEM((resample_kernel<T, TIO_NEAREST, true, false>))
    TIO_SET_SMEM((resample_kernel<T, TIO_LINEAR, true, true>))
    TIO_SET_SMEM((resample_kernel<T, TIO_LINEAR, true, false>))
  }
#undef TIO_SET_SMEM
  if (mode == TIO_NEAREST) {
    if (has_cp) return launch_fill<T, TIO_NEAREST, true>(a, grid, block, smem, st);
    return launch_fill<T, TIO_NEAREST, false>(a, grid, block, smem, st);
  }
  if (has_cp) return launch_fill<T, TIO_LINEAR, true>(a, grid, block, smem, st);
  return launch_fill<T, TIO_LINEAR, false>(a, grid, block, smem, st);
}

// ---- min of sample 0 (fill value "minimum") --------------------------------

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  // valid for any mix of signs once *addr was initialised to +inf
  if (v >= 0.0f)
    atomicMin((int*)addr, __float_as_int(v));
  else
    atomicMax((unsigned int*)addr, __float_as_uint(v));
}

__global__ void min_init_kernel(float* fill, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) fill[c] = __int_as_float(0x7f800000);
}

template <bool VEC>
__global__ void __launch_bounds__(256) min_kernel(const float* __restrict__ src, int64_t n,
                                                  float* fill) {
  const int c = blockIdx.y;
  const float* base = src + (int64_t)c * n;
  float m = __int_as_float(0x7f800000);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    const float4* p = (const float4*)base;
    for (int64_t t = t0; t < (n >> 2); t += stride) {
      float4 v = __ldg(p + t);
      m = fminf(fminf(m, v.x), fminf(v.y, fminf(v.z, v.w)));
    }
  } else {
    for (int64_t t = t0; t < n; t += stride) m = fminf(m, __ldg(base + t));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, s));
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t < 8; ++t) m = fminf(m, part[t]);
    atomic_min_float(fill + c, m);
  }
}

}  // namespace tio

extern "C" int tio_resample(const void* src, void* dst, int dtype, int B, int C, int I, int J,
                            int K, int OI, int OJ, int OK, const float* mat, const float* cp,
                            const uint8_t* flags, int ni, int nj, int nk,
                            const float* spacing_in, const float* spacing_out,
                            int affine_first, int mode, const float* fill, void* stream) {
  using namespace tio;
  TIO_CHECK_ARG(src && dst && mat, "tio_resample: null src/dst/mat");
  TIO_CHECK_ARG(src != dst, "tio_resample: src and dst must not alias");
  TIO_CHECK_ARG(B > 0 && C > 0 && I > 0 && J > 0 && K > 0 && OI > 0 && OJ > 0 && OK > 0,
                "tio_resample: non-positive shape");
  TIO_CHECK_ARG(mode == TIO_NEAREST || mode == TIO_LINEAR, "tio_resample: bad mode %d", mode);
  TIO_CHECK_ARG(spacing_in && spacing_out, "tio_resample: null spacing");
  TIO_CHECK_ARG(!cp || (ni >= 2 && nj >= 2 && nk >= 2), "tio_resample: control grid < 2 per axis");
  TIO_CHECK_ARG((int64_t)B * ((OI + TI - 1) / TI) <= 65535 && (OJ + TJ - 1) / TJ <= 65535,
                "tio_resample: grid too large (B*ceil(OI/16) and ceil(OJ/4) must be <= 65535)");
  ResampleArgs a;
  a.src = src; a.dst = dst; a.mat = mat; a.cp = cp; a.flags = flags; a.fill = fill;
  a.B = B; a.C = C; a.I = I; a.J = J; a.K = K; a.OI = OI; a.OJ = OJ; a.OK = OK;
  a.ni = ni; a.nj = nj; a.nk = nk;
  auto scale = [](int n_in, int n_out) {
    if (n_in == n_out) return 1.0f;  // ATen short-circuit == identity weights
    return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f;
  };
  a.sc_i = cp ? scale(ni, OI) : 0.f; a.sc_j = cp ? scale(nj, OJ) : 0.f; a.sc_k = cp ? scale(nk, OK) : 0.f;
  const int dims[3] = {I, J, K};
  for (int t = 0; t < 3; ++t) {
    a.sp_in[t] = spacing_in[t]; a.sp_out[t] = spacing_out[t];
    a.nm1[t] = (float)(dims[t] - 1 > 1 ? dims[t] - 1 : 1);
    a.sm1[t] = (float)(dims[t] - 1);
  }
  a.affine_first = affine_first;
  a.cp_in_smem = cp && ((size_t)ni * nj * nk * 12 <= 96 * 1024);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case TIO_F32: launch_typed<float>(a, mode, st); break;
    case TIO_U8: launch_typed<uint8_t>(a, mode, st); break;
    case TIO_I8: launch_typed<int8_t>(a, mode, st); break;
    case TIO_I16: launch_typed<int16_t>(a, mode, st); break;
    case TIO_I32: launch_typed<int32_t>(a, mode, st); break;
    case TIO_I64: launch_typed<int64_t>(a, mode, st); break;
    default: TIO_CHECK_ARG(false, "tio_resample: unknown dtype %d", dtype);
  }
  TIO_CHECK_LAUNCH();
  return 0;
}

extern "C" int tio_min_sample0(const float* src, int C, int64_t n, float* fill, void* stream) {
  using namespace tio;
  TIO_CHECK_ARG(src && fill && C > 0 && n > 0, "tio_min_sample0: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  min_init_kernel<<<(C + 31) / 32, 32, 0, st>>>(fill, C);
  const bool vec = (((uintptr_t)src & 15) == 0) && ((n & 3) == 0);
  int64_t work = vec ? (n >> 2) : n;
  int blocks = (int)((work + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  if (vec)
    min_kernel<true><<<dim3(blocks, C), 256, 0, st>>>(src, n, fill);
  else
    min_kernel<false><<<dim3(blocks, C), 256, 0, st>>>(src, n, fill);
  TIO_CHECK_LAUNCH();
  return 0;
}
