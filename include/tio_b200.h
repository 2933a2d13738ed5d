/*
 * tio_b200.h — C-ABI of the B200-native 3-D augmentation hot path.
 *
 * Drop-in boundary for the TorchIO v2 (2.0.0a2 @ 2b019d2) transform kernels.
 * The reference has no FFI layer: its seam is the Python method
 *   Transform.apply_transform(batch, params)        transforms/transform.py:408-427
 * and, one level down, the plain-tensor helpers each entry point below
 * replaces (cited per function; paths relative to src/torchio/).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - Volumes are contiguous (B, C, I, J, K), K fastest, in DEVICE memory.
 *   - Parameter tables are DEVICE pointers unless marked "host"; callers pack
 *     them into one pinned staging buffer and upload it once per launch.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default).
 *   - Every function returns 0 on success, non-zero on error;
 *     tio_last_error() returns a thread-local message.  No exceptions cross
 *     the boundary, no global mutable state, re-entrant from several threads
 *     (the reference calls transforms from Queue's ThreadPoolExecutor,
 *     data/queue.py:119-123).
 *   - Outputs are caller-allocated; the library keeps no pointer after return.
 */
#ifndef TIO_B200_H
#define TIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIO_ABI_VERSION 1

/* element types accepted by tio_resample (images: F32; label maps: the rest) */
enum tio_dtype {
  TIO_F32 = 0,
  TIO_U8 = 1,
  TIO_I8 = 2,
  TIO_I16 = 3,
  TIO_I32 = 4,
  TIO_I64 = 5
};

enum tio_interp { TIO_NEAREST = 0, TIO_LINEAR = 1, TIO_LABEL_PV = 2 };
/* OR into `mode`: keep the reference's fp32 rounding sequence of the sampling coordinates on
 * every tile.  Without it, fp32 trilinear tiles whose taps all lie inside the volume evaluate the
 * same mapping with one fma per axis (differs from the reference's own coordinate noise by
 * <= ~2e-5 voxel); label maps, border tiles and fill decisions are always exact. */
#define TIO_EXACT_COORDS 0x100

/* per-element flag bits for tio_resample */
#define TIO_FLAG_PASSTHROUGH 1u /* copy the row bit-exactly (spatial.py:1101-1106) */
#define TIO_FLAG_ELASTIC 2u     /* the element has a control-point grid */

const char* tio_last_error(void);
int tio_abi_version(void);

/*
 * K1 — fused resample: affine matrix + trilinear control-point displacement +
 * 8-tap / nearest gather + out-of-bounds fill + pass-through rows, one pass.
 *
 * Replaces _build_sampling_grid + _sample_batch[_per_sample]
 *   (transforms/spatial/spatial.py:1504-1579, 1651-1731, 1764-1857)
 * i.e. arange/meshgrid/cat/matmul, F.interpolate(trilinear, align_corners),
 * two F.grid_sample(zeros, align_corners=True) passes and torch.where.
 *
 *   src, dst   (B, C, I, J, K) / (B, C, OI, OJ, OK) of `dtype`
 *   mat        [B][12] fp32: rows 0..2 of inv(A_in) @ inv(T) @ A_out
 *              (float64 product cast to fp32, spatial.py:1594-1601)
 *   cp         [B][ni][nj][nk][3] fp32 displacements in mm, or NULL
 *   flags      [B] bytes (TIO_FLAG_*), or NULL (= 0 for every element)
 *   spacing_in/out  host float[3]: fp32 casts of the affine column norms
 *              (spatial.py:1559-1568)
 *   affine_first    spatial.py:1570-1577
 *   mode       TIO_NEAREST | TIO_LINEAR (spatial.py:150-153), optionally | TIO_EXACT_COORDS;
 *              TIO_LABEL_PV = label_interpolation="label" with the default linear one-hot
 *              interpolation, fused (spatial.py:1275-1389 without materialising the one-hot
 *              channels): C must be 1, fill[0] = default_pad_label (required); per output voxel
 *              the trilinear weight of every label among the 8 taps is accumulated in
 *              grid_sample's corner order, the largest wins (smallest label on ties, as
 *              argmax over torch.unique's ascending channels), and voxels whose in-bounds
 *              weight is not > 0.5 take the pad label.  u8 / i16 / i32 take the TMA tile path
 *              (box_hint >= 0 and a workspace), any other dtype the general gather kernel.
 *   fill       [C] fp32 per-channel fill, or NULL = skip the mask step
 *              (the reference skips it only for a python-float 0.0 fill,
 *              spatial.py:2072-2076).  The mask is always the TRILINEAR
 *              in-bounds weight sum, even for nearest data (:1722-1727).
 *   box_hint   host int selecting the kernel: < 0 = general gather kernel only
 *              (exact mul+add tap sum); 0 = TMA tile path with the default
 *              24^3 input box; 20 / 22 / 24 / 28 / 32 = TMA tile path with that box edge
 *              (callers that know the matrices pick the smallest box covering
 *              the pre-image of a 16^3 output tile).  The tile path applies to
 *              fp32 + TIO_LINEAR with K % 4 == 0; anything else, and any tile
 *              whose pre-image does not fit, uses the general kernel.
 *   workspace  device scratch of tio_resample_workspace_bytes(B, OI, OJ, OK) bytes,
 *              16-byte aligned (per-tile records of the TMA path); may be NULL,
 *              which selects the general kernel.  The library allocates nothing.
 * src and dst must not alias.
 */
size_t tio_resample_workspace_bytes(int B, int OI, int OJ, int OK);
int tio_resample(const void* src, void* dst, int dtype,
                 int B, int C, int I, int J, int K,
                 int OI, int OJ, int OK,
                 const float* mat, const float* cp, const uint8_t* flags,
                 int ni, int nj, int nk,
                 const float* spacing_in, const float* spacing_out,
                 int affine_first, int mode, const float* fill, int box_hint,
                 void* workspace, size_t workspace_bytes, void* stream);

/*
 * The materialised form of label_interpolation="label", for the combinations the fused mode
 * does not cover (antialias=True blurs the one-hot channels before they are sampled,
 * spatial.py:1367-1368):
 *   tio_onehot        dst (B, n, vox) fp32 = (src (B, vox) == labels[c])  (spatial.py:1362-1365);
 *                     `labels` = the n distinct values of the batch, ascending, as device
 *                     int64 (integer dtypes) or fp32 (TIO_F32) values
 *   tio_label_argmax  dst (B, vox) of `dtype` = labels[argmax_c sampled (B, n, vox)] (first maximum),
 *                     or pad_label where the sequential channel sum is not > 0.5 (spatial.py:1378-1389)
 */
int tio_onehot(const void* src, int dtype, int B, int64_t vox, const void* labels, int n,
               float* dst, void* stream);
int tio_label_argmax(const float* sampled, int B, int n, int64_t vox, const void* labels,
                     float pad_label, void* dst, int dtype, void* stream);

/*
 * Patch extraction for the Queue path: gathers `n` patches of size (pi,pj,pk) with
 * corners `corners[n][3]` (device int32, voxel indices, corner + size <= shape —
 * validated by the caller) from one volume `src` (C,I,J,K) into a dense block `dst`
 * (n,C,pi,pj,pk).  `elem_bytes` in {1,2,4,8} (any dtype; bytes are moved verbatim).
 * Replaces PatchSampler._extract_patch's per-patch views (data/sampler.py:54-67,
 * 198-223) + the per-patch copies of collate_subjects' torch.stack
 * (loader.py:15-24) with one pass over the patch bytes.
 */
int tio_crop_patches(const void* src, void* dst, int elem_bytes, int C, int I, int J, int K,
                     int n, const int32_t* corners, int pi, int pj, int pk, void* stream);

/*
 * Flip / Crop / Pad as one index-remap copy of a (B,C,I,J,K) batch into (B,C,OI,OJ,OK):
 * source index along an axis = output index - off (off = voxels padded before; negative =
 * voxels cropped), indices outside the volume follow `mode` (0 constant -> `*fill`,
 * `elem_bytes` bytes on the HOST; 1 replicate; 2 reflect; 3 circular — F.pad's modes),
 * then the axis is reversed when the element's bit in `flip[b]` is set (bit 0 I, 1 J, 2 K;
 * device array or NULL).  Replaces torch.flip + torch.where (spatial/flip.py:233-263),
 * the crop slice (crop.py:84-101) and F.pad (_padding.py:73-104).  Any dtype by size.
 */
int tio_remap(const void* src, void* dst, int elem_bytes, int B, int C, int I, int J, int K,
              int OI, int OJ, int OK, int off_i, int off_j, int off_k, int mode,
              const void* fill, const uint8_t* flip, void* stream);

/*
 * Parameter-table upload without the copy engine: an SM kernel reads `bytes`
 * from page-locked host memory (`host_pinned`, a cudaHostAlloc/cudaHostRegister
 * pointer, device-visible under unified addressing) and writes them to
 * `dst_device`, ordered on `stream` like any launch.  The reference builds these
 * tables on the host and moves them with `.to(device)` inside each transform
 * (e.g. spatial.py:1548-1551, blur.py:292-328); when a batch is streamed through
 * the device in slices, such small cudaMemcpyAsync calls queue behind the bulk
 * volume copies on the copy engine and stall the kernels that need them.
 */
int tio_upload(const void* host_pinned, void* dst_device, size_t bytes, void* stream);

/*
 * Per-channel minimum of batch element 0 -> fill[C] on the device, no host
 * sync.  Replaces _batch_fill_value("minimum") = tensor.min().item()
 * (spatial.py:2054-2060, 2094-2095).  `src` is (B, C, n) fp32, n = I*J*K.
 */
int tio_min_sample0(const float* src, int C, int64_t n, float* fill, void* stream);

/*
 * K2 — bias field: dst = src * exp(trilerp_align_corners(coarse)) (or / for
 * the inverse).  Replaces _apply_bias_per_element / _generate_bias_field
 * (transforms/intensity/bias_field.py:201-255, 296-341).
 *   coarse    [B][C][si][sj][sk] fp32, drawn on the host by torch.normal from
 *             the recorded seeds (bias_field.py:281-293,316-329)
 *   identity  [B] bytes, non-zero = copy the row exactly (std == 0), or NULL
 * In-place (src == dst) allowed.
 */
int tio_bias_field(const float* src, float* dst, int B, int C, int I, int J, int K,
                   const float* coarse, int si, int sj, int sk,
                   const uint8_t* identity, int divide, void* stream);

/*
 * K3 — separable Gaussian blur with replicate (clamp) addressing, axes I, J,
 * K in that order.  Replaces _gaussian_smooth{,_shared,_per_element}
 * (transforms/intensity/blur.py:129-252) = 3 x (F.pad replicate + F.conv3d).
 *   taps      [3][B][2R+1] fp32, centred, normalised on the host exactly as
 *             blur.py:179-183 / 292-328 (zero beyond each element's radius,
 *             delta kernel where sigma <= 0)
 *   radius    [3][B] int32: the element's own radius on that axis (0 = skip)
 *   R         table half-width (max radius over the whole table)
 *   axes_mask host int, bit a set = axis a is active for at least one element
 *             (lets the library skip whole passes without reading `radius`)
 *   identity  [B] bytes: rows with all sigma <= 0 are copied exactly
 *   scratch   device buffer of B*C*I*J*K floats (may be NULL when only the I
 *             axis is active)
 * src and dst must not alias.
 */
int tio_blur(const float* src, float* dst, float* scratch,
             int B, int C, int I, int J, int K,
             const float* taps, const int32_t* radius, int R, int axes_mask,
             const uint8_t* identity, void* stream);

/*
 * K4a — exact replay of torch's CPU `randn` stream on the device:
 * mt19937(seed) -> 24-bit uniforms -> 16-wide Box-Muller blocks (ATen normal_fill;
 * the stream the reference depends on through torch.randn(generator=CPU),
 * noise.py:166-178).  Writes stream elements [offset, offset+n) to z[0..n).
 * Requires offset % 16 == 0, n % 16 == 0, n >= 16 (ragged tails and tiny draws
 * stay on the host), and offset + n <= 2^31 words.
 *   table      device copy of the jump-ahead table built once by
 *              tio_mt19937_build_table (host, ~2 s; depends only on MT19937, so
 *              callers cache it; tio_mt19937_table_bytes() gives its size)
 *   workspace  device scratch of tio_randn_mt19937_workspace_bytes(offset, n)
 * The uniforms are bit-identical to torch.s; normals agree to <= 4e-6 absolute
 * (CUDA libm vs the host.s log/sin/cos).
 */
size_t tio_mt19937_table_bytes(void);
int tio_mt19937_build_table(void* host_blob, size_t bytes);
size_t tio_randn_mt19937_workspace_bytes(uint64_t offset, uint64_t n);
int tio_randn_mt19937(uint64_t seed, uint64_t offset, uint64_t n, float* z,
                      const void* table, void* workspace, size_t workspace_bytes,
                      void* stream);

/*
 * K4 — additive Gaussian / Rician noise.  Replaces _sample_noise + add +
 * _restore_gated_out (transforms/intensity/noise.py:98-178).
 *   dst = src + (mean[b] + std[b] * z)                       (Gaussian)
 *   dst = sqrt((src + n1)^2 + n2^2), n_i = mean + std * z_i  (Rician)
 *   z, z2    standard normals, same shape as src (z2 NULL unless Rician); the
 *            reference draws them with torch.randn on a CPU mt19937 generator
 *   keep     [B] bytes or NULL; rows with keep == 0 are copied exactly
 * In-place allowed.
 */
int tio_noise(const float* src, float* dst, int B, int64_t per_elem,
              const float* mean, const float* std, const uint8_t* keep,
              const float* z, const float* z2, void* stream);

/*
 * K4b — same, with normals generated in registers from Philox4x32-7 keyed by
 * (seed, global element index): statistically equivalent, NOT the reference
 * stream.  `rician` selects the two-draw variant.
 */
int tio_noise_philox(const float* src, float* dst, int B, int64_t per_elem,
                     const float* mean, const float* std, const uint8_t* keep,
                     uint64_t seed, int rician, void* stream);

/*
 * K5 — gamma: dst = sign(src) * |src| ^ gamma[b].  Replaces
 * data.sign() * data.abs().pow(gamma)  (transforms/intensity/gamma.py:88-90).
 * In-place allowed.
 */
int tio_gamma(const float* src, float* dst, int B, int64_t per_elem,
              const float* gamma, void* stream);

/*
 * Data-derived parameters of Standardize / Normalize, computed where the batch lives
 * (the reference reads batch element 0 on the host: standardize.py:52-79, normalize.py:121-139,
 * 332-366, _statistics.py:11-45).
 *   tio_moments    out3 (device doubles) = {sum, sum of squares, count} of the `n` values at `src`
 *                  for which mask[t] != 0 (mask NULL = all); fp64 accumulation, one pass
 *   tio_quantiles  for each of the m <= 2 quantiles q (host doubles in [0,1]): index = q*(count-1),
 *                  lower = floor(index); values[2t], values[2t+1] = the order statistics of rank
 *                  lower and min(lower+1, count-1) (what torch.kthvalue(lower+1 / lower+2) returns),
 *                  weights[t] = index - lower, *count = number of selected values.  Exact (3-level
 *                  radix select over the order-preserving integer image of fp32), three passes.
 *                  workspace: tio_quantiles_workspace_bytes() device bytes, 16-byte aligned.
 *   tio_rescale    dst = ((clamp(src, lo, hi) - sub[b]) / div[b]) * mul[b] + add[b] over (B, per_elem),
 *                  each step rounded to fp32 like the reference's separate elementwise ops
 *                  (normalize.py:176-181, standardize.py:93; the inverses :271-297, :139-141).
 *                  `flags` selects the steps: 1 clamp, 2 sub, 4 div, 8 mul, 16 add (tables for
 *                  unselected steps may be NULL); keep[b] == 0 copies the row.  In-place allowed.
 */
int tio_moments(const float* src, const uint8_t* mask, int64_t n, double* out3, void* stream);
size_t tio_quantiles_workspace_bytes(void);
int tio_quantiles(const float* src, const uint8_t* mask, int64_t n, const double* q_host, int m,
                  float* values, double* weights, double* count, void* workspace,
                  size_t workspace_bytes, void* stream);
int tio_rescale(const float* src, float* dst, int B, int64_t per_elem, float lo, float hi,
                const float* sub, const float* div, const float* mul, const float* add,
                const uint8_t* keep, int flags, void* stream);

/*
 * Fused intensity chain: what Compose([BiasField, Blur, Noise, Gamma]) computes,
 * in two HBM passes.  Any stage may be absent (NULL table / noise_mode 0):
 *   v   = src * exp(trilerp(coarse))   (/ when bias_divide)  if coarse != NULL
 *   v   = blur_I(blur_J(blur_K(v)))                          if taps   != NULL
 *   v   = v + mean[b] + std[b] * n  (or Rician)               if noise_mode != 0
 *   dst = sign(v) |v|^gamma[b]                                if gamma  != NULL
 * Equal to running K2, K3, K4, K5 one after another up to fp32 summation order
 * (the separable passes commute; the reference order is I, J, K).
 *   noise_mode 1: normals supplied in z (z2 for the second Rician draw)
 *   noise_mode 2: Philox4x32-7 keyed by philox_seed (NOT the reference stream)
 *   per-element identity rows (bias_identity[b], all radii 0, keep[b] == 0,
 *   gamma[b] == 1) pass through every stage as bit-exact copies
 *   scratch: B*C*I*J*K floats, required when axes_mask has bit 1 or 2 (J/K)
 * src, dst, scratch must be distinct when blur is active.
 */
int tio_intensity_fused(const float* src, float* dst, float* scratch,
                        int B, int C, int I, int J, int K,
                        const float* coarse, int si, int sj, int sk,
                        const uint8_t* bias_identity, int bias_divide,
                        const float* taps, const int32_t* radius, int R, int axes_mask,
                        const float* mean, const float* std, const uint8_t* keep,
                        const float* z, const float* z2,
                        uint64_t philox_seed, int noise_mode, int rician,
                        const float* gamma, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TIO_B200_H */
