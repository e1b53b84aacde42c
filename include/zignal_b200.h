/*
 * zignal_b200.h -- C ABI of libzignal_b200.so: the B200 (sm_100a) implementation of zignal's
 * per-pixel image hot path.  This is the drop-in boundary: plain pointers and sizes, no C++ or
 * torch types.  A host shim in the reference's language (zig/zignal_b200.zig; C++ mirror in
 * zignal_b200/host/zignal.hpp; Python ctypes mirror in zignal_b200/) maps zignal's own
 * Image(T)/Matrix(T) method signatures onto these entry points.
 *
 * The reference has no FFI on this path (callers invoke monomorphised Zig generics), so each entry
 * point cites the reference *method* it replaces, file:line under arrufat/zignal src/.
 *
 * Conventions
 *  - zb_image mirrors Image(T) (image.zig:97-102): row-major, `stride` in PIXELS, views allowed.
 *  - `pixfmt` replaces the comptime pixel type T.
 *  - zb_* image ops take DEVICE pointers and enqueue on `stream` (a cudaStream_t, NULL = default
 *    stream); they return after enqueueing.  zb_host_* twins take HOST pointers, stage through the
 *    library's device scratch and return when the result is back in host memory (the literal drop-in
 *    for an `Image.data` that lives in host memory).
 *  - Every function returns a zb_status (0 = ok) named after the Zig error it maps to.
 *  - Transcendental inputs (Gaussian taps, cos/sin of a rotation angle, Lanczos LUT) are computed
 *    on the host and cross the boundary as data, so every consumer sees identical values.
 */
#ifndef ZIGNAL_B200_H
#define ZIGNAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZB_VERSION_MAJOR 0
#define ZB_VERSION_MINOR 1

typedef void* zb_stream; /* cudaStream_t */

/* Image(T): image.zig:97-102.  data: device pointer (zb_*) or host pointer (zb_host_*). */
typedef struct zb_image {
    void*    data;
    uint32_t rows;
    uint32_t cols;
    uint64_t stride; /* in pixels; == cols when contiguous (image.zig:355) */
} zb_image;

/* comptime pixel type T -> runtime tag */
typedef enum zb_pixfmt {
    ZB_PIX_U8 = 0,      /* Image(u8)                                                         */
    ZB_PIX_F32 = 1,     /* Image(f32)                                                        */
    ZB_PIX_RGB8 = 2,    /* Image(Rgb(u8)):  3 bytes r,g,b          color.zig:286-290         */
    ZB_PIX_RGBA8 = 3,   /* Image(Rgba(u8)): packed r,g,b,a         color.zig:400-405         */
    ZB_PIX_RGBAF32 = 4  /* Image(Rgba(f32)): 16-byte packed struct; filters act per channel  */
} zb_pixfmt;

/* BorderMode: border.zig:10-19 */
typedef enum zb_border { ZB_BORDER_ZERO = 0, ZB_BORDER_REPLICATE = 1, ZB_BORDER_MIRROR = 2, ZB_BORDER_WRAP = 3 } zb_border;

/* Interpolation: interpolation.zig:53-68 (union tag order; mitchell carries b, c) */
typedef enum zb_interp {
    ZB_INTERP_NEAREST = 0, ZB_INTERP_BILINEAR = 1, ZB_INTERP_BICUBIC = 2,
    ZB_INTERP_CATMULL_ROM = 3, ZB_INTERP_MITCHELL = 4, ZB_INTERP_LANCZOS = 5
} zb_interp;

/* geometry/transforms.zig:10,118,197 */
typedef enum zb_xform { ZB_XFORM_SIMILARITY = 0, ZB_XFORM_AFFINE = 1, ZB_XFORM_PROJECTIVE = 2 } zb_xform;

/* svd.zig:6-17 */
typedef enum zb_svd_mode { ZB_SVD_NO_U = 0, ZB_SVD_SKINNY_U = 1, ZB_SVD_FULL_U = 2 } zb_svd_mode;

/* Zig error names (image.zig:637,970,531-536; Matrix.zig:52-62; fdm.zig:114,142-143; pca.zig) */
typedef enum zb_status {
    ZB_OK = 0,
    ZB_ERR_DIMENSION_MISMATCH = 1, /* error.DimensionMismatch */
    ZB_ERR_INVALID_SIGMA = 2,      /* error.InvalidSigma      */
    ZB_ERR_UNSUPPORTED = 3,        /* (compile error in Zig: type not supported by this op) */
    ZB_ERR_NOT_CONVERGED = 4,      /* error.NotConverged / error.SvdFailed */
    ZB_ERR_INVALID_ARGUMENT = 5,
    ZB_ERR_OUT_OF_MEMORY = 6,      /* error.OutOfMemory       */
    ZB_ERR_DEVICE_FAILURE = 7,     /* CUDA / NCCL failure (new: error.DeviceFailure) */
    ZB_ERR_INVALID_SCALE_FACTOR = 8, /* error.InvalidScaleFactor (image.zig:531) */
    ZB_ERR_INVALID_DIMENSIONS = 9, /* error.InvalidDimensions (image.zig:536) */
    ZB_ERR_NO_TARGET_SET = 10,     /* fdm.zig:142 */
    ZB_ERR_NO_SOURCE_SET = 11,     /* fdm.zig:143 */
    ZB_ERR_INSUFFICIENT_DATA = 12, /* pca.zig:114-115 NoVectors / InsufficientData */
    ZB_ERR_INVALID_COMPONENTS = 13, /* pca.zig:123 */
    ZB_ERR_INVALID_THRESHOLD = 14, /* error.InvalidThreshold (edges.zig:225-226) */
    ZB_ERR_INVALID_PERCENTILE = 15, /* error.InvalidPercentile (order_statistic_blur.zig:49) */
    ZB_ERR_INVALID_TRIM = 16,      /* error.InvalidTrim (order_statistic_blur.zig:161) */
    ZB_ERR_IMAGE_TOO_SMALL = 17,   /* error.ImageTooSmall (metrics.zig:61) */
    ZB_ERR_NOT_SQUARE = 18,        /* error.NotSquare (eigen.zig:36) */
    ZB_ERR_NOT_SYMMETRIC = 19,     /* error.NotSymmetric (eigen.zig:54) */
    ZB_ERR_NOT_FINITE = 20         /* error.NotFinite (eigen.zig:49) */
} zb_status;

/* ------------------------------------------------------------------------------------------------
 * Runtime: device, streams, memory.  These back the Zig `std.mem.Allocator` vtable of the shim
 * (Image(T).init(dev_alloc, ...) / deinit, image.zig:124-158) and the pinned-host allocator.
 * ---------------------------------------------------------------------------------------------- */
int         zb_version(void);                 /* major*1000 + minor */
const char* zb_status_name(int status);
const char* zb_last_error(void);              /* thread-local text of the last CUDA failure */
int zb_device_count(int* count);
int zb_set_device(int ordinal);
int zb_get_device(int* ordinal);
int zb_sm_count(int* count);
int zb_stream_create(zb_stream* out);
int zb_stream_destroy(zb_stream s);
int zb_stream_synchronize(zb_stream s);
int zb_malloc(void** out, size_t bytes, zb_stream s);      /* cudaMallocAsync (stream-ordered pool) */
int zb_free(void* p, zb_stream s);                         /* cudaFreeAsync */
int zb_malloc_host(void** out, size_t bytes);              /* pinned host memory */
int zb_free_host(void* p);
/* Image-shaped copies (row pitch honoured on both sides); direction from the pointer kinds. */
/* Image.setBorder(rect, zeroes(T))   image.zig:198-229: zero everything outside [l, r) x [t, b) (the rect is clipped to the
 * image; an empty intersection zeroes the whole image).  Four strided memsets on the stream. */
int zb_set_border_zero(zb_image* img, int pixfmt, uint32_t l, uint32_t t, uint32_t r, uint32_t b, zb_stream s);
int zb_upload(const zb_image* host_src, zb_image* dev_dst, int pixfmt, zb_stream s);
int zb_download(const zb_image* dev_src, zb_image* host_dst, int pixfmt, zb_stream s);
int zb_copy(const zb_image* dev_src, zb_image* dev_dst, int pixfmt, zb_stream s); /* Image.copy, image.zig:375-392 */
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t zb_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Filters
 * ---------------------------------------------------------------------------------------------- */
/* Host helper, image.zig:972-990: radius = ceil(3 sigma); taps exp(-x^2/(2 sigma^2)) in f32,
 * normalised by a sequential f32 sum.  Returns the tap count in *n (0 for sigma == 0). */
int zb_gaussian_taps(float sigma, float* taps, int cap, int* n);

/* Image.convolveSeparable(out, allocator, kernel_x, kernel_y, border)   image.zig:935-951,
 * convolution.zig:313-438.  kx/ky are HOST arrays.  In-place (src->data == dst->data) is allowed. */
int zb_conv_separable(const zb_image* src, zb_image* dst, int pixfmt,
                      const float* kx, int nx, const float* ky, int ny, int border, zb_stream s);
/* Sharding extension (no reference counterpart): the same operation restricted to the OUTPUT rows
 * [row_begin, row_end) of dst.  src is still the whole image, so border handling and every value are identical to
 * the full call.  A row-block owner convolves the rows that do not depend on its halo while the halo exchange is in
 * flight, then the few boundary rows.  row_end == 0 means "to the last row". */
int zb_conv_separable_rows(const zb_image* src, zb_image* dst, int pixfmt,
                           const float* kx, int nx, const float* ky, int ny, int border,
                           uint32_t row_begin, uint32_t row_end, zb_stream s);
/* Image.convolve(out, allocator, kernel, border)   image.zig:917-931, convolution.zig:198-301.
 * kernel: HOST kh*kw row-major f32 (the comptime 2-D array after `as(f32, .)`). */
int zb_convolve(const zb_image* src, zb_image* dst, int pixfmt,
                const float* kernel, int kh, int kw, int border, zb_stream s);
/* Image.gaussianBlur(out, allocator, sigma)   image.zig:954-994 (always .mirror). */
int zb_gaussian_blur(const zb_image* src, zb_image* dst, int pixfmt, float sigma, zb_stream s);
/* Image.boxBlur(out, allocator, radius)   image.zig:635-648, integral.zig:148-269. */
int zb_box_blur(const zb_image* src, zb_image* dst, int pixfmt, uint32_t radius, zb_stream s);
/* Image.sharpen(out, allocator, radius)   image.zig:785-799, integral.zig:273-422. */
int zb_sharpen(const zb_image* src, zb_image* dst, int pixfmt, uint32_t radius, zb_stream s);
/* Image.integral: one f32 summed-area plane of a scalar image (integral.zig:41-78). sat: device rows*cols f32. */
int zb_integral_plane(const zb_image* src, int pixfmt, float* sat, zb_stream s);

/* ------------------------------------------------------------------------------------------------
 * Resampling and geometry
 * ---------------------------------------------------------------------------------------------- */
/* Image.resize(out, allocator, method)   image.zig:523-525, interpolation.zig:89-191. */
int zb_resize(const zb_image* src, zb_image* dst, int pixfmt, int method, float mitchell_b, float mitchell_c, zb_stream s);
/* Image.rotateBounds(angle)   image.zig:570, transforms.zig:112-149 (host math). */
int zb_rotate_bounds(uint32_t rows, uint32_t cols, float angle, uint32_t* out_rows, uint32_t* out_cols);
/* Image.rotateInto(out, angle, method, border)   image.zig:564, transforms.zig:163-212. */
int zb_rotate_into(const zb_image* src, zb_image* dst, int pixfmt, float angle,
                   int method, float mitchell_b, float mitchell_c, int border, zb_stream s);
/* Same, with cos(angle) / sin(angle) supplied by the caller (transforms.zig:190-191 computes them with
 * Zig's @cos/@sin; passing them as data makes every consumer use identical values).  `angle` still
 * selects the orthogonal fast paths (transforms.zig:165-187). */
int zb_rotate_into_cs(const zb_image* src, zb_image* dst, int pixfmt, float angle, float cos_a, float sin_a,
                      int method, float mitchell_b, float mitchell_c, int border, zb_stream s);
/* Batched rotateInto: n images of identical shape stored back to back (image i starts
 * i*image_pitch_px pixels after image 0).  Same arithmetic per image as zb_rotate_into_cs. */
int zb_rotate_into_batch(const zb_image* src0, uint64_t src_image_pitch_px, zb_image* dst0, uint64_t dst_image_pitch_px,
                         uint32_t n_images, int pixfmt, float angle, float cos_a, float sin_a,
                         int method, float mitchell_b, float mitchell_c, int border, zb_stream s);
/* Image.warp(out, transform, method)   image.zig:621, transforms.zig:522-531 with
 * {Similarity,Affine,Projective}Transform.project  geometry/transforms.zig:39,147,224.
 * m (HOST): similarity/affine {m00,m01,m10,m11,b0,b1}; projective 9 values row-major. */
int zb_warp(const zb_image* src, zb_image* dst, int pixfmt, int xform_kind, const float* m,
            int method, float mitchell_b, float mitchell_c, zb_stream s);

/* Image.sobel(out, allocator)   image.zig:999-1009, edges.zig:33-73: gradient magnitude of the luma into an Image(u8)
 * (src: U8, F32, RGB8 or RGBA8; dst is always an 8-bit gray image of the same shape). */
int zb_sobel(const zb_image* src, zb_image* dst_u8, int pixfmt, zb_stream s);
/* Image.canny(out, allocator, sigma, low_threshold, high_threshold)   image.zig:1041-1063, edges.zig:212-274: luma -> Gaussian
 * (.replicate) -> Sobel -> non-maximum suppression -> double threshold + hysteresis, all on the device; dst receives 0 / 255.
 * Non-finite parameter: ZB_ERR_INVALID_ARGUMENT (error.InvalidParameter); sigma < 0: ZB_ERR_INVALID_SIGMA; negative thresholds
 * or low >= high: ZB_ERR_INVALID_THRESHOLD.  The hysteresis iterates to a fixed point, so this call waits for the stream. */
int zb_canny(const zb_image* src, zb_image* dst_u8, int pixfmt, float sigma, float low_threshold, float high_threshold, zb_stream s);

/* Order-statistic filters   image.zig:650-790, image/order_statistic_blur.zig:22-197 on 8-bit images (U8, RGB8, RGBA8; every channel
 * independently, :199-229).  One entry point for the family:
 *   Image.percentileBlur(out, radius, percentile, border)   mode ZB_ORDER_PERCENTILE, param = percentile in [0, 1]
 *   Image.medianBlur(out, radius)                            = percentile 0.5 with ZB_BORDER_MIRROR (:22-29)
 *   Image.minBlur / maxBlur(out, radius, border)             = percentile 0.0 / 1.0 (:83-101)
 *   Image.midpointBlur(out, radius, border)                  mode ZB_ORDER_MIDPOINT (param ignored)
 *   Image.alphaTrimmedMeanBlur(out, radius, trim, border)    mode ZB_ORDER_ALPHA_TRIMMED, param = trim fraction in [0, 0.5)
 * radius 0 copies; src may alias dst.  Errors in the reference's order: DimensionMismatch, (empty image: ok), InvalidTrim, (radius 0:
 * copy), InvalidPercentile, Unsupported (pixel type; also a window that no longer fits one SM's shared memory: radius beyond ~100 for
 * Rgba, ~220 for gray). */
enum { ZB_ORDER_PERCENTILE = 0, ZB_ORDER_MIDPOINT = 1, ZB_ORDER_ALPHA_TRIMMED = 2 };
int zb_order_blur(const zb_image* src, zb_image* dst, int pixfmt, uint32_t radius, int mode, double param, int border, zb_stream s);

/* Image.motionBlur(out, allocator, .{ .linear = .{ .angle, .distance } })   image/motion_blur.zig:65-250.  cos / sin of the angle cross the
 * ABI as data (like rotateInto).  distance 0 copies; |sin| < 0.001 or |cos| < 0.001 is the reference's convolveSeparable branch (uniform
 * kernel, .replicate: the fused / tile convolution kernels); any other angle is the per-pixel line integral of bilinear samples
 * (src must not alias dst there). */
int zb_motion_blur_linear(const zb_image* src, zb_image* dst, int pixfmt, float angle, float cos_a, float sin_a, uint32_t distance, zb_stream s);
/* Image.motionBlur(..., .{ .radial_zoom | .radial_spin = .{ .center_x, .center_y, .strength } })   motion_blur.zig:252-436; spin != 0 selects
 * radial_spin.  Zoom is bit-exact; spin evaluates atan2 / cos / sin per sample on the device (last-bit libm differences: 1e-5 relative,
 * at most one 8-bit step on isolated pixels; on the outermost pixel ring the reference's own bounds test sits within an ulp of the image
 * edge, so a sample may be kept by one libm and dropped by another). */
int zb_motion_blur_radial(const zb_image* src, zb_image* dst, int pixfmt, float center_x, float center_y, float strength, int spin, zb_stream s);

/* Image.convert(allocator, TargetType) / convertInto(TargetType, out)   image.zig:396-421: out[r, c] = convertColor(TargetType, self[r, c])
 * (color.zig:108-151) between any two of the five pixel formats (equal formats copy, :398).  Bit-identical to the reference: integer BT.709
 * luma for 8-bit colours, separately rounded f32 arithmetic for float ones, u8 <-> float components as v / 255 and
 * round(255 * clamp(v, 0, 1)). */
int zb_convert(const zb_image* src, int src_pixfmt, zb_image* dst, int dst_pixfmt, zb_stream s);

/* Image.psnr(other) / Image.ssim(other) / Image.meanPixelError(other)   image.zig:1105-1147, image/metrics.zig:10-165: f64 quality
 * metrics of two device images of the same pixel format (U8, F32, RGB8, RGBA8, RGBAF32); *out is a HOST double, so each call waits
 * for the stream.  8-bit psnr / meanPixelError reproduce the reference's value exactly (integer sums); float formats and ssim
 * associate the final f64 sum differently (relative difference ~1e-15).  ssim: images under 11x11 give ZB_ERR_IMAGE_TOO_SMALL. */
int zb_psnr(const zb_image* a, const zb_image* b, int pixfmt, double* out, zb_stream s);
int zb_ssim(const zb_image* a, const zb_image* b, int pixfmt, double* out, zb_stream s);
int zb_mean_pixel_error(const zb_image* a, const zb_image* b, int pixfmt, double* out, zb_stream s);

/* Image.extract(out, rect, angle, method, border)   image.zig / transforms.zig:232-283: resample the rectangle (l, t, r, b in source
 * coordinates, rotated by `angle` CCW around its centre; cos/sin cross the ABI as data like rotateInto) into dst; an axis-aligned
 * rect of dst's own size takes the copyRect path (:465-518).  Image.crop(rect) (:216-222) is
 * zb_extract(rect, 0, cos 1, sin 0, NEAREST, ZERO) into a round(height) x round(width) image. */
int zb_extract(const zb_image* src, zb_image* dst, int pixfmt, float rect_l, float rect_t, float rect_r, float rect_b, float angle,
               float cos_a, float sin_a, int method, float mitchell_b, float mitchell_c, int border, zb_stream s);

/* Image.insert(source, rect, angle, method, .none)   transforms.zig:293-376, the complement of extract: `source` (same pixel type as
 * `self`) is resampled into the rotated rectangle of `self`; pixels outside the rectangle are not touched.  (Blend modes:
 * zb_insert_blend; a source of another pixel type: zb_insert_from.) */
int zb_insert(zb_image* self, const zb_image* source, int pixfmt, float rect_l, float rect_t, float rect_r, float rect_b, float angle,
              float cos_a, float sin_a, int method, float mitchell_b, float mitchell_c, zb_stream s);
/* Image.insert(source, rect, angle, method, blend_mode) with a Blending mode (blending.zig:8-22, enum order): Rgba(u8) samples are
 * composited onto the Rgba(u8) destination by blendColors(u8, dest, sample, mode) (blending.zig:26-156, f32 arithmetic, Porter-Duff
 * "over" for the alpha); for every other pixel type the sample is assigned, as image.zig:67-95 assignPixel does. */
enum { ZB_BLEND_NONE = 0, ZB_BLEND_NORMAL, ZB_BLEND_MULTIPLY, ZB_BLEND_SCREEN, ZB_BLEND_OVERLAY, ZB_BLEND_SOFT_LIGHT, ZB_BLEND_HARD_LIGHT,
       ZB_BLEND_COLOR_DODGE, ZB_BLEND_COLOR_BURN, ZB_BLEND_DARKEN, ZB_BLEND_LIGHTEN, ZB_BLEND_DIFFERENCE, ZB_BLEND_EXCLUSION };
int zb_insert_blend(zb_image* self, const zb_image* source, int pixfmt, float rect_l, float rect_t, float rect_r, float rect_b, float angle,
                    float cos_a, float sin_a, int method, float mitchell_b, float mitchell_c, int blend_mode, zb_stream s);
/* Image.insert(source: anytype, ...) with a source of ANOTHER pixel type (transforms.zig:293, image.zig:67-95): samples are taken in the
 * source's type and converted per pixel with convertColor(DestType, sample); an Rgba(u8) source with a blend mode composites through
 * Rgba(u8): dest = convertColor(DestType, blend(convertColor(Rgba, dest), sample)).  Equal pixel types forward to zb_insert_blend. */
int zb_insert_from(zb_image* self, int self_pixfmt, const zb_image* source, int source_pixfmt, float rect_l, float rect_t, float rect_r,
                   float rect_b, float angle, float cos_a, float sin_a, int method, float mitchell_b, float mitchell_c, int blend_mode,
                   zb_stream s);

/* ------------------------------------------------------------------------------------------------
 * Linear algebra behind fdm / pca
 * ---------------------------------------------------------------------------------------------- */
/* Matrix.gemm(trans_a, other, trans_b, alpha, beta, c)   Matrix.zig:696-822.
 * All matrices DEVICE, row-major, contiguous.  c may be NULL.  out is a_rows x b_cols. */
int zb_gemm_f32(const float* a, uint32_t a_rows, uint32_t a_cols, int trans_a,
                const float* b, uint32_t b_rows, uint32_t b_cols, int trans_b,
                float alpha, float beta, const float* c, float* out, zb_stream s);
int zb_gemm_f64(const double* a, uint32_t a_rows, uint32_t a_cols, int trans_a,
                const double* b, uint32_t b_rows, uint32_t b_cols, int trans_b,
                double alpha, double beta, const double* c, double* out, zb_stream s);
/* Pca.fit's centering step (pca.zig:135-154): mean[j] = sum_i x[i][j] / n (when compute_mean != 0; f64 accumulation),
 * centered = x - mean (skipped when NULL).  Pca.transform (pca.zig:300-308) calls it with compute_mean = 0.  DEVICE pointers. */
int zb_center_columns_f32(const float* x, uint32_t n, uint32_t dim, float* mean, int compute_mean, float* centered, zb_stream s);
int zb_center_columns_f64(const double* x, uint32_t n, uint32_t dim, double* mean, int compute_mean, double* centered, zb_stream s);
/* Matrix.svd / SMatrix.svd   Matrix.zig:1570, SMatrix.zig:804, svd.zig:80-496.  HOST matrices.
 * a: m x n row-major, m >= n.  u: m x (mode==FULL ? m : n) or NULL; s: n (descending); v: n x n or NULL.
 * *converged receives 0 or a non-zero failure mark (svd.zig:79).
 * Computed by a one-sided Jacobi method with a parallel pair ordering on the GPU (n >= 24; smaller matrices run the same
 * algorithm on the host), not by the reference's sequential Golub-Reinsch: singular values agree to sqrt(eps) (to high relative
 * accuracy, in fact), U and V are orthonormal and reconstruct A, but the SIGN of each (u_i, v_i) pair and the basis inside a
 * cluster of equal singular values are the decomposition's usual freedom and may differ from the reference's. */
int zb_svd_f64(const double* a, uint32_t m, uint32_t n, int mode, int with_v, double* u, double* s, double* v, uint64_t* converged);
int zb_svd_f32(const float* a, uint32_t m, uint32_t n, int mode, int with_v, float* u, float* s, float* v, uint64_t* converged);
/* The same decomposition for a matrix that already lives on the device (Pca.fit: the covariance from zb_gemm never leaves the
 * GPU, pca.zig:331-425).  d_a: m x n row-major DEVICE (not modified); d_u: m x n skinny U or NULL; d_s: n; d_v: n x n or NULL --
 * all DEVICE.  Returns after the stream has finished. */
int zb_svd_dev_f64(const double* d_a, uint32_t m, uint32_t n, double* d_u, double* d_s, double* d_v, uint64_t* converged, zb_stream s);
int zb_svd_dev_f32(const float* d_a, uint32_t m, uint32_t n, float* d_u, float* d_s, float* d_v, uint64_t* converged, zb_stream s);
/* Matrix.eigh(allocator)   matrix/eigen.zig:34-136: symmetric eigendecomposition (HOST pointers, row-major n x n).  values[n] ascending,
 * vectors[n * n] with the matching unit eigenvectors as columns.  ZB_ERR_NOT_SQUARE, ZB_ERR_NOT_FINITE (NaN / inf entry),
 * ZB_ERR_NOT_SYMMETRIC (|a_ij - a_ji| > max|a| * sqrt(eps)), in the reference's order.  Two-sided Jacobi with a parallel pair
 * ordering on the GPU (n >= 24, host below): same eigenvalues to eps * |A|, eigenvectors up to sign. */
int zb_eigh_f64(const double* a, uint32_t rows, uint32_t cols, double* values, double* vectors);
int zb_eigh_f32(const float* a, uint32_t rows, uint32_t cols, float* values, float* vectors);

/* ------------------------------------------------------------------------------------------------
 * Feature distribution matching   fdm.zig:19-275
 * ---------------------------------------------------------------------------------------------- */
typedef struct zb_fdm zb_fdm;                       /* FeatureDistributionMatching(T) state, fdm.zig:22-39 */
int zb_fdm_create(zb_fdm** out, int pixfmt);        /* .init   fdm.zig:42 ; pixfmt in {U8, RGB8, RGBA8} (fdm.zig:20) */
int zb_fdm_destroy(zb_fdm* f);                      /* .deinit fdm.zig:61 */
int zb_fdm_set_target(zb_fdm* f, const zb_image* target, zb_stream s);  /* fdm.zig:68  (device image, contiguous) */
int zb_fdm_set_source(zb_fdm* f, zb_image* source);                      /* fdm.zig:127 */
int zb_fdm_update(zb_fdm* f, zb_stream s);                               /* fdm.zig:141 (in place on source) */
int zb_fdm_match(zb_fdm* f, zb_image* source, const zb_image* target, zb_stream s); /* fdm.zig:133 */
/* The statistics pass alone: exact integer moment sums of a u8 image (n, sum x_i, sum x_i x_j, is_gray)
 * -- the quantities one all-reduce combines across GPUs.  sums: HOST u64[11] =
 * {n, Sr, Sg, Sb, Srr, Srg, Srb, Sgg, Sgb, Sbb, non_gray_count}.  as_luma selects fdm.zig:157-162. */
int zb_fdm_moments(const zb_image* img, int pixfmt, int as_luma, uint64_t* sums11, zb_stream s);
/* Install externally reduced moments (multi-GPU: after the all-reduce) as target / source statistics. */
int zb_fdm_set_target_moments(zb_fdm* f, const uint64_t* sums11);
int zb_fdm_update_with_moments(zb_fdm* f, const uint64_t* source_sums11, zb_stream s);
/* zb_fdm_update queues its three kernels (moments, the 3x3 solve, the map) on the stream and returns; the solve's result
 * (0, or ZB_ERR_NOT_CONVERGED = error.SvdFailed, fdm.zig:216) stays on the device.  This call waits for the stream and
 * returns it. */
int zb_fdm_status(zb_fdm* f, zb_stream s);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, a large image sharded into row blocks.
 * The reference is single-process, so there is no reference counterpart: these entry points are what a
 * Zig host running one process per GPU binds to shard Image.convolveSeparable / gaussianBlur (image.zig:935-994),
 * any other neighbourhood filter, fdm.update (fdm.zig:141-273) and batches of rotate / resize.
 *
 * Plumbing: NCCL (dlopen'ed at run time) bootstraps the communicator and exchanges CUDA IPC handles; the data
 * path is NVLink peer memory.  The RGBA f32 convolution kernel copies the `half` edge rows of the row
 * neighbours from their memory into its own halo rows in a prologue -- one launch per step, no separate exchange
 * -- and carries the whole synchronisation in a pair of flags per neighbour (blocks need halo_cap >= 8 for this
 * kernel); every other filter uses zb_shard_halo_exchange (one pull
 * kernel over NVLink; NCCL send/recv when IPC mappings are unavailable) followed by its ordinary entry point
 * on the extended block.  All ranks must issue the same sequence of zb_shard_* calls (SPMD).
 * ---------------------------------------------------------------------------------------------- */
typedef struct zb_shard_comm zb_shard_comm;     /* one per process / GPU */
typedef struct zb_shard_image zb_shard_image;   /* a global image = `world` row blocks stacked in rank order */
#define ZB_SHARD_ID_BYTES 128
/* Rank 0 creates the id and hands it to the other ranks out of band (the host's own launcher); then every rank
 * creates its communicator on its current device.  world == 1 needs no id (and no NCCL). */
int zb_shard_unique_id(uint8_t* id128);
int zb_shard_comm_create(zb_shard_comm** out, int rank, int world, const uint8_t* id128);
int zb_shard_comm_destroy(zb_shard_comm* c);
int zb_shard_comm_info(const zb_shard_comm* c, int* rank, int* world, int* peer_access /* 1: IPC peer mappings in use */);
/* Waits for the stream; ZB_ERR_DEVICE_FAILURE if a kernel gave up waiting for a neighbour (a rank died or the ranks'
 * call sequences diverged) instead of hanging the GPU. */
int zb_shard_status(zb_shard_comm* c, zb_stream s);
/* Symmetric allocation (collective, same order and size on every rank): device memory every rank can address. */
int zb_shard_alloc(zb_shard_comm* c, size_t bytes, void** out);
int zb_shard_free(zb_shard_comm* c, void* p);
/* Describe this rank's row block (inside a zb_shard_alloc allocation, with `halo_cap` spare rows of the same stride above
 * and below it) of the global image (collective: the ranks learn each other's block heights and addresses). */
int zb_shard_image_create(zb_shard_comm* c, const zb_image* block, uint32_t halo_cap, int pixfmt, zb_shard_image** out);
int zb_shard_image_destroy(zb_shard_image* img);
int zb_shard_image_block(const zb_shard_image* img, zb_image* block);   /* this rank's block as a plain zb_image */
/* Contiguous share [lo, hi) of n_items for `rank` (batches of rotate / resize / warp: no exchange).  Pure host arithmetic. */
int zb_shard_split(uint32_t n_items, int rank, int world, uint32_t* lo, uint32_t* hi);
/* Fill the `reach` halo rows next to each row neighbour with the neighbour's edge rows (global edges are left to the filter's own
 * border mode; ZB_BORDER_WRAP closes the ring).  One kernel on `s`; when it completes, no neighbour is still reading this block. */
int zb_shard_halo_exchange(zb_shard_comm* c, zb_shard_image* img, uint32_t reach, int border, zb_stream s);
/* The block plus the halo rows that zb_shard_halo_exchange filled, as a zb_image for any zb_* filter; *interior_first = index of
 * the block's first row inside the view.  Rows of the result within `reach` of a neighbour-side end are halo rows (discard). */
int zb_shard_view(const zb_shard_image* img, uint32_t reach, int border, zb_image* view, uint32_t* interior_first);
/* Image.convolveSeparable / gaussianBlur of the GLOBAL image; src / dst: the same partition.  Bit-identical to the single-GPU
 * call on the whole image. */
int zb_shard_conv_separable(zb_shard_comm* c, const zb_shard_image* src, zb_shard_image* dst,
                            const float* kx, int nx, const float* ky, int ny, int border, zb_stream s);
int zb_shard_gaussian_blur(zb_shard_comm* c, const zb_shard_image* src, zb_shard_image* dst, float sigma, zb_stream s);
/* In-place sum over all ranks of a device buffer (dtype 0 = f32, 1 = f64, 2 = u64): PCA's partial X^T X products (pca.zig:338),
 * moment sums.  NCCL all-reduce. */
int zb_shard_allreduce(zb_shard_comm* c, void* dev_buf, size_t count, int dtype, zb_stream s);
/* fdm on a row-sharded image: every rank passes its block.  set_target: local moments + one all-reduce; update (zb_fdm_set_source
 * with the local block first): the statistics kernel all-gathers the 11 sums over NVLink peer memory in its last block, solves and
 * the map follows -- nothing leaves the device.  The result equals the single-GPU result on the whole image bit for bit. */
int zb_shard_fdm_set_target(zb_shard_comm* c, zb_fdm* f, const zb_image* target_block, zb_stream s);
int zb_shard_fdm_update(zb_shard_comm* c, zb_fdm* f, zb_stream s);
/* Testing aid: 0 = automatic, 1 = force the NCCL send/recv halo exchange, 2 = force the peer pull kernel for every format. */
int zb_shard_tune_path(int path);
/* Introspection: %globaltimer stamps (ns) of this rank's last zb_shard_conv_separable kernel on the fused path -- [0] first CTA
 * started, [1] both neighbours' sources seen complete, [2] all neighbour rows copied into the halo, [3] last CTA finished its
 * rows, [4] both neighbours seen done reading this block's edge rows.  Waits for the stream. */
int zb_shard_debug_times(zb_shard_comm* c, uint64_t* out8, zb_stream s);

/* ------------------------------------------------------------------------------------------------
 * Host-pointer twins (H2D + op + D2H inside the call; returns when dst is valid on the host).
 * ---------------------------------------------------------------------------------------------- */
int zb_host_conv_separable(const zb_image* src, zb_image* dst, int pixfmt,
                           const float* kx, int nx, const float* ky, int ny, int border);
int zb_host_convolve(const zb_image* src, zb_image* dst, int pixfmt, const float* kernel, int kh, int kw, int border);
int zb_host_gaussian_blur(const zb_image* src, zb_image* dst, int pixfmt, float sigma);
int zb_host_box_blur(const zb_image* src, zb_image* dst, int pixfmt, uint32_t radius);
int zb_host_sharpen(const zb_image* src, zb_image* dst, int pixfmt, uint32_t radius);
int zb_host_resize(const zb_image* src, zb_image* dst, int pixfmt, int method, float mitchell_b, float mitchell_c);
int zb_host_rotate_into(const zb_image* src, zb_image* dst, int pixfmt, float angle, int method,
                        float mitchell_b, float mitchell_c, int border);
int zb_host_warp(const zb_image* src, zb_image* dst, int pixfmt, int xform_kind, const float* m,
                 int method, float mitchell_b, float mitchell_c);
int zb_host_fdm_match(zb_image* source, const zb_image* target, int pixfmt);

/* ------------------------------------------------------------------------------------------------
 * Tuning / introspection (not part of the drop-in surface)
 * ---------------------------------------------------------------------------------------------- */
/* Selects the arithmetic of the fused f32 separable kernel: 0 = FFMA (default, <= 1e-6 rel. from the
 * reference), 1 = unfused mul+add in the reference's order (bit-exact with it for finite data). */
int zb_set_exact_f32(int on);
/* Forces the generic (two-pass through HBM) separable path; used by tests to cross-check kernels. */
int zb_set_force_generic(int on);
/* Kernel tuning knobs for experiments ("conv.stages" 2|3, "conv.f32x2" 0|1, "conv.band_rows" >= 64,
 * "conv.variant" -1 auto | 0 phase-synchronous | 1 warp-specialised, "conv.u8_fmath" 0|1,
 * "host.band_rows": rows per PCIe band of the pipelined host-pointer path, 0 = stage the whole image,
 * "conv.u8_dp" 0|1, "conv.edge_fast" 0|1 (x borders of .replicate / .mirror as in-stage copies), "sobel.tile" 0|1, "jacobi.cluster" 0|1 (SVD inside one cluster's shared memory when it fits), "rotate.tile" 0 gather kernel |
 * 1 shared-memory tile kernel where it applies). */
int zb_tune(const char* key, int value);
/* Name of the kernel variant the last zb_conv_separable call selected on this thread. */
const char* zb_last_kernel(void);
/* Jacobi sweeps the last zb_svd_* / zb_eigh_* call on this thread needed (60 = the limit: not converged). */
int zb_last_sweeps(void);

#ifdef __cplusplus
}
#endif
#endif /* ZIGNAL_B200_H */
