/*
 * zignal_oracle.h -- C interface of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement (C++17, scalar, f32/f64/integer
 * arithmetic written out exactly as the reference does it) of the arrufat/zignal hot path that
 * zignal_b200 replaces.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load it; the product (zignal_b200/) never links, imports or calls it.
 *
 * Parity status: the reference is Zig (needs nightly >= 0.17.0-dev.1441, build.zig.zon:5) and
 * no Zig compiler exists in the build container, so the reference itself cannot be run here.
 * The oracle is pinned against every known-answer fact the reference's own tests hold for this
 * path (ported in tests/test_oracle_kat.py).  Those KATs pin semantics (borders,
 * rounding, fixed point, SAT order), not large-image numerics; see DESIGN.md "Oracle".
 *
 * All citations are file:line into /root/reference/src/.
 */
#ifndef ZIGNAL_ORACLE_H
#define ZIGNAL_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Image(T) mirror: image.zig:97-102.  stride is in PIXELS (not bytes). */
typedef struct zo_image {
    void*    data;
    uint32_t rows;
    uint32_t cols;
    uint64_t stride;
} zo_image;

/* BorderMode, border.zig:10-19 (enum order). */
enum { ZO_BORDER_ZERO = 0, ZO_BORDER_REPLICATE = 1, ZO_BORDER_MIRROR = 2, ZO_BORDER_WRAP = 3 };

/* Interpolation union tag order, interpolation.zig:53-68. */
enum {
    ZO_INTERP_NEAREST = 0, ZO_INTERP_BILINEAR = 1, ZO_INTERP_BICUBIC = 2,
    ZO_INTERP_CATMULL_ROM = 3, ZO_INTERP_MITCHELL = 4, ZO_INTERP_LANCZOS = 5
};

/* Pixel formats the hot path is exercised with. */
enum {
    ZO_PIX_U8 = 0,      /* Image(u8)                                  */
    ZO_PIX_F32 = 1,     /* Image(f32)                                 */
    ZO_PIX_RGB8 = 2,    /* Image(Rgb(u8)),  3 bytes r,g,b   color.zig:286 */
    ZO_PIX_RGBA8 = 3,   /* Image(Rgba(u8)), packed r,g,b,a  color.zig:400 */
    ZO_PIX_RGBAF32 = 4  /* 4 x f32 interleaved; the reference cannot convolve this type
                           (convolution.zig:432), so the oracle defines it per channel as
                           Image(f32) planes */
};

/* Transform kinds for warp (geometry/transforms.zig:10,118,197). */
enum { ZO_XFORM_SIMILARITY = 0, ZO_XFORM_AFFINE = 1, ZO_XFORM_PROJECTIVE = 2 };

/* SVD Mode, svd.zig:6-17. */
enum { ZO_SVD_NO_U = 0, ZO_SVD_SKINNY_U = 1, ZO_SVD_FULL_U = 2 };

/* Status codes (Zig error names). */
enum {
    ZO_OK = 0, ZO_ERR_DIMENSION_MISMATCH = 1, ZO_ERR_INVALID_SIGMA = 2, ZO_ERR_UNSUPPORTED = 3,
    ZO_ERR_NOT_CONVERGED = 4, ZO_ERR_INVALID_ARGUMENT = 5 /* also error.InvalidParameter */,
    ZO_ERR_INVALID_THRESHOLD = 14 /* edges.zig:225-226 */, ZO_ERR_INVALID_PERCENTILE = 15, ZO_ERR_INVALID_TRIM = 16 /* order_statistic_blur.zig:15-20 */,
    ZO_ERR_IMAGE_TOO_SMALL = 17 /* metrics.zig:61 */, ZO_ERR_NOT_SQUARE = 18, ZO_ERR_NOT_SYMMETRIC = 19, ZO_ERR_NOT_FINITE = 20 /* eigen.zig:34-54 */
};

/* Number of OpenMP threads the row-parallel loops may use (1 = the reference's behaviour). */
void zo_set_threads(int n);
int  zo_get_threads(void);
int  zo_hw_threads(void);
/* benchmark helper: thread-parallel first-touch fill of an f32 plane with uniform [0, 1) values */
void zo_parallel_fill_f32(float* p, uint64_t rows, uint64_t cols, uint64_t seed);

/* border.zig:46-63.  Returns -1 for "null" (contributes zero). */
int64_t zo_resolve_index(int64_t idx, int64_t length, int border);

/* meta.zig:110-135 clamp(u8, f32) and convolution.zig:18-22 divClampU8 (for KATs). */
uint8_t zo_clamp_u8_f32(float v);
uint8_t zo_div_clamp_u8(int64_t accum, int64_t scale);

/* image.zig:954-990: radius = ceil(3 sigma), taps exp(-x^2/(2 sigma^2)) in f32, normalised by a
 * sequential f32 sum.  Writes 2*radius+1 taps; returns the tap count, 0 for sigma == 0,
 * -1 for sigma < 0, -2 if cap is too small. */
int zo_gaussian_taps(float sigma, float* taps, int cap);

/* convolution.zig:313-438 (+ :441-647).  pixfmt in {U8, F32, RGB8, RGBA8, RGBAF32}. */
int zo_conv_separable(const zo_image* src, zo_image* dst, int pixfmt,
                      const float* kx, int nx, const float* ky, int ny, int border);

/* convolution.zig:198-301.  kernel is kh*kw row-major f32 (the reference's comptime array after
 * `as(f32, ...)`, :102).  pixfmt in {U8, F32, RGB8, RGBA8}. */
int zo_convolve(const zo_image* src, zo_image* dst, int pixfmt,
                const float* kernel, int kh, int kw, int border);

/* image.zig:954-994. */
int zo_gaussian_blur(const zo_image* src, zo_image* dst, int pixfmt, float sigma);

/* integral.zig:41-78: f32 summed-area table of one scalar plane (pixfmt U8 or F32). */
int zo_integral_plane(const zo_image* src, int pixfmt, float* sat /* rows*cols */);
/* image.zig:635-648 / integral.zig:148-269; pixfmt in {U8, F32, RGB8, RGBA8, RGBAF32}. */
int zo_box_blur(const zo_image* src, zo_image* dst, int pixfmt, uint32_t radius);
/* image.zig:999-1009 / edges.zig:33-73: Sobel magnitude into an Image(u8) (src: U8, F32, RGB8 or RGBA8). */
int zo_sobel(const zo_image* src, zo_image* dst_u8, int pixfmt);
/* image.zig:1041-1063 / edges.zig:212-274: Canny edges (0 / 255) into an Image(u8) (src: U8, F32, RGB8 or RGBA8). */
int zo_canny(const zo_image* src, zo_image* dst_u8, int pixfmt, float sigma, float low_threshold, float high_threshold);
/* image.zig:650-790 / order_statistic_blur.zig: percentileBlur (param = percentile in [0, 1]; medianBlur = 0.5 with .mirror, minBlur = 0,
 * maxBlur = 1), midpointBlur (param ignored), alphaTrimmedMeanBlur (param = trim fraction in [0, 0.5)).  pixfmt U8, RGB8 or RGBA8. */
enum { ZO_ORDER_PERCENTILE = 0, ZO_ORDER_MIDPOINT = 1, ZO_ORDER_ALPHA_TRIMMED = 2 };
int zo_order_blur(const zo_image* src, zo_image* dst, int pixfmt, uint32_t radius, int mode, double param, int border);
/* image.zig:1105-1147 / metrics.zig: f64 image-quality metrics of two images of the same pixel format (U8, F32, RGB8, RGBA8, RGBAF32). */
int zo_psnr(const zo_image* a, const zo_image* b, int pixfmt, double* out);
int zo_ssim(const zo_image* a, const zo_image* b, int pixfmt, double* out);
int zo_mean_pixel_error(const zo_image* a, const zo_image* b, int pixfmt, double* out);
void zo_ssim_window(double* w121);
/* motion_blur.zig:115-250: the diagonal branch of MotionBlur.linear (distance > 0; cos / sin of the angle are data) and :252-436 radial
 * zoom (spin = 0) / spin (spin = 1); center in normalised [0, 1] coordinates.  Any of the five pixel formats. */
int zo_motion_blur_line(const zo_image* src, zo_image* dst, int pixfmt, float cos_angle, float sin_angle, uint32_t distance);
int zo_motion_blur_radial(const zo_image* src, zo_image* dst, int pixfmt, float center_x, float center_y, float strength, int spin);
/* image.zig:396-421 Image.convert / convertInto: per pixel convertColor(Target, px) (color.zig:108-151) between any two of the five pixel
 * formats (the same format copies). */
int zo_convert(const zo_image* src, int src_pixfmt, zo_image* dst, int dst_pixfmt);
/* image.zig:785-799 / integral.zig:273-422. */
int zo_sharpen(const zo_image* src, zo_image* dst, int pixfmt, uint32_t radius);

/* interpolation.zig:72-84.  Samples one pixel; returns 1 and writes the pixel (pixel-size bytes)
 * or returns 0 for null. */
int zo_interpolate(const zo_image* src, int pixfmt, float x, float y, int method,
                   float mitchell_b, float mitchell_c, int border, void* out_pixel);
/* interpolation.zig:89-191 (plane path for RGB8/RGBA8, generic path otherwise). */
int zo_resize(const zo_image* src, zo_image* dst, int pixfmt, int method,
              float mitchell_b, float mitchell_c);

/* transforms.zig:112-149.  angle-only (uses cosf/sinf of this libm). */
void zo_rotate_bounds(uint32_t rows, uint32_t cols, float angle, uint32_t* out_rows, uint32_t* out_cols);
/* transforms.zig:163-212; fast-path class of `angle`: 0 general, 1 = 0deg, 2 = 90, 3 = 180, 4 = 270. */
int zo_rotate_class(float angle);
/* transforms.zig:163-212.  cos_a / sin_a are passed in (computed by the caller) so that the GPU
 * path and the oracle consume identical values; the orthogonal fast paths key off `angle`. */
int zo_rotate_into(const zo_image* src, zo_image* dst, int pixfmt, float angle, float cos_a, float sin_a,
                   int method, float mitchell_b, float mitchell_c, int border);
/* transforms.zig:232-283 (extract; rect = l, t, r, b in source coordinates, angle CCW around its centre; cos/sin passed in)
 * including the copyRect fast path :465-518.  crop (:216-222) = extract(rect, 0, nearest, zero) into round(height) x round(width). */
int zo_extract(const zo_image* src, zo_image* dst, int pixfmt, float rl, float rt, float rr, float rb, float angle, float cos_a,
               float sin_a, int method, float mitchell_b, float mitchell_c, int border);
/* transforms.zig:293-376 (insert) for blend_mode == .none and a source of the destination's pixel type: `self` is modified in place. */
int zo_insert(zo_image* self, const zo_image* source, int pixfmt, float rl, float rt, float rr, float rb, float angle, float cos_a,
              float sin_a, int method, float mitchell_b, float mitchell_c);
/* blending.zig:8-22 Blending (enum order) and :26-156 blendColors(u8, ...); insert with a blend mode blends Rgba(u8) samples into an
 * Rgba(u8) destination (image.zig:67-95 assignPixel) and assigns for every other pixel type. */
enum { ZO_BLEND_NONE = 0, ZO_BLEND_NORMAL, ZO_BLEND_MULTIPLY, ZO_BLEND_SCREEN, ZO_BLEND_OVERLAY, ZO_BLEND_SOFT_LIGHT, ZO_BLEND_HARD_LIGHT,
       ZO_BLEND_COLOR_DODGE, ZO_BLEND_COLOR_BURN, ZO_BLEND_DARKEN, ZO_BLEND_LIGHTEN, ZO_BLEND_DIFFERENCE, ZO_BLEND_EXCLUSION };
void zo_blend_rgba8(const uint8_t* base4, const uint8_t* overlay4, int mode, uint8_t* out4);
int zo_insert_blend(zo_image* self, const zo_image* source, int pixfmt, float rl, float rt, float rr, float rb, float angle, float cos_a,
                    float sin_a, int method, float mb, float mc, int blend);
/* transforms.zig:522-531 with project() of geometry/transforms.zig:39,147,224.
 * m is row-major: similarity/affine: {m00,m01,m10,m11,b0,b1}; projective: 9 values. */
int zo_warp(const zo_image* src, zo_image* dst, int pixfmt, int xform_kind, const float* m,
            int method, float mitchell_b, float mitchell_c);

/* svd.zig:149-496 on a row-major m x n matrix (m >= n).  u is m x (mode==FULL ? m : n) (ignored
 * for NO_U), s is n, v is n x n (only if with_v).  Returns `converged` (0 = ok, k = failed). */
/* matrix/eigen.zig:34-136 Matrix.eigh: row-major n x n symmetric input; values[n] ascending, vectors[n*n] with eigenvectors as columns. */
int zo_eigh_f64(const double* a, uint32_t rows, uint32_t cols, double* values, double* vectors);
int zo_eigh_f32(const float* a, uint32_t rows, uint32_t cols, float* values, float* vectors);
int64_t zo_svd_f64(const double* a, uint32_t m, uint32_t n, int mode, int with_v,
                   double* u, double* s, double* v);
int64_t zo_svd_f32(const float* a, uint32_t m, uint32_t n, int mode, int with_v,
                   float* u, float* s, float* v);

/* Matrix.zig:696-822: C(out) = alpha*op(A)*op(B) + beta*C.  c may be NULL.  vec_len is the host
 * vector length the reference would use (std.simd.suggestVectorLength): pass 0 for "this build". */
int zo_gemm_f32(const float* a, uint32_t a_rows, uint32_t a_cols, int trans_a,
                const float* b, uint32_t b_rows, uint32_t b_cols, int trans_b,
                float alpha, float beta, const float* c, float* out, int vec_len);
int zo_gemm_f64(const double* a, uint32_t a_rows, uint32_t a_cols, int trans_a,
                const double* b, uint32_t b_rows, uint32_t b_cols, int trans_b,
                double alpha, double beta, const double* c, double* out, int vec_len);

/* stats.zig:234-322 CovarianceStats(3,f64) streamed over an image the way fdm.zig:74-90 does.
 * out: mean[3], cov[9] (unbiased), is_gray. */
int zo_fdm_stats(const zo_image* img, int pixfmt, int as_luma, double* mean3, double* cov9, int* is_gray);
/* fdm.zig:133-273: match(source, target); source modified in place.  pixfmt in {U8, RGB8, RGBA8}. */
int zo_fdm_match(zo_image* source, const zo_image* target, int pixfmt);
/* The colour-branch map alone (fdm.zig:257-271) with a given W (3x3 row-major) and bias. */
int zo_fdm_apply(zo_image* source, int pixfmt, const double* w9, const double* bias3);

/* pca.zig:104-164,331-425.  data n x dim row-major.  components: dim x k, eigenvalues: k,
 * mean: dim.  Returns 0 or an error. */
int zo_pca_fit_f64(const double* data, uint32_t n, uint32_t dim, uint32_t k,
                   double* mean, double* components, double* eigenvalues);
int zo_pca_fit_f32(const float* data, uint32_t n, uint32_t dim, uint32_t k,
                   float* mean, float* components, float* eigenvalues);
/* pca.zig:291-312 */
int zo_pca_transform_f64(const double* data, uint32_t n, uint32_t dim, uint32_t k,
                         const double* mean, const double* components, double* out);
int zo_pca_transform_f32(const float* data, uint32_t n, uint32_t dim, uint32_t k,
                         const float* mean, const float* components, float* out);

#ifdef __cplusplus
}
#endif
#endif
