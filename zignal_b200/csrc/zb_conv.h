// zb_conv.h -- internal interface between the convolution translation units.
#pragma once
#include "zb_internal.h"

namespace zb {

constexpr int kMaxTaps = 1023;     // per separable axis (generic path)
constexpr int kMaxTaps2D = 1024;   // kh*kw (generic dense path)

// zb_conv_generic.cu
int conv_separable_generic(const zb_image* src, zb_image* dst, int pixfmt, const float* kx, int nx, const float* ky, int ny,
                           int border, cudaStream_t s);
int convolve_generic(const zb_image* src, zb_image* dst, int pixfmt, const float* kernel, int kh, int kw, int border, cudaStream_t s);

// zb_conv_fused.cu: single-pass (read once, write once) separable convolution of interleaved RGBA f32.
// Returns ZB_ERR_UNSUPPORTED when the configuration is outside the fused kernel's envelope; the caller
// then uses the generic path.  [row0, row1) restricts the OUTPUT rows produced (row1 < 0: all) -- the host pipeline
// computes a band as soon as the rows it reads have been uploaded.
int conv_separable_fused_rgbaf32(const zb_image* src, zb_image* dst, const float* kx, int nx, const float* ky, int ny, int border,
                                 bool exact, cudaStream_t s, int row0 = 0, int row1 = -1);

// zb_conv_fused_u8.cu: single-pass separable convolution of interleaved Rgba(u8) (i32 accumulators, provably overflow-free taps).
int conv_separable_fused_rgba8(const zb_image* src, zb_image* dst, const float* kx, int nx, const float* ky, int ny, int border,
                               cudaStream_t s, int row0 = 0, int row1 = -1);

// zb_conv_tile_u8.cu: single-pass (shared-memory tile) separable convolution of any 8-bit format / alignment / border mode.
int conv_separable_tile_u8(const zb_image* src, zb_image* dst, int channels, const float* kx, int nx, const float* ky, int ny, int border,
                           cudaStream_t s);
// Dense kernels up to 7 x 7 on 8-bit images from shared-memory tiles (zb_conv_tile_u8.cu); ZB_ERR_UNSUPPORTED outside its envelope.
int convolve_tile_u8(const zb_image* src, zb_image* dst, int channels, const int32_t* ki, int kh, int kw, int border, cudaStream_t s);
// Image.sobel of an 8-bit image (gray / Rgb / Rgba) from shared-memory byte tiles (zb_conv_tile_u8.cu); ZB_ERR_UNSUPPORTED outside its envelope.
int sobel_tile_u8(const zb_image* src, zb_image* dst, int channels, cudaStream_t s);

}  // namespace zb
