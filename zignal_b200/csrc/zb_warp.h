// zb_warp.h -- internal declarations shared by zb_resize.cu / zb_warp.cu.
#pragma once
#include "zb_internal.h"

namespace zb {
// Device copy of the reference's 1025-entry Lanczos3 LUT (interpolation.zig:256-267), built once per device on the host.
int lanczos_lut_device(const float** out, cudaStream_t s);
int resize_dispatch(const zb_image* src, zb_image* dst, int pixfmt, int method, float mb, float mc, cudaStream_t s);
}  // namespace zb
