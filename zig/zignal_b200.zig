//! zignal_b200.zig -- the Zig side of the drop-in boundary: `extern fn` declarations of the C ABI in
//! include/zignal_b200.h, a `std.mem.Allocator` over device memory, and `Image(T)`-shaped wrappers whose
//! bodies replace the reference's CPU loops (src/image.zig:635-648, 785-799, 917-994, 523-541, 558-623)
//! with calls into libzignal_b200.so.
//!
//! NOTE: no Zig toolchain exists in the build image (zignal needs nightly >= 0.17.0-dev.1441), so this
//! file is reviewed against the reference's signatures but not compiled here.  The compiled and tested
//! boundary is the C ABI itself; see INTEGRATION.md for how a maintainer wires this file in.

const std = @import("std");
const Allocator = std.mem.Allocator;

pub const c = struct {
    pub const Stream = ?*anyopaque;
    pub const ZbImage = extern struct { data: ?*anyopaque, rows: u32, cols: u32, stride: u64 };

    pub extern fn zb_status_name(status: c_int) [*:0]const u8;
    pub extern fn zb_malloc(out: *?*anyopaque, bytes: usize, s: Stream) c_int;
    pub extern fn zb_free(p: ?*anyopaque, s: Stream) c_int;
    pub extern fn zb_malloc_host(out: *?*anyopaque, bytes: usize) c_int;
    pub extern fn zb_free_host(p: ?*anyopaque) c_int;
    pub extern fn zb_upload(host_src: *const ZbImage, dev_dst: *ZbImage, pixfmt: c_int, s: Stream) c_int;
    pub extern fn zb_download(dev_src: *const ZbImage, host_dst: *ZbImage, pixfmt: c_int, s: Stream) c_int;
    pub extern fn zb_copy(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, s: Stream) c_int;
    pub extern fn zb_stream_synchronize(s: Stream) c_int;

    pub extern fn zb_conv_separable(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, kx: [*]const f32, nx: c_int, ky: [*]const f32, ny: c_int, border: c_int, s: Stream) c_int;
    pub extern fn zb_convolve(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, kernel: [*]const f32, kh: c_int, kw: c_int, border: c_int, s: Stream) c_int;
    pub extern fn zb_gaussian_blur(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, sigma: f32, s: Stream) c_int;
    pub extern fn zb_box_blur(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, radius: u32, s: Stream) c_int;
    pub extern fn zb_sharpen(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, radius: u32, s: Stream) c_int;
    pub extern fn zb_resize(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, method: c_int, b: f32, cc: f32, s: Stream) c_int;
    pub extern fn zb_rotate_bounds(rows: u32, cols: u32, angle: f32, out_rows: *u32, out_cols: *u32) c_int;
    pub extern fn zb_rotate_into_cs(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, angle: f32, cos_a: f32, sin_a: f32, method: c_int, b: f32, cc: f32, border: c_int, s: Stream) c_int;
    pub extern fn zb_warp(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, kind: c_int, m: [*]const f32, method: c_int, b: f32, cc: f32, s: Stream) c_int;
    // widening rows (SURVEY 8f): extract / crop, insert (.none), sobel, setBorder(zeroes), and the sharding extension
    pub extern fn zb_extract(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, l: f32, t: f32, r: f32, b: f32, angle: f32, cos_a: f32, sin_a: f32, method: c_int, mb: f32, mc: f32, border: c_int, s: Stream) c_int;
    pub extern fn zb_insert(self: *ZbImage, source: *const ZbImage, pixfmt: c_int, l: f32, t: f32, r: f32, b: f32, angle: f32, cos_a: f32, sin_a: f32, method: c_int, mb: f32, mc: f32, s: Stream) c_int;
    pub extern fn zb_insert_blend(self: *ZbImage, source: *const ZbImage, pixfmt: c_int, l: f32, t: f32, r: f32, b: f32, angle: f32, cos_a: f32, sin_a: f32, method: c_int, mb: f32, mc: f32, blend_mode: c_int, s: Stream) c_int;
    pub extern fn zb_sobel(src: *const ZbImage, dst_u8: *ZbImage, pixfmt: c_int, s: Stream) c_int;
    pub extern fn zb_order_blur(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, radius: u32, mode: c_int, param: f64, border: c_int, s: Stream) c_int;
    pub extern fn zb_motion_blur_linear(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, angle: f32, cos_a: f32, sin_a: f32, distance: u32, s: Stream) c_int;
    pub extern fn zb_motion_blur_radial(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, center_x: f32, center_y: f32, strength: f32, spin: c_int, s: Stream) c_int;
    pub extern fn zb_eigh_f64(a: [*]const f64, rows: u32, cols: u32, values: [*]f64, vectors: [*]f64) c_int;
    pub extern fn zb_eigh_f32(a: [*]const f32, rows: u32, cols: u32, values: [*]f32, vectors: [*]f32) c_int;
    pub extern fn zb_convert(src: *const ZbImage, src_pixfmt: c_int, dst: *ZbImage, dst_pixfmt: c_int, s: Stream) c_int;
    pub extern fn zb_psnr(a: *const ZbImage, b: *const ZbImage, pixfmt: c_int, out: *f64, s: Stream) c_int;
    pub extern fn zb_ssim(a: *const ZbImage, b: *const ZbImage, pixfmt: c_int, out: *f64, s: Stream) c_int;
    pub extern fn zb_mean_pixel_error(a: *const ZbImage, b: *const ZbImage, pixfmt: c_int, out: *f64, s: Stream) c_int;
    pub extern fn zb_canny(src: *const ZbImage, dst_u8: *ZbImage, pixfmt: c_int, sigma: f32, low_threshold: f32, high_threshold: f32, s: Stream) c_int;
    pub extern fn zb_set_border_zero(img: *ZbImage, pixfmt: c_int, l: u32, t: u32, r: u32, b: u32, s: Stream) c_int;
    pub extern fn zb_conv_separable_rows(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, kx: [*]const f32, nx: c_int, ky: [*]const f32, ny: c_int, border: c_int, row_begin: u32, row_end: u32, s: Stream) c_int;

    // host-pointer twins: Image.data in host memory, H2D + op + D2H inside the call
    pub extern fn zb_host_gaussian_blur(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, sigma: f32) c_int;
    pub extern fn zb_host_conv_separable(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, kx: [*]const f32, nx: c_int, ky: [*]const f32, ny: c_int, border: c_int) c_int;
    pub extern fn zb_host_box_blur(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, radius: u32) c_int;
    pub extern fn zb_host_sharpen(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, radius: u32) c_int;
    pub extern fn zb_host_resize(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, method: c_int, b: f32, cc: f32) c_int;
    pub extern fn zb_host_rotate_into(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, angle: f32, method: c_int, b: f32, cc: f32, border: c_int) c_int;
    pub extern fn zb_host_warp(src: *const ZbImage, dst: *ZbImage, pixfmt: c_int, kind: c_int, m: [*]const f32, method: c_int, b: f32, cc: f32) c_int;
};

/// Zig error set the status codes map onto (names as in the reference: image.zig:637,970,531-536; fdm.zig:114,142).
pub const Error = error{ DimensionMismatch, InvalidSigma, Unsupported, NotConverged, InvalidArgument, OutOfMemory, DeviceFailure, InvalidScaleFactor, InvalidDimensions, NoTargetSet, NoSourceSet, InsufficientData, InvalidComponents, InvalidThreshold, InvalidPercentile, InvalidTrim, ImageTooSmall, NotSquare, NotSymmetric, NotFinite };

pub fn check(status: c_int) Error!void {
    return switch (status) {
        0 => {},
        1 => error.DimensionMismatch,
        2 => error.InvalidSigma,
        3 => error.Unsupported,
        4 => error.NotConverged,
        6 => error.OutOfMemory,
        7 => error.DeviceFailure,
        8 => error.InvalidScaleFactor,
        9 => error.InvalidDimensions,
        10 => error.NoTargetSet,
        11 => error.NoSourceSet,
        12 => error.InsufficientData,
        13 => error.InvalidComponents,
        14 => error.InvalidThreshold,
        15 => error.InvalidPercentile,
        16 => error.InvalidTrim,
        17 => error.ImageTooSmall,
        18 => error.NotSquare,
        19 => error.NotSymmetric,
        20 => error.NotFinite,
        else => error.InvalidArgument,
    };
}

/// comptime pixel type -> zb_pixfmt (zignal's color.zig structs; `Rgba(f32)` is the 16-byte packed struct)
pub fn pixfmtOf(comptime T: type) c_int {
    if (T == u8) return 0;
    if (T == f32) return 1;
    const info = @typeInfo(T);
    if (info == .@"struct") {
        if (@sizeOf(T) == 3) return 2; // Rgb(u8)
        if (@sizeOf(T) == 4) return 3; // Rgba(u8)
        if (@sizeOf(T) == 16) return 4; // Rgba(f32)
    }
    @compileError("zignal_b200: unsupported pixel type " ++ @typeName(T));
}

/// `std.mem.Allocator` over device memory (cudaMallocAsync / cudaFreeAsync), so that
/// `Image(T).init(dev.allocator(), rows, cols)` / `deinit` (image.zig:124-158) work unchanged.
pub const DeviceAllocator = struct {
    stream: c.Stream = null,

    pub fn allocator(self: *DeviceAllocator) Allocator {
        return .{ .ptr = self, .vtable = &.{ .alloc = alloc, .resize = resize, .remap = remap, .free = free } };
    }
    fn alloc(ctx: *anyopaque, len: usize, _: std.mem.Alignment, _: usize) ?[*]u8 {
        const self: *DeviceAllocator = @ptrCast(@alignCast(ctx));
        var p: ?*anyopaque = null;
        if (c.zb_malloc(&p, len, self.stream) != 0) return null;
        return @ptrCast(p);
    }
    fn resize(_: *anyopaque, _: []u8, _: std.mem.Alignment, _: usize, _: usize) bool {
        return false;
    }
    fn remap(_: *anyopaque, _: []u8, _: std.mem.Alignment, _: usize, _: usize) ?[*]u8 {
        return null;
    }
    fn free(ctx: *anyopaque, buf: []u8, _: std.mem.Alignment, _: usize) void {
        const self: *DeviceAllocator = @ptrCast(@alignCast(ctx));
        _ = c.zb_free(buf.ptr, self.stream);
    }
};

/// The hot methods of `Image(T)` re-implemented over the C ABI.  `host = true` treats `img.data` as host
/// memory (literal drop-in); `host = false` as device memory obtained from `DeviceAllocator`.
pub fn Ops(comptime Image: type, comptime T: type, comptime host: bool) type {
    return struct {
        fn raw(img: Image) c.ZbImage {
            return .{ .data = @ptrCast(img.data.ptr), .rows = img.rows, .cols = img.cols, .stride = img.stride };
        }
        pub fn gaussianBlur(self: Image, out: Image, sigma: f32) Error!void { // image.zig:954
            var a = raw(self);
            var d = raw(out);
            try check(if (host) c.zb_host_gaussian_blur(&a, &d, pixfmtOf(T), sigma) else c.zb_gaussian_blur(&a, &d, pixfmtOf(T), sigma, null));
        }
        pub fn convolveSeparable(self: Image, out: Image, kernel_x: []const f32, kernel_y: []const f32, border: anytype) Error!void { // image.zig:935
            var a = raw(self);
            var d = raw(out);
            const b: c_int = @intFromEnum(border);
            try check(if (host)
                c.zb_host_conv_separable(&a, &d, pixfmtOf(T), kernel_x.ptr, @intCast(kernel_x.len), kernel_y.ptr, @intCast(kernel_y.len), b)
            else
                c.zb_conv_separable(&a, &d, pixfmtOf(T), kernel_x.ptr, @intCast(kernel_x.len), kernel_y.ptr, @intCast(kernel_y.len), b, null));
        }
        pub fn boxBlur(self: Image, out: Image, radius: u32) Error!void { // image.zig:635
            var a = raw(self);
            var d = raw(out);
            try check(if (host) c.zb_host_box_blur(&a, &d, pixfmtOf(T), radius) else c.zb_box_blur(&a, &d, pixfmtOf(T), radius, null));
        }
        pub fn sharpen(self: Image, out: Image, radius: usize) Error!void { // image.zig:785
            var a = raw(self);
            var d = raw(out);
            const r: u32 = @intCast(@min(radius, std.math.maxInt(u32)));
            try check(if (host) c.zb_host_sharpen(&a, &d, pixfmtOf(T), r) else c.zb_sharpen(&a, &d, pixfmtOf(T), r, null));
        }
        /// `method` is the reference's `Interpolation` union (interpolation.zig:53-68)
        pub fn resize(self: Image, out: Image, method: anytype) void { // image.zig:523 (infallible in the reference)
            var a = raw(self);
            var d = raw(out);
            const tag: c_int = @intFromEnum(std.meta.activeTag(method));
            const bc: [2]f32 = switch (method) {
                .mitchell => |m| .{ m.b, m.c },
                else => .{ 0, 0 },
            };
            _ = if (host) c.zb_host_resize(&a, &d, pixfmtOf(T), tag, bc[0], bc[1]) else c.zb_resize(&a, &d, pixfmtOf(T), tag, bc[0], bc[1], null);
        }
        pub fn rotateInto(self: Image, out: Image, angle: f32, method: anytype, border: anytype) void { // image.zig:564
            var a = raw(self);
            var d = raw(out);
            const tag: c_int = @intFromEnum(std.meta.activeTag(method));
            // cos/sin are computed here, with Zig's own @cos/@sin (transforms.zig:190-191), and cross the ABI as data
            _ = c.zb_rotate_into_cs(&a, &d, pixfmtOf(T), angle, @cos(angle), @sin(angle), tag, 0, 0, @intFromEnum(border), null);
        }

        // ---- device-resident only (host = false): the SURVEY 8(f) callers.  `OutU8` is the caller's Image(u8). ----
        pub fn sobel(self: Image, out: anytype) Error!void { // image.zig:999
            var a = raw(self);
            var d: c.ZbImage = .{ .data = @ptrCast(out.data.ptr), .rows = out.rows, .cols = out.cols, .stride = out.stride };
            try check(c.zb_sobel(&a, &d, pixfmtOf(T), null));
        }
        pub fn canny(self: Image, out: anytype, sigma: f32, low_threshold: f32, high_threshold: f32) Error!void { // image.zig:1041
            var a = raw(self);
            var d: c.ZbImage = .{ .data = @ptrCast(out.data.ptr), .rows = out.rows, .cols = out.cols, .stride = out.stride };
            try check(c.zb_canny(&a, &d, pixfmtOf(T), sigma, low_threshold, high_threshold, null));
        }
        pub fn medianBlur(self: Image, out: Image, radius: usize) Error!void { // image.zig:650 (percentile 0.5, .mirror)
            try orderBlur(self, out, radius, 0, 0.5, 2);
        }
        pub fn percentileBlur(self: Image, out: Image, radius: usize, percentile: f64, border: anytype) Error!void { // image.zig:672
            try orderBlur(self, out, radius, 0, percentile, @intFromEnum(border));
        }
        pub fn minBlur(self: Image, out: Image, radius: usize, border: anytype) Error!void { // image.zig:696
            try orderBlur(self, out, radius, 0, 0.0, @intFromEnum(border));
        }
        pub fn maxBlur(self: Image, out: Image, radius: usize, border: anytype) Error!void { // image.zig:719
            try orderBlur(self, out, radius, 0, 1.0, @intFromEnum(border));
        }
        pub fn midpointBlur(self: Image, out: Image, radius: usize, border: anytype) Error!void { // image.zig:742
            try orderBlur(self, out, radius, 1, 0.0, @intFromEnum(border));
        }
        pub fn alphaTrimmedMeanBlur(self: Image, out: Image, radius: usize, trim_fraction: f64, border: anytype) Error!void { // image.zig:767
            try orderBlur(self, out, radius, 2, trim_fraction, @intFromEnum(border));
        }
        fn orderBlur(self: Image, out: Image, radius: usize, mode: c_int, param: f64, border: c_int) Error!void {
            var a = raw(self);
            var d = raw(out);
            try check(c.zb_order_blur(&a, &d, pixfmtOf(T), @intCast(@min(radius, std.math.maxInt(u32))), mode, param, border, null));
        }
        /// `blur` is the reference's `MotionBlur` union (motion_blur.zig:12-55)
        pub fn motionBlur(self: Image, out: Image, blur: anytype) Error!void {
            var a = raw(self);
            var d = raw(out);
            switch (blur) {
                .linear => |l| try check(c.zb_motion_blur_linear(&a, &d, pixfmtOf(T), l.angle, @cos(l.angle), @sin(l.angle), @intCast(l.distance), null)),
                .radial_zoom => |z| try check(c.zb_motion_blur_radial(&a, &d, pixfmtOf(T), z.center_x, z.center_y, z.strength, 0, null)),
                .radial_spin => |z| try check(c.zb_motion_blur_radial(&a, &d, pixfmtOf(T), z.center_x, z.center_y, z.strength, 1, null)),
            }
        }
        /// `rect` is the reference's Rectangle(f32); `blend_mode` its `Blending` enum (blending.zig:8-22, same order as ZB_BLEND_*)
        pub fn insert(self: Image, source: Image, rect: anytype, angle: f32, method: anytype, blend_mode: anytype) void { // image.zig:604
            var d = raw(self);
            var a = raw(source);
            const tag: c_int = @intFromEnum(std.meta.activeTag(method));
            _ = c.zb_insert_blend(&d, &a, pixfmtOf(T), rect.l, rect.t, rect.r, rect.b, angle, @cos(angle), @sin(angle), tag, 0, 0, @intFromEnum(blend_mode), null);
        }
        pub fn extract(self: Image, out: Image, rect: anytype, angle: f32, method: anytype, border: anytype) void { // image.zig:594
            var a = raw(self);
            var d = raw(out);
            const tag: c_int = @intFromEnum(std.meta.activeTag(method));
            _ = c.zb_extract(&a, &d, pixfmtOf(T), rect.l, rect.t, rect.r, rect.b, angle, @cos(angle), @sin(angle), tag, 0, 0, @intFromEnum(border), null);
        }
        pub fn psnr(self: Image, other: Image) Error!f64 { // image.zig:1105
            var a = raw(self);
            var b = raw(other);
            var out: f64 = 0;
            try check(c.zb_psnr(&a, &b, pixfmtOf(T), &out, null));
            return out;
        }
        pub fn ssim(self: Image, other: Image) Error!f64 { // image.zig:1126
            var a = raw(self);
            var b = raw(other);
            var out: f64 = 0;
            try check(c.zb_ssim(&a, &b, pixfmtOf(T), &out, null));
            return out;
        }
        pub fn meanPixelError(self: Image, other: Image) Error!f64 { // image.zig:1145
            var a = raw(self);
            var b = raw(other);
            var out: f64 = 0;
            try check(c.zb_mean_pixel_error(&a, &b, pixfmtOf(T), &out, null));
            return out;
        }
        /// Image.convertInto(TargetType, out) (image.zig:396)
        pub fn convertInto(self: Image, comptime TargetType: type, out: anytype) Error!void {
            var a = raw(self);
            var d: c.ZbImage = .{ .data = @ptrCast(out.data.ptr), .rows = out.rows, .cols = out.cols, .stride = out.stride };
            try check(c.zb_convert(&a, pixfmtOf(T), &d, pixfmtOf(TargetType), null));
        }
    };
}
