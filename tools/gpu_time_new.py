"""Device timings (CUDA events, warm, median of 5) of the 8(f) additions at 4096x4096: Canny, median / min / alpha-trimmed, ssim / psnr."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import zignal_b200 as zb  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    rng = np.random.default_rng(0)
    n = 4096
    yy, xx = np.mgrid[0:n, 0:n]
    gray = ((np.sin(xx / 37.0) * np.cos(yy / 53.0) * 100 + 128) + rng.normal(0, 3, (n, n))).clip(0, 255).astype(np.uint8)
    rgba = rng.integers(0, 256, (n, n, 4), dtype=np.uint8)
    g, c = zb.Image.from_numpy(gray), zb.Image.from_numpy(rgba)
    c2 = zb.Image.from_numpy(np.roll(rgba, 1, axis=1))
    out8 = zb.Image.init(n, n, zb.PixFmt.U8, device="cuda")
    outg = zb.Image.init(n, n, zb.PixFmt.U8, device="cuda")
    outc = zb.Image.init(n, n, zb.PixFmt.RGBA8, device="cuda")
    res = {"size": [n, n]}
    res["canny_gray_sigma1.4_ms"] = timed(lambda: g.canny(1.4, 20.0, 60.0, out=out8))
    res["canny_rgba_sigma1.0_ms"] = timed(lambda: c.canny(1.0, 20.0, 60.0, out=out8))
    res["median_r1_gray_ms"] = timed(lambda: g.median_blur(1, out=outg))
    res["median_r2_gray_ms"] = timed(lambda: g.median_blur(2, out=outg))
    res["median_r3_gray_ms"] = timed(lambda: g.median_blur(3, out=outg))
    res["median_r5_gray_ms"] = timed(lambda: g.median_blur(5, out=outg))
    res["median_r2_rgba_ms"] = timed(lambda: c.median_blur(2, out=outc))
    res["min_r3_rgba_ms"] = timed(lambda: c.min_blur(3, out=outc))
    res["alpha_trim_r2_gray_ms"] = timed(lambda: g.alpha_trimmed_mean_blur(2, 0.2, out=outg))
    from zignal_b200.compose import motion_blur_linear, motion_blur_radial
    res["motion_linear_diag_d15_rgba_ms"] = timed(lambda: motion_blur_linear(c, outc, 0.6, 15))
    res["motion_linear_horiz_d15_rgba_ms"] = timed(lambda: motion_blur_linear(c, outc, 0.0, 15))
    res["motion_zoom_rgba_ms"] = timed(lambda: motion_blur_radial(c, outc, 0.5, 0.5, 0.5))
    res["motion_spin_rgba_ms"] = timed(lambda: motion_blur_radial(c, outc, 0.5, 0.5, 0.5, spin=True))
    src = zb.Image.from_numpy(rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8))
    res["insert_blend_overlay_1024_into_4096_ms"] = timed(lambda: c.insert(src, (500.0, 400.0, 2500.0, 2400.0), 0.3, zb.Interpolation.BILINEAR,
                                                                            blend=zb.Blending.OVERLAY))
    outf = zb.Image.init(n, n, zb.PixFmt.RGBAF32, device="cuda")
    res["convert_rgba8_to_rgbaf32_ms"] = timed(lambda: c.convert(zb.PixFmt.RGBAF32, out=outf))
    res["convert_rgba8_to_u8_ms"] = timed(lambda: c.convert(zb.PixFmt.U8, out=outg))
    res["psnr_rgba_ms"] = timed(lambda: c.psnr(c2))
    res["ssim_rgba_ms"] = timed(lambda: c.ssim(c2))
    res["ssim_gray_ms"] = timed(lambda: g.ssim(outg))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
