"""Generates tests/golden/golden.json: SHA-256 of the oracle's outputs on seeded inputs.

The reference (Zig) cannot be built in this image, so these fixtures pin the ORACLE (regression guard
on the restatement) and give the GPU tests answers that do not require running the oracle at test
time.  Inputs come from numpy's PCG64 (`default_rng(seed)`): integers(0,256) for u8, random() for f32.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib as zo  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rand_image(rng, shape, dtype):
    if dtype == "u8":
        return rng.integers(0, 256, shape, dtype=np.uint8)
    return rng.random(shape, dtype=np.float32)


out = {"_about": "sha256 of oracle outputs; see make_golden.py", "conv": {}, "box": {}, "resize": {}, "rotate": {}, "fdm": {}}

for name, shape, dtype, sigma, seed in [
    ("c2_small_rgbaf32_sigma2.25", (512, 768, 4), "f32", 2.25, 2),
    ("rgba8_sigma1.4", (300, 401, 4), "u8", 1.4, 12),
    ("rgb8_sigma2.25", (257, 263, 3), "u8", 2.25, 13),
    ("u8_sigma0.8", (199, 301), "u8", 0.8, 14),
    ("f32_sigma3.0", (128, 160), "f32", 3.0, 15),
]:
    rng = np.random.default_rng(seed)
    img = rand_image(rng, shape, dtype)
    res = zo.gaussian_blur(img, sigma)
    out["conv"][name] = {"shape": list(shape), "dtype": dtype, "sigma": sigma, "seed": seed, "input_sha256": sha(img), "output_sha256": sha(res)}

for name, shape, dtype, radius, seed in [("c1_box3x3_512_u8", (512, 512), "u8", 1, 1), ("rgba8_r4", (320, 333, 4), "u8", 4, 21), ("f32_r2", (100, 120), "f32", 2, 22)]:
    rng = np.random.default_rng(seed)
    img = rand_image(rng, shape, dtype)
    out["box"][name] = {"shape": list(shape), "dtype": dtype, "radius": radius, "seed": seed, "input_sha256": sha(img),
                        "box_sha256": sha(zo.box_blur(img, radius)), "sharpen_sha256": sha(zo.sharpen(img, radius))}

for name, shape, dst, method, seed in [("c3_small_bicubic_4to1_rgb", (1024, 1024, 3), (256, 256), "bicubic", 3),
                                       ("rgba_lanczos_up", (90, 70, 4), (200, 131), "lanczos", 31),
                                       ("rgb_bilinear_down", (333, 222, 3), (100, 77), "bilinear", 32),
                                       ("rgb_mitchell", (120, 99, 3), (63, 140), "mitchell", 33),
                                       ("rgba_catmull", (64, 64, 4), (48, 80), "catmull_rom", 34),
                                       ("rgb_nearest", (50, 60, 3), (125, 31), "nearest", 35)]:
    rng = np.random.default_rng(seed)
    img = rand_image(rng, shape, "u8")
    out["resize"][name] = {"shape": list(shape), "dst": list(dst), "method": method, "seed": seed, "input_sha256": sha(img),
                           "output_sha256": sha(zo.resize(img, dst, method))}

for name, shape, angle, method, border, seed in [("c4_small_rgba_45deg", (270, 480, 4), float(np.float32(np.pi / 4)), "bilinear", "zero", 4),
                                                  ("rgb_0.3_bicubic_mirror", (100, 140, 3), 0.3, "bicubic", "mirror", 41)]:
    rng = np.random.default_rng(seed)
    img = rand_image(rng, shape, "u8")
    a = np.float32(angle)
    cs = (np.float32(np.cos(np.float64(a))), np.float32(np.sin(np.float64(a))))
    res = zo.rotate(img, a, method, border, cos_sin=cs)
    out["rotate"][name] = {"shape": list(shape), "angle": float(a), "cos": float(cs[0]), "sin": float(cs[1]), "method": method, "border": border,
                           "seed": seed, "input_sha256": sha(img), "out_shape": list(res.shape), "output_sha256": sha(res)}

for name, shape, seed in [("c5_small_rgb", (256, 256, 3), 5), ("rgba", (100, 150, 4), 51), ("gray", (64, 64), 52)]:
    rng = np.random.default_rng(seed)
    src = rand_image(rng, shape, "u8")
    tgt = (rand_image(rng, shape, "u8") // 2 + 40).astype(np.uint8)
    if len(shape) == 3:
        tgt[..., 1] = (tgt[..., 1].astype(np.int32) * 3 // 4 + tgt[..., 0] // 4).astype(np.uint8)
    out["fdm"][name] = {"shape": list(shape), "seed": seed, "source_sha256": sha(src), "target_sha256": sha(tgt), "output_sha256": sha(zo.fdm_match(src, tgt))}

(HERE / "golden.json").write_text(json.dumps(out, indent=1) + "\n")
print("wrote", HERE / "golden.json")
