"""The C++ host mirror (zignal_b200/host/zignal.hpp) must compile against include/zignal_b200.h and link against the library: a small
program instantiates Image<T> / DeviceImage<T> for every pixel type (so every wrapper is type-checked against the C ABI) and runs the
host-only entry points (status names, Matrix.eigh).  No GPU is touched."""
import shutil
import subprocess
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIB_DIR = ROOT / "zignal_b200" / "lib"

PROGRAM = textwrap.dedent(r"""
    #include <cstdio>
    #include "zignal_b200/host/zignal.hpp"
    using namespace zignal;
    template <typename T> void instantiate() {          // never called: forces overload resolution of every wrapper against the C ABI
        Image<T> h, ho;
        h.gaussianBlur(ho, 1.0f); h.boxBlur(ho, 1); h.sharpen(ho, 1); h.resize(ho, Interpolation::bicubic());
        h.rotateInto(ho, 0.5f, Interpolation::bilinear(), BorderMode::zero);
        DeviceImage<T> d, o;
        DeviceImage<uint8_t> e;
        d.copy(o); d.gaussianBlur(o, 1.0f); d.convolveSeparable(o, {1.0f}, {1.0f}, BorderMode::mirror);
        const float k[3][3] = {{0, 0, 0}, {0, 1, 0}, {0, 0, 0}};
        d.convolve(o, k, BorderMode::wrap); d.boxBlur(o, 2); d.sharpen(o, 2);
        d.medianBlur(o, 1); d.percentileBlur(o, 1, 0.3, BorderMode::zero); d.minBlur(o, 1, BorderMode::mirror); d.maxBlur(o, 1, BorderMode::mirror);
        d.midpointBlur(o, 1, BorderMode::replicate); d.alphaTrimmedMeanBlur(o, 1, 0.1, BorderMode::replicate);
        d.motionBlurLinear(o, 0.7f, 5); d.motionBlurRadial(o, 0.5f, 0.5f, 0.5f, true);
        d.resize(o, Interpolation::lanczos()); d.rotateInto(o, 0.1f, Interpolation::nearest(), BorderMode::zero);
        const float m[6] = {1, 0, 0, 1, 0, 0};
        d.warp(o, ZB_XFORM_AFFINE, m, Interpolation::bilinear());
        d.extract(o, Rect{0, 0, 4, 4}, 0.0f, Interpolation::bilinear(), BorderMode::mirror);
        d.insert(o, Rect{0, 0, 4, 4}, 0.0f, Interpolation::bilinear(), Blending::overlay);
        d.sobel(e); d.canny(e, 1.0f, 10.0f, 30.0f);
        (void)d.psnr(o); (void)d.ssim(o); (void)d.meanPixelError(o);
        d.template convertInto<RgbaF32>(DeviceImage<RgbaF32>{});
        (void)d.view(0, 0, 1, 1);
    }
    template void instantiate<uint8_t>();
    template void instantiate<float>();
    template void instantiate<Rgb8>();
    template void instantiate<Rgba8>();
    template void instantiate<RgbaF32>();
    int main() {
        std::printf("%s %s %s\n", zb_status_name(ZB_ERR_INVALID_THRESHOLD), zb_status_name(ZB_ERR_NOT_SYMMETRIC), zb_status_name(ZB_ERR_IMAGE_TOO_SMALL));
        Eigh e = eigh({2, 1, 1, 2}, 2);
        std::printf("%.6f %.6f\n", e.values[0], e.values[1]);
        try { eigh({0, 1, 2, 0}, 2); } catch (const Error& err) { std::printf("%d %s\n", err.status, err.what()); }
        return 0;
    }
""")


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_cpp_mirror_compiles_links_and_runs(tmp_path):
    if not (LIB_DIR / "libzignal_b200.so").exists():
        pytest.fail("libzignal_b200.so missing: run __graft_entry__.build()")
    src = tmp_path / "mirror.cpp"
    src.write_text(PROGRAM)
    exe = tmp_path / "mirror"
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    cmd = [gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT), str(src), "-o", str(exe), "-L", str(LIB_DIR), "-lzignal_b200",
           f"-Wl,-rpath,{LIB_DIR}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "InvalidThreshold NotSymmetric ImageTooSmall"
    assert lines[1] == "1.000000 3.000000"
    assert lines[2].startswith("19 ") and "NotSymmetric" in lines[2]
