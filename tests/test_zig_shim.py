"""The Zig side of the boundary (zig/zignal_b200.zig) cannot be compiled here (no Zig toolchain), so its agreement with the C ABI is
checked textually: every entry point include/zignal_b200.h declares must have an `extern fn` in the shim (the block is generated
from the header and must not be stale), and the wrappers the north-star names must exist with the reference's method names."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
ZIG = (ROOT / "zig" / "zignal_b200.zig").read_text()


def test_every_header_entry_point_is_declared_in_the_zig_shim():
    sys.path.insert(0, str(ROOT))
    from zignal_b200 import _ffi
    declared = set(re.findall(r"pub extern fn (zb_[a-z0-9_]+)\(", ZIG))
    header = set(_ffi.declared_symbols())
    assert header - declared == set(), sorted(header - declared)
    assert declared - header == set(), sorted(declared - header)


def test_generated_extern_block_is_in_sync_with_the_header():
    res = subprocess.run([sys.executable, str(ROOT / "tools" / "gen_zig_externs.py"), "--check"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr


def test_the_named_path_has_wrappers_with_the_reference_names():
    # Image(T) methods (image.zig:523-621, 635, 785, 917-994)
    for name in ("gaussianBlur", "convolveSeparable", "convolve", "boxBlur", "sharpen", "resize", "scale", "rotate", "rotateInto",
                 "rotateBounds", "warp", "extract", "insert"):
        assert re.search(rf"pub fn {name}\(self: Image", ZIG), name
    # Matrix / SMatrix (Matrix.zig:696, 1570; SMatrix.zig:804), fdm (fdm.zig:42-141), pca (pca.zig:78-312)
    for pattern in (r"pub fn gemm\(self: Matrix", r"pub fn svd\(self: Matrix", r"pub fn eigh\(self: Matrix", r"pub fn smatrixSvd\(",
                    r"pub fn FeatureDistributionMatching\(", r"pub fn setTarget\(", r"pub fn setSource\(", r"pub fn match\(", r"pub fn update\(",
                    r"pub fn Pca\(", r"pub fn fit\(", r"pub fn project\(", r"pub fn projectInto\(", r"pub fn reconstruct\(", r"pub fn transform\(",
                    r"pub const Shard = struct", r"pub fn gaussianBlur\(self: Self, shard: \*Shard"):
        assert re.search(pattern, ZIG), pattern
    # the status codes map onto the header's enum values
    header = (ROOT / "include" / "zignal_b200.h").read_text()
    for name, value in re.findall(r"ZB_ERR_([A-Z_]+) = (\d+)", header):
        zig_name = "".join(w.capitalize() for w in name.lower().split("_"))
        if zig_name == "InvalidArgument":
            continue
        assert re.search(rf"{value} => error\.{zig_name}", ZIG), (name, value)
