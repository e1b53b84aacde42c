"""Shared helpers for the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import hashlib
import json
from pathlib import Path

import numpy as np

BORDERS = ["zero", "replicate", "mirror", "wrap"]
METHODS = ["nearest", "bilinear", "bicubic", "catmull_rom", "mitchell", "lanczos"]
GOLDEN = Path(__file__).resolve().parent / "golden" / "golden.json"


def rand_image(rng, shape, dtype):
    if dtype == np.uint8:
        return rng.integers(0, 256, shape, dtype=np.uint8)
    return rng.random(shape, dtype=np.float32)


def rel_err(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-6))) if want.size else 0.0


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden():
    return json.loads(GOLDEN.read_text())


def border_enum(zb, name):
    return zb.BorderMode[name.upper()]


def method_enum(zb, name):
    return zb.Interpolation[name.upper()]
