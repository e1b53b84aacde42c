"""Generates tests/golden/golden_8f.json: SHA-256 of the oracle's outputs for the SURVEY 8(f) additions (cases in tests/golden_8f.py).
Like golden.json these pin the ORACLE (the Zig reference cannot be built here) and give the GPU tests fixed answers.
Run:  python tests/golden/make_golden_8f.py"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib as zo  # noqa: E402
from golden_8f import CASES, eigh_case  # noqa: E402
from gpu_utils import sha  # noqa: E402

out = {"_about": "sha256 of oracle outputs for the 8(f) additions; see make_golden_8f.py", "cases": {}}
for name, (build, oracle, _device) in CASES.items():
    img = build()
    res = oracle(img)
    out["cases"][name] = {"input_sha256": sha(img), "shape": list(res.shape), "dtype": str(res.dtype), "output_sha256": sha(res)}
vals, vecs = zo.eigh(eigh_case())
out["eigh_12x12_f64"] = {"input_sha256": sha(eigh_case()), "values_sha256": sha(vals), "vectors_sha256": sha(vecs)}
(HERE / "golden_8f.json").write_text(json.dumps(out, indent=1) + "\n")
print("wrote", HERE / "golden_8f.json", len(out["cases"]), "cases")
