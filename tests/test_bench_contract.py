"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints one JSON line with the keys the driver reads,
non-zero ranks of a torchrun launch stay silent, and the product arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=str(ROOT))


def test_reference_arm_prints_the_contract_line():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    baseline = json.loads((ROOT / "BASELINE.json").read_text())
    assert line["impl"] == "reference" and baseline["metric"].startswith(line["metric"]) and line["unit"] == "Mpixels/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    res = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return                                   # on a GPU box the driver runs the real thing
    res = _run(["--steps", "1", "--warmup", "3", "--no-e2e", "--no-cpu-baseline"], timeout=300)
    assert res.returncode != 0 and res.stdout.strip() == ""      # no JSON line from a CPU fallback: there is none
