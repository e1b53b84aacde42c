"""GPU outputs must hash to the committed golden fixtures (integer formats are bit-exact)."""
import numpy as np
import pytest

from gpu_utils import golden, rand_image, sha
from test_golden_oracle import _fdm_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available()
    import zignal_b200 as zb
    return zb


def test_box_resize_rotate_fixtures(zb):
    g = golden()
    for name, c in g["box"].items():
        img = rand_image(np.random.default_rng(c["seed"]), tuple(c["shape"]), np.uint8 if c["dtype"] == "u8" else np.float32)
        assert sha(zb.Image.from_numpy(img).box_blur(c["radius"]).to_numpy()) == c["box_sha256"], name
        assert sha(zb.Image.from_numpy(img).sharpen(c["radius"]).to_numpy()) == c["sharpen_sha256"], name
    for name, c in g["resize"].items():
        img = rand_image(np.random.default_rng(c["seed"]), tuple(c["shape"]), np.uint8)
        dev = zb.Image.from_numpy(img)
        out = dev.resize(zb.Image.init(c["dst"][0], c["dst"][1], dev.pixfmt), zb.Interpolation[c["method"].upper()])
        assert sha(out.to_numpy()) == c["output_sha256"], name
    for name, c in g["rotate"].items():
        img = rand_image(np.random.default_rng(c["seed"]), tuple(c["shape"]), np.uint8)
        out = zb.Image.from_numpy(img).rotate(np.float32(c["angle"]), zb.Interpolation[c["method"].upper()], zb.BorderMode[c["border"].upper()],
                                              cos_sin=(np.float32(c["cos"]), np.float32(c["sin"])))
        assert list(out.to_numpy().shape) == c["out_shape"] and sha(out.to_numpy()) == c["output_sha256"], name


def test_fdm_fixtures(zb):
    from zignal_b200.fdm import FeatureDistributionMatching
    for name, c in golden()["fdm"].items():
        src, tgt = _fdm_pair(c)
        f = FeatureDistributionMatching(zb.image.pixfmt_of_array(src))
        s = zb.Image.from_numpy(src)
        f.match(s, zb.Image.from_numpy(tgt))
        f.deinit()
        assert sha(s.to_numpy()) == c["output_sha256"], name


def test_8f_fixtures(zb):
    """The device results of the 8(f) additions hash to tests/golden/golden_8f.json (all listed operations are bit-exact)."""
    import json
    from pathlib import Path

    from golden_8f import CASES
    g = json.loads((Path(__file__).resolve().parent / "golden" / "golden_8f.json").read_text())
    for name, (build, _oracle, device) in CASES.items():
        img = build()
        res = device(zb, img)
        assert list(res.shape) == g["cases"][name]["shape"] and str(res.dtype) == g["cases"][name]["dtype"], name
        assert sha(res) == g["cases"][name]["output_sha256"], name
