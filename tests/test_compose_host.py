"""Host-side scalar set-up of the composed callers (no GPU): pyramid level plan as pyramid.zig:55-89 computes it."""
import numpy as np


def test_pyramid_level_plan_matches_reference_formulas():
    from zignal_b200.compose import ImagePyramid
    plan = ImagePyramid.level_plan(480, 640, 8, 1.2, 1.6)            # buildDefault on a VGA frame
    assert len(plan) == 7
    for i, (rows, cols, sigma) in enumerate(plan, start=1):
        scale = np.float32(1.2) ** np.float32(i)
        assert rows == int(np.float32(480) / scale) and cols == int(np.float32(640) / scale)
        assert sigma is not None and abs(sigma - 1.6 * np.sqrt(float(scale) ** 2 - 1.0)) < 1e-5
    # a tiny blur sigma skips the blur (sigma <= 0.5, pyramid.zig:80), and the pyramid stops below 8 pixels (:61)
    assert ImagePyramid.level_plan(64, 64, 3, 1.05, 0.5)[0][2] is None
    assert len(ImagePyramid.level_plan(20, 20, 8, 1.5, 1.6)) == 2


def test_letterbox_rect_kats():
    """image/tests/resize.zig:15-90: 8x4 (cols x rows) into 6x6 -> content 6 wide, 3 tall at (0, 1); 3x9 into 12x4 -> 1 wide, 4 tall, centred."""
    from zignal_b200.compose import letterbox_rect
    r = letterbox_rect(4, 8, 6, 6)                       # wide image into a square: vertical padding
    assert (r.r - r.l, r.b - r.t, r.l, r.t) == (6, 3, 0, 1)
    r = letterbox_rect(9, 3, 4, 12)                      # tall image into a wide frame: horizontal padding, content height 4
    assert (r.r - r.l, r.b - r.t, r.l, r.t) == (1, 4, (12 - 1) // 2, 0)
    assert letterbox_rect(10, 20, 5, 10) is None         # same aspect ratio: no letterboxing
