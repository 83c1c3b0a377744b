"""The persistent kernel hands pairs from CTA to CTA (time slices) and leaves the last ones to a cluster kernel; neither
may change a single bit of a result: a batch larger than the grid must equal the same pairs run in small launches
and in launches of one pair (one CTA per pair from start to end)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dvo_mod():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import dense_visual_odometry_b200 as m
    return m


@pytest.mark.parametrize("weights,depth", [("none", False), ("tdist", False), ("huber", True)])
def test_large_batch_equals_small_launches_bitwise(dvo_mod, weights, depth):
    from dense_visual_odometry_b200.synthetic import make_pairs_numpy
    m = dvo_mod
    h, w, levels, n = 120, 160, 3, 700      # 700 pairs > 296 CTAs: time slices + tail kernel
    d = make_pairs_numpy(range(n), height=h, width=w)
    K = d["K"]
    cam = m.RGBDCameraModel(np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32), d["depth_scale"])
    frames = (d["bgr_prev"], d["depth_prev"], d["bgr_cur"], d["depth_cur"])
    big = m.PairBatchAligner(cam, h, w, levels, max_pairs=n, weights=weights, use_depth_residual=depth)
    big.build(*(x.copy() for x in frames))
    qt_big, st_big = big.estimate()
    assert np.all(np.isfinite(qt_big)) and not (st_big["flags"] & 1).any()
    small = m.PairBatchAligner(cam, h, w, levels, max_pairs=100, weights=weights, use_depth_residual=depth)
    for lo in range(0, n, 100):
        small.build(*(x[lo:lo + 100].copy() for x in frames))
        qt, st = small.estimate()
        assert np.array_equal(qt, qt_big[lo:lo + 100]), f"pairs {lo}.. differ between the large batch and small launches"
        assert np.array_equal(st["iters"], st_big["iters"][lo:lo + 100])
    # launches of ONE pair never yield: one CTA runs the pair from start to end (the small launches above hand their
    # last pairs to the tail kernel's clusters like the large one)
    one = m.PairBatchAligner(cam, h, w, levels, max_pairs=1, weights=weights, use_depth_residual=depth)
    for i in (0, 1, 350, 699):
        one.build(*(x[i:i + 1].copy() for x in frames))
        qt, st = one.estimate()
        assert np.array_equal(qt[0], qt_big[i]), f"pair {i} differs between the batch and a one-pair launch"
        assert np.array_equal(st["iters"][0], st_big["iters"][i])
    # and from run to run
    qt2, _ = big.estimate()
    assert np.array_equal(qt2, qt_big)
