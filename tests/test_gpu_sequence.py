"""GPU tests of BASELINE.json configs[2]: a synthetic frame stream along a smooth trajectory, t-distribution weights,
through SequenceAligner.  Small streams are compared pose by pose with the oracle (which is pinned to the real
reference for these weights); a full-size stream is checked through properties that do not need the oracle:
agreement with the pairwise batch path bit for bit, and closeness to the known motion."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import dvo_oracle as O  # noqa: E402

POSE_TOL = 1e-4


def _Km(K):
    return np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], dtype=np.float32)


@pytest.fixture(scope="module")
def dvo_mod():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import dense_visual_odometry_b200 as m
    return m


@pytest.mark.parametrize("weights", ["tdist", "none"])
def test_synthetic_stream_vs_oracle(dvo_mod, weights):
    from dense_visual_odometry_b200.synthetic import make_sequence
    m = dvo_mod
    s = make_sequence(7, height=120, width=160)
    Km = _Km(s["K"])
    cam = m.RGBDCameraModel(Km, s["depth_scale"])
    seq = m.SequenceAligner(cam, 120, 160, 3, max_frames=7, weights=weights)
    qt, stats = seq.align(s["bgr"], s["depth"].copy(), chunk_frames=3)   # chunked: pairs straddle chunk borders
    assert not stats["flags"].any()
    ref = O.OracleDVO(Km, s["depth_scale"], 3, weights=O.W_TDIST_REF if weights == "tdist" else O.W_NONE)
    ref.step(s["bgr"][0], s["depth"][0].copy())
    for p in range(6):
        Tr = ref.step(s["bgr"][p + 1], s["depth"][p + 1].copy())
        # iteration counts are printed, not asserted: the stop rule |d err| < 1e-6 sits at err's float32 resolution
        print("pair", p, "iters", stats["iters"][p][:3].tolist(), ref.last_result.iters)
        assert np.abs(qt[p, :4] - Tr.q).max() < POSE_TOL and np.abs(qt[p, 4:] - Tr.t).max() < POSE_TOL


def test_full_size_stream_properties(dvo_mod):
    """96 frames at 640x480, t-distribution weights, frames generated on the device: (1) the stream path (pyramids
    built once per frame, slots p / p+1) gives bit for bit the poses of the pairwise batch path on the same frames;
    (2) every pose is close to the known motion; (3) host (pinned) inputs give the same bits as resident ones."""
    import torch
    from dense_visual_odometry_b200.synthetic import make_sequence
    m = dvo_mod
    N = 96
    s = make_sequence(N, device=torch.device("cuda", 0))
    cam = m.RGBDCameraModel(_Km(s["K"]), s["depth_scale"])
    seq = m.SequenceAligner(cam, 480, 640, 4, max_frames=N, weights="tdist")
    qt, stats = seq.align(s["bgr"], s["depth"].clone(), chunk_frames=40)
    assert not stats["flags"].any()
    xi = np.stack([m.Se3.from_qt(q).log().reshape(6) for q in qt])
    err = np.abs(xi - s["xi"]).max(axis=1)
    print("stream: max / median twist error vs truth", err.max(), np.median(err))
    assert err.max() < 2e-3 and np.median(err) < 5e-4   # the reference's stop rule ends short of the exact motion
    al = m.PairBatchAligner(cam, 480, 640, 4, max_pairs=N - 1, weights="tdist")
    qb, sb = al.align(s["bgr"][:-1], s["depth"][:-1].clone(), s["bgr"][1:], s["depth"][1:].clone())
    assert np.array_equal(qb, qt) and np.array_equal(sb["iters"], stats["iters"])
    qh, _ = seq.align(s["bgr"].cpu().numpy(), s["depth"].cpu().numpy(), chunk_frames=32)
    assert np.array_equal(qh, qt)
