#!/usr/bin/env python
"""Small cases for memory-safety / race evidence: odd image sizes, every pass type, cluster mode.

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py          # where the tool is available
    DVO_B200_LIB=build/exp/libdvo_chk.so python tools/sanitize_cases.py       # own bounds check (-DDVO_BOUNDS_CHECK build)

The alignment kernel deliberately reads past plane ends (unclamped taps, pipeline over-run, L1 touches) inside the
slack rows dvo_create allocates, and reduces through distributed shared memory with hand-placed cluster barriers.
compute-sanitizer is closed on the GPU pool of this project, so the evidence is the library's own instrumentation
(every load / prefetch address of the streaming pass tested against the allocation extents, zero misses expected;
DVO_DEBUG_SHRINK_ROWS=n is the negative control) plus run-to-run bit-identity of every case (a race would show)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import dense_visual_odometry_b200 as dvo
    from dense_visual_odometry_b200.synthetic import make_pairs_numpy
    h, w, levels = 77, 101, 3
    d = make_pairs_numpy([1, 2, 3], height=h, width=w)
    K = d["K"]
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], dtype=np.float32)
    cam = dvo.RGBDCameraModel(Km, d["depth_scale"])
    frames = (d["bgr_prev"], d["depth_prev"], d["bgr_cur"], d["depth_cur"])
    cases = [dict(), dict(oob_mode="strict"), dict(weights="tdist"), dict(weights="huber"), dict(weights="huber_mad"),
             dict(use_depth_residual=True), dict(approximate_image2_gradient=True), dict(threads_per_block=256)]
    total = 0

    def violations(handle):
        try:
            return handle.bounds_violations()
        except dvo.DvoError:
            return None

    for kw in cases:
        al = dvo.PairBatchAligner(cam, h, w, levels, max_pairs=3, max_iterations=6, **kw)
        qt, st = al.align(*(x.copy() for x in frames))
        assert np.all(np.isfinite(qt)), kw
        for _ in range(2):   # run-to-run bit identity
            q2, _ = al.align(*(x.copy() for x in frames))
            assert np.array_equal(qt, q2), ("not deterministic", kw)
        v = violations(al.handle)
        total += v or 0
        print("batch", kw, "iters", st["iters"][:, :levels].sum(1).tolist(), "bounds violations", v, flush=True)
        del al
    for kw in (dict(cluster_size=8, weights="tdist"), dict(cluster_size=16, weights="tdist"), dict(cluster_size=8),
               dict(cluster_size=4, use_depth_residual=True), dict(cluster_size=2, weights="huber")):
        est = dvo.get_dvo("robust-dvo", cam, dvo.Se3.identity(), levels=levels, max_iterations=6, **kw)
        est.step(d["bgr_prev"][0], d["depth_prev"][0].copy())
        T = est.step(d["bgr_cur"][0], d["depth_cur"][0].copy())
        assert T is not None, kw
        for _ in range(2):
            est.step(d["bgr_prev"][0], d["depth_prev"][0].copy())
            T2 = est.step(d["bgr_cur"][0], d["depth_cur"][0].copy())
            assert np.array_equal(T.exp(), T2.exp()), ("not deterministic", kw)
        v = violations(est._h)
        total += v or 0
        print("cluster", kw, "iters", est.last_stats["iters"][0][:levels].tolist(), "bounds violations", v, flush=True)
        del est
    seq = dvo.SequenceAligner(cam, h, w, levels, max_frames=4, max_iterations=6, weights="tdist")
    bgr = np.concatenate([d["bgr_prev"][:2], d["bgr_cur"][:2]])
    dep = np.concatenate([d["depth_prev"][:2], d["depth_cur"][:2]])
    qt, st = seq.align(bgr, dep, chunk_frames=2)
    v = violations(seq.handle)
    total += v or 0
    print("sequence iters", st["iters"][:, :levels].sum(1).tolist(), "bounds violations", v, flush=True)
    print("sanitize cases done: total bounds violations", total if v is not None else "n/a (library built without the check)")


if __name__ == "__main__":
    main()
