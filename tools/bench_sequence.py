#!/usr/bin/env python
"""BASELINE.json configs[2]: a synthetic 640x480 sequence of N frames (default 1000), t-distribution weights, full
coarse-to-fine Gauss-Newton per pair, through SequenceAligner (every frame's pyramid built once, all N-1 pairs in
flight on one GPU).  Prints one JSON line; not the headline bench (bench.py is).  Run on the GPU box:
    python tools/bench_sequence.py --frames 1000 --weights tdist"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--weights", default="tdist", choices=["none", "tdist", "tdist_mean", "huber", "huber_mad"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--depth-residual", action="store_true")
    args = ap.parse_args()
    import torch
    import dense_visual_odometry_b200 as dvo
    from dense_visual_odometry_b200.synthetic import make_sequence
    from dense_visual_odometry_b200.sharding import chain_poses

    dev = torch.device("cuda", 0)
    H, W, L, N = 480, 640, 4, args.frames
    s = make_sequence(N, device=dev)
    K = s["K"]
    cam = dvo.RGBDCameraModel(np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32), s["depth_scale"])
    kw = dict(weights=args.weights)
    if args.depth_residual:
        kw["use_depth_residual"] = True
    seq = dvo.SequenceAligner(cam, H, W, L, max_frames=N, **kw)
    host = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in (s["bgr"], s["depth"])]
    host[0].copy_(s["bgr"])
    host[1].copy_(s["depth"])
    torch.cuda.synchronize(dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return out, float(np.median(ts))

    (qt_r, st_r), ms_res = timed(lambda: seq.align(s["bgr"], s["depth"].clone()))
    (qt_h, st_h), ms_e2e = timed(lambda: seq.align(host[0], host[1]))
    assert np.array_equal(qt_r, qt_h), "host and resident inputs must give the same poses"
    xi = np.stack([dvo.Se3.from_qt(q).log().reshape(6) for q in qt_r])
    err = np.abs(xi - s["xi"]).max(axis=1)
    # absolute trajectory from the relative poses vs the true one (camera centres)
    est_abs = chain_poses(qt_r)
    from dense_visual_odometry_b200.synthetic import se3_exp
    gt = [np.zeros(3)]
    Tcur = np.eye(4)
    for k in range(N - 1):
        R, t = se3_exp(s["xi"][k])
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        Tcur = Tcur @ np.linalg.inv(T)          # base_dense_visual_odometry.py:79: pose <- pose * T^-1
        gt.append(Tcur[:3, 3].copy())
    gt_t = np.stack(gt)
    est_t = np.stack([np.asarray(T.tvec, np.float64).reshape(3) for T in est_abs])
    out = {"metric": "pose estimates/sec, one 640x480 sequence, 4-level pyramid", "frames": N, "pairs": N - 1,
           "weights": args.weights, "depth_residual": bool(args.depth_residual),
           "value_resident": (N - 1) / (ms_res / 1e3), "value_e2e_host_pinned": (N - 1) / (ms_e2e / 1e3), "unit": "pose/s",
           "ms_resident": ms_res, "ms_e2e": ms_e2e,
           "gn_iterations_per_pose_mean": float(st_r["iters"][:, :L].sum(1).mean()),
           "flags_nonzero": int((st_r["flags"] != 0).sum()),
           "max_abs_twist_error_vs_truth": float(err.max()), "median_abs_twist_error_vs_truth": float(np.median(err)),
           "trajectory_end_drift_m": float(np.linalg.norm(est_t[-1] - gt_t[-1])),
           "trajectory_length_m": float(np.linalg.norm(np.diff(gt_t, axis=0), axis=1).sum()),
           "gpu_launches": seq.launch_count()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
