#!/usr/bin/env python
"""Dev helper: kernel-only timing of the batched estimate for one build of libdvo_b200.so.

    DVO_B200_LIB=path/to/lib.so python tools/kbench.py --pairs 1184 --reps 5 [--save ref.npz | --ref ref.npz]

Pyramids are built once (resident), the estimate is launched `reps` times and timed with the library's own
CUDA events.  Prints one JSON line: kernel ms, algorithmic GB/s (12 B/px/iteration) and fraction of the measured
HBM peak, pose / iteration-count differences against a reference run."""
import argparse, json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1184)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--weights", default="none")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--depth-residual", action="store_true")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--fixed-iters", type=int, default=0,
                    help="run exactly this many iterations at every level (no convergence test): equal work for builds "
                         "whose results differ")
    ap.add_argument("--save")
    ap.add_argument("--ref")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch
    import dense_visual_odometry_b200 as dvo
    from dense_visual_odometry_b200.synthetic import make_pairs_torch, TUM_FR1, TUM_DEPTH_SCALE
    dev = torch.device("cuda", 0)
    s = a.width / 640.0
    Km = np.array([[TUM_FR1[0] * s, 0, TUM_FR1[2] * s], [0, TUM_FR1[1] * s, TUM_FR1[3] * s], [0, 0, 1]], dtype=np.float32)
    cam = dvo.RGBDCameraModel(Km, TUM_DEPTH_SCALE)
    data = make_pairs_torch(range(a.pairs), dev, height=a.height, width=a.width)
    extra = {}
    if a.fixed_iters:
        extra = dict(max_iterations=a.fixed_iters, tolerance=-1.0, max_increased_steps_allowed=1 << 20)
    al = dvo.PairBatchAligner(cam, a.height, a.width, a.levels, max_pairs=a.pairs, weights=a.weights,
                              threads_per_block=a.threads, blocks_per_sm=a.blocks_per_sm,
                              use_depth_residual=a.depth_residual, **extra)
    al.build(data["bgr_prev"], data["depth_prev"], data["bgr_cur"], data["depth_cur"])
    ms = []
    for _ in range(a.reps + 1):
        qt, st = al.estimate()
        ms.append(al.last_kernel_ms())
    ms = ms[1:]
    px, h, w = [], a.height, a.width
    for _ in range(a.levels):
        px.append(h * w)
        h, w = (h + 1) // 2, (w + 1) // 2
    it = st["iters"][:, :a.levels].astype(np.int64)
    alg = float((it * np.array(px)[None]).sum() * 12)
    peak = 6547.2
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    best = min(ms)
    out = {"tag": a.tag, "lib": os.environ.get("DVO_B200_LIB", "default"), "pairs": a.pairs, "weights": a.weights,
           "kernel_ms_min": best, "kernel_ms_mean": float(np.mean(ms)), "alg_GBps": alg / best / 1e6,
           "frac": alg / best / 1e6 / peak, "iters_mean": float(it.sum(1).mean()),
           "pose_s": 1e3 * a.pairs / best}
    xi = data["xi"]
    err = [float(np.abs(dvo.Se3.from_qt(qt[j]).log().reshape(6) - xi[j]).max()) for j in range(a.pairs)]
    out["max_twist_err_vs_truth"] = max(err)
    if a.save:
        np.savez(a.save, qt=qt, iters=it)
    if a.ref:
        r = np.load(a.ref)
        out["max_pose_diff_vs_ref"] = float(np.abs(r["qt"] - qt).max())
        out["iters_changed"] = int((r["iters"] != it).any(1).sum())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
