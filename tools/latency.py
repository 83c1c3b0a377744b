#!/usr/bin/env python
"""Dev helper: single-pair latency of the estimate (resident pyramids) for the CTA shapes / cluster sizes."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import dense_visual_odometry_b200 as m
from dense_visual_odometry_b200.synthetic import make_pairs_numpy, TUM_FR1, TUM_DEPTH_SCALE

d = make_pairs_numpy([0], height=480, width=640)
Km = np.array([[TUM_FR1[0], 0, TUM_FR1[2]], [0, TUM_FR1[1], TUM_FR1[3]], [0, 0, 1]], dtype=np.float32)
cam = m.RGBDCameraModel(Km, TUM_DEPTH_SCALE)
dev = torch.device("cuda", 0)
args = [torch.as_tensor(d[k]).to(dev) for k in ("bgr_prev", "depth_prev", "bgr_cur", "depth_cur")]
ref = None
W = dict(weights=sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] != "depth" else {}
if len(sys.argv) > 1 and sys.argv[1] == "depth":
    W = dict(use_depth_residual=True)
for name, kw in [("1 CTA x 128 thr", dict(threads_per_block=128)), ("1 CTA x 256 thr", dict(threads_per_block=256)),
                 ("cluster 2", dict(cluster_size=2)), ("cluster 4", dict(cluster_size=4)),
                 ("cluster 8", dict(cluster_size=8)), ("cluster 16", dict(cluster_size=16))]:
    al = m.PairBatchAligner(cam, 480, 640, 4, max_pairs=1, **kw, **W)
    al.build(*args)
    for _ in range(3):
        qt, st = al.estimate()
    ms = []
    for _ in range(10):
        al.estimate(to_host=False)
        torch.cuda.synchronize()
        ms.append(al.last_kernel_ms())
    if ref is None:
        ref = qt
    print(f"{name:18s} kernel {np.median(ms):7.3f} ms  iters {st['iters'][0][:4].tolist()}  |dqt| vs first {np.abs(qt - ref).max():.2e}")
