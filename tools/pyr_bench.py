#!/usr/bin/env python
"""Dev helper: time of the pyramid build (both roles of a pair batch) with resident inputs, CUDA events."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import dense_visual_odometry_b200 as dvo
from dense_visual_odometry_b200.synthetic import make_pairs_torch, TUM_FR1, TUM_DEPTH_SCALE

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
dev = torch.device("cuda", 0)
Km = np.array([[TUM_FR1[0], 0, TUM_FR1[2]], [0, TUM_FR1[1], TUM_FR1[3]], [0, 0, 1]], dtype=np.float32)
cam = dvo.RGBDCameraModel(Km, TUM_DEPTH_SCALE)
d = make_pairs_torch(range(pairs), dev)
al = dvo.PairBatchAligner(cam, 480, 640, 4, max_pairs=pairs)
ms = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    al.build(d["bgr_prev"], d["depth_prev"], d["bgr_cur"], d["depth_cur"])
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print(f"pyramid build, {pairs} pairs: min {min(ms[1:]):.3f} ms, mean {np.mean(ms[1:]):.3f} ms")
