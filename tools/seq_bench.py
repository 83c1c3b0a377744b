#!/usr/bin/env python
"""Dev helper: the 1000-frame t-distribution sequence of bench.py (configs[2]) from pinned host memory, for a list of
chunk_frames values (default: the library's)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
import dense_visual_odometry_b200 as dvo
from dense_visual_odometry_b200.synthetic import make_sequence

dev = torch.device("cuda", 0)
frames = 1000
s = make_sequence(frames, device=dev)
cam = bench.camera_for(dvo, bench.W)
seq = dvo.SequenceAligner(cam, bench.H, bench.W, bench.LEVELS, max_frames=frames, weights="tdist")
host = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in (s["bgr"], s["depth"])]
for hb, x in zip(host, (s["bgr"], s["depth"])):
    hb.copy_(x)
torch.cuda.synchronize()
qt_r, _ = seq.align(s["bgr"], s["depth"].clone())
for cf in [int(a) for a in sys.argv[1:]] or [256]:
    ms = []
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        qt, _ = seq.align(host[0], host[1], chunk_frames=cf)
        ms.append(1e3 * (time.perf_counter() - t0))
    print(f"chunk_frames {cf}: e2e min {min(ms[1:]):.2f} ms = {(frames - 1) / min(ms[1:]) * 1e3:.0f} pose/s, bitwise equal to resident: {np.array_equal(qt, qt_r)}")
