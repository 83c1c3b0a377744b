#!/usr/bin/env python
"""Dev helper: end-to-end pose/s of PairBatchAligner.align from pinned host memory for a list of chunk_pairs values.
    python tools/e2e_chunks.py HEIGHT WIDTH LEVELS PAIRS DEPTH(0|1) chunk [chunk ...]"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
import dense_visual_odometry_b200 as dvo
from dense_visual_odometry_b200.synthetic import make_pairs_torch

h, w, levels, pairs, depth = (int(a) for a in sys.argv[1:6])
dev = torch.device("cuda", 0)
cam = bench.camera_for(dvo, w)
data = make_pairs_torch(range(pairs), dev, height=h, width=w)
al = dvo.PairBatchAligner(cam, h, w, levels, max_pairs=pairs, use_depth_residual=bool(depth))
tensors = (data["bgr_prev"], data["depth_prev"], data["bgr_cur"], data["depth_cur"])
hb = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in tensors]
for b, x in zip(hb, tensors):
    b.copy_(x)
nbytes = sum(x.numel() * x.element_size() for x in hb)
al.build(*tensors)
qt_r, _ = al.estimate()
for c in (int(a) for a in sys.argv[6:]):
    ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        qt, _ = al.align(*hb, chunk_pairs=c)
        ms.append(1e3 * (time.perf_counter() - t0))
    m = min(ms[1:])
    print(f"{w}x{h} {pairs} pairs, chunk_pairs {c}: {m:.1f} ms = {pairs / m * 1e3:.0f} pose/s, {nbytes / m / 1e6:.1f} GB/s, equal to resident: {np.array_equal(qt, qt_r)}")
