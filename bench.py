#!/usr/bin/env python
"""Benchmark of the photometric-alignment hot path (BASELINE.json: pose estimates/s at 640x480, 4 levels).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference itself on the host cores

A "step" is one pass of the hot path over one batch of synthetic frame pairs: gray conversion + depth clamp + median
pyramids + gradient records for both frames of every pair, then the full coarse-to-fine Gauss-Newton estimate of
every pair.  The headline workload is BASELINE.json configs[3]: 4096 independent 640x480 pairs SHARDED over the N GPUs
(strong scaling: rank g takes pairs shard_range(4096, g, N)); `value` is measured with the frames already resident in
HBM, `e2e` is the same work through the public API with HOST (pinned) buffers, H2D/D2H inside the timed region.
Extra records of the same line: `weak_scaling` (4096 pairs per GPU), `configs` (the other BASELINE.json configs on one
GPU: Huber pair/batch, 1000-frame t-distribution sequence, 1280x720 and 1920x1080 five-level photometric + depth),
`cpu_baseline` (the real reference when baseline/_ref is importable, and the NumPy port).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

H, W, LEVELS = 480, 640, 4
TOTAL_PAIRS = 4096
METRIC = "pose estimates/sec at 640x480, 4-level pyramid"
UNIT = "pose/s"
B_PX = 12  # algorithmic bytes per pixel per GN iteration: I1 u8 + D1 u16 + I2 u8 + gx f32 + gy f32


def level_pixels(h=H, w=W, levels=LEVELS):
    px = []
    for _ in range(levels):
        px.append(h * w)
        h, w = (h + 1) // 2, (w + 1) // 2   # image_pyramid.py:21 (ceil division)
    return px


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ CPU side
def _cpu_worker(args):
    """One pose estimate with the oracle port on one process (BLAS pinned to one thread)."""
    seed, weights, approx, depth = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    import dense_visual_odometry_b200  # noqa: F401
    from dense_visual_odometry_b200.synthetic import make_pairs_numpy
    from oracle import dvo_oracle as O
    d = make_pairs_numpy([seed], height=H, width=W)
    K = d["K"]
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], dtype=np.float32)
    wmode = {"none": O.W_NONE, "tdist": O.W_TDIST_REF, "huber": O.W_HUBER, "huber_mad": O.W_HUBER_MAD}[weights]
    t0 = time.perf_counter()
    est = O.OracleDVO(Km, d["depth_scale"], LEVELS, weights=wmode, approximate_image2_gradient=approx,
                      use_depth_residual=depth)
    est.step(d["bgr_prev"][0], d["depth_prev"][0].copy())
    T = est.step(d["bgr_cur"][0], d["depth_cur"][0].copy())
    dt = time.perf_counter() - t0
    return seed, np.concatenate([T.q, T.t]).astype(np.float32), dt, est.last_result.iters


def cpu_pool(n):
    import multiprocessing as mp
    return mp.get_context("spawn").Pool(n)


def run_port_sample(pool, seeds, weights, approx=False, depth=False):
    """Estimates len(seeds) pairs in parallel with the NumPy port; returns (pairs/s, results, seconds)."""
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, [(s, weights, approx, depth) for s in seeds])
    dt = time.perf_counter() - t0
    return len(seeds) / dt, res, dt


def run_reference_sample(seeds, use_weighter=False, warm=True):
    """Estimates the pairs one after the other with the REAL reference (get_dvo(...).step, its own numba / BLAS / OpenCV
    threading over all host cores).  Returns (pairs/s, poses, seconds, numba threads), timing excludes import + JIT."""
    import dense_visual_odometry_b200  # noqa: F401
    from dense_visual_odometry_b200.synthetic import make_pairs_numpy
    from oracle import reference_runner as R
    ns = R.load("inclusive")
    data = [make_pairs_numpy([s], height=H, width=W) for s in seeds]
    if warm:   # JIT warm-up (a throw-away pose)
        d = data[0]
        R.estimate_pair(d["K"], d["depth_scale"], LEVELS, d["bgr_prev"][0], d["depth_prev"][0], d["bgr_cur"][0],
                        d["depth_cur"][0], use_weighter)
    out = []
    t0 = time.perf_counter()
    for d in data:
        out.append(R.estimate_pair(d["K"], d["depth_scale"], LEVELS, d["bgr_prev"][0], d["depth_prev"][0],
                                   d["bgr_cur"][0], d["depth_cur"][0], use_weighter))
    dt = time.perf_counter() - t0
    return len(seeds) / dt, out, dt, ns.threads


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores: the unmodified
    reference (baseline/_ref) when it is importable, else the NumPy port pinned to it (one pair per process)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference_runner as R
    cores = host_cores()
    use_real = R.available() and args.weights in ("none", "tdist") and not args.depth_residual and not args.port_only
    times = []
    if use_real:
        n_per_step = 1
        R.load("inclusive")
        for s in range(args.warmup + args.steps):
            v, _, dt, threads = run_reference_sample([1000 * s], use_weighter=(args.weights == "tdist"), warm=(s == 0))
            if s >= args.warmup:
                times.append(dt)
        kind, workers = "reference", threads
        sample = (f"1 synthetic 640x480 pair per step through the unmodified reference (get_dvo('robust-dvo').step, "
                  f"numba/BLAS/OpenCV threads = {threads}), {args.steps} steps; out-of-image guard applied externally "
                  f"(SURVEY F1/F2)")
    else:
        workers = max(1, min(cores, args.cpu_workers or cores))
        n_per_step = workers
        pool = cpu_pool(workers)
        try:
            for s in range(args.warmup + args.steps):
                seeds = [1000 * s + i for i in range(workers)]
                v, _, dt = run_port_sample(pool, seeds, args.weights, args.approximate_gradient, args.depth_residual)
                if s >= args.warmup:
                    times.append(dt)
        finally:
            pool.close()
        kind = "port"
        sample = (f"{workers} synthetic 640x480 pairs per step (one per process, BLAS 1 thread each), {args.steps} "
                  f"steps; oracle/dvo_oracle.py (NumPy port pinned to the reference)")
    total = sum(times)
    value = n_per_step * args.steps / total
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "pairs_per_step": n_per_step, "levels": LEVELS,
                   "weights": args.weights},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(out)


# ------------------------------------------------------------------------------------------------ GPU side
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


def workload_name(args):
    return (f"{TOTAL_PAIRS} independent synthetic 640x480 RGB-D pairs with known SE(3) motion sharded over the GPUs "
            f"(BASELINE.json configs[3]; pair type of configs[1]), {LEVELS}-level pyramid, weights={args.weights}"
            + (", approximate_image2_gradient" if getattr(args, "approximate_gradient", False) else "")
            + (", photometric + depth residual" if getattr(args, "depth_residual", False) else ""))


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_pixel_iteration():
    """DRAM bytes per pixel-iteration of align_kernel from this round's ncu --set full capture."""
    p = ROOT / "profiles" / "r2" / "ncu_traffic.json"
    try:
        return json.loads(p.read_text())
    except Exception:
        return None


def algorithmic_bytes(iters, px, weights="none", depth=False):
    """sum over pairs and levels of iterations x pixels x 12 B; + 2 B/px (D2) with the depth residual;
    Huber/MAD + 4 B/px for the residual pre-pass of every iteration.  The t-distribution weights add one residual-only
    pass (4 B/px) per LEVEL plus the repeated passes of rejected speculations; neither is counted (the device does not
    report how many there were), so that fraction is a lower bound."""
    pxit = float((np.asarray(iters, np.int64) * np.asarray(px)[None, :]).sum())
    b = B_PX + (2 if depth else 0) + (4 if weights == "huber_mad" else 0)
    return pxit * b, pxit


def roofline_record(stats, px, kernel_ms, weights="none", depth=False, levels=LEVELS, depth_valid=None):
    """depth_valid: fraction of the previous frames' level-0 pixels that have depth.  The kernel walks only those
    (point lists); `achieved` keeps SURVEY 8d's dense count (every pixel of a level, as the reference's planes are
    read), and `frac_touched_pixels_only` says what the fraction is when only walked pixels are counted."""
    iters = stats["iters"][:, :levels]
    ab, pxit = algorithmic_bytes(iters, px, weights, depth)
    peak, peak_src = measured_peak()
    achieved = ab / (kernel_ms / 1e3) / 1e9
    rec = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "kernel": "align_kernel", "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": ab,
           "pixel_iterations_per_launch": pxit, "peak_source": peak_src,
           "gn_iterations_per_pose_mean": float(iters.sum(1).mean())}
    if depth_valid is not None:
        rec["depth_valid_fraction"] = float(depth_valid)
        rec["frac_touched_pixels_only"] = float(depth_valid) * achieved / peak
    t = ncu_traffic_per_pixel_iteration()
    rec["traffic"] = None
    if t and weights == "none" and not depth:
        rec["traffic"] = t["dram_bytes_per_pixel_iteration"] * pxit / 1e9
        rec["traffic_note"] = ("GB per launch: (dram__bytes_read.sum + dram__bytes_write.sum) per pixel-iteration of the "
                               f"same kernel in this round's ncu --set full capture ({t.get('source', 'profiles/r2')}) x the "
                               "pixel-iterations of this launch")
    return rec


def twist_errors(dvo, qt, xi_true):
    return np.array([float(np.abs(dvo.Se3.from_qt(qt[j]).log().reshape(6) - xi_true[j]).max()) for j in range(len(qt))])


def camera_for(dvo, width):
    from dense_visual_odometry_b200.synthetic import TUM_FR1, TUM_DEPTH_SCALE
    s = width / 640.0
    Km = np.array([[TUM_FR1[0] * s, 0, TUM_FR1[2] * s], [0, TUM_FR1[1] * s, TUM_FR1[3] * s], [0, 0, 1]], dtype=np.float32)
    return dvo.RGBDCameraModel(Km, TUM_DEPTH_SCALE)


def time_batch(torch, al, tensors, B, steps, warmup, dev, world=1, gather_total=0):
    """`steps` timed passes (pyramids of both frames + estimate of B pairs [+ the pose gather]) with resident inputs.
    Returns (ms of the timed region, mean kernel ms, launches, qt, stats) of this rank."""
    import torch.distributed as dist
    from dense_visual_odometry_b200.sharding import gather_poses
    bp, dp, bc, dc = (t[:B] for t in tensors)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        al.build(bp, dp, bc, dc)
        qt, st = al.estimate(to_host=False)
        if world > 1:
            gather_poses(qt, gather_total)
    barrier()
    l0 = al.launch_count()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        al.build(bp, dp, bc, dc)
        k_ev[s][0].record()
        qt, st = al.estimate(to_host=False)
        k_ev[s][1].record()
        if world > 1:
            gather_poses(qt, gather_total)   # the path's only collective: [B,7] poses per rank (NCCL)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    return ms, kernel_ms, al.launch_count() - l0, qt, st


def max_over_ranks(torch, x, dev, world):
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def time_e2e(torch, al, hb, B, steps, warmup, dev, world, chunk):
    import torch.distributed as dist
    hbB = [x[:B] for x in hb]
    for _ in range(max(1, min(warmup, 2))):
        al.align(*hbB, chunk_pairs=chunk)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        qt_e, st_e = al.align(*hbB, chunk_pairs=chunk)   # H2D of all four buffers, kernels, D2H of poses + stats, stream sync
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sec = max_over_ranks(torch, time.perf_counter() - t0, dev, world)
    h2d = int(sum(x.numel() * x.element_size() for x in hbB))
    return sec, h2d, int(B * (7 * 4 + 128)), qt_e


def h2d_ceiling(torch, hb, dst, B, dev, world, reps=3):
    """Bare pinned host->device copies of the same four buffers, all ranks at once: what the host memory system /
    PCIe can deliver to this many GPUs together, with no kernel running."""
    import torch.distributed as dist
    best = None
    nbytes = sum(x[:B].numel() * x.element_size() for x in hb)
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for d, s in zip(dst, hb):
            d[:B].copy_(s[:B], non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        dt = max_over_ranks(torch, dt, dev, world)
        best = dt if best is None else min(best, dt)
    return {"per_rank_gbs": nbytes / best / 1e9, "aggregate_gbs": world * nbytes / best / 1e9, "ranks": world,
            "bytes_per_rank": int(nbytes)}


# ---- the other BASELINE.json configs, one GPU each ------------------------------------------------------------
def config_batch(torch, dvo, dev, height, width, levels, pairs, steps, weights="none", depth=False, e2e=True):
    """Resident (+ end-to-end) throughput and roofline of one batch workload."""
    from dense_visual_odometry_b200.synthetic import make_pairs_torch
    cam = camera_for(dvo, width)
    data = make_pairs_torch(range(pairs), dev, height=height, width=width)
    al = dvo.PairBatchAligner(cam, height, width, levels, max_pairs=pairs, weights=weights, use_depth_residual=depth)
    tensors = (data["bgr_prev"], data["depth_prev"], data["bgr_cur"], data["depth_cur"])
    ms, kernel_ms, launches, qt, st = time_batch(torch, al, tensors, pairs, steps, 3, dev)
    stats = dvo.stats_to_numpy(st.cpu().numpy())
    qt_h = qt.cpu().numpy()
    px = level_pixels(height, width, levels)
    err = twist_errors(dvo, qt_h, data["xi"])
    rec = {"workload": f"{pairs} synthetic {width}x{height} pairs, {levels}-level pyramid, weights={weights}"
                       + (", photometric + depth residual" if depth else ""),
           "value": pairs * steps / (ms / 1e3), "unit": UNIT, "steps": steps, "ms_per_step": ms / steps,
           "roofline": roofline_record(stats, px, kernel_ms, weights, depth, levels,
                                       depth_valid=float((tensors[1].view(torch.int16) != 0).float().mean().item())),
           "gpu_launches": int(launches),
           "max_abs_twist_error_vs_truth": float(err.max()), "frac_within_1e-4": float((err < 1e-4).mean()),
           "flags_nonzero": int((stats["flags"] != 0).sum())}
    if e2e:
        hb = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in tensors]
        for hbuf, x in zip(hb, tensors):
            hbuf.copy_(x)
        torch.cuda.synchronize(dev)
        sec, h2d, d2h, qt_e = time_e2e(torch, al, hb, pairs, max(2, steps // 2), 1, dev, 1, 256 if width <= 640 else 148)
        rec["e2e"] = {"value": pairs * max(2, steps // 2) / sec, "unit": UNIT, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "matches_resident": bool(np.array_equal(qt_e, qt_h))}
        del hb
    del al, data, tensors
    gc.collect()
    torch.cuda.empty_cache()
    return rec


def config_single_pair(dvo, torch, dev, weights):
    """configs[1]: ONE synthetic 640x480 pair through the reference's own call, get_dvo(...).step(color, depth)."""
    from dense_visual_odometry_b200.synthetic import make_pairs_numpy
    d = make_pairs_numpy([0], height=H, width=W)
    cam = camera_for(dvo, W)
    est = dvo.get_dvo("robust-dvo", cam, dvo.Se3.identity(), levels=LEVELS, weights=weights)
    lat, kms = [], []
    for i in range(7):
        est.step(d["bgr_prev"][0], d["depth_prev"][0].copy())
        t0 = time.perf_counter()
        T = est.step(d["bgr_cur"][0], d["depth_cur"][0].copy())
        lat.append(1e3 * (time.perf_counter() - t0))
        kms.append(est._h.last_estimate_ms())
    err = float(np.abs(T.log().reshape(6) - d["xi"][0]).max())
    return {"single_pair_step_ms": float(np.median(lat[2:])), "single_pair_kernel_ms": float(np.median(kms[2:])),
            "abs_twist_error_vs_truth": err,
            "note": "one 640x480 pair through get_dvo(...).step(color, depth): H2D of the frame, pyramids, estimate on one "
                    "thread-block cluster (16 CTAs on a B200), D2H of the pose"}


def config_sequence(torch, dvo, dev, frames, weights, reps=3):
    """configs[2]: one synthetic 640x480 stream of `frames` frames through SequenceAligner (every frame's pyramid built
    once, every pair a full coarse-to-fine estimate from the identity guess, as step() does)."""
    from dense_visual_odometry_b200.synthetic import make_sequence
    s = make_sequence(frames, device=dev)
    cam = camera_for(dvo, W)
    seq = dvo.SequenceAligner(cam, H, W, LEVELS, max_frames=frames, weights=weights)
    host = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in (s["bgr"], s["depth"])]
    host[0].copy_(s["bgr"])
    host[1].copy_(s["depth"])
    torch.cuda.synchronize(dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize(dev)
            ts.append(1e3 * (time.perf_counter() - t0))
        return out, float(np.median(ts))

    l0 = seq.launch_count()
    (qt_r, st_r), ms_res = timed(lambda: seq.align(s["bgr"], s["depth"].clone()))
    launches = (seq.launch_count() - l0) // (reps + 1)
    (qt_h, st_h), ms_e2e = timed(lambda: seq.align(host[0], host[1]))
    err = twist_errors(dvo, qt_r, s["xi"])
    px = level_pixels()
    iters = st_r["iters"][:, :LEVELS]
    ab, pxit = algorithmic_bytes(iters, px, weights, False)
    peak, _ = measured_peak()
    n = frames - 1
    rec = {"workload": f"one synthetic 640x480 sequence of {frames} frames, weights={weights}, full coarse-to-fine "
                       f"Gauss-Newton per pair (SequenceAligner)",
           "value": n / (ms_res / 1e3), "unit": UNIT, "ms_resident": ms_res,
           "e2e": {"value": n / (ms_e2e / 1e3), "unit": UNIT, "ms": ms_e2e,
                   "h2d_bytes_per_step": int(host[0].numel() + 2 * host[1].numel()), "d2h_bytes_per_step": int(n * (28 + 128)),
                   "matches_resident": bool(np.array_equal(qt_r, qt_h))},
           "roofline_whole_call": {"achieved": ab / (ms_res / 1e3) / 1e9, "peak": peak, "frac": ab / (ms_res / 1e3) / 1e9 / peak,
                                   "unit": "GB/s", "note": "algorithmic bytes (12 B/px/iteration) / wall time of the whole "
                                                           "call, pyramids and the three streams' overlap included"},
           "gn_iterations_per_pose_mean": float(iters.sum(1).mean()), "gpu_launches": int(launches),
           "max_abs_twist_error_vs_truth": float(err.max()), "flags_nonzero": int((st_r["flags"] != 0).sum())}
    del seq, s, host
    gc.collect()
    torch.cuda.empty_cache()
    return rec


def other_configs(torch, dvo, dev, args):
    out = {}
    t0 = time.perf_counter()
    out["c1_huber"] = config_batch(torch, dvo, dev, H, W, LEVELS, 1184, 3, weights="huber")
    out["c1_huber"]["single_pair"] = config_single_pair(dvo, torch, dev, "huber")
    out["c2_tdist_sequence_1000"] = config_sequence(torch, dvo, dev, 1000, "tdist")
    out["c2_tdist_batch"] = config_batch(torch, dvo, dev, H, W, LEVELS, 1184, 3, weights="tdist", e2e=False)
    out["c4_1280x720_depth"] = config_batch(torch, dvo, dev, 720, 1280, 5, 592, 2, depth=True)
    out["c4_1920x1080_depth"] = config_batch(torch, dvo, dev, 1080, 1920, 5, 592, 2, depth=True)
    out["seconds"] = time.perf_counter() - t0
    return out


def gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    local_cpus = []
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep the version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
        if not args.no_numa_bind:
            # one process per GPU: keep this rank's pinned staging memory (and its copies) on the GPU's own NUMA node
            from dense_visual_odometry_b200.sharding import bind_to_gpu_local_cpus
            local_cpus = bind_to_gpu_local_cpus(local_rank)

    import dense_visual_odometry_b200 as dvo
    from dense_visual_odometry_b200.synthetic import make_pairs_numpy, make_pairs_torch
    from dense_visual_odometry_b200.sharding import shard_range

    total = args.total_pairs
    lo, hi = shard_range(total, rank, world)
    B = hi - lo                                   # strong scaling: this rank's shard of the 4096 pairs
    do_weak = world > 1 and not args.no_weak
    Bw = args.pairs_weak if do_weak else B        # weak scaling: that many pairs on EVERY GPU
    Bmax = max(B, Bw)
    # the CPU baseline is timed on rank 0 at N = 1 only
    n_cpu = min(args.cpu_pairs or min(host_cores(), 32), B) if (rank == 0 and world == 1 and not args.no_cpu) else 0
    cam = camera_for(dvo, W)
    # rank r renders pairs [r * Bmax, (r+1) * Bmax); its strong-scaling shard is the first B of them
    base = rank * Bmax
    data = make_pairs_torch(range(base, base + Bmax), dev, height=H, width=W)
    if n_cpu:   # pairs [0, n_cpu) of rank 0 are rendered with NumPy so the CPU baseline sees bit-identical inputs
        cpu_data = make_pairs_numpy(range(base, base + n_cpu), height=H, width=W)
        for k in ("bgr_prev", "depth_prev", "bgr_cur", "depth_cur"):
            data[k][:n_cpu] = torch.as_tensor(cpu_data[k]).to(dev)
    tensors = (data["bgr_prev"], data["depth_prev"], data["bgr_cur"], data["depth_cur"])

    al = dvo.PairBatchAligner(cam, H, W, LEVELS, max_pairs=Bmax, device=local_rank, weights=args.weights,
                              threads_per_block=args.threads, blocks_per_sm=args.blocks_per_sm,
                              prefetch_rows=args.prefetch_rows, approximate_image2_gradient=args.approximate_gradient,
                              use_depth_residual=args.depth_residual)

    # ---------------- device-resident leg, strong scaling (the headline) ----------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, kernel_ms, launches, qt, st = time_batch(torch, al, tensors, B, args.steps, args.warmup, dev, world, total)
    clocks = sampler.stop() if sampler else None
    ms_max = max_over_ranks(torch, ms, dev, world)
    value = total * args.steps / (ms_max / 1e3)
    qt_h = qt.cpu().numpy()
    stats = dvo.stats_to_numpy(st.cpu().numpy())
    px = level_pixels()
    roof = roofline_record(stats, px, kernel_ms, args.weights, args.depth_residual,
                           depth_valid=float((tensors[1][:B].view(torch.int16) != 0).float().mean().item()))

    # ---------------- end-to-end leg: host (pinned) buffers through the public API ----------------
    hb = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in tensors]
    for hbuf, x in zip(hb, tensors):
        hbuf.copy_(x)
    torch.cuda.synchronize(dev)
    e2e_steps = args.e2e_steps or args.steps
    e2e_s, h2d, d2h, qt_e = time_e2e(torch, al, hb, B, e2e_steps, args.warmup, dev, world, args.chunk_pairs)
    e2e_value = total * e2e_steps / e2e_s
    e2e_match = bool(np.array_equal(qt_e, qt_h))
    ceiling = h2d_ceiling(torch, hb, tensors, B, dev, world)
    pair_bytes = h2d / B
    ceiling["pose_s_at_ceiling"] = ceiling["aggregate_gbs"] * 1e9 / pair_bytes

    # ---------------- weak scaling record (N > 1): 4096 pairs on every GPU ----------------------------
    weak = None
    if do_weak:
        wms, wk, _, wqt, wst = time_batch(torch, al, tensors, Bw, args.steps, 2, dev, world, world * Bw)
        wms = max_over_ranks(torch, wms, dev, world)
        wst_np = dvo.stats_to_numpy(wst.cpu().numpy())
        ws, _, _, _ = time_e2e(torch, al, hb, Bw, max(2, e2e_steps // 2), 1, dev, world, args.chunk_pairs)
        weak = {"scaling": "weak", "pairs_per_gpu": Bw, "global_pairs": world * Bw,
                "value": world * Bw * args.steps / (wms / 1e3), "unit": UNIT, "ms_per_step": wms / args.steps,
                "roofline_frac_rank0": roofline_record(wst_np, px, wk, args.weights, args.depth_residual)["frac"],
                "e2e_value": world * Bw * max(2, e2e_steps // 2) / ws}

    # ---------------- accuracy of the timed workload + CPU baselines (rank 0, N = 1) -----------------------
    pose_err = twist_errors(dvo, qt_h, data["xi"][:B])
    cpu = None
    cpu_port = None
    parity = None
    if rank == 0 and n_cpu:
        cores = host_cores()
        workers = max(1, min(cores, n_cpu))
        pool = cpu_pool(workers)
        try:
            run_port_sample(pool, [base + i for i in range(workers)], args.weights, args.approximate_gradient,
                            args.depth_residual)  # warm-up
            v, res, dt = run_port_sample(pool, [base + i for i in range(n_cpu)], args.weights, args.approximate_gradient,
                                         args.depth_residual)
        finally:
            pool.close()
        dmax = max(float(np.abs(r[1] - qt_h[r[0] - base]).max()) for r in res)
        parity = {"pairs": n_cpu, "max_abs_pose_diff_vs_oracle": dmax, "tolerance": 1e-4, "ok": dmax < 1e-4}
        cpu_port = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
                    "sample": f"{n_cpu} of the timed pairs (seeds {base}..{base + n_cpu - 1}), one process per pair, BLAS 1 "
                              f"thread each, {dt:.1f} s wall; oracle/dvo_oracle.py (NumPy port pinned to the reference)"}
        cpu = cpu_port
        from oracle import reference_runner as R
        if R.available() and args.weights in ("none", "tdist") and not args.depth_residual and not args.approximate_gradient:
            try:
                nref = args.ref_pairs
                v, poses, dt, threads = run_reference_sample([base + i for i in range(nref)],
                                                             use_weighter=(args.weights == "tdist"))
                dref = max(float(np.abs(p - qt_h[i]).max()) for i, p in enumerate(poses) if p is not None)
                cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                       "sample": f"the first {nref} of the timed pairs through the unmodified reference "
                                 f"(get_dvo('robust-dvo').step, numba/BLAS/OpenCV threads = {threads}), one after the other, "
                                 f"{dt:.1f} s wall after a throw-away JIT warm-up pose",
                       "max_abs_pose_diff_vs_b200": dref}
                parity["max_abs_pose_diff_vs_reference"] = dref
                parity["reference_pairs"] = nref
            except Exception as e:   # the port's number stands
                cpu_port["reference_error"] = repr(e)[:200]

    launches_total = int(launches)
    configs = None
    latency = None
    if rank == 0:
        latency = config_single_pair(dvo, torch, dev, args.weights) if not args.depth_residual else None
    del al, hb, tensors
    del data
    gc.collect()
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_configs:
        configs = other_configs(torch, dvo, dev, args)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "total_pairs": total, "pairs_per_gpu": B,
                       "levels": LEVELS, "weights": args.weights, "parallelism": f"pairs sharded over {world} GPU(s)",
                       "host_cpu_binding": (f"rank 0 bound to {len(local_cpus)} GPU-local cores" if local_cpus else "none"),
                       "l2_policy": "inputs larger than L2 (%.2f GB of frames + %.2f GB of pyramids per GPU)" % (
                           2 * B * H * W * 5 / 1e9, B * 437760 * (3 + 8 + 3 + 8) / 1e9),
                       "threads_per_block": args.threads or 128},
            "roofline": roof,
            "cpu_baseline": cpu,
            "cpu_baseline_port": cpu_port if (cpu is not cpu_port) else None,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "steps": e2e_steps, "matches_resident": e2e_match, "h2d_ceiling": ceiling},
            "gpu_launches": launches_total,
            "clocks": clocks,
            "weak_scaling": weak,
            "configs": configs,
            "accuracy": {"max_abs_twist_error_vs_truth": float(pose_err.max()),
                         "frac_within_1e-4": float((pose_err < 1e-4).mean()),
                         "flags_nonzero": int((stats["flags"] != 0).sum())},
            "parity": parity,
            "latency": latency,
        }
        emit_json(out)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner when
    the environment sets NCCL_DEBUG), so file descriptor 1 is pointed at stderr for the whole run and the JSON line
    goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(out):
    line = (json.dumps(out) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--total-pairs", type=int, default=TOTAL_PAIRS,
                    help="strong scaling: frame pairs per step over ALL GPUs (BASELINE.json configs[3])")
    ap.add_argument("--pairs-weak", type=int, default=4096, help="weak-scaling record (N > 1): frame pairs per GPU")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the records of the other BASELINE.json configs")
    ap.add_argument("--weights", default="none", choices=["none", "tdist", "huber", "huber_mad"])
    ap.add_argument("--cpu-pairs", type=int, default=0,
                    help="pairs of the batch also estimated by the CPU port (0 = one per host core, at most 32)")
    ap.add_argument("--ref-pairs", type=int, default=3, help="pairs also estimated by the real reference (about 6 s each)")
    ap.add_argument("--cpu-workers", type=int, default=0)
    ap.add_argument("--port-only", action="store_true", help="--impl reference: time the NumPy port even if the reference is installed")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--prefetch-rows", type=int, default=0)
    ap.add_argument("--approximate-gradient", action="store_true",
                    help="the reference's approximate_image2_gradient=True mode (not the headline configuration)")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="N > 1: do not pin each rank to the CPU cores local to its GPU")
    ap.add_argument("--chunk-pairs", type=int, default=256, help="end-to-end leg: pairs per upload/compute chunk")
    ap.add_argument("--depth-residual", action="store_true",
                    help="photometric + depth residual (extension, BASELINE.json configs[4]; not the headline configuration)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
