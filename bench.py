#!/usr/bin/env python
"""Benchmark of the photometric-alignment hot path (BASELINE.json: pose estimates/s at 640x480, 4 levels).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

A "step" is one pass of the hot path over one batch of synthetic frame pairs: gray conversion + depth
clamp + median pyramids + Sobel planes for both frames of every pair, then the full coarse-to-fine
Gauss-Newton estimate of every pair.  `value` is measured with the frames already resident in HBM; `e2e`
is the same work through the public API with HOST (pinned) buffers, H2D/D2H inside the timed region.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
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
METRIC = "pose estimates/sec at 640x480, 4-level pyramid"
UNIT = "pose/s"
B_PX = 12  # algorithmic bytes per pixel per GN iteration: I1 u8 + D1 u16 + I2 u8 + gx f32 + gy f32


def level_pixels():
    px, h, w = [], H, W
    for _ in range(LEVELS):
        px.append(h * w)
        h, w = (h + 1) // 2, (w + 1) // 2
    return px


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


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_sample(pool, seeds, weights, approx=False, depth=False):
    """Estimates len(seeds) pairs in parallel; returns (pairs/s, results)."""
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, [(s, weights, approx, depth) for s in seeds])
    dt = time.perf_counter() - t0
    return len(seeds) / dt, res, dt


def reference_arm(args):
    """--impl reference: the reference algorithm (oracle port; the Python reference itself cannot travel
    to the GPU box) on all host cores, one pair per process per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    workers = max(1, min(cores, args.cpu_workers or cores))
    pool = cpu_pool(workers)
    times = []
    try:
        for s in range(args.warmup + args.steps):
            seeds = [1000 * s + i for i in range(workers)]
            v, _, dt = run_cpu_sample(pool, seeds, args.weights, args.approximate_gradient, args.depth_residual)
            if s >= args.warmup:
                times.append(dt)
    finally:
        pool.close()
    total = sum(times)
    value = workers * args.steps / total
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "pairs_per_step": workers, "levels": LEVELS,
                   "weights": args.weights},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"{workers} synthetic 640x480 pairs per step (one per process, BLAS 1 thread each), "
                                   f"{args.steps} steps; oracle/dvo_oracle.py (NumPy port pinned to the reference)"},
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
    return (f"batch of independent synthetic 640x480 RGB-D pairs with known SE(3) motion (BASELINE.json configs[1] "
            f"pair type, batched as configs[3]), {LEVELS}-level pyramid, weights={args.weights}"
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


def ncu_traffic_per_pair():
    """DRAM bytes per frame pair of align_kernel from the committed ncu capture (profiles/r1/ncu_traffic.json)."""
    p = ROOT / "profiles" / "r1" / "ncu_traffic.json"
    try:
        return float(json.loads(p.read_text())["dram_bytes_per_pair"])
    except Exception:
        return None


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
    from dense_visual_odometry_b200.synthetic import make_pairs_numpy, make_pairs_torch, TUM_FR1, TUM_DEPTH_SCALE

    B = args.pairs
    # the CPU baseline is timed on rank 0 at N = 1 only
    n_cpu = min(args.cpu_pairs or min(host_cores(), 32), B) if (rank == 0 and world == 1) else 0
    base = rank * B
    # pairs [0, n_cpu) of rank 0 are rendered with NumPy so the CPU baseline sees bit-identical inputs
    Km = np.array([[TUM_FR1[0], 0, TUM_FR1[2]], [0, TUM_FR1[1], TUM_FR1[3]], [0, 0, 1]], dtype=np.float32)
    cam = dvo.RGBDCameraModel(Km, TUM_DEPTH_SCALE)
    data = make_pairs_torch(range(base, base + B), dev, height=H, width=W)
    if n_cpu:
        cpu_data = make_pairs_numpy(range(base, base + n_cpu), height=H, width=W)
        for k in ("bgr_prev", "depth_prev", "bgr_cur", "depth_cur"):
            data[k][:n_cpu] = torch.as_tensor(cpu_data[k]).to(dev)
    bp, dp, bc, dc = data["bgr_prev"], data["depth_prev"], data["bgr_cur"], data["depth_cur"]

    al = dvo.PairBatchAligner(cam, H, W, LEVELS, max_pairs=B, device=local_rank, weights=args.weights,
                              threads_per_block=args.threads, blocks_per_sm=args.blocks_per_sm,
                              prefetch_rows=args.prefetch_rows, approximate_image2_gradient=args.approximate_gradient,
                              use_depth_residual=args.depth_residual)
    from dense_visual_odometry_b200.sharding import gather_poses

    def step_resident():
        al.build(bp, dp, bc, dc)
        qt, st = al.estimate(to_host=False)
        if world > 1:
            gather_poses(qt, world * B)   # the path's only collective: [B,7] poses per rank (NCCL)
        return qt, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident leg -------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = al.launch_count()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        al.build(bp, dp, bc, dc)
        k_ev[s][0].record()
        qt, st = al.estimate(to_host=False)
        k_ev[s][1].record()
        if world > 1:
            gather_poses(qt, world * B)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = al.launch_count() - l0
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1e3)

    qt_h = qt.cpu().numpy()
    stats = dvo.stats_to_numpy(st.cpu().numpy())
    px = level_pixels()
    iters = stats["iters"][:, :LEVELS].astype(np.int64)
    algo_bytes = float((iters * np.array(px)[None, :]).sum() * B_PX)
    extra = 0.0
    if args.weights == "tdist":   # + the residual pre-pass (4 B/px gathers + 4 B/px store) and one scale pass (4 B/px)
        extra = float((iters * np.array(px)[None, :]).sum() * 12)
    elif args.weights == "huber_mad":   # + the residual pre-pass that feeds the median (I1 1 + D1 2 + I2 1 B/px)
        extra = float((iters * np.array(px)[None, :]).sum() * 4)
    if args.depth_residual:   # + D2 u16 per pixel per iteration; the separate pass re-reads D1 (not counted)
        extra += float((iters * np.array(px)[None, :]).sum() * 2)
    peak, peak_src = measured_peak()
    achieved = (algo_bytes + extra) / (kernel_ms / 1e3) / 1e9

    # ---------------- end-to-end leg: host (pinned) buffers through the public API ----------------
    hb = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in (bp, dp, bc, dc)]
    for hbuf, x in zip(hb, (bp, dp, bc, dc)):
        hbuf.copy_(x)
    torch.cuda.synchronize(dev)
    for _ in range(max(1, min(args.warmup, 2))):
        al.align(*hb, chunk_pairs=args.chunk_pairs)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = args.e2e_steps or args.steps
    for _ in range(e2e_steps):
        qt_e, st_e = al.align(*hb, chunk_pairs=args.chunk_pairs)       # H2D of all four buffers, kernels, D2H of poses + stats, stream sync
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t.item())
    h2d = int(sum(x.numel() * x.element_size() for x in hb))
    d2h = int(B * (7 * 4 + 128))
    e2e_match = bool(np.array_equal(qt_e, qt_h))

    # ---------------- accuracy of the timed workload + CPU baseline (rank 0) -----------------------
    xi_true = data["xi"]
    pose_err = []
    for j in range(B):
        T = dvo.Se3.from_qt(qt_h[j])
        pose_err.append(float(np.abs(T.log().reshape(6) - xi_true[j]).max()))
    cpu = None
    parity = None
    if rank == 0 and n_cpu and not args.no_cpu:
        cores = host_cores()
        workers = max(1, min(cores, n_cpu))
        pool = cpu_pool(workers)
        try:
            run_cpu_sample(pool, [base + i for i in range(workers)], args.weights, args.approximate_gradient,
                           args.depth_residual)  # warm-up
            v, res, dt = run_cpu_sample(pool, [base + i for i in range(n_cpu)], args.weights, args.approximate_gradient,
                                        args.depth_residual)
        finally:
            pool.close()
        dmax = max(float(np.abs(r[1] - qt_h[r[0] - base]).max()) for r in res)
        parity = {"pairs": n_cpu, "max_abs_pose_diff_vs_oracle": dmax, "tolerance": 1e-4, "ok": dmax < 1e-4}
        cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
               "sample": f"{n_cpu} of the timed pairs (seeds {base}..{base + n_cpu - 1}), one process per pair, BLAS 1 "
                         f"thread each, {dt:.1f} s wall; oracle/dvo_oracle.py (NumPy port pinned to the reference)"}

    latency = None
    if rank == 0:
        # single-pair latency through the reference's own call, step(color, depth) with host arrays (cluster mode)
        est = dvo.get_dvo("robust-dvo", cam, dvo.Se3.identity(), levels=LEVELS, weights=args.weights,
                          approximate_image2_gradient=args.approximate_gradient,
                          use_depth_residual=args.depth_residual)
        f0 = (bp[0].cpu().numpy(), dp[0].cpu().numpy())
        f1 = (bc[0].cpu().numpy(), dc[0].cpu().numpy())
        lat, kms = [], []
        for i in range(7):
            est.step(f0[0], f0[1].copy())
            t0 = time.perf_counter()
            est.step(f1[0], f1[1].copy())
            lat.append(1e3 * (time.perf_counter() - t0))
            kms.append(est._h.last_estimate_ms())
        latency = {"single_pair_step_ms": float(np.median(lat[2:])), "single_pair_kernel_ms": float(np.median(kms[2:])),
                   "cluster_size": 1 if (args.weights == "huber_mad" or args.depth_residual) else 8,
                   "note": "one 640x480 pair through get_dvo(...).step(color, depth): H2D of the frame, pyramids, "
                           "estimate on one thread-block cluster, D2H of the pose"}
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "pairs_per_gpu": B, "global_pairs": world * B,
                       "levels": LEVELS, "weights": args.weights, "parallelism": f"pairs sharded over {world} GPU(s)",
                       "host_cpu_binding": (f"rank 0 bound to {len(local_cpus)} GPU-local cores" if local_cpus else "none"),
                       "l2_policy": "inputs larger than L2 (%.2f GB of frames + %.2f GB of pyramids per GPU)" % (
                           2 * B * H * W * 5 / 1e9, 2 * B * 437760 * 11 / 1e9),
                       "threads_per_block": args.threads or 128},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (ncu_traffic_per_pair() * B / 1e9) if (ncu_traffic_per_pair() and args.weights == "none") else None,
                         "traffic_note": "GB per launch: dram__bytes_read+write per pair from the ncu --set full capture "
                                         "(profiles/r1/ncu_traffic.json, 1184 pairs) x pairs in this launch",
                         "kernel": "align_kernel", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": algo_bytes + extra, "peak_source": peak_src,
                         "gn_iterations_per_pose_mean": float(iters.sum(1).mean())},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "matches_resident": e2e_match},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "accuracy": {"max_abs_twist_error_vs_truth": float(np.max(pose_err)),
                         "frac_within_1e-4": float(np.mean(np.array(pose_err) < 1e-4)),
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
    ap.add_argument("--pairs", type=int, default=4096, help="frame pairs per GPU per step (BASELINE.json configs[3])")
    ap.add_argument("--weights", default="none", choices=["none", "tdist", "huber", "huber_mad"])
    ap.add_argument("--cpu-pairs", type=int, default=0,
                    help="pairs of the batch also estimated by the CPU oracle (0 = one per host core, at most 32)")
    ap.add_argument("--cpu-workers", type=int, default=0)
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
