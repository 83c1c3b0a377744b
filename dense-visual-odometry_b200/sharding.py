"""Pair sharding over ranks: frame pairs are independent, so rank g of G estimates a
contiguous block of pair indices and the only collective is one final gather of [n,7] poses (+ stats).
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the pair indices owned by `rank`: contiguous blocks, sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n_pairs < 0:
        raise ValueError("n_pairs must be >= 0")
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sequence_shard_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the FRAME indices a rank needs to estimate its share of the n_frames-1 consecutive pairs of a
    sequence: contiguous chunks with a one-frame overlap (pair p aligns frame p against frame p+1)."""
    lo, hi = shard_range(max(n_frames - 1, 0), rank, world)
    return (lo, hi + 1) if hi > lo else (lo, lo)


def gather_poses(local, n_pairs: int, group=None):
    """All ranks contribute their [n_local, C] block (poses qt [.,7], or any per-pair rows); every rank gets the
    [n_pairs, C] table in pair order.  Uses all_gather_into_tensor on equal-size padded blocks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(n_pairs, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {hi - lo}")
    width = -(-n_pairs // world)  # ceil: every block padded to the largest shard
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    rows = []
    for r in range(world):
        a, b = shard_range(n_pairs, r, world)
        rows.append(out[r * width: r * width + (b - a)])
    return torch.cat(rows, dim=0)


def chain_poses(relative_qt, initial=None):
    """Absolute trajectory from relative estimates: cur = cur * T^-1 (base_dense_visual_odometry.py:79)."""
    from .lie import Se3

    cur = initial.copy() if initial is not None else Se3.identity()
    out = [cur.copy()]
    for row in relative_qt:
        cur = cur * Se3.from_qt(row).inverse()
        out.append(cur.copy())
    return out


def bind_to_gpu_local_cpus(device_index: int):
    """Restricts this process to the CPU cores NVML reports as local to GPU `device_index` (same NUMA node / PCIe root),
    so that the pinned staging buffers it allocates afterwards land in that node's memory.  With one process per GPU
    on a multi-socket host this keeps every rank's host-to-device copies off the inter-socket link.  Returns the core
    list, or [] if NVML or the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        allowed = os.sched_getaffinity(0)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1 and (w * 64 + b) in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return []
