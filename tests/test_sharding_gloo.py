"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: pair sharding and the final pose gather."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    import dense_visual_odometry_b200  # noqa: F401
    from dense_visual_odometry_b200.sharding import gather_poses, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_pairs, rank, world)
    # every pair's "pose" encodes its global index, so the gathered order can be checked
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones((1, 7)) + torch.arange(7) * 0.125
    full = gather_poses(local, n_pairs)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the timing reduction bench.py uses
    np.save(Path(out_dir) / f"r{rank}.npy", full.numpy())
    np.save(Path(out_dir) / f"t{rank}.npy", t.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [8, 7, 1])
def test_gather_poses_world2(tmp_path, n_pairs):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_pairs, str(tmp_path)), nprocs=2, join=True)
    want = np.arange(n_pairs, dtype=np.float32)[:, None] * np.ones((1, 7), np.float32) + np.arange(7) * 0.125
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"r{r}.npy"), want)
        assert np.load(tmp_path / f"t{r}.npy")[0] == 2.0


def test_shard_ranges_cover_and_are_disjoint():
    import dense_visual_odometry_b200  # noqa: F401
    from dense_visual_odometry_b200.sharding import sequence_shard_range, shard_range
    for n in (0, 1, 5, 4096, 4097):
        for world in (1, 2, 3, 8):
            got = [shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            assert max(h - l for l, h in got) - min(h - l for l, h in got) <= 1
    # 10 frames = 9 pairs over 2 ranks: frames [0,6) and [5,10): one frame of overlap
    assert sequence_shard_range(10, 0, 2) == (0, 6)
    assert sequence_shard_range(10, 1, 2) == (5, 10)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_chain_poses_matches_reference_rule():
    import dense_visual_odometry_b200 as m
    from dense_visual_odometry_b200.sharding import chain_poses
    xi = np.array([0.01, -0.02, 0.03, 0.004, -0.003, 0.002], dtype=np.float32).reshape(6, 1)
    T = m.Se3.from_se3(xi)
    qt = m.pose_to_qt(T)
    traj = chain_poses([qt, qt])
    want = m.Se3.identity() * T.inverse() * T.inverse()
    np.testing.assert_allclose(m.pose_to_qt(traj[-1]), m.pose_to_qt(want), atol=1e-7)
    assert len(traj) == 3
