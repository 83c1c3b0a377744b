"""CPU tests of the sequence / trajectory harness (reference: src/test_dvo.py)."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import dense_visual_odometry_b200 as m  # noqa: E402
from dense_visual_odometry_b200 import dataset as D  # noqa: E402


def test_tum_association_follows_the_reference(tmp_path):
    # rgb frames at 0.00, 0.03, 0.07, 0.10; depth at 0.01, 0.08, 0.11 -> rgb 0 and 1 both pick depth 0 (the first
    # claim wins, src/test_dvo.py:154-155), rgb 2 -> depth 1, rgb 3 -> depth 2
    rgb_ts = np.array([0.00, 0.03, 0.07, 0.10])
    dep_ts = np.array([0.01, 0.08, 0.11])
    gt_ts = np.array([0.0, 0.02, 0.05, 0.075, 0.1, 0.2])
    ri, di, gi = D.associate_tum(rgb_ts, dep_ts, gt_ts)
    assert ri.tolist() == [0, 2, 3] and di.tolist() == [0, 1, 2]
    # frame timestamps 0.005, 0.075, 0.105 -> nearest ground truth
    assert gi.tolist() == [0, 3, 4]
    (tmp_path / "rgb.txt").write_text("# color images\n" + "".join(f"{t:.2f} rgb/{i}.png\n" for i, t in enumerate(rgb_ts)))
    (tmp_path / "depth.txt").write_text("# depth\n" + "".join(f"{t:.2f} depth/{i}.png\n" for i, t in enumerate(dep_ts)))
    (tmp_path / "groundtruth.txt").write_text(
        "# ts tx ty tz qx qy qz qw\n" + "".join(f"{t} {i} 0 0 0 0 0 1\n" for i, t in enumerate(gt_ts)))
    seq = D.load_tum_sequence(tmp_path)
    assert [Path(p).name for p in seq["rgb"]] == ["0.png", "2.png", "3.png"]
    assert [Path(p).name for p in seq["depth"]] == ["0.png", "1.png", "2.png"]
    np.testing.assert_array_equal(seq["gt_qt"][:, 4], [0, 3, 4])        # tx encodes the ground-truth row
    np.testing.assert_array_equal(seq["gt_qt"][:, 0], [1, 1, 1])        # qw moved to the front
    assert D.load_tum_sequence(tmp_path, size=2)["timestamps"].shape == (2,)
    with pytest.raises(FileNotFoundError):
        D.load_tum_sequence(tmp_path / "missing")


def test_trajectory_round_trip_and_metrics(tmp_path):
    rng = np.random.default_rng(1)
    rel = [m.pose_to_qt(m.Se3.from_se3(rng.uniform(-0.05, 0.05, (6, 1)).astype(np.float32))) for _ in range(12)]
    traj = m.chain_poses(rel)
    ts = np.arange(len(traj)) * 0.033 + 1305031102.175304
    D.write_tum_trajectory(tmp_path / "traj.txt", ts, traj)
    ts2, qt2 = D.read_tum_trajectory(tmp_path / "traj.txt")
    np.testing.assert_allclose(ts2, ts, rtol=0, atol=1e-6)
    np.testing.assert_allclose(qt2, np.stack([m.pose_to_qt(p) for p in traj]), atol=1e-6)
    # ATE is invariant to a rigid transform of the estimate; RPE of a trajectory against itself is zero
    xyz = np.stack([p.tvec.reshape(3) for p in traj]).astype(np.float64)
    Rz = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    assert D.ate_rmse(xyz @ Rz.T + np.array([1.0, -2.0, 0.5]), xyz) < 1e-9
    assert D.ate_rmse(xyz + np.array([0.1, 0.0, 0.0]), xyz, align=False) == pytest.approx(0.1)
    assert D.rpe_translation(traj, traj) == pytest.approx(0.0, abs=1e-7)
    drift = m.chain_poses([r + np.array([0, 0, 0, 0, 0.01, 0, 0], np.float32) for r in rel])
    assert D.rpe_translation(drift, traj) > 1e-3


def test_test_format_loader(tmp_path, testdata_frames):
    cv2 = pytest.importorskip("cv2")
    f = testdata_frames
    (tmp_path / "rgb").mkdir()
    (tmp_path / "depth").mkdir()
    gt = {}
    T = np.eye(4)
    T[:3, 3] = [0.1, 0.2, 0.3]
    for i in range(3):
        cv2.imwrite(str(tmp_path / "rgb" / f"{i + 1}.png"), f["bgr"][i])
        cv2.imwrite(str(tmp_path / "depth" / f"{i + 1}.png"), f["depth"][i])
        gt[str(i + 1)] = {"rgb": f"rgb/{i + 1}.png", "depth": f"depth/{i + 1}.png", "transformation": T.tolist()}
    (tmp_path / "ground_truth.json").write_text(json.dumps(gt))
    seq = D.load_test_sequence(tmp_path)
    assert len(seq["rgb"]) == 3 and seq["gt_qt"].shape == (3, 7)
    np.testing.assert_allclose(seq["gt_qt"][0], [1, 0, 0, 0, 0.1, 0.2, 0.3], atol=1e-7)
    colors, depths = D.read_frames(seq)
    np.testing.assert_array_equal(colors, f["bgr"][:3])      # PNG is lossless: the exact frames come back, BGR
    np.testing.assert_array_equal(depths, f["depth"][:3])
    assert D.read_frames(seq, bgr=False)[0][0, 0, 0, 0] == f["bgr"][0][0, 0, 2]


def test_synthetic_sequence_is_consistent_and_deterministic():
    """BASELINE.json configs[2] generator: the relative twists chain to the absolute poses the frames were rendered
    from, frame 0 is the identity view, and the stream is reproducible."""
    from dense_visual_odometry_b200.synthetic import make_sequence, make_sequence_motions, se3_exp
    xi, R_abs, t_abs = make_sequence_motions(6, seed=5)
    assert xi.shape == (5, 6) and R_abs.shape == (6, 3, 3) and t_abs.shape == (6, 3)
    assert np.allclose(R_abs[0], np.eye(3)) and not t_abs[0].any()
    R, t = np.eye(3), np.zeros(3)
    for k in range(5):
        Rk, tk = se3_exp(xi[k])
        R, t = Rk @ R, Rk @ t + tk
        assert np.allclose(R, R_abs[k + 1]) and np.allclose(t, t_abs[k + 1])
    assert np.abs(xi[:, :3]).max() <= 0.02 and np.abs(xi[:, 3:]).max() <= 0.01
    a = make_sequence(3, height=48, width=64)
    b = make_sequence(3, height=48, width=64)
    assert a["bgr"].shape == (3, 48, 64, 3) and a["depth"].shape == (3, 48, 64) and a["depth"].dtype == np.uint16
    assert np.array_equal(a["bgr"], b["bgr"]) and np.array_equal(a["depth"], b["depth"])
    assert (a["depth"] == 0).mean() > 0.02          # every frame has depth holes
    assert not np.array_equal(a["bgr"][0], a["bgr"][1])


def test_gpu_local_cpu_binding_is_harmless_without_nvml():
    """The N > 1 host helper must never raise or shrink the affinity to nothing on a box without a GPU."""
    import os
    from dense_visual_odometry_b200.sharding import bind_to_gpu_local_cpus
    before = os.sched_getaffinity(0)
    cpus = bind_to_gpu_local_cpus(0)
    after = os.sched_getaffinity(0)
    assert isinstance(cpus, list)
    assert after == before or (cpus and set(cpus) == after)
    os.sched_setaffinity(0, before)
