"""The callers' side of the path: sequence loading, trajectory writing and
trajectory error metrics, mirroring the reference's benchmark runner `src/test_dvo.py`.

  * TUM RGB-D directories (`rgb.txt`, `depth.txt`, `groundtruth.txt`): parsing and the reference's timestamp
    association (src/test_dvo.py:86-206): every rgb frame takes its nearest depth frame, duplicates are dropped,
    the ground-truth pose nearest to the mean of the two timestamps is attached;
  * the reference's own "test" format (`ground_truth.json` + `camera_intrinsics.yaml`, src/test_dvo.py:209-280);
  * TUM trajectory files `timestamp tx ty tz qx qy qz qw` (src/test_dvo.py:336-345);
  * absolute trajectory error (rigid alignment, Horn) and relative pose error;
  * `run_sequence`: frames -> relative poses -> absolute trajectory, either streaming through
    `RobustDVOB200.step` (what src/test_dvo.py:305-309 does) or as one batch through `SequenceAligner`.

Image decoding uses OpenCV if it is installed; everything else is NumPy.  Nothing here runs on the GPU.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .lie import Se3, So3, pose_to_qt
from .sharding import chain_poses


# ------------------------------------------------------------------------------------------------ TUM lists
def parse_tum_list(path) -> Tuple[np.ndarray, List[List[str]]]:
    """`timestamp field...` lines, '#' comments skipped (src/test_dvo.py:122-147)."""
    ts, rows = [], []
    for line in Path(path).read_text().splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        parts = line.replace(",", " ").split()
        ts.append(float(parts[0]))
        rows.append(parts[1:])
    return np.asarray(ts, dtype=np.float64), rows


def associate_tum(rgb_ts: np.ndarray, depth_ts: np.ndarray, gt_ts: Optional[np.ndarray] = None):
    """The reference's association (src/test_dvo.py:149-169).  Returns (rgb_idx, depth_idx, gt_idx or None)."""
    rgb_ts = np.asarray(rgb_ts, dtype=np.float64)
    depth_ts = np.asarray(depth_ts, dtype=np.float64)
    closest = np.abs(rgb_ts.reshape(-1, 1) - depth_ts.reshape(1, -1)).argmin(axis=1)
    depth_idx, rgb_idx = np.unique(closest, return_index=True)   # first rgb frame claiming each depth frame
    gt_idx = None
    if gt_ts is not None and len(gt_ts):
        frame_ts = 0.5 * (rgb_ts[rgb_idx] + depth_ts[depth_idx])
        gt_idx = np.abs(frame_ts.reshape(-1, 1) - np.asarray(gt_ts, dtype=np.float64).reshape(1, -1)).argmin(axis=1)
    return rgb_idx, depth_idx, gt_idx


def tum_pose_to_qt(fields: Sequence[str]) -> np.ndarray:
    """`tx ty tz qx qy qz qw` -> [qw qx qy qz tx ty tz] (the reference rolls the quaternion, src/test_dvo.py:138-141)."""
    v = np.asarray([float(x) for x in fields[:7]], dtype=np.float32)
    return np.concatenate([[v[6]], v[3:6], v[0:3]]).astype(np.float32)


def load_tum_sequence(data_dir, size: Optional[int] = None) -> Dict:
    """Associated frame list of a TUM RGB-D directory (no pixels are read here)."""
    d = Path(data_dir)
    for name in ("rgb.txt", "depth.txt"):
        if not (d / name).exists():
            raise FileNotFoundError(f"Expected TUM RGB-D dataset to contain a file named '{name}' at '{d}'")
    rgb_ts, rgb_rows = parse_tum_list(d / "rgb.txt")
    dep_ts, dep_rows = parse_tum_list(d / "depth.txt")
    gt_ts, gt_rows = (parse_tum_list(d / "groundtruth.txt") if (d / "groundtruth.txt").exists()
                      else (np.zeros(0), []))
    ri, di, gi = associate_tum(rgb_ts, dep_ts, gt_ts if len(gt_ts) else None)
    if size is not None:
        ri, di = ri[:size], di[:size]
        gi = gi[:size] if gi is not None else None
    out = {"type": "TUM", "rgb": [str(d / rgb_rows[i][0]) for i in ri], "depth": [str(d / dep_rows[i][0]) for i in di],
           "rgb_timestamps": rgb_ts[ri], "depth_timestamps": dep_ts[di]}
    if gi is not None:
        out["timestamps"] = gt_ts[gi]
        out["gt_qt"] = np.stack([tum_pose_to_qt(gt_rows[i]) for i in gi])
    else:
        out["timestamps"] = 0.5 * (rgb_ts[ri] + dep_ts[di])
        out["gt_qt"] = None
    return out


def load_test_sequence(data_dir, size: Optional[int] = None) -> Dict:
    """The reference's custom format (src/test_dvo.py:209-280): ground_truth.json maps frame ids to
    {"rgb", "depth", "transformation"}; the transformation is a 4x4 camera-to-world matrix in the shipped test
    data (a 6-vector twist is accepted too)."""
    d = Path(data_dir)
    data = json.loads((d / "ground_truth.json").read_text())
    keys = sorted(data, key=lambda k: int(k) if str(k).isdigit() else k)
    if size is not None:
        keys = keys[:size]
    rgb, depth, gt = [], [], []
    for k in keys:
        v = data[k]
        rgb.append(str(d / v["rgb"]))
        depth.append(str(d / v["depth"]))
        if gt is not None and "transformation" in v:
            a = np.asarray(v["transformation"], dtype=np.float64)
            if a.size == 16:
                M = a.reshape(4, 4)
                gt.append(pose_to_qt(Se3(So3(M[:3, :3].copy()), M[:3, 3].reshape(3, 1).astype(np.float32))))
            else:
                gt.append(pose_to_qt(Se3.from_se3(a.reshape(6, 1).astype(np.float32))))
        else:
            gt = None
    return {"type": "TEST", "rgb": rgb, "depth": depth, "timestamps": np.arange(len(rgb), dtype=np.float64),
            "gt_qt": np.stack(gt) if gt else None, "camera_intrinsics": str(d / "camera_intrinsics.yaml")}


def read_frames(seq: Dict, bgr: bool = True):
    """Decodes the images of a sequence dict: ([N,H,W,3] u8, [N,H,W] u16).  `step()` assumes BGR channel order
    (base_dense_visual_odometry.py:58); the reference's runner hands it RGB (src/test_dvo.py:183, a bit-rot the
    survey records) -- pass bgr=False to reproduce that."""
    import cv2
    colors, depths = [], []
    for c, z in zip(seq["rgb"], seq["depth"]):
        img = cv2.imread(c, cv2.IMREAD_ANYCOLOR)
        dep = cv2.imread(z, cv2.IMREAD_UNCHANGED)
        if img is None or dep is None:
            raise FileNotFoundError(f"could not read '{c}' / '{z}'")
        colors.append(img if bgr else cv2.cvtColor(img, cv2.COLOR_BGR2RGB))
        depths.append(dep.astype(np.uint16))
    return np.stack(colors), np.stack(depths)


# ------------------------------------------------------------------------------------------------ trajectories
def write_tum_trajectory(path, timestamps, poses) -> None:
    """`# timestamp tx ty tz qx qy qz qw` (src/test_dvo.py:336-345).  poses: Se3 objects or [.,7] qt rows."""
    with Path(path).open("w") as fp:
        fp.write("# timestamp tx ty tz qx qy qz qw\n")
        for ts, p in zip(timestamps, poses):
            qt = np.asarray(p, dtype=np.float64).reshape(-1) if not hasattr(p, "tvec") else pose_to_qt(p).astype(np.float64)
            fp.write(" ".join(repr(float(x)) for x in (ts, qt[4], qt[5], qt[6], qt[1], qt[2], qt[3], qt[0])) + "\n")


def read_tum_trajectory(path) -> Tuple[np.ndarray, np.ndarray]:
    ts, rows = parse_tum_list(path)
    return ts, (np.stack([tum_pose_to_qt(r) for r in rows]) if rows else np.zeros((0, 7), np.float32))


def align_rigid(src: np.ndarray, dst: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Least-squares rotation + translation taking src points [N,3] onto dst (Horn / Kabsch, no scale)."""
    src, dst = np.asarray(src, np.float64), np.asarray(dst, np.float64)
    cs, cd = src.mean(0), dst.mean(0)
    U, _, Vt = np.linalg.svd((dst - cd).T @ (src - cs))
    S = np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt))])
    R = U @ S @ Vt
    return R, cd - R @ cs


def ate_rmse(est_xyz: np.ndarray, gt_xyz: np.ndarray, align: bool = True) -> float:
    """Absolute trajectory error (RMSE of positions after rigid alignment)."""
    est_xyz, gt_xyz = np.asarray(est_xyz, np.float64), np.asarray(gt_xyz, np.float64)
    if align and len(est_xyz) >= 3:
        R, t = align_rigid(est_xyz, gt_xyz)
        est_xyz = est_xyz @ R.T + t
    return float(np.sqrt(((est_xyz - gt_xyz) ** 2).sum(1).mean()))


def rpe_translation(est: Sequence[Se3], gt: Sequence[Se3], delta: int = 1) -> float:
    """Relative pose error, translational RMSE over frame pairs `delta` apart."""
    errs = []
    for i in range(len(est) - delta):
        de = est[i].inverse() * est[i + delta]
        dg = gt[i].inverse() * gt[i + delta]
        errs.append(float(np.linalg.norm((dg.inverse() * de).tvec)))
    return float(np.sqrt(np.mean(np.square(errs)))) if errs else float("nan")


# ------------------------------------------------------------------------------------------------ runner
def run_sequence(colors: np.ndarray, depths: np.ndarray, camera_model, levels: int = 4, initial_pose: Optional[Se3] = None,
                 batch: bool = True, gt_qt: Optional[np.ndarray] = None, **dvo_kwargs) -> Dict:
    """Estimates a whole sequence.  Returns the fields of the reference's report (src/test_dvo.py:327-329):
    estimated_transforms (twists of the relative poses; identity for frame 0), estimated_poses (twists of the
    absolute poses), errors (distance to the ground-truth position, or "N/A"), plus `trajectory` (Se3 list)."""
    from .estimator import RobustDVOB200, SequenceAligner

    n = colors.shape[0]
    init = initial_pose if initial_pose is not None else Se3.identity()
    failed = []                       # frames without a pose of their own (they repeat the previous one)
    if batch:
        h, w = depths.shape[1:]
        seq = SequenceAligner(camera_model, h, w, levels, max_frames=n, **dvo_kwargs)
        rel, stats = seq.align(colors, depths)
        rel = [pose_to_qt(Se3.identity())] + [r for r in rel]
        # A pair whose estimate is not finite: step() returns None for it, KEEPS the previous frame and aligns the next
        # frame against that one (base_dense_visual_odometry.py:75-85).  The same here: the frame gets the identity
        # and the following frame is re-estimated against the last good one (its pyramids are still resident).
        good = 0                      # last frame with a pose
        for f in range(1, n):
            if good == f - 1:         # the batch estimated exactly this pair
                q = rel[f] if (np.all(np.isfinite(rel[f])) and not int(stats["flags"][f - 1]) & 1) else None
            else:
                q = seq.estimate_pair(good, f)
            if q is not None:
                rel[f] = q
                good = f
            else:
                failed.append(f)
                rel[f] = pose_to_qt(Se3.identity())
    else:
        dvo = RobustDVOB200(camera_model, init, levels, **dvo_kwargs)
        rel = []
        for i in range(n):
            T = dvo.step(colors[i], depths[i])
            if T is None:
                failed.append(i)
            rel.append(pose_to_qt(T if T is not None else Se3.identity()))
    traj = chain_poses(rel[1:], init)
    errors = []
    for i, p in enumerate(traj):
        if gt_qt is not None:
            errors.append(float(np.linalg.norm(p.tvec.reshape(3) - np.asarray(gt_qt[i][4:], dtype=np.float64))))
        else:
            errors.append("N/A")
    return {"estimated_transforms": [Se3.from_qt(r).log().reshape(-1).tolist() for r in rel],
            "estimated_poses": [p.log().reshape(-1).tolist() for p in traj], "errors": errors, "trajectory": traj,
            "failed_frames": failed}
