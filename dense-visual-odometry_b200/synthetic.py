"""Synthetic RGB-D frame pairs with exactly known SE(3) motion.

A tilted textured plane is rendered analytically (ray/plane intersection) from the previous
camera (identity) and from the current camera, so intensity and depth are exact for any motion
and resolution.  Scene parameters come from `numpy.random.default_rng(seed)`; rendering runs in
NumPy or, for large batches that should never cross PCIe, in torch on the GPU.

The motion returned is the transform the estimator recovers: X_cur = R X_prev + t, i.e. the
reference's `T_{t-1 -> t}` (core/base_dense_visual_odometry.py:72-79 in the reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

TUM_FR1 = (517.3, 516.5, 318.6, 239.5)   # tests/test_data/camera_intrinsics.yaml of the reference
TUM_DEPTH_SCALE = 0.0002


@dataclass
class Scene:
    """Parameters of one synthetic pair (all small host arrays)."""
    normal: np.ndarray      # (3,) unit plane normal in the previous camera frame
    offset: float           # plane: normal . X = offset
    e1: np.ndarray          # (3,) in-plane basis
    e2: np.ndarray
    freq: np.ndarray        # (S,2) cycles per metre along e1/e2
    phase: np.ndarray       # (S,)
    amp: np.ndarray         # (S,)
    R: np.ndarray           # (3,3) motion rotation
    t: np.ndarray           # (3,) motion translation
    xi: np.ndarray          # (6,) twist [v; w] whose exponential is (R, t)
    holes: np.ndarray       # (H/8+1, W/8+1) bool, True = depth hole in that 8x8 block (both frames differ)
    holes2: np.ndarray


def so3_exp(phi: np.ndarray) -> np.ndarray:
    th = float(np.linalg.norm(phi))
    K = np.array([[0, -phi[2], phi[1]], [phi[2], 0, -phi[0]], [-phi[1], phi[0], 0]], dtype=np.float64)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + (math.sin(th) / th) * K + ((1 - math.cos(th)) / th ** 2) * (K @ K)


def se3_exp(xi: np.ndarray):
    """Twist [v; w] -> (R, t) with t = V(w) v (Barfoot's convention, as the reference's Se3.from_se3)."""
    v = np.asarray(xi[:3], dtype=np.float64)
    w = np.asarray(xi[3:], dtype=np.float64)
    th = float(np.linalg.norm(w))
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    R = so3_exp(w)
    if th < 1e-12:
        V = np.eye(3) + 0.5 * K
    else:
        V = np.eye(3) + ((1 - math.cos(th)) / th ** 2) * K + ((th - math.sin(th)) / th ** 3) * (K @ K)
    return R, V @ v


def make_scene(seed: int, height: int, width: int, fx: float, n_sin: int = 8, trans_mag: float = 0.02,
               rot_mag: float = 0.01, hole_frac: float = 0.10) -> Scene:
    rng = np.random.default_rng(seed)
    n = np.array([0.3, -0.15, -1.0]) + rng.uniform(-0.05, 0.05, 3) * np.array([1, 1, 0])
    n /= np.linalg.norm(n)
    offset = -rng.uniform(2.2, 2.6)
    e1 = np.cross(n, [0.0, 1.0, 0.0])
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(n, e1)
    # wavelengths of 8..128 level-0 pixels of a 640-wide image at ~2.4 m, scaled with resolution
    px = 2.4 / fx
    wl = px * np.exp(rng.uniform(np.log(8.0), np.log(128.0), n_sin)) * (width / 640.0)
    ang = rng.uniform(0, np.pi, n_sin)
    freq = np.stack([np.cos(ang), np.sin(ang)], axis=1) / wl[:, None]
    phase = rng.uniform(0, 2 * np.pi, n_sin)
    amp = rng.uniform(0.5, 1.0, n_sin) * np.sqrt(wl / wl.max())
    amp /= amp.sum()
    xi = np.concatenate([rng.uniform(-trans_mag, trans_mag, 3), rng.uniform(-rot_mag, rot_mag, 3)])
    R, t = se3_exp(xi)
    hb, wb = (height + 7) // 8, (width + 7) // 8
    holes = rng.random((hb, wb)) < hole_frac
    holes2 = rng.random((hb, wb)) < hole_frac
    return Scene(n, offset, e1, e2, freq, phase, amp, R, t, xi, holes, holes2)


def _render(xp, scene_arrays, height, width, K, depth_scale, device=None):
    """Render previous and current frames for a stack of scenes.  xp is numpy or torch."""
    is_torch = xp.__name__ == "torch"
    fx, fy, cx, cy = K

    def arr(a):
        if is_torch:
            return xp.as_tensor(np.asarray(a), dtype=xp.float32, device=device)
        return np.asarray(a, dtype=np.float32)

    n, off, e1, e2, freq, phase, amp, R, t = [arr(a) for a in scene_arrays]
    B = n.shape[0]
    if is_torch:
        u = xp.arange(width, dtype=xp.float32, device=device)
        v = xp.arange(height, dtype=xp.float32, device=device)
    else:
        u = np.arange(width, dtype=np.float32)
        v = np.arange(height, dtype=np.float32)
    xn = ((u - cx) / fx)[None, None, :]          # (1,1,W)
    yn = ((v - cy) / fy)[None, :, None]          # (1,H,1)

    def tex(X, Y, Z):
        a = X * e1[:, 0, None, None] + Y * e1[:, 1, None, None] + Z * e1[:, 2, None, None]
        b = X * e2[:, 0, None, None] + Y * e2[:, 1, None, None] + Z * e2[:, 2, None, None]
        acc = xp.zeros_like(a)
        for k in range(freq.shape[1]):
            arg = 2 * math.pi * (a * freq[:, k, 0, None, None] + b * freq[:, k, 1, None, None]) \
                + phase[:, k, None, None]
            acc = acc + amp[:, k, None, None] * xp.sin(arg)
        g = 127.5 + 107.5 * acc
        g = xp.clip(xp.floor(g + 0.5), 0, 255)
        return g

    def to_dn(Z):
        dn = xp.floor(Z / depth_scale + 0.5)
        return xp.clip(dn, 0, 65535)

    # previous frame: camera at identity, X = s * (xn, yn, 1), n.X = off
    denom = n[:, 0, None, None] * xn + n[:, 1, None, None] * yn + n[:, 2, None, None]
    s1 = off[:, None, None] / denom
    g1 = tex(xn * s1, yn * s1, s1)
    d1 = to_dn(s1)
    # current frame: X_prev = R^T (s v - t)
    Rn = xp.einsum("bij,bj->bi", R, n) if is_torch else np.einsum("bij,bj->bi", R, n)
    nRt = (Rn * t).sum(-1)                       # n . R^T t == (R n) . t
    denom2 = Rn[:, 0, None, None] * xn + Rn[:, 1, None, None] * yn + Rn[:, 2, None, None]
    s2 = (off + nRt)[:, None, None] / denom2
    Xc = xn * s2 - t[:, 0, None, None]
    Yc = yn * s2 - t[:, 1, None, None]
    Zc = s2 - t[:, 2, None, None]
    Xp = R[:, 0, 0, None, None] * Xc + R[:, 1, 0, None, None] * Yc + R[:, 2, 0, None, None] * Zc
    Yp = R[:, 0, 1, None, None] * Xc + R[:, 1, 1, None, None] * Yc + R[:, 2, 1, None, None] * Zc
    Zp = R[:, 0, 2, None, None] * Xc + R[:, 1, 2, None, None] * Yc + R[:, 2, 2, None, None] * Zc
    g2 = tex(Xp, Yp, Zp)
    d2 = to_dn(s2)
    return g1, d1, g2, d2, B


def _stack(scenes):
    return (np.stack([s.normal for s in scenes]), np.array([s.offset for s in scenes]),
            np.stack([s.e1 for s in scenes]), np.stack([s.e2 for s in scenes]),
            np.stack([s.freq for s in scenes]), np.stack([s.phase for s in scenes]),
            np.stack([s.amp for s in scenes]), np.stack([s.R for s in scenes]), np.stack([s.t for s in scenes]))


def _hole_mask_np(scenes, height, width, which):
    m = np.stack([np.kron(getattr(s, which), np.ones((8, 8), dtype=bool))[:height, :width] for s in scenes])
    return m


def make_pairs_numpy(seeds, height=480, width=640, K=TUM_FR1, depth_scale=TUM_DEPTH_SCALE, **scene_kw):
    """Returns dict with bgr_prev/bgr_cur (B,H,W,3) u8, depth_prev/depth_cur (B,H,W) u16, xi (B,6), R, t."""
    scale = width / 640.0
    Ks = (K[0] * scale, K[1] * scale, K[2] * scale, K[3] * scale) if scene_kw.pop("scale_K", True) else K
    scenes = [make_scene(int(s), height, width, Ks[0], **scene_kw) for s in seeds]
    g1, d1, g2, d2, _ = _render(np, _stack(scenes), height, width, Ks, depth_scale)
    d1 = d1.astype(np.uint16)
    d2 = d2.astype(np.uint16)
    d1[_hole_mask_np(scenes, height, width, "holes")] = 0
    d2[_hole_mask_np(scenes, height, width, "holes2")] = 0
    g1 = g1.astype(np.uint8)
    g2 = g2.astype(np.uint8)
    return dict(bgr_prev=np.repeat(g1[..., None], 3, axis=-1), depth_prev=d1,
                bgr_cur=np.repeat(g2[..., None], 3, axis=-1), depth_cur=d2,
                xi=np.stack([s.xi for s in scenes]).astype(np.float64),
                R=np.stack([s.R for s in scenes]), t=np.stack([s.t for s in scenes]), K=Ks,
                depth_scale=depth_scale)


def make_pairs_torch(seeds, device, height=480, width=640, K=TUM_FR1, depth_scale=TUM_DEPTH_SCALE, chunk=32,
                     **scene_kw):
    """Same scenes rendered with torch on `device`; tensors stay on the device."""
    import torch

    scale = width / 640.0
    Ks = (K[0] * scale, K[1] * scale, K[2] * scale, K[3] * scale) if scene_kw.pop("scale_K", True) else K
    seeds = list(seeds)
    B = len(seeds)
    bgr_prev = torch.empty((B, height, width, 3), dtype=torch.uint8, device=device)
    bgr_cur = torch.empty((B, height, width, 3), dtype=torch.uint8, device=device)
    depth_prev = torch.empty((B, height, width), dtype=torch.uint16, device=device)
    depth_cur = torch.empty((B, height, width), dtype=torch.uint16, device=device)
    xis, Rs, ts = [], [], []
    for c0 in range(0, B, chunk):
        scenes = [make_scene(int(s), height, width, Ks[0], **scene_kw) for s in seeds[c0:c0 + chunk]]
        g1, d1, g2, d2, _ = _render(torch, _stack(scenes), height, width, Ks, depth_scale, device=device)
        h1 = torch.as_tensor(_hole_mask_np(scenes, height, width, "holes"), device=device)
        h2 = torch.as_tensor(_hole_mask_np(scenes, height, width, "holes2"), device=device)
        d1 = d1.to(torch.int32).masked_fill(h1, 0)
        d2 = d2.to(torch.int32).masked_fill(h2, 0)
        n = len(scenes)
        bgr_prev[c0:c0 + n] = g1.to(torch.uint8)[..., None]
        bgr_cur[c0:c0 + n] = g2.to(torch.uint8)[..., None]
        depth_prev[c0:c0 + n] = d1.to(torch.uint16)
        depth_cur[c0:c0 + n] = d2.to(torch.uint16)
        xis += [s.xi for s in scenes]
        Rs += [s.R for s in scenes]
        ts += [s.t for s in scenes]
    return dict(bgr_prev=bgr_prev, depth_prev=depth_prev, bgr_cur=bgr_cur, depth_cur=depth_cur,
                xi=np.stack(xis).astype(np.float64), R=np.stack(Rs), t=np.stack(ts), K=Ks, depth_scale=depth_scale)


def make_sequence_motions(n_frames: int, seed: int = 1000, trans_mag: float = 0.02, rot_mag: float = 0.01,
                          smooth: float = 0.8):
    """Frame-to-frame twists of a smooth trajectory: AR(1)-smoothed draws of the same
    magnitude as the pair generator's.  Returns (xi_rel [n-1,6], R_abs [n,3,3], t_abs [n,3]) with
    X_k = R_abs[k] X_0 + t_abs[k] and X_{k+1} = exp(xi_rel[k]) X_k."""
    rng = np.random.default_rng(seed)
    mag = np.concatenate([np.full(3, trans_mag), np.full(3, rot_mag)])
    xi = rng.uniform(-1, 1, 6) * mag
    xis = []
    R_abs, t_abs = [np.eye(3)], [np.zeros(3)]
    for _ in range(n_frames - 1):
        xi = smooth * xi + (1 - smooth) * rng.uniform(-1, 1, 6) * mag
        R, t = se3_exp(xi)
        xis.append(xi.copy())
        R_abs.append(R @ R_abs[-1])
        t_abs.append(R @ t_abs[-1] + t)
    return np.array(xis).reshape(-1, 6), np.stack(R_abs), np.stack(t_abs)


def make_sequence(n_frames: int, device=None, height=480, width=640, K=TUM_FR1, depth_scale=TUM_DEPTH_SCALE,
                  scene_seed: int = 1000, motion_seed: int = 1000, chunk: int = 32, hole_frac: float = 0.10,
                  **motion_kw):
    """A stream of `n_frames` views of ONE textured plane along a smooth trajectory (BASELINE.json configs[2]).
    device=None renders with NumPy (host arrays), otherwise with torch on that device (tensors stay there).
    Returns dict(bgr [N,H,W,3] u8, depth [N,H,W] u16, xi [N-1,6] = the twist taking frame k's camera to frame k+1's,
    K, depth_scale).  Every frame gets its own 8x8-block depth holes."""
    scale = width / 640.0
    Ks = (K[0] * scale, K[1] * scale, K[2] * scale, K[3] * scale)
    base = make_scene(scene_seed, height, width, Ks[0], hole_frac=hole_frac)
    xi_rel, R_abs, t_abs = make_sequence_motions(n_frames, motion_seed, **motion_kw)
    rng = np.random.default_rng(motion_seed + 7)
    hb, wb = (height + 7) // 8, (width + 7) // 8
    if device is None:
        xp = np
        bgr = np.empty((n_frames, height, width, 3), np.uint8)
        depth = np.empty((n_frames, height, width), np.uint16)
    else:
        import torch
        xp = torch
        bgr = torch.empty((n_frames, height, width, 3), dtype=torch.uint8, device=device)
        depth = torch.empty((n_frames, height, width), dtype=torch.uint16, device=device)
    for c0 in range(0, n_frames, chunk):
        n = min(chunk, n_frames - c0)
        rep = lambda a: np.stack([np.asarray(a)] * n)  # noqa: E731
        arrays = (rep(base.normal), np.full(n, base.offset), rep(base.e1), rep(base.e2), rep(base.freq),
                  rep(base.phase), rep(base.amp), R_abs[c0:c0 + n], t_abs[c0:c0 + n])
        _, _, g2, d2, _ = _render(xp, arrays, height, width, Ks, depth_scale, device=device)
        holes = np.stack([np.kron(rng.random((hb, wb)) < hole_frac, np.ones((8, 8), dtype=bool))[:height, :width]
                          for _ in range(n)])
        if device is None:
            d = d2.astype(np.uint16)
            d[holes] = 0
            depth[c0:c0 + n] = d
            bgr[c0:c0 + n] = g2.astype(np.uint8)[..., None]
        else:
            d = d2.to(xp.int32).masked_fill(xp.as_tensor(holes, device=device), 0)
            depth[c0:c0 + n] = d.to(xp.uint16)
            bgr[c0:c0 + n] = g2.to(xp.uint8)[..., None]
    return dict(bgr=bgr, depth=depth, xi=xi_rel.astype(np.float64), K=Ks, depth_scale=depth_scale)
