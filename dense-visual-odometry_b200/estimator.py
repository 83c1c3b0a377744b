"""Host-side mirror of the reference's estimator interface over the C ABI (include/dvo_b200.h).

`RobustDVOB200` keeps the reference's surface — constructor kwargs of `BaseRobustDVO`
(core/robust_dense_visual_odometry/base_robust_dvo.py:34-83), `step()` / `current_pose`
(core/base_dense_visual_odometry.py:54-91) and the four backend hooks
(base_robust_dvo.py:91-135) — while all arithmetic of the path runs in libdvo_b200.so on the GPU.
`PairBatchAligner` is the throughput form: B independent frame pairs per call.

PyTorch is used for device/pinned memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import Optional, Tuple

import numpy as np

from . import _cabi
from .lie import Se3, pose_to_qt

logger = logging.getLogger(__name__)

_WEIGHTS = {"none": _cabi.W_NONE, "tdist": _cabi.W_TDIST_REF, "tdist_mean": _cabi.W_TDIST_REF, "huber": _cabi.W_HUBER,
            "huber_mad": _cabi.W_HUBER_MAD}
_OOB = {"inclusive": _cabi.OOB_INCLUSIVE, "strict": _cabi.OOB_STRICT}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _cabi.DvoError("a CUDA device is required: this path has no CPU fallback")
    return torch


def _stream_ptr(torch, device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _intrinsics_of(camera_model) -> Tuple[float, float, float, float, float]:
    K = np.asarray(camera_model.intrinsics, dtype=np.float32)
    return float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), float(camera_model.depth_scale)


class _Handle:
    """Owns one dvo_handle."""

    def __init__(self, device: int, height: int, width: int, levels: int, max_frames: int, max_pairs: int, cfg):
        self.lib = _cabi.load()
        self.ptr = C.c_void_p()
        rc = self.lib.dvo_create(C.byref(self.ptr), device, height, width, levels, max_frames, max_pairs,
                                 C.byref(cfg))
        if rc != 0:
            msg = self.lib.dvo_last_error(self.ptr)
            if self.ptr:
                self.lib.dvo_destroy(self.ptr)
            self.ptr = None   # __del__ must not destroy it again
            raise _cabi.DvoError(f"dvo_create failed with status {rc}: {msg.decode() if msg else ''}")
        self.height, self.width, self.levels = height, width, levels
        self.max_frames, self.max_pairs = max_frames, max_pairs

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.ptr, *args)
        _cabi.check(self.lib, self.ptr, rc, name)

    def level_shape(self, level):
        h, w = C.c_int(), C.c_int()
        self.call("dvo_level_shape", level, C.byref(h), C.byref(w))
        return h.value, w.value

    def clamp_threshold(self) -> int:
        t = C.c_int()
        self.call("dvo_depth_clamp_threshold", C.byref(t))
        return t.value

    def launch_count(self) -> int:
        return int(self.lib.dvo_launch_count(self.ptr))

    def last_estimate_ms(self) -> float:
        ms = C.c_float()
        self.call("dvo_last_estimate_ms", C.byref(ms))
        return ms.value

    def bounds_violations(self) -> int:
        """Debug builds (-DDVO_BOUNDS_CHECK): out-of-allocation addresses the alignment kernel has formed so far."""
        n = C.c_ulonglong()
        self.call("dvo_debug_bounds_violations", C.byref(n))
        return int(n.value)

    def close(self):
        if getattr(self, "ptr", None):
            self.lib.dvo_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_config(use_weighter=False, max_increased_steps_allowed=0, sigma=None, tolerance=1e-6, max_iterations=100,
                weights: Optional[str] = None, oob_mode="inclusive", huber_k=None, max_distance=5.0,
                threads_per_block=0, blocks_per_sm=0, prefetch_rows=0, approximate_image2_gradient=False,
                cluster_size=0, use_depth_residual=False, depth_weight=None):
    lib = _cabi.load()
    cfg = _cabi.dvo_config()
    lib.dvo_default_config(C.byref(cfg))
    if weights is None:
        weights = "tdist" if use_weighter else "none"
    if weights not in _WEIGHTS:
        raise ValueError(f"weights must be one of {list(_WEIGHTS)}, got '{weights}'")
    if oob_mode not in _OOB:
        raise ValueError(f"oob_mode must be one of {list(_OOB)}, got '{oob_mode}'")
    cfg.weights = _WEIGHTS[weights]
    cfg.tdist_mean = 1 if weights == "tdist_mean" else 0   # extension: textbook scale (mean), SURVEY F3
    cfg.oob_mode = _OOB[oob_mode]
    cfg.max_iterations = int(max_iterations)
    cfg.max_increased_steps = int(max_increased_steps_allowed)
    cfg.tolerance = float(tolerance)
    cfg.sigma_prior = float(sigma) if sigma is not None else -1.0
    if huber_k is not None:
        cfg.huber_k = float(huber_k)
    elif weights == "huber_mad":
        cfg.huber_k = 1.345   # the tuning constant c of k = c * 1.4826 * MAD
    cfg.max_distance = float(max_distance)
    cfg.threads_per_block = int(threads_per_block)
    cfg.blocks_per_sm = int(blocks_per_sm)
    cfg.approximate_image2_gradient = 1 if approximate_image2_gradient else 0
    cfg.cluster_size = int(cluster_size)
    # extension (parity unpinned): photometric + depth residual, lambda_Z = depth_weight
    cfg.use_depth_residual = 1 if use_depth_residual else 0
    if depth_weight is not None:
        cfg.depth_weight = float(depth_weight)
    if int(prefetch_rows) > 32:   # the planes carry slack rows for at most this read-ahead (dvo_create rejects more)
        raise ValueError(f"prefetch_rows must be <= 32, got {prefetch_rows}")
    cfg.reserved[0] = int(prefetch_rows)  # tuning knob: L1 prefetch distance in rows (0 = default, < 0 = off)
    return cfg


def stats_to_numpy(buf: np.ndarray):
    """uint8 [B,128] -> dict of arrays (dvo_pair_stats)."""
    raw = np.ascontiguousarray(buf).view(np.int32).reshape(-1, 32)
    L = _cabi.DVO_MAX_LEVELS
    return dict(iters=raw[:, :L].copy(), n_valid=raw[:, L:2 * L].copy(),
                err=raw[:, 2 * L:3 * L].copy().view(np.float32), flags=raw[:, 3 * L].copy())


class RobustDVOB200:
    """Drop-in for `RobustDVOCPU` / `RobustDVOGPU` on a B200.

    Constructor arguments follow base_robust_dvo.py:34-83 (plus `height`/`width` as in
    gpu_robust_dense_visual_odometry.py:17; if omitted the device state is created on the first frame).
    Extras, all defaulting to reference behaviour: `weights` ("none" | "tdist" | "huber"), `oob_mode`
    ("inclusive" | "strict"), `huber_k`, `max_distance`, `device`, `cluster_size` (CTAs sharing the
    pair: 1, 2, 4, 8 or 16; default -1 = the largest the device can co-schedule, 16 on a B200), `use_depth_residual` / `depth_weight` (photometric + depth residual, an extension the
    reference does not have; with weights "none" or "huber").
    """

    def __init__(self, camera_model, initial_pose, levels: int, use_weighter: bool = False,
                 max_increased_steps_allowed: int = 0, sigma: float = None, tolerance: float = 1e-6,
                 max_iterations: int = 100, approximate_image2_gradient: bool = False, height: int = None,
                 width: int = None, weights: Optional[str] = None, oob_mode: str = "inclusive",
                 huber_k: float = None, max_distance: float = 5.0, device: int = 0, cluster_size: int = -1,
                 use_depth_residual: bool = False, depth_weight: float = None):
        if levels < 1 or levels > _cabi.DVO_MAX_LEVELS:
            raise ValueError(f"levels must be in [1, {_cabi.DVO_MAX_LEVELS}], got {levels}")
        self._camera_model = camera_model
        self._initial_pose = initial_pose
        self._current_pose = self._as_local(initial_pose)
        self._last_pose = None
        self._last_estimated_transform = None
        self._levels = int(levels)
        self._sigma = sigma
        self._max_distance = max_distance
        self._device = device
        # one pair at a time: a thread-block cluster of `cluster_size` CTAs shares the pair (kernel 6.7 ms -> 1.55 ms per
        # 640x480 pose at 8, 1.27 ms at 16); the Huber/MAD weights fall back to one 256-thread CTA
        self._cfg = make_config(use_weighter, max_increased_steps_allowed, sigma, tolerance, max_iterations, weights,
                                oob_mode, huber_k, max_distance, threads_per_block=256,
                                approximate_image2_gradient=approximate_image2_gradient, cluster_size=cluster_size,
                                use_depth_residual=use_depth_residual, depth_weight=depth_weight)
        self._approx = bool(approximate_image2_gradient)
        self._h: Optional[_Handle] = None
        self._have_prev = False
        self._prev_slot = 0           # step(): slot holding the previous frame
        self._hook_slots = (2, 3)     # _build_pyramids hook: (prev, cur)
        self._host_prev = None        # lazily fetched (gray, depth) of the previous frame
        self._upload_done = None      # event: the last frame's copies out of the pinned staging buffers have finished
        self.last_stats = None
        if height is not None and width is not None:
            self._ensure(int(height), int(width))

    # ------------------------------------------------------------------ plumbing
    @staticmethod
    def _as_local(pose) -> Se3:
        return Se3.from_qt(pose_to_qt(pose))

    def _ensure(self, height: int, width: int):
        if self._h is not None:
            if (self._h.height, self._h.width) != (height, width):
                raise ValueError(f"frame size changed from {(self._h.height, self._h.width)} to {(height, width)}")
            return
        torch = _torch()
        self._torch = torch
        self._h = _Handle(self._device, height, width, self._levels, 4, 1, self._cfg)
        fx, fy, cx, cy, scale = _intrinsics_of(self._camera_model)
        self._h.call("dvo_set_intrinsics", fx, fy, cx, cy, scale)
        self._clamp_thr = self._h.clamp_threshold()
        dev = torch.device("cuda", self._device)
        self._dev = dev
        self._pin_bgr = torch.empty((height, width, 3), dtype=torch.uint8).pin_memory()
        self._pin_depth = torch.empty((height, width), dtype=torch.uint16).pin_memory()
        self._pin_qt = torch.empty((3, 7), dtype=torch.float32).pin_memory()   # init, last, out
        self._pin_stats = torch.empty((_cabi.STATS_BYTES,), dtype=torch.uint8).pin_memory()

    @property
    def levels(self) -> int:
        return self._levels

    @property
    def current_pose(self):
        return self._current_pose

    # ------------------------------------------------------------------ reference API
    def step(self, color_image: np.ndarray, depth_image: np.ndarray, init_guess=None, **kwargs):
        """base_dense_visual_odometry.py:54-87.  Returns the Se3 taking the previous camera frame to the
        current one (identity for the first frame), or None if the estimate is not finite."""
        if color_image.ndim != 3 or color_image.shape[2] != 3 or color_image.dtype != np.uint8:
            raise ValueError("color_image must be HxWx3 uint8 (BGR)")
        if depth_image.shape != color_image.shape[:2]:
            raise ValueError("depth_image must be HxW")
        h, w = depth_image.shape
        self._ensure(h, w)
        torch = self._torch
        if depth_image.dtype != np.uint16:
            # the reference accepts any integer dtype here (its own unit test passes uint8)
            depth16 = depth_image.astype(np.uint16)
        else:
            depth16 = depth_image
        st = _stream_ptr(torch, self._dev)
        cur_slot = 1 - self._prev_slot
        # the previous frame's asynchronous copies out of the pinned staging buffers must be over before they are
        # rewritten (the first-frame path returns without synchronising the stream)
        if self._upload_done is not None:
            self._upload_done.synchronize()
        self._pin_bgr.numpy()[...] = color_image
        self._pin_depth.numpy()[...] = depth16
        self._h.call("dvo_build_pyramids_host", cur_slot, C.c_void_p(self._pin_bgr.data_ptr()),
                     C.c_void_p(self._pin_depth.data_ptr()), 1, 1, st)
        self._upload_done = torch.cuda.Event()
        self._upload_done.record(torch.cuda.current_stream(self._dev))
        # the reference zeroes far depth in the caller's array (base_dense_visual_odometry.py:59)
        if self._clamp_thr < 65536:
            depth_image[depth_image >= self._clamp_thr] = 0

        if not self._have_prev:
            transform = Se3.identity()
        else:
            qt = self._pin_qt.numpy()
            init_ptr = None
            if init_guess is not None:
                qt[0] = pose_to_qt(init_guess)
                init_ptr = C.c_void_p(self._pin_qt[0].data_ptr())
            last_ptr = None
            if self._sigma is not None and self._last_estimated_transform is not None:
                qt[1] = pose_to_qt(self._last_estimated_transform)
                last_ptr = C.c_void_p(self._pin_qt[1].data_ptr())
            self._h.call("dvo_estimate_host", self._prev_slot, cur_slot, 1, init_ptr, last_ptr,
                         C.c_void_p(self._pin_qt[2].data_ptr()), C.c_void_p(self._pin_stats.data_ptr()), st)
            torch.cuda.current_stream(self._dev).synchronize()
            out = qt[2].copy()
            self.last_stats = stats_to_numpy(self._pin_stats.numpy()[None, :])
            transform = Se3.from_qt(out) if np.all(np.isfinite(out)) else None

        if transform is not None:
            self._last_pose = self._current_pose.copy()
            self._last_estimated_transform = transform.copy()
            self._current_pose = self._current_pose * transform.inverse()
            self._prev_slot = cur_slot
            self._have_prev = True
            self._host_prev = None
        else:
            logger.warning("DVO could not estimate transform, trying luck on next frame..")
        return transform

    # ------------------------------------------------------------------ backend hooks
    def _fetch_prev(self):
        """Host copies of the previous frame's gray / clamped depth (level 0 of its slot)."""
        if self._host_prev is None:
            if not self._have_prev:
                return None, None
            self._host_prev = self.get_pyramid_level(self._prev_slot, 0)[:2]
        return self._host_prev

    @property
    def _gray_image_prev(self):
        return self._fetch_prev()[0]

    @property
    def _depth_image_prev(self):
        return self._fetch_prev()[1]

    def _build_pyramids(self, gray_image: np.ndarray, depth_image: np.ndarray):
        """base_robust_dvo.py:119-125 / cpu_...py:44-52: previous-frame pyramids from the stored previous
        frame, current-frame pyramids (and Sobel planes) from the arguments."""
        if not self._have_prev:
            raise NotImplementedError("no previous frame: call step() with a first frame before _build_pyramids")
        torch = self._torch
        st = _stream_ptr(torch, self._dev)
        g = torch.as_tensor(np.ascontiguousarray(gray_image, dtype=np.uint8)).to(self._dev)
        d = torch.as_tensor(np.ascontiguousarray(depth_image).astype(np.uint16)).to(self._dev)
        ps, cs = self._hook_slots
        pg, pd = self._fetch_prev()
        pgt = torch.as_tensor(pg).to(self._dev)
        pdt = torch.as_tensor(pd).to(self._dev)
        self._h.call("dvo_build_pyramids_gray", ps, C.c_void_p(pgt.data_ptr()), C.c_void_p(pdt.data_ptr()), 1,
                     1 if self._approx else 0, st)
        self._h.call("dvo_build_pyramids_gray", cs, C.c_void_p(g.data_ptr()), C.c_void_p(d.data_ptr()), 1, 2, st)
        torch.cuda.current_stream(self._dev).synchronize()
        self._hook_ready = True

    def _setup(self, level: int):
        """Sobel planes of every level were built with the pyramid; nothing to do per level."""
        if not 0 <= level < self._levels:
            raise IndexError(f"'level' out of range [0, {self._levels - 1}], got {level} instead")

    def _cleanup(self):
        pass

    def compute_residuals_and_jacobian(self, estimate, level: int = 0):
        """base_robust_dvo.py:91-117: (residuals Nx1 f32, jacobian Nx6 f32, depth mask HxW bool) in masked
        row-major pixel order, from the pyramids of the last `_build_pyramids` call."""
        if not getattr(self, "_hook_ready", False):
            raise NotImplementedError("Call to _build_pyramids did not correctly set pyramids")
        r, J, mask, valid, _ = self.residuals_dense(estimate, level, self._hook_slots[0], self._hook_slots[1])
        v = valid.reshape(-1)
        return r.reshape(-1, 1)[v], J.reshape(-1, 6)[v], mask

    def residuals_dense(self, estimate, level: int, prev_slot: int, cur_slot: int):
        """Dense dump of one level: r [H,W], J [H,W,6], depth mask, warp-valid mask, acc[29] (float64)."""
        torch = self._torch
        hl, wl = self._h.level_shape(level)
        st = _stream_ptr(torch, self._dev)
        r = torch.empty((hl, wl), dtype=torch.float32, device=self._dev)
        J = torch.empty((hl, wl, 6), dtype=torch.float32, device=self._dev)
        m = torch.empty((hl, wl), dtype=torch.uint8, device=self._dev)
        v = torch.empty((hl, wl), dtype=torch.uint8, device=self._dev)
        acc = torch.empty((_cabi.DVO_ACC_TERMS,), dtype=torch.float64, device=self._dev)
        qt = np.ascontiguousarray(pose_to_qt(estimate))
        self._h.call("dvo_residuals_jacobian", prev_slot, cur_slot, level, qt.ctypes.data_as(C.c_void_p),
                     C.c_void_p(r.data_ptr()), C.c_void_p(J.data_ptr()), C.c_void_p(m.data_ptr()),
                     C.c_void_p(v.data_ptr()), C.c_void_p(acc.data_ptr()), st)
        torch.cuda.current_stream(self._dev).synchronize()
        return (r.cpu().numpy(), J.cpu().numpy(), m.cpu().numpy().astype(bool), v.cpu().numpy().astype(bool),
                acc.cpu().numpy())

    def depth_residuals_dense(self, estimate, level: int, prev_slot: int, cur_slot: int):
        """Dense dump of the depth-residual extension at one level: r_Z [H,W] (NaN = undefined), J_Z [H,W,6],
        valid_Z mask, acc[29] (float64: the term's share of the normal equations, depth_weight included)."""
        torch = self._torch
        hl, wl = self._h.level_shape(level)
        st = _stream_ptr(torch, self._dev)
        r = torch.empty((hl, wl), dtype=torch.float32, device=self._dev)
        J = torch.empty((hl, wl, 6), dtype=torch.float32, device=self._dev)
        v = torch.empty((hl, wl), dtype=torch.uint8, device=self._dev)
        acc = torch.empty((_cabi.DVO_ACC_TERMS,), dtype=torch.float64, device=self._dev)
        qt = np.ascontiguousarray(pose_to_qt(estimate))
        self._h.call("dvo_depth_residuals_jacobian", prev_slot, cur_slot, level, qt.ctypes.data_as(C.c_void_p),
                     C.c_void_p(r.data_ptr()), C.c_void_p(J.data_ptr()), C.c_void_p(v.data_ptr()),
                     C.c_void_p(acc.data_ptr()), st)
        torch.cuda.current_stream(self._dev).synchronize()
        return r.cpu().numpy(), J.cpu().numpy(), v.cpu().numpy().astype(bool), acc.cpu().numpy()

    def get_pyramid_level(self, slot: int, level: int):
        """(gray u8, depth u16, gx f32, gy f32) of one level of a frame slot, as host arrays."""
        torch = self._torch
        hl, wl = self._h.level_shape(level)
        st = _stream_ptr(torch, self._dev)
        g = torch.empty((hl, wl), dtype=torch.uint8, device=self._dev)
        d = torch.empty((hl, wl), dtype=torch.uint16, device=self._dev)
        gx = torch.empty((hl, wl), dtype=torch.float32, device=self._dev)
        gy = torch.empty((hl, wl), dtype=torch.float32, device=self._dev)
        self._h.call("dvo_get_pyramid", slot, level, C.c_void_p(g.data_ptr()), C.c_void_p(d.data_ptr()),
                     C.c_void_p(gx.data_ptr()), C.c_void_p(gy.data_ptr()), st)
        torch.cuda.current_stream(self._dev).synchronize()
        return g.cpu().numpy(), d.cpu().numpy(), gx.cpu().numpy(), gy.cpu().numpy()


    def get_point_list(self, slot: int, level: int):
        """(z f32, col i32, row i32, intensity u8) of the frame slot's point list at one level, in list order: the
        pixels with depth, as the alignment kernel walks them when the frame is the previous frame of a pair."""
        torch = self._torch
        hl, wl = self._h.level_shape(level)
        st = _stream_ptr(torch, self._dev)
        z = torch.empty(hl * wl, dtype=torch.float32, device=self._dev)
        col = torch.empty(hl * wl, dtype=torch.int32, device=self._dev)
        row = torch.empty(hl * wl, dtype=torch.int32, device=self._dev)
        inten = torch.empty(hl * wl, dtype=torch.uint8, device=self._dev)
        n = torch.zeros(2, dtype=torch.int32, device=self._dev)
        self._h.call("dvo_get_point_list", slot, level, C.c_void_p(z.data_ptr()), C.c_void_p(col.data_ptr()),
                     C.c_void_p(row.data_ptr()), C.c_void_p(inten.data_ptr()), C.c_void_p(n.data_ptr()), st)
        torch.cuda.current_stream(self._dev).synchronize()
        k = int(n[0].item())
        return z[:k].cpu().numpy(), col[:k].cpu().numpy(), row[:k].cpu().numpy(), inten[:k].cpu().numpy()


class PairBatchAligner:
    """B independent frame pairs per call (BASELINE.json configs 2-4): one persistent kernel launch runs
    every pair's coarse-to-fine Gauss-Newton on the device.

    Inputs may be CUDA tensors (resident path) or host arrays / pinned tensors (end-to-end path: the
    copies are part of the call).  Frame slots: previous frames 0..B-1, current frames B..2B-1.
    """

    def __init__(self, camera_model, height: int, width: int, levels: int, max_pairs: int, device: int = 0,
                 **cfg_kwargs):
        torch = _torch()
        self._torch = torch
        self._dev = torch.device("cuda", device)
        self.max_pairs = int(max_pairs)
        self.levels = int(levels)
        self._cfg = make_config(**cfg_kwargs)
        # roles of the frames (dvo_build_pyramids, with_gradients): previous frames 0 (their own tap records are read
        # only with approximate_image2_gradient: 1), current frames 2 (tap records only)
        self._prev_grad = 1 if cfg_kwargs.get("approximate_image2_gradient") else 0
        self._h = _Handle(device, height, width, levels, 2 * self.max_pairs, self.max_pairs, self._cfg)
        fx, fy, cx, cy, scale = _intrinsics_of(camera_model)
        self._h.call("dvo_set_intrinsics", fx, fy, cx, cy, scale)
        self.clamp_threshold = self._h.clamp_threshold()
        self._qt = torch.empty((self.max_pairs, 7), dtype=torch.float32, device=self._dev)
        self._stats = torch.empty((self.max_pairs, _cabi.STATS_BYTES), dtype=torch.uint8, device=self._dev)
        self._pin_qt = torch.empty((self.max_pairs, 7), dtype=torch.float32).pin_memory()
        self._pin_stats = torch.empty((self.max_pairs, _cabi.STATS_BYTES), dtype=torch.uint8).pin_memory()

    @property
    def handle(self) -> _Handle:
        return self._h

    def _ptr(self, t):
        return C.c_void_p(t.data_ptr())

    def build(self, bgr_prev, depth_prev, bgr_cur, depth_cur):
        """Gray conversion, depth clamp (in place on device inputs), pyramids and gradient planes."""
        torch = self._torch
        B = bgr_prev.shape[0]
        if B > self.max_pairs:
            raise ValueError(f"batch {B} exceeds max_pairs {self.max_pairs}")
        st = _stream_ptr(torch, self._dev)
        if isinstance(bgr_prev, np.ndarray) or not bgr_prev.is_cuda:
            bp, dp, bc, dc = self._as_host_tensors(bgr_prev, depth_prev, bgr_cur, depth_cur)
            self._h.call("dvo_build_pyramids_host", 0, self._ptr(bp), self._ptr(dp), B, self._prev_grad, st)
            self._h.call("dvo_build_pyramids_host", self.max_pairs, self._ptr(bc), self._ptr(dc), B, 2, st)
        else:
            self._h.call("dvo_build_pyramids", 0, self._ptr(bgr_prev), self._ptr(depth_prev), B, self._prev_grad, st)
            self._h.call("dvo_build_pyramids", self.max_pairs, self._ptr(bgr_cur), self._ptr(depth_cur), B, 2, st)
        self._B = B

    def _as_host_tensors(self, *arrays):
        torch = self._torch
        out = tuple((torch.as_tensor(a) if isinstance(a, np.ndarray) else a).contiguous() for a in arrays)
        self._keep = out  # the copies are asynchronous: keep the sources alive
        return out

    def estimate(self, init_qt=None, to_host: bool = True):
        """Runs the GN kernel on the pairs of the last build().  Returns (qt [B,7], stats dict) as host arrays
        when to_host, else the device tensors (no synchronisation)."""
        torch = self._torch
        B = self._B
        st = _stream_ptr(torch, self._dev)
        init_ptr = None
        if init_qt is not None:
            self._init = torch.as_tensor(np.asarray(init_qt, dtype=np.float32)).to(self._dev).contiguous()
            init_ptr = self._ptr(self._init)
        self._h.call("dvo_estimate", 0, self.max_pairs, B, init_ptr, None, self._ptr(self._qt), self._ptr(self._stats),
                     st)
        if not to_host:
            return self._qt[:B], self._stats[:B]
        self._pin_qt[:B].copy_(self._qt[:B], non_blocking=True)
        self._pin_stats[:B].copy_(self._stats[:B], non_blocking=True)
        torch.cuda.current_stream(self._dev).synchronize()
        return self._pin_qt[:B].numpy().copy(), stats_to_numpy(self._pin_stats[:B].numpy())

    def align(self, bgr_prev, depth_prev, bgr_cur, depth_cur, init_qt=None, chunk_pairs: int = 256):
        """build() + estimate().  Host inputs are pipelined: the batch is cut into chunks of `chunk_pairs` pairs;
        one stream uploads chunk after chunk, three compute streams take the chunks round-robin, each waiting for
        its upload and then building the pyramids and running the Gauss-Newton kernel while later chunks are still
        on the bus (kernels of different chunks share the GPU).  The result does
        not depend on the chunking: every pair is estimated by one CTA with a fixed order of operations."""
        host = isinstance(bgr_prev, np.ndarray) or not bgr_prev.is_cuda
        B = bgr_prev.shape[0]
        if not host or B <= chunk_pairs:
            self.build(bgr_prev, depth_prev, bgr_cur, depth_cur)
            return self.estimate(init_qt)
        torch = self._torch
        if B > self.max_pairs:
            raise ValueError(f"batch {B} exceeds max_pairs {self.max_pairs}")
        bp, dp, bc, dc = self._as_host_tensors(bgr_prev, depth_prev, bgr_cur, depth_cur)
        init_dev = None
        if init_qt is not None:
            init_dev = torch.as_tensor(np.asarray(init_qt, dtype=np.float32)).to(self._dev).contiguous()
            self._init = init_dev
        if not hasattr(self, "_streams"):
            self._streams = [torch.cuda.Stream(self._dev) for _ in range(3)]
            self._copy_stream = torch.cuda.Stream(self._dev)
        cur = torch.cuda.current_stream(self._dev)
        cs = self._copy_stream
        cs.wait_stream(cur)
        for s in self._streams:
            s.wait_stream(cur)
        cp = C.c_void_p(cs.cuda_stream)
        # all uploads go back to back on one stream (the copy engine never idles); chunk k's compute stream
        # waits for its upload's event, then builds and estimates while later chunks are still arriving.  Staging
        # slots are per frame, so chunks of one call never share them; the previous call was synchronised.
        for k, lo in enumerate(range(0, B, chunk_pairs)):
            n = min(chunk_pairs, B - lo)
            s = self._streams[k % 3]
            sp = C.c_void_p(s.cuda_stream)
            self._h.call("dvo_upload_frames", lo, self._ptr(bp[lo]), self._ptr(dp[lo]), n, cp)
            self._h.call("dvo_upload_frames", self.max_pairs + lo, self._ptr(bc[lo]), self._ptr(dc[lo]), n, cp)
            ev = torch.cuda.Event()
            ev.record(cs)
            s.wait_event(ev)
            self._h.call("dvo_build_pyramids_staged", lo, n, self._prev_grad, sp)
            self._h.call("dvo_build_pyramids_staged", self.max_pairs + lo, n, 2, sp)
            init_ptr = self._ptr(init_dev[lo]) if init_dev is not None else None
            self._h.call("dvo_estimate", lo, self.max_pairs + lo, n, init_ptr, None, self._ptr(self._qt[lo]),
                         self._ptr(self._stats[lo]), sp)
            with torch.cuda.stream(s):
                self._pin_qt[lo:lo + n].copy_(self._qt[lo:lo + n], non_blocking=True)
                self._pin_stats[lo:lo + n].copy_(self._stats[lo:lo + n], non_blocking=True)
        for s in self._streams:
            s.synchronize()
            cur.wait_stream(s)
        cur.wait_stream(cs)
        self._B = B
        return self._pin_qt[:B].numpy().copy(), stats_to_numpy(self._pin_stats[:B].numpy())

    def last_kernel_ms(self) -> float:
        return self._h.last_estimate_ms()

    def launch_count(self) -> int:
        return self._h.launch_count()


def sequence_launch_groups(n_frames: int, chunk_frames: int):
    """How SequenceAligner.align cuts a stream of host frames: a list of groups, each a list of (lo, hi) upload chunks
    of at most chunk_frames frames; one estimate is launched per group.  Single chunks first (the GPU starts while
    most frames are still on the bus), and everything within two chunks of the end as ONE group (a launch with more
    pairs than CTAs is balanced over the SMs by the kernel's work queue and finished by its tail kernel)."""
    if chunk_frames < 1:
        raise ValueError("chunk_frames must be positive")
    chunks = [(lo, min(lo + chunk_frames, n_frames)) for lo in range(0, n_frames, chunk_frames)]
    groups, k = [], 0
    while k < len(chunks):
        if n_frames - chunks[k][0] <= 2 * chunk_frames:
            groups.append(chunks[k:])
            break
        groups.append(chunks[k:k + 1])
        k += 1
    return groups


class SequenceAligner:
    """A stream of N frames -> N-1 relative poses in one go (BASELINE.json configs[2]).

    Frame f lives in frame slot f and its pyramids and gradient planes are built ONCE; pair p aligns slot p
    (previous frame) against slot p+1 (current frame), i.e. `dvo_estimate(prev_base=0, cur_base=1)`.  Every pair
    starts from the identity guess exactly as `step()` does (base_dense_visual_odometry.py:56,67), so the pairs
    are independent and run as one batch; the `sigma` motion prior would make them serially dependent and is
    rejected here (use RobustDVOB200.step for that).  Host inputs are pipelined over three streams like
    PairBatchAligner.align."""

    def __init__(self, camera_model, height: int, width: int, levels: int, max_frames: int, device: int = 0,
                 **cfg_kwargs):
        if cfg_kwargs.get("sigma") is not None:
            raise ValueError("the sigma prior couples consecutive pairs; use RobustDVOB200.step()")
        torch = _torch()
        self._torch = torch
        self._dev = torch.device("cuda", device)
        self.max_frames = int(max_frames)
        self.levels = int(levels)
        if self.max_frames < 2:
            raise ValueError("a sequence needs at least two frames")
        self._cfg = make_config(**cfg_kwargs)
        self._h = _Handle(device, height, width, levels, self.max_frames, self.max_frames - 1, self._cfg)
        fx, fy, cx, cy, scale = _intrinsics_of(camera_model)
        self._h.call("dvo_set_intrinsics", fx, fy, cx, cy, scale)
        n = self.max_frames - 1
        self._qt = torch.empty((n, 7), dtype=torch.float32, device=self._dev)
        self._stats = torch.empty((n, _cabi.STATS_BYTES), dtype=torch.uint8, device=self._dev)
        self._pin_qt = torch.empty((n, 7), dtype=torch.float32).pin_memory()
        self._pin_stats = torch.empty((n, _cabi.STATS_BYTES), dtype=torch.uint8).pin_memory()
        self._streams = [torch.cuda.Stream(self._dev) for _ in range(3)]
        self._copy_stream = torch.cuda.Stream(self._dev)

    @property
    def handle(self) -> _Handle:
        return self._h

    def align(self, bgr, depth, chunk_frames: int = 256):
        """bgr [N,H,W,3] u8, depth [N,H,W] u16 (host arrays / pinned tensors or CUDA tensors).
        Returns (qt [N-1,7], stats dict) as host arrays; qt[p] takes frame p's camera to frame p+1's."""
        torch = self._torch
        N = bgr.shape[0]
        if N < 2 or N > self.max_frames:
            raise ValueError(f"need 2..{self.max_frames} frames, got {N}")
        host = isinstance(bgr, np.ndarray) or not bgr.is_cuda
        if host:
            bgr = (torch.as_tensor(bgr) if isinstance(bgr, np.ndarray) else bgr).contiguous()
            depth = (torch.as_tensor(depth) if isinstance(depth, np.ndarray) else depth).contiguous()
            self._keep = (bgr, depth)
        else:
            # nothing to overlap with: one launch for all pairs, so that the kernel's work queue balances them over
            # the SMs and its tail kernel finishes the last ones (chunks exist to hide the upload of host frames)
            chunk_frames = N
        cur = torch.cuda.current_stream(self._dev)
        cs = self._copy_stream
        cs.wait_stream(cur)
        for s in self._streams:
            s.wait_stream(cur)
        cp = C.c_void_p(cs.cuda_stream)
        # Frames arrive in chunks of chunk_frames (upload on the copy stream, pyramids as soon as a chunk has landed).
        # Estimates are launched per GROUP of chunks: single chunks first, so that the GPU starts while most frames are
        # still on the bus (a launch of <= 296 pairs runs one CTA per pair to completion, and launches of different
        # groups share the SMs), but everything within two chunks of the end goes into ONE launch: with more pairs
        # than CTAs the kernel's work queue and tail kernel balance the finish over all SMs, where short launches would
        # each end on their longest pair.
        groups = sequence_launch_groups(N, chunk_frames)
        built = None  # event: the previous group's pyramids are complete
        for gi, group in enumerate(groups):
            s = self._streams[gi % 3]
            sp = C.c_void_p(s.cuda_stream)
            for lo, hi in group:
                if host:   # uploads back to back on the copy stream, the compute stream waits for its chunk's event
                    self._h.call("dvo_upload_frames", lo, C.c_void_p(bgr[lo].data_ptr()),
                                 C.c_void_p(depth[lo].data_ptr()), hi - lo, cp)
                    up = torch.cuda.Event()
                    up.record(cs)
                    s.wait_event(up)
                    self._h.call("dvo_build_pyramids_staged", lo, hi - lo, 1, sp)
                else:
                    self._h.call("dvo_build_pyramids", lo, C.c_void_p(bgr[lo].data_ptr()),
                                 C.c_void_p(depth[lo].data_ptr()), hi - lo, 1, sp)
            ev = torch.cuda.Event()
            ev.record(s)
            glo, ghi = group[0][0], group[-1][1]
            p0, p1 = max(glo - 1, 0), ghi - 1   # pairs whose current frame is in this group
            if p1 > p0:
                if built is not None and p0 < glo:
                    s.wait_event(built)       # pair glo-1 reads frame glo-1, built on another stream
                self._h.call("dvo_estimate", p0, p0 + 1, p1 - p0, None, None, C.c_void_p(self._qt[p0].data_ptr()),
                             C.c_void_p(self._stats[p0].data_ptr()), sp)
                with torch.cuda.stream(s):
                    self._pin_qt[p0:p1].copy_(self._qt[p0:p1], non_blocking=True)
                    self._pin_stats[p0:p1].copy_(self._stats[p0:p1], non_blocking=True)
            built = ev
        for s in self._streams:
            s.synchronize()
            cur.wait_stream(s)
        cur.wait_stream(cs)
        return self._pin_qt[:N - 1].numpy().copy(), stats_to_numpy(self._pin_stats[:N - 1].numpy())

    def estimate_pair(self, prev_frame: int, cur_frame: int):
        """One more estimate between two frames of the last align() call (their pyramids are still resident):
        qt [7] taking frame prev_frame's camera to frame cur_frame's, or None if it is not finite."""
        torch = self._torch
        if not (0 <= prev_frame < self.max_frames and 0 <= cur_frame < self.max_frames):
            raise ValueError("frame index out of range")
        st = _stream_ptr(torch, self._dev)
        self._h.call("dvo_estimate", prev_frame, cur_frame, 1, None, None, C.c_void_p(self._qt[0].data_ptr()),
                     C.c_void_p(self._stats[0].data_ptr()), st)
        q = self._qt[0].cpu().numpy().copy()
        return q if np.all(np.isfinite(q)) else None

    def launch_count(self) -> int:
        return self._h.launch_count()


def robust_dvo_factory(use_gpu: bool = True, **kwargs):
    """Mirror of core/robust_dense_visual_odometry/__init__.py:5-25 with the B200 backend as the GPU branch.
    There is no CPU branch here: `use_gpu=False` is an error, not a fallback."""
    if not use_gpu:
        raise ValueError("dense_visual_odometry_b200 has no CPU backend; use the reference for use_gpu=False")
    return RobustDVOB200(**kwargs)


_SUPPORTED_METHODS = {"robust-dvo": robust_dvo_factory}


def get_dvo(method: str, camera_model, init_pose, **kwargs):
    """Mirror of core/__init__.py:14-40 (same registry name, same error wrapping)."""
    if method not in _SUPPORTED_METHODS:
        raise ValueError("Not supported method '{}', available options are '{}'".format(
            method, list(_SUPPORTED_METHODS.keys())))
    try:
        return _SUPPORTED_METHODS[method](camera_model=camera_model, initial_pose=init_pose, **kwargs)
    except Exception as e:
        raise ValueError((
            "Could not dynamically load method '{}' with parameters '{}'".format(method, kwargs),
            ", got the following exception: {}".format(e)))
