"""TEST / BENCH INFRASTRUCTURE ONLY: runs the UNMODIFIED reference implementation (pfontana96/dense-visual-odometry)
when an installed copy can be imported, for `bench.py --impl reference` / `cpu_baseline` and for the golden-vector
generator.  Never imported by the product package.

Where the reference is looked for, in order: $DVO_REFERENCE_SRC, baseline/_ref (the offline `pip install --target`
of /root/reference that `__graft_entry__.build()` makes in the build container; git-ignored, travels to the GPU box),
/root/reference/src.

External shims, none of which touches a reference file:
  1. NUMBA_ENABLE_CUDASIM=1 before import: the eager @cuda.jit decorators of cuda/residuals_kernel.py import
     without a GPU driver (and identically with one).
  2. numpy.bool8 = numpy.bool_ (annotation used by the reference, gone in NumPy 2).
  3. RobustDVOCPU.interpolate_bilinear (cpu_robust_dense_visual_odometry.py:202-254) replaced by the same body plus
     the missing out-of-image `continue` (the unmodified function reads out of bounds): inclusive
     (default) or strict.
"""
from __future__ import annotations

import math
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
_CANDIDATES = [os.environ.get("DVO_REFERENCE_SRC"), str(ROOT / "baseline" / "_ref"), "/root/reference/src"]

_state = {}


def find_source():
    for c in _CANDIDATES:
        if c and (Path(c) / "dense_visual_odometry" / "core" / "__init__.py").exists():
            return Path(c)
    return None


def available() -> bool:
    return find_source() is not None


def load(guard: str = "inclusive"):
    """Imports the reference with the shims and returns a namespace of its public pieces."""
    if "ns" in _state:
        set_guard(guard)
        return _state["ns"]
    src = find_source()
    if src is None:
        raise ImportError("the reference is not installed (baseline/_ref, $DVO_REFERENCE_SRC, /root/reference/src)")
    os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
    import numpy as np
    np.bool8 = np.bool_
    sys.path.insert(0, str(src))
    import numba as nb
    from dense_visual_odometry.core import get_dvo
    from dense_visual_odometry.camera_model import RGBDCameraModel
    from dense_visual_odometry.utils.lie_algebra import Se3, So3
    from dense_visual_odometry.core.robust_dense_visual_odometry.cpu_robust_dense_visual_odometry import RobustDVOCPU

    sigs = ['float32[:,:](uint8[:,:], float32[:,:])', 'float32[:,:](float32[:,:], float32[:,:])']

    @nb.njit(sigs, parallel=True, fastmath=True)
    def interp_inclusive(image, pixels_coordinates):
        N = pixels_coordinates.shape[0]
        height, width = image.shape
        out = np.empty((N, 1), dtype=np.float32)
        for i in nb.prange(N):
            x, y = pixels_coordinates[i]
            if not ((x >= 0) and (y >= 0) and (x <= width - 1) and (y <= height - 1)):
                out[i, 0] = np.nan
                continue
            x0 = int(math.floor(x))
            y0 = int(math.floor(y))
            x1 = x0 + 1
            y1 = y0 + 1
            w00 = (x1 - x) * (y1 - y)
            w01 = (x1 - x) * (y - y0)
            w10 = (x - x0) * (y1 - y)
            w11 = (x - x0) * (y - y0)
            x1c = min(x1, width - 1)
            y1c = min(y1, height - 1)
            out[i, 0] = (
                (w00 * image[y0, x0] + w01 * image[y1c, x0] + w10 * image[y0, x1c] + w11 * image[y1c, x1c]) /
                ((x1 - x0) * (y1 - y0))
            )
        return out

    @nb.njit(sigs, parallel=True, fastmath=True)
    def interp_strict(image, pixels_coordinates):
        N = pixels_coordinates.shape[0]
        height, width = image.shape
        out = np.empty((N, 1), dtype=np.float32)
        for i in nb.prange(N):
            x, y = pixels_coordinates[i]
            x0 = int(math.floor(x))
            y0 = int(math.floor(y))
            x1 = x0 + 1
            y1 = y0 + 1
            if (x0 < 0) or (y0 < 0) or (x1 >= width) or (y1 >= height) or not (x == x) or not (y == y):
                out[i, 0] = np.nan
                continue
            w00 = (x1 - x) * (y1 - y)
            w01 = (x1 - x) * (y - y0)
            w10 = (x - x0) * (y1 - y)
            w11 = (x - x0) * (y - y0)
            out[i, 0] = (
                (w00 * image[y0, x0] + w01 * image[y1, x0] + w10 * image[y0, x1] + w11 * image[y1, x1]) /
                ((x1 - x0) * (y1 - y0))
            )
        return out

    class NS:
        pass

    ns = NS()
    ns.source = src
    ns.get_dvo, ns.RGBDCameraModel, ns.Se3, ns.So3, ns.RobustDVOCPU = get_dvo, RGBDCameraModel, Se3, So3, RobustDVOCPU
    ns.interp = {"inclusive": interp_inclusive, "strict": interp_strict}
    ns.threads = nb.get_num_threads()
    _state["ns"] = ns
    set_guard(guard)
    return ns


def set_guard(mode: str):
    ns = _state["ns"]
    ns.RobustDVOCPU.interpolate_bilinear = staticmethod(ns.interp[mode])


def estimate_pair(K4, depth_scale, levels, bgr0, d0, bgr1, d1, use_weighter=False, guard="inclusive"):
    """One pose estimate through the reference's own public API: get_dvo(...).step() twice.
    Returns (q[4], t[3]) of the relative transform."""
    import numpy as np
    ns = load(guard)
    Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], dtype=np.float32)
    cam = ns.RGBDCameraModel(Km, depth_scale)
    dvo = ns.get_dvo("robust-dvo", cam, ns.Se3.identity(), levels=levels, use_weighter=use_weighter)
    dvo.step(bgr0, d0.copy())
    T = dvo.step(bgr1, d1.copy())
    if T is None:
        return None
    return np.concatenate([T.so3.quat.reshape(4), T.tvec.reshape(3)]).astype(np.float32)
