"""CPU oracle for the per-frame photometric alignment hot path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy-only restatement of the reference's CPU path
(pfontana96/dense-visual-odometry, `RobustDVOCPU`), written from the algorithm, not
copied.  It exists to CHECK the CUDA path: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The product
(`dense-visual-odometry_b200/`) never imports anything from `oracle/`.

Parity pinning: `tests/golden/make_golden.py` runs the *real* reference (imported from
/root/reference with the three external shims of SURVEY.md §8c) and commits its outputs
as `tests/golden/*.npz`; `tests/test_oracle_vs_golden.py` checks this restatement against
those files on CPU.  Extensions that the reference does not have (Huber weights, the
depth residual) are marked "parity unpinned" where they are defined.

Every function cites the reference lines it restates (paths relative to
/root/reference/src/dense_visual_odometry/).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

F32 = np.float32
LIE_EPS = 1e-6  # utils/lie_algebra/base_special_group.py:8

OOB_INCLUSIVE = 0  # valid iff 0 <= x <= W-1 and 0 <= y <= H-1 (taps clamped; their weight is 0)
OOB_STRICT = 1     # valid iff x0 >= 0, y0 >= 0, x0+1 < W, y0+1 < H (docstring of cpu_...py:222-223)

W_NONE = 0
W_TDIST_REF = 1    # weighter/t_weighter.py as written (scale is a SUM)
W_HUBER = 2        # extension, parity unpinned
W_HUBER_MAD = 3    # extension, parity unpinned: Huber threshold c * 1.4826 * MAD(r), re-estimated every iteration


# --------------------------------------------------------------------------------------
# a1: colour conversion and depth clamp        core/base_dense_visual_odometry.py:58-59
# --------------------------------------------------------------------------------------
def bgr_to_gray(bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(COLOR_BGR2GRAY) for u8: fixed point, 15 fractional bits, round half up.

    base_dense_visual_odometry.py:58.  OpenCV is a third-party dependency (pinned
    opencv-python~=4.6.0 in requirements.txt:8); its published integer formula is
    (B*3735 + G*19235 + R*9798 + 2^14) >> 15, checked bit-exact against cv2 4.13 in
    tests/golden/make_golden.py.
    """
    b = bgr[..., 0].astype(np.uint32)
    g = bgr[..., 1].astype(np.uint32)
    r = bgr[..., 2].astype(np.uint32)
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


def depth_clamp_threshold(depth_scale: float, max_distance: float = 5.0) -> int:
    """Smallest u16 digital number d with float64(d * depth_scale) > max_distance.

    base_dense_visual_odometry.py:59 evaluates `(depth * scale) > max_distance` in
    float64; the set of zeroed DNs is the upper range [threshold, 65535].  Returns 65536
    if nothing is clamped.
    """
    dn = np.arange(65536, dtype=np.uint16)
    hit = (dn * depth_scale) > max_distance
    idx = np.flatnonzero(hit)
    return int(idx[0]) if idx.size else 65536


def clamp_depth(depth: np.ndarray, depth_scale: float, max_distance: float = 5.0) -> np.ndarray:
    """Out-of-place version of base_dense_visual_odometry.py:59."""
    out = depth.copy()
    out[(out * depth_scale) > max_distance] = 0
    return out


# --------------------------------------------------------------------------------------
# a2: median pyramid                                      utils/image_pyramid.py:19-21
# --------------------------------------------------------------------------------------
def median3(img: np.ndarray) -> np.ndarray:
    """3x3 median with replicated border (cv2.medianBlur(img, 3) for u8 and u16)."""
    p = np.pad(img, 1, mode="edge")
    h, w = img.shape
    stack = np.stack([p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)], axis=0)
    return np.partition(stack, 4, axis=0)[4]


def median3_down(img: np.ndarray) -> np.ndarray:
    """pyrDownMedianSmooth: median then keep even rows/cols (image_pyramid.py:19-21)."""
    return np.ascontiguousarray(median3(img)[::2, ::2])


def build_pyramid(img: np.ndarray, levels: int) -> List[np.ndarray]:
    """ImagePyramid.__init__ (image_pyramid.py:36-54): level 0 is the image itself."""
    pyr = [img]
    for _ in range(1, levels):
        pyr.append(median3_down(pyr[-1]))
    return pyr


# --------------------------------------------------------------------------------------
# a9: Sobel gradients                                          utils/jacobian.py:47-73
# --------------------------------------------------------------------------------------
def sobel3(img: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """cv2.Sobel(ksize=3, CV_32F, BORDER_REFLECT), dx and dy; no normalisation (gain 8).

    BORDER_REFLECT with a one-pixel halo equals edge replication.
    """
    p = np.pad(img.astype(np.int32), 1, mode="edge")
    h, w = img.shape

    def s(dy, dx):
        return p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]

    gx = (s(-1, 1) + 2 * s(0, 1) + s(1, 1)) - (s(-1, -1) + 2 * s(0, -1) + s(1, -1))
    gy = (s(1, -1) + 2 * s(1, 0) + s(1, 1)) - (s(-1, -1) + 2 * s(-1, 0) + s(-1, 1))
    return gx.astype(F32), gy.astype(F32)


# --------------------------------------------------------------------------------------
# a3/a4/a7: camera model                                   camera_model.py:62-79,171-252
# --------------------------------------------------------------------------------------
def intrinsics_at(K: np.ndarray, level: int) -> np.ndarray:
    """RGBDCameraModel.at (camera_model.py:62-79): 3x3 f32 K of pyramid level `level`."""
    K = np.asarray(K, dtype=F32)[:3, :3]
    if level == 0:
        return K.copy()
    e = -level
    S = np.array([[2.0 ** e, 0, 2.0 ** (e - 1) - 0.5],
                  [0, 2.0 ** e, 2.0 ** (e - 1) - 0.5],
                  [0, 0, 1]], dtype=F32)
    return np.dot(S, K).astype(F32)


def deproject(depth: np.ndarray, K_l: np.ndarray, depth_scale: float) -> Tuple[np.ndarray, np.ndarray]:
    """RGBDCameraModel.deproject (camera_model.py:171-226): P (4xN f32) and the H*W mask."""
    h, w = depth.shape
    mask = (depth != 0).reshape(-1)
    z = (depth.reshape(-1) * depth_scale)[mask].astype(F32)
    xs, ys = np.meshgrid(np.arange(w, dtype=F32), np.arange(h, dtype=F32))
    xs = xs.reshape(-1)[mask]
    ys = ys.reshape(-1)[mask]
    Kinv = np.linalg.inv(K_l.astype(F32))
    pts = np.dot(Kinv, np.vstack((xs, ys, np.ones_like(z))))
    P = np.vstack((pts[0] * z, pts[1] * z, z, np.ones_like(z))).astype(F32)
    return P, mask.reshape(h, w)


def project(P: np.ndarray, K_l: np.ndarray) -> np.ndarray:
    """RGBDCameraModel.project (camera_model.py:228-252): (K P) / (K P)_z, 3xN f32."""
    K34 = np.zeros((3, 4), dtype=F32)
    K34[:, :3] = K_l
    uv = np.dot(K34, P)
    uv /= uv[2, :]
    return uv


# --------------------------------------------------------------------------------------
# a5: Jacobian of the warp at the UNtransformed point            utils/jacobian.py:7-44
# --------------------------------------------------------------------------------------
def warp_jacobian(P: np.ndarray, K_l: np.ndarray) -> np.ndarray:
    """compute_jacobian_of_warp_function: N x 2 x 6 f32, twist order [v_x v_y v_z w_x w_y w_z]."""
    fx = F32(K_l[0, 0])
    fy = F32(K_l[1, 1])
    x, y, z = P[0], P[1], P[2]
    n = P.shape[1]
    Jw = np.zeros((n, 2, 6), dtype=F32)
    z2 = z * z
    Jw[:, 0, 0] = fx / z
    Jw[:, 0, 2] = -fx * x / z2
    Jw[:, 0, 3] = -fx * (x * y) / z2
    Jw[:, 0, 4] = (fx * (1 + (x.astype(np.float64) ** 2 / z2.astype(np.float64)))).astype(F32)
    Jw[:, 0, 5] = -fx * y / z
    Jw[:, 1, 1] = fy / z
    Jw[:, 1, 2] = -fy * y / z2
    Jw[:, 1, 3] = (-fy * (1 + (y.astype(np.float64) ** 2 / z2.astype(np.float64)))).astype(F32)
    Jw[:, 1, 4] = fy * (x * y) / z2
    Jw[:, 1, 5] = fy * x / z
    return Jw


# --------------------------------------------------------------------------------------
# a8: bilinear sampling with an explicit out-of-bounds rule          cpu_...py:202-254
# --------------------------------------------------------------------------------------
def interp_bilinear(img: np.ndarray, xy: np.ndarray, oob_mode: int = OOB_INCLUSIVE) -> np.ndarray:
    """RobustDVOCPU.interpolate_bilinear with the missing out-of-image `continue`.

    img: H x W (u8 or f32); xy: N x 2 f32 (x, y).  Returns N f32, NaN where invalid.
    Weights are formed in float64 from the f32 coordinates, as Numba does for
    `int64 - float32` (cpu_...py:243-246); the divisor (x1-x0)(y1-y0) is 1.
    """
    h, w = img.shape
    x = xy[:, 0].astype(F32)
    y = xy[:, 1].astype(F32)
    with np.errstate(invalid="ignore"):
        if oob_mode == OOB_STRICT:
            x0f = np.floor(x)
            y0f = np.floor(y)
            valid = (x0f >= 0) & (y0f >= 0) & (x0f + 1 < w) & (y0f + 1 < h)
        else:
            valid = (x >= 0) & (y >= 0) & (x <= w - 1) & (y <= h - 1)
    out = np.full(x.shape[0], np.nan, dtype=F32)
    if not valid.any():
        return out
    xv = x[valid].astype(np.float64)
    yv = y[valid].astype(np.float64)
    x0 = np.floor(xv).astype(np.int64)
    y0 = np.floor(yv).astype(np.int64)
    x1 = x0 + 1
    y1 = y0 + 1
    w00 = (x1 - xv) * (y1 - yv)
    w01 = (x1 - xv) * (yv - y0)
    w10 = (xv - x0) * (y1 - yv)
    w11 = (xv - x0) * (yv - y0)
    x1c = np.minimum(x1, w - 1)
    y1c = np.minimum(y1, h - 1)
    im = img.astype(np.float64)
    val = w00 * im[y0, x0] + w01 * im[y1c, x0] + w10 * im[y0, x1c] + w11 * im[y1c, x1c]
    out[valid] = val.astype(F32)
    return out


# --------------------------------------------------------------------------------------
# a6/a16: SO(3)/SE(3) as the reference stores them (wxyz quaternion + t, never renormalised)
# utils/lie_algebra/special_orthogonal_group.py, special_euclidean_group.py, common.py
# --------------------------------------------------------------------------------------
def _wrap(a):
    """common.py:30-44."""
    return (a + np.pi) % (2 * np.pi) - np.pi


def quat_to_R(q: np.ndarray) -> np.ndarray:
    """So3.exp (special_orthogonal_group.py:158-188): R from wxyz quaternion, f32 arithmetic."""
    w, x, y, z = [F32(v) for v in np.asarray(q, dtype=F32).reshape(4)]
    two = F32(2)
    one = F32(1)
    return np.array([
        [two * (w * w + x * x) - one, two * (x * y - w * z), two * (x * z + w * y)],
        [two * (x * y + w * z), two * (w * w + y * y) - one, two * (y * z - w * x)],
        [two * (x * z - w * y), two * (y * z + w * x), two * (w * w + z * z) - one]], dtype=F32)


def quat_log(q: np.ndarray) -> np.ndarray:
    """So3.log (special_orthogonal_group.py:190-209): phi (3,) f32 from a wxyz quaternion."""
    q = np.asarray(q, dtype=F32).reshape(4)
    w = q[0]
    vec = q[1:]
    n = F32(np.linalg.norm(vec))
    if n < LIE_EPS:
        return np.zeros(3, dtype=F32)
    theta = F32(_wrap(F32(2 * math.atan2(float(n), float(w))) / n))
    return (theta * vec).astype(F32)


def phi_to_quat(phi: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """So3.__init__ for a (3,1) rotation vector (special_orthogonal_group.py:33-50, 65-86).

    Returns (quat wxyz f32, phi f32 after angle wrapping).
    """
    phi = np.asarray(phi, dtype=F32).reshape(3)
    theta = F32(np.linalg.norm(phi))
    if theta < LIE_EPS:
        return np.array([1, 0, 0, 0], dtype=F32), np.zeros(3, dtype=F32)
    a = phi / theta
    theta = F32(_wrap(theta))
    phi_w = (theta * a).astype(F32)
    th = F32(np.linalg.norm(phi_w))
    ax = phi_w / th
    s = math.sin(float(th) / 2)
    c = math.cos(float(th) / 2)
    q = np.array([c, s * float(ax[0]), s * float(ax[1]), s * float(ax[2])], dtype=F32)
    return q, phi_w


def hat3(phi: np.ndarray) -> np.ndarray:
    m = np.zeros((3, 3), dtype=F32)
    m[0, 1] = -phi[2]
    m[0, 2] = phi[1]
    m[1, 0] = phi[2]
    m[1, 2] = -phi[0]
    m[2, 0] = -phi[1]
    m[2, 1] = phi[0]
    return m


def R_to_quat(R: np.ndarray) -> np.ndarray:
    """So3._SE3_to_quat (special_orthogonal_group.py:88-128), float64 like the reference."""
    R = np.asarray(R, dtype=np.float64)
    q = np.zeros(4)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        t = math.sqrt(1 + t)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1] = (R[2, 1] - R[1, 2]) * t
        q[2] = (R[0, 2] - R[2, 0]) * t
        q[3] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[k, j] - R[j, k]) * t
        q[1 + j] = (R[j, i] + R[i, j]) * t
        q[1 + k] = (R[k, i] + R[i, k]) * t
    return q


@dataclass
class Pose:
    """A reference `Se3`: wxyz quaternion (not renormalised) + translation."""
    q: np.ndarray = field(default_factory=lambda: np.array([1, 0, 0, 0], dtype=F32))
    t: np.ndarray = field(default_factory=lambda: np.zeros(3, dtype=F32))

    def copy(self) -> "Pose":
        return Pose(self.q.copy(), self.t.copy())

    def matrix(self) -> np.ndarray:
        """Se3.exp (special_euclidean_group.py:35-52): 4x4 f32; R = I when |phi| < 1e-6."""
        T = np.eye(4, dtype=F32)
        phi = quat_log(self.q)
        if abs(np.linalg.norm(phi)) >= LIE_EPS:
            T[:3, :3] = quat_to_R(self.q)
        T[:3, 3] = np.asarray(self.t, dtype=F32).reshape(3)
        return T

    def compose(self, right: "Pose") -> "Pose":
        """Se3.__mul__ (special_euclidean_group.py:91-96) with quat_mult (common.py:51-73)."""
        a = np.asarray(self.q).reshape(4)
        b = np.asarray(right.q).reshape(4)
        w = a[0] * b[0] - np.dot(a[1:], b[1:])
        v = a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:])
        q = np.concatenate(([w], v)).astype(F32)
        t = np.asarray(self.t).reshape(3) + np.dot(quat_to_R(self.q), np.asarray(right.t).reshape(3))
        return Pose(q, t)

    def inverse(self) -> "Pose":
        """Se3.inverse (special_euclidean_group.py:79-81): quaternion rebuilt from R^T (f64)."""
        Rt = quat_to_R(self.q).T.copy()
        qi = R_to_quat(Rt)
        t = -np.dot(quat_to_R(qi), np.asarray(self.t).reshape(3))
        return Pose(qi, t)

    def log(self) -> np.ndarray:
        """Se3.log (special_euclidean_group.py:54-77): xi (6,) f32."""
        xi = np.zeros(6, dtype=F32)
        phi = quat_log(self.q)
        theta = np.linalg.norm(phi)
        t = np.asarray(self.t, dtype=np.float64).reshape(3)
        if abs(theta) < LIE_EPS:
            xi[:3] = t
            return xi
        a = (phi / theta).astype(F32)
        # So3(a) wraps/normalises `a` (a unit vector, i.e. a 1 rad rotation vector)
        _, a_phi = phi_to_quat(a)
        a_hat = hat3(a_phi)
        th2 = theta / 2
        A = th2 * np.cos(th2) / np.sin(th2)
        V_inv = A * np.eye(3, dtype=F32) + (1 - A) * np.outer(a_phi, a_phi) - th2 * a_hat
        xi[:3] = np.dot(V_inv, t)
        xi[3:] = phi
        return xi


def pose_from_xi(xi: np.ndarray) -> Pose:
    """Se3.from_se3 (special_euclidean_group.py:105-123)."""
    xi = np.asarray(xi, dtype=F32).reshape(6)
    ups = xi[:3]
    q, phi = phi_to_quat(xi[3:])
    theta = F32(np.linalg.norm(phi))
    if theta < LIE_EPS:
        return Pose(np.array([1, 0, 0, 0], dtype=F32), ups.copy())
    ph = hat3(phi)
    ph2 = np.dot(ph, ph)
    th = float(theta)
    c1 = (1 - math.cos(th)) / (th ** 2)
    c2 = (th - math.sin(th)) / (th ** 3)
    V = (np.eye(3, dtype=F32) + F32(c1) * ph + F32(c2) * ph2).astype(F32)
    return Pose(q, np.dot(V, ups).astype(F32))


# --------------------------------------------------------------------------------------
# a12: reference t-distribution weighter                        weighter/t_weighter.py
# --------------------------------------------------------------------------------------
def tdist_lambda(r2: np.ndarray, dof: float = 5.0, init_sigma: float = 5.0, tol: float = 1e-3,
                 max_iter: int = 50, mean: bool = False) -> float:
    """TDistributionWeighter.weight's fixed point (t_weighter.py:21-34) with the SUMMED scale
    of `_compute_scale` (t_weighter.py:36-47).  Returns the final lambda.
    mean=True is the textbook scale (an extension, parity unpinned): the mean instead of the sum."""
    r2 = r2.astype(np.float64).reshape(-1)
    last = 1.0 / (init_sigma ** 2)
    cur = last
    for _ in range(max_iter):
        sigma2 = float(np.sum(r2 * ((dof + 1) / (dof + r2 * last))))
        if mean and r2.size:
            sigma2 /= r2.size
        cur = 1.0 / sigma2 if sigma2 != 0 else float("inf")
        if abs(cur - last) < tol:
            break
        last = cur
    return cur


def tdist_weights(r2: np.ndarray, dof: float = 5.0, **kw) -> np.ndarray:
    lam = tdist_lambda(r2, dof=dof, **kw)
    return ((dof + 1) / (dof + r2 * F32(lam))).astype(F32)


def huber_weights(r: np.ndarray, k: float) -> np.ndarray:
    """Extension (parity unpinned): w = 1 if |r| <= k else k/|r|, fixed k."""
    a = np.abs(r)
    wgt = np.ones_like(a, dtype=F32)
    big = a > k
    wgt[big] = (F32(k) / a[big]).astype(F32)
    return wgt


# --------------------------------------------------------------------------------------
# a10/a11/a13: residuals, Jacobian, normal equations                  cpu_...py:134-200
# --------------------------------------------------------------------------------------
@dataclass
class LevelData:
    K: np.ndarray
    gray_prev: np.ndarray
    depth_prev: np.ndarray
    gray_cur: np.ndarray
    gx: np.ndarray
    gy: np.ndarray
    P: np.ndarray = None
    mask: np.ndarray = None
    Jw: np.ndarray = None
    i1: np.ndarray = None
    J_approx: np.ndarray = None  # approximate_image2_gradient: Jacobian fixed per level, from I1's gradients
    depth_cur: np.ndarray = None  # depth-residual extension only: the current frame's depth level (u16)
    depth_scale: float = 0.0


def prepare_level(K, depth_scale, gray_prev, depth_prev, gray_cur, level, approximate: bool = False,
                  depth_cur: Optional[np.ndarray] = None) -> LevelData:
    """_setup (cpu_...py:54-77) + the pose-independent part of compute_residuals_and_jacobian.

    approximate=True is the reference's `approximate_image2_gradient` mode (cpu_...py:60-77, :160-165): the image
    Jacobian uses the Sobel gradients of the PREVIOUS image at the unwarped pixel, J = [I1x(x) I1y(x)] J_w, computed
    once per level."""
    K_l = intrinsics_at(K, level)
    gx, gy = sobel3(gray_prev if approximate else gray_cur)
    ld = LevelData(K_l, gray_prev, depth_prev, gray_cur, gx, gy)
    ld.depth_cur = depth_cur
    ld.depth_scale = depth_scale
    ld.P, ld.mask = deproject(depth_prev, K_l, depth_scale)
    ld.Jw = warp_jacobian(ld.P, K_l)
    ld.i1 = gray_prev[ld.mask]
    if approximate:
        g1x = gx[ld.mask].astype(F32)
        g1y = gy[ld.mask].astype(F32)
        ld.J_approx = (g1x[:, None] * ld.Jw[:, 0, :] + g1y[:, None] * ld.Jw[:, 1, :]).astype(F32)
    return ld


def residuals_and_jacobian(ld: LevelData, T: np.ndarray, oob_mode: int = OOB_INCLUSIVE):
    """compute_residuals_and_jacobian (cpu_...py:134-200).

    Returns r (N',) f32, J (N',6) f32, depth mask (H,W) bool, warp-valid (N,) bool where N is the
    number of depth-valid pixels in row-major order and N' the number of those that warp inside I2.
    """
    Pw = np.dot(T.astype(F32), ld.P)
    uv = project(Pw, ld.K)
    xy = np.ascontiguousarray(uv[:2].T)
    i2 = interp_bilinear(ld.gray_cur, xy, oob_mode)
    valid = ~np.isnan(i2)
    if ld.J_approx is not None:   # cpu_...py:187-188
        J = ld.J_approx[valid]
    else:
        xyv = xy[valid]
        gxv = interp_bilinear(ld.gx, xyv, oob_mode)
        gyv = interp_bilinear(ld.gy, xyv, oob_mode)
        Jw = ld.Jw[valid]
        J = (gxv[:, None] * Jw[:, 0, :] + gyv[:, None] * Jw[:, 1, :]).astype(F32)
    r = (i2[valid] - ld.i1[valid]).astype(F32)
    return r, J, ld.mask, valid


def depth_residuals_and_jacobian(ld: LevelData, T: np.ndarray, oob_mode: int = OOB_INCLUSIVE):
    """Depth (geometric) residual: an EXTENSION, the reference has none -- PARITY UNPINNED.  This is the
    definition the CUDA kernel implements (align_kernel.cuh, depth_pair_math), in float64:

      (u', v') = project(T P)  exactly as for the photometric term                      cpu_...py:173-176
      valid_Z  = photometric-valid and u' < W-1 and v' < H-1 (no clamped tap) and the four taps of the current
                 frame's depth level around (u', v') are non-zero
      Z2       = scale * bilinear(D2)(u', v');   r_Z = Z2 - (T P)_z
      grad Z2  = derivative of the bilinear patch: dZ/du = scale ((1-wy)(d10-d00) + wy (d11-d01)), dZ/dv likewise
      J_Z      = [dZ/du dZ/dv] J_w - [0 0 1 Y -X 0], J_w and (X, Y) at the UNtransformed point, the reference's
                 convention for the photometric Jacobian (utils/jacobian.py:37-40)

    Returns r_Z (Nz,) f32, J_Z (Nz,6) f32, valid_Z (N,) bool over the depth-valid pixels in row-major order."""
    Pw = np.dot(T.astype(F32), ld.P)
    uv = project(Pw, ld.K)
    x = uv[0].astype(F32)
    y = uv[1].astype(F32)
    h, w = ld.depth_cur.shape
    i2 = interp_bilinear(ld.gray_cur, np.ascontiguousarray(uv[:2].T), oob_mode)
    with np.errstate(invalid="ignore"):
        inside = ~np.isnan(i2) & (x < w - 1) & (y < h - 1)
    xs = np.where(inside, x, 0).astype(np.float64)
    ys = np.where(inside, y, 0).astype(np.float64)
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.floor(ys).astype(np.int64)
    wx = xs - x0
    wy = ys - y0
    D = ld.depth_cur.astype(np.float64)
    d00, d10, d01, d11 = D[y0, x0], D[y0, x0 + 1], D[y0 + 1, x0], D[y0 + 1, x0 + 1]
    valid = inside & (d00 != 0) & (d10 != 0) & (d01 != 0) & (d11 != 0)
    sc = float(ld.depth_scale)
    z2 = sc * ((1 - wx) * (1 - wy) * d00 + wx * (1 - wy) * d10 + (1 - wx) * wy * d01 + wx * wy * d11)
    rz = z2 - Pw[2].astype(np.float64)
    dzu = sc * ((1 - wy) * (d10 - d00) + wy * (d11 - d01))
    dzv = sc * ((1 - wx) * (d01 - d00) + wx * (d11 - d10))
    Jw = ld.Jw.astype(np.float64)
    J = dzu[:, None] * Jw[:, 0, :] + dzv[:, None] * Jw[:, 1, :]
    X = ld.P[0].astype(np.float64)
    Y = ld.P[1].astype(np.float64)
    J[:, 2] -= 1.0
    J[:, 3] -= Y
    J[:, 4] += X
    return rz[valid].astype(F32), J[valid].astype(F32), valid


MAD_BINS = 2048      # |r| is histogrammed in 1/8 intensity steps (covers [0, 256))
MAD_BIN_SCALE = 8.0


def huber_mad_threshold(r: np.ndarray, c: float = 1.345) -> float:
    """Huber threshold from the residuals' median absolute value, as the CUDA kernel defines it: |r| is binned in
    1/8 intensity steps, the LOWER median bin (the ceil(n/2)-th smallest value) is taken at its centre, and
    k = c * 1.4826 * MAD, never below 1e-3.  Extension: the reference has no Huber weights (parity unpinned)."""
    n = r.size
    if n == 0:
        return 1e-3
    bins = np.minimum((np.abs(r.astype(F32)) * F32(MAD_BIN_SCALE)).astype(np.int64), MAD_BINS - 1)
    hist = np.bincount(bins, minlength=MAD_BINS)
    target = (n + 1) // 2
    b = int(np.searchsorted(np.cumsum(hist), target, side="left"))
    mad = F32((F32(b) + F32(0.5)) / F32(MAD_BIN_SCALE))
    return float(max(F32(c) * F32(1.4826) * mad, F32(1e-3)))


def normal_equations(r: np.ndarray, J: np.ndarray, weights: int = W_NONE, huber_k: float = 1.345 * 5.0,
                     tdist_kw: Optional[dict] = None):
    """base_robust_dvo.py:168-188.  Returns H (6,6) f32, b (6,) f32, err f32."""
    Jt = J.T.copy()
    if weights == W_NONE:
        err = np.mean(r ** 2) if r.size else F32(np.nan)
        Jw_, rw = J, r
    else:
        r2 = r * r
        if weights == W_TDIST_REF:
            wgt = tdist_weights(r2, **(tdist_kw or {}))
        elif weights == W_HUBER_MAD:
            wgt = huber_weights(r, huber_mad_threshold(r))
        else:
            wgt = huber_weights(r, huber_k)
        err = np.mean(wgt * r2) if r.size else F32(np.nan)
        rw = wgt * r
        Jw_ = wgt[:, None] * J
    H = Jt @ Jw_
    b = -(Jt @ rw)
    return H.astype(F32), b.astype(F32), F32(err)


def solve6(H: np.ndarray, b: np.ndarray) -> np.ndarray:
    """scipy.linalg.lstsq(lapack_driver="gelsy") at base_robust_dvo.py:196-198.

    SciPy is third-party (pinned scipy~=1.7.3, requirements.txt:9).  gelsy returns the
    minimum-norm least-squares solution with rcond = eps(f32); for the symmetric positive
    definite H of this path that is the unique solution, which NumPy's SVD-based lstsq also gives.
    """
    x, *_ = np.linalg.lstsq(H.astype(F32), b.astype(F32).reshape(6, 1), rcond=np.finfo(F32).eps)
    return x.reshape(6).astype(F32)


@dataclass
class EstimateResult:
    pose: Pose
    xi: np.ndarray
    iters: List[int]
    err_last: List[float]
    n_valid: List[int]
    trace: List[List[float]]


def estimate_pose(K, depth_scale, gray_prev_pyr, depth_prev_pyr, gray_cur_pyr, levels, init: Optional[Pose] = None,
                  weights: int = W_NONE, tolerance: float = 1e-6, max_iterations: int = 100,
                  max_increased_steps_allowed: int = 0, sigma: Optional[float] = None,
                  last_transform: Optional[Pose] = None, oob_mode: int = OOB_INCLUSIVE,
                  huber_k: float = 1.345 * 5.0, tdist_kw: Optional[dict] = None,
                  approximate_image2_gradient: bool = False, depth_cur_pyr=None,
                  depth_weight: float = 2500.0) -> EstimateResult:
    """BaseRobustDVO._step (base_robust_dvo.py:137-236): coarse-to-fine Gauss-Newton.

    depth_cur_pyr (extension, parity unpinned): the current frame's depth pyramid; when given, the depth term of
    depth_residuals_and_jacobian joins the normal equations: H += lambda J_Z^T J_Z, b -= lambda J_Z^T r_Z and
    err += lambda sum r_Z^2 / n with n the photometric residual count."""
    est = (init or Pose()).copy()
    iters = [0] * levels
    err_last = [float("nan")] * levels
    n_valid = [0] * levels
    trace: List[List[float]] = [[] for _ in range(levels)]
    for level in range(levels - 1, -1, -1):
        old = (last_transform or Pose()).copy()
        err_prev = np.finfo("float32").max
        inc_count = 0
        ld = prepare_level(K, depth_scale, gray_prev_pyr[level], depth_prev_pyr[level], gray_cur_pyr[level], level,
                           approximate_image2_gradient,
                           depth_cur_pyr[level] if depth_cur_pyr is not None else None)
        for i in range(max_iterations):
            r, J, _, valid = residuals_and_jacobian(ld, est.matrix(), oob_mode)
            H, b, err = normal_equations(r, J, weights, huber_k, tdist_kw)
            if depth_cur_pyr is not None:
                rz, Jz, _ = depth_residuals_and_jacobian(ld, est.matrix(), oob_mode)
                lam = F32(depth_weight)
                H = (H + lam * (Jz.T @ Jz)).astype(F32)
                b = (b - lam * (Jz.T @ rz)).astype(F32)
                err = F32(err + lam * F32(np.sum(rz.astype(np.float64) ** 2)) / F32(max(r.size, 1)))
            if sigma is not None:
                inv_cov = (1 / sigma) * np.eye(6, dtype=F32)
                H = H + inv_cov
                old_log = old.log()
                b = b + inv_cov @ old_log
                err = err + 0.5 * sigma * np.linalg.norm(old_log)
            iters[level] = i + 1
            err_last[level] = float(err)
            n_valid[level] = int(valid.sum())
            trace[level].append(float(err))
            inc = pose_from_xi(solve6(H, b))
            err_diff = err - err_prev
            if abs(err_diff) < tolerance:
                break
            if err_diff < 0.0:
                est = inc.compose(est)
                err_prev = err
                if sigma is not None:
                    old = inc.inverse().compose(old)
                inc_count = 0
            else:
                inc_count += 1
            if inc_count > max_increased_steps_allowed:
                break
    return EstimateResult(est, est.log(), iters, err_last, n_valid, trace)


class OracleDVO:
    """BaseDenseVisualOdometry.step (base_dense_visual_odometry.py:54-87) on top of estimate_pose."""

    def __init__(self, K, depth_scale, levels, initial_pose: Optional[Pose] = None, use_weighter=False,
                 max_increased_steps_allowed=0, sigma=None, tolerance=1e-6, max_iterations=100,
                 max_distance=5.0, oob_mode=OOB_INCLUSIVE, weights: Optional[int] = None, huber_k=1.345 * 5.0,
                 approximate_image2_gradient=False, use_depth_residual=False, depth_weight=2500.0):
        self.use_depth_residual = use_depth_residual
        self.depth_weight = depth_weight
        self.K = np.asarray(K, dtype=F32)[:3, :3]
        self.depth_scale = depth_scale
        self.levels = levels
        self.current_pose = (initial_pose or Pose()).copy()
        self.weights = weights if weights is not None else (W_TDIST_REF if use_weighter else W_NONE)
        self.kw = dict(tolerance=tolerance, max_iterations=max_iterations,
                       max_increased_steps_allowed=max_increased_steps_allowed, sigma=sigma, oob_mode=oob_mode,
                       huber_k=huber_k, approximate_image2_gradient=approximate_image2_gradient)
        self.max_distance = max_distance
        self._gray_prev = None
        self._depth_prev = None
        self._last = None
        self.last_result: Optional[EstimateResult] = None

    def step(self, color_image, depth_image, init_guess: Optional[Pose] = None) -> Pose:
        gray = bgr_to_gray(color_image)
        depth_image[(depth_image * self.depth_scale) > self.max_distance] = 0
        if self._gray_prev is None:
            T = Pose()
        else:
            res = estimate_pose(self.K, self.depth_scale, build_pyramid(self._gray_prev, self.levels),
                                build_pyramid(self._depth_prev, self.levels), build_pyramid(gray, self.levels),
                                self.levels, init=init_guess, weights=self.weights, last_transform=self._last,
                                depth_cur_pyr=(build_pyramid(depth_image, self.levels)
                                               if self.use_depth_residual else None),
                                depth_weight=self.depth_weight, **self.kw)
            self.last_result = res
            T = res.pose
        self._last = T.copy()
        self.current_pose = self.current_pose.compose(T.inverse())
        self._gray_prev = gray
        self._depth_prev = depth_image
        return T
