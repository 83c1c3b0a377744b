"""CPU checks of the oracle's depth (geometric) residual, an extension the reference does not have (SURVEY F4,
PARITY UNPINNED: there is no golden vector to pin it to).  What can be checked without a reference: the term
vanishes where it must, its Jacobian is the derivative of its residual where the reference's own
J_w-at-the-untransformed-point convention is exact (at the identity), and the estimator still finds the known
motion of a synthetic scene with it."""
import numpy as np

from oracle import dvo_oracle as O
from dense_visual_odometry_b200.synthetic import make_pairs_numpy


def _Km(K):
    return np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], dtype=np.float32)


def _level0(seed=3, h=120, w=160):
    d = make_pairs_numpy([seed], height=h, width=w)
    Km = _Km(d["K"])
    gp, gc = O.bgr_to_gray(d["bgr_prev"][0]), O.bgr_to_gray(d["bgr_cur"][0])
    dp = O.clamp_depth(d["depth_prev"][0].copy(), d["depth_scale"])
    dc = O.clamp_depth(d["depth_cur"][0].copy(), d["depth_scale"])
    return d, Km, gp, gc, dp, dc


def test_depth_term_vanishes_on_identical_frames():
    d, Km, gp, gc, dp, dc = _level0()
    ld = O.prepare_level(Km, d["depth_scale"], gp, dp, gp, 0, depth_cur=dp)
    rz, Jz, valid = O.depth_residuals_and_jacobian(ld, np.eye(4, dtype=np.float32))
    # valid exactly where the pixel and its right / lower / diagonal neighbours have depth
    nz = dp != 0
    expect = np.zeros_like(nz)
    expect[:-1, :-1] = nz[:-1, :-1] & nz[:-1, 1:] & nz[1:, :-1] & nz[1:, 1:]
    # (the warped coordinate of pixel (u, v) at the identity is (u, v) up to float32 rounding of the projection)
    got = np.zeros(nz.size, bool)
    got[np.flatnonzero(nz.reshape(-1))[valid]] = True
    assert (got.reshape(nz.shape) != expect).mean() < 0.02
    assert rz.size > 0.5 * nz.sum()
    # Z2 at (almost) the pixel centre is the pixel's own depth, up to the sub-pixel rounding of the projection
    # (|du| ~ 1e-5 px) times the local depth slope
    assert np.abs(rz).max() < 2e-5
    assert Jz.shape == (rz.size, 6) and np.isfinite(Jz).all()


def test_depth_jacobian_is_the_derivative_at_identity():
    d, Km, gp, gc, dp, dc = _level0(h=240, w=320)
    ld = O.prepare_level(Km, d["depth_scale"], gp, dp, gc, 0, depth_cur=dc)
    n = ld.P.shape[1]
    rz, Jz, v = O.depth_residuals_and_jacobian(ld, np.eye(4, dtype=np.float32))

    def full(r, vv):
        a = np.full(n, np.nan)
        a[vv] = r
        return a

    eps = 2e-4
    for i in range(6):
        xi = np.zeros(6, np.float32)
        xi[i] = eps
        rp, _, vp = O.depth_residuals_and_jacobian(ld, O.pose_from_xi(xi).matrix())
        xi[i] = -eps
        rm, _, vm = O.depth_residuals_and_jacobian(ld, O.pose_from_xi(xi).matrix())
        both = v & vp & vm
        fd = (full(rp, vp)[both] - full(rm, vm)[both]) / (2 * eps)
        J = full(Jz[:, i], v)[both]
        # the bilinear patch is piecewise: compare robustly (medians), not pixel by pixel
        assert np.median(np.abs(fd - J)) < 0.03 * max(np.median(np.abs(J)), 0.05), i
        assert np.corrcoef(fd, J)[0, 1] > 0.9 or np.std(J) < 1e-3, i   # depth is quantised to 0.2 mm: noisy slopes


def test_estimate_with_depth_term_finds_the_motion():
    d, Km, gp, gc, dp, dc = _level0(seed=8, h=240, w=320)
    pyr = lambda a: O.build_pyramid(a, 4)  # noqa: E731
    plain = O.estimate_pose(Km, d["depth_scale"], pyr(gp), pyr(dp), pyr(gc), 4)
    both = O.estimate_pose(Km, d["depth_scale"], pyr(gp), pyr(dp), pyr(gc), 4, depth_cur_pyr=pyr(dc))
    assert np.abs(both.xi - d["xi"][0]).max() < 2e-3
    assert np.abs(both.xi - plain.xi).max() > 1e-7     # the term takes part
    # lambda = 0 reduces to the photometric estimate exactly
    zero = O.estimate_pose(Km, d["depth_scale"], pyr(gp), pyr(dp), pyr(gc), 4, depth_cur_pyr=pyr(dc), depth_weight=0.0)
    assert np.array_equal(zero.pose.q, plain.pose.q) and np.array_equal(zero.pose.t, plain.pose.t)
