"""GPU tests of the depth (geometric) residual extension (BASELINE.json configs[4]: "photometric + depth residual").

PARITY UNPINNED: the reference has no depth residual, so there are no golden vectors.  The CUDA path is
compared with the oracle's float64 restatement of the same definition (oracle/dvo_oracle.py,
depth_residuals_and_jacobian) on the same inputs.  Tolerances: validity mask bit-exact (it is decided on the same
bit-identical warped coordinates as the photometric mask); r_Z and J_Z within 1e-5 of the plane's max magnitude
(+ 2e-7 m, the float32 resolution of a 2-3 m depth); final pose within 1e-4 rad / 1e-4 m."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import dvo_oracle as O  # noqa: E402

POSE_TOL = 1e-4


def _Km(K):
    return np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], dtype=np.float32)


@pytest.fixture(scope="module")
def dvo_mod():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import dense_visual_odometry_b200 as m
    return m


def _dense(vals, valid_n, mask):
    """(Nz, ...) values over valid depth pixels -> dense [H,W,...] with NaN / 0 elsewhere."""
    h, w = mask.shape
    full_valid = np.zeros(h * w, bool)
    idx = np.flatnonzero(mask.reshape(-1))[valid_n]
    full_valid[idx] = True
    out = np.zeros((h * w,) + vals.shape[1:], np.float64)
    out[idx] = vals
    return out.reshape((h, w) + vals.shape[1:]), full_valid.reshape(h, w)


def _compare_depth_level(est, ld, pose, lv, oob, lam, report):
    rz, Jz, vz = O.depth_residuals_and_jacobian(ld, pose.exp(), oob)
    rd, vd = _dense(rz.astype(np.float64), vz, ld.mask)
    Jd, _ = _dense(Jz.astype(np.float64), vz, ld.mask)
    gr, gJ, gv, acc = est.depth_residuals_dense(pose, lv, est._hook_slots[0], est._hook_slots[1])
    n_mis = int((gv != vd).sum())
    report[f"L{lv}_valid_mismatch"] = n_mis
    assert n_mis == 0, f"depth-term validity differs at {n_mis} pixels"
    assert vd.sum() > 0
    dr = np.abs(gr[vd] - rd[vd])
    rs = float(np.abs(rd[vd]).max())
    dJ = np.abs(gJ[vd] - Jd[vd])
    Js = float(np.abs(Jd[vd]).max())
    report[f"L{lv}_r_abs"] = float(dr.max())
    report[f"L{lv}_J_rel"] = float(dJ.max() / Js)
    assert dr.max() <= 1e-5 * rs + 2e-7, f"depth residuals differ by {dr.max()} m (scale {rs})"
    assert dJ.max() <= 1e-5 * Js, f"depth Jacobians differ by {dJ.max()} (scale {Js})"
    assert np.isnan(gr[~vd]).all() and not gJ[~vd].any()
    J64, r64 = Jz.astype(np.float64), rz.astype(np.float64)
    H = lam * (J64.T @ J64)
    g = lam * (J64.T @ r64)
    iu = np.triu_indices(6)
    np.testing.assert_allclose(acc[:21], H[iu], rtol=2e-5, atol=2e-5 * np.abs(H).max())
    np.testing.assert_allclose(acc[21:27], g, rtol=2e-5, atol=2e-5 * np.abs(g).max())
    np.testing.assert_allclose(acc[27], lam * (r64 ** 2).sum(), rtol=2e-5)
    assert acc[28] == rz.shape[0]


@pytest.mark.parametrize("oob", ["inclusive", "strict"])
def test_depth_term_dense_vs_oracle_testdata(dvo_mod, testdata_frames, oob):
    """r_Z, J_Z, validity and the term's share of the normal equations, per pixel, four levels, three poses, on the
    reference's own test frames (real sensor depth with holes)."""
    m = dvo_mod
    f = testdata_frames
    Km = _Km(f["K"])
    lam = 400.0
    cam = m.RGBDCameraModel(Km, f["depth_scale"])
    est = m.get_dvo("robust-dvo", cam, m.Se3.identity(), levels=4, oob_mode=oob, use_depth_residual=True,
                    depth_weight=lam)
    est.step(f["bgr"][0], f["depth"][0].copy())
    d1 = O.clamp_depth(f["depth"][1], f["depth_scale"])
    est._build_pyramids(O.bgr_to_gray(f["bgr"][1]), d1)
    gp0 = O.build_pyramid(O.bgr_to_gray(f["bgr"][0]), 4)
    dp0 = O.build_pyramid(O.clamp_depth(f["depth"][0], f["depth_scale"]), 4)
    gp1 = O.build_pyramid(O.bgr_to_gray(f["bgr"][1]), 4)
    dp1 = O.build_pyramid(d1, 4)
    poses = [m.Se3.identity(),
             m.Se3.from_se3(np.array([[0.0017], [-0.0072], [-0.0108], [0.005], [0.0065], [0.0034]], np.float32)),
             m.Se3.from_se3(np.array([[0.05], [-0.03], [0.08], [-0.04], [0.03], [0.06]], np.float32))]
    report = {}
    for lv in range(4):
        ld = O.prepare_level(Km, f["depth_scale"], gp0[lv], dp0[lv], gp1[lv], lv, depth_cur=dp1[lv])
        for pose in poses:
            _compare_depth_level(est, ld, pose, lv, O.OOB_STRICT if oob == "strict" else O.OOB_INCLUSIVE, lam, report)
    print("depth-term dense parity:", oob, report)


def test_depth_pose_testdata_sequence_vs_oracle(dvo_mod, testdata_frames):
    """step() with the depth term on the reference's test frames against the oracle with the same option."""
    m = dvo_mod
    f = testdata_frames
    Km = _Km(f["K"])
    cam = m.RGBDCameraModel(Km, f["depth_scale"])
    est = m.get_dvo("robust-dvo", cam, m.Se3.identity(), levels=4, use_depth_residual=True)
    ref = O.OracleDVO(Km, f["depth_scale"], 4, use_depth_residual=True)
    plain = O.OracleDVO(Km, f["depth_scale"], 4)
    for i in range(4):
        T = est.step(f["bgr"][i], f["depth"][i].copy())
        Tr = ref.step(f["bgr"][i], f["depth"][i].copy())
        Tp = plain.step(f["bgr"][i], f["depth"][i].copy())
        if i == 0:
            continue
        assert T is not None and est.last_stats["flags"][0] == 0
        print("pair", i, "iters", est.last_stats["iters"][0][:4].tolist(), ref.last_result.iters)
        # (iteration counts may differ by a few at the end of a level: the stop rule |d err| < 1e-6 sits at the
        # float32 resolution of err, and the two sides sum lambda r_Z^2 in different orders)
        assert np.abs(T.so3.quat.reshape(4) - Tr.q).max() < POSE_TOL
        assert np.abs(T.tvec.reshape(3) - Tr.t).max() < POSE_TOL
        # the option is live: the estimate is not the photometric-only one
        assert max(np.abs(Tr.q - Tp.q).max(), np.abs(Tr.t - Tp.t).max()) > 1e-6


@pytest.mark.parametrize("weights,oob", [("none", "inclusive"), ("huber", "strict")])
def test_depth_pose_batch_synthetic_vs_oracle_and_truth(dvo_mod, golden_dir, weights, oob):
    """PairBatchAligner (persistent kernel, 128 threads x 2 CTAs) with the depth term: every pair against the oracle,
    the known motion, and independence of the batch position."""
    from dense_visual_odometry_b200.synthetic import make_pairs_numpy
    m = dvo_mod
    d = make_pairs_numpy([21, 22, 23], height=240, width=320)
    Km = _Km(d["K"])
    cam = m.RGBDCameraModel(Km, d["depth_scale"])
    al = m.PairBatchAligner(cam, 240, 320, 4, max_pairs=3, weights=weights, oob_mode=oob, use_depth_residual=True)
    qt, stats = al.align(d["bgr_prev"], d["depth_prev"].copy(), d["bgr_cur"], d["depth_cur"].copy())
    assert not stats["flags"].any()
    wm = {"none": O.W_NONE, "huber": O.W_HUBER}[weights]
    om = O.OOB_STRICT if oob == "strict" else O.OOB_INCLUSIVE
    for p in range(3):
        gp, gc = O.bgr_to_gray(d["bgr_prev"][p]), O.bgr_to_gray(d["bgr_cur"][p])
        dp = O.clamp_depth(d["depth_prev"][p].copy(), d["depth_scale"])
        dc = O.clamp_depth(d["depth_cur"][p].copy(), d["depth_scale"])
        res = O.estimate_pose(Km, d["depth_scale"], O.build_pyramid(gp, 4), O.build_pyramid(dp, 4),
                              O.build_pyramid(gc, 4), 4, weights=wm, oob_mode=om, depth_cur_pyr=O.build_pyramid(dc, 4))
        print("pair", p, "iters", stats["iters"][p][:4].tolist(), res.iters)
        assert np.abs(qt[p, :4] - res.pose.q).max() < POSE_TOL and np.abs(qt[p, 4:] - res.pose.t).max() < POSE_TOL
        xi = m.Se3.from_qt(qt[p]).log().reshape(6)
        assert np.abs(xi - d["xi"][p]).max() < 2e-3   # the reference's stop rule ends short of the true motion
    # a pair's result does not depend on its position in the batch
    al1 = m.PairBatchAligner(cam, 240, 320, 4, max_pairs=1, weights=weights, oob_mode=oob, use_depth_residual=True)
    q1, _ = al1.align(d["bgr_prev"][2:3], d["depth_prev"][2:3].copy(), d["bgr_cur"][2:3], d["depth_cur"][2:3].copy())
    assert np.array_equal(q1[0], qt[2])


@pytest.mark.parametrize("h,w", [(720, 1280), (1080, 1920)])
def test_depth_high_resolution_five_levels(dvo_mod, h, w):
    """BASELINE.json configs[4]: a 1280x720 and a 1920x1080 pair, 5-level pyramid, photometric + depth residual: the
    pose against the oracle, and the depth term per pixel at the coarsest and the finest level."""
    from dense_visual_odometry_b200.synthetic import make_pairs_numpy, TUM_FR1
    m = dvo_mod
    s = w / 640.0
    K = (TUM_FR1[0] * s, TUM_FR1[1] * s, TUM_FR1[2] * s, TUM_FR1[3] * s)
    d = make_pairs_numpy([5], height=h, width=w, K=K)
    Km = _Km(K)
    cam = m.RGBDCameraModel(Km, d["depth_scale"])
    est = m.get_dvo("robust-dvo", cam, m.Se3.identity(), levels=5, use_depth_residual=True)
    est.step(d["bgr_prev"][0], d["depth_prev"][0].copy())
    T = est.step(d["bgr_cur"][0], d["depth_cur"][0].copy())
    assert T is not None and est.last_stats["flags"][0] == 0
    ref = O.OracleDVO(Km, d["depth_scale"], 5, use_depth_residual=True)
    ref.step(d["bgr_prev"][0], d["depth_prev"][0].copy())
    Tr = ref.step(d["bgr_cur"][0], d["depth_cur"][0].copy())
    print("iters", est.last_stats["iters"][0][:5].tolist(), ref.last_result.iters)
    assert np.abs(T.so3.quat.reshape(4) - Tr.q).max() < POSE_TOL
    assert np.abs(T.tvec.reshape(3) - Tr.t).max() < POSE_TOL
    gp, gc = O.bgr_to_gray(d["bgr_prev"][0]), O.bgr_to_gray(d["bgr_cur"][0])
    dp = O.clamp_depth(d["depth_prev"][0].copy(), d["depth_scale"])
    dc = O.clamp_depth(d["depth_cur"][0].copy(), d["depth_scale"])
    pg, pd, cg, cd = (O.build_pyramid(a, 5) for a in (gp, dp, gc, dc))
    est = m.get_dvo("robust-dvo", cam, m.Se3.identity(), levels=5, use_depth_residual=True)  # hooks: previous = stored
    est.step(d["bgr_prev"][0], d["depth_prev"][0].copy())
    est._build_pyramids(gc, dc)
    report = {}
    for lv in (4, 0):
        ld = O.prepare_level(Km, d["depth_scale"], pg[lv], pd[lv], cg[lv], lv, depth_cur=cd[lv])
        _compare_depth_level(est, ld, m.Se3.identity(), lv, O.OOB_INCLUSIVE, 2500.0, report)
    print(f"{w}x{h} depth-term parity:", report)


def test_depth_option_errors_are_loud(dvo_mod):
    m = dvo_mod
    cam = m.RGBDCameraModel(_Km((100.0, 100.0, 32.0, 24.0)), 0.001)
    for kw in (dict(weights="tdist"), dict(weights="huber_mad"), dict(approximate_image2_gradient=True)):
        with pytest.raises(ValueError):
            m.get_dvo("robust-dvo", cam, m.Se3.identity(), levels=2, height=48, width=64, use_depth_residual=True, **kw)
