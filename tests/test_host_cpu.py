"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol, the Lie /
camera mirrors agree with the reference's golden vectors, and host-side errors are raised without a GPU."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import dense_visual_odometry_b200 as m
    return m


def test_cabi_exports_every_declared_symbol(built):
    from dense_visual_odometry_b200 import _cabi
    header = (ROOT / "include" / "dvo_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dvo_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dvo_b200.h but not exported"
    assert declared == set(_cabi.SYMBOLS), (declared ^ set(_cabi.SYMBOLS))


def test_default_config_matches_reference_defaults(built):
    import ctypes as C
    from dense_visual_odometry_b200 import _cabi
    cfg = _cabi.dvo_config()
    _cabi.load().dvo_default_config(C.byref(cfg))
    # base_robust_dvo.py:34-38, t_weighter.py:14, base_dense_visual_odometry.py:27
    assert (cfg.max_iterations, cfg.max_increased_steps, cfg.weights, cfg.oob_mode) == (100, 0, 0, 0)
    assert cfg.tolerance == pytest.approx(1e-6) and cfg.sigma_prior < 0
    assert (cfg.tdist_dof, cfg.tdist_init_sigma, cfg.tdist_max_iterations) == (5.0, 5.0, 50)
    assert cfg.tdist_tolerance == pytest.approx(1e-3) and cfg.max_distance == 5.0


def test_null_handle_calls_fail_without_gpu(built):
    from dense_visual_odometry_b200 import _cabi
    lib = _cabi.load()
    assert lib.dvo_destroy(None) < 0
    assert lib.dvo_set_intrinsics(None, 1.0, 1.0, 0.0, 0.0, 1.0) < 0
    assert lib.dvo_last_error(None) == b"null handle"
    assert lib.dvo_launch_count(None) == 0


def test_lie_mirror_vs_reference_golden(built, golden_dir):
    m = built
    ka = np.load(golden_dir / "known_answers.npz")
    xis = ka["lie_xi"]
    poses = [m.Se3.from_se3(x.reshape(6, 1)) for x in xis]
    np.testing.assert_allclose(np.stack([p.so3.quat.reshape(4) for p in poses]), ka["lie_q"], atol=1e-7)
    np.testing.assert_allclose(np.stack([p.tvec.reshape(3) for p in poses]), ka["lie_t"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(np.stack([p.exp() for p in poses]), ka["lie_T"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(np.stack([p.log().reshape(6) for p in poses]), ka["lie_log"], rtol=1e-5, atol=2e-6)
    prod = [poses[i] * poses[i + 1] for i in range(7)]
    np.testing.assert_allclose(np.stack([p.so3.quat.reshape(4) for p in prod]), ka["lie_prod_q"], atol=1e-7)
    np.testing.assert_allclose(np.stack([p.tvec.reshape(3) for p in prod]), ka["lie_prod_t"], rtol=1e-6, atol=1e-7)
    inv = [p.inverse() for p in poses]
    np.testing.assert_allclose(np.stack([p.so3.quat.reshape(4) for p in inv]), ka["lie_inv_q"], atol=1e-7)
    np.testing.assert_allclose(np.stack([p.tvec.reshape(3) for p in inv]), ka["lie_inv_t"], rtol=1e-6, atol=1e-7)
    # identities (reference tests test_special_euclidean_group.py)
    I = m.Se3.identity()
    assert np.array_equal(I.exp(), np.eye(4, dtype=np.float32)) and not I.log().any()
    assert (poses[2] * poses[2].inverse()) == I
    qt = m.pose_to_qt(poses[3])
    assert m.Se3.from_qt(qt) == poses[3]


def test_camera_model_mirror(built, golden_dir, testdata_frames):
    m = built
    ka = np.load(golden_dir / "known_answers.npz")
    K = testdata_frames["K"]
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], dtype=np.float32)
    cam = m.RGBDCameraModel(Km, testdata_frames["depth_scale"])
    assert cam.intrinsics.shape == (3, 4) and cam.intrinsics.dtype == np.float32
    for lv in (0, 2):
        P, mask = cam.deproject(ka["dsmall"], return_mask=True, level=lv)
        np.testing.assert_array_equal(mask, ka[f"mask_L{lv}"])
        np.testing.assert_allclose(P, ka[f"P_L{lv}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(cam.project(P.copy(), level=lv), ka[f"uv_L{lv}"], rtol=1e-6, atol=1e-5)
    with pytest.raises(AssertionError):
        m.RGBDCameraModel(np.eye(4), 1.0)
    with pytest.raises(AssertionError):
        m.RGBDCameraModel(np.eye(3), -1.0)
    assert m.RGBDCameraModel.load_from_yaml(Path("/nonexistent.yaml")) is None


def test_factory_errors_without_gpu(built):
    m = built
    cam = m.RGBDCameraModel(np.eye(3, dtype=np.float32), 1.0)
    with pytest.raises(ValueError):
        m.get_dvo("loftr", cam, m.Se3.identity(), levels=1)
    with pytest.raises(ValueError):
        m.get_dvo("robust-dvo", cam, m.Se3.identity(), levels=1, use_gpu=False)
    with pytest.raises(ValueError):   # constructor errors are wrapped like the reference wraps them
        m.get_dvo("robust-dvo", cam, m.Se3.identity(), levels=99)


def test_product_does_not_import_oracle():
    """The shipped package must never route through oracle/ (no CPU fallback)."""
    pkg = ROOT / "dense-visual-odometry_b200"
    for p in pkg.glob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p


def test_ctypes_structs_match_the_c_header(built, tmp_path):
    """The ctypes mirrors of dvo_config / dvo_pair_stats have the size and field offsets a C compiler gives the
    structs of include/dvo_b200.h (the header is plain C: it must compile with gcc, no CUDA)."""
    import ctypes as C
    import shutil
    import subprocess
    from dense_visual_odometry_b200 import _cabi
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    fields = {"dvo_config": [f[0] for f in _cabi.dvo_config._fields_],
              "dvo_pair_stats": [f[0] for f in _cabi.dvo_pair_stats._fields_]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "dvo_b200.h"}"', 'int main(void) {']
    for s, names in fields.items():
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for n in names:
            lines.append(f'printf("{s}.{n} %zu\\n", offsetof({s}, {n}));')
    lines += ['printf("DVO_MAX_LEVELS %d\\n", DVO_MAX_LEVELS);', 'printf("DVO_ACC_TERMS %d\\n", DVO_ACC_TERMS);',
              'printf("W %d %d %d %d\\n", DVO_W_NONE, DVO_W_TDIST_REF, DVO_W_HUBER, DVO_W_HUBER_MAD);',
              'printf("OOB %d %d\\n", DVO_OOB_INCLUSIVE, DVO_OOB_STRICT);', 'return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True)
    out = dict(l.split(" ", 1) for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for s, names in fields.items():
        cls = getattr(_cabi, s)
        assert int(out[s]) == C.sizeof(cls), s
        for n in names:
            assert int(out[f"{s}.{n}"]) == getattr(cls, n).offset, f"{s}.{n}"
    assert int(out["DVO_MAX_LEVELS"]) == _cabi.DVO_MAX_LEVELS and int(out["DVO_ACC_TERMS"]) == _cabi.DVO_ACC_TERMS
    assert out["W"].split() == [str(v) for v in (_cabi.W_NONE, _cabi.W_TDIST_REF, _cabi.W_HUBER, _cabi.W_HUBER_MAD)]
    assert out["OOB"].split() == [str(_cabi.OOB_INCLUSIVE), str(_cabi.OOB_STRICT)]


def test_sequence_launch_groups(built):
    """SequenceAligner.align: upload chunks cover the frames once, in order; single-chunk groups first, everything
    within two chunks of the end as one group (so the last launch has more pairs than CTAs whenever the stream allows)."""
    from dense_visual_odometry_b200.estimator import sequence_launch_groups as groups
    for n, c in ((1000, 256), (600, 256), (512, 256), (300, 256), (2, 256), (1000, 64), (1025, 256), (999, 1000)):
        g = groups(n, c)
        flat = [ch for grp in g for ch in grp]
        assert flat[0][0] == 0 and flat[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(flat, flat[1:]))          # contiguous, no overlap
        assert all(0 < hi - lo <= c for lo, hi in flat)
        assert all(len(grp) == 1 for grp in g[:-1])                        # single chunks before the last group
        assert n - g[-1][0][0] <= 2 * c                                    # the last group: within two chunks of the end
        if len(g) > 1:
            assert n - g[-2][0][0] > 2 * c
    assert groups(1000, 256) == [[(0, 256)], [(256, 512)], [(512, 768), (768, 1000)]]
    assert groups(1000, 1000) == [[(0, 1000)]]
    with pytest.raises(ValueError):
        groups(10, 0)
