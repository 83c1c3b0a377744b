"""Generate golden vectors by running the UNMODIFIED reference sources from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

External shims applied, none of which touches a reference file:
  1. NUMBA_ENABLE_CUDASIM=1 before import, so the eager @cuda.jit decorators import without a GPU.
  2. numpy.bool8 = numpy.bool_ (annotation used by the reference, gone in NumPy 2).
  3. RobustDVOCPU.interpolate_bilinear replaced by the same body plus the missing out-of-image
     `continue` (the unmodified function reads out of bounds).  Two variants: inclusive
     (default) and strict.
Everything else (pyramids, Sobel, deprojection, J_w, weighter, lstsq, Lie algebra, GN loop) is the
reference's own code.
"""
import hashlib
import json
import math
import os
import sys
from pathlib import Path

os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
import numpy as np  # noqa: E402

np.bool8 = np.bool_

REF = Path(os.environ.get("DVO_REFERENCE_SRC", "/root/reference/src"))
ROOT = Path(__file__).resolve().parents[2]
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(REF))
sys.path.insert(0, str(ROOT))

import cv2  # noqa: E402
import numba as nb  # noqa: E402

from dense_visual_odometry.core import get_dvo  # noqa: E402
from dense_visual_odometry.camera_model import RGBDCameraModel  # noqa: E402
from dense_visual_odometry.utils.lie_algebra import Se3, So3  # noqa: E402
from dense_visual_odometry.core.robust_dense_visual_odometry.cpu_robust_dense_visual_odometry import RobustDVOCPU  # noqa: E402
from dense_visual_odometry.utils.jacobian import compute_gradients, compute_jacobian_of_warp_function  # noqa: E402
from dense_visual_odometry.utils.image_pyramid import ImagePyramid  # noqa: E402

import dense_visual_odometry_b200  # noqa: E402,F401  (repo-root shim that loads the hyphenated package dir)
from dense_visual_odometry_b200.synthetic import make_pairs_numpy  # noqa: E402

_SIGS = ['float32[:,:](uint8[:,:], float32[:,:])', 'float32[:,:](float32[:,:], float32[:,:])']


@nb.njit(_SIGS, parallel=True, fastmath=True)
def _interp_inclusive(image, pixels_coordinates):
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


@nb.njit(_SIGS, parallel=True, fastmath=True)
def _interp_strict(image, pixels_coordinates):
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


def set_guard(mode):
    RobustDVOCPU.interpolate_bilinear = staticmethod(_interp_inclusive if mode == "inclusive" else _interp_strict)


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def camera(K, scale):
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], dtype=np.float32)
    return RGBDCameraModel(Km, scale)


def run_pair(cam, levels, bgr0, d0, bgr1, d1, capture_levels=True, stride=97, **kw):
    """One reference pose estimate with per-call instrumentation of compute_residuals_and_jacobian."""
    dvo = get_dvo("robust-dvo", cam, Se3.identity(), levels=levels, **kw)
    calls = []
    first = {}
    orig = dvo.compute_residuals_and_jacobian

    def spy(estimate, level=0):
        r, J, mask = orig(estimate=estimate, level=level)
        calls.append((level, float(np.mean(r ** 2)) if r.size else float("nan"), int(r.shape[0])))
        if capture_levels and level not in first:
            Jt = J.T.copy()
            H = Jt @ J
            b = -Jt @ r
            n = r.shape[0]
            idx = np.arange(0, n, stride if n > 20000 else 1)
            first[level] = dict(T=estimate.exp().copy(), H=H, b=b.reshape(-1), err=np.float32(np.mean(r ** 2)),
                                n=n, n_depth=int(mask.sum()), idx=idx, r=r.reshape(-1)[idx].copy(),
                                J=J[idx].copy(), mask_digest=digest(mask.astype(np.uint8)))
        return r, J, mask

    dvo.compute_residuals_and_jacobian = spy
    lambdas = []
    if dvo._weighter is not None:
        worig = dvo._weighter.weight

        def wspy(residuals_squared):
            w = worig(residuals_squared=residuals_squared)
            # w = (dof+1)/(dof + r2*lam)  ->  lam from the largest residual
            k = int(np.argmax(residuals_squared))
            r2 = float(residuals_squared.reshape(-1)[k])
            lambdas.append((6.0 / float(w.reshape(-1)[k]) - 5.0) / r2 if r2 > 0 else 0.0)
            return w

        dvo._weighter.weight = wspy
    d0 = d0.copy()
    d1 = d1.copy()
    dvo.step(bgr0, d0)
    T = dvo.step(bgr1, d1)
    iters = [sum(1 for c in calls if c[0] == lv) for lv in range(levels)]
    errs = [[c[1] for c in calls if c[0] == lv] for lv in range(levels)]
    nval = [[c[2] for c in calls if c[0] == lv] for lv in range(levels)]
    out = dict(q=T.so3.quat.reshape(4).astype(np.float32), t=T.tvec.reshape(3).astype(np.float32),
               xi=T.log().reshape(6).astype(np.float32), T=T.exp().astype(np.float32),
               iters=np.array(iters), err_last=np.array([e[-1] for e in errs], dtype=np.float32),
               n_last=np.array([n[-1] for n in nval]),
               cur_pose_T=dvo.current_pose.exp().astype(np.float32))
    for lv in range(levels):
        out[f"errs_L{lv}"] = np.array(errs[lv], dtype=np.float32)
    if lambdas:
        out["lambdas"] = np.array(lambdas)
    for lv, f in first.items():
        for k, v in f.items():
            out[f"L{lv}_{k}"] = np.asarray(v)
    return out, dvo


def load_test_frames():
    td = REF.parent / "tests" / "test_data"
    gt = json.loads((td / "ground_truth.json").read_text())
    bgr, depth, poses = [], [], []
    for k in sorted(gt, key=int):
        bgr.append(cv2.imread(str(td / gt[k]["rgb"]), cv2.IMREAD_ANYCOLOR))
        depth.append(cv2.imread(str(td / gt[k]["depth"]), cv2.IMREAD_ANYDEPTH))
        poses.append(np.array(gt[k]["transformation"]))
    return np.stack(bgr), np.stack(depth), np.stack(poses)


def main():
    set_guard("inclusive")
    K = (517.3, 516.5, 318.6, 239.5)
    scale = 0.0002
    cam = camera(K, scale)

    # ---------------- inputs: the reference's own test frames --------------------------------
    bgr, depth, gt = load_test_frames()
    np.savez_compressed(OUT / "frames_testdata.npz", bgr=bgr, depth=depth, gt=gt, K=np.array(K), depth_scale=scale)

    # ---------------- primitives: gray, clamp, pyramids, Sobel digests (bit-exact class) ------
    prim = {}
    for i in range(bgr.shape[0]):
        gray = cv2.cvtColor(bgr[i], cv2.COLOR_BGR2GRAY)
        d = depth[i].copy()
        d[(d * scale) > 5.0] = 0
        gp = ImagePyramid(4, gray)
        dp = ImagePyramid(4, d)
        prim[f"f{i}_clamped"] = int((d != depth[i]).sum())
        for lv in range(4):
            prim[f"f{i}_gray_L{lv}"] = digest(gp.at(lv))
            prim[f"f{i}_depth_L{lv}"] = digest(dp.at(lv))
            gx, gy = compute_gradients(gp.at(lv), kernel_size=3)
            prim[f"f{i}_gx_L{lv}"] = digest(gx)
            prim[f"f{i}_gy_L{lv}"] = digest(gy)
    # odd-sized pyramid + gray lattice check
    rng = np.random.default_rng(7)
    odd8 = rng.integers(0, 256, (77, 101), dtype=np.uint8)
    odd16 = rng.integers(0, 65536, (77, 101), dtype=np.uint16)
    odd16[rng.random((77, 101)) < 0.3] = 0
    for lv, (a, b) in enumerate(zip(ImagePyramid(4, odd8)._pyramid, ImagePyramid(4, odd16)._pyramid)):
        prim[f"odd_gray_L{lv}"] = digest(a)
        prim[f"odd_depth_L{lv}"] = digest(b)
        prim[f"odd_shape_L{lv}"] = list(a.shape)
        gx, gy = compute_gradients(a, kernel_size=3)
        prim[f"odd_gx_L{lv}"] = digest(gx)
        prim[f"odd_gy_L{lv}"] = digest(gy)
    lat = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    prim["lattice_gray"] = digest(cv2.cvtColor(lat, cv2.COLOR_BGR2GRAY))
    for lv in range(5):
        prim[f"K_L{lv}"] = [float(v) for v in cam.at(lv)[:3, :3].reshape(-1)]
    (OUT / "primitives.json").write_text(json.dumps(prim, indent=1))

    # ---------------- known-answer: J_w, deproject/project on random points --------------------
    dsmall = rng.integers(0, 20000, (12, 16)).astype(np.uint16)
    dsmall[rng.random((12, 16)) < 0.25] = 0
    ka = {}
    for lv in (0, 2):
        P, mask = cam.deproject(dsmall, return_mask=True, level=lv)
        ka[f"P_L{lv}"] = P
        ka[f"mask_L{lv}"] = mask
        ka[f"Jw_L{lv}"] = compute_jacobian_of_warp_function(P, cam.at(lv))
        ka[f"uv_L{lv}"] = cam.project(P.copy(), level=lv)
    ka["dsmall"] = dsmall
    # Lie algebra known answers
    xis = rng.uniform(-0.05, 0.05, (8, 6)).astype(np.float32)
    xis[0, 3:] = 0
    xis[1] *= 30
    ka["lie_xi"] = xis
    ka["lie_q"] = np.stack([Se3.from_se3(x.reshape(6, 1)).so3.quat.reshape(4) for x in xis])
    ka["lie_t"] = np.stack([Se3.from_se3(x.reshape(6, 1)).tvec.reshape(3) for x in xis])
    ka["lie_T"] = np.stack([Se3.from_se3(x.reshape(6, 1)).exp() for x in xis])
    ka["lie_log"] = np.stack([Se3.from_se3(x.reshape(6, 1)).log().reshape(6) for x in xis])
    prod = [Se3.from_se3(xis[i].reshape(6, 1)) * Se3.from_se3(xis[i + 1].reshape(6, 1)) for i in range(7)]
    ka["lie_prod_q"] = np.stack([p.so3.quat.reshape(4) for p in prod])
    ka["lie_prod_t"] = np.stack([p.tvec.reshape(3) for p in prod])
    inv = [Se3.from_se3(x.reshape(6, 1)).inverse() for x in xis]
    ka["lie_inv_q"] = np.stack([p.so3.quat.reshape(4) for p in inv]).astype(np.float64)
    ka["lie_inv_t"] = np.stack([p.tvec.reshape(3) for p in inv]).astype(np.float64)
    np.savez_compressed(OUT / "known_answers.npz", **ka)

    # ---------------- full pose estimates on the 9 test pairs ---------------------------------
    for i in range(bgr.shape[0] - 1):
        out, _ = run_pair(cam, 4, bgr[i], depth[i], bgr[i + 1], depth[i + 1], capture_levels=(i in (0, 3, 4)))
        np.savez_compressed(OUT / f"pose_testdata_{i + 1}_{i + 2}.npz", **out)
        print("pair", i + 1, i + 2, out["iters"], out["xi"])

    # variants on pair 1->2: strict guard, t-dist weighter, prior
    set_guard("strict")
    out, _ = run_pair(cam, 4, bgr[0], depth[0], bgr[1], depth[1], capture_levels=True)
    np.savez_compressed(OUT / "pose_testdata_1_2_strict.npz", **out)
    print("strict", out["iters"], out["xi"])
    set_guard("inclusive")
    out, _ = run_pair(cam, 4, bgr[0], depth[0], bgr[1], depth[1], capture_levels=False, use_weighter=True)
    np.savez_compressed(OUT / "pose_testdata_1_2_tdist.npz", **out)
    print("tdist", out["iters"], out["xi"], out["lambdas"][:3])

    # prior (sigma): needs a previous estimate, so run three frames through one estimator
    dvo = get_dvo("robust-dvo", cam, Se3.identity(), levels=4, sigma=1e-9)
    seq = []
    for i in range(3):
        T = dvo.step(bgr[i], depth[i].copy())
        seq.append(np.concatenate([T.so3.quat.reshape(4), T.tvec.reshape(3)]))
    np.savez_compressed(OUT / "pose_testdata_seq3_sigma.npz", qt=np.stack(seq).astype(np.float32), sigma=1e-9)
    print("sigma seq", seq[1], seq[2])

    # ---------------- synthetic pairs (inputs committed, so the GPU box needs no generator match) ---
    for name, (h, w, lv, seeds) in {"syn640": (480, 640, 4, [0, 1]), "syn160": (120, 160, 3, [2, 3, 4]),
                                    "syn101": (77, 101, 3, [5, 6])}.items():
        data = make_pairs_numpy(seeds, height=h, width=w)
        camS = camera(data["K"], data["depth_scale"])
        save = dict(gray_prev=data["bgr_prev"][..., 0], gray_cur=data["bgr_cur"][..., 0],
                    depth_prev=data["depth_prev"], depth_cur=data["depth_cur"], xi_true=data["xi"],
                    K=np.array(data["K"]), depth_scale=data["depth_scale"], levels=lv)
        for j in range(len(seeds)):
            for wname, kw in (("none", {}), ("tdist", {"use_weighter": True})):
                if wname == "tdist" and j > 0:
                    continue
                out, _ = run_pair(camS, lv, data["bgr_prev"][j], data["depth_prev"][j], data["bgr_cur"][j],
                                  data["depth_cur"][j], capture_levels=(wname == "none"), **kw)
                for k, v in out.items():
                    save[f"p{j}_{wname}_{k}"] = v
                print(name, j, wname, out["iters"], out["xi"], "true", data["xi"][j])
        np.savez_compressed(OUT / f"pose_{name}.npz", **save)

    # the reference's 10x10 unit test scene (test_cpu_robust_dense_visual_odometry.py:20-44)
    g = np.full((10, 10), 150, dtype=np.uint8)
    g[:5, :5] = 50
    c = np.repeat(g[..., None], 3, axis=-1)
    d = np.ones((10, 10), dtype=np.uint8)
    d[:5, :5] = 3
    camI = RGBDCameraModel(np.eye(3, dtype=np.float32), 1.0)
    dvo = get_dvo("robust-dvo", camI, Se3.identity(), levels=1)
    dvo.step(color_image=c, depth_image=d)
    dvo._build_pyramids(gray_image=g, depth_image=d)
    dvo._setup(level=0)
    r, J, m = dvo.compute_residuals_and_jacobian(estimate=Se3.identity(), level=0)
    np.savez_compressed(OUT / "unit10x10.npz", r=r, J=J, mask=m)
    print("unit10", r.shape, float(np.abs(r).max()))


if __name__ == "__main__":
    main()
