"""Golden vectors of the reference's `approximate_image2_gradient=True` mode (cpu_...py:60-77, :160-165, :187-188),
produced by the UNMODIFIED reference sources like make_golden.py (same three external shims).

    python tests/golden/make_golden_approx.py     # writes tests/golden/pose_*_approx.npz
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden as G  # noqa: E402  (sets up the shims and the reference imports)


def main():
    G.set_guard("inclusive")
    frames = np.load(G.OUT / "frames_testdata.npz")
    bgr, depth = frames["bgr"], frames["depth"]
    cam = G.camera(tuple(frames["K"]), float(frames["depth_scale"]))
    for i in (0, 3):
        out, _ = G.run_pair(cam, 4, bgr[i], depth[i], bgr[i + 1], depth[i + 1], capture_levels=True,
                            approximate_image2_gradient=True)
        np.savez_compressed(G.OUT / f"pose_testdata_{i + 1}_{i + 2}_approx.npz", **out)
        print("approx pair", i + 1, i + 2, out["iters"], out["xi"])
    out, _ = G.run_pair(cam, 4, bgr[0], depth[0], bgr[1], depth[1], capture_levels=False,
                        approximate_image2_gradient=True, use_weighter=True)
    np.savez_compressed(G.OUT / "pose_testdata_1_2_approx_tdist.npz", **out)
    print("approx tdist", out["iters"], out["xi"])
    syn = np.load(G.OUT / "pose_syn160.npz")
    camS = G.camera(tuple(syn["K"]), float(syn["depth_scale"]))
    rep = lambda a: np.ascontiguousarray(np.repeat(a[..., None], 3, axis=-1))  # noqa: E731
    save = {}
    for j in range(syn["gray_prev"].shape[0]):
        out, _ = G.run_pair(camS, int(syn["levels"]), rep(syn["gray_prev"][j]), syn["depth_prev"][j],
                            rep(syn["gray_cur"][j]), syn["depth_cur"][j], capture_levels=False,
                            approximate_image2_gradient=True)
        for k in ("q", "t", "xi", "iters", "err_last"):
            save[f"p{j}_{k}"] = out[k]
        print("approx syn160", j, out["iters"], out["xi"])
    np.savez_compressed(G.OUT / "pose_syn160_approx.npz", **save)


if __name__ == "__main__":
    main()
