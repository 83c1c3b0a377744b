"""Pinhole RGB-D camera model with the reference's interface (camera_model.py in the reference).
The per-level intrinsics the kernels use are derived inside the library (dvo_set_intrinsics); this class
carries K and the depth scale across the API and offers the same helpers for host-side callers."""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import numpy as np


class RGBDCameraModel:
    INTRINSICS_KEYWORD = "intrinsics"
    DEPTH_SCALE_KEYWORD = "depth_scale"

    def __init__(self, intrinsics: np.ndarray, depth_scale: float, distorssion_coeffs=None, distorssion_model=None):
        intrinsics = np.asarray(intrinsics)
        if intrinsics.shape != (3, 3):
            raise AssertionError(f"Expected a 3x3 'intrinsics', got {intrinsics.shape} instead")
        if not depth_scale >= 0.0:
            raise AssertionError("Expected 'scale' to be a positive floating point, got '{:.3f}' instead".format(
                depth_scale))
        self._intrinsics = np.zeros((3, 4), dtype=np.float32)
        self._intrinsics[:3, :3] = intrinsics
        self.depth_scale = depth_scale
        self.distorssion_coeffs = distorssion_coeffs
        self.distorsion_model = distorssion_model

    @property
    def intrinsics(self) -> np.ndarray:
        return self._intrinsics

    def at(self, level: int) -> np.ndarray:
        """3x4 float32 K of pyramid level `level` (half-pixel-centred decimation)."""
        if level < 0:
            raise AssertionError(f"Expected 'level' to be >= 0, got '{level}' instead")
        if level == 0:
            return self._intrinsics
        s = 2.0 ** (-level)
        o = 2.0 ** (-level - 1) - 0.5
        S = np.array([[s, 0, o], [0, s, o], [0, 0, 1]], dtype=np.float32)
        out = np.zeros((3, 4), dtype=np.float32)
        out[:3, :3] = np.dot(S, self._intrinsics[:3, :3])
        return out

    @classmethod
    def load_from_yaml(cls, filepath: Path) -> Optional["RGBDCameraModel"]:
        import yaml
        filepath = Path(filepath)
        if not filepath.exists():
            return None
        with filepath.open("r") as fp:
            data = yaml.load(fp, yaml.Loader)
        try:
            K = np.array(data[cls.INTRINSICS_KEYWORD], dtype=np.float32)
            scale = data[cls.DEPTH_SCALE_KEYWORD]
        except KeyError:
            return None
        return cls(K, scale, data.get("distorssion_coefficients"), data.get("distorssion_model"))

    def deproject(self, depth_image: np.ndarray, return_mask: bool = False, level: int = 0):
        """Host helper (4xN float32 homogeneous points of the non-zero depth pixels)."""
        h, w = depth_image.shape
        mask = (depth_image != 0).reshape(-1)
        z = (depth_image.reshape(-1) * self.depth_scale)[mask].astype(np.float32)
        xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
        pix = np.vstack((xs.reshape(-1)[mask], ys.reshape(-1)[mask], np.ones_like(z)))
        rays = np.dot(np.linalg.inv(self.at(level)[:3, :3]), pix)
        cloud = np.vstack((rays[0] * z, rays[1] * z, z, np.ones_like(z)))
        return (cloud, mask.reshape(h, w)) if return_mask else cloud

    def project(self, pointcloud: np.ndarray, level: int = 0) -> np.ndarray:
        uv = np.dot(self.at(level), pointcloud)
        uv /= uv[2, :]
        return uv
