"""SO(3)/SE(3) value types with the reference's interface (utils/lie_algebra/ in the reference):
rotation stored as a wxyz quaternion that is never renormalised, translation as a (3,1) float32
column, twists ordered [v; w].  Host-side only (pose chaining, API types); the per-iteration SE(3)
updates of the estimator run on the GPU (csrc/se3_device.cuh)."""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-6
_F = np.float32


def _wrap(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def _hat(v):
    x, y, z = (float(c) for c in np.asarray(v).reshape(3))
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=_F)


class So3:
    """Rotation.  Accepts a (4,1) wxyz quaternion, a (3,1) rotation vector or a (3,3) matrix."""

    def __init__(self, rot: np.ndarray, quat_repr: str = "wxyz"):
        if not isinstance(rot, np.ndarray):
            raise AssertionError(f"Expected 'rot' to be a numpy array, got {type(rot)} instead")
        if quat_repr != "wxyz":
            raise AssertionError("only the 'wxyz' quaternion layout is supported")
        self.quat_repr = quat_repr
        self._phi = None
        if rot.shape == (4, 1):
            self._q = rot.copy()
        elif rot.shape == (3, 1):
            self._q, self._phi = self._from_rotvec(rot)
        elif rot.shape == (3, 3):
            if not (np.allclose(rot @ rot.T, np.eye(3), atol=1e-3) and abs(np.linalg.det(rot) - 1) < 1e-3):
                raise AssertionError(f"Got invalid rotation matrix '{rot.tolist()}'")
            self._q = self._from_matrix(rot)
        else:
            raise ValueError(f"Expected 'rot' to have shape (4, 1), (3, 1) or (3, 3), got '{rot.shape}' instead")

    @staticmethod
    def _from_rotvec(phi):
        theta = np.linalg.norm(phi)
        if theta < EPS:
            return np.array([[1], [0], [0], [0]], dtype=_F), np.zeros((3, 1), dtype=_F)
        axis = phi / theta
        phi_w = _wrap(theta) * axis
        th = np.linalg.norm(phi_w)
        ax = (phi_w / th).reshape(3)
        s, c = math.sin(th / 2), math.cos(th / 2)
        return np.array([c, s * ax[0], s * ax[1], s * ax[2]], dtype=_F).reshape(4, 1), phi_w

    @staticmethod
    def _from_matrix(R):
        q = np.zeros(4)
        tr = R[0, 0] + R[1, 1] + R[2, 2]
        if tr > 0:
            t = math.sqrt(1 + tr)
            q[0] = 0.5 * t
            t = 0.5 / t
            q[1:] = [(R[2, 1] - R[1, 2]) * t, (R[0, 2] - R[2, 0]) * t, (R[1, 0] - R[0, 1]) * t]
        else:
            i = 1 if R[1, 1] > R[0, 0] else 0
            if R[2, 2] > R[i, i]:
                i = 2
            j, k = (i + 1) % 3, (i + 2) % 3
            t = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1)
            q[1 + i] = 0.5 * t
            t = 0.5 / t
            q[0] = (R[k, j] - R[j, k]) * t
            q[1 + j] = (R[j, i] + R[i, j]) * t
            q[1 + k] = (R[k, i] + R[i, k]) * t
        return q.reshape(4, 1)

    @property
    def quat(self):
        return self._q

    def exp(self) -> np.ndarray:
        """3x3 float32 rotation matrix straight from the stored quaternion."""
        w, x, y, z = self._q.flatten()
        return np.array([
            [2 * (w * w + x * x) - 1, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 2 * (w * w + y * y) - 1, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 2 * (w * w + z * z) - 1]], dtype=_F)

    def log(self) -> np.ndarray:
        if self._phi is None:
            w = self._q[0, 0]
            v = self._q[1:, 0]
            n = np.linalg.norm(v)
            if n < EPS:
                self._phi = np.zeros((3, 1), dtype=_F)
            else:
                self._phi = _wrap(2 * math.atan2(n, w) / n) * v.reshape(3, 1)
        return self._phi

    def hat(self) -> np.ndarray:
        return _hat(self.log())

    @property
    def theta(self):
        return np.linalg.norm(self.log())

    def inverse(self) -> "So3":
        return So3(self.exp().T.copy())

    def copy(self) -> "So3":
        return So3(self._q.copy())

    @classmethod
    def identity(cls) -> "So3":
        return cls(np.zeros((3, 1), dtype=_F))

    def __mul__(self, right: "So3") -> "So3":
        if not isinstance(right, So3):
            raise AssertionError(f"Got invalid type '{type(right)}', expected So3")
        a, b = self._q, right.quat
        w = a[0] * b[0] - np.dot(a[1:].T, b[1:]).flatten()
        v = a[0, 0] * b[1:, 0] + b[0, 0] * a[1:, 0] + np.cross(a[1:, 0], b[1:, 0])
        return So3(np.concatenate((w, v)).astype(_F).reshape(4, 1))

    def __eq__(self, other) -> bool:
        return bool(np.allclose(self.log(), other.log(), atol=EPS))


class Se3:
    """Rigid transform: So3 + (3,1) translation."""

    def __init__(self, so3: So3, tvec: np.ndarray):
        if not hasattr(so3, "quat"):
            raise AssertionError(f"Expected 'so3' to be of type 'So3', got '{type(so3)}' instead")
        if tvec.shape != (3, 1):
            raise AssertionError(f"Expected 'tvec' to have shape '(3, 1)', got '{tvec.shape}' instead")
        self._so3 = so3
        self._t = tvec.copy()

    @property
    def so3(self):
        return self._so3

    @property
    def tvec(self):
        return self._t

    def exp(self) -> np.ndarray:
        """4x4 float32 matrix; the rotation block is the identity for rotation vectors below 1e-6."""
        T = np.eye(4, dtype=_F)
        if abs(np.linalg.norm(self._so3.log())) >= EPS:
            T[:3, :3] = self._so3.exp()
        T[:3, 3] = self._t.flatten()
        return T

    def log(self) -> np.ndarray:
        xi = np.zeros((6, 1), dtype=_F)
        phi = self._so3.log()
        theta = np.linalg.norm(phi)
        if abs(theta) < EPS:
            xi[:3, 0] = self._t.flatten()
            return xi
        a = So3(phi.reshape(3, 1) / theta)
        al = a.log()
        h2 = theta / 2
        A = h2 * np.cos(h2) / np.sin(h2)
        V_inv = A * np.eye(3, dtype=_F) + (1 - A) * np.dot(al, al.T) - h2 * a.hat()
        xi[:3, 0] = np.dot(V_inv, self._t).flatten()
        xi[3:, 0] = phi.flatten()
        return xi

    def inverse(self) -> "Se3":
        inv = self._so3.inverse()
        return Se3(inv, -np.dot(inv.exp(), self._t))

    def copy(self) -> "Se3":
        return Se3(self._so3.copy(), self._t.copy())

    @classmethod
    def identity(cls) -> "Se3":
        return cls(So3.identity(), np.zeros((3, 1), dtype=_F))

    @classmethod
    def from_se3(cls, xi: np.ndarray) -> "Se3":
        if xi.shape != (6, 1):
            raise AssertionError(f"Expected 'xi' to have shape '(6, 1)', got '{xi.shape}' instead")
        v = xi[:3]
        so3 = So3(xi[3:])
        theta = so3.theta
        if theta < EPS:
            return cls(So3.identity(), v)
        K = so3.hat()
        V = (np.eye(3, dtype=_F) + ((1 - math.cos(theta)) / theta ** 2) * K +
             ((theta - math.sin(theta)) / theta ** 3) * np.dot(K, K))
        return cls(So3(xi[3:]), np.dot(V, v))

    @classmethod
    def from_qt(cls, qt) -> "Se3":
        """From the C ABI's 7-float layout [qw qx qy qz tx ty tz]."""
        qt = np.asarray(qt, dtype=_F).reshape(7)
        return cls(So3(qt[:4].reshape(4, 1).copy()), qt[4:].reshape(3, 1).copy())

    def __mul__(self, right: "Se3") -> "Se3":
        if not hasattr(right, "tvec"):
            raise AssertionError(f"Expected 'right' to be of type 'Se3', got '{type(right)}' instead.")
        return Se3(self._so3 * right.so3, self._t + np.dot(self._so3.exp(), right.tvec))

    def __eq__(self, other) -> bool:
        return bool(np.allclose(self.log(), other.log(), atol=EPS))


def pose_to_qt(pose) -> np.ndarray:
    """Any Se3-like object (this module's or the reference's) -> 7 float32 [q wxyz, t]."""
    q = np.asarray(pose.so3.quat, dtype=np.float64).reshape(4)
    if getattr(pose.so3, "quat_repr", "wxyz") == "xyzw":
        q = np.roll(q, 1)
    return np.concatenate([q, np.asarray(pose.tvec, dtype=np.float64).reshape(3)]).astype(_F)
