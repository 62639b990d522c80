"""
Numeric Rot3 / Pose3 value types for the Python `Values` of symforce_b200.opt -- the two Lie groups the device
factor kinds take.  Storage order and group operations follow the reference's generated numeric classes:

  Rot3  storage [qx, qy, qz, qw]                gen/python/sym/rot3.py, gen/python/sym/ops/rot3/{group,lie_group}_ops.py
  Pose3 storage [qx, qy, qz, qw, tx, ty, tz]    gen/python/sym/pose3.py, gen/python/sym/ops/pose3/{group,lie_group}_ops.py
        tangent [rotation (3), translation (3)]; retract is decoupled: q * exp(w), t + v (pose3 lie_group_ops.py:79-116)

Host-side numpy only (building problems, reading results); nothing here is on the optimization path.
"""
import math

import numpy as np

K_DEFAULT_EPSILON = 10 * np.finfo(np.float64).eps  # sym::kDefaultEpsilon<double>, gen/cpp/sym/util/epsilon.h:31


def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ])


class Rot3:
    """Unit quaternion [x, y, z, w]."""

    STORAGE_DIM = 4
    TANGENT_DIM = 3

    def __init__(self, q=None):
        if q is None:
            q = [0.0, 0.0, 0.0, 1.0]
        q = np.asarray(q, dtype=np.float64).reshape(-1)
        if q.shape != (4,):
            raise IndexError(f"Rot3 expects 4 storage elements, got shape {q.shape}")
        self.data = q.copy()

    # -- storage -------------------------------------------------------------------------------------------------
    @classmethod
    def from_storage(cls, vec):
        return cls(vec)

    def to_storage(self):
        return [float(x) for x in self.data]

    @classmethod
    def identity(cls):
        return cls()

    # -- constructors --------------------------------------------------------------------------------------------
    @classmethod
    def from_yaw_pitch_roll(cls, yaw=0.0, pitch=0.0, roll=0.0):
        """R = Rz(yaw) Ry(pitch) Rx(roll)  (symforce/geo/rot3.py: from_yaw_pitch_roll)."""
        cy, sy = math.cos(0.5 * yaw), math.sin(0.5 * yaw)
        cp, sp = math.cos(0.5 * pitch), math.sin(0.5 * pitch)
        cr, sr = math.cos(0.5 * roll), math.sin(0.5 * roll)
        qz = np.array([0.0, 0.0, sy, cy])
        qy = np.array([0.0, sp, 0.0, cp])
        qx = np.array([sr, 0.0, 0.0, cr])
        return cls(_quat_mul(_quat_mul(qz, qy), qx))

    @classmethod
    def from_tangent(cls, vec, epsilon=K_DEFAULT_EPSILON):
        """exp map, rot3 lie_group_ops.py: from_tangent (theta = sqrt(eps^2 + |v|^2))."""
        v = np.asarray(vec, dtype=np.float64).reshape(-1)
        th = math.sqrt(epsilon ** 2 + float(v @ v))
        s = math.sin(0.5 * th) / th
        return cls([s * v[0], s * v[1], s * v[2], math.cos(0.5 * th)])

    @classmethod
    def from_rotation_matrix(cls, R):
        R = np.asarray(R, dtype=np.float64)
        tr = R[0, 0] + R[1, 1] + R[2, 2]
        if tr > 0:
            s = 2.0 * math.sqrt(1.0 + tr)
            q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
        else:
            i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
            j, k = (i + 1) % 3, (i + 2) % 3
            s = 2.0 * math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k])
            q = [0.0] * 4
            q[i] = 0.25 * s
            q[j] = (R[j, i] + R[i, j]) / s
            q[k] = (R[k, i] + R[i, k]) / s
            q[3] = (R[k, j] - R[j, k]) / s
        q = np.array(q)
        return cls(q / np.linalg.norm(q))

    # -- group ---------------------------------------------------------------------------------------------------
    def inverse(self):
        x, y, z, w = self.data
        return Rot3([-x, -y, -z, w])

    def compose(self, other):
        return Rot3(_quat_mul(self.data, other.data))

    def between(self, other):
        return self.inverse().compose(other)

    def to_rotation_matrix(self):
        x, y, z, w = self.data
        return np.array([
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ])

    def __mul__(self, other):
        if isinstance(other, Rot3):
            return self.compose(other)
        v = np.asarray(other, dtype=np.float64)
        return (self.to_rotation_matrix() @ v.reshape(3, -1)).reshape(v.shape)

    # -- Lie group -----------------------------------------------------------------------------------------------
    def to_tangent(self, epsilon=K_DEFAULT_EPSILON):
        """log map, rot3 lie_group_ops.py: to_tangent (w clamped to [-1+eps, 1-eps], sign-fixed)."""
        x, y, z, w = self.data
        sign = 1.0 if w >= 0 else -1.0  # 2*min(0, sign(w)) + 1 in the generated code
        wc = min(abs(w), 1.0 - epsilon)
        s = 2.0 * sign * math.acos(wc) / math.sqrt(1.0 - wc * wc)
        return np.array([s * x, s * y, s * z])

    def retract(self, vec, epsilon=K_DEFAULT_EPSILON):
        return self.compose(Rot3.from_tangent(vec, epsilon))

    def local_coordinates(self, other, epsilon=K_DEFAULT_EPSILON):
        return self.between(other).to_tangent(epsilon)

    def __repr__(self):
        return "<Rot3 [%s]>" % ", ".join("%.8g" % x for x in self.data)


class Pose3:
    """Rotation + translation, storage [q (4), t (3)]."""

    STORAGE_DIM = 7
    TANGENT_DIM = 6

    def __init__(self, R=None, t=None):
        self.R = R if R is not None else Rot3()
        self.t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64).reshape(3).copy()

    @property
    def data(self):
        return np.concatenate([self.R.data, self.t])

    @classmethod
    def from_storage(cls, vec):
        v = np.asarray(vec, dtype=np.float64).reshape(-1)
        if v.shape != (7,):
            raise IndexError(f"Pose3 expects 7 storage elements, got shape {v.shape}")
        return cls(Rot3(v[:4]), v[4:])

    def to_storage(self):
        return [float(x) for x in self.data]

    @classmethod
    def identity(cls):
        return cls()

    def rotation(self):
        return self.R

    def position(self):
        return self.t.copy()

    @classmethod
    def from_tangent(cls, vec, epsilon=K_DEFAULT_EPSILON):
        v = np.asarray(vec, dtype=np.float64).reshape(-1)
        return cls(Rot3.from_tangent(v[:3], epsilon), v[3:6])

    def to_tangent(self, epsilon=K_DEFAULT_EPSILON):
        return np.concatenate([self.R.to_tangent(epsilon), self.t])

    def inverse(self):
        Ri = self.R.inverse()
        return Pose3(Ri, -(Ri * self.t))

    def compose(self, other):
        return Pose3(self.R.compose(other.R), self.R * other.t + self.t)

    def between(self, other):
        return self.inverse().compose(other)

    def __mul__(self, other):
        if isinstance(other, Pose3):
            return self.compose(other)
        v = np.asarray(other, dtype=np.float64)
        return (self.R * v.reshape(3)) + self.t

    def retract(self, vec, epsilon=K_DEFAULT_EPSILON):
        v = np.asarray(vec, dtype=np.float64).reshape(-1)
        return Pose3(self.R.retract(v[:3], epsilon), self.t + v[3:6])

    def local_coordinates(self, other, epsilon=K_DEFAULT_EPSILON):
        return np.concatenate([self.R.local_coordinates(other.R, epsilon), other.t - self.t])

    def __repr__(self):
        return "<Pose3 R=%r t=[%s]>" % (self.R, ", ".join("%.8g" % x for x in self.t))
