"""Restatement of ase.quaternions.Quaternion (ase 3.22.1 public behaviour).

Convention: q = [w, x, y, z]; `rotate(v)` returns R(q)·v; `a * b` is the Hamilton
product; `from_euler_angles(a, b, c, mode)` = q_z(c) · q_{y|x}(b) · q_z(a).
"""
import numpy as np


class Quaternion:
    def __init__(self, qin=(1, 0, 0, 0)):
        assert len(qin) == 4
        self.q = np.array(qin)

    def __str__(self):
        return self.q.__str__()

    def __mul__(self, other):
        sw, sx, sy, sz = self.q
        ow, ox, oy, oz = other.q
        return Quaternion(
            [
                sw * ow - sx * ox - sy * oy - sz * oz,
                sw * ox + sx * ow + sy * oz - sz * oy,
                sw * oy + sy * ow + sz * ox - sx * oz,
                sw * oz + sz * ow + sx * oy - sy * ox,
            ]
        )

    def conjugate(self):
        return Quaternion(self.q * np.array([1.0, -1.0, -1.0, -1.0]))

    def rotation_matrix(self):
        w, x, y, z = self.q
        ww, xx, yy, zz = w * w, x * x, y * y, z * z
        wx, wy, wz = w * x, w * y, w * z
        xy, xz, yz = x * y, x * z, y * z
        return np.array(
            [
                [ww + xx - yy - zz, 2 * (xy - wz), 2 * (xz + wy)],
                [2 * (xy + wz), ww - xx + yy - zz, 2 * (yz - wx)],
                [2 * (xz - wy), 2 * (yz + wx), ww - xx - yy + zz],
            ]
        )

    def rotate(self, vector):
        return np.dot(self.rotation_matrix(), np.array(vector))

    def axis_angle(self):
        sinth2 = np.linalg.norm(self.q[1:])
        if sinth2 == 0:
            return np.array([0, 0, 1.0]), 0.0
        theta = np.arctan2(sinth2, self.q[0]) * 2
        return self.q[1:] / sinth2, theta

    def arc_distance(self, other):
        return np.arccos(np.clip(np.dot(self.q, other.q), -1, 1))

    @staticmethod
    def rotate_byq(q, vector):
        return Quaternion(q).rotate(vector)

    @staticmethod
    def from_axis_angle(n, theta):
        n = np.array(n, float) / np.linalg.norm(n)
        return Quaternion(np.concatenate([[np.cos(theta / 2.0)], np.sin(theta / 2.0) * n]))

    @staticmethod
    def from_euler_angles(a, b, c, mode="zyz"):
        q_a = Quaternion.from_axis_angle([0, 0, 1], a)
        q_c = Quaternion.from_axis_angle([0, 0, 1], c)
        if mode == "zyz":
            q_b = Quaternion.from_axis_angle([0, 1, 0], b)
        elif mode == "zxz":
            q_b = Quaternion.from_axis_angle([1, 0, 0], b)
        else:
            raise ValueError("Invalid Euler angles mode {0}".format(mode))
        return q_c * q_b * q_a
