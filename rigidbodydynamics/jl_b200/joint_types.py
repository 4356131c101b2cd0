"""The eight joint types of the reference (src/joint_types/*.jl) as host-side descriptors.

Only what the flattener and the state randomiser need lives here: nq/nv, the constant
parameters (axes), the C-ABI type code, and ``rand_configuration`` / ``zero_configuration``.
The joint kinematics themselves (joint_transform, motion_subspace, joint_twist, q̇ = N(q) v) are
device code in ``csrc/rbd_joints.cuh`` and, independently, CPU code in ``oracle/``.

Every type has a constant motion subspace in the frame after the joint and zero joint bias
acceleration (e.g. revolute.jl:70-74, quaternion_floating.jl:98-102, planar.jl:102-106), which
is what makes the one-hot-subspace canonicalisation in ``csrc/rbd_model.cpp`` possible.
"""
from __future__ import annotations

import numpy as np

from .spatial import random_unit_quaternion, quat_to_mrp

# C-ABI joint type codes: must match include/rbd_b200.h (RBD_JOINT_*).
JOINT_REVOLUTE = 0
JOINT_PRISMATIC = 1
JOINT_FIXED = 2
JOINT_PLANAR = 3
JOINT_QUATERNION_FLOATING = 4
JOINT_SPQUAT_FLOATING = 5
JOINT_QUATERNION_SPHERICAL = 6
JOINT_SINCOS_REVOLUTE = 7


def _unit(a):
    a = np.asarray(a, float).reshape(3)
    n = np.linalg.norm(a)
    if n == 0:
        raise ValueError("joint axis must be non-zero")
    return a / n


class JointType:
    code = -1
    nq = 0
    nv = 0
    isfloating = False

    def params9(self):
        """9 doubles of per-type constants for rbd_model_desc.jparam."""
        return np.zeros(9)

    def rand_configuration(self, rng):
        return np.zeros(self.nq)

    def zero_configuration(self):
        return np.zeros(self.nq)

    def flip_direction(self):
        return self

    def __repr__(self):
        return type(self).__name__


class Revolute(JointType):
    """revolute.jl: q = angle, v = rate; axis normalised at construction (:14-17)."""
    code, nq, nv = JOINT_REVOLUTE, 1, 1

    def __init__(self, axis):
        self.axis = _unit(axis)

    def params9(self):
        p = np.zeros(9)
        p[:3] = self.axis
        return p

    def rand_configuration(self, rng):                       # revolute.jl:54-57
        return rng.standard_normal(1)

    def flip_direction(self):
        return type(self)(-self.axis)

    @staticmethod
    def rand(rng):
        return Revolute(_unit(rng.standard_normal(3)))

    def __repr__(self):
        return f"{type(self).__name__}(axis={self.axis.tolist()})"


class Prismatic(Revolute):
    """prismatic.jl: q = displacement along axis."""
    code = JOINT_PRISMATIC

    @staticmethod
    def rand(rng):
        return Prismatic(_unit(rng.standard_normal(3)))


class SinCosRevolute(Revolute):
    """sin_cos_revolute.jl: q = [sin θ, cos θ], v = θ̇."""
    code, nq, nv = JOINT_SINCOS_REVOLUTE, 2, 1

    def rand_configuration(self, rng):                       # sin_cos_revolute.jl:55-58
        q = rng.standard_normal(2)
        return q / np.linalg.norm(q)

    def zero_configuration(self):
        return np.array([0.0, 1.0])

    @staticmethod
    def rand(rng):
        return SinCosRevolute(_unit(rng.standard_normal(3)))


class Fixed(JointType):
    """fixed.jl: no motion; nq = nv = 0."""
    code = JOINT_FIXED

    @staticmethod
    def rand(rng):
        return Fixed()


class Planar(JointType):
    """planar.jl: q = [x, y, θ]; v = [ẋ_body, ẏ_body, θ̇] (linear part in the AFTER frame, :7-21)."""
    code, nq, nv = JOINT_PLANAR, 3, 3

    def __init__(self, x_axis, y_axis):
        self.x_axis = _unit(x_axis)
        self.y_axis = _unit(y_axis)
        if abs(self.x_axis @ self.y_axis) > 100 * np.finfo(float).eps:
            raise ValueError("Planar: x and y axes must be orthogonal")      # planar.jl:36
        self.rot_axis = np.cross(self.x_axis, self.y_axis)

    def params9(self):
        return np.concatenate([self.x_axis, self.y_axis, self.rot_axis])

    def rand_configuration(self, rng):                       # planar.jl:57-63
        return np.array([rng.random() - 0.5, rng.random() - 0.5, rng.standard_normal()])

    @staticmethod
    def rand(rng):                                           # planar.jl:45-50
        x = _unit(rng.standard_normal(3))
        y = _unit(rng.standard_normal(3))
        y = _unit(y - (x @ y) * x)
        return Planar(x, y)

    def __repr__(self):
        return f"Planar(x={self.x_axis.tolist()}, y={self.y_axis.tolist()})"


class QuaternionFloating(JointType):
    """quaternion_floating.jl: q = [w x y z, p], v = [ω; v_lin] in the body frame (:9-17)."""
    code, nq, nv = JOINT_QUATERNION_FLOATING, 7, 6
    isfloating = True

    def rand_configuration(self, rng):                       # quaternion_floating.jl:175-180
        return np.concatenate([random_unit_quaternion(rng), rng.random(3) - 0.5])

    def zero_configuration(self):
        return np.array([1.0, 0, 0, 0, 0, 0, 0])

    @staticmethod
    def rand(rng):
        return QuaternionFloating()


class SPQuatFloating(JointType):
    """spquat_floating.jl: q = [MRP (3), p], v = body twist."""
    code, nq, nv = JOINT_SPQUAT_FLOATING, 6, 6
    isfloating = True

    def rand_configuration(self, rng):                       # spquat_floating.jl:178-183
        q = random_unit_quaternion(rng)
        if q[0] < 0:            # principal value: keeps |MRP| <= 1
            q = -q
        return np.concatenate([quat_to_mrp(q), rng.random(3) - 0.5])

    @staticmethod
    def rand(rng):
        return SPQuatFloating()


class QuaternionSpherical(JointType):
    """quaternion_spherical.jl: q = unit quaternion [w x y z], v = body angular velocity."""
    code, nq, nv = JOINT_QUATERNION_SPHERICAL, 4, 3

    def rand_configuration(self, rng):                       # quaternion_spherical.jl:110-114
        return random_unit_quaternion(rng)

    def zero_configuration(self):
        return np.array([1.0, 0, 0, 0])

    @staticmethod
    def rand(rng):
        return QuaternionSpherical()


JOINT_TYPE_BY_CODE = {
    c.code: c for c in (Revolute, Prismatic, Fixed, Planar, QuaternionFloating, SPQuatFloating,
                        QuaternionSpherical, SinCosRevolute)
}
