"""Kinematics by-products of the hot path's outward sweep, batched.  Same names and meaning as the
reference, every quantity in the mechanism's root frame, 6-vectors as [angular; linear]:

    transform_to_root               -> transforms_to_root(_)             src/mechanism_state.jl:687-714
    center_of_mass                  -> center_of_mass                    src/mechanism_algorithms.jl:30-49
    kinetic_energy                  -> kinetic_energy                    src/mechanism_state.jl:886-888, 989-994
    gravitational_potential_energy  -> gravitational_potential_energy    src/mechanism_state.jl:897-903, 996-1000
    momentum / momentum_rate_bias   -> momentum / momentum_rate_bias     src/mechanism_state.jl:878-884, 975-987
    momentum_matrix!                -> momentum_matrix(_)                src/mechanism_algorithms.jl:313-327
    path + geometric_jacobian!      -> path, geometric_jacobian(_)       src/graphs/tree_path.jl, mechanism_algorithms.jl:80-100

All are one call of ``rbd_kinematics`` (include/rbd_b200.h) on the current CUDA stream; ``kinematics_`` exposes the fused
form (any subset of outputs from a single launch).  There is no CPU path.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _cabi
from .algorithms import DimensionMismatch, _stream
from .mechanism import Mechanism, RigidBody
from .state import MechanismState, _DT

__all__ = ["TreePath", "path", "kinematics_", "transforms_to_root", "transforms_to_root_", "center_of_mass", "kinetic_energy",
           "gravitational_potential_energy", "momentum", "momentum_rate_bias", "momentum_matrix", "momentum_matrix_",
           "geometric_jacobian", "geometric_jacobian_"]

_ROWS = {"transforms_to_root": lambda s: 12 * len(s.mechanism.joints), "center_of_mass": lambda s: 3,
         "kinetic_energy": lambda s: 1, "gravitational_potential_energy": lambda s: 1, "momentum": lambda s: 6,
         "momentum_rate_bias": lambda s: 6, "momentum_matrix": lambda s: 6 * s.nv, "geometric_jacobian": lambda s: 6 * s.nv}
_NEEDS_V = ("kinetic_energy", "momentum", "momentum_rate_bias")


@dataclass
class TreePath:
    """``TreePath`` (src/graphs/tree_path.jl): the joints between ``source`` and ``target`` with their traversal
    directions, stored as one sign per tree joint: -1 = up (towards the root, source side), +1 = down, 0 = not on the path."""
    source: RigidBody
    target: RigidBody
    sign: np.ndarray            # int8 [number of tree joints]


def path(mechanism: Mechanism, source: RigidBody, target: RigidBody) -> TreePath:
    """``path(mechanism, from, to)`` (src/mechanism.jl:146-151 -> graphs/tree_path.jl:60-95): up from ``source`` to the
    lowest common ancestor, then down to ``target``."""
    index = {id(j.successor): i for i, j in enumerate(mechanism.joints)}

    def ancestors(body):                # joints from `body` up to the root
        out = []
        while body is not mechanism.root_body:
            i = index[id(body)]
            out.append(i)
            body = mechanism.joints[i].predecessor
        return out

    up, down = ancestors(source), ancestors(target)
    while up and down and up[-1] == down[-1]:      # drop the common part above the lowest common ancestor
        up.pop()
        down.pop()
    sign = np.zeros(len(mechanism.joints), np.int8)
    sign[up] = -1
    sign[down] = 1
    return TreePath(source, target, sign)


def kinematics_(state: MechanismState, path_: Optional[TreePath] = None, **outs: Optional[torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Fused form: fill any subset of {transforms_to_root, center_of_mass, kinetic_energy, gravitational_potential_energy,
    momentum, momentum_rate_bias, momentum_matrix, geometric_jacobian} ([rows, B] tensors of the state's dtype) with one
    kernel launch."""
    state.check_modcount()
    lib = _cabi.load_library()
    ko = _cabi.RbdKinematicsOut()
    for name, t in outs.items():
        if name not in _ROWS:
            raise TypeError(f"unknown kinematics output {name!r}")
        if t is None:
            continue
        rows = _ROWS[name](state)
        if t.dtype != state.dtype or t.device != state.q.device:
            raise TypeError(f"{name}: dtype/device must match the state")
        if t.dim() != 2 or t.shape[0] != rows or t.shape[1] != state.batch:
            raise DimensionMismatch(f"{name} has wrong size: expected ({rows}, {state.batch}), got {tuple(t.shape)}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be [rows, B] contiguous (batch index fastest)")
        setattr(ko, name, t.data_ptr())
    want_jac = outs.get("geometric_jacobian") is not None
    if want_jac and path_ is None:
        raise ValueError("geometric_jacobian needs a path")
    sign = None
    if want_jac:
        sign = np.ascontiguousarray(path_.sign, np.int8)
        if sign.shape != (len(state.mechanism.joints),):
            raise DimensionMismatch("path does not belong to this mechanism")
    _cabi.check(lib.rbd_kinematics(state.handle.ptr, _DT[state.dtype], state.batch, state.batch, state.q.data_ptr(),
                                   state.v.data_ptr(), None if sign is None else sign.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.byref(ko), _stream()))
    return {k: t for k, t in outs.items() if t is not None}


def _alloc(state: MechanismState, name: str) -> torch.Tensor:
    return torch.empty((_ROWS[name](state), state.batch), dtype=state.dtype, device=state.q.device)


def _one(state: MechanismState, name: str, out: Optional[torch.Tensor] = None, path_: Optional[TreePath] = None):
    out = _alloc(state, name) if out is None else out
    kinematics_(state, path_, **{name: out})
    return out


def transforms_to_root_(out: torch.Tensor, state: MechanismState):
    """``transform_to_root(state, body)`` of every non-root body: rows 12 i .. 12 i + 11 = rotation (row-major 9) and
    translation (3) of the successor of tree joint i."""
    return _one(state, "transforms_to_root", out)


def transforms_to_root(state: MechanismState):
    return _one(state, "transforms_to_root")


def center_of_mass(state: MechanismState):
    return _one(state, "center_of_mass")


def kinetic_energy(state: MechanismState):
    return _one(state, "kinetic_energy")[0]


def gravitational_potential_energy(state: MechanismState):
    return _one(state, "gravitational_potential_energy")[0]


def momentum(state: MechanismState):
    return _one(state, "momentum")


def momentum_rate_bias(state: MechanismState):
    return _one(state, "momentum_rate_bias")


def momentum_matrix_(out: torch.Tensor, state: MechanismState):
    """``momentum_matrix!(A, state)``: [6 nv, B], column k of A at rows 6 k .. 6 k + 5."""
    return _one(state, "momentum_matrix", out)


def momentum_matrix(state: MechanismState):
    return _one(state, "momentum_matrix")


def geometric_jacobian_(out: torch.Tensor, state: MechanismState, path_: TreePath):
    """``geometric_jacobian!(J, state, path)`` in the root frame: [6 nv, B], column k at rows 6 k .. 6 k + 5."""
    return _one(state, "geometric_jacobian", out, path_)


def geometric_jacobian(state: MechanismState, path_: TreePath):
    return _one(state, "geometric_jacobian", None, path_)
