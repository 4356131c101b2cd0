"""The C-ABI library loads on a machine without a GPU, exports every symbol include/rbd_b200.h declares, and its host-side
logic (model flattening, status codes) behaves like the reference's error paths.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

import rigidbodydynamics.jl_b200 as rbd
from rigidbodydynamics.jl_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "rbd_b200.h")).read()
    declared = set(re.findall(r"\b(rbd_[a-z_]+)\s*\(", header))
    declared -= {"rbd_status", "rbd_dtype"}
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(_cabi.SYMBOLS) == declared


def test_model_info_matches_reference_layout(built):
    mech = rbd.load_model("atlas", floating=True)
    d = mech.flatten()
    h = _cabi.ModelHandle(d)
    assert (h.info.nb, h.info.nq, h.info.nv) == (31, 37, 36)
    assert list(h.info.qstart[:31]) == list(d.qstart) and list(h.info.vstart[:31]) == list(d.vstart)
    order = list(h.info.eval_order[:31])
    assert sorted(order) == list(range(31)) and order[0] == 0
    # depth-first preorder: every body's parent precedes it
    pos = {j: p for p, j in enumerate(order)}
    assert all(d.parent[j] < 0 or pos[d.parent[j]] < pos[j] for j in range(31))
    assert h.info.general_path == 0 and h.info.max_branch_depth == 1      # pending slots after Sethi-Ullman child ordering


def test_status_codes(built):
    lib = rbd.load_library()
    mech = rbd.load_model("iiwa14")
    d = mech.flatten()
    # loops -> RBD_ELOOP with the reference's message (mechanism_algorithms.jl:549)
    with pytest.raises(rbd.RbdError) as e:
        _cabi.ModelHandle(d, num_non_tree_joints=1)
    assert e.value.status == _cabi.RBD_ELOOP and "tree Mechanisms" in str(e.value)
    # malformed parent array -> RBD_EINVAL
    bad = mech.flatten()
    bad.parent = bad.parent.copy(); bad.parent[0] = 3
    with pytest.raises(rbd.RbdError) as e:
        _cabi.ModelHandle(bad)
    assert e.value.status == _cabi.RBD_EINVAL
    # stale handle -> RBD_ESTALE (ModificationCountMismatch, util.jl:56-72)
    h = _cabi.ModelHandle(d)
    h.check_modcount(d.modcount)
    with pytest.raises(rbd.RbdError) as e:
        h.check_modcount(d.modcount + 1)
    assert e.value.status == _cabi.RBD_ESTALE
    # argument checks happen before any CUDA call
    assert lib.rbd_dynamics(None, 0, 1, 1, None, None, None, None, None, None, None) == _cabi.RBD_EINVAL
    assert lib.rbd_dynamics(h.ptr, 7, 1, 1, None, None, None, None, None, None, None) == _cabi.RBD_EINVAL
    assert lib.rbd_dynamics(h.ptr, 0, 8, 4, None, None, None, None, None, None, None) == _cabi.RBD_EDIM
    assert lib.rbd_status_string(_cabi.RBD_EDIM) == b"RBD_EDIM"


def test_too_many_bodies_is_unsupported(built):
    rng = np.random.default_rng(0)
    mech = rbd.rand_chain_mechanism(rng, [rbd.Revolute] * 65)
    with pytest.raises(rbd.RbdError) as e:
        _cabi.ModelHandle(mech.flatten())
    assert e.value.status == _cabi.RBD_EUNSUPPORTED


def test_kinematics_and_integrate_argument_checks(built):
    """rbd_kinematics / rbd_integrate validate their arguments on the host, before any CUDA call (no GPU needed)."""
    lib = rbd.load_library()
    h = _cabi.ModelHandle(rbd.load_model("iiwa14").flatten())
    ko = _cabi.RbdKinematicsOut()
    assert ctypes.sizeof(ko) == 8 * ctypes.sizeof(ctypes.c_void_p)           # eight output pointers, as in the header
    dummy = ctypes.c_void_p(16)                                               # never dereferenced by the checks below
    assert lib.rbd_kinematics(None, 0, 1, 1, dummy, None, None, ctypes.byref(ko), None) == _cabi.RBD_EINVAL
    assert lib.rbd_kinematics(h.ptr, 0, 1, 1, dummy, None, None, None, None) == _cabi.RBD_EINVAL              # out == NULL
    assert lib.rbd_kinematics(h.ptr, _cabi.RBD_DUAL64X6, 1, 1, dummy, None, None, ctypes.byref(ko), None) == _cabi.RBD_EUNSUPPORTED
    assert lib.rbd_kinematics(h.ptr, 0, 4, 2, dummy, None, None, ctypes.byref(ko), None) == _cabi.RBD_EDIM    # ld < B
    assert lib.rbd_kinematics(h.ptr, 0, 0, 0, None, None, None, ctypes.byref(ko), None) == _cabi.RBD_OK       # empty batch
    assert lib.rbd_kinematics(h.ptr, 0, 1, 1, None, None, None, ctypes.byref(ko), None) == _cabi.RBD_EINVAL   # q == NULL
    ko.kinetic_energy = 16
    assert lib.rbd_kinematics(h.ptr, 0, 1, 1, dummy, None, None, ctypes.byref(ko), None) == _cabi.RBD_EINVAL  # needs v
    ko.kinetic_energy = None
    ko.geometric_jacobian = 16
    assert lib.rbd_kinematics(h.ptr, 0, 1, 1, dummy, None, None, ctypes.byref(ko), None) == _cabi.RBD_EINVAL  # needs path_sign
    bad = (ctypes.c_int8 * 7)(1, 1, 2, 0, 0, 0, 0)
    assert lib.rbd_kinematics(h.ptr, 0, 1, 1, dummy, None, bad, ctypes.byref(ko), None) == _cabi.RBD_EINVAL   # sign not in {-1,0,1}
    assert b"path_sign" in lib.rbd_last_error()
    assert lib.rbd_integrate(None, 0, 1, 1, dummy, dummy, None, 1e-3, 1, None) == _cabi.RBD_EINVAL
    assert lib.rbd_integrate(h.ptr, 0, 4, 2, dummy, dummy, None, 1e-3, 1, None) == _cabi.RBD_EDIM


def test_derivative_and_precompile_entry_points_without_a_gpu(built):
    """Host-side behaviour of the round-2 entry points that needs no device: argument checks of rbd_dynamics_derivatives (empty batch,
    NULL pointers, Dual dtype, size mismatch) and ahead-of-time compilation of the model-specialised kernels for the derivative solve
    and for mass_matrix! (NVRTC runs without a GPU; skipped when NVRTC is not installed)."""
    lib = rbd.load_library()
    h = _cabi.ModelHandle(rbd.load_model("iiwa14").flatten())
    fake = ctypes.c_void_p(64)
    args = lambda dtype, B, ld, q: (h.ptr, dtype, B, ld, q, fake, None, fake, fake, fake, None)     # noqa: E731
    assert lib.rbd_dynamics_derivatives(*args(_cabi.RBD_F64, 0, 0, None)) == _cabi.RBD_OK          # empty batch: nothing touched
    assert lib.rbd_dynamics_derivatives(*args(_cabi.RBD_F64, 4, 4, None)) == _cabi.RBD_EINVAL      # q NULL
    assert lib.rbd_dynamics_derivatives(*args(_cabi.RBD_DUAL64X6, 4, 4, fake)) == _cabi.RBD_EUNSUPPORTED
    assert lib.rbd_dynamics_derivatives(*args(_cabi.RBD_F32, 8, 4, fake)) == _cabi.RBD_EDIM        # ld < B
    assert lib.rbd_dynamics_derivatives(None, _cabi.RBD_F32, 4, 4, fake, fake, None, fake, fake, fake, None) == _cabi.RBD_EINVAL
    assert lib.rbd_model_precompile_derivatives(None, _cabi.RBD_F64) == _cabi.RBD_EINVAL
    assert lib.rbd_model_precompile_derivatives(h.ptr, _cabi.RBD_DUAL64X6) == _cabi.RBD_EINVAL
    rc = lib.rbd_model_precompile_derivatives(h.ptr, _cabi.RBD_F64)
    if rc == _cabi.RBD_EUNSUPPORTED and b"NVRTC" in lib.rbd_last_error():
        pytest.skip("NVRTC not available")
    assert rc == _cabi.RBD_OK, lib.rbd_last_error()
    assert lib.rbd_model_precompile_derivatives(h.ptr, _cabi.RBD_F32) == _cabi.RBD_OK
    h.precompile(_cabi.RBD_F32, _cabi.RBD_SPEC_MASS_MATRIX | _cabi.RBD_SPEC_MASS_MATRIX_LOWER, load=False)
    h.precompile(_cabi.RBD_F64, _cabi.RBD_SPEC_MASS_MATRIX, load=False)
    h.close()
