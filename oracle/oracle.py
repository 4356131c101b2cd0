"""TEST INFRASTRUCTURE -- ctypes binding of oracle/librbd_oracle.so (CPU restatement of the reference path).

Arrays are numpy, structure-of-arrays ``[rows, B]`` C-contiguous (batch index fastest), float32 or float64.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "librbd_oracle.so")


def build_oracle(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("rbd_oracle.cpp", "rbd_oracle.hpp")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "librbd_oracle.so"])
    return _LIB


def _load():
    lib = ctypes.CDLL(build_oracle())
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.rbdo_model_create.restype = vp
    lib.rbdo_model_create.argtypes = [i32, vp, vp, vp, vp, vp, vp]
    lib.rbdo_model_destroy.argtypes = [vp]
    lib.rbdo_nq.argtypes = [vp]
    lib.rbdo_nv.argtypes = [vp]
    lib.rbdo_dynamics.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp, i32, i32]
    lib.rbdo_dynamics_dual6.argtypes = [vp, i64, vp, vp, vp, vp, i32, i32]
    lib.rbdo_integrate.argtypes = [vp, i64, vp, vp, vp, ctypes.c_double, i32, i32]
    lib.rbdo_inverse_dynamics.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, i32]
    lib.rbdo_inverse_dynamics_bodies.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp]
    lib.rbdo_contact_dynamics.argtypes = [vp, i32, i64, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.rbdo_mass_matrix.argtypes = [vp, i32, i64, vp, vp, i32]
    lib.rbdo_kinematics.argtypes = [vp, i32, i64] + [vp] * 11 + [i32]
    return lib


_lib = None


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """CPU oracle bound to one flattened model (a ``ModelDesc`` from ``Mechanism.flatten()``)."""

    def __init__(self, desc):
        global _lib
        if _lib is None:
            _lib = _load()
        self.desc = desc
        parent = np.ascontiguousarray(desc.parent, np.int32)
        jtype = np.ascontiguousarray(desc.jtype, np.int32)
        X = np.ascontiguousarray(desc.X_tree, np.float64)
        jp = np.ascontiguousarray(desc.jparam, np.float64)
        inr = np.ascontiguousarray(desc.inertia, np.float64)
        g = np.ascontiguousarray(desc.gravity, np.float64)
        self._h = _lib.rbdo_model_create(desc.nb, _ptr(parent), _ptr(jtype), _ptr(X), _ptr(jp), _ptr(inr), _ptr(g))
        self.nq, self.nv, self.nb = desc.nq, desc.nv, desc.nb
        assert _lib.rbdo_nq(self._h) == self.nq and _lib.rbdo_nv(self._h) == self.nv

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rbdo_model_destroy(self._h)
            self._h = None

    @staticmethod
    def _prep(a, rows, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dt)
        if a.ndim == 1:
            a = a.reshape(rows, 1)
        assert a.shape[0] == rows, (a.shape, rows)
        return a

    @staticmethod
    def _code(dt):
        return 0 if np.dtype(dt) == np.float32 else 1

    def dynamics(self, q, v, tau=None, wext=None, *, algo="reference", want_qd=False, nthreads=1, dtype=None):
        """algo='reference': RNEA bias + CRBA + Cholesky (the reference's dynamics!); algo='aba': world-frame ABA."""
        dt = np.dtype(dtype or np.asarray(q).dtype)
        q = self._prep(q, self.nq, dt); v = self._prep(v, self.nv, dt)
        tau = self._prep(tau, self.nv, dt); wext = self._prep(wext, self.nb * 6, dt)
        B = q.shape[1]
        vd = np.empty((self.nv, B), dt)
        qd = np.empty((self.nq, B), dt) if want_qd else None
        rc = _lib.rbdo_dynamics(self._h, self._code(dt), B, _ptr(q), _ptr(v), _ptr(tau), _ptr(wext), _ptr(vd), _ptr(qd),
                                0 if algo == "reference" else 1, nthreads)
        if rc != 0:
            raise np.linalg.LinAlgError("mass matrix not positive definite")
        return (vd, qd) if want_qd else vd

    def dynamics_dual6(self, q, v, tau=None, *, algo="reference", nthreads=1):
        """dynamics! on Dual{Float64,6} inputs: arrays [rows, B, 7] float64 (value, 6 partials)."""
        q = np.ascontiguousarray(q, np.float64); v = np.ascontiguousarray(v, np.float64)
        tau = None if tau is None else np.ascontiguousarray(tau, np.float64)
        assert q.shape[0] == self.nq and q.shape[2] == 7 and v.shape[0] == self.nv
        B = q.shape[1]
        vd = np.empty((self.nv, B, 7))
        rc = _lib.rbdo_dynamics_dual6(self._h, B, _ptr(q), _ptr(v), _ptr(tau), _ptr(vd), 0 if algo == "reference" else 1, nthreads)
        if rc != 0:
            raise np.linalg.LinAlgError("mass matrix not positive definite")
        return vd

    def integrate(self, q, v, tau=None, *, dt=1e-4, nsteps=1, nthreads=1):
        """simulate(): ``nsteps`` Munthe-Kaas RK4 steps with constant torques (fp64); returns (q, v)."""
        q = np.array(self._prep(q, self.nq, np.float64)); v = np.array(self._prep(v, self.nv, np.float64))
        tau = self._prep(tau, self.nv, np.float64)
        rc = _lib.rbdo_integrate(self._h, q.shape[1], _ptr(q), _ptr(v), _ptr(tau), float(dt), int(nsteps), nthreads)
        if rc != 0:
            raise np.linalg.LinAlgError("mass matrix not positive definite")
        return q, v

    def inverse_dynamics(self, q, v, vd, wext=None, *, nthreads=1, dtype=None):
        dt = np.dtype(dtype or np.asarray(q).dtype)
        q = self._prep(q, self.nq, dt); v = self._prep(v, self.nv, dt)
        vd = self._prep(vd, self.nv, dt); wext = self._prep(wext, self.nb * 6, dt)
        B = q.shape[1]
        tau = np.empty((self.nv, B), dt)
        _lib.rbdo_inverse_dynamics(self._h, self._code(dt), B, _ptr(q), _ptr(v), _ptr(vd), _ptr(wext), _ptr(tau), nthreads)
        return tau

    def inverse_dynamics_bodies(self, q, v, vd=None, wext=None, *, dtype=None):
        """The per-body caches inverse_dynamics! fills: (accelerations, jointwrenches), each [6*nb, B], root frame, rows
        6i..6i+5 = [angular; linear] / [torque; force] of the successor of tree joint i (mechanism_algorithms.jl:387-459)."""
        dt = np.dtype(dtype or np.asarray(q).dtype)
        q = self._prep(q, self.nq, dt); v = self._prep(v, self.nv, dt)
        vd = self._prep(vd, self.nv, dt); wext = self._prep(wext, self.nb * 6, dt)
        B = q.shape[1]
        acc = np.empty((6 * self.nb, B), dt); jw = np.empty((6 * self.nb, B), dt)
        _lib.rbdo_inverse_dynamics_bodies(self._h, self._code(dt), B, _ptr(q), _ptr(v), _ptr(vd), _ptr(wext), _ptr(acc), _ptr(jw))
        return acc, jw

    def contact_dynamics(self, q, v, contact, s=None, *, dtype=None):
        """contact_dynamics! (mechanism_algorithms.jl:680-723) for a ``ContactDesc``-like object (body, location, normal_model,
        friction_model, halfspace).  Returns (contactwrenches [6*nb, B], state_derivatives [ns, B], state after the resets [ns, B])."""
        dt = np.dtype(dtype or np.asarray(q).dtype)
        q = self._prep(q, self.nq, dt); v = self._prep(v, self.nv, dt)
        B = q.shape[1]
        body = np.ascontiguousarray(contact.body, np.int32)
        loc = np.ascontiguousarray(contact.location, np.float64); hc = np.ascontiguousarray(contact.normal_model, np.float64)
        fr = np.ascontiguousarray(contact.friction_model, np.float64); hs = np.ascontiguousarray(contact.halfspace, np.float64)
        npnt, nh = len(body), len(hs)
        ns = 3 * npnt * nh
        s = np.zeros((ns, B), dt) if s is None else np.array(self._prep(s, ns, dt), copy=True)
        sd = np.zeros((ns, B), dt); wr = np.empty((6 * self.nb, B), dt)
        _lib.rbdo_contact_dynamics(self._h, self._code(dt), B, _ptr(q), _ptr(v), npnt, _ptr(body), _ptr(loc), _ptr(hc), _ptr(fr), nh,
                                   _ptr(hs), _ptr(s), _ptr(sd), _ptr(wr))
        return wr, sd, s

    def dynamics_bias(self, q, v, wext=None, *, nthreads=1, dtype=None):
        return self.inverse_dynamics(q, v, None, wext, nthreads=nthreads, dtype=dtype)

    def mass_matrix(self, q, *, nthreads=1, dtype=None):
        """Returns [nv*nv, B]; entry (i, j) of sample b at row i + j*nv (column-major like M.data), both triangles filled."""
        dt = np.dtype(dtype or np.asarray(q).dtype)
        q = self._prep(q, self.nq, dt)
        B = q.shape[1]
        M = np.empty((self.nv * self.nv, B), dt)
        _lib.rbdo_mass_matrix(self._h, self._code(dt), B, _ptr(q), _ptr(M), nthreads)
        return M

    def kinematics(self, q, v=None, sign=None, *, want=("transforms", "com", "ke", "pe", "momentum", "mrb", "A", "J"),
                   nthreads=1, dtype=None):
        """Kinematics by-products in the root frame.  Returns a dict of [rows, B] arrays:
        transforms [12*nb] (R row-major, p per body), com [3], ke [1], pe [1], momentum [6], mrb (momentum_rate_bias) [6],
        A (momentum matrix) [6*nv], J (geometric jacobian of the path given by ``sign`` [nb] in {-1, 0, +1}) [6*nv]."""
        dt = np.dtype(dtype or np.asarray(q).dtype)
        q = self._prep(q, self.nq, dt); v = self._prep(v, self.nv, dt)
        B = q.shape[1]
        rows = {"transforms": 12 * self.nb, "com": 3, "ke": 1, "pe": 1, "momentum": 6, "mrb": 6, "A": 6 * self.nv, "J": 6 * self.nv}
        out = {k: (np.empty((rows[k], B), dt) if k in want else None) for k in rows}
        if v is None:
            for k in ("ke", "momentum", "mrb"):
                out[k] = None
        sg = None if sign is None else np.ascontiguousarray(sign, np.int8)
        if sg is None:
            out["J"] = None
        _lib.rbdo_kinematics(self._h, self._code(dt), B, _ptr(q), _ptr(v), _ptr(sg), *[_ptr(out[k]) for k in rows], nthreads)
        return {k: a for k, a in out.items() if a is not None}
