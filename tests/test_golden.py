"""Committed fixtures (tests/golden/, written by tools/make_golden.py).

  * double_pendulum_closed_form.json: reference-derived known answers (closed form of test/test_double_pendulum.jl:40-75 and the
    quick-start numbers of SURVEY 8(c)(1)); CPU tier: the oracle must reproduce them to the reference's own 1e-12; GPU tier: the
    CUDA path (fp64) must too.
  * <model>_seed<k>.npz: frozen fp64 oracle outputs on seeded inputs; CPU tier: the oracle and the host-compiled device code still
    reproduce them; GPU tier: every entry point of the C ABI against them (fp64 1e-9, fp32 2e-5, relative as in test_gpu_parity).
"""
import json
import os

import numpy as np
import pytest

import rigidbodydynamics.jl_b200 as rbd
from oracle import Oracle
from tests import hostsim
from tests.util import double_pendulum, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NPZ = [("atlas", True, 17), ("valkyrie", True, 18), ("iiwa14", False, 19)]
KIN = {"transforms": "transforms_to_root", "com": "center_of_mass", "ke": "kinetic_energy", "pe": "gravitational_potential_energy",
       "momentum": "momentum", "mrb": "momentum_rate_bias", "A": "momentum_matrix", "J": "geometric_jacobian"}


def _closed_form():
    with open(os.path.join(GOLD, "double_pendulum_closed_form.json")) as f:
        return json.load(f)


def _pendulum(p):
    return double_pendulum(p["I1"], p["I2"], p["lc1"], p["lc2"], p["l1"], p["m1"], p["m2"], p["g"])


def test_oracle_reproduces_closed_form_fixture():
    doc = _closed_form()
    for params, cases in ((doc["test_parameters"], doc["test_cases"]), (doc["quickstart_parameters"], [doc["quickstart"]])):
        o = Oracle(_pendulum(params).flatten())
        for c in cases:
            q, v, vd = (np.array(c[k]) for k in ("q", "v", "vd"))
            assert np.allclose(o.mass_matrix(q).reshape(2, 2), c["M"], rtol=0, atol=1e-12)
            assert np.allclose(o.dynamics_bias(q, v).ravel(), c["bias"], rtol=0, atol=1e-12)
            assert np.allclose(o.inverse_dynamics(q, v, vd).ravel(), c["tau"], rtol=0, atol=1e-12)
            assert np.allclose(o.dynamics(q, v).ravel(), c["vd_passive"], rtol=0, atol=1e-10)
            assert abs(o.kinematics(q, v, want=("ke",))["ke"][0, 0] - c["kinetic_energy"]) < 1e-12
    quick = doc["quickstart"]                      # the numbers quoted in SURVEY 8(c)(1)
    assert np.allclose(quick["M"], [[2.587060994002885, 0.7935304970014425], [0.7935304970014425, 0.333]], atol=1e-14)
    assert np.allclose(quick["vd_passive"], [2.935110215118255, -17.068157341777777], atol=1e-12)


@pytest.mark.parametrize("name,floating,seed", NPZ)
def test_oracle_and_device_code_reproduce_regression_vectors(name, floating, seed):
    g = np.load(os.path.join(GOLD, f"{name}_seed{seed}.npz"))
    mech = rbd.load_model(name, floating=floating)
    desc = mech.flatten()
    o = Oracle(desc)
    q, v, tau, vd, w = g["q"], g["v"], g["tau"], g["vd_in"], g["wext"]
    assert rel_err(o.dynamics(q, v, tau), g["dynamics"]) < 1e-12
    assert rel_err(o.inverse_dynamics(q, v, vd, w), g["inverse_dynamics_wext"]) < 1e-12
    assert rel_err(o.mass_matrix(q), g["mass_matrix"]) < 1e-12
    # device code compiled for the host
    assert rel_err(hostsim.dynamics(desc, q, v, tau), g["dynamics"]) < 1e-9
    assert rel_err(hostsim.dynamics(desc, q, v, tau, w), g["dynamics_wext"]) < 1e-9
    assert rel_err(hostsim.inverse_dynamics(desc, q, v, vd), g["inverse_dynamics"]) < 1e-10
    assert rel_err(hostsim.mass_matrix(desc, q), g["mass_matrix"]) < 1e-10
    kin = hostsim.kinematics(desc, q, v, g["path_sign"])
    for k in KIN:
        assert rel_err(kin[k], g["kin_" + k]) < 1e-10, k


@pytest.mark.gpu
def test_gpu_closed_form_fixture(built):
    import torch
    doc = _closed_form()
    for params, cases in ((doc["test_parameters"], doc["test_cases"]), (doc["quickstart_parameters"], [doc["quickstart"]])):
        mech = _pendulum(params)
        B = len(cases)
        st = rbd.MechanismState(mech, B, torch.float64)
        st.q.copy_(torch.tensor([c["q"] for c in cases], dtype=torch.float64).T)
        st.v.copy_(torch.tensor([c["v"] for c in cases], dtype=torch.float64).T)
        vd = torch.tensor([c["vd"] for c in cases], dtype=torch.float64, device="cuda").T.contiguous()
        assert np.allclose(rbd.mass_matrix(st).cpu().numpy().T.reshape(B, 2, 2), [c["M"] for c in cases], rtol=0, atol=1e-12)
        assert np.allclose(rbd.dynamics_bias(st).cpu().numpy().T, [c["bias"] for c in cases], rtol=0, atol=1e-12)
        assert np.allclose(rbd.inverse_dynamics(st, vd).cpu().numpy().T, [c["tau"] for c in cases], rtol=0, atol=1e-12)
        res = rbd.DynamicsResult(mech, B, torch.float64)
        rbd.dynamics_(res, st)
        assert np.allclose(res.vd.cpu().numpy().T, [c["vd_passive"] for c in cases], rtol=0, atol=1e-10)
        assert np.allclose(rbd.kinetic_energy(st).cpu().numpy(), [c["kinetic_energy"] for c in cases], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name,floating,seed", NPZ)
def test_gpu_regression_vectors(built, name, floating, seed):
    import torch
    g = np.load(os.path.join(GOLD, f"{name}_seed{seed}.npz"))
    mech = rbd.load_model(name, floating=floating)
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 2e-5)):
        cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype).cuda()
        B = g["q"].shape[1]
        st = rbd.MechanismState(mech, B, dtype)
        st.q.copy_(cu(g["q"])); st.v.copy_(cu(g["v"]))
        tau, vd, w = cu(g["tau"]), cu(g["vd_in"]), cu(g["wext"])
        res = rbd.DynamicsResult(mech, B, dtype)
        num = lambda t: t.double().cpu().numpy()
        rbd.dynamics_(res, st, tau)
        assert rel_err(num(res.vd), g["dynamics"]) < tol
        assert np.abs(num(res.qd) - g["qdot"]).max() < (1e-12 if dtype == torch.float64 else 1e-5)
        rbd.dynamics_(res, st, tau, w)
        assert rel_err(num(res.vd), g["dynamics_wext"]) < tol
        assert rel_err(num(rbd.inverse_dynamics(st, vd)), g["inverse_dynamics"]) < tol
        assert rel_err(num(rbd.inverse_dynamics(st, vd, w)), g["inverse_dynamics_wext"]) < tol
        assert rel_err(num(rbd.dynamics_bias(st)), g["dynamics_bias"]) < tol
        assert rel_err(num(rbd.mass_matrix(st)), g["mass_matrix"]) < tol
        p = rbd.TreePath(None, None, g["path_sign"])
        outs = {KIN[k]: torch.empty(g["kin_" + k].shape, dtype=dtype, device="cuda") for k in KIN}
        rbd.kinematics_(st, p, **outs)
        ktol = 1e-10 if dtype == torch.float64 else 2e-5
        for k in KIN:
            assert rel_err(num(outs[KIN[k]]), g["kin_" + k]) < ktol, k
