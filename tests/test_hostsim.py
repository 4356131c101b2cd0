"""CPU tier: the per-sample DEVICE algorithms (csrc/rbd_device.cuh, compiled for the host by tests/hostsim) against the oracle.
This checks the kernels' math -- body-frame ABA, one-hot subspaces, depth-first ordering, pending slots -- without a GPU."""
import numpy as np
import pytest

import rigidbodydynamics.jl_b200 as rbd
from oracle import Oracle
from tests import hostsim
from tests.util import axis_aligned_tree, config_distance, make_duals, rand_inputs, randmech, rel_err

MODELS = [("atlas", True), ("atlas", False), ("valkyrie", True), ("iiwa14", False), ("double_pendulum", False)]


@pytest.mark.parametrize("name,floating", MODELS)
def test_aba_matches_oracle_fp64(name, floating):
    mech = rbd.load_model(name, floating=floating)
    d = mech.flatten()
    q, v, tau, _, _ = rand_inputs(mech, 12, 17)
    ref, ref_qd = Oracle(d).dynamics(q, v, tau, want_qd=True)
    got, got_qd = hostsim.dynamics(d, q, v, tau, want_qd=True)
    assert rel_err(got, ref) < 1e-11
    assert np.abs(got_qd - ref_qd).max() < 1e-14
    # zero torques = the ConstVector default
    assert rel_err(hostsim.dynamics(d, q, v), Oracle(d).dynamics(q, v)) < 1e-11


@pytest.mark.parametrize("name,floating", MODELS[:4])
def test_aba_fp32_accuracy(name, floating):
    """fp32 body-frame ABA stays within 5e-6 of the fp64 oracle (the world-frame fp32 forms are ~1e-4)."""
    mech = rbd.load_model(name, floating=floating)
    d = mech.flatten()
    q, v, tau, _, _ = rand_inputs(mech, 32, 23)
    ref = Oracle(d).dynamics(q, v, tau)
    got = hostsim.dynamics(d, q.astype(np.float32), v.astype(np.float32), tau.astype(np.float32))
    assert rel_err(got, ref) < 5e-6


@pytest.mark.parametrize("seed", range(6))
def test_aba_general_trees_all_joint_types(seed):
    mech = randmech(seed, shuffle=seed % 2 == 1)
    d = mech.flatten()
    info = hostsim.info(d)
    assert info["general"] == 1 and sorted(info["order"]) == list(range(d.nb))
    q, v, tau, _, _ = rand_inputs(mech, 5, seed)
    ref, ref_qd = Oracle(d).dynamics(q, v, tau, want_qd=True)
    got, got_qd = hostsim.dynamics(d, q, v, tau, want_qd=True)
    assert rel_err(got, ref) < 1e-10
    assert np.abs(got_qd - ref_qd).max() < 1e-13


@pytest.mark.parametrize("jt", [rbd.Revolute, rbd.Prismatic, rbd.Planar, rbd.QuaternionFloating, rbd.SPQuatFloating,
                                rbd.QuaternionSpherical, rbd.SinCosRevolute, rbd.Fixed])
def test_aba_single_joint_type_chains(jt):
    rng = np.random.default_rng(5)
    types = [jt] * 4 if jt is not rbd.Fixed else [rbd.Revolute, rbd.Fixed, rbd.Revolute, rbd.Fixed, rbd.Prismatic]
    mech = rbd.rand_chain_mechanism(rng, types)
    d = mech.flatten()
    q, v, tau, _, _ = rand_inputs(mech, 4, 9)
    assert rel_err(hostsim.dynamics(d, q, v, tau), Oracle(d).dynamics(q, v, tau)) < 1e-10


@pytest.mark.parametrize("name,floating", MODELS)
def test_rnea_crba_wext_match_oracle(name, floating):
    """inverse_dynamics!, dynamics_bias!, mass_matrix! and external wrenches (root-frame wrench on every body, as in
    perf/runbenchmarks.jl:49-67) through the device code."""
    mech = rbd.load_model(name, floating=floating)
    d = mech.flatten()
    o = Oracle(d)
    q, v, tau, vd, w = rand_inputs(mech, 7, 29, wext=True)
    assert rel_err(hostsim.inverse_dynamics(d, q, v, vd), o.inverse_dynamics(q, v, vd)) < 1e-13
    assert rel_err(hostsim.inverse_dynamics(d, q, v, vd, w), o.inverse_dynamics(q, v, vd, w)) < 1e-13
    assert rel_err(hostsim.inverse_dynamics(d, q, v, None, w), o.dynamics_bias(q, v, w)) < 1e-13
    M = hostsim.mass_matrix(d, q)
    assert not np.isnan(M).any()                 # every entry written, including the structural zeros
    assert rel_err(M, o.mass_matrix(q)) < 1e-13
    assert rel_err(hostsim.dynamics(d, q, v, tau, w), o.dynamics(q, v, tau, w)) < 1e-10


@pytest.mark.parametrize("seed", range(6))
def test_rnea_crba_wext_general_trees(seed):
    mech = randmech(seed, shuffle=seed % 2 == 1)
    d = mech.flatten()
    o = Oracle(d)
    q, v, tau, vd, w = rand_inputs(mech, 4, seed, wext=True)
    assert rel_err(hostsim.inverse_dynamics(d, q, v, vd, w), o.inverse_dynamics(q, v, vd, w)) < 1e-12
    M = hostsim.mass_matrix(d, q)
    assert not np.isnan(M).any() and rel_err(M, o.mass_matrix(q)) < 1e-12
    assert rel_err(hostsim.dynamics(d, q, v, tau, w), o.dynamics(q, v, tau, w)) < 1e-10


def test_device_identities_fd_id_roundtrip():
    """dynamics / inverse dynamics round trip (test_mechanism_algorithms.jl:729-740) and M v̇ + c = tau, device code only."""
    mech = rbd.load_model("atlas", floating=True)
    d = mech.flatten()
    q, v, tau, _, w = rand_inputs(mech, 5, 40, wext=True)
    acc = hostsim.dynamics(d, q, v, tau, w)
    assert np.abs(hostsim.inverse_dynamics(d, q, v, acc, w) - tau).max() < 1e-9
    M = hostsim.mass_matrix(d, q).reshape(d.nv, d.nv, -1)
    c = hostsim.inverse_dynamics(d, q, v, None, w)
    assert np.abs(np.einsum("jib,jb->ib", M, acc) + c - tau).max() < 1e-8


@pytest.mark.parametrize("name,floating", [("atlas", True), ("iiwa14", False)])
def test_dual_number_dynamics(name, floating):
    """Config 4: dynamics! on Dual{Float64,6} inputs through the device code == the oracle's dual-number run of the
    reference's CRBA + RNEA + Cholesky path, and == central finite differences of the fp64 path."""
    mech = rbd.load_model(name, floating=floating)
    d = mech.flatten()
    o = Oracle(d)
    q, v, tau, _, _ = rand_inputs(mech, 3, 4)
    Q, V, T = make_duals(mech, q, v, tau, 5)
    ref = o.dynamics_dual6(Q, V, T)
    got = hostsim.dynamics_dual(d, Q, V, T)
    assert not np.isnan(got).any()
    assert np.abs(got[..., 0] - ref[..., 0]).max() / np.abs(ref[..., 0]).max() < 1e-11
    assert np.abs(got[..., 1:] - ref[..., 1:]).max() / np.abs(ref[..., 1:]).max() < 1e-9
    h = 1e-6
    fd = (o.dynamics(q + h * Q[..., 3], v + h * V[..., 3], tau + h * T[..., 3])
          - o.dynamics(q - h * Q[..., 3], v - h * V[..., 3], tau - h * T[..., 3])) / (2 * h)
    assert np.abs(fd - got[..., 3]).max() / np.abs(fd).max() < 1e-6


@pytest.mark.parametrize("name,floating", [("atlas", True), ("iiwa14", False), ("double_pendulum", False)])
def test_rk4_step_matches_oracle(name, floating):
    """simulate / MuntheKaasIntegrator (src/ode_integrators.jl:233-300): device coordinate maps + device ABA vs the oracle's
    restatement with the reference's closed forms, fp64, several steps."""
    mech = rbd.load_model(name, floating=floating)
    d = mech.flatten()
    q, v, tau, _, _ = rand_inputs(mech, 3, 12)
    for tq in (None, tau):
        qr, vr = Oracle(d).integrate(q, v, tq, dt=1e-3, nsteps=5)
        qg, vg = hostsim.integrate(d, q, v, tq, dt=1e-3, nsteps=5)
        assert config_distance(mech, qg, qr) < 1e-11
        assert rel_err(vg, vr) < 1e-10


@pytest.mark.parametrize("seed", range(3))
def test_rk4_step_all_joint_types(seed):
    mech = randmech(seed, shuffle=seed % 2 == 1)
    d = mech.flatten()
    q, v, tau, _, _ = rand_inputs(mech, 2, seed)
    qr, vr = Oracle(d).integrate(q, v, tau, dt=5e-4, nsteps=3)
    qg, vg = hostsim.integrate(d, q, v, tau, dt=5e-4, nsteps=3)
    assert config_distance(mech, qg, qr) < 1e-10
    assert rel_err(vg, vr) < 1e-9


def test_rk4_fp32_small_angle_series():
    """fp32: the Taylor branch of the dexp^-1 coefficients keeps the floating-base step accurate (1e-5 of the fp64 oracle)."""
    mech = rbd.load_model("atlas", floating=True)
    d = mech.flatten()
    q, v, tau, _, _ = rand_inputs(mech, 4, 2)
    qr, vr = Oracle(d).integrate(q, v, tau, dt=1e-3, nsteps=4)
    qg, vg = hostsim.integrate(d, q.astype(np.float32), v.astype(np.float32), tau.astype(np.float32), dt=1e-3, nsteps=4)
    assert config_distance(mech, qg, qr) < 2e-5
    assert rel_err(vg, vr) < 2e-4


def test_energy_conservation_passive_pendulum():
    """test/test_simulate.jl:5-13: total energy of the passive double pendulum changes by < 1e-3 over 0.1 s at dt = 1e-2."""
    from tests.util import double_pendulum
    mech = double_pendulum()
    d = mech.flatten()
    o = Oracle(d)

    def energy(q, v):
        M = o.mass_matrix(q).reshape(2, 2)
        z1 = -0.5 * np.cos(q[0, 0]); z2 = -np.cos(q[0, 0]) - 0.5 * np.cos(q[0, 0] + q[1, 0])
        return 0.5 * v[:, 0] @ M @ v[:, 0] + 9.81 * (z1 + z2)
    q, v = np.array([[0.3], [0.4]]), np.array([[1.0], [2.0]])
    q1, v1 = hostsim.integrate(d, q, v, None, dt=1e-2, nsteps=10)
    assert abs(energy(q1, v1) - energy(q, v)) < 1e-3


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_axis_aligned_trees_take_the_fast_classes(seed):
    mech = axis_aligned_tree(seed)
    desc = mech.flatten()
    flags = hostsim.flags(desc)
    npar, nperp, nzr = (sum(1 for f in flags if f & b) for b in (32, 64, 128))
    assert npar + nperp >= 5, (npar, nperp)                       # the fast classes really are exercised
    q, v, tau, vd, w = rand_inputs(mech, 5, seed, wext=True)
    o = Oracle(desc)
    assert rel_err(hostsim.dynamics(desc, q, v, tau), o.dynamics(q, v, tau)) < 1e-9
    assert rel_err(hostsim.dynamics(desc, q, v, tau, w), o.dynamics(q, v, tau, w)) < 1e-9
    assert rel_err(hostsim.dynamics(desc, q.astype(np.float32), v.astype(np.float32), tau.astype(np.float32)),
                   o.dynamics(q, v, tau)) < 5e-4
    assert rel_err(hostsim.inverse_dynamics(desc, q, v, vd, w), o.inverse_dynamics(q, v, vd, w)) < 1e-10
    kin = hostsim.kinematics(desc, q, v, None, want=("transforms", "A"))
    ref = o.kinematics(q, v, None, want=("transforms", "A"))
    assert rel_err(kin["transforms"], ref["transforms"]) < 1e-11 and rel_err(kin["A"], ref["A"]) < 1e-10
