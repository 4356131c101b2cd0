"""Pins the CPU oracle (oracle/) against everything the reference's own tests hold for the hot path.

  * closed-form double pendulum  -- test/test_double_pendulum.jl:2-11,51-65,72-75 (atol 1e-12)
  * quick-start state values     -- examples/1: M, c, inverse dynamics, forward dynamics
  * identity / property tests    -- test/test_mechanism_algorithms.jl:564-572 (KE), 729-740 (FD o ID), 742-753 (bias = ID(0)),
                                    600-614 (columns of M = ID(e_k) - ID(0))
The reference holds no stored numeric vectors for Atlas-sized models; for those the oracle is pinned by the
identities plus the agreement of two independent formulations (CRBA + RNEA + Cholesky vs world-frame ABA).
"""
import numpy as np
import pytest

import rigidbodydynamics.jl_b200 as rbd
from oracle import Oracle
from tests.util import double_pendulum, rand_inputs, randmech, rel_err


def test_double_pendulum_closed_form():
    """The textbook M, C, G of test/test_double_pendulum.jl:51-65 with its parameters (:2-11)."""
    lc1, l1, m1, I1, lc2, l2, m2, I2, g = -0.5, -1.0, 1.0, 0.333, -1.0, -2.0, 1.0, 1.33, -9.81
    mech = double_pendulum(I1, I2, lc1, lc2, l1, m1, m2, g)
    o = Oracle(mech.flatten())
    rng = np.random.default_rng(6)
    for _ in range(10):
        q1, q2 = rng.standard_normal(2)
        v1, v2 = rng.random(2)
        vd = rng.random(2)
        c2, s1, s2, s12 = np.cos(q2), np.sin(q1), np.sin(q2), np.sin(q1 + q2)
        M = np.array([[I1 + I2 + m2 * l1 ** 2 + 2 * m2 * l1 * lc2 * c2, I2 + m2 * l1 * lc2 * c2],
                      [I2 + m2 * l1 * lc2 * c2, I2]])
        C = np.array([[-2 * m2 * l1 * lc2 * s2 * v2, -m2 * l1 * lc2 * s2 * v2], [m2 * l1 * lc2 * s2 * v1, 0]])
        G = np.array([m1 * g * lc1 * s1 + m2 * g * (l1 * s1 + lc2 * s12), m2 * g * lc2 * s12])
        q, v = np.array([q1, q2]), np.array([v1, v2])
        assert np.allclose(o.mass_matrix(q).reshape(2, 2), M, rtol=0, atol=1e-12)
        tau = o.inverse_dynamics(q, v, vd).ravel()
        assert np.allclose(tau, M @ vd + C @ v + G, rtol=0, atol=1e-12)
        # forward dynamics inverts it (both the reference's Cholesky path and the independent ABA)
        for algo in ("reference", "aba"):
            assert np.allclose(o.dynamics(q, v, tau, algo=algo).ravel(), vd, rtol=0, atol=1e-10)


def test_quickstart_state_values():
    """Config 1 acceptance numbers: examples/1 quick-start pendulum at q=(0.3,0.4), v=(1,2)."""
    o = Oracle(double_pendulum().flatten())
    q, v = np.array([0.3, 0.4]), np.array([1.0, 2.0])
    assert np.allclose(o.mass_matrix(q).reshape(2, 2),
                       [[2.587060994002885, 0.7935304970014425], [0.7935304970014425, 0.333]], atol=1e-12)
    assert np.allclose(o.dynamics_bias(q, v).ravel(), [5.950794227687885, 3.3545969270552], atol=1e-12)
    assert np.allclose(o.inverse_dynamics(q, v, np.array([1.0, 2.0])).ravel(),
                       [10.124916215693656, 4.814127424056642], atol=1e-12)
    assert np.allclose(o.dynamics(q, v).ravel(), [2.935110215118255, -17.068157341777777], atol=1e-12)


def test_urdf_double_pendulum_matches_api_model():
    """test/test_double_pendulum.jl:78-99: the URDF model gives the same M and tau (incl. the SinCosRevolute variant)."""
    api = Oracle(double_pendulum().flatten())
    urdf = Oracle(rbd.load_model("double_pendulum").flatten())
    jt = rbd.default_urdf_joint_types()
    jt["continuous"] = rbd.SinCosRevolute
    sincos = Oracle(rbd.load_model("double_pendulum", joint_types=jt).flatten())
    rng = np.random.default_rng(3)
    q, v, vd = rng.standard_normal((2, 5)), rng.random((2, 5)), rng.random((2, 5))
    qsc = np.stack([np.sin(q[0]), np.cos(q[0]), np.sin(q[1]), np.cos(q[1])])
    assert np.allclose(api.mass_matrix(q), urdf.mass_matrix(q), atol=1e-12)
    assert np.allclose(api.mass_matrix(q), sincos.mass_matrix(qsc), atol=1e-12)
    assert np.allclose(api.inverse_dynamics(q, v, vd), urdf.inverse_dynamics(q, v, vd), atol=1e-12)
    assert np.allclose(api.inverse_dynamics(q, v, vd), sincos.inverse_dynamics(qsc, v, vd), atol=1e-12)


@pytest.mark.parametrize("seed", [17, 25, 40, 41])
def test_identities_random_mechanism(seed):
    """The reference's cross-algorithm identities on its 25-joint random trees (all joint types, external wrenches)."""
    mech = randmech(seed, shuffle=seed % 2 == 1)
    o = Oracle(mech.flatten())
    B = 6
    q, v, tau, vd, w = rand_inputs(mech, B, seed, wext=True)
    nv = o.nv
    # dynamics / inverse dynamics round trip (test_mechanism_algorithms.jl:729-740, atol 1e-10)
    acc = o.dynamics(q, v, tau, w)
    assert np.abs(o.inverse_dynamics(q, v, acc, w) - tau).max() < 1e-10
    # the independent ABA agrees with the reference's CRBA + RNEA + Cholesky path
    assert rel_err(o.dynamics(q, v, tau, w, algo="aba"), acc) < 1e-10
    # dynamics_bias == inverse_dynamics(v̇ = 0) (:742-753)
    c = o.dynamics_bias(q, v, w)
    assert np.abs(c - o.inverse_dynamics(q, v, np.zeros((nv, B)), w)).max() < 1e-12
    # columns of M: tau(e_k) - tau(0) (:600-614), symmetry, M v̇ + c = tau
    M = o.mass_matrix(q).reshape(nv, nv, B).transpose(1, 0, 2)      # [i, j, b]
    assert np.abs(M - M.transpose(1, 0, 2)).max() == 0
    for k in range(0, nv, 7):
        e = np.zeros((nv, B)); e[k] = 1
        assert np.abs(o.inverse_dynamics(q, v, e, w) - c - M[:, k, :]).max() < 1e-10
    assert np.abs(np.einsum("ijb,jb->ib", M, acc) + c - tau).max() < 1e-9
    for b in range(B):
        assert np.linalg.eigvalsh(M[:, :, b]).min() > 0


def test_kinetic_energy_identity_chain():
    """1/2 v' M v equals the sum of body kinetic energies (:564-572), via the power identity on a gravity-free chain:
    with g = 0 and zero torque, d/dt (1/2 v' M v) = 0  <=>  v . (M v̇ + Ṁ v / 2) = 0, checked through v . c = v . (C v)
    where v'(Ṁ - 2C)v = 0 (skew symmetry, :616-652)."""
    rng = np.random.default_rng(5)
    mech = rbd.rand_chain_mechanism(rng, [rbd.Revolute] * 6)
    mech.gravitational_acceleration[:] = 0
    o = Oracle(mech.flatten())
    q, v, _, _, _ = rand_inputs(mech, 4, 5)
    M = lambda qq: o.mass_matrix(qq).reshape(o.nv, o.nv, -1).transpose(1, 0, 2)
    h = 1e-6
    Mdot = (M(q + h * v) - M(q - h * v)) / (2 * h)                   # q̇ = v for revolute joints
    c = o.dynamics_bias(q, v)
    lhs = np.einsum("ib,ib->b", v, c)
    rhs = 0.5 * np.einsum("ib,ijb,jb->b", v, Mdot, v)
    assert np.allclose(lhs, rhs, atol=1e-6)


def test_float32_oracle_runs():
    mech = rbd.load_model("iiwa14")
    o = Oracle(mech.flatten())
    q, v, tau, vd, _ = rand_inputs(mech, 8, 2)
    a64 = o.dynamics(q, v, tau)
    a32 = o.dynamics(q, v, tau, dtype=np.float32)
    assert a32.dtype == np.float32 and rel_err(a32, a64) < 1e-3


# ----------------------------------------------------------------------------------------------------------------------------
# Further identities of the reference's test-suite restated against the oracle (they involve no stored numbers, so they pin the
# oracle on mechanisms of any size).  tests/test_gpu_parity.py repeats (a), (b) and the closed form through the GPU path.
# ----------------------------------------------------------------------------------------------------------------------------
def _hat(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])


def momentum_balance_residual(mech, q, v, vd, wext, tau, kin):
    """test/test_mechanism_algorithms.jl:707-727 for one sample: floating-joint wrench (moved to the root frame) + gravity wrench
    + sum of external wrenches - (A v̇ + momentum_rate_bias).  `kin` = oracle-style kinematics dict of that sample."""
    nv = v.shape[0]
    T = kin["transforms"][:12]                                  # body of tree joint 0 = the floating base
    R, p = T[:9].reshape(3, 3), T[9:12]
    n_b, f_b = tau[:3], tau[3:6]                                # QuaternionFloating: tau = body-frame wrench [torque; force]
    f_w = R @ f_b
    n_w = R @ n_b + np.cross(p, f_w)
    mg = mech.mass() * mech.gravitational_acceleration
    grav = np.concatenate([np.cross(kin["com"], mg), mg])
    ext = wext.reshape(-1, 6).sum(0)                            # already in the root frame
    hdot = kin["A"].reshape(nv, 6).T @ vd + kin["mrb"]
    return np.concatenate([n_w, f_w]) + grav + ext - hdot


@pytest.mark.parametrize("seed", [39, 52])
def test_external_wrench_momentum_balance(seed):
    rng = np.random.default_rng(seed)
    mech = rbd.rand_floating_tree_mechanism(rng, [rbd.Revolute] * 10 + [rbd.Planar] * 10 + [rbd.SinCosRevolute] * 5)
    o = Oracle(mech.flatten())
    B = 5
    q, v, _, vd, w = rand_inputs(mech, B, seed, wext=True)
    tau = o.inverse_dynamics(q, v, vd, w)
    kin = o.kinematics(q, v, None, want=("transforms", "com", "mrb", "A"))
    for b in range(B):
        r = momentum_balance_residual(mech, q[:, b], v[:, b], vd[:, b], w[:, b], tau[:, b], {k: a[:, b] for k, a in kin.items()})
        assert np.abs(r).max() < 1e-9


def body_twists(o, mech, q, v):
    """twist_wrt_world of every non-root body = geometric Jacobian of the path root -> body times v; [nb, 6, B]."""
    out = []
    for body in (j.successor for j in mech.joints):
        sign = rbd.path(mech, mech.root_body, body).sign
        J = o.kinematics(q, v, sign, want=("J",))["J"].reshape(o.nv, 6, -1)
        out.append(np.einsum("kcb,kb->cb", J, v))
    return np.stack(out)


@pytest.mark.parametrize("seed", [43, 44])
def test_power_flow(seed):
    """:773-798: tau . v + sum_b w_ext,b . twist_b == d/dt (kinetic + gravitational potential energy) along the solution."""
    mech = randmech(seed)
    o = Oracle(mech.flatten())
    B = 4
    q, v, tau, _, w = rand_inputs(mech, B, seed, wext=True)
    vd, qd = o.dynamics(q, v, tau, w, want_qd=True)
    tw = body_twists(o, mech, q, v)
    power = np.einsum("ib,ib->b", tau, v) + np.einsum("ncb,ncb->b", w.reshape(-1, 6, B), tw)
    h = 1e-6

    def energy(qq, vv):
        k = o.kinematics(qq, vv, None, want=("ke", "pe"))
        return (k["ke"] + k["pe"]).ravel()
    dE = (energy(q + h * qd, v + h * vd) - energy(q - h * qd, v - h * vd)) / (2 * h)
    assert np.allclose(power, dE, rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("seed", [45, 46])
def test_mass_matrix_rate_minus_two_coriolis_is_skew_on_random_trees(seed):
    """:616-652 on the 25-joint trees with every joint type: v'(Ṁ - 2C)v = 0, i.e. v . c = 1/2 v' Ṁ v without gravity, with
    Ṁ the derivative of M along q̇ = N(q) v."""
    mech = randmech(seed, shuffle=True)
    mech.gravitational_acceleration[:] = 0
    o = Oracle(mech.flatten())
    q, v, tau, _, _ = rand_inputs(mech, 3, seed)
    _, qd = o.dynamics(q, v, tau, want_qd=True)
    M = lambda qq: o.mass_matrix(qq).reshape(o.nv, o.nv, -1).transpose(1, 0, 2)
    h = 1e-6
    Mdot = (M(q + h * qd) - M(q - h * qd)) / (2 * h)
    lhs = np.einsum("ib,ib->b", v, o.dynamics_bias(q, v))
    rhs = 0.5 * np.einsum("ib,ijb,jb->b", v, Mdot, v)
    assert np.allclose(lhs, rhs, rtol=1e-6, atol=1e-5)


def free_body_closed_form(J, c, m, g, quat, v):
    """Newton-Euler equations of ONE free rigid body written out with 3-vectors (textbook form, independent of the oracle's
    spatial algebra): body-frame twist v = [w; u] of the body origin, inertia J about the origin, centre of mass c.
        m (u̇ + w x u + ẇ x c + w x (w x c)) = m R' g
        J ẇ + w x J w + m c x (u̇ + w x u) = m c x R' g
    Returns v̇ = [ẇ; u̇] -- what dynamics! must give for a QuaternionFloating joint with zero torque."""
    w_, x, y, z = quat
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w_ * z), 2 * (x * z + w_ * y)],
                  [2 * (x * y + w_ * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w_ * x)],
                  [2 * (x * z - w_ * y), 2 * (y * z + w_ * x), 1 - 2 * (x * x + y * y)]])
    w, u = v[:3], v[3:]
    gb = R.T @ g
    # unknowns [ẇ; u̇]:  [[J, m ĉ], [-m ĉ, m 1]] [ẇ; u̇] = rhs
    A = np.block([[J, m * _hat(c)], [-m * _hat(c), m * np.eye(3)]])
    rhs = np.concatenate([m * np.cross(c, gb) - np.cross(w, J @ w) - m * np.cross(c, np.cross(w, u)),
                          m * gb - m * np.cross(w, u) - m * np.cross(w, np.cross(w, c))])
    return np.linalg.solve(A, rhs)


def free_body_mechanism(rng, com_offset=True):
    mech = rbd.Mechanism(rbd.RigidBody("world"), gravity=(0.3, -0.2, -9.81))
    Jc = rng.random((3, 3)); Jc = Jc @ Jc.T + np.eye(3)            # inertia about the centre of mass
    c = rng.standard_normal(3) * (0.3 if com_offset else 0.0)
    m = 2.5
    J = Jc + m * (c @ c * np.eye(3) - np.outer(c, c))              # parallel axes: about the frame origin
    mech.attach(mech.root_body, rbd.RigidBody("body", rbd.SpatialInertia(J, m * c, m)), rbd.Joint("free", rbd.QuaternionFloating()))
    return mech, J, c, m


@pytest.mark.parametrize("com_offset", [False, True])
def test_free_rigid_body_closed_form(com_offset):
    """Euler's equations (com at the origin) / the general Newton-Euler form (com offset) for a single QuaternionFloating body:
    pins the floating-joint conventions (body-frame twist, world-frame gravity) that the planar pendulum cannot."""
    rng = np.random.default_rng(7)
    mech, J, c, m = free_body_mechanism(rng, com_offset)
    o = Oracle(mech.flatten())
    q, v, _, _, _ = rand_inputs(mech, 6, 8)
    got = o.dynamics(q, v, None)
    for b in range(6):
        ref = free_body_closed_form(J, c, m, mech.gravitational_acceleration, q[:4, b], v[:, b])
        assert np.allclose(got[:, b], ref, rtol=0, atol=1e-11)
        if not com_offset:                                           # Euler's equations proper
            assert np.allclose(J @ got[:3, b] + np.cross(v[:3, b], J @ v[:3, b]), 0, atol=1e-11)
