"""CPU tier for the kinematics by-products.

1. The oracle's restatement is pinned by the same identities the reference's own tests use (no reference-produced vectors
   exist for these):
     double pendulum kinetic energy, closed form            test/test_double_pendulum.jl:51-52, 72-73   (atol 1e-12)
     1/2 v' M v == kinetic_energy                            test/test_mechanism_algorithms.jl:564-572   (1e-12)
     A(q) v == momentum == sum of body momenta               :527-546, :686                              (1e-12)
     d/dt (A v) == A v̇ + momentum_rate_bias                  :677-704 (ForwardDiff there, central differences here)
     d PE / d q == gravity term of inverse_dynamics          :654-675
     J(path) v == relative twist of target w.r.t. source     :310-344 (twists checked against finite-differenced transforms)
2. The device code (csrc/rbd_kin.cuh compiled for the host) must agree with that oracle for every joint type.
"""
import numpy as np
import pytest

import rigidbodydynamics.jl_b200 as rbd
from oracle import Oracle
from tests import hostsim
from tests.util import double_pendulum, rand_inputs, randmech


def _sign_to_body(mech, i):
    """path(mechanism, root, successor of joint i)"""
    return rbd.path(mech, mech.root_body, mech.joints[i].successor).sign


def test_double_pendulum_energy_closed_form():
    I1 = I2 = 0.333; lc1 = lc2 = -0.5; l1 = -1.0; m1 = m2 = 1.0; g = -9.81
    mech = double_pendulum(I1, I2, lc1, lc2, l1, m1, m2, g)
    q = np.array([[0.3], [0.4]]); v = np.array([[1.0], [2.0]])
    k = Oracle(mech.flatten()).kinematics(q, v)
    c2 = np.cos(q[1, 0])
    T1 = 0.5 * I1 * v[0, 0] ** 2
    T2 = (0.5 * (m2 * l1 ** 2 + I2 + 2 * m2 * l1 * lc2 * c2) * v[0, 0] ** 2 + 0.5 * I2 * v[1, 0] ** 2
          + (I2 + m2 * l1 * lc2 * c2) * v[0, 0] * v[1, 0])
    assert abs(k["ke"][0, 0] - (T1 + T2)) < 1e-12
    # potential energy from the geometry: link coms hang at z = lc1 cos q1 and l1 cos q1 + lc2 cos(q1 + q2)
    z1 = lc1 * np.cos(0.3); z2 = l1 * np.cos(0.3) + lc2 * np.cos(0.7)
    assert abs(k["pe"][0, 0] - (-g) * (m1 * z1 + m2 * z2)) < 1e-12
    assert np.allclose(k["com"][:, 0], [(m1 * lc1 * np.sin(0.3) + m2 * (l1 * np.sin(0.3) + lc2 * np.sin(0.7))) / 2, 0,
                                         (m1 * z1 + m2 * z2) / 2], atol=1e-12)


@pytest.mark.parametrize("seed", [32, 33, 38])
def test_oracle_momentum_and_energy_identities(seed):
    mech = randmech(seed)
    desc = mech.flatten()
    orc = Oracle(desc)
    q, v, _, vd, _ = rand_inputs(mech, 3, seed)
    k = orc.kinematics(q, v)
    nv = desc.nv
    A = k["A"].reshape(nv, 6, -1)                         # [k, c, b]
    Av = np.einsum("kcb,kb->cb", A, v)
    assert np.abs(Av - k["momentum"]).max() < 1e-11 * max(1, np.abs(Av).max())
    M = orc.mass_matrix(q).reshape(nv, nv, -1)
    ke = 0.5 * np.einsum("ib,ijb,jb->b", v, M, v)
    assert np.abs(ke - k["ke"][0]).max() < 1e-11 * max(1, ke.max())
    # potential energy and centre of mass from the transforms and the body inertias
    T = k["transforms"].reshape(desc.nb, 12, -1)
    mc = np.zeros((3, q.shape[1])); mass = 0.0
    for i in range(desc.nb):
        m = desc.inertia[i, 12]; c = desc.inertia[i, 9:12]       # cross_part = m * com
        R = T[i, :9].reshape(3, 3, -1)
        mc += np.einsum("ijb,j->ib", R, c) + m * T[i, 9:]
        mass += m
    assert np.abs(mc / mass - k["com"]).max() < 1e-12
    assert np.abs(-(desc.gravity @ mc) - k["pe"][0]).max() < 1e-10


def test_oracle_momentum_rate_and_gravity_term_by_central_differences():
    rng = np.random.default_rng(37)
    mech = rbd.rand_tree_mechanism(rng, [rbd.Revolute] * 10 + [rbd.Prismatic] * 10)     # q̇ = v for these joints
    desc = mech.flatten()
    orc = Oracle(desc)
    q, v, _, vd, _ = rand_inputs(mech, 2, 38)
    nv = desc.nv
    eps = 1e-6

    def Av(qq, vv):
        A = orc.kinematics(qq, vv, want=("A",))["A"].reshape(nv, 6, -1)
        return np.einsum("kcb,kb->cb", A, vv)
    hdot_fd = (Av(q + eps * v, v + eps * vd) - Av(q - eps * v, v - eps * vd)) / (2 * eps)
    k = orc.kinematics(q, v)
    hdot = np.einsum("kcb,kb->cb", k["A"].reshape(nv, 6, -1), vd) + k["mrb"]
    assert np.abs(hdot - hdot_fd).max() < 1e-6 * max(1, np.abs(hdot).max())
    # gravity term: d PE / d q = inverse_dynamics(q, v = 0, v̇ = 0)
    g = orc.inverse_dynamics(q, 0 * v, 0 * v)
    for j in range(nv):
        dq = np.zeros_like(q); dq[j] = eps
        dpe = (orc.kinematics(q + dq, None, want=("pe",))["pe"] - orc.kinematics(q - dq, None, want=("pe",))["pe"]) / (2 * eps)
        assert np.abs(dpe[0] - g[j]).max() < 1e-6 * max(1, np.abs(g).max())


def test_oracle_jacobian_twist_against_differenced_transforms():
    mech = randmech(25)
    desc = mech.flatten()
    orc = Oracle(desc)
    q, v, _, _, _ = rand_inputs(mech, 2, 25)
    qd = orc.dynamics(q, v, want_qd=True)[1]
    eps = 1e-6
    Tp = orc.kinematics(q + eps * qd, None, want=("transforms",))["transforms"].reshape(desc.nb, 12, -1)
    Tm = orc.kinematics(q - eps * qd, None, want=("transforms",))["transforms"].reshape(desc.nb, 12, -1)
    T0 = orc.kinematics(q, None, want=("transforms",))["transforms"].reshape(desc.nb, 12, -1)
    twists = []
    for i in range(desc.nb):
        J = orc.kinematics(q, v, _sign_to_body(mech, i), want=("J",))["J"].reshape(desc.nv, 6, -1)
        tw = np.einsum("kcb,kb->cb", J, v)               # twist of body i w.r.t. the world, root frame
        twists.append(tw)
        Rdot = ((Tp[i, :9] - Tm[i, :9]) / (2 * eps)).reshape(3, 3, -1)
        pdot = (Tp[i, 9:] - Tm[i, 9:]) / (2 * eps)
        R = T0[i, :9].reshape(3, 3, -1)
        W = np.einsum("ijb,kjb->ikb", Rdot, R)           # Rdot R^T = hat(omega)
        omega = np.stack([W[2, 1], W[0, 2], W[1, 0]])
        assert np.abs(omega - tw[:3]).max() < 1e-6 * max(1, np.abs(tw).max())
        vo = pdot - np.cross(omega.T, T0[i, 9:].T).T     # velocity of the body-fixed point at the root origin
        assert np.abs(vo - tw[3:]).max() < 1e-6 * max(1, np.abs(tw).max())
    # path between two arbitrary bodies: J v = twist(target) - twist(source)     (relative_twist, :310-326)
    rng = np.random.default_rng(5)
    for _ in range(20):
        a, b = rng.choice(desc.nb, 2, replace=False)
        p = rbd.path(mech, mech.joints[a].successor, mech.joints[b].successor)
        J = orc.kinematics(q, v, p.sign, want=("J",))["J"].reshape(desc.nv, 6, -1)
        rel = np.einsum("kcb,kb->cb", J, v)
        assert np.abs(rel - (twists[b] - twists[a])).max() < 1e-11 * max(1, np.abs(rel).max())


def test_zero_configuration_transforms_are_products_of_tree_transforms():
    mech = rbd.load_model("atlas", floating=True)
    desc = mech.flatten()
    q = mech.zero_configuration().reshape(-1, 1)
    T = Oracle(desc).kinematics(q, None, want=("transforms",))["transforms"].reshape(desc.nb, 12)
    H = {-1: np.eye(4)}
    for i in range(desc.nb):
        X = np.eye(4); X[:3, :3] = desc.X_tree[i, :9].reshape(3, 3); X[:3, 3] = desc.X_tree[i, 9:]
        H[i] = H[int(desc.parent[i])] @ X
        assert np.allclose(T[i, :9].reshape(3, 3), H[i][:3, :3], atol=1e-13)
        assert np.allclose(T[i, 9:], H[i][:3, 3], atol=1e-13)


@pytest.mark.parametrize("name,floating", [("atlas", True), ("valkyrie", False), ("iiwa14", False), ("double_pendulum", False)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_device_code_matches_oracle_named_models(name, floating, dt):
    mech = rbd.load_model(name, floating=floating)
    _compare(mech, 17, dt)


@pytest.mark.parametrize("seed", [17, 18, 19, 20])
def test_device_code_matches_oracle_all_joint_types(seed):
    _compare(randmech(seed, shuffle=seed % 2 == 1), seed, np.float64)


def _compare(mech, seed, dt):
    desc = mech.flatten()
    q, v, _, _, _ = rand_inputs(mech, 4, seed)
    rng = np.random.default_rng(seed)
    a, b = rng.choice(desc.nb, 2, replace=False) if desc.nb > 1 else (0, 0)
    sign = rbd.path(mech, mech.joints[a].successor, mech.joints[b].successor).sign if desc.nb > 1 else _sign_to_body(mech, 0)
    ref = Oracle(desc).kinematics(q, v, sign)
    got = hostsim.kinematics(desc, q.astype(dt), v.astype(dt), sign)
    tol = 1e-11 if dt == np.float64 else 2e-5
    assert set(got) == set(ref)
    for k in ref:
        err = np.abs(got[k] - ref[k]).max() / max(1.0, np.abs(ref[k]).max())
        assert err < tol, (k, err)
    # subsets: no velocity, single outputs
    sub = hostsim.kinematics(desc, q.astype(dt), None, None, want=("com", "transforms"))
    assert set(sub) == {"com", "transforms"}
    assert np.abs(sub["com"] - ref["com"]).max() < tol * max(1.0, np.abs(ref["com"]).max())
    only_a = hostsim.kinematics(desc, q.astype(dt), None, None, want=("A",))
    assert np.abs(only_a["A"] - ref["A"]).max() < tol * max(1.0, np.abs(ref["A"]).max())
