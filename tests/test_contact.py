"""Soft point contact: CPU tier + GPU parity.

The oracle's restatement of contact_dynamics! (oracle/rbd_oracle.cpp, mechanism_algorithms.jl:680-723 + contact.jl) is pinned by
  * the reference's HalfSpace3D test                          test/test_contact.jl:2-16
  * an independent closed form for a single free body (numpy, written from the formulas in contact.jl)
  * the reference's two simulation tests, re-run on the oracle with a plain RK4 driver:
      "elastic ball drop"  energy balance + bounces             test/test_simulate.jl:34-89
      "inclined plane"     stick above / slip below mu_crit     test/test_simulate.jl:91-125
Then the device code (csrc/rbd_kin.cuh compiled for the host) and the CUDA kernel through the C ABI must agree with it.
"""
import numpy as np
import pytest

import rigidbodydynamics.jl_b200 as rbd
from oracle import Oracle
from tests import hostsim
from tests.util import rand_inputs, randmech


def _free_body(inertia=None, rng=None):
    mech = rbd.Mechanism(rbd.RigidBody("world"))
    body = rbd.RigidBody("body", inertia if inertia is not None else rbd.SpatialInertia.rand(rng))
    mech.attach(mech.root_body, body, rbd.Joint("floating", rbd.QuaternionFloating()))
    return mech, body


def _quat_rot(w, x, y, z):
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _with_contacts(mech, seed, npoints=5, nhalf=2, scale=0.3):
    """Random contact points on random bodies and half-spaces that cut through the mechanism, so that a good share of the
    (point, half-space) pairs is in contact and a good share is not."""
    rng = np.random.default_rng(seed)
    for _ in range(npoints):
        body = mech.joints[int(rng.integers(len(mech.joints)))].successor
        model = rbd.SoftContactModel(rbd.hunt_crossley_hertz(k=float(rng.uniform(1e2, 1e3)), alpha=float(rng.uniform(0, 0.5))),
                                     rbd.ViscoelasticCoulombModel(float(rng.uniform(0.2, 1.5)), float(rng.uniform(10, 100)),
                                                                  float(rng.uniform(5, 50))))
        rbd.add_contact_point(body, rbd.ContactPoint(rng.standard_normal(3) * scale, model))
    for _ in range(nhalf):
        rbd.add_environment_primitive(mech, rbd.HalfSpace3D(rng.standard_normal(3) * 0.2, rng.standard_normal(3)))
    return rbd.contact_desc(mech)


# ------------------------------------------------------------------------------------------------------------------
# oracle pinning
# ------------------------------------------------------------------------------------------------------------------
def test_halfspace_separation_like_reference():
    """test/test_contact.jl:2-16: inside <=> below the plane; the outward normal is the gradient of the separation."""
    rng = np.random.default_rng(4)
    mech, body = _free_body(rng=rng)
    pt = rng.random(3)
    model = rbd.SoftContactModel(rbd.HuntCrossleyModel(1.0, 0.0, 1.0), rbd.ViscoelasticCoulombModel(0.0, 0.0, 1.0))
    rbd.add_contact_point(body, rbd.ContactPoint(np.zeros(3), model))
    rbd.add_environment_primitive(mech, rbd.HalfSpace3D(pt, [0, 0, 2.0]))        # normalised on construction
    cd = rbd.contact_desc(mech)
    assert np.allclose(cd.halfspace[0, 3:], [0, 0, 1])
    orc = Oracle(mech.flatten())
    B = 100
    q = np.zeros((7, B)); q[0] = 1; q[4:] = rng.standard_normal((3, B))
    wr, sd, s = orc.contact_dynamics(q, np.zeros((6, B)), cd)
    # k = 1, n = 1, no damping, no friction: force = penetration * normal, so the force IS the (negative) separation when inside
    sep = q[6] - pt[2]
    assert np.array_equal(wr[5] > 0, sep < 0)
    assert np.allclose(wr[5], np.maximum(-sep, 0), atol=1e-15)
    assert np.allclose(wr[3:5], 0)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_free_body_closed_form(seed):
    """Everything written out by hand for one floating body: T = (R(quat), p), point velocity = R (omega x l + v)."""
    rng = np.random.default_rng(seed)
    mech, body = _free_body(rng=rng)
    cd = _with_contacts(mech, seed + 10, npoints=3, nhalf=2, scale=0.5)
    orc = Oracle(mech.flatten())
    B = 64
    q, v, _, _, _ = rand_inputs(mech, B, seed)
    q[4:] *= 0.3
    s0 = rng.standard_normal((cd.nstates, B)) * 0.05
    wr, sd, s1 = orc.contact_dynamics(q, v, cd, s0)
    ncontact = 0
    for b in range(B):
        R = _quat_rot(*q[:4, b]); p = q[4:, b]
        w_exp = np.zeros(6)
        for i in range(cd.npoints):
            pt = R @ cd.location[i] + p
            vel = R @ (np.cross(v[:3, b], cd.location[i]) + v[3:, b])
            k, lam, n = cd.normal_model[i]; mu, kf, bf = cd.friction_model[i]
            for h in range(cd.nhalfspaces):
                hp, hn = cd.halfspace[h, :3], cd.halfspace[h, 3:]
                row = 3 * (i * cd.nhalfspaces + h)
                sep = (pt - hp) @ hn
                if sep <= 0:
                    ncontact += 1
                    z, zd = -sep, -(vel @ hn)
                    fn = max(lam * z ** n * zd + k * z ** n, 0.0)
                    x = s0[row:row + 3, b]
                    ft = -kf * x - bf * (vel + zd * hn)
                    if ft @ ft > (mu * fn) ** 2:
                        ft = ft * np.sqrt((mu * fn) ** 2 / (ft @ ft))
                    assert np.linalg.norm(ft) <= mu * fn * (1 + 1e-12) + 1e-15            # inside the friction cone
                    f = fn * hn + ft
                    w_exp += np.concatenate([np.cross(pt, f), f])
                    assert np.allclose(sd[row:row + 3, b], (-kf * x - ft) / bf, rtol=1e-12, atol=1e-12)
                    assert np.array_equal(s1[row:row + 3, b], s0[row:row + 3, b])
                else:
                    assert np.all(s1[row:row + 3, b] == 0) and np.all(sd[row:row + 3, b] == 0)    # reset! / zero!
        assert np.allclose(wr[:, b], w_exp, rtol=1e-12, atol=1e-10)
    assert 0.15 * B * cd.npoints * cd.nhalfspaces < ncontact < 0.85 * B * cd.npoints * cd.nhalfspaces


def _simulate_oracle(mech, cd, q, v, t_final, dt, record=None, s=None):
    """simulate(state, t_final; dt) on the oracle: classic RK4 on (q, v, s) with q̇ from configuration_derivative and the
    quaternion renormalised after each step (the reference uses the Munthe-Kaas form of the same tableau; for the tests below
    the body does not rotate, so the two coincide)."""
    orc = Oracle(mech.flatten())
    s = np.zeros((cd.nstates, 1)) if s is None else s          # the additional state lives in the MechanismState

    def f(q, v, s):
        wr, sd, _ = orc.contact_dynamics(q, v, cd, s)
        vd, qd = orc.dynamics(q, v, None, wr, want_qd=True)
        return qd, vd, sd

    n = int(round(t_final / dt))
    for step in range(n):
        if record is not None:
            record(step * dt, q, v)
        _, _, s = orc.contact_dynamics(q, v, cd, s)          # the resets of pairs that left contact stick to the state
        k1 = f(q, v, s)
        k2 = f(q + 0.5 * dt * k1[0], v + 0.5 * dt * k1[1], s + 0.5 * dt * k1[2])
        k3 = f(q + 0.5 * dt * k2[0], v + 0.5 * dt * k2[1], s + 0.5 * dt * k2[2])
        k4 = f(q + dt * k3[0], v + dt * k3[1], s + dt * k3[2])
        q = q + dt / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
        v = v + dt / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
        s = s + dt / 6 * (k1[2] + 2 * k2[2] + 2 * k3[2] + k4[2])
        q[:4] /= np.linalg.norm(q[:4])
    return q, v, s


def test_elastic_ball_drop_energy_balance():
    """test/test_simulate.jl:34-89: a body with a contact point at its centre of mass dropped on the floor with a conservative
    normal model (alpha = 0): kinetic + gravitational + elastic energy stays constant (atol 1e-2) and the ball bounces."""
    rng = np.random.default_rng(61)
    mech, body = _free_body(rng=rng)
    com = body.inertia.cross_part / body.inertia.mass
    model = rbd.SoftContactModel(rbd.hunt_crossley_hertz(alpha=0.0), rbd.ViscoelasticCoulombModel(0.5, 1e3, 1e3))
    rbd.add_contact_point(body, rbd.ContactPoint(com, model))
    rbd.add_environment_primitive(mech, rbd.HalfSpace3D(np.zeros(3), [0, 0, 1.0]))
    cd = rbd.contact_desc(mech)
    orc = Oracle(mech.flatten())
    z0 = 0.05
    q = np.zeros((7, 1)); q[0] = 1; q[4:, 0] = [1.0, 2.0, z0 - com[2]]
    v = np.zeros((6, 1))
    energies, vz = [], []

    def record(t, q, v):
        k = orc.kinematics(q, v, want=("com", "ke", "pe"))
        pen = max(-k["com"][2, 0], 0.0)
        n = model.normal.n
        energies.append(model.normal.k * pen ** (n + 1) / (n + 1) + k["ke"][0, 0] + k["pe"][0, 0])
        vz.append(v[5, 0])           # no rotation: the linear velocity of the body frame is the point velocity

    _simulate_oracle(mech, cd, q, v, 0.5, 1e-3, record)
    energies = np.asarray(energies)
    assert np.abs(energies - energies[0]).max() < 1e-2
    sg = np.sign(vz)
    assert np.count_nonzero(sg[1:] != sg[:-1]) > 3


@pytest.mark.parametrize("stick", [True, False])
def test_inclined_plane_stick_slip(stick):
    """test/test_simulate.jl:91-125: a point mass on a plane inclined by theta sticks for mu > tan(theta) and slides below; a
    second, irrelevant half-space far below must not matter (#211)."""
    theta = 0.5
    mu = np.tan(theta) + (1e-2 if stick else -1e-2)
    mech, body = _free_body(rbd.SpatialInertia(np.eye(3), np.zeros(3), 2.0))
    rbd.add_environment_primitive(mech, rbd.HalfSpace3D(np.zeros(3), [np.sin(theta), 0, np.cos(theta)]))
    rbd.add_environment_primitive(mech, rbd.HalfSpace3D([0, 0, -100.0], [0, 0, 1.0]))
    model = rbd.SoftContactModel(rbd.hunt_crossley_hertz(k=50e3, alpha=1.0), rbd.ViscoelasticCoulombModel(mu, 50e3, 1e4))
    rbd.add_contact_point(body, rbd.ContactPoint(np.zeros(3), model))
    cd = rbd.contact_desc(mech)
    q = np.zeros((7, 1)); q[0] = 1
    v = np.zeros((6, 1))
    q, v, s = _simulate_oracle(mech, cd, q, v, 1.0, 1e-3)       # settle
    x1 = q[4:, 0].copy()
    q, v, s = _simulate_oracle(mech, cd, q, v, 0.5, 1e-3, s=s)
    x2 = q[4:, 0]
    if stick:
        assert np.allclose(x1, x2, atol=1e-4)
    else:
        assert not np.allclose(x1, x2, atol=5e-2)


# ------------------------------------------------------------------------------------------------------------------
# device code on the CPU vs the oracle
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [32, 33, 35, 38])
def test_hostsim_contact_matches_oracle(seed):
    mech = randmech(seed)
    cd = _with_contacts(mech, seed, npoints=7, nhalf=3)
    desc = mech.flatten()
    q, v, _, _, _ = rand_inputs(mech, 24, seed)
    s0 = np.random.default_rng(seed).standard_normal((cd.nstates, 24)) * 0.05
    wr_o, sd_o, s_o = Oracle(desc).contact_dynamics(q, v, cd, s0)
    wr, sd, s = hostsim.contact(desc, q, v, cd, s0)
    sc = max(1.0, np.abs(wr_o).max())
    assert np.abs(wr_o).max() > 0 and np.any(s_o != s0)
    assert np.abs(wr - wr_o).max() < 1e-11 * sc
    assert np.abs(sd - sd_o).max() < 1e-11 * max(1.0, np.abs(sd_o).max())
    assert np.array_equal(s, s_o)
    # fp32 device code against the fp64 oracle
    wr32, sd32, _ = hostsim.contact(desc, q.astype(np.float32), v.astype(np.float32), cd, s0.astype(np.float32))
    assert np.abs(wr32 - wr_o).max() < 2e-4 * sc


def test_contact_desc_state_layout():
    """num_additional_states = 3 per (contact point, half-space), body / point / half-space order (mechanism.jl:143-149)."""
    mech = randmech(33)
    assert rbd.num_contact_states(mech) == 0
    cd = _with_contacts(mech, 1, npoints=4, nhalf=2)
    assert rbd.num_contact_states(mech) == 24 == cd.nstates
    assert np.all(np.diff(cd.body) >= 0)


# ------------------------------------------------------------------------------------------------------------------
# CUDA kernel through the C ABI vs the oracle
# ------------------------------------------------------------------------------------------------------------------
def _gpu_state(mech, q, v, dtype):
    import torch
    st = rbd.MechanismState(mech, batch=q.shape[1], dtype=dtype)
    st.q.copy_(torch.from_numpy(q).to(dtype)); st.v.copy_(torch.from_numpy(v).to(dtype))
    return st


@pytest.mark.gpu
@pytest.mark.parametrize("seed,dtype_name", [(32, "float64"), (35, "float64"), (38, "float64"), (33, "float32"), (38, "float32")])
def test_gpu_contact_matches_oracle(seed, dtype_name):
    import torch
    dtype = getattr(torch, dtype_name)
    mech = randmech(seed)
    cd = _with_contacts(mech, seed, npoints=9, nhalf=3)
    B = 777                                                        # ragged: not a multiple of the block size
    q, v, _, _, _ = rand_inputs(mech, B, seed)
    s0 = np.random.default_rng(seed).standard_normal((cd.nstates, B)) * 0.05
    if dtype == torch.float32:
        q = q.astype(np.float32).astype(np.float64); v = v.astype(np.float32).astype(np.float64)
        s0 = s0.astype(np.float32).astype(np.float64)
    wr_o, sd_o, s_o = Oracle(mech.flatten()).contact_dynamics(q, v, cd, s0)
    st = _gpu_state(mech, q, v, dtype)
    s = torch.from_numpy(s0).to(dtype).cuda()
    sd = torch.full_like(s, float("nan")); wr = torch.full((6 * len(mech.joints), B), float("nan"), dtype=dtype, device="cuda")
    rbd.contact_dynamics_(st, wr, s, sd)
    assert rbd.launch_info().kernels_launched == 1
    tol = 1e-11 if dtype == torch.float64 else 2e-4
    sc = max(1.0, np.abs(wr_o).max())
    assert np.abs(wr.cpu().numpy() - wr_o).max() < tol * sc
    assert np.abs(sd.cpu().numpy() - sd_o).max() < tol * max(1.0, np.abs(sd_o).max())
    got_s = s.cpu().numpy()
    if dtype == torch.float64:
        assert np.array_equal(got_s, s_o)
    else:       # a pair exactly at the boundary may fall on the other side in fp32; everything else is copied or zeroed exactly
        assert np.mean(got_s != s_o) < 1e-3
    # state = None means zeros
    wr2 = torch.empty_like(wr)
    rbd.contact_dynamics_(st, wr2)
    wr_z, _, _ = Oracle(mech.flatten()).contact_dynamics(q, v, cd, None)
    assert np.abs(wr2.cpu().numpy() - wr_z).max() < tol * max(1.0, np.abs(wr_z).max())


@pytest.mark.gpu
def test_gpu_dynamics_with_contact_atlas():
    """dynamics! on a mechanism with contact points (mechanism_algorithms.jl:845-866): Atlas with four contact points per foot
    standing in / above a floor, plus caller-supplied external wrenches; compared with the oracle's contact_dynamics + dynamics."""
    import torch
    mech = rbd.load_model("atlas", floating=True)
    model = rbd.SoftContactModel(rbd.hunt_crossley_hertz(), rbd.ViscoelasticCoulombModel(0.8, 20e3, 100.0))
    for foot in ("l_foot", "r_foot"):
        body = mech.findbody(foot)
        for x in (-0.08, 0.17):
            for y in (-0.06, 0.06):
                rbd.add_contact_point(body, rbd.ContactPoint(np.array([x, y, -0.08]), model))
    rbd.add_environment_primitive(mech, rbd.HalfSpace3D(np.zeros(3), [0, 0, 1.0]))
    cd = rbd.contact_desc(mech)
    assert cd.npoints == 8 and cd.nstates == 24
    B = 4096
    rng = np.random.default_rng(7)
    q, v, tau, _, wext = rand_inputs(mech, B, 7, wext=True)
    q[:4] = [[1.0], [0], [0], [0]] + 0.05 * rng.standard_normal((4, B)); q[:4] /= np.linalg.norm(q[:4], axis=0)
    q[4:6] = rng.standard_normal((2, B)); q[6] = 0.93 + 0.03 * rng.standard_normal(B)      # pelvis height: feet around z = 0
    q[7:] *= 0.1; v *= 0.2
    s0 = rng.standard_normal((cd.nstates, B)) * 1e-3
    orc = Oracle(mech.flatten())
    wr_o, sd_o, s_o = orc.contact_dynamics(q, v, cd, s0)
    frac = np.mean(np.abs(wr_o).reshape(-1, 6, B).sum(1) > 0)
    assert 0.005 < frac < 0.2, frac                                   # only the feet, and not all of them, touch
    vd_o = orc.dynamics(q, v, tau, wr_o + wext)
    st = _gpu_state(mech, q, v, torch.float64)
    res = rbd.DynamicsResult(mech, B, torch.float64)
    s = torch.from_numpy(s0).cuda(); sd = torch.empty_like(s)
    rbd.dynamics_contact_(res, st, torch.from_numpy(tau).cuda(), torch.from_numpy(wext).cuda(), s, sd)
    assert np.abs(res.contactwrenches.cpu().numpy() - wr_o).max() < 1e-10 * max(1.0, np.abs(wr_o).max())
    assert np.abs(res.vd.cpu().numpy() - vd_o).max() < 1e-8 * max(1.0, np.abs(vd_o).max())
    assert np.abs(sd.cpu().numpy() - sd_o).max() < 1e-10 * max(1.0, np.abs(sd_o).max())
    assert np.array_equal(s.cpu().numpy(), s_o)


@pytest.mark.gpu
def test_gpu_contact_errors():
    import torch
    mech = randmech(33)
    cd = _with_contacts(mech, 3, npoints=2, nhalf=1)
    q, v, _, _, _ = rand_inputs(mech, 8, 0)
    st = _gpu_state(mech, q, v, torch.float64)
    wr = torch.empty((6 * len(mech.joints), 8), dtype=torch.float64, device="cuda")
    with pytest.raises(rbd.DimensionMismatch):
        rbd.contact_dynamics_(st, wr[:, :4].contiguous())
    bad = rbd.ContactDesc(cd.body.copy(), cd.location, cd.normal_model, cd.friction_model, cd.halfspace)
    bad.body[0] = 99
    with pytest.raises(rbd.RbdError):
        rbd.contact_dynamics_(st, wr, contact=bad)
    many = rbd.ContactDesc(np.zeros(33, np.int32), np.zeros((33, 3)), np.ones((33, 3)), np.ones((33, 3)), cd.halfspace)
    with pytest.raises(rbd.RbdError) as e:
        rbd.contact_dynamics_(st, wr, contact=many)
    assert e.value.status == 6          # RBD_EUNSUPPORTED: fall back to the reference
    # empty batch and a mechanism without contact points are fine
    empty = rbd.ContactDesc(np.zeros(0, np.int32), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 6)))
    rbd.contact_dynamics_(st, wr, contact=empty)
    assert float(wr.abs().max()) == 0.0
