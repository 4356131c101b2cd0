#!/usr/bin/env python
"""Runs ONE entry point twice (first call warms up / loads the module, the second is the one ncu captures with `-s 1 -c 1` on the
kernel-name regex).  usage: prof_family.py <what> [log2 batch]
  what: aba_f32 | aba_f64 | id_f32 | bias_f32 | crba_f32 | crba_lower_f32 | kin_A | kin_J | kin_small | dual | rk4 | bodies | aba_ext | iiwa_id | iiwa_crba"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rigidbodydynamics.jl_b200 as rbd

what = sys.argv[1]
B = 1 << int(sys.argv[2] if len(sys.argv) > 2 else 18)
rng = np.random.default_rng(1)
name = "iiwa14" if what.startswith("iiwa") else "atlas"
mech = rbd.load_model(name, floating=(name == "atlas"))
dt = torch.float64 if what.endswith("f64") else torch.float32
if what == "dual":
    dt = torch.float64
st = rbd.MechanismState(mech, B, dt)
rbd.rand_(st, rng)
nv, nq, nb = st.nv, st.nq, len(mech.joints)
x = torch.rand((nv, B), dtype=dt, device="cuda")
res = rbd.DynamicsResult(mech, B, dt)
out = torch.empty_like(x)


def call():
    if what in ("aba_f32", "aba_f64"):
        rbd.dynamics_(res, st, x, want_qd=False)
    elif what in ("id_f32", "iiwa_id"):
        rbd.inverse_dynamics_(out, st, x)
    elif what == "bias_f32":
        rbd.dynamics_bias_(out, st)
    elif what in ("crba_f32", "iiwa_crba"):
        rbd.mass_matrix_(call.M, st)
    elif what == "crba_lower_f32":
        rbd.mass_matrix_(call.M, st, uplo="L")
    elif what == "kin_A":
        rbd.momentum_matrix_(call.A, st)
    elif what == "kin_J":
        rbd.geometric_jacobian_(call.A, st, call.pth)
    elif what == "kin_small":
        rbd.kinematics_(st, None, **call.small)
    elif what == "dual":
        rbd.dynamics_dual_(call.od, call.std, call.qd, call.vdual, call.tdual)
    elif what == "rk4":
        rbd.simulate_(st, 1e-4, x, dt=1e-4)
    elif what == "bodies":
        rbd.inverse_dynamics_(out, st, x, jointwrenchesout=call.jw, accelerations=call.acc)
    elif what == "aba_ext":
        rbd.dynamics_(res, st, x, call.wext, want_qd=False)


if what.startswith("crba") or what == "iiwa_crba":
    call.M = torch.empty((nv * nv, B), dtype=dt, device="cuda")
if what in ("kin_A", "kin_J"):
    call.A = torch.empty((6 * nv, B), dtype=dt, device="cuda")
    call.pth = rbd.path(mech, mech.findbody("r_foot"), mech.findbody("l_hand"))
if what == "kin_small":
    call.small = {k: torch.empty((r, B), dtype=dt, device="cuda") for k, r in
                  (("center_of_mass", 3), ("kinetic_energy", 1), ("gravitational_potential_energy", 1), ("momentum", 6), ("momentum_rate_bias", 6))}
if what == "dual":
    call.std = rbd.MechanismState(mech, 1, torch.float64)
    call.qd = torch.zeros((nq, B, 7), dtype=torch.float64, device="cuda")
    call.qd[..., 0] = st.q
    call.qd[4:, :, 1:] = torch.rand((nq - 4, B, 6), dtype=torch.float64, device="cuda")
    call.vdual = torch.rand((nv, B, 7), dtype=torch.float64, device="cuda")
    call.tdual = torch.rand((nv, B, 7), dtype=torch.float64, device="cuda")
    call.od = torch.empty((nv, B, 7), dtype=torch.float64, device="cuda")
if what == "bodies":
    call.jw = torch.empty((6 * nb, B), dtype=dt, device="cuda"); call.acc = torch.empty_like(call.jw)
if what == "aba_ext":
    call.wext = torch.rand((6 * nb, B), dtype=dt, device="cuda")
for _ in range(2):
    call()
torch.cuda.synchronize()
print(what, B, "specialised", rbd.launch_info().specialised, "kernels", rbd.launch_info().kernels_launched)
