"""Workload for compute-sanitizer (profiles/sanitizer_*.txt): every kernel family once, at batch sizes that take the production
paths (shared-memory + Tensor-Memory kernel pair at 2^16; single kernels at 257)."""
import numpy as np
import torch
import rigidbodydynamics.jl_b200 as rbd

rng = np.random.default_rng(0)
mech = rbd.load_model("atlas", floating=True)
for dtype in (torch.float32, torch.float64):
    for B in (257, 1 << 16):
        st = rbd.MechanismState(mech, B, dtype)
        rbd.rand_(st, rng)
        tau = torch.rand((st.nv, B), dtype=dtype, device="cuda")
        wext = torch.rand((6 * len(mech.joints), B), dtype=dtype, device="cuda")
        res = rbd.DynamicsResult(mech, B, dtype)
        rbd.dynamics_(res, st, tau)
        rbd.dynamics_(res, st, tau, wext)
        out = torch.empty_like(tau)
        rbd.inverse_dynamics_(out, st, tau)
        rbd.inverse_dynamics_(out, st, tau, wext)
        rbd.dynamics_bias_(out, st)
        if B == 257:
            rbd.mass_matrix(st)
            p = rbd.path(mech, mech.findbody("r_foot"), mech.findbody("l_hand"))
            outs = {k: torch.empty((r, B), dtype=dtype, device="cuda") for k, r in
                    (("transforms_to_root", 12 * len(mech.joints)), ("center_of_mass", 3), ("kinetic_energy", 1),
                     ("gravitational_potential_energy", 1), ("momentum", 6), ("momentum_rate_bias", 6),
                     ("momentum_matrix", 6 * st.nv), ("geometric_jacobian", 6 * st.nv))}
            rbd.kinematics_(st, p, **outs)
            rbd.simulate_(st, 2e-4, tau, dt=1e-4)
        torch.cuda.synchronize()
        print("ok", dtype, B, rbd.launch_info().kernels_launched)
