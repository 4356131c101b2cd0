"""Times rbd_dynamics_derivatives (analytic dv̇/dq, dv̇/dv) against the Dual{Float64,6} sweeps that produce the same Jacobians
(2 nv / 6 sweeps per sample), kernel time by CUDA events.  usage: time_deriv.py [model] [log2 batch ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rigidbodydynamics.jl_b200 as rbd  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "atlas"
logs = [int(a) for a in sys.argv[2:]] or [13, 15, 16]
m = rbd.load_model(name, floating=(name != "iiwa14"))
for dtype in (torch.float64, torch.float32):
    for lb in logs:
        B = 1 << lb
        st = rbd.MechanismState(m, B, dtype)
        rbd.rand_(st, np.random.default_rng(1))
        nv = st.nv
        tau = torch.rand((nv, B), dtype=dtype, device="cuda")
        res = rbd.DynamicsResult(m, B, dtype)
        dq = torch.empty((nv * nv, B), dtype=dtype, device="cuda")
        dv = torch.empty_like(dq)
        for _ in range(2):
            rbd.dynamics_derivatives_(dq, dv, res, st, tau)
        torch.cuda.synchronize()
        n = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            rbd.dynamics_derivatives_(dq, dv, res, st, tau)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        info = rbd.launch_info()
        out_bytes = 2 * nv * nv * B * dq.element_size()
        line = (f"{name} {str(dtype)[6:]} B=2^{lb}: {ms:.3f} ms  {B / ms / 1e3:.2f} M Jacobian pairs/s  "
                f"= {B / ms / 1e3 * (2 * nv / 6):.1f} M dual-sweep equivalents/s; output {out_bytes / ms / 1e6:.0f} GB/s; "
                f"{info.kernels_launched} launches")
        if dtype == torch.float64 and lb <= 13:
            Q = torch.zeros((st.nq, B, 7), dtype=torch.float64, device="cuda"); Q[..., 0] = st.q
            V = torch.zeros((nv, B, 7), dtype=torch.float64, device="cuda"); V[..., 0] = st.v
            T = torch.zeros((nv, B, 7), dtype=torch.float64, device="cuda"); T[..., 0] = tau
            out = torch.empty((nv, B, 7), dtype=torch.float64, device="cuda")
            rbd.dynamics_dual_(out, st, Q, V, T)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                rbd.dynamics_dual_(out, st, Q, V, T)
            e1.record()
            torch.cuda.synchronize()
            msd = e0.elapsed_time(e1) / n
            sweeps = -(-2 * nv // 6)
            line += f" | Dual sweep {msd:.3f} ms ({B / msd / 1e3:.1f} M sweeps/s) x {sweeps} sweeps = {msd * sweeps:.2f} ms -> speed-up {msd * sweeps / ms:.1f}x"
        print(line, flush=True)
        del dq, dv, res, st, tau
        torch.cuda.empty_cache()
