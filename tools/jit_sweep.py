#!/usr/bin/env python
"""Throughput of the model-specialised kernels vs. program size: revolute chains of n joints (fp32 forward dynamics, batch 2^20).
Run once with RBD_JIT=1 and once with RBD_JIT=0 (generic kernels) -- the environment is read when the library loads.
Prints one JSON line per n with the generated program's size, so the instruction-cache cliff is visible."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rigidbodydynamics.jl_b200 as rbd
from tests import hostsim

B = 1 << 20
for n in [int(x) for x in (sys.argv[1:] or ["4", "7", "10", "14", "18", "24", "31"])]:
    rng = np.random.default_rng(n)
    mech = rbd.rand_chain_mechanism(rng, [rbd.Revolute] * n)
    _, stats = hostsim.spec_source(mech.flatten(), "aba", np.float32, True, False, 1)
    st = rbd.MechanismState(mech, B, torch.float32)
    rbd.rand_(st, rng)
    tau = torch.rand((n, B), dtype=torch.float32, device="cuda")
    res = rbd.DynamicsResult(mech, B, torch.float32)
    fn = lambda: rbd.dynamics_(res, st, tau, want_qd=False)
    fn(); torch.cuda.synchronize()
    info = rbd.launch_info()
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"n": n, "jit": os.environ.get("RBD_JIT", "1"), "specialised": info.specialised, "kernels": info.kernels_launched,
                      "nodes_live": stats["nodes_live"], "ms": round(ms, 4), "Mevals_s": round(B / ms / 1e3, 1),
                      "Gnodes_s": round(B / ms / 1e6 * stats["nodes_live"] / 1e3, 1)}), flush=True)
