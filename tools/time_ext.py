"""Times dynamics! / inverse_dynamics! WITH an external wrench on every body (the reference's own benchmark variant,
perf/runbenchmarks.jl:59-67) on Atlas, fp32, batch 2^20; RBD_JIT=0 gives the generic kernels for comparison."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rigidbodydynamics.jl_b200 as rbd  # noqa: E402

m = rbd.load_model("atlas", floating=True)
B = 1 << 20
st = rbd.MechanismState(m, B, torch.float32)
rbd.rand_(st, np.random.default_rng(1))
x = torch.rand((st.nv, B), dtype=torch.float32, device="cuda")
w = torch.rand((6 * len(m.joints), B), dtype=torch.float32, device="cuda")
res = rbd.DynamicsResult(m, B, torch.float32)
out = torch.empty_like(x)
for name, fn in (("dynamics+wext", lambda: rbd.dynamics_(res, st, x, w, want_qd=False)), ("inverse_dynamics+wext", lambda: rbd.inverse_dynamics_(out, st, x, w))):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    info = rbd.launch_info()
    print(f"atlas {name}: {ms:.3f} ms  {B / ms / 1e3:.1f} M evals/s  specialised={info.specialised} launches={info.kernels_launched} "
          f"grid={info.grid}x{info.block}", flush=True)
