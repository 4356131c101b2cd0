"""Times inverse_dynamics! (with / without external wrenches) on Atlas and the 7-DoF arm, fp32, batch 2^20 (kernel time, CUDA events)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rigidbodydynamics.jl_b200 as rbd  # noqa: E402

for name, fl in (("atlas", True), ("iiwa14", False)):
    m = rbd.load_model(name, floating=fl)
    B = 1 << 20
    st = rbd.MechanismState(m, B, torch.float32)
    rbd.rand_(st, np.random.default_rng(1))
    vd = torch.rand((st.nv, B), dtype=torch.float32, device="cuda")
    out = torch.empty_like(vd)
    w = torch.rand((6 * len(m.joints), B), dtype=torch.float32, device="cuda")
    for wx in (None, w):
        for _ in range(3):
            rbd.inverse_dynamics_(out, st, vd, wx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            rbd.inverse_dynamics_(out, st, vd, wx)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(name, "inverse_dynamics", "with wext" if wx is not None else "", f"{B / ms / 1e3:.1f} M evals/s")
