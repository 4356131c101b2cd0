"""Times mass_matrix! (full / lower) on Atlas (2^18) and the 7-DoF arm (2^20), fp32, kernel time by CUDA events; RBD_JIT=0 gives
the generic kernel for comparison."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rigidbodydynamics.jl_b200 as rbd  # noqa: E402

for name, fl, lb, dt in (("atlas", True, 18, torch.float32), ("iiwa14", False, 20, torch.float32), ("atlas", True, 17, torch.float64)):
    m = rbd.load_model(name, floating=fl)
    B = 1 << lb
    st = rbd.MechanismState(m, B, dt)
    rbd.rand_(st, np.random.default_rng(1))
    nv = st.nv
    M = torch.empty((nv * nv, B), dtype=dt, device="cuda")
    for uplo in ("full", "L"):
        for _ in range(3):
            rbd.mass_matrix_(M, st, uplo=uplo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            rbd.mass_matrix_(M, st, uplo=uplo)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        info = rbd.launch_info()
        out_b = (nv * nv if uplo == "full" else nv * (nv + 1) // 2) * M.element_size()
        print(f"{name} {str(dt)[6:]} mass_matrix {uplo}: {ms:.3f} ms  {B / ms / 1e3:.1f} M evals/s  output {B * out_b / ms / 1e6:.0f} GB/s  "
              f"specialised={info.specialised} launches={info.kernels_launched} grid={info.grid}x{info.block} bps={info.blocks_per_sm}", flush=True)
