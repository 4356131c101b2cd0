"""Times the kinematics by-products bench.py reports (geometric jacobian foot -> hand, com + energies + momenta fused, momentum
matrix) on Atlas fp32 2^20, kernel time by CUDA events; RBD_JIT=0 gives the generic kernel."""
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
A = torch.empty((6 * 36, B), dtype=torch.float32, device="cuda")
pth = rbd.path(m, m.findbody("r_foot"), m.findbody("l_hand"))
small = {k: torch.empty((r, B), dtype=torch.float32, device="cuda") for k, r in
         (("center_of_mass", 3), ("kinetic_energy", 1), ("gravitational_potential_energy", 1), ("momentum", 6), ("momentum_rate_bias", 6))}
tr = torch.empty((12 * len(m.joints), B), dtype=torch.float32, device="cuda")
for name, fn, byts in (("geometric_jacobian", lambda: rbd.geometric_jacobian_(A, st, pth), (37 + 216) * 4),
                       ("com+energies+momenta", lambda: rbd.kinematics_(st, None, **small), (37 + 36 + 17) * 4),
                       ("transforms_to_root", lambda: rbd.transforms_to_root_(tr, st), (37 + 12 * 36) * 4),
                       ("momentum_matrix", lambda: rbd.momentum_matrix_(A, st), (37 + 216) * 4)):
    for _ in range(3):
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
    print(f"atlas {name}: {ms:.3f} ms  {B / ms / 1e3:.1f} M evals/s  {B * byts / ms / 1e6:.0f} GB/s  specialised={info.specialised} "
          f"launches={info.kernels_launched}", flush=True)
