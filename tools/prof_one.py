"""Run ONE Atlas floating fp32 dynamics call (batch from argv, default 2^18) -- the target of the ncu captures in profiles/."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rigidbodydynamics.jl_b200 as rbd

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
mech = rbd.load_model("atlas", floating=True)
st = rbd.MechanismState(mech, B, torch.float32)
rbd.rand_(st, np.random.default_rng(1))
tau = torch.rand((st.nv, B), dtype=torch.float32, device="cuda")
res = rbd.DynamicsResult(mech, B, torch.float32)
for _ in range(2):
    rbd.dynamics_(res, st, tau, want_qd=False)
torch.cuda.synchronize()
