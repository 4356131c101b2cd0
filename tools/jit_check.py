#!/usr/bin/env python
"""Times one entry point on Atlas / iiwa14 with the model-specialised (NVRTC) kernels or the generic ones (RBD_JIT=0) and
checks a strided sub-sample against the fp64 oracle.  usage: jit_check.py [model] [algo] [dtype] [log2 batch]"""
import sys, os, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rigidbodydynamics.jl_b200 as rbd
from oracle import Oracle

name = sys.argv[1] if len(sys.argv) > 1 else "atlas"
algo = sys.argv[2] if len(sys.argv) > 2 else "dynamics"
dtype = {"f32": torch.float32, "f64": torch.float64}[sys.argv[3] if len(sys.argv) > 3 else "f32"]
B = 1 << int(sys.argv[4] if len(sys.argv) > 4 else 20)
mech = rbd.load_model(name, floating=(name != "iiwa14"))
rng = np.random.default_rng(1)
st = rbd.MechanismState(mech, B, dtype)
rbd.rand_(st, rng)
nv = st.nv
x = torch.rand((nv, B), dtype=dtype, device="cuda")
res = rbd.DynamicsResult(mech, B, dtype)
out = torch.empty((nv, B), dtype=dtype, device="cuda")
if algo == "dynamics":
    fn = lambda: rbd.dynamics_(res, st, x, want_qd=False)
elif algo == "id":
    fn = lambda: rbd.inverse_dynamics_(out, st, x)
t0 = time.time(); fn(); torch.cuda.synchronize(); first = time.time() - t0
info = rbd.launch_info()
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 20
e0.record()
for _ in range(steps): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
idx = torch.arange(0, B, max(1, B // 512), device="cuda")
o = Oracle(mech.flatten())
qn, vn, xn = (t[:, idx].double().cpu().numpy() for t in (st.q, st.v, x))
if algo == "dynamics":
    ref = o.dynamics(qn, vn, xn); got = res.vd[:, idx].double().cpu().numpy()
else:
    ref = o.inverse_dynamics(qn, vn, xn); got = out[:, idx].double().cpu().numpy()
err = float((np.abs(got - ref).max(0) / np.maximum(1.0, np.abs(ref).max(0))).max())
print(json.dumps({"model": name, "algo": algo, "dtype": str(dtype), "B": B, "jit_env": os.environ.get("RBD_JIT", "1"),
                  "specialised": info.specialised, "kernels": info.kernels_launched, "grid": info.grid, "bps": info.blocks_per_sm,
                  "first_call_s": round(first, 3), "ms": round(ms, 4), "Mevals_s": round(B / ms / 1e3, 1), "rel_err": err}))
