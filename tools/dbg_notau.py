import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rigidbodydynamics.jl_b200 as rbd
from oracle import Oracle
from tests.util import rand_inputs, rel_err
mech = rbd.load_model("atlas", floating=True)
B = 1 << 16
base = rand_inputs(mech, 512, 11)
q, v, tau, vd = (np.tile(a, (1, B // 512)) for a in base[:4])
o = Oracle(mech.flatten())
st = rbd.MechanismState(mech, B, torch.float32)
st.q.copy_(torch.from_numpy(q).float()); st.v.copy_(torch.from_numpy(v).float())
res = rbd.DynamicsResult(mech, B, torch.float32)
rbd.dynamics_(res, st, None, want_qd=False)
torch.cuda.synchronize()
info = rbd.launch_info()
got = res.vd.double().cpu().numpy()
ref = o.dynamics(q[:, :512], v[:, :512], None)
err = np.abs(got[:, :512] - ref) / np.maximum(1.0, np.abs(ref).max(0))
print("specialised", info.specialised, "kernels", info.kernels_launched, "rel_err", err.max(), "worst row", err.max(1).argmax(), "nan", np.isnan(got).sum())
print("per-row max err", np.round(err.max(1), 7)[:12])
blocks = got.reshape(36, -1, 512)
print("blocks equal to first:", np.array_equal(blocks, np.broadcast_to(blocks[:, :1], blocks.shape)), "max diff across blocks", np.abs(blocks - blocks[:, :1]).max())
