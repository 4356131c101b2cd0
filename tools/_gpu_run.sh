set -x
python -m pytest tests -m gpu -x -q -k "simulate or rk4 or schedule" 2>&1 | grep -v "^E    " | tail -5
python - <<'PY'
import numpy as np, torch, time, sys
sys.path.insert(0, ".")
import rigidbodydynamics.jl_b200 as rbd
from oracle import Oracle
mech = rbd.load_model("atlas", floating=True)
B = 1 << 20
st = rbd.MechanismState(mech, B, torch.float32); rbd.rand_(st, np.random.default_rng(1))
tau = torch.rand((36, B), dtype=torch.float32, device="cuda")
# parity of the vectorised path at a large batch: first 64 samples vs oracle after 3 steps
q0 = st.q[:, :64].double().cpu().numpy(); v0 = st.v[:, :64].double().cpu().numpy(); t0 = tau[:, :64].double().cpu().numpy()
s2 = rbd.MechanismState(mech, B, torch.float32); s2.q.copy_(st.q); s2.v.copy_(st.v)
rbd.simulate_(s2, 3e-3 - 1e-9, tau, dt=1e-3)
qr, vr = Oracle(mech.flatten()).integrate(q0, v0, t0, dt=1e-3, nsteps=3)
print("max |dq|", np.abs(s2.q[:, :64].double().cpu().numpy() - qr).max(), "max |dv|", np.abs(s2.v[:, :64].double().cpu().numpy() - vr).max())
def run():
    rbd.simulate_(st, 1e-4, tau, dt=1e-4)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print("RK4 step ms", e0.elapsed_time(e1) / 10, "launches", rbd.launch_info().kernels_launched)
PY
