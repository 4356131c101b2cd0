#!/usr/bin/env python
"""Runs rbd_dynamics_derivatives twice on Atlas (first call warms up); for ncu launch lists / captures.
usage: prof_deriv.py [f64|f32] [log2 batch] [model]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rigidbodydynamics.jl_b200 as rbd

dt = torch.float32 if (len(sys.argv) > 1 and sys.argv[1] == "f32") else torch.float64
B = 1 << int(sys.argv[2] if len(sys.argv) > 2 else 15)
name = sys.argv[3] if len(sys.argv) > 3 else "atlas"
mech = rbd.load_model(name, floating=(name != "iiwa14"))
st = rbd.MechanismState(mech, B, dt)
rbd.rand_(st, np.random.default_rng(1))
nv = st.nv
tau = torch.rand((nv, B), dtype=dt, device="cuda")
res = rbd.DynamicsResult(mech, B, dt)
dq = torch.empty((nv * nv, B), dtype=dt, device="cuda")
dv = torch.empty_like(dq)
for _ in range(2):
    rbd.dynamics_derivatives_(dq, dv, res, st, tau)
    torch.cuda.synchronize()
