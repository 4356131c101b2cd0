"""Writes the committed fixtures under tests/golden/ (run from the repo root: `python tools/make_golden.py`).

1. double_pendulum_closed_form.json -- REFERENCE-DERIVED known answers: the closed-form mass matrix, Coriolis matrix, gravity
   vector and link kinetic energies that the reference's own test asserts to 1e-12 (test/test_double_pendulum.jl:2-11, 40-75),
   evaluated with numpy for its parameters at seeded states, plus the quick-start state of examples/1.  Nothing
   here comes from this repository's kernels or oracle.
2. <model>_seed<k>.npz -- REGRESSION vectors: seeded inputs (numpy PCG64, the distributions of SURVEY 8(d)) and the fp64 outputs of
   oracle/ (the CPU restatement of the reference's algorithm) for Atlas / Valkyrie / the 7-DoF arm.  The reference itself cannot
   be executed in this environment (no Julia), so these are NOT reference-produced; they freeze the oracle that the reference's
   identities pin (tests/test_oracle.py), and let the GPU tier check the CUDA path without running the oracle.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def closed_form(q, v, vd, p):
    lc1, l1, m1, I1, lc2, m2, I2, g = (p[k] for k in ("lc1", "l1", "m1", "I1", "lc2", "m2", "I2", "g"))
    q1, q2 = q
    v1, v2 = v
    c2, s1, s2, s12 = np.cos(q2), np.sin(q1), np.sin(q2), np.sin(q1 + q2)
    M = np.array([[I1 + I2 + m2 * l1 ** 2 + 2 * m2 * l1 * lc2 * c2, I2 + m2 * l1 * lc2 * c2], [I2 + m2 * l1 * lc2 * c2, I2]])
    C = np.array([[-2 * m2 * l1 * lc2 * s2 * v2, -m2 * l1 * lc2 * s2 * v2], [m2 * l1 * lc2 * s2 * v1, 0.0]])
    G = np.array([m1 * g * lc1 * s1 + m2 * g * (l1 * s1 + lc2 * s12), m2 * g * lc2 * s12])
    T1 = 0.5 * I1 * v1 ** 2
    T2 = 0.5 * (m2 * l1 ** 2 + I2 + 2 * m2 * l1 * lc2 * c2) * v1 ** 2 + 0.5 * I2 * v2 ** 2 + (I2 + m2 * l1 * lc2 * c2) * v1 * v2
    tau = M @ np.asarray(vd) + C @ np.asarray(v) + G
    bias = C @ np.asarray(v) + G
    vd_passive = np.linalg.solve(M, -bias)
    return {"q": list(q), "v": list(v), "vd": list(vd), "M": M.tolist(), "bias": bias.tolist(), "tau": tau.tolist(),
            "kinetic_energy": float(T1 + T2), "vd_passive": vd_passive.tolist()}


def main():
    os.makedirs(OUT, exist_ok=True)
    test_params = dict(lc1=-0.5, l1=-1.0, m1=1.0, I1=0.333, lc2=-1.0, m2=1.0, I2=1.33, g=-9.81)        # test_double_pendulum.jl:2-11
    quick_params = dict(lc1=-0.5, l1=-1.0, m1=1.0, I1=0.333, lc2=-0.5, m2=1.0, I2=0.333, g=-9.81)      # examples/1
    rng = np.random.default_rng(6)
    cases = [closed_form(rng.standard_normal(2), rng.random(2), rng.random(2), test_params) for _ in range(8)]
    doc = {"source": "closed form of test/test_double_pendulum.jl:40-75 evaluated by tools/make_golden.py (numpy fp64)",
           "test_parameters": test_params, "test_cases": cases,
           "quickstart_parameters": quick_params,
           "quickstart": closed_form([0.3, 0.4], [1.0, 2.0], [1.0, 2.0], quick_params)}
    with open(os.path.join(OUT, "double_pendulum_closed_form.json"), "w") as f:
        json.dump(doc, f, indent=1)

    import rigidbodydynamics.jl_b200 as rbd
    from oracle import Oracle
    from tests.util import rand_inputs
    for name, floating, seed in (("atlas", True, 17), ("valkyrie", True, 18), ("iiwa14", False, 19)):
        mech = rbd.load_model(name, floating=floating)
        desc = mech.flatten()
        o = Oracle(desc)
        B = 16
        q, v, tau, vd, w = rand_inputs(mech, B, seed, wext=True)
        sign = rbd.path(mech, mech.joints[-1].successor, mech.joints[min(2, desc.nb - 1)].successor).sign
        kin = o.kinematics(q, v, sign)
        np.savez_compressed(
            os.path.join(OUT, f"{name}_seed{seed}.npz"), q=q, v=v, tau=tau, vd_in=vd, wext=w, path_sign=sign,
            dynamics=o.dynamics(q, v, tau), dynamics_wext=o.dynamics(q, v, tau, w), qdot=o.dynamics(q, v, tau, want_qd=True)[1],
            inverse_dynamics=o.inverse_dynamics(q, v, vd), inverse_dynamics_wext=o.inverse_dynamics(q, v, vd, w),
            dynamics_bias=o.dynamics_bias(q, v), mass_matrix=o.mass_matrix(q),
            **{"kin_" + k: a for k, a in kin.items()})
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
