"""Generates tests/golden/oracle_*.npz: small input/output vectors of the CPU oracle (oracle/sll_oracle.c), so the
GPU parity tests also run against COMMITTED numbers and not only against an oracle rebuilt on the test box.
The reference itself cannot be run here (Fortran, no compiler in the image); the oracle is pinned to the
reference by the golden file reffile_bsl_vp_3d3v_cart_dd.dat and the reference's analytic known-answer tests
(tests/test_oracle.py).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261017)

# 1. batched advection on a 4D field, every axis, every method
shape = (16, 8, 12, 8)
f0 = np.asfortranarray(rng.standard_normal(shape))
out = {"f0": f0}
for axis in range(4):
    v_axis = (axis + 2) % 4
    disp = rng.uniform(-0.9, 0.9, shape[v_axis])
    if v_axis > axis:
        dsel = (int(np.prod(shape[axis + 1:v_axis], dtype=np.int64)), shape[v_axis], 1, 1, 1, 0)
    else:
        dsel = (1, 1, 0, int(np.prod(shape[:v_axis], dtype=np.int64)), shape[v_axis], 1)
    out[f"disp{axis}"] = disp
    out[f"dsel{axis}"] = np.array(dsel)
    for method, order in (("spline", 4), ("lagrange_fixed", 7), ("lagrange_centered", 6)):
        if method != "spline" and axis in (1, 3):
            continue
        out[f"{method}{order}_axis{axis}"] = orc.advect_axis(f0.copy(order="F"), axis, method, order, disp, dsel)
np.savez_compressed(os.path.join(HERE, "oracle_advect4d.npz"), **out)

# 2. line-granular objects
n = 64
x = np.arange(n + 1) * (2 * np.pi / n)
line = 2.0 * (np.sin(x) + 2.5 + np.cos(x))
lines = {"line": line,
         "adv_spline": orc.advect_1d_periodic_constant("spline", n, 0.0, 2 * np.pi, 4, 1.3, 0.1, line),
         "adv_lagrange6": orc.advect_1d_periodic_constant("lagrange", n, 0.0, 2 * np.pi, 6, 1.3, 0.1, line),
         "spline_disp": orc.spline_interpolate_array_disp(line, 0.0, 2 * np.pi, -1.2 * (2 * np.pi / n))}
np.savez_compressed(os.path.join(HERE, "oracle_lines.npz"), **lines)

# 3. Poisson 2D on random data (exercises the non-Hermitian Nyquist handling) and 1D
rho2 = np.asfortranarray(rng.standard_normal((16, 12)))
e1, e2 = orc.poisson_2d(rho2.copy(order="F"), 16, 12, 0.0, 4 * np.pi, 0.0, 2 * np.pi)
rho1 = rng.standard_normal(33)
rho1[-1] = rho1[0]
np.savez_compressed(os.path.join(HERE, "oracle_poisson.npz"), rho2=rho2, e1=e1, e2=e2, rho1=rho1,
                    e1d=orc.poisson_1d(rho1, 0.0, 4 * np.pi))

# 4. simulation traces
rows4, f4 = orc.sim4d([16, 16, 32, 32], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1, 5, want_f=True)
rows2 = orc.sim2d(64, 64, 0.0, 4 * np.pi, -6.0, 6.0, 0, 0.5, 1e-3, 0.1, 20)
np.savez_compressed(os.path.join(HERE, "oracle_traces.npz"), rows4=rows4, rows2=rows2)
print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])
