"""Launches each hot kernel a few times on the 128^4 field (for ncu captures; never a bench number)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selalib_b200 as sb  # noqa: E402

n = int(os.environ.get("SLLB_BENCH_N", "128"))
sb.init(0)
S = sb.Sim4d([n] * 4, [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1)
F = S.field()
v = np.linspace(-2.3, 2.3, n)
E = 1e-2 * np.sin(np.arange(n * n))
reps = int(os.environ.get("SLLB_PROF_REPS", "2"))
for _ in range(reps):
    F.advect_axis(0, sb.METHOD_SPLINE, 4, v, 1.0, (n, n, 1, 1, 1, 0))
    F.advect_axis(1, sb.METHOD_SPLINE, 4, v, 1.0, (1, n, 1, 1, 1, 0))
    F.advect_axis(2, sb.METHOD_SPLINE, 4, E, 1.0, (1, 1, 0, 1, n * n, 1))
    F.advect_axis(3, sb.METHOD_SPLINE, 4, E, 1.0, (1, 1, 0, 1, n * n, 1))
    F.reduce_velocity(2, 1.0)
    F.advect_plane(v, (n, n, 1, 1, 1, 0), 1.0, v, (n, n, 1, 1, 1, 0), 1.0, rho_scale=1.0)
    F.advect_axis(1, sb.METHOD_LAGRANGE_FIXED, 7, v, 0.3, (1, n, 1, 1, 1, 0))
    F.advect_axis(0, sb.METHOD_LAGRANGE_FIXED, 7, v, 0.3, (n, n, 1, 1, 1, 0))
sb.synchronize()
print("done")
