"""Launches the contiguous-axis (eta1) kernels and the K9/K10 kernels a few times on a 6D block (for ncu captures;
never a bench number)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selalib_b200 as sb  # noqa: E402

shape = [int(v) for v in os.environ.get("SLLB_DD_SHAPE", "32,32,32,20,20,20").split(",")]
sb.init(0)
F = sb.Field(shape)
rng = np.random.default_rng(20261017)
F.upload(np.asfortranarray(rng.standard_normal(int(np.prod(shape))).reshape(shape, order="F")))
v = -6.0 + 12.0 / shape[3] * np.arange(shape[3])
disp = -v * 0.13
shift, _, _ = sb.spline_dd_blocks(disp)
dsel0 = (shape[1] * shape[2], shape[3], 1, 1, 1, 0)
dsel1 = (shape[2], shape[4], 1, 1, 1, 0)
nx3 = shape[0] * shape[1] * shape[2]
E = rng.uniform(-0.9, 0.9, nx3)
for _ in range(int(os.environ.get("SLLB_PROF_REPS", "2"))):
    F.advect_axis(0, sb.METHOD_LAGRANGE_FIXED, 7, disp, 1.0, dsel0)
    F.advect_axis(0, sb.METHOD_SPLINE, 4, disp, 1.0, dsel0)
    F.advect_axis_spline_dd(0, disp, dsel0, shift=shift)
    F.advect_axis_spline_dd(1, disp, dsel1)
    F.advect_axis_spline_dd(3, E, (1, 1, 0, 1, nx3, 1))
sb.synchronize()
print("done")
