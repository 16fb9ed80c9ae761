"""Launches the fused Lagrange plane kernel (K2d) a couple of times on a 6D block (for ncu; never a bench number)."""
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
v0 = -0.13 * (-6.0 + 12.0 / shape[3] * np.arange(shape[3]))
v1 = -0.13 * (-6.0 + 12.0 / shape[4] * np.arange(shape[4]))
ds0 = (shape[1] * shape[2], shape[3], 1, 1, 1, 0)
ds1 = (shape[2] * shape[3], shape[4], 1, 1, 1, 0)
for _ in range(2):
    F.advect_plane(v0, ds0, 1.0, v1, ds1, 1.0, method=sb.METHOD_LAGRANGE_FIXED, order=7)
sb.synchronize()
print("done")
