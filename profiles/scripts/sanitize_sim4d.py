"""Small 2D2V (16^2 x 32^2) and 1D1V runs + the plane kernel at 32^4, for compute-sanitizer (never a bench number)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import selalib_b200 as sb  # noqa: E402

sb.init(0)
S = sb.Sim4d([16, 16, 32, 32], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1)
rows = S.run(2)
print("sim4d 16^2x32^2 rows", rows[-1])
S.destroy()
S = sb.Sim4d([32, 32, 32, 32], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1)
rows = S.run(2)
print("sim4d 32^4 rows (plane kernel)", rows[-1])
S.destroy()
# the round-2 kernels: 64-point lines (chunked strided kernel with the fused diagnostics, recorded step), order 6 splines,
# the dup_velocity_planes mode, the 1D1V loop with its mode columns, the parallel-variant Poisson solve
S = sb.Sim4d([16, 16, 64, 64], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1)
rows = S.run(4)
print("sim4d 16^2x64^2 rows (fused diagnostics, recorded step)", rows[-1])
S.destroy()
S = sb.Sim4d([16, 16, 32, 32], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1, dup_velocity_planes=True,
             method=[sb.METHOD_SPLINE] * 4, order=[6, 4, 4, 8])
rows = S.run(2)
print("sim4d dup planes + order 6/8 splines", rows[-1])
S.destroy()
S2 = sb.Sim2d(64, 64, 0.0, 4 * np.pi, -6.0, 6.0, 0, 0.5, 1e-3, 0.1)
S2.run(3)
print("sim2d thdiag", S2.thdiag(3)[:8])
S2.destroy()
P = sb.Poisson((64, 48), (0.0, 0.0), (1.0, 2.0), par=True)
print("poisson par", float(np.abs(P.solve(np.random.default_rng(0).standard_normal((65, 49)))).max()))
P.destroy()
sb.synchronize()
print("done, launches", sb.launch_count())
