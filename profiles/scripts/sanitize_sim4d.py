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
sb.synchronize()
print("done, launches", sb.launch_count())
