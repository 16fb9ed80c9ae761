"""C2: 1D1V two-stream instability 1024 x 1024 fp64, fixed 7-point Lagrange in both directions, Strang VTV, single B200
(sim_bsl_vp_1d1v_cart semantics).  8 MB of f: the problem lives in L2, the step is bound by kernel launches, not HBM.
One JSON line: point-updates/s per advection pass over whole steps, and microseconds per step."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selalib_b200 as sb  # noqa: E402

n = int(os.environ.get("SLLB_C2_N", "1024"))
steps = int(os.environ.get("SLLB_C2_STEPS", "2000"))
spline = os.environ.get("SLLB_C2_METHOD", "lagrange7") == "spline"   # C1: 128 x 128 Landau damping with cubic splines
sb.init(0)
res = {"workload": (f"1D1V Landau damping {n}x{n} fp64, periodic cubic splines, Strang VTV, dt=0.1, k=0.5, eps=1e-3, v in [-6,6]" if spline else
                    f"1D1V two-stream {n}x{n} fp64, Lagrange fixed 7-point, Strang VTV, dt=0.002, k=0.5, eps=0.01, v in [-6,6]"),
       "passes_per_step": 3, "points": n * n}
for graphs in ([0, 1] if hasattr(sb, "set_cuda_graphs") else [None]):
    if graphs is not None:
        sb.set_cuda_graphs(graphs)
    if spline:
        S = sb.Sim2d(n, n, 0.0, 4 * np.pi, -6.0, 6.0, 0, 0.5, 1e-3, 0.1, method=sb.METHOD_SPLINE, order=4)
    else:
        S = sb.Sim2d(n, n, 0.0, 4 * np.pi, -6.0, 6.0, 1, 0.5, 0.01, 0.002, method=sb.METHOD_LAGRANGE_FIXED, order=7)
    S.run(200, diagnostics=False) if "diagnostics" in S.run.__code__.co_varnames else S.run(200)
    torch.cuda.synchronize()
    sb.launch_count_reset()
    t0 = time.perf_counter()
    S.run(steps, diagnostics=False) if "diagnostics" in S.run.__code__.co_varnames else S.run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    key = "default" if graphs is None else ("cuda_graph" if graphs else "stream_launches")
    res[key] = {"us_per_step": 1e6 * dt / steps, "point_updates_per_s": 3.0 * n * n * steps / dt,
                "kernel_launches_per_step": sb.launch_count() / steps}
    rows = S.run(1)
    res[key]["mass"] = float(rows[0, 1])
    S.destroy()
print(json.dumps(res))
