"""A/B timing of the advection kernels' tuning knobs on the 128^4 field (CUDA events, 20 launches each)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selalib_b200 as sb  # noqa: E402

n = int(os.environ.get("SLLB_BENCH_N", "128"))
sb.init(0)
F = sb.Field([n] * 4)
F.upload(np.asfortranarray(np.random.default_rng(0).random((n,) * 4)))
v = torch.linspace(-2.3, 2.3, n, dtype=torch.float64, device="cuda")
E = (1e-2 * torch.sin(torch.arange(n * n, dtype=torch.float64, device="cuda"))).contiguous()
pts = float(n) ** 4


def timeit(call, reps=20):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        call()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def axis_call(axis, method=sb.METHOD_SPLINE, order=4):
    if axis < 2:
        dsel = (n if axis == 0 else 1, n, 1, 1, 1, 0)
        return lambda: F.advect_axis(axis, method, order, v.data_ptr(), 0.3 if method else 1.0, dsel, on_device=True)
    return lambda: F.advect_axis(axis, method, order, E.data_ptr(), 1.0, (1, 1, 0, 1, n * n, 1), on_device=True)


out = {}
for split in (1, 2, 4, 8):
    for staging in (1, 2):
        sb.set_spline_split(split); sb.set_staging(staging)
        for axis in (1, 2, 3):
            ms = timeit(axis_call(axis))
            out[f"spline split={split} staging={'tma' if staging == 1 else 'cpasync'} x{axis + 1}"] = [ms, 16 * pts / ms / 1e6]
sb.set_spline_split(-1); sb.set_staging(0)
out["spline x1 (contig)"] = [timeit(axis_call(0))] * 1
out["spline x1 (contig)"].append(16 * pts / out["spline x1 (contig)"][0] / 1e6)
for order in (3, 7, 11):
    for axis in (0, 1, 3):
        ms = timeit(axis_call(axis, sb.METHOD_LAGRANGE_FIXED, order))
        out[f"lagrange{order} x{axis + 1}"] = [ms, 16 * pts / ms / 1e6]
ms = timeit(lambda: F.reduce_velocity(2, 1.0))
out["reduce_velocity (incl. D2H of rho)"] = [ms, 8 * pts / ms / 1e6]
for k, (ms, gbs) in out.items():
    print(f"{k:50s} {ms:8.4f} ms  {gbs:8.1f} GB/s")
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "ab_kernels.json"), "w"), indent=1)
