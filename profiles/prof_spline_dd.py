"""CUDA-event timing of the K9 kernels (local cubic spline with halo cells) per axis of a 6D block, next to the fixed
7-point Lagrange pass on the same block: GB/s at 16 B/point/pass and the fraction of the measured HBM peak.
One JSON line.  With SLLB_PROF_ONLY=1 it only launches every kernel twice (for ncu)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selalib_b200 as sb  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shape = [int(v) for v in os.environ.get("SLLB_DD_SHAPE", "32,32,32,20,20,20").split(",")]
only = os.environ.get("SLLB_PROF_ONLY", "0") == "1"
sb.init(0)
D = sb.Dd6d(None, shape)
F = D.field()
npts = int(np.prod(shape, dtype=np.int64))
rng = np.random.default_rng(20261017)
F.upload(np.asfortranarray(rng.standard_normal(npts).reshape(shape, order="F")))
nx3 = shape[0] * shape[1] * shape[2]
E = rng.uniform(-0.9, 0.9, nx3)
peak = None
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def dsel_x(axis):
    stride = int(np.prod(shape[axis + 1:axis + 3], dtype=np.int64))
    return (stride, shape[axis + 3], 1, 1, 1, 0)


def timed(fn, reps):
    fn(); fn(); sb.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(torch.cuda.default_stream())
    for _ in range(reps):
        fn()
    e1.record(torch.cuda.default_stream())
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {"shape": shape, "points": npts, "peak_gbs": peak, "kernels": {}}
reps = 2 if only else 10
for axis in range(6):
    if axis < 3:
        v = -6.0 + 12.0 / shape[axis + 3] * np.arange(shape[axis + 3])
        disp = -v * 0.13
        shift, _, _ = sb.spline_dd_blocks(disp)
        cases = {"spline_dd_wrap": lambda a=axis, d=disp, s=shift: F.advect_axis_spline_dd(a, d, dsel_x(a), shift=s),
                 "lagrange7": lambda a=axis, d=disp: F.advect_axis(a, sb.METHOD_LAGRANGE_FIXED, 7, d, 1.0, dsel_x(a)),
                 "spline27": lambda a=axis, d=disp: F.advect_axis(a, sb.METHOD_SPLINE, 4, d, 1.0, dsel_x(a))}
    else:
        dsel = (1, 1, 0, 1, nx3, 1)
        cases = {"spline_dd_wrap": lambda a=axis: F.advect_axis_spline_dd(a, E, dsel),
                 "lagrange7": lambda a=axis: F.advect_axis(a, sb.METHOD_LAGRANGE_FIXED, 7, E, 1.0, dsel),
                 "spline27": lambda a=axis: F.advect_axis(a, sb.METHOD_SPLINE, 4, E, 1.0, dsel)}
        if shape[axis] >= 18:
            def halo(a=axis):
                sb.dd6d_set_force_halo(True)
                D.advect_axis_spline(a, E, dsel=dsel, hw=(1, 1))
                sb.dd6d_set_force_halo(False)
            cases["spline_dd_halo_path(prepare+pack+kernel)"] = halo
    for name, fn in cases.items():
        try:
            ms = timed(fn, reps)
        except sb.SllbError as e:
            res["kernels"][f"axis{axis}:{name}"] = {"error": str(e)}
            continue
        gbs = 16.0 * npts / (ms * 1e-3) / 1e9
        res["kernels"][f"axis{axis}:{name}"] = {"ms": ms, "gbs_at_16B_per_point": gbs, "frac_of_measured_hbm": gbs / peak if peak else None}
print(json.dumps(res))
