"""A/B timing of the T-stage plane kernel variants on the 128^4 field (CUDA events, 20 launches each):
register-resident (ept 0) vs in-place (ept 16 / 32), with and without the fused charge density, against the two
separate passes + reduction they replace."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import selalib_b200 as sb  # noqa: E402
from selalib_b200.capi import DispT, dp, vp, lib  # noqa: E402

n = int(os.environ.get("SLLB_BENCH_N", "128"))
sb.init(0)
F = sb.Field([n] * 4)
F.upload(np.asfortranarray(np.random.default_rng(0).random((n,) * 4)))
v = torch.linspace(-2.3, 2.3, n, dtype=torch.float64, device="cuda")
rho = torch.empty(n * n, dtype=torch.float64, device="cuda")
pts = float(n) ** 4


def disp(dsel):
    d = DispT()
    d.values = C.cast(vp(v.data_ptr()), dp); d.nvalues = 0; d.values_on_device = 1; d.scale = 1.0
    d.odiv, d.omod, d.ostr, d.idiv, d.imod, d.istr = dsel
    return d


d0, d1 = disp((n, n, 1, 1, 1, 0)), disp((n, n, 1, 1, 1, 0))


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


def plane(with_rho):
    rc = lib().sllb_advect_plane(F.h, 0, 4, C.byref(d0), C.byref(d1), C.c_double(1.0),
                                 C.cast(vp(rho.data_ptr()), dp) if with_rho else None)
    assert rc == 0, sb.last_error()


for ept in ((0,) if os.environ.get("SLLB_AB_EPT0_ONLY") else (0, 16, 32)):
    sb.set_plane_kernel(True, ept)
    for with_rho in (False, True):
        ms = timeit(lambda: plane(with_rho))
        print(f"plane kernel ept={ept:2d} rho={int(with_rho)}   {ms:8.4f} ms   {2 * 16 * pts / ms / 1e6:8.1f} GB/s-equivalent (2 passes)")
sb.set_plane_kernel(True, 0)
if os.environ.get("SLLB_AB_EPT0_ONLY"):
    sys.exit(0)
ms1 = timeit(lambda: F.advect_axis(0, sb.METHOD_SPLINE, 4, v.data_ptr(), 1.0, (n, n, 1, 1, 1, 0), on_device=True))
ms2 = timeit(lambda: F.advect_axis(1, sb.METHOD_SPLINE, 4, v.data_ptr(), 1.0, (1, n, 1, 1, 1, 0), on_device=True))
rho2 = torch.empty(n * n, dtype=torch.float64, device="cuda")
ms3 = timeit(lambda: lib().sllb_reduce_velocity(F.h, 2, C.c_double(1.0), C.cast(vp(rho2.data_ptr()), dp)))
print(f"separate: x1 pass {ms1:.4f} + x2 pass {ms2:.4f} + reduction {ms3:.4f} = {ms1 + ms2 + ms3:.4f} ms")
