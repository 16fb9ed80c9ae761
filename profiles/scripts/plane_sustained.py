"""T-stage plane kernel (fused density) launched back to back for a few seconds: ms per launch of every block of 100
launches next to the SM clock and power draw sampled by nvidia-smi every 50 ms -- does the kernel slow down when the
power cap pulls the clock?"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import selalib_b200 as sb  # noqa: E402
from selalib_b200.capi import DispT, dp, vp, lib  # noqa: E402

n = 128
sb.init(0)
F = sb.Field([n] * 4)
F.upload(np.asfortranarray(np.random.default_rng(0).random((n,) * 4)))
v = torch.linspace(-2.3, 2.3, n, dtype=torch.float64, device="cuda")
rho = torch.empty(n * n, dtype=torch.float64, device="cuda")


def disp(dsel):
    d = DispT()
    d.values = C.cast(vp(v.data_ptr()), dp); d.nvalues = 0; d.values_on_device = 1; d.scale = 1.0
    d.odiv, d.omod, d.ostr, d.idiv, d.imod, d.istr = dsel
    return d


d0, d1 = disp((n, n, 1, 1, 1, 0)), disp((n, n, 1, 1, 1, 0))
which = sys.argv[1] if len(sys.argv) > 1 else "plane"


def call():
    if which == "plane":
        assert lib().sllb_advect_plane(F.h, 0, 4, C.byref(d0), C.byref(d1), C.c_double(1.0), C.cast(vp(rho.data_ptr()), dp)) == 0
    else:
        F.advect_axis(2, sb.METHOD_SPLINE, 4, v.data_ptr(), 1.0, (1, 1, 0, 1, 1, 0), on_device=True)


smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                       stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
for _ in range(3):
    call()
torch.cuda.synchronize()
blocks = 30
ev = [torch.cuda.Event(enable_timing=True) for _ in range(blocks + 1)]
ev[0].record()
for b in range(blocks):
    for _ in range(100):
        call()
    ev[b + 1].record()
torch.cuda.synchronize()
time.sleep(0.2)
smi.terminate()
out = smi.communicate()[0].strip().splitlines()
print(which, "ms per launch by block of 100:", " ".join(f"{ev[b].elapsed_time(ev[b + 1]) / 100:.3f}" for b in range(blocks)))
print("clock MHz / W every 50 ms:", " ".join(l.replace(", ", "/") for l in out))
