"""Builds and loads tests/host/spline15_host.cpp: the per-line device functions of K9 compiled for the host
(test infrastructure; lets the CPU suite check the CUDA path's arithmetic against the oracle)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libspline15_host.so")
        deps = [os.path.join(_HERE, "spline15_host.cpp"),
                os.path.join(_HERE, "..", "..", "selalib_b200", "csrc", "sllb_spline15.cuh"),
                os.path.join(_HERE, "..", "..", "selalib_b200", "csrc", "sllb_hermite.cuh"),
                os.path.join(_HERE, "..", "..", "selalib_b200", "csrc", "sllb_lagrange.cuh")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, deps[0]])
        _LIB = C.CDLL(so)
        _LIB.emu_spline_dd_line.restype = C.c_int
    return _LIB


def spline_dd_line(line, nblk, si, alpha, hwl=1, hwr=1, mode=0):
    line = np.ascontiguousarray(line, dtype=np.float64)
    out = np.empty_like(line)
    dp = C.POINTER(C.c_double)
    rc = lib().emu_spline_dd_line(line.ctypes.data_as(dp), out.ctypes.data_as(dp), C.c_int(line.size), C.c_int(nblk),
                                  C.c_int(si), C.c_double(alpha), C.c_int(hwl), C.c_int(hwr), C.c_int(mode))
    assert rc == 0, rc
    return out


def hermite_line(line, delta, alpha, inplace=True, slopes=None):
    line = np.ascontiguousarray(line, dtype=np.float64)
    out = np.empty_like(line)
    dp = C.POINTER(C.c_double)
    hs = slopes is not None
    lib().emu_hermite_line.restype = C.c_int
    rc = lib().emu_hermite_line(line.ctypes.data_as(dp), out.ctypes.data_as(dp), C.c_int(line.size), C.c_double(delta),
                                C.c_double(alpha), C.c_int(1 if inplace else 0), C.c_int(hs),
                                C.c_double(slopes[0] if hs else 0.0), C.c_double(slopes[1] if hs else 0.0))
    assert rc == 0, rc
    return out


def lagrange_line(line, disp, stencil):
    line = np.ascontiguousarray(line, dtype=np.float64)
    out = np.empty_like(line)
    dp = C.POINTER(C.c_double)
    lib().emu_lagrange_line.restype = C.c_int
    rc = lib().emu_lagrange_line(line.ctypes.data_as(dp), out.ctypes.data_as(dp), C.c_int(line.size), C.c_double(disp), C.c_int(stencil))
    assert rc == 0, rc
    return out


def spline_dd_prepare(local, si):
    local = np.ascontiguousarray(local, dtype=np.float64)
    a, b = C.c_double(), C.c_double()
    lib().emu_spline_dd_prepare(local.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(local.size), C.c_int(si), C.byref(a), C.byref(b))
    return a.value, b.value


def spline_dd_piece(local, halo_l, halo_r, si, alpha, rem_d, rem_c):
    dp = C.POINTER(C.c_double)
    local = np.ascontiguousarray(local, dtype=np.float64)
    hl = np.ascontiguousarray(halo_l, dtype=np.float64); hr = np.ascontiguousarray(halo_r, dtype=np.float64)
    out = np.empty_like(local)
    lib().emu_spline_dd_piece.restype = C.c_int
    rc = lib().emu_spline_dd_piece(local.ctypes.data_as(dp), hl.ctypes.data_as(dp), C.c_int(hl.size), hr.ctypes.data_as(dp),
                                   C.c_int(hr.size), C.c_int(local.size), C.c_int(si), C.c_double(alpha), C.c_double(rem_d),
                                   C.c_double(rem_c), out.ctypes.data_as(dp))
    assert rc == 0, rc
    return out
