"""ctypes loader for the CPU oracle (oracle/sll_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by selalib_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, n) for n in ("sll_oracle.c", "sll_oracle_halo.c", "sll_oracle_split.c", "sll_oracle_hermite.c", "Makefile")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_sim6d_run.restype = C.c_int
        _LIB.orc_sim4d_run.restype = C.c_int
        _LIB.orc_sim2d_run.restype = C.c_int
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(dp)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP team size of the oracle's loops over lines (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def spline_interpolate_array_disp(data, xmin, xmax, alpha, fast=-1):
    data = _f(data); out = np.empty_like(data)
    lib().orc_cubic_spline_interpolate_array_disp(C.c_int(data.size), C.c_double(xmin), C.c_double(xmax),
                                                  C.c_int(fast), _p(data), C.c_double(alpha), _p(out))
    return out


def spline_interpolate_array_disp_inplace(data, xmin, xmax, alpha, fast=-1):
    data = _f(data).copy()
    lib().orc_cubic_spline_interpolate_array_disp_inplace(C.c_int(data.size), C.c_double(xmin), C.c_double(xmax),
                                                          C.c_int(fast), _p(data), C.c_double(alpha))
    return data


def spline_coeffs(data, fast=-1):
    data = _f(data); c = np.empty(data.size + 3)
    lib().orc_spline_compute_interpolant_periodic(_p(data), C.c_int(data.size), C.c_int(fast), _p(c))
    return c


def periodic_interp(u, alpha, kind="spline", order=4):
    u = _f(u); out = np.empty_like(u)
    fn = lib().orc_periodic_interp_spline if kind == "spline" else lib().orc_periodic_interp_lagrange
    fn(C.c_int(u.size), C.c_int(order), _p(u), C.c_double(alpha), _p(out))
    return out


def advect_1d_periodic_constant(kind, num_cells, xmin, xmax, order, A, dt, inp):
    inp = _f(inp); out = np.empty_like(inp)
    lib().orc_advect_1d_periodic_constant(C.c_int(0 if kind == "spline" else 1), C.c_int(num_cells),
                                          C.c_double(xmin), C.c_double(xmax), C.c_int(order), C.c_double(A),
                                          C.c_double(dt), _p(inp), _p(out), C.c_int(inp.size))
    return out


def advect_1d_bsl_constant(npts, eta_min, eta_max, A, dt, inp, fast=1):
    """sll_t_advector_1d_bsl%advect_1d_constant with explicit-Euler periodic characteristics + cubic-spline interpolator"""
    inp = _f(inp); out = np.empty_like(inp)
    assert inp.size == npts
    lib().orc_advect_1d_bsl_constant(C.c_int(npts), C.c_double(eta_min), C.c_double(eta_max), C.c_int(fast),
                                     C.c_double(A), C.c_double(dt), _p(inp), _p(out))
    return out


def lagr_coeff(s, p):
    pp = np.zeros(11)
    rc = lib().orc_lagr_coeff(C.c_int(s), C.c_double(p), _p(pp))
    assert rc == 0
    return pp[:s]


def lagrange(variant, fi, p, s):
    fi = _f(fi); fp = np.full_like(fi, np.nan)
    fn = getattr(lib(), "orc_lagrange_" + variant)
    fn.restype = C.c_int
    rc = fn(_p(fi), _p(fp), C.c_int(fi.size), C.c_double(p), C.c_int(s))
    if rc != 0:
        raise ValueError("Lagrange stencil not implemented")
    return fp


def lagrange_centered_barycentric(fi, xmin, xmax, d, periodic_last, alpha):
    fi = _f(fi); out = np.empty_like(fi)
    lib().orc_lagrange_centered_barycentric(_p(fi), _p(out), C.c_int(fi.size if periodic_last else fi.size + 1),
                                            C.c_double(xmin), C.c_double(xmax), C.c_int(d),
                                            C.c_int(periodic_last), C.c_double(alpha))
    return out


def poisson_1d(rhs, xmin, xmax):
    rhs = _f(rhs); field = np.empty_like(rhs)
    lib().orc_poisson_1d_periodic_solve(C.c_int(rhs.size - 1), C.c_double(xmin), C.c_double(xmax), _p(rhs), _p(field))
    return field


def poisson_2d(rho, nc_x, nc_y, x_min, x_max, y_min, y_max, want_phi=False):
    """rho: Fortran-ordered (ld1, ld2) array."""
    rho = np.asfortranarray(rho, dtype=np.float64)
    ld1, ld2 = rho.shape
    ex = np.zeros_like(rho, order="F"); ey = np.zeros_like(rho, order="F"); phi = np.zeros_like(rho, order="F")
    lib().orc_poisson_2d_periodic_solve_e(C.c_int(nc_x), C.c_int(nc_y), C.c_double(x_min), C.c_double(x_max),
                                          C.c_double(y_min), C.c_double(y_max), _p(rho), C.c_int(ld1), C.c_int(ld2),
                                          _p(ex), _p(ey), _p(phi) if want_phi else None)
    return (ex, ey, phi) if want_phi else (ex, ey)


def poisson_2d_par(rho, ncx, ncy, Lx, Ly):
    """sll_s_poisson_2d_periodic_par_solve (Delta phi = rho); rho: Fortran-ordered (ncx+1, ncy+1) array"""
    rho = np.asfortranarray(rho, dtype=np.float64)
    assert rho.shape == (ncx + 1, ncy + 1)
    phi = np.zeros_like(rho, order="F")
    lib().orc_poisson_2d_periodic_par_solve(C.c_int(ncx), C.c_int(ncy), C.c_double(Lx), C.c_double(Ly), _p(rho), _p(phi))
    return phi


def poisson_3d(rho, Lx, Ly, Lz):
    rho = np.asfortranarray(rho, dtype=np.float64)
    nx, ny, nz = rho.shape
    outs = [np.zeros_like(rho, order="F") for _ in range(4)]
    lib().orc_poisson_3d_periodic_solve(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_double(Lx), C.c_double(Ly),
                                        C.c_double(Lz), _p(rho), *[_p(o) for o in outs])
    return outs  # phi, ex, ey, ez


def reduction_34(f, delta3, delta4):
    f = np.asfortranarray(f, dtype=np.float64)
    n1, n2, n3, n4 = f.shape
    out = np.zeros((n1, n2), order="F")
    lib().orc_reduction_4d_to_2d_direction34(_p(f), C.c_int(n1), C.c_int(n2), C.c_int(n3), C.c_int(n4),
                                             C.c_double(delta3), C.c_double(delta4), _p(out))
    return out


def charge_density_6d(f, volume_v):
    f = np.asfortranarray(f, dtype=np.float64)
    n = (C.c_int * 6)(*f.shape)
    rho = np.zeros(f.shape[:3], order="F")
    lib().orc_charge_density_6d(_p(f), n, C.c_double(volume_v), _p(rho))
    return rho


METHODS = {"spline": 0, "fft_spline": 1, "fft_lagrange": 2, "lagrange_fixed": 3, "lagrange_centered": 4}


def advect_axis(f, axis, method, order, disp, dsel):
    """Advect the Fortran-ordered array `f` in place along `axis`.

    The displacement (in cells, out(i) = f(i + disp)) of line (o, in) is
    disp[((o // odiv) % omod) * ostr + ((in // idiv) % imod) * istr], dsel = those six ints.
    """
    assert f.flags.f_contiguous
    shape = f.shape
    inner = int(np.prod(shape[:axis], dtype=np.int64))
    outer = int(np.prod(shape[axis + 1:], dtype=np.int64))
    disp = _f(disp)
    L = C.c_long
    lib().orc_advect_axis(_p(f), L(outer), C.c_int(shape[axis]), L(inner), C.c_int(METHODS[method]), C.c_int(order),
                          _p(disp), *[L(int(v)) for v in dsel])
    return f


def advect_axis_sub(f, axis, method, order, disp, dsel, frac):
    """Same as advect_axis on the first 1/frac of the lines (bounded sample for CPU timing)."""
    assert f.flags.f_contiguous
    shape = f.shape
    inner = int(np.prod(shape[:axis], dtype=np.int64))
    outer = int(np.prod(shape[axis + 1:], dtype=np.int64))
    oc, ic = (max(1, outer // frac), inner) if outer > 1 else (outer, max(1, inner // frac))
    disp = _f(disp)
    L = C.c_long
    lib().orc_advect_axis_sub(_p(f), L(outer), C.c_int(shape[axis]), L(inner), L(oc), L(ic), C.c_int(METHODS[method]),
                              C.c_int(order), _p(disp), *[L(int(v)) for v in dsel])
    return oc * ic * shape[axis]


def sim6d(n, v_max, xmax, stencil_x, stencil_v, delta_t, nsteps, alpha, kx, vth=(1.0, 1.0, 1.0),
          time_in_phase=True, want_f=False, advector=0, vblk=(1, 1, 1)):
    """advector: 0 fixed, 1 centered, 2 spline; vblk: emulated ring ranks along eta4..6 (local splines depend on it)"""
    nn = (C.c_int * 6)(*n)
    rows = np.zeros((nsteps + 1, 14))
    f = np.zeros(tuple(n), order="F") if want_f else None
    lib().orc_sim6d_run_ex.restype = C.c_int
    rc = lib().orc_sim6d_run_ex(nn, C.c_double(v_max), (C.c_double * 3)(*xmax), C.c_int(stencil_x), C.c_int(stencil_v),
                                C.c_double(delta_t), C.c_int(nsteps), C.c_double(alpha), (C.c_double * 3)(*kx),
                                (C.c_double * 3)(*vth), C.c_int(1 if time_in_phase else 0), _p(rows),
                                _p(f) if want_f else None, C.c_int(advector), (C.c_int * 3)(*vblk))
    assert rc == 0, rc
    return (rows, f) if want_f else rows


SPLIT_CASES = ["SLL_STRANG_VTV", "SLL_STRANG_TVT", "SLL_LIE_TV", "SLL_LIE_VT", "SLL_TRIPLE_JUMP_TVT", "SLL_TRIPLE_JUMP_VTV",
               "SLL_ORDER6_VTV", "SLL_ORDER6_TVT", "SLL_ORDER6VP_TVT", "SLL_ORDER6VP_VTV", "SLL_ORDER6VPnew_TVT",
               "SLL_ORDER6VPnew1_VTV", "SLL_ORDER6VPnew2_VTV", "SLL_ORDER6VP2D_VTV", "SLL_ORDER6VPOT_VTV",
               "SLL_ORDER6VPOTnew1_VTV", "SLL_ORDER6VPOTnew2_VTV", "SLL_ORDER6VPOTnew3_VTV"]


def splitting_coeff(split, dt):
    """(steps, split_begin_T, dim_split_V) of sll_f_new_time_splitting_coeff for case number / name `split`"""
    if isinstance(split, str):
        split = SPLIT_CASES.index(split)
    s = np.zeros(32); nb = C.c_int(); bt = C.c_int(); dv = C.c_int()
    lib().orc_splitting_coeff.restype = C.c_int
    rc = lib().orc_splitting_coeff(C.c_int(split), C.c_double(dt), _p(s), C.byref(nb), C.byref(bt), C.byref(dv))
    if rc != 0:
        raise ValueError("split_case not defined")
    nT = (nb.value + (1 if bt.value else 0)) // 2
    nV = nb.value - nT
    return s[:nT + nV * dv.value].copy(), bool(bt.value), dv.value


def compute_w_hermite(r, s):
    w = np.zeros(s - r + 1)
    lib().orc_compute_w_hermite(C.c_int(r), C.c_int(s), _p(w))
    return w


def compute_jacobian(E1, E2, factor, r=-2, s=2):
    E1 = np.asfortranarray(E1, dtype=np.float64); E2 = np.asfortranarray(E2, dtype=np.float64)
    jac = np.zeros_like(E1, order="F")
    lib().orc_compute_jacobian(_p(E1), _p(E2), C.c_int(E1.shape[0] - 1), C.c_int(E1.shape[1] - 1), C.c_double(factor),
                               C.c_int(r), C.c_int(s), _p(jac))
    return jac


def sim4d(nc, xmin, xmax, kx1, kx2, eps, dt, nsteps, split=0, method=0, order=4, want_f=False, stencil=(-2, 2),
          want_thdiag=False, cells_only=False, want_fields=False):
    """split: case number (0 Strang VTV, 1 Strang TVT, 2 Lie TV, ... see SPLIT_CASES) or the namelist's name;
    cells_only: the v_max planes are reset to the v_min planes after every T stage (see orc_sim4d_run_ex2)"""
    if isinstance(split, str):
        split = SPLIT_CASES.index(split)
    rows = np.zeros((nsteps + 1, 6))
    thd = np.zeros((nsteps + 1, 13))
    f = np.zeros(tuple(c + 1 for c in nc), order="F") if want_f else None
    fields = np.zeros((nc[0] + 1, nc[1] + 1, 3), order="F") if want_fields else None
    lib().orc_sim4d_run_ex2.restype = C.c_int
    rc = lib().orc_sim4d_run_ex2((C.c_int * 4)(*nc), (C.c_double * 4)(*xmin), (C.c_double * 4)(*xmax), C.c_double(kx1),
                                 C.c_double(kx2), C.c_double(eps), C.c_double(dt), C.c_int(nsteps), C.c_int(split),
                                 C.c_int(method), C.c_int(order), _p(rows), _p(f) if want_f else None,
                                 C.c_int(stencil[0]), C.c_int(stencil[1]), _p(thd), C.c_int(1 if cells_only else 0),
                                 _p(fields) if want_fields else None)
    assert rc == 0, rc
    out = (rows,)
    if want_f:
        out += (f,)
    if want_thdiag:
        out += (thd,)
    if want_fields:   # rho, E1, E2 of the last field solve, duplicated end points included
        out += (fields,)
    return out if len(out) > 1 else rows


def sim2d(nc_x1, nc_x2, x1_min, x1_max, x2_min, x2_max, init, kmode, eps, dt, nsteps, method=0, order=4,
          want_f=False):
    rows = np.zeros((nsteps, 8))
    f = np.zeros((nc_x1 + 1, nc_x2 + 1), order="F") if want_f else None
    E = np.zeros(nc_x1 + 1)
    rc = lib().orc_sim2d_run(C.c_int(nc_x1), C.c_int(nc_x2), C.c_double(x1_min), C.c_double(x1_max),
                             C.c_double(x2_min), C.c_double(x2_max), C.c_int(init), C.c_double(kmode), C.c_double(eps),
                             C.c_double(dt), C.c_int(nsteps), C.c_int(method), C.c_int(order), _p(rows),
                             _p(f) if want_f else None, _p(E))
    assert rc == 0
    return (rows, f, E) if want_f else rows


# ---- local cubic spline with halo cells (sll_m_cubic_spline_halo_1d, sll_m_advection_6d_spline_dd_slim) ----
SKIP = -2 ** 31


def halo_prepare_exchange(fdata, si):
    fdata = _f(fdata); d0 = C.c_double(); c2 = C.c_double()
    lib().orc_halo_prepare_exchange(_p(fdata), C.c_int(si), C.c_int(fdata.size), C.byref(d0), C.byref(c2))
    return d0.value, c2.value


def halo_finish_boundary_conditions(fdata, si, d0, c2):
    fdata = _f(fdata); d0 = C.c_double(d0); c2 = C.c_double(c2)
    lib().orc_halo_finish_boundary_conditions(_p(fdata), C.c_int(si), C.c_int(fdata.size), C.byref(d0), C.byref(c2))
    return d0.value, c2.value


def halo_compute_interpolant(fin, np_):
    fin = _f(fin); d = np.zeros(np_ + 3); coeffs = np.zeros(np_ + 3)
    lib().orc_halo_compute_interpolant(_p(fin), C.c_int(np_), _p(d), _p(coeffs))
    return coeffs


def halo_eval_disp(coeffs, alpha, np_):
    coeffs = _f(coeffs); out = np.zeros(np_)
    lib().orc_halo_eval_disp(_p(coeffs), C.c_double(alpha), C.c_int(np_), _p(out))
    return out


def halo_periodic(data, alpha):
    """sll_s_cubic_spline_halo_1d_periodic on the n periodic cells of `data`"""
    n = data.size
    fin = np.zeros(n + 3); fin[:n] = data
    out = np.zeros(n)
    lib().orc_halo_periodic(_p(fin), C.c_double(alpha), C.c_int(n), _p(out))
    return out


def make_blocks_spline(disp):
    disp = _f(disp); n = disp.size
    shift = np.zeros(n, dtype=np.int32); alpha = np.zeros(n)
    lib().orc_make_blocks_spline.restype = C.c_int
    nb = lib().orc_make_blocks_spline(C.c_int(n), _p(disp), shift.ctypes.data_as(ip), _p(alpha))
    return shift, alpha, nb


def spline_dd_advect_axis(f, axis, nblk, disp, dsel, shifts=None):
    """Local-spline pass along `axis` of the Fortran-ordered array f (in place) with the axis cut into nblk ring
    pieces; disp / shifts indexed like advect_axis."""
    assert f.flags.f_contiguous
    shape = f.shape
    inner = int(np.prod(shape[:axis], dtype=np.int64))
    outer = int(np.prod(shape[axis + 1:], dtype=np.int64))
    disp = _f(disp)
    L = C.c_long
    sh = None
    if shifts is not None:
        sh = np.ascontiguousarray(shifts, dtype=np.int32)
    lib().orc_spline_dd_advect_axis.restype = C.c_int
    rc = lib().orc_spline_dd_advect_axis(_p(f), L(outer), C.c_int(shape[axis]), L(inner), C.c_int(nblk), _p(disp),
                                         sh.ctypes.data_as(ip) if sh is not None else None, *[L(int(v)) for v in dsel])
    if rc != 0:
        raise ValueError("local spline: too few points per piece for the 15-term boundary series (np > 15, np >= 16 - si, "
                         "np >= 17 + si)")
    return f


# ---- cubic splines with Hermite boundary conditions (sll_m_cubic_splines, sll_p_hermite) ----
def hermite_coeffs(data, xmin, xmax, fast=1, slopes=None):
    data = _f(data); c = np.zeros(data.size + 4)
    hs = slopes is not None
    lib().orc_spline_hermite_compute_interpolant(_p(data), C.c_int(data.size), C.c_double(xmin), C.c_double(xmax), C.c_int(fast),
                                                 C.c_int(hs), C.c_double(slopes[0] if hs else 0.0), C.c_int(hs),
                                                 C.c_double(slopes[1] if hs else 0.0), _p(c))
    return c[:data.size + 3]


def spline_eval_array(coeffs, npts, xmin, xmax, x):
    x = _f(np.atleast_1d(x)); out = np.zeros(x.size); coeffs = _f(coeffs)
    lib().orc_spline_eval_array(_p(coeffs), C.c_int(npts), C.c_double(xmin), C.c_double(xmax), _p(x), C.c_int(x.size), _p(out))
    return out


def hermite_interpolate_array_disp(data, xmin, xmax, alpha, fast=1, slopes=None, inplace=False):
    """sll_t_cubic_spline_interpolator_1d with sll_p_hermite: interpolate_array_disp / _inplace"""
    data = _f(data).copy(); out = np.zeros_like(data)
    hs = slopes is not None
    sl, sr = (slopes if hs else (0.0, 0.0))
    if inplace:
        lib().orc_hermite_interpolate_array_disp_inplace(C.c_int(data.size), C.c_double(xmin), C.c_double(xmax), C.c_int(fast),
                                                         C.c_int(hs), C.c_double(sl), C.c_int(hs), C.c_double(sr), _p(data),
                                                         C.c_double(alpha))
        return data
    lib().orc_hermite_interpolate_array_disp(C.c_int(data.size), C.c_double(xmin), C.c_double(xmax), C.c_int(fast), C.c_int(hs),
                                             C.c_double(sl), C.c_int(hs), C.c_double(sr), _p(data), C.c_double(alpha), _p(out))
    return out


def hermite_advect_axis(f, axis, xmin, xmax, disp, dsel, fast=1, inplace=True):
    """every line of the Fortran-ordered f along `axis`: Hermite-spline interpolation at x_i + disp (physical units)"""
    assert f.flags.f_contiguous
    shape = f.shape
    inner = int(np.prod(shape[:axis], dtype=np.int64))
    outer = int(np.prod(shape[axis + 1:], dtype=np.int64))
    disp = _f(disp)
    L = C.c_long
    lib().orc_hermite_advect_axis(_p(f), L(outer), C.c_int(shape[axis]), L(inner), C.c_double(xmin), C.c_double(xmax),
                                  C.c_int(fast), C.c_int(1 if inplace else 0), _p(disp), *[L(int(v)) for v in dsel])
    return f
