"""ctypes binding of include/sll_b200.h (test / bench harness side of the C ABI)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "lib", "libsllb200.so")

METHOD_SPLINE, METHOD_LAGRANGE_FIXED, METHOD_LAGRANGE_CENTERED = 0, 1, 2
ADVECTOR_FIXED, ADVECTOR_CENTERED, ADVECTOR_SPLINE = 0, 1, 2
SHIFT_SKIP = -2 ** 31
ADV_PERIODIC_SPLINE, ADV_PERIODIC_LAGRANGE, ADV_BSL = 0, 1, 2
(INTERP_CUBIC_SPLINE, INTERP_LAGRANGE_CENTERED, INTERP_LAGRANGE_FIXED, INTERP_PERIODIC_SPLINE,
 INTERP_PERIODIC_LAGRANGE) = range(5)
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = 1, 2, 3, 4
BC_PERIODIC, BC_HERMITE = 0, 1

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
vp = C.c_void_p


class SllbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sllb error {code}: {msg}")
        self.code = code


class DispT(C.Structure):
    _fields_ = [("values", dp), ("nvalues", C.c_int64), ("values_on_device", C.c_int), ("scale", C.c_double),
                ("odiv", C.c_int64), ("omod", C.c_int64), ("ostr", C.c_int64),
                ("idiv", C.c_int64), ("imod", C.c_int64), ("istr", C.c_int64)]


class Sim4dParams(C.Structure):
    _fields_ = [("nc", C.c_int * 4), ("xmin", C.c_double * 4), ("xmax", C.c_double * 4),
                ("kx1", C.c_double), ("kx2", C.c_double), ("eps", C.c_double), ("dt", C.c_double),
                ("split", C.c_int), ("method", C.c_int), ("order", C.c_int), ("stencil_r", C.c_int), ("stencil_s", C.c_int),
                ("method_axis", C.c_int * 4), ("order_axis", C.c_int * 4), ("dup_velocity_planes", C.c_int)]


class Sim6dParams(C.Structure):
    _fields_ = [("n", C.c_int * 6), ("v_max", C.c_double), ("x_max", C.c_double * 3),
                ("stencil_x", C.c_int), ("stencil_v", C.c_int), ("delta_t", C.c_double),
                ("alpha", C.c_double), ("kx", C.c_double * 3), ("v_thermal", C.c_double * 3),
                ("time_in_phase", C.c_int), ("advector", C.c_int)]


_LIB = None


def lib():
    """Load the CUDA library; fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(selalib_b200 has no CPU fallback)")
        try:
            # torch bundles a newer NCCL under the same soname (libnccl.so.2); it has to be the one that gets
            # loaded, otherwise a later `import torch` in this process fails to resolve its NCCL symbols
            import torch  # noqa: F401
        except Exception:
            pass
        _LIB = C.CDLL(_SO, mode=C.RTLD_GLOBAL)
        _LIB.sllb_last_error.restype = C.c_char_p
        _LIB.sllb_launch_count.restype = C.c_int64
    return _LIB


def last_error():
    return lib().sllb_last_error().decode()


def _ck(rc):
    if rc != 0:
        raise SllbError(rc, last_error())


def _p(a):
    return a.ctypes.data_as(dp)


def _ints(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


def init(device=0):
    _ck(lib().sllb_init(C.c_int(device)))


def device_count():
    n = C.c_int(0)
    _ck(lib().sllb_device_count(C.byref(n)))
    return n.value


def synchronize():
    _ck(lib().sllb_synchronize())


def launch_count():
    return int(lib().sllb_launch_count())


def launch_count_reset():
    lib().sllb_launch_count_reset()


def set_staging(mode):
    _ck(lib().sllb_set_staging(C.c_int(mode)))


def set_plane_kernel(on, points_per_thread=0):
    _ck(lib().sllb_set_plane_kernel(C.c_int(1 if on else 0), C.c_int(points_per_thread)))


def set_plane_variant(tmem_accumulators=-1, const_extents=1):
    _ck(lib().sllb_set_plane_variant(C.c_int(tmem_accumulators), C.c_int(const_extents)))


def set_fused_remap(on):
    _ck(lib().sllb_set_fused_remap(C.c_int(1 if on else 0)))


def set_remap_rotation(on):
    _ck(lib().sllb_set_remap_rotation(C.c_int(1 if on else 0)))


def dd6d_set_exchange_timing(on):
    _ck(lib().sllb_dd6d_set_exchange_timing(C.c_int(1 if on else 0)))


def set_v_overlap(on):
    _ck(lib().sllb_set_v_overlap(C.c_int(int(on))))


def set_poisson_direct(on):
    """2D Poisson on small grids: 1 = three dense-DFT kernels (default), 0 = cuFFT"""
    _ck(lib().sllb_set_poisson_direct(C.c_int(1 if on else 0)))


def set_phase_timers(on):
    """opt-in per-phase CUDA events inside sllb_sim4d_run (Sim4d.phase_ms8)"""
    _ck(lib().sllb_set_phase_timers(C.c_int(1 if on else 0)))


def set_cuda_graphs(on):
    _ck(lib().sllb_set_cuda_graphs(C.c_int(1 if on else 0)))


def set_spline_split(chunks):
    _ck(lib().sllb_set_spline_split(C.c_int(chunks)))


# ---------------------------------------------------------------------------------------------
# line-granular drop-in objects (mirror sll_t_advector_1d_periodic / sll_c_interpolator_1d)
# ---------------------------------------------------------------------------------------------
class Advector1dPeriodic:
    """sll_t_advector_1d_periodic (sll_m_advection_1d_periodic.F90:41-130); kind=ADV_BSL: sll_t_advector_1d_bsl
    (sll_m_advection_1d_BSL.F90) with explicit-Euler periodic characteristics + cubic-spline interpolator."""

    def __init__(self, num_cells, xmin, xmax, kind=ADV_PERIODIC_SPLINE, order=4):
        self.h = vp()
        _ck(lib().sllb_adv1d_create(C.c_int(kind), C.c_int(num_cells), C.c_double(xmin), C.c_double(xmax),
                                    C.c_int(order), C.byref(self.h)))

    def advect_1d_constant(self, A, dt, inp, out=None):
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        if out is None:
            out = np.empty_like(inp)
        _ck(lib().sllb_adv1d_advect_constant(self.h, C.c_double(A), C.c_double(dt), _p(inp), _p(out), C.c_int(inp.size)))
        return out

    def delete(self):
        if self.h:
            lib().sllb_adv1d_delete(self.h)
            self.h = vp()

    __del__ = delete


class Interpolator1d:
    """sll_c_interpolator_1d implementations (interpolate_array_disp[_inplace])."""

    def __init__(self, kind, num_points, xmin, xmax, d_or_order=4, periodic_last=1, fast_algorithm=1, bc=0):
        self.h = vp()
        _ck(lib().sllb_interp1d_create(C.c_int(kind), C.c_int(num_points), C.c_double(xmin), C.c_double(xmax),
                                       C.c_int(bc), C.c_int(d_or_order), C.c_int(periodic_last),
                                       C.c_int(fast_algorithm), C.byref(self.h)))

    def set_slopes(self, slope_left, slope_right):
        _ck(lib().sllb_interp1d_set_slopes(self.h, C.c_double(slope_left), C.c_double(slope_right)))

    def interpolate_array_disp(self, num_pts, data, alpha):
        data = np.ascontiguousarray(data, dtype=np.float64)
        out = np.empty(num_pts)
        _ck(lib().sllb_interp1d_array_disp(self.h, C.c_int(num_pts), _p(data), C.c_double(alpha), _p(out)))
        return out

    def interpolate_array_disp_inplace(self, num_pts, data, alpha):
        assert data.dtype == np.float64 and data.flags.c_contiguous
        _ck(lib().sllb_interp1d_array_disp_inplace(self.h, C.c_int(num_pts), _p(data), C.c_double(alpha)))
        return data

    def delete(self):
        if self.h:
            lib().sllb_interp1d_delete(self.h)
            self.h = vp()

    __del__ = delete


# ---------------------------------------------------------------------------------------------
# batched device-resident API
# ---------------------------------------------------------------------------------------------
def _disp(owner, values, scale, dsel, on_device=False):
    d = DispT()
    if on_device:
        d.values = C.cast(vp(values), dp); d.nvalues = 0; d.values_on_device = 1
    else:
        values = np.ascontiguousarray(values, dtype=np.float64)
        owner._keep = values
        d.values = _p(values); d.nvalues = values.size; d.values_on_device = 0
    d.scale = scale
    d.odiv, d.omod, d.ostr, d.idiv, d.imod, d.istr = [int(v) for v in dsel]
    return d


def spline_dd_blocks(disp):
    """make_blocks_spline: (shift int32[n], alpha[n], nblocks)"""
    disp = np.ascontiguousarray(disp, dtype=np.float64)
    shift = np.zeros(disp.size, dtype=np.int32); alpha = np.zeros(disp.size); nb = C.c_int(0)
    _ck(lib().sllb_spline_dd_blocks(C.c_int(disp.size), _p(disp), shift.ctypes.data_as(C.POINTER(C.c_int32)), _p(alpha), C.byref(nb)))
    return shift, alpha, nb.value


def lagrange_dd_blocks(disp, stencil):
    """make_blocks_lagrange: (box int32[n], nblocks, halo widths (nblocks, 2))"""
    disp = np.ascontiguousarray(disp, dtype=np.float64)
    box = np.zeros(disp.size, dtype=np.int32); nb = C.c_int(0); hw = (C.c_int * (2 * stencil + 2))()
    _ck(lib().sllb_lagrange_dd_blocks(C.c_int(disp.size), C.c_int(stencil), _p(disp), box.ctypes.data_as(C.POINTER(C.c_int32)),
                                      C.byref(nb), hw))
    return box, nb.value, np.array(hw[:2 * nb.value]).reshape(-1, 2)


class Field:
    def __init__(self, extents=None, handle=None):
        self.owns = handle is None
        if handle is None:
            self.h = vp()
            _ck(lib().sllb_field_create(C.c_int(len(extents)), _ints(extents), C.byref(self.h)))
        else:
            self.h = handle
        nd = C.c_int(0)
        ext = (C.c_int * 6)()
        _ck(lib().sllb_field_extents(self.h, C.byref(nd), ext))
        self.extents = tuple(ext[i] for i in range(nd.value))

    def upload(self, host, dup_last=None):
        host = np.asfortranarray(host, dtype=np.float64)
        dup = _ints(dup_last) if dup_last is not None else None
        exp = tuple(e + (dup_last[i] if dup_last is not None else 0) for i, e in enumerate(self.extents))
        assert host.shape == exp, (host.shape, exp)
        _ck(lib().sllb_field_upload(self.h, _p(host), dup))
        return self

    def download(self, dup_last=None):
        exp = tuple(e + (dup_last[i] if dup_last is not None else 0) for i, e in enumerate(self.extents))
        out = np.empty(exp, order="F")
        _ck(lib().sllb_field_download(self.h, _p(out), _ints(dup_last) if dup_last is not None else None))
        return out

    def device_ptr(self):
        p = dp()
        _ck(lib().sllb_field_device_ptr(self.h, C.byref(p)))
        return C.cast(p, vp).value

    def advect_axis(self, axis, method, order, values, scale=1.0, dsel=(1, 1, 0, 1, 1, 0), on_device=False):
        d = DispT()
        if on_device:
            d.values = C.cast(vp(values), dp); d.nvalues = 0; d.values_on_device = 1
        else:
            values = np.ascontiguousarray(values, dtype=np.float64)
            self._keep = values
            d.values = _p(values); d.nvalues = values.size; d.values_on_device = 0
        d.scale = scale
        d.odiv, d.omod, d.ostr, d.idiv, d.imod, d.istr = [int(v) for v in dsel]
        _ck(lib().sllb_advect_axis(self.h, C.c_int(axis), C.c_int(method), C.c_int(order), C.byref(d)))

    def advect_axis_spline_dd(self, axis, values, dsel=(1, 1, 0, 1, 1, 0), scale=1.0, shift=None):
        """local cubic spline (NUM_TERMS = 15) along an unsplit axis; shift: int32 table indexed like values or None"""
        d = _disp(self, values, scale, dsel)
        sh = None
        if shift is not None:
            sh = np.ascontiguousarray(shift, dtype=np.int32)
            assert sh.size == d.nvalues
        _ck(lib().sllb_advect_axis_spline_dd(self.h, C.c_int(axis), C.byref(d), sh.ctypes.data_as(C.POINTER(C.c_int32)) if sh is not None else None))

    def advect_axis_hermite(self, axis, xmin, xmax, values, dsel=(1, 1, 0, 1, 1, 0), scale=1.0, inplace=True):
        """Hermite-BC cubic spline along a non-periodic axis; values = displacements in physical units"""
        d = _disp(self, values, scale, dsel)
        _ck(lib().sllb_advect_axis_hermite(self.h, C.c_int(axis), C.c_double(xmin), C.c_double(xmax), C.byref(d), C.c_int(1 if inplace else 0)))

    def advect_plane(self, values0, dsel0, scale0, values1, dsel1, scale1, rho_scale=None, method=METHOD_SPLINE, order=4):
        """K1c: spline passes along axes 0 and 1 in one sweep; returns rho (host) when rho_scale is given."""
        ds = []
        keep = []
        for values, dsel, scale in ((values0, dsel0, scale0), (values1, dsel1, scale1)):
            d = DispT()
            values = np.ascontiguousarray(values, dtype=np.float64)
            keep.append(values)
            d.values = _p(values); d.nvalues = values.size; d.values_on_device = 0; d.scale = scale
            d.odiv, d.omod, d.ostr, d.idiv, d.imod, d.istr = [int(v) for v in dsel]
            ds.append(d)
        rho_dev = None
        if rho_scale is not None:
            import torch
            rho_dev = torch.empty(self.extents[0] * self.extents[1], dtype=torch.float64, device="cuda")
        _ck(lib().sllb_advect_plane(self.h, C.c_int(method), C.c_int(order), C.byref(ds[0]), C.byref(ds[1]),
                                    C.c_double(rho_scale if rho_scale is not None else 0.0),
                                    C.cast(vp(rho_dev.data_ptr()), dp) if rho_dev is not None else None))
        if rho_dev is not None:
            return rho_dev.cpu().numpy().reshape(self.extents[:2], order="F")
        return None

    def advect_axis_affine(self, axis, method, order, v_axis, vmin, dv, scale):
        _ck(lib().sllb_advect_axis_affine(self.h, C.c_int(axis), C.c_int(method), C.c_int(order), C.c_int(v_axis),
                                          C.c_double(vmin), C.c_double(dv), C.c_double(scale)))

    def advect_axis_field(self, axis, method, order, d_field_ptr, nfield_axes, scale):
        _ck(lib().sllb_advect_axis_field(self.h, C.c_int(axis), C.c_int(method), C.c_int(order),
                                         C.cast(vp(d_field_ptr), dp), C.c_int(nfield_axes), C.c_double(scale)))

    def reduce_velocity(self, nx_axes, scale):
        out = np.empty(self.extents[:nx_axes], order="F")
        _ck(lib().sllb_reduce_velocity_host(self.h, C.c_int(nx_axes), C.c_double(scale), _p(out)))
        return out

    def moments(self, nv, w1, w2):
        out = np.zeros(3 + 2 * nv)
        w1 = np.ascontiguousarray(w1, dtype=np.float64); w2 = np.ascontiguousarray(w2, dtype=np.float64)
        _ck(lib().sllb_moments(self.h, C.c_int(nv), _p(w1), _p(w2), _p(out)))
        return out

    def destroy(self):
        if self.owns and self.h:
            lib().sllb_field_destroy(self.h)
            self.h = vp()

    __del__ = destroy


class Poisson:
    def __init__(self, n, xmin, xmax, par=False):
        """par=True (2D): sll_t_poisson_2d_periodic_par, Delta phi = rho, potential only"""
        self.h = vp()
        self.n = tuple(int(v) for v in n)
        self.par = bool(par)
        if par:
            assert len(n) == 2
            _ck(lib().sllb_poisson2d_par_create(C.c_int(n[0]), C.c_int(n[1]), C.c_double(xmax[0] - xmin[0]),
                                                C.c_double(xmax[1] - xmin[1]), C.byref(self.h)))
        elif len(n) == 1:
            _ck(lib().sllb_poisson1d_create(C.c_int(n[0]), C.c_double(xmin[0]), C.c_double(xmax[0]), C.byref(self.h)))
        elif len(n) == 2:
            _ck(lib().sllb_poisson2d_create(C.c_int(n[0]), C.c_int(n[1]), C.c_double(xmin[0]), C.c_double(xmax[0]),
                                            C.c_double(xmin[1]), C.c_double(xmax[1]), C.byref(self.h)))
        else:
            _ck(lib().sllb_poisson3d_create(C.c_int(n[0]), C.c_int(n[1]), C.c_int(n[2]), C.c_double(xmax[0] - xmin[0]),
                                            C.c_double(xmax[1] - xmin[1]), C.c_double(xmax[2] - xmin[2]), C.byref(self.h)))

    def solve(self, rho, want_phi=True):
        """rho: Fortran-ordered host array with nc or nc+1 points per axis. Returns (phi, e1[, e2[, e3]])."""
        rho = np.asfortranarray(rho, dtype=np.float64)
        dim = len(self.n)
        ld = _ints(rho.shape)
        outs = [np.zeros_like(rho, order="F") for _ in range(4)]
        ptrs = [_p(o) for o in outs]
        if not want_phi or dim == 1:
            ptrs[0] = None
        for k in range(dim + 1, 4):
            ptrs[k] = None
        if self.par:
            ptrs[1] = ptrs[2] = None
        _ck(lib().sllb_poisson_solve_host(self.h, _p(rho), ld, *ptrs))
        return outs[0] if self.par else tuple(outs[: dim + 1])

    def destroy(self):
        if self.h:
            lib().sllb_poisson_destroy(self.h)
            self.h = vp()

    __del__ = destroy


# ---------------------------------------------------------------------------------------------
# layouts / remap plans (host logic, no device needed)
# ---------------------------------------------------------------------------------------------
def factorize_in_two_powers_of_two(n):
    a, b = C.c_int(0), C.c_int(0)
    _ck(lib().sllb_factorize_in_two_powers_of_two(C.c_int(n), C.byref(a), C.byref(b)))
    return a.value, b.value


def layout4d_boxes(global_ext, procs, nranks):
    out = (C.c_int * (8 * nranks))()
    _ck(lib().sllb_layout4d_boxes(_ints(global_ext), _ints(procs), C.c_int(nranks), out))
    return np.array(out[:]).reshape(nranks, 4, 2)


def remap4d_plan(global_ext, procs_from, procs_to, nranks, rank):
    s = (C.c_int * (8 * nranks))(); r = (C.c_int * (8 * nranks))()
    _ck(lib().sllb_remap4d_plan(_ints(global_ext), _ints(procs_from), _ints(procs_to), C.c_int(nranks), C.c_int(rank), s, r))
    return np.array(s[:]).reshape(nranks, 4, 2), np.array(r[:]).reshape(nranks, 4, 2)


def set_process_grid(nranks):
    g = (C.c_int * 6)()
    _ck(lib().sllb_set_process_grid(C.c_int(nranks), g))
    return tuple(g[:])


class Comm:
    """NCCL communicator, one process per GPU.  The 128-byte id travels over torch.distributed."""

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        _ck(lib().sllb_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, id_bytes, nranks, rank):
        self.h = vp()
        self.nranks, self.rank = nranks, rank
        buf = (C.c_ubyte * 128)(*id_bytes)
        _ck(lib().sllb_comm_create(buf, C.c_int(nranks), C.c_int(rank), C.byref(self.h)))

    def destroy(self):
        if self.h:
            lib().sllb_comm_destroy(self.h)
            self.h = vp()


class Dist4d:
    def __init__(self, comm, global_ext):
        self.h = vp()
        _ck(lib().sllb_dist4d_create(comm.h if comm is not None else None, _ints(global_ext), C.byref(self.h)))

    def field(self, which):
        f = vp()
        _ck(lib().sllb_dist4d_field(self.h, C.c_int(which), C.byref(f)))
        return Field(handle=f)

    def box(self, which):
        b = (C.c_int * 8)()
        _ck(lib().sllb_dist4d_box(self.h, C.c_int(which), b))
        return np.array(b[:]).reshape(4, 2)

    def remap(self, direction):
        _ck(lib().sllb_dist4d_remap(self.h, C.c_int(direction)))

    def p2p(self):
        e = C.c_int(0)
        _ck(lib().sllb_dist4d_p2p(self.h, C.byref(e)))
        return bool(e.value)

    def advect_remap(self, src, axis, method, order, values_ptr, scale=1.0, dsel=(1, 1, 0, 1, 1, 0)):
        """Fused pass: advect layout `src` along `axis`, result lands in the other layout (device disp values)."""
        d = DispT()
        d.values = C.cast(vp(values_ptr), dp); d.nvalues = 0; d.values_on_device = 1
        d.scale = scale
        d.odiv, d.omod, d.ostr, d.idiv, d.imod, d.istr = [int(v) for v in dsel]
        _ck(lib().sllb_dist4d_advect_remap(self.h, C.c_int(src), C.c_int(axis), C.c_int(method), C.c_int(order), C.byref(d)))

    def destroy(self):
        if self.h:
            lib().sllb_dist4d_destroy(self.h)
            self.h = vp()


# ---------------------------------------------------------------------------------------------
# simulations
# ---------------------------------------------------------------------------------------------
def splitting_case(name):
    """namelist split_case string -> SLLB_SPLIT_* number"""
    k = C.c_int(-1)
    _ck(lib().sllb_splitting_case_from_name(name.encode(), C.byref(k)))
    return k.value


def splitting_coeff(split, dt):
    """(split_step array, split_begin_T, dim_split_V, nb_split_step) of sll_f_new_time_splitting_coeff"""
    if isinstance(split, str):
        split = splitting_case(split)
    steps = np.zeros(32); n = C.c_int(); nb = C.c_int(); bt = C.c_int(); dv = C.c_int()
    _ck(lib().sllb_splitting_coeff(C.c_int(split), C.c_double(dt), _p(steps), C.byref(n), C.byref(nb), C.byref(bt), C.byref(dv)))
    return steps[:n.value].copy(), bool(bt.value), dv.value, nb.value


def compute_w_hermite(r, s):
    w = np.zeros(s - r + 1)
    _ck(lib().sllb_compute_w_hermite(C.c_int(r), C.c_int(s), _p(w)))
    return w


def format_g20_12(x):
    buf = C.create_string_buffer(24)
    _ck(lib().sllb_format_g20_12(C.c_double(x), buf))
    return buf.value.decode()


def sim4d_run_namelist(filename, thdiag_path, comm=None):
    """sim_bsl_vp_2d2v_cart_poisson_serial <filename>: run the namelist, write the thdiag file"""
    _ck(lib().sllb_sim4d_run_namelist(filename.encode(), comm.h if comm is not None else None, thdiag_path.encode()))


def sim2d_run_namelist(filename, outdir=None):
    """sim_bsl_vp_1d1v_cart <filename>: run the namelist, write thdiag.dat, the .bdat files and the restart files"""
    _ck(lib().sllb_sim2d_run_namelist(filename.encode(), outdir.encode() if outdir else None))


def format_g(x, w, d):
    buf = C.create_string_buffer(w + 1)
    _ck(lib().sllb_format_g(C.c_double(x), C.c_int(w), C.c_int(d), buf))
    return buf.value.decode()


class Sim4d:
    def __init__(self, nc, xmin, xmax, kx1, kx2, eps, dt, split=0, method=METHOD_SPLINE, order=4, comm=None, stencil=(0, 0),
                 dup_velocity_planes=False):
        if isinstance(split, str):
            split = splitting_case(split)
        p = Sim4dParams()
        p.dup_velocity_planes = 1 if dup_velocity_planes else 0
        p.nc[:] = nc; p.xmin[:] = xmin; p.xmax[:] = xmax
        p.kx1, p.kx2, p.eps, p.dt = kx1, kx2, eps, dt
        p.split = split
        p.stencil_r, p.stencil_s = stencil
        if isinstance(method, (tuple, list)):      # per-axis advectors (advector_x1..x4 / order_x1..x4)
            orders = order if isinstance(order, (tuple, list)) else [order] * 4
            p.method_axis[:] = list(method); p.order_axis[:] = list(orders)
            p.method, p.order = method[0], orders[0]
        else:
            p.method, p.order = method, order
        self.h = vp()
        self.nc = tuple(int(c) for c in nc)
        _ck(lib().sllb_sim4d_create(C.byref(p), comm.h if comm is not None else None, C.byref(self.h)))

    def run(self, nsteps, diagnostics=True):
        rows = np.zeros((nsteps, 6))
        _ck(lib().sllb_sim4d_run(self.h, C.c_int(nsteps), C.c_int(1 if diagnostics else 0), _p(rows) if diagnostics else None))
        return rows

    def stream_step(self, host_next_in=None, host_prev_out=None):
        """ensemble streaming: arguments are raw host pointers (ints) or None"""
        _ck(lib().sllb_sim4d_stream_step(self.h, C.cast(vp(host_next_in), dp) if host_next_in else None,
                                         C.cast(vp(host_prev_out), dp) if host_prev_out else None))

    def thdiag(self):
        row = np.zeros(13)
        _ck(lib().sllb_sim4d_thdiag(self.h, _p(row)))
        return row

    def diagnostics(self):
        row = np.zeros(6)
        _ck(lib().sllb_sim4d_diagnostics(self.h, _p(row)))
        return row

    def field(self):
        f = vp()
        _ck(lib().sllb_sim4d_field(self.h, C.byref(f)))
        return Field(handle=f)

    def box(self, which):
        b = (C.c_int * 8)()
        _ck(lib().sllb_sim4d_box(self.h, C.c_int(which), b))
        return np.array(b[:]).reshape(4, 2)

    def phase_ms(self):
        out = np.zeros(6)
        _ck(lib().sllb_sim4d_phase_ms6(self.h, _p(out)))
        return out

    def fields(self):
        """rho, E1, E2 of the last field solve (N1 x N2 periodic cells each)"""
        n1, n2 = self.nc[0], self.nc[1]
        out = [np.zeros((n1, n2), order="F") for _ in range(3)]
        _ck(lib().sllb_sim4d_fields_host(self.h, _p(out[0]), _p(out[1]), _p(out[2])))
        return out

    def checksum(self):
        """(sum w f, sum w f^2) over the global field, w from the global index: equal on any number of ranks"""
        out = np.zeros(2)
        _ck(lib().sllb_sim4d_checksum(self.h, _p(out)))
        return out

    def phase_ms8(self):
        out = np.zeros(8)
        _ck(lib().sllb_sim4d_phase_ms8(self.h, _p(out)))
        return out

    def destroy(self):
        if self.h:
            lib().sllb_sim4d_destroy(self.h)
            self.h = vp()


class Sim2d:
    def __init__(self, nc_x1, nc_x2, x1_min, x1_max, x2_min, x2_max, init, kmode, eps, dt, method=METHOD_SPLINE, order=4):
        self.h = vp()
        _ck(lib().sllb_sim2d_create(C.c_int(nc_x1), C.c_int(nc_x2), C.c_double(x1_min), C.c_double(x1_max),
                                    C.c_double(x2_min), C.c_double(x2_max), C.c_int(init), C.c_double(kmode),
                                    C.c_double(eps), C.c_double(dt), C.c_int(method), C.c_int(order), C.byref(self.h)))

    @classmethod
    def from_namelist(cls, filename):
        """the namelist of sim_bsl_vp_1d1v_cart; returns (sim, number_iterations, freq_diag_time, nb_mode)"""
        self = cls.__new__(cls)
        self.h = vp()
        nit, fdt, nbm = C.c_int(0), C.c_int(0), C.c_int(0)
        _ck(lib().sllb_sim2d_create_from_namelist(filename.encode(), C.byref(self.h), C.byref(nit), C.byref(fdt), C.byref(nbm)))
        return self, nit.value, fdt.value, nbm.value

    def run(self, nsteps, diagnostics=True):
        rows = np.zeros((nsteps, 8))
        _ck(lib().sllb_sim2d_run(self.h, C.c_int(nsteps), _p(rows) if diagnostics else None))
        return rows

    def field(self):
        f = vp()
        _ck(lib().sllb_sim2d_field(self.h, C.byref(f)))
        return Field(handle=f)

    def set_splitting(self, split):
        if isinstance(split, str):
            split = splitting_case(split)
        _ck(lib().sllb_sim2d_set_splitting(self.h, C.c_int(split)))

    def set_advectors(self, method_x1, order_x1, method_x2, order_x2):
        _ck(lib().sllb_sim2d_set_advectors(self.h, C.c_int(method_x1), C.c_int(order_x1), C.c_int(method_x2), C.c_int(order_x2)))

    def set_time(self, t):
        _ck(lib().sllb_sim2d_set_time(self.h, C.c_double(t)))

    def fields(self):
        """rho and E of the current state (N1 periodic cells each)"""
        n1 = self.field().extents[0]
        rho, e = np.zeros(n1), np.zeros(n1)
        _ck(lib().sllb_sim2d_fields_host(self.h, _p(rho), _p(e)))
        return rho, e

    def thdiag(self, nb_mode):
        """one row of the reference's thdiag.dat: 8 integrals, Re/Im of rho^_k, f_hat_x2(k), k = 0..nb_mode"""
        row = np.zeros(8 + 3 * (nb_mode + 1))
        _ck(lib().sllb_sim2d_thdiag(self.h, C.c_int(nb_mode), _p(row)))
        return row

    def write_restart(self, path):
        _ck(lib().sllb_sim2d_write_restart(self.h, path.encode()))

    def read_restart(self, path):
        t = C.c_double(0.0)
        _ck(lib().sllb_sim2d_read_restart(self.h, path.encode(), C.byref(t)))
        return t.value

    def destroy(self):
        if self.h:
            lib().sllb_sim2d_destroy(self.h)
            self.h = vp()


class Dd6d:
    """sll_t_decomposition_slim_6d + halo exchange (sll_m_decomposition.F90:835-869,1715-2030)."""

    def __init__(self, comm, global_ext, procs=None):
        self.h = vp()
        _ck(lib().sllb_dd6d_create(comm.h if comm is not None else None, _ints(global_ext),
                                   _ints(procs) if procs is not None else None, C.byref(self.h)))
        arrs = [(C.c_int * 6)() for _ in range(6)]
        _ck(lib().sllb_dd6d_layout(self.h, *arrs))
        self.procs, self.coords, self.mn, self.nw, self.left, self.right = [tuple(a[:]) for a in arrs]

    def field(self):
        f = vp()
        _ck(lib().sllb_dd6d_field(self.h, C.byref(f)))
        return Field(handle=f)

    def p2p(self):
        e = C.c_int(0)
        _ck(lib().sllb_dd6d_p2p(self.h, C.byref(e)))
        return bool(e.value)

    def halo_exchange(self, axis, hw_left, hw_right):
        _ck(lib().sllb_dd6d_halo_exchange(self.h, C.c_int(axis), C.c_int(hw_left), C.c_int(hw_right)))
        self._halo = (axis, hw_left, hw_right)

    def halo(self, side):
        axis, hl, hr = self._halo
        shp = list(self.nw); shp[axis] = hl if side == 0 else hr
        out = np.empty(shp, order="F")
        if out.size:
            _ck(lib().sllb_dd6d_halo_download(self.h, C.c_int(side), _p(out)))
        return out

    def exchange_ms(self):
        ms = C.c_double(0)
        _ck(lib().sllb_dd6d_exchange_ms(self.h, C.byref(ms)))
        return ms.value

    def advect_axis(self, axis, stencil, values, scale=1.0, dsel=(1, 1, 0, 1, 1, 0), on_device=False):
        d = DispT()
        if on_device:
            d.values = C.cast(vp(values), dp); d.nvalues = 0; d.values_on_device = 1
        else:
            values = np.ascontiguousarray(values, dtype=np.float64)
            self._keep = values
            d.values = _p(values); d.nvalues = values.size; d.values_on_device = 0
        d.scale = scale
        d.odiv, d.omod, d.ostr, d.idiv, d.imod, d.istr = [int(v) for v in dsel]
        _ck(lib().sllb_dd6d_advect_axis(self.h, C.c_int(axis), C.c_int(stencil), C.byref(d)))

    def advect_axis_spline(self, axis, values, scale=1.0, dsel=(1, 1, 0, 1, 1, 0), shift=None, hw=(1, 1), on_device=False):
        d = _disp(self, values, scale, dsel, on_device)
        sh = None
        if shift is not None:
            sh = np.ascontiguousarray(shift, dtype=np.int32)
        _ck(lib().sllb_dd6d_advect_axis_spline(self.h, C.c_int(axis), C.byref(d),
                                               sh.ctypes.data_as(C.POINTER(C.c_int32)) if sh is not None else None,
                                               C.c_int(hw[0]), C.c_int(hw[1])))

    def destroy(self):
        if self.h:
            lib().sllb_dd6d_destroy(self.h)
            self.h = vp()


def dd6d_plan(nranks, rank, global_ext, procs=None):
    """Host-only decomposition of `rank`: dict(procs, coords, mn, nw, left, right)."""
    arrs = [(C.c_int * 6)() for _ in range(6)]
    _ck(lib().sllb_dd6d_plan(C.c_int(nranks), C.c_int(rank), _ints(global_ext), _ints(procs) if procs is not None else None, *arrs))
    return dict(zip(("procs", "coords", "mn", "nw", "left", "right"), [tuple(a[:]) for a in arrs]))


def dd6d_set_halo_p2p(on):
    _ck(lib().sllb_dd6d_set_halo_p2p(C.c_int(1 if on else 0)))


def dd6d_set_halo_chunks(chunks):
    _ck(lib().sllb_dd6d_set_halo_chunks(C.c_int(chunks)))


def dd6d_set_force_halo(on):
    _ck(lib().sllb_dd6d_set_force_halo(C.c_int(1 if on else 0)))


class Sim6d:
    def __init__(self, n, v_max, x_max, stencil_x, stencil_v, delta_t, alpha, kx, v_thermal=(1.0, 1.0, 1.0),
                 time_in_phase=True, comm=None, process_grid=None, advector=ADVECTOR_FIXED):
        p = Sim6dParams()
        p.n[:] = n; p.v_max = v_max; p.x_max[:] = x_max
        p.stencil_x, p.stencil_v, p.delta_t = stencil_x, stencil_v, delta_t
        p.alpha = alpha; p.kx[:] = kx; p.v_thermal[:] = v_thermal
        p.time_in_phase = 1 if time_in_phase else 0
        p.advector = advector
        self.h = vp()
        _ck(lib().sllb_sim6d_create_dist(C.byref(p), comm.h if comm is not None else None,
                                         _ints(process_grid) if process_grid is not None else None, C.byref(self.h)))

    def run(self, nsteps, first=None):
        """rows written by this call: the first call on a handle includes the t = 0 row (the library knows which call
        this is; `first` is kept for old callers and ignored)."""
        nrows = C.c_int(0)
        _ck(lib().sllb_sim6d_run_rows(self.h, C.c_int(nsteps), C.byref(nrows)))
        rows = np.zeros((nrows.value, 14))
        _ck(lib().sllb_sim6d_run(self.h, C.c_int(nsteps), _p(rows)))
        return rows

    def field(self):
        f = vp()
        _ck(lib().sllb_sim6d_field(self.h, C.byref(f)))
        return Field(handle=f)

    def layout(self):
        d = vp()
        _ck(lib().sllb_sim6d_decomposition(self.h, C.byref(d)))
        arrs = [(C.c_int * 6)() for _ in range(6)]
        _ck(lib().sllb_dd6d_layout(d, *arrs))
        return dict(zip(("procs", "coords", "mn", "nw", "left", "right"), [tuple(a[:]) for a in arrs]))

    def set_clocks(self, on=True):
        """the reference's stopwatch table (labels P, PC, PF, D, X, X1..X3, V, X4..X6, H4..H6)"""
        _ck(lib().sllb_sim6d_set_clocks(self.h, C.c_int(1 if on else 0)))

    def write_clocks(self, path):
        _ck(lib().sllb_sim6d_write_clocks(self.h, path.encode()))

    def advect_x(self):
        _ck(lib().sllb_sim6d_advect_x(self.h))

    def advect_v(self, dt):
        _ck(lib().sllb_sim6d_advect_v(self.h, C.c_double(dt)))

    def fields(self):
        _ck(lib().sllb_sim6d_fields(self.h))

    def diagnostics(self, time):
        row = np.zeros(14)
        _ck(lib().sllb_sim6d_diagnostics(self.h, C.c_double(time), _p(row)))
        return row

    def halo_ms(self, reset=True):
        ms = C.c_double(0)
        _ck(lib().sllb_sim6d_halo_ms(self.h, C.byref(ms), C.c_int(1 if reset else 0)))
        return ms.value

    def destroy(self):
        if self.h:
            lib().sllb_sim6d_destroy(self.h)
            self.h = vp()
