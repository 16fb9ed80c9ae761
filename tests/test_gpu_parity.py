"""GPU parity tests: every CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Tolerance (north_star): <= 1e-12 relative to max|f| per advection step, fp64."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-12
SEED = 20261017


@pytest.fixture(scope="module")
def sb():
    import selalib_b200 as s
    s.init(0)
    return s


@pytest.fixture(scope="module")
def orc():
    from oracle import orc as o
    return o


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def landau_line(n, L=4 * np.pi, eps=0.05):
    x = np.arange(n) * L / n
    return (1 + eps * np.cos(0.5 * x)) * np.exp(-0.5 * ((x - L / 2) / 1.3) ** 2)


# ---------------------------------------------------------------------------------------------
# a1/a2: advect_1d_constant, line by line
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [32, 64, 128, 256, 1024])
@pytest.mark.parametrize("data", ["random", "landau"])
def test_advector_periodic_spline(sb, orc, n, data):
    rng = np.random.default_rng(SEED + n)
    xmin, xmax = 0.0, 4 * np.pi
    adv = sb.Advector1dPeriodic(n, xmin, xmax, sb.ADV_PERIODIC_SPLINE, 4)
    for A, dt in [(0.73, 0.1), (-5.9, 0.1), (0.0, 0.1), (37.3, 0.5), (-1e-9, 1.0)]:
        f = rng.standard_normal(n) if data == "random" else landau_line(n)
        fin = np.append(f, f[0])
        ref = orc.advect_1d_periodic_constant("spline", n, xmin, xmax, 4, A, dt, fin)  # the sims' FFT advector
        ref2 = orc.spline_interpolate_array_disp(fin, xmin, xmax, -A * dt)             # direct spline
        out = adv.advect_1d_constant(A, dt, fin)
        assert out.shape == (n + 1,) and out[-1] == out[0]
        assert relerr(out, ref) < TOL and relerr(out, ref2) < TOL
        out_n = adv.advect_1d_constant(A, dt, f)          # n = num_cells (no duplicate)
        assert np.array_equal(out_n, out[:-1])
        buf = fin.copy()                                   # in/out aliasing as the sims use it
        adv.advect_1d_constant(A, dt, buf, buf)
        assert np.array_equal(buf, out)
    ones = np.ones(n + 1)
    assert np.abs(adv.advect_1d_constant(0.3, 0.1, ones) - 1.0).max() < 2e-15   # test_advection_1d_periodic.F90
    adv.delete()


@pytest.mark.parametrize("order", [4, 6, 8])
def test_advector_periodic_lagrange(sb, orc, order):
    rng = np.random.default_rng(SEED + order)
    n, xmin, xmax = 96, -6.0, 6.0
    adv = sb.Advector1dPeriodic(n, xmin, xmax, sb.ADV_PERIODIC_LAGRANGE, order)
    for A, dt in [(0.4, 0.1), (-3.3, 0.25), (12.0, 0.5)]:
        f = rng.standard_normal(n + 1); f[-1] = f[0]
        ref = orc.advect_1d_periodic_constant("lagrange", n, xmin, xmax, order, A, dt, f)
        assert relerr(adv.advect_1d_constant(A, dt, f), ref) < TOL
    adv.delete()


@pytest.mark.parametrize("n", [32, 64, 200])
def test_advector_bsl(sb, orc, n):
    """a4: sll_t_advector_1d_bsl%advect_1d_constant (explicit-Euler periodic characteristics + cubic-spline
    interpolate_array) against the oracle's restatement of that object chain."""
    rng = np.random.default_rng(SEED + 7 * n)
    xmin, xmax = -1.0, 2.5
    adv = sb.Advector1dPeriodic(n, xmin, xmax, sb.ADV_BSL, 4)
    for A, dt in [(1.0, 0.1), (-0.37, 0.3), (0.0, 0.1), (55.5, 0.2), (3.5 / n, 1.0)]:
        f = rng.standard_normal(n + 1); f[-1] = f[0]
        ref = orc.advect_1d_bsl_constant(n + 1, xmin, xmax, A, dt, f)
        out = adv.advect_1d_constant(A, dt, f)
        assert out.shape == (n + 1,)
        assert relerr(out, ref) < TOL
    ones = np.ones(n + 1)                                  # test_advection_1d_bsl.F90: err < 1e-15 on input = 1
    assert np.abs(adv.advect_1d_constant(1.0, 0.1, ones) - 1.0).max() < 2e-15
    adv.delete()


@pytest.mark.parametrize("order", [6, 8])
@pytest.mark.parametrize("n", [8, 20, 64, 100, 128, 512])
def test_advector_periodic_spline_order_6_8(sb, orc, n, order):
    """a3: sll_s_periodic_interp with sll_p_spline of order 6 / 8 (the reference's test_periodic_interpolation.F90 runs
    order 8): in-block recursive-filter solve == the reference's FFT diagonalisation, short lines included (the periodic
    sums close with 1/(1 - z^N) there)."""
    rng = np.random.default_rng(SEED + 31 * n + order)
    xmin, xmax = 0.0, 4 * np.pi
    adv = sb.Advector1dPeriodic(n, xmin, xmax, sb.ADV_PERIODIC_SPLINE, order)
    for A, dt in [(0.73, 0.1), (-5.9, 0.1), (0.0, 0.1), (37.3, 0.5), (-1e-9, 1.0)]:
        f = rng.standard_normal(n)
        fin = np.append(f, f[0])
        ref = orc.advect_1d_periodic_constant("spline", n, xmin, xmax, order, A, dt, fin)
        out = adv.advect_1d_constant(A, dt, fin)
        assert out.shape == (n + 1,) and out[-1] == out[0]
        assert relerr(out, ref) < TOL, (A, dt, relerr(out, ref))
    ones = np.ones(n + 1)
    assert np.abs(adv.advect_1d_constant(0.3, 0.1, ones) - 1.0).max() < 1e-14
    adv.delete()


@pytest.mark.parametrize("order", [6, 8])
def test_periodic_spline_order_6_8_whole_array_passes(sb, orc, order):
    """the same splines as whole-array passes on every axis of a 4D field (strided tiles and the contiguous axis)"""
    rng = np.random.default_rng(SEED + order)
    shape = (32, 16, 24, 40)
    f0 = np.asfortranarray(rng.standard_normal(shape))
    F = sb.Field(shape)
    for axis in range(4):
        v_axis = (axis + 2) % 4
        disp = rng.uniform(-2.5, 2.5, shape[v_axis])
        if v_axis > axis:
            stride = int(np.prod(shape[axis + 1:v_axis], dtype=np.int64))
            dsel = (stride, shape[v_axis], 1, 1, 1, 0)
        else:
            dsel = (1, 1, 0, int(np.prod(shape[:v_axis], dtype=np.int64)), shape[v_axis], 1)
        ref = orc.advect_axis(f0.copy(order="F"), axis, "fft_spline", order, disp, dsel)
        F.upload(f0)
        F.advect_axis(axis, sb.METHOD_SPLINE, order, disp, 1.0, dsel)
        assert relerr(F.download(), ref) < TOL, axis
    F.destroy()


def test_advector_unsupported(sb):
    with pytest.raises(sb.SllbError) as e:
        sb.Advector1dPeriodic(64, 0.0, 1.0, sb.ADV_PERIODIC_SPLINE, 10)
    assert e.value.code == 2
    with pytest.raises(sb.SllbError):
        sb.Advector1dPeriodic(64, 0.0, 1.0, sb.ADV_PERIODIC_LAGRANGE, 5)
    with pytest.raises(sb.SllbError):
        sb.Interpolator1d(sb.INTERP_LAGRANGE_FIXED, 65, 0.0, 1.0, d_or_order=7)


# ---------------------------------------------------------------------------------------------
# a5/a8: interpolate_array_disp[_inplace]
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [27, 64, 100, 513])
def test_cubic_spline_interpolator(sb, orc, n):
    rng = np.random.default_rng(SEED + 7 * n)
    xmin, xmax = 0.0, 2 * np.pi
    dx = (xmax - xmin) / n
    itp = sb.Interpolator1d(sb.INTERP_CUBIC_SPLINE, n + 1, xmin, xmax)
    for alpha in [-1.2 * dx, 0.37 * dx, 0.0, 3 * dx, -17.81 * dx]:
        f = rng.standard_normal(n + 1); f[-1] = f[0]
        ref = orc.spline_interpolate_array_disp(f, xmin, xmax, alpha)
        out = itp.interpolate_array_disp(n + 1, f, alpha)
        assert relerr(out, ref) < TOL
        ref_ip = orc.spline_interpolate_array_disp_inplace(f, xmin, xmax, alpha)
        buf = f.copy()
        itp.interpolate_array_disp_inplace(n + 1, buf, alpha)
        assert relerr(buf, ref_ip) < TOL
    # reference KAT (test_cubic_spline_interpolator_1d.F90): analytic error < 1e-6 at n = 64
    if n == 64:
        x = xmin + dx * np.arange(n + 1)
        fa = lambda t: 2.0 * (np.sin(t) + 2.5 + np.cos(t))
        assert np.abs(itp.interpolate_array_disp(n + 1, fa(x), -1.2 * dx) - fa(x - 1.2 * dx)).max() < 1e-6
    itp.delete()


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("periodic_last", [0, 1])
def test_lagrange_fixed_interpolator(sb, orc, d, periodic_last):
    rng = np.random.default_rng(SEED + d)
    n, s = 100, 2 * d + 1
    npts = n + 1
    xmin, xmax = 0.0, float(n)
    itp = sb.Interpolator1d(sb.INTERP_LAGRANGE_FIXED, npts, xmin, xmax, d_or_order=d, periodic_last=periodic_last)
    for alpha in [0.2, -0.73, 0.999]:
        f = rng.standard_normal(npts); f[-1] = f[0]
        if periodic_last:
            ref = orc.lagrange("fixed_periodicl", f, alpha, s)
            out = itp.interpolate_array_disp(npts, f, alpha)
        else:
            ref = orc.lagrange("fixed_periodic", f[:-1], alpha, s)
            out = itp.interpolate_array_disp(n, f[:-1], alpha)
        assert relerr(out, ref) < TOL
    itp.delete()


@pytest.mark.parametrize("d", [2, 3, 4])
def test_lagrange_centered_interpolator(sb, orc, d):
    rng = np.random.default_rng(SEED + 31 * d)
    n = 100
    itp = sb.Interpolator1d(sb.INTERP_LAGRANGE_CENTERED, n + 1, 0.0, float(n), d_or_order=d)
    for alpha in [0.2, -1.7, 3.4, 2.0, 0.0]:
        f = rng.standard_normal(n + 1); f[-1] = f[0]
        ref = orc.lagrange_centered_barycentric(f, 0.0, float(n), d, 1, alpha)
        assert relerr(itp.interpolate_array_disp(n + 1, f, alpha), ref) < TOL
    itp.delete()


# ---------------------------------------------------------------------------------------------
# batched axis advection vs the oracle's per-line loops
# ---------------------------------------------------------------------------------------------
def _dsel_for(shape, axis, v_axis):
    """selector: displacement depends on index along v_axis only"""
    if v_axis > axis:
        stride = int(np.prod(shape[axis + 1:v_axis], dtype=np.int64))
        return (stride, shape[v_axis], 1, 1, 1, 0)
    stride = int(np.prod(shape[:v_axis], dtype=np.int64))
    return (1, 1, 0, stride, shape[v_axis], 1)


CASES_4D = [((32, 32, 16, 16), "tma-able"), ((20, 12, 16, 10), "ragged"), ((64, 8, 8, 40), "mixed")]


@pytest.mark.parametrize("shape,label", CASES_4D)
@pytest.mark.parametrize("method,order", [("spline", 4), ("lagrange_fixed", 3), ("lagrange_fixed", 7),
                                          ("lagrange_fixed", 11), ("lagrange_centered", 4), ("lagrange_centered", 8)])
@pytest.mark.parametrize("staging", [0, 2])
def test_advect_axis_4d(sb, orc, shape, label, method, order, staging):
    rng = np.random.default_rng(SEED + sum(shape) + order)
    mcode = {"spline": sb.METHOD_SPLINE, "lagrange_fixed": sb.METHOD_LAGRANGE_FIXED,
             "lagrange_centered": sb.METHOD_LAGRANGE_CENTERED}[method]
    sb.set_staging(staging)
    try:
        f0 = np.asfortranarray(rng.standard_normal(shape))
        F = sb.Field(shape)
        for axis in range(4):
            if shape[axis] < 8:
                continue
            v_axis = (axis + 2) % 4
            amp = 0.9 if method == "lagrange_fixed" else 2.7
            disp = rng.uniform(-amp, amp, shape[v_axis])
            dsel = _dsel_for(shape, axis, v_axis)
            ref = orc.advect_axis(f0.copy(order="F"), axis, method, order, disp, dsel)
            F.upload(f0)
            F.advect_axis(axis, mcode, order, disp, 1.0, dsel)
            out = F.download()
            assert relerr(out, ref) < TOL, (axis, label)
        # displacement from a "field" over the two fastest axes (v-advection, K5)
        for axis in (2, 3):
            if shape[axis] < 8:
                continue
            E = rng.uniform(-1, 1, shape[0] * shape[1])
            scale = 0.8
            dsel = (1, 1, 0, 1, shape[0] * shape[1], 1)
            ref = orc.advect_axis(f0.copy(order="F"), axis, method, order, E * scale, dsel)
            F.upload(f0)
            F.advect_axis(axis, mcode, order, E, scale, dsel)
            assert relerr(F.download(), ref) < TOL
        F.destroy()
    finally:
        sb.set_staging(0)


@pytest.mark.parametrize("split", [1, 2, 4, 8, -1])
@pytest.mark.parametrize("staging", [0, 2])
def test_spline_strided_split_kernel(sb, orc, split, staging):
    """the chunked (P warps per line) strided spline kernel, every split factor, TMA and cp.async staging"""
    rng = np.random.default_rng(SEED + 100 + split)
    sb.set_spline_split(split)
    sb.set_staging(staging)
    try:
        for shape in [(32, 64, 4, 8), (20, 128, 6), (64, 32, 128)]:
            f0 = np.asfortranarray(rng.standard_normal(shape))
            F = sb.Field(shape)
            for axis in range(1, len(shape)):
                if shape[axis] < 64:
                    continue
                nin = int(np.prod(shape[:axis]))
                disp = rng.uniform(-40, 40, nin)
                dsel = (1, 1, 0, 1, nin, 1)
                ref = orc.advect_axis(f0.copy(order="F"), axis, "spline", 4, disp, dsel)
                F.upload(f0)
                F.advect_axis(axis, sb.METHOD_SPLINE, 4, disp, 1.0, dsel)
                assert relerr(F.download(), ref) < TOL, (shape, axis, split)
            F.destroy()
    finally:
        sb.set_spline_split(-1)
        sb.set_staging(0)


@pytest.mark.parametrize("split", [1, 2, 4, 8, -1])
def test_spline_contig_split_kernel(sb, orc, split):
    """the chunked contiguous-axis spline kernel (one bulk TMA copy per 32-line tile, skewed chunk starts):
    every split factor, full and ragged tiles, large integer shifts, line lengths with and without the
    multiple-of-16 pitch, and the fallbacks (odd N, cp.async staging)."""
    rng = np.random.default_rng(SEED + 200 + split)
    sb.set_spline_split(split)
    try:
        for shape in [(128, 32, 4), (128, 37), (64, 33, 3), (256, 40), (100, 50), (72, 64), (65, 31), (512, 32), (32, 96)]:
            f0 = np.asfortranarray(rng.standard_normal(shape))
            F = sb.Field(shape)
            nl = int(np.prod(shape[1:]))
            disp = rng.uniform(-1.5 * shape[0], 1.5 * shape[0], nl)
            dsel = (1, nl, 1, 1, 1, 0)
            ref = orc.advect_axis(f0.copy(order="F"), 0, "spline", 4, disp, dsel)
            for staging in (0, 2):
                sb.set_staging(staging)
                F.upload(f0)
                F.advect_axis(0, sb.METHOD_SPLINE, 4, disp, 1.0, dsel)
                assert relerr(F.download(), ref) < TOL, (shape, split, staging)
            F.destroy()
    finally:
        sb.set_spline_split(-1)
        sb.set_staging(0)


@pytest.mark.parametrize("shape", [(128, 128, 5, 3), (64, 64, 7, 5), (32, 64, 4, 3), (128, 64, 3, 2), (64, 32, 150)])
@pytest.mark.parametrize("ept", [0, 16, 32])
def test_advect_plane_kernel(sb, orc, shape, ept):
    """K1c: x1 and x2 passes (+ the charge density) in one sweep == two oracle passes + a plain sum."""
    rng = np.random.default_rng(SEED + 300 + sum(shape))
    f0 = np.asfortranarray(rng.standard_normal(shape))
    nd = len(shape)
    n3 = shape[2]
    n4 = shape[3] if nd > 3 else 1
    v3 = rng.uniform(-70, 70, n3)
    v4 = rng.uniform(-70, 70, n4)
    d0 = (shape[1], n3, 1, 1, 1, 0)            # axis-0 lines: o = x2 + N2*(x3 + N3*x4) -> x3
    d1 = (n3, n4, 1, 1, 1, 0)                  # axis-1 lines: o = x3 + N3*x4 -> x4
    ref = orc.advect_axis(f0.copy(order="F"), 0, "spline", 4, v3 * 0.9, d0)
    ref = orc.advect_axis(ref, 1, "spline", 4, v4 * 1.1, d1)
    F = sb.Field(shape)
    sb.set_plane_kernel(True, ept)
    try:
        F.upload(f0)
        rho = F.advect_plane(v3, d0, 0.9, v4, d1, 1.1, rho_scale=0.25)
        out = F.download()
        assert relerr(out, ref) < TOL
        rho_ref = 0.25 * ref.reshape(shape[0], shape[1], -1).sum(axis=2)
        assert np.abs(rho - rho_ref).max() < 1e-12 * np.abs(ref).max() * (n3 * n4)
        F.upload(f0)
        assert F.advect_plane(v3, d0, 0.9, v4, d1, 1.1) is None
        assert relerr(F.download(), ref) < TOL
    finally:
        sb.set_plane_kernel(True, 0)
    F.destroy()


@pytest.mark.parametrize("shape", [(128, 128, 19, 17), (64, 64, 41, 23)])
@pytest.mark.parametrize("tmem,const_extents", [(-1, 1), (0, 1), (1, 1), (0, 0), (1, 0)])
def test_advect_plane_kernel_variants_many_planes(sb, orc, shape, tmem, const_extents):
    """K1c with more planes than CTAs (every CTA accumulates its charge density over several planes): the instantiations
    with compile-time extents and the run-time-extent kernel, accumulators in tensor memory and in registers + shared
    memory, all against two oracle passes + a plain sum."""
    rng = np.random.default_rng(SEED + 700 + sum(shape))
    f0 = np.asfortranarray(rng.standard_normal(shape))
    n3, n4 = shape[2], shape[3]
    v3 = rng.uniform(-70, 70, n3)
    v4 = rng.uniform(-70, 70, n4)
    d0 = (shape[1], n3, 1, 1, 1, 0)
    d1 = (n3, n4, 1, 1, 1, 0)
    ref = orc.advect_axis(f0.copy(order="F"), 0, "spline", 4, v3 * 0.9, d0)
    ref = orc.advect_axis(ref, 1, "spline", 4, v4 * 1.1, d1)
    rho_ref = 0.25 * ref.reshape(shape[0], shape[1], -1).sum(axis=2)
    F = sb.Field(shape)
    sb.set_plane_variant(tmem, const_extents)
    try:
        F.upload(f0)
        rho = F.advect_plane(v3, d0, 0.9, v4, d1, 1.1, rho_scale=0.25)
        assert relerr(F.download(), ref) < TOL
        assert np.abs(rho - rho_ref).max() < 1e-12 * np.abs(ref).max() * (n3 * n4)
        F.upload(f0)
        assert F.advect_plane(v3, d0, 0.9, v4, d1, 1.1) is None
        assert relerr(F.download(), ref) < TOL
    finally:
        sb.set_plane_variant(-1, 1)
    F.destroy()


def test_advect_plane_kernel_unsupported(sb):
    F = sb.Field((48, 40, 4))
    with pytest.raises(sb.SllbError) as ei:
        F.advect_plane(np.zeros(4), (40, 4, 1, 1, 1, 0), 1.0, np.zeros(1), (1, 1, 0, 1, 1, 0), 1.0)
    assert ei.value.code == 2
    F.destroy()
    F = sb.Field((64, 64, 4))
    with pytest.raises(sb.SllbError) as ei:   # displacement varying inside a plane
        F.advect_plane(np.zeros(64), (1, 64, 1, 1, 1, 0), 1.0, np.zeros(1), (1, 1, 0, 1, 1, 0), 1.0)
    assert ei.value.code == 2
    F.destroy()


def test_sim4d_fused_stage_kernels_match_separate_passes(sb):
    """the T-stage plane kernel + line-sum charge density vs. the separate passes + full reduction"""
    a4 = ([32, 64, 32, 32], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-2, 0.1)
    out = {}
    for on in (True, False):
        sb.set_plane_kernel(on)
        try:
            S = sb.Sim4d(*a4)
            rows = S.run(4)
            f = S.field().download()
            S.destroy()
        finally:
            sb.set_plane_kernel(True)
        out[on] = (rows, f)
    assert np.abs(out[True][0] / out[False][0] - 1).max() < 1e-10
    assert relerr(out[True][1], out[False][1]) < TOL


def test_advect_axis_6d(sb, orc):
    rng = np.random.default_rng(SEED)
    shape = (8, 10, 8, 8, 12, 8)
    f0 = np.asfortranarray(rng.standard_normal(shape))
    F = sb.Field(shape)
    for axis in range(6):
        v_axis = (axis + 3) % 6
        disp = rng.uniform(-0.9, 0.9, shape[v_axis])
        dsel = _dsel_for(shape, axis, v_axis)
        for method, mcode, order in (("lagrange_fixed", sb.METHOD_LAGRANGE_FIXED, 5), ("spline", sb.METHOD_SPLINE, 4)):
            ref = orc.advect_axis(f0.copy(order="F"), axis, method, order, disp, dsel)
            F.upload(f0)
            F.advect_axis(axis, mcode, order, disp, 1.0, dsel)
            assert relerr(F.download(), ref) < TOL, (axis, method)
    # affine helper == explicit array
    F.upload(f0)
    F.advect_axis_affine(1, sb.METHOD_LAGRANGE_FIXED, 7, 4, -6.0, 1.0, -0.07)
    a = F.download()
    v = (-6.0 + np.arange(shape[4])) * -0.07
    ref = orc.advect_axis(f0.copy(order="F"), 1, "lagrange_fixed", 7, v, _dsel_for(shape, 1, 4))
    assert relerr(a, ref) < TOL
    F.destroy()


def test_advect_long_lines_and_small(sb, orc):
    rng = np.random.default_rng(SEED + 5)
    for shape in [(1024, 24), (24, 1024), (8, 8), (2048, 3)]:
        f0 = np.asfortranarray(rng.standard_normal(shape))
        F = sb.Field(shape)
        for axis in (0, 1):
            if shape[axis] < 8:
                continue
            disp = rng.uniform(-3, 3, shape[1 - axis])
            dsel = _dsel_for(shape, axis, 1 - axis)
            for method, mcode, order in (("spline", sb.METHOD_SPLINE, 4), ("lagrange_fixed", sb.METHOD_LAGRANGE_FIXED, 7)):
                d = disp if method == "spline" else disp / 4
                ref = orc.advect_axis(f0.copy(order="F"), axis, method, order, d, dsel)
                F.upload(f0)
                F.advect_axis(axis, mcode, order, d, 1.0, dsel)
                assert relerr(F.download(), ref) < TOL, (shape, axis, method)
        F.destroy()
    F = sb.Field((4, 16))
    with pytest.raises(sb.SllbError) as e:       # line shorter than 8 points
        F.advect_axis(0, sb.METHOD_SPLINE, 4, np.zeros(16), 1.0, (1, 16, 1, 1, 1, 0))
    assert e.value.code == 2
    with pytest.raises(sb.SllbError):            # stencil not implemented
        F.advect_axis(1, sb.METHOD_LAGRANGE_FIXED, 13, np.zeros(4), 1.0, (1, 1, 0, 1, 4, 1))
    F.destroy()


# ---------------------------------------------------------------------------------------------
# field transfer, reductions, Poisson
# ---------------------------------------------------------------------------------------------
def test_field_upload_download_duplicates(sb):
    rng = np.random.default_rng(SEED)
    shape = (9, 8, 10, 8)
    core = rng.standard_normal(shape)
    F = sb.Field(shape)
    dup = [1, 1, 0, 1]
    big = np.pad(core, [(0, d) for d in dup], mode="wrap")
    F.upload(big, dup)
    assert np.array_equal(F.download(), core)
    assert np.array_equal(F.download(dup), big)
    assert np.array_equal(F.download([1, 1, 1, 1]), np.pad(core, [(0, 1)] * 4, mode="wrap"))
    F.destroy()


def test_reduce_and_moments(sb, orc):
    rng = np.random.default_rng(SEED)
    shape = (24, 20, 16, 12)
    f = rng.standard_normal(shape)
    F = sb.Field(shape).upload(f)
    # trapezoid over duplicated end points (sll_m_reduction.F90:187-272) == delta3*delta4 * plain sum
    fdup = np.pad(f, [(0, 0), (0, 0), (0, 1), (0, 1)], mode="wrap")
    ref = orc.reduction_34(fdup, 0.3, 0.7)
    out = F.reduce_velocity(2, 0.3 * 0.7)
    assert relerr(out, ref) < 1e-13
    w1 = np.concatenate([np.linspace(-6, 5, 16), np.linspace(-3, 3, 12)])
    m = F.moments(2, w1, w1 ** 2)
    s0 = f.sum(axis=(0, 1))
    assert abs(m[0] - f.sum()) < 1e-10 and abs(m[1] - np.abs(f).sum()) < 1e-10 and abs(m[2] - (f * f).sum()) < 1e-10
    assert abs(m[3] - (s0 * w1[:16, None]).sum()) < 1e-10 and abs(m[4] - (s0 * w1[None, 16:]).sum()) < 1e-10
    assert abs(m[5] - (s0 * w1[:16, None] ** 2).sum()) < 1e-10
    F.destroy()
    f6 = rng.standard_normal((8, 6, 4, 5, 4, 3))
    F6 = sb.Field(f6.shape).upload(f6)
    assert relerr(F6.reduce_velocity(3, -0.25), orc.charge_density_6d(f6, 0.25)) < 1e-13
    F6.destroy()


def test_poisson_1d(sb, orc):
    nc, m = 128, 4
    x = np.arange(nc + 1) * 2 * np.pi / nc
    P = sb.Poisson([nc], [0.0], [2 * np.pi])
    (_, E) = P.solve(m * m * np.sin(m * x))
    assert np.abs(E + m * np.cos(m * x)).max() <= 1e-13      # test_poisson_1d_periodic.F90
    rng = np.random.default_rng(SEED)
    rho = rng.standard_normal(nc + 1); rho[-1] = rho[0]
    assert relerr(P.solve(rho)[1], orc.poisson_1d(rho, 0.0, 2 * np.pi)) < 1e-13
    P.destroy()


@pytest.mark.parametrize("n1,n2", [(128, 128), (16, 12), (64, 32)])
def test_poisson_2d(sb, orc, n1, n2):
    rng = np.random.default_rng(SEED + n1)
    P = sb.Poisson([n1, n2], [0.0, 0.0], [3.0, 5.0])
    rho = rng.standard_normal((n1 + 1, n2 + 1)); rho[-1, :] = rho[0, :]; rho[:, -1] = rho[:, 0]
    phi, e1, e2 = P.solve(rho)
    rex, rey, rphi = orc.poisson_2d(rho, n1, n2, 0.0, 3.0, 0.0, 5.0, want_phi=True)
    # random data has O(1) Nyquist content: exercises the FFTW c2r semantics
    assert relerr(e1, rex) < 1e-13 and relerr(e2, rey) < 1e-13 and relerr(phi, rphi) < 1e-13
    P.destroy()
    if n1 == 128:   # test_poisson_2d_periodic.F90 known answer
        mode = 2
        x = np.arange(n1 + 1) * 2 * np.pi / n1
        X1, X2 = np.meshgrid(x, x, indexing="ij")
        P = sb.Poisson([n1, n2], [0.0, 0.0], [2 * np.pi, 2 * np.pi])
        phi, e1, e2 = P.solve(-2.0 * mode ** 3 * np.sin(mode * X1) * np.cos(mode * X2))
        assert np.abs(mode * np.sin(mode * X1) * np.cos(mode * X2) + phi).max() <= 1e-13
        assert np.abs(mode ** 2 * np.cos(mode * X1) * np.cos(mode * X2) - e1).max() <= 1e-13
        assert np.abs(-mode ** 2 * np.sin(mode * X1) * np.sin(mode * X2) - e2).max() <= 1e-13
        P.destroy()


def test_poisson_3d(sb, orc):
    rng = np.random.default_rng(SEED)
    n = (16, 12, 8)
    L = (4 * np.pi, 3.0, 7.0)
    rho = rng.standard_normal(n)
    P = sb.Poisson(n, [0, 0, 0], L)
    phi, ex, ey, ez = P.solve(rho)
    r = orc.poisson_3d(rho, *L)
    for a, b in zip((phi, ex, ey, ez), r):
        assert relerr(a, b) < 1e-13
    P.destroy()


# ---------------------------------------------------------------------------------------------
# simulations: trace parity with the oracle's restatement of the reference time loops
# ---------------------------------------------------------------------------------------------
def test_sim6d_golden_and_oracle(sb, orc):
    """G1: reffile_bsl_vp_3d3v_cart_dd.dat, reference tolerance 5e-7 (sll_m_sim_6d_utilities.F90:663)."""
    gold = np.loadtxt(os.path.join(os.path.dirname(__file__), "golden", "reffile_bsl_vp_3d3v_cart_dd.dat"))
    S = sb.Sim6d([16] * 6, 6.0, [12.5663706144] * 3, 3, 3, 0.01, 0.01, [0.499999999998376] * 3)
    rows = S.run(2)
    assert rows.shape == (3, 14)
    assert np.abs(rows - gold).max() < 5e-7
    orows, of = orc.sim6d([16] * 6, 6.0, [12.5663706144] * 3, 3, 3, 0.01, 2, 0.01, [0.499999999998376] * 3, want_f=True)
    assert np.abs(rows - orows).max() < 1e-12
    f = S.field().download()
    assert relerr(f, of) < TOL
    S.destroy()


def test_sim6d_lagrange7(sb, orc):
    n = [12, 10, 8, 12, 10, 8]
    S = sb.Sim6d(n, 6.0, [4 * np.pi] * 3, 7, 7, 0.01, 0.01, [0.5] * 3)
    rows = S.run(2)
    orows, of = orc.sim6d(n, 6.0, [4 * np.pi] * 3, 7, 7, 0.01, 2, 0.01, [0.5] * 3, want_f=True)
    assert np.abs(rows - orows).max() < 1e-12
    assert relerr(S.field().download(), of) < TOL
    S.destroy()


def test_sim2d_c1_trace(sb, orc):
    """C1: 1D1V Landau 128x128, spline order 4, Strang VTV, dt = 0.1 (sim_bsl_vp_1d1v_cart)."""
    args = (128, 128, 0.0, 4 * np.pi, -6.0, 6.0, 0, 0.5, 1e-3, 0.1)
    nsteps = 100
    S = sb.Sim2d(*args)
    rows = S.run(nsteps)
    orows, of, oE = orc.sim2d(*args, nsteps, method=0, want_f=True)
    f = S.field().download([1, 1])
    # per-step parity compounds over 100 steps x 3 passes; field values agree far below 1e-12
    assert relerr(f[:, :-1], of[:, :-1]) < 1e-12
    for col, name in [(1, "mass"), (2, "l1"), (4, "l2"), (5, "ekin"), (7, "etot")]:
        assert np.abs(rows[:, col] / orows[:, col] - 1).max() < 1e-10, name
    # field energy spans orders of magnitude; E itself agrees to ~1e-12 absolute (rounding + the end-point term of
    # the trapezoid rule, DESIGN.md "periodic cells only"), i.e. 1e-9 of the trace maximum and 1e-6 per sample
    assert np.abs(rows[:, 6] - orows[:, 6]).max() / orows[:, 6].max() < 1e-9
    assert np.abs(rows[:, 6] / orows[:, 6] - 1).max() < 1e-6
    assert np.abs(rows[:, 3] - orows[:, 3]).max() < 1e-9      # momentum ~ 0
    S.destroy()


def test_sim2d_c2_lagrange7_small(sb, orc):
    """C2 at reduced size: two-stream, Lagrange fixed 7-pt both axes, max|v| dt/dx <= 1."""
    nx = 256
    dx = 4 * np.pi / nx
    dt = 0.9 * dx / 6.0
    args = (nx, nx, 0.0, 4 * np.pi, -6.0, 6.0, 1, 0.5, 0.01, dt)
    S = sb.Sim2d(*args, method=sb.METHOD_LAGRANGE_FIXED, order=7)
    rows = S.run(20)
    orows, of, _ = orc.sim2d(*args, 20, method=3, order=7, want_f=True)
    assert relerr(S.field().download([1, 1])[:, :-1], of[:, :-1]) < 1e-12
    assert np.abs(rows[:, 1] / orows[:, 1] - 1).max() < 1e-10
    assert np.abs(rows[:, 6] - orows[:, 6]).max() / orows[:, 6].max() < 1e-9
    S.destroy()


@pytest.mark.parametrize("split", [0, 1])
def test_sim4d_trace(sb, orc, split):
    """C3 semantics at 16^2 x 32^2 (the shipped vpsim4d nml sizes): spline order 4, Strang."""
    nc = [16, 16, 32, 32]
    xmin, xmax = [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6]
    S = sb.Sim4d(nc, xmin, xmax, 0.5, 0.5, 1e-3, 0.1, split=split)
    row0 = S.diagnostics()
    rows = S.run(10)
    orows, of = orc.sim4d(nc, xmin, xmax, 0.5, 0.5, 1e-3, 0.1, 10, split=split, method=0, want_f=True)
    assert abs(row0[1] / orows[0, 1] - 1) < 1e-10
    f = S.field().download([1, 1, 1, 1])
    # the oracle carries the duplicated velocity end planes like the reference (they are advected in x with
    # +vmax instead of -vmax and re-synchronised by every V stage); compare the periodic cells
    assert relerr(f[:-1, :-1, :-1, :-1], of[:-1, :-1, :-1, :-1]) < 1e-10
    assert np.abs(rows[:, 1] / orows[1:, 1] - 1).max() < 1e-6     # field energy (end-plane term)
    assert np.abs(rows[:, 3] / orows[1:, 3] - 1).max() < 1e-8     # mass
    assert np.abs(rows[:, 2] / orows[1:, 2] - 1).max() < 1e-8     # kinetic energy
    S.destroy()


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE configs), size independent
# ---------------------------------------------------------------------------------------------
def test_full_size_properties_2d2v_64(sb):
    n = 64
    S = sb.Sim4d([n] * 4, [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1)
    F = S.field()
    f0 = F.download()
    mass0 = f0.sum()
    # (1) integer-cell shifts are exact permutations for the spline (dx = 0 weights are 1/6, 4/6, 1/6 of an
    #     interpolating spline): compare against np.roll to rounding
    F.advect_axis(0, sb.METHOD_SPLINE, 4, np.full(1, 3.0), 1.0)
    assert relerr(F.download(), np.roll(f0, -3, axis=0)) < 1e-13
    # (2) advecting by +d then -d returns to the start up to the interpolation error, and conserves mass
    F.upload(f0)
    for axis in range(4):
        F.advect_axis(axis, sb.METHOD_SPLINE, 4, np.full(1, 0.37), 1.0)
    f1 = F.download()
    assert abs(f1.sum() / mass0 - 1) < 1e-12
    # (3) linearity: A(f + 2g) = A(f) + 2 A(g)
    rng = np.random.default_rng(SEED)
    g0 = np.asfortranarray(rng.standard_normal(f0.shape))
    disp = rng.uniform(-2, 2, n)
    outs = []
    for data in (f0, g0, f0 + 2 * g0):
        F.upload(data)
        F.advect_axis(2, sb.METHOD_SPLINE, 4, disp, 1.0, (1, 1, 0, 1, n, 1))
        outs.append(F.download())
    assert relerr(outs[2], outs[0] + 2 * outs[1]) < 1e-13
    # (4) constants are preserved exactly-ish by every method
    F.upload(np.ones_like(f0))
    for axis, (m, o) in enumerate([(sb.METHOD_SPLINE, 4), (sb.METHOD_LAGRANGE_FIXED, 7), (sb.METHOD_LAGRANGE_CENTERED, 6),
                                   (sb.METHOD_SPLINE, 4)]):
        F.advect_axis(axis, m, o, disp, 0.3, (1, 1, 0, 1, 1, 0) if axis else (1, n, 1, 1, 1, 0))
    assert np.abs(F.download() - 1).max() < 1e-14
    S.destroy()


# ---------------------------------------------------------------------------------------------
# committed golden vectors (tests/golden/oracle_*.npz)
# ---------------------------------------------------------------------------------------------
def test_committed_golden_vectors(sb):
    """the CUDA path against COMMITTED oracle outputs (no oracle call on this box)"""
    G = os.path.join(os.path.dirname(__file__), "golden")
    d = np.load(os.path.join(G, "oracle_advect4d.npz"))
    f0 = np.asfortranarray(d["f0"])
    codes = {"spline": sb.METHOD_SPLINE, "lagrange_fixed": sb.METHOD_LAGRANGE_FIXED, "lagrange_centered": sb.METHOD_LAGRANGE_CENTERED}
    F = sb.Field(f0.shape)
    n = 0
    for key in d.files:
        if "_axis" not in key:
            continue
        name, axis = key.rsplit("_axis", 1)
        axis = int(axis)
        method = name.rstrip("0123456789")
        order = int(name[len(method):])
        F.upload(f0)
        F.advect_axis(axis, codes[method], order, d[f"disp{axis}"], 1.0, tuple(int(v) for v in d[f"dsel{axis}"]))
        assert relerr(F.download(), d[key]) < TOL, key
        n += 1
    assert n == 8
    F.destroy()
    ln = np.load(os.path.join(G, "oracle_lines.npz"))
    adv = sb.Advector1dPeriodic(64, 0.0, 2 * np.pi, sb.ADV_PERIODIC_SPLINE, 4)
    assert relerr(adv.advect_1d_constant(1.3, 0.1, ln["line"]), ln["adv_spline"]) < TOL
    adv.delete()
    adv = sb.Advector1dPeriodic(64, 0.0, 2 * np.pi, sb.ADV_PERIODIC_LAGRANGE, 6)
    assert relerr(adv.advect_1d_constant(1.3, 0.1, ln["line"]), ln["adv_lagrange6"]) < TOL
    adv.delete()
    itp = sb.Interpolator1d(sb.INTERP_CUBIC_SPLINE, 65, 0.0, 2 * np.pi)
    assert relerr(itp.interpolate_array_disp(65, ln["line"], -1.2 * (2 * np.pi / 64)), ln["spline_disp"]) < TOL
    itp.delete()
    po = np.load(os.path.join(G, "oracle_poisson.npz"))
    P = sb.Poisson([16, 12], [0.0, 0.0], [4 * np.pi, 2 * np.pi])
    _, e1, e2 = P.solve(po["rho2"])
    assert np.abs(e1 - po["e1"]).max() < 1e-12 * np.abs(po["e1"]).max()
    assert np.abs(e2 - po["e2"]).max() < 1e-12 * np.abs(po["e2"]).max()
    P.destroy()
    tr = np.load(os.path.join(G, "oracle_traces.npz"))
    S = sb.Sim4d([16, 16, 32, 32], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1)
    rows = S.run(5)
    S.destroy()
    assert np.abs(rows / tr["rows4"][1:] - 1).max() < 1e-8
    S2 = sb.Sim2d(64, 64, 0.0, 4 * np.pi, -6.0, 6.0, 0, 0.5, 1e-3, 0.1)
    r2 = S2.run(20)
    S2.destroy()
    sel = [0, 1, 2, 4, 5, 7]      # momentum (~1e-17) and the decaying potential energy are compared absolutely
    assert np.abs(r2[:, sel] / tr["rows2"][:, sel] - 1).max() < 1e-8
    assert np.abs(r2[:, 3] - tr["rows2"][:, 3]).max() < 1e-12
    assert np.abs(r2[:, 6] - tr["rows2"][:, 6]).max() < 1e-9 * tr["rows2"][:, 6].max()
