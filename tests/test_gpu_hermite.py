"""GPU parity tests of K10, cubic-spline interpolation with Hermite boundary conditions (SURVEY.md section 8(f) rank 3:
sll_m_cubic_splines.F90 sll_p_hermite through sll_t_cubic_spline_interpolator_1d), through the C ABI against the oracle
(oracle/sll_oracle_hermite.c).  Tolerance 1e-12 * max|f| per pass."""
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


def test_kat_reference_unit_test(sb):
    """test_cubic_splines.F90:57-138 (bc = 1): f = exp(sin x), 33 points, exact end slopes: the spline reproduces the
    grid values (1e-14) and the mid-cell value to 2e-5 -- through interpolate_array_disp with alpha = 0 and -delta/2"""
    npts = 33
    delta = 2 * np.pi / 32
    x = np.arange(npts) * delta
    data = np.exp(np.sin(x)); deriv = np.cos(x) * np.exp(np.sin(x))
    I = sb.Interpolator1d(sb.INTERP_CUBIC_SPLINE, npts, 0.0, 2 * np.pi, bc=sb.BC_HERMITE)
    I.set_slopes(deriv[0], deriv[-1])
    out = I.interpolate_array_disp(npts, data, 0.0)
    assert np.abs(out[:32] - data[:32]).max() <= 1e-14
    out = I.interpolate_array_disp(npts, data, -0.5 * delta)
    xg = (16 - 0.5) * delta
    assert abs(out[16] - np.exp(np.sin(xg))) <= 2e-5
    I.delete()


@pytest.mark.parametrize("npts", [27, 33, 129])
def test_line_objects_vs_oracle(sb, orc, npts):
    """sll_t_cubic_spline_interpolator_1d with sll_p_hermite: interpolate_array_disp and _inplace, finite-difference slopes"""
    rng = np.random.default_rng(SEED + npts)
    xmin, xmax = -6.0, 6.0
    delta = (xmax - xmin) / (npts - 1)
    I = sb.Interpolator1d(sb.INTERP_CUBIC_SPLINE, npts, xmin, xmax, bc=sb.BC_HERMITE)
    for ac in (0.0, 0.41, -0.41, 1.0, -2.0, 3.3, -5.7):
        data = rng.standard_normal(npts)
        ref = orc.hermite_interpolate_array_disp(data, xmin, xmax, ac * delta, inplace=False)
        assert relerr(I.interpolate_array_disp(npts, data, ac * delta), ref) <= TOL
        ref = orc.hermite_interpolate_array_disp(data, xmin, xmax, ac * delta, inplace=True)
        d2 = data.copy()
        I.interpolate_array_disp_inplace(npts, d2, ac * delta)
        assert relerr(d2, ref) <= TOL
    I.delete()
    with pytest.raises(sb.SllbError) as e:
        sb.Interpolator1d(sb.INTERP_CUBIC_SPLINE, 20, xmin, xmax, bc=sb.BC_HERMITE)
    assert e.value.code == sb.ERR_UNSUPPORTED


@pytest.mark.parametrize("shape", [(32, 4, 33, 40), (12, 5, 65, 27), (64, 32, 33, 33)])
@pytest.mark.parametrize("staging", [0, 2])
@pytest.mark.parametrize("inplace", [True, False])
def test_batched_v_passes_vs_oracle(sb, orc, shape, staging, inplace):
    """the velocity passes of bsl_vp_2d2v_cart (:520-545): alpha = E(x1,x2) dt per line, axes 2 and 3 (strided), and the
    contiguous axis 0 for the kernel's other tile shape"""
    rng = np.random.default_rng(SEED)
    f0 = np.asfortranarray(rng.standard_normal(shape))
    xmin, xmax = -6.0, 6.0
    F = sb.Field(shape)
    sb.set_staging(staging)
    try:
        n12 = shape[0] * shape[1]
        for axis in (2, 3):
            delta = (xmax - xmin) / (shape[axis] - 1)
            E = rng.uniform(-2.5, 2.5, n12) * delta
            dsel = (1, 1, 0, 1, n12, 1)
            ref = orc.hermite_advect_axis(f0.copy(order="F"), axis, xmin, xmax, E, dsel, inplace=inplace)
            F.upload(f0)
            F.advect_axis_hermite(axis, xmin, xmax, E, dsel, inplace=inplace)
            assert relerr(F.download(), ref) <= TOL, axis
        if shape[0] >= 27:
            delta = (xmax - xmin) / (shape[0] - 1)
            v = rng.uniform(-2.5, 2.5, shape[2]) * delta
            dsel = (shape[1], shape[2], 1, 1, 1, 0)
            ref = orc.hermite_advect_axis(f0.copy(order="F"), 0, xmin, xmax, v, dsel, inplace=inplace)
            F.upload(f0)
            F.advect_axis_hermite(0, xmin, xmax, v, dsel, inplace=inplace)
            assert relerr(F.download(), ref) <= TOL
    finally:
        sb.set_staging(0)
    F.destroy()


def test_properties(sb):
    """zero displacement is the identity to rounding; a linear function is reproduced exactly inside the domain (cubic
    splines with exact finite-difference slopes) and clamped outside"""
    shape = (32, 8, 65)
    xmin, xmax = -6.0, 6.0
    delta = (xmax - xmin) / 64
    v = xmin + delta * np.arange(65)
    f0 = np.asfortranarray(np.broadcast_to(2.0 * v + 1.0, shape).copy())
    F = sb.Field(shape)
    F.upload(f0)
    F.advect_axis_hermite(2, xmin, xmax, np.zeros(1), (1, 1, 0, 1, 1, 0))
    assert np.abs(F.download() - f0).max() < 1e-13
    F.upload(f0)
    F.advect_axis_hermite(2, xmin, xmax, np.array([0.3 * delta]), (1, 1, 0, 1, 1, 0))
    got = F.download()
    assert np.abs(got[:, :, :-1] - (2.0 * (v[:-1] + 0.3 * delta) + 1.0)).max() < 1e-12
    assert np.abs(got[:, :, -1] - (2.0 * xmax + 1.0)).max() < 1e-12     # foot beyond xmax: boundary value
    F.destroy()
