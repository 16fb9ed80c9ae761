"""Pins the oracle's restatement of the Hermite-BC cubic spline (oracle/sll_oracle_hermite.c) with the reference's
known-answer test and checks the two algorithms and the two displacement entry points against each other."""
import numpy as np
import pytest

from oracle import orc

RNG = np.random.default_rng(20261017)


def test_kat_cubic_spline_1d_hermite():
    """test_cubic_splines.F90:57-138 with bc = 1: np = 32, f = exp(sin x) on [0, 2 pi], exact end slopes;
    grid values reproduced to 1e-14, value at the mid-cell point (np/2 - 1/2) delta to 2e-5."""
    npts = 33
    x = np.arange(npts) * 2 * np.pi / 32
    data = np.exp(np.sin(x)); deriv = np.cos(x) * np.exp(np.sin(x))
    c = orc.hermite_coeffs(data, 0.0, 2 * np.pi, slopes=(deriv[0], deriv[-1]))
    vals = orc.spline_eval_array(c, npts, 0.0, 2 * np.pi, x[:32])
    assert np.abs(vals - data[:32]).max() <= 1e-14
    xg = (32 / 2 - 0.5) * 2 * np.pi / 32
    assert abs(orc.spline_eval_array(c, npts, 0.0, 2 * np.pi, xg)[0] - np.exp(np.sin(xg))) <= 2e-5


@pytest.mark.parametrize("npts", [27, 33, 64, 129])
def test_fast_algorithm_is_the_hermite_spline(npts):
    """compute_spline_1D_hermite_aux (27-term start, :583-652): the coefficients satisfy the interpolation conditions
    (c_{i-1} + 4 c_i + c_{i+1})/6 = f_i at every point and the two end-slope conditions (c_2 - c_0)/(2 delta) = slope_l,
    (c_{np+1} - c_{np-1})/(2 delta) = slope_r with the 5-point finite-difference slopes (:176-181).
    (The matrix of the reference's non-fast path, :340-352, pairs the first row with c_0, c_1 instead of c_1, c_2 and
    does not produce this spline; it is only reachable with fewer than 27 points or fast_algorithm = .false. and is not
    offered by the GPU path.)"""
    xmin, xmax = -6.0, 6.0
    delta = (xmax - xmin) / (npts - 1)
    data = RNG.standard_normal(npts)
    c = orc.hermite_coeffs(data, xmin, xmax, fast=1)
    scale = np.abs(c).max()
    assert np.abs((c[0:npts] + 4 * c[1:npts + 1] + c[2:npts + 2]) / 6 - data).max() < 1e-13 * scale
    sl = (-(25 / 12) * data[0] + 4 * data[1] - 3 * data[2] + (4 / 3) * data[3] - 0.25 * data[4]) / delta
    sr = (0.25 * data[-5] - (4 / 3) * data[-4] + 3 * data[-3] - 4 * data[-2] + (25 / 12) * data[-1]) / delta
    assert abs((c[2] - c[0]) / (2 * delta) - sl) < 1e-12 * scale / delta
    assert abs((c[npts + 1] - c[npts - 1]) / (2 * delta) - sr) < 1e-12 * scale / delta


def test_disp_entry_points():
    """interpolate_array_disp_inplace clamps the feet to [xmin, xmax]; interpolate_array_disp agrees with it except at the
    one point whose foot lies in the ghost cell beyond xmax (eval_disp evaluates the ghost cell there, :2640-2643)"""
    npts, xmin, xmax = 65, -6.0, 6.0
    delta = (xmax - xmin) / (npts - 1)
    data = np.exp(-0.5 * np.linspace(xmin, xmax, npts) ** 2) + 0.1 * RNG.standard_normal(npts)
    for alpha in (0.0, 0.3 * delta, -0.3 * delta, 2.6 * delta, -3.2 * delta):
        a = orc.hermite_interpolate_array_disp(data, xmin, xmax, alpha, inplace=True)
        b = orc.hermite_interpolate_array_disp(data, xmin, xmax, alpha, inplace=False)
        dcell = int(np.floor(alpha / delta))
        same = np.ones(npts, bool)
        if dcell >= 1 or (dcell == 0 and alpha > 0):
            same[npts - 1 - dcell] = False if dcell >= 1 else True
        assert np.abs(a[same] - b[same]).max() < 1e-13
        if alpha == 0.0:
            assert np.abs(a - data).max() < 1e-14
        # feet beyond the boundary take the boundary value
        if alpha < 0:
            assert np.abs(a[:-dcell - 1] - a[0]).max() == 0 if dcell < -1 else True


@pytest.mark.parametrize("npts", [27, 33, 65, 130])
@pytest.mark.parametrize("inplace", [True, False])
def test_device_line_functions_match_oracle(npts, inplace):
    """hermite_coeffs_line / hermite_eval_point (what the K10 kernels run per line, sllb_hermite.cuh compiled for the
    host) vs the restated reference: 1e-12 max|f|"""
    from host import emu
    xmin, xmax = -6.0, 6.0
    delta = (xmax - xmin) / (npts - 1)
    for alpha_cells in (0.0, 0.37, -0.37, 1.0, -1.0, 2.25, -3.6, 0.999, 40.0, -40.0):
        if abs(alpha_cells) > npts and not inplace:
            continue                    # eval_disp's loops run past the array there
        data = RNG.standard_normal(npts)
        alpha = alpha_cells * delta
        got = emu.hermite_line(data, delta, alpha, inplace=inplace)
        ref = orc.hermite_interpolate_array_disp(data, xmin, xmax, alpha, inplace=inplace)
        scale = max(np.abs(ref).max(), np.abs(data).max())
        assert np.abs(got - ref).max() <= 1e-12 * scale, (alpha_cells, np.abs(got - ref).argmax())
    data = RNG.standard_normal(npts)
    got = emu.hermite_line(data, delta, 0.3 * delta, inplace=inplace, slopes=(0.7, -1.3))
    ref = orc.hermite_interpolate_array_disp(data, xmin, xmax, 0.3 * delta, inplace=inplace, slopes=(0.7, -1.3))
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
