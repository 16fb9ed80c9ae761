"""Pins the oracle's restatement of the local cubic spline with halo cells (oracle/sll_oracle_halo.c) with the
reference's own known-answer test and checks the emulated decomposition."""
import numpy as np
import pytest

from oracle import orc

RNG = np.random.default_rng(20261017)


def _kat_data(n=64):
    x = np.arange(n + 1) * 2 * np.pi / n
    return 2.0 * (np.sin(x) + 2.5 + np.cos(x)), 2 * np.pi / n


def _halo_version(pdata, n, si, alpha):
    """test_halo_version of test_cubic_spline_halo_1d.F90:110-143"""
    d0, c2 = orc.halo_prepare_exchange(pdata[:n], si)
    d0, c2 = orc.halo_finish_boundary_conditions(pdata[:n], si, d0, c2)
    fin = np.zeros(n + 3)
    fin[0] = d0; fin[n + 2] = c2
    for i in range(1, n + 2):
        fin[i] = pdata[(si + i - 1) % n]
    coeffs = orc.halo_compute_interpolant(fin, n)
    return orc.halo_eval_disp(coeffs, alpha, n), c2


@pytest.mark.parametrize("si", [-2, -1, 0, 1, 2])
def test_kat_halo_vs_global_spline(si):
    """test_cubic_spline_halo_1d.F90:28-82: n = 64, f = 2(sin x + 2.5 + cos x), alpha = 0.25, tolerance 4e-9
    (the value the reference chose for NUM_TERMS = 15) against sll_t_cubic_spline_interpolator_1d."""
    n, alpha = 64, 0.25
    pdata, delta = _kat_data(n)
    out, c2 = _halo_version(pdata, n, si, alpha)
    ref = orc.spline_interpolate_array_disp(pdata, 0.0, 2 * np.pi, (si + alpha) * delta)
    err = np.abs(out - ref[:n]).max()
    assert err <= 4e-9
    assert err > 1e-13          # the 15-term truncation is really there (a 27-term version would be ~1e-15)
    # the test's own 27-term closed form of c_np2 (:122-127)
    i = np.arange(1, 28)
    c27 = pdata[(1 + si) % n] + np.sum((np.sqrt(3.0) - 2.0) ** i * (pdata[(n + 1 - i + si) % n] + pdata[(1 + i + si) % n]))
    assert abs(c2 - np.sqrt(3.0) * c27) < 4e-8


def test_kat_halo_periodic():
    """test_cubic_spline_halo_1d.F90:84-90: sll_s_cubic_spline_halo_1d_periodic vs the global spline, 4e-9"""
    n, alpha = 64, 0.25
    pdata, delta = _kat_data(n)
    out = orc.halo_periodic(pdata[:n].copy(), alpha)
    ref = orc.spline_interpolate_array_disp(pdata, 0.0, 2 * np.pi, alpha * delta)
    assert np.abs(out - ref[:n]).max() <= 4e-9


def test_constant_preserved_to_truncation():
    out, _ = _halo_version(np.full(65, 3.0), 64, 0, 0.4)
    assert np.abs(out - 3.0).max() < 3 * 4e-9


def test_make_blocks_spline_matches_floor_and_skips_zero():
    """make_blocks_spline (sll_m_advection_6d_spline_dd_slim.F90:202-287) on the simulation's displacement
    -v dt/dx: every index gets floor(disp) except the v = 0 line, which belongs to no block."""
    n = 32
    v = -6.0 + 12.0 / n * np.arange(n)
    for dt_dx in (0.05, 0.2, 0.45):
        disp = -v * dt_dx
        shift, alpha, nb = orc.make_blocks_spline(disp)
        zero = np.where(disp == 0.0)[0]
        assert zero.size == 1 and shift[zero[0]] == orc.SKIP
        keep = disp != 0.0
        notint = keep & (disp != np.floor(disp))
        assert np.array_equal(shift[notint], np.floor(disp[notint]).astype(np.int32))
        assert nb == int(np.floor(disp[0])) - int(np.floor(disp[-1])) + 1
        assert np.allclose(alpha, disp - np.floor(disp), rtol=0, atol=0)
    # increasing displacement
    shift, alpha, nb = orc.make_blocks_spline(v * 0.2)
    keep = (v != 0.0)
    assert np.array_equal(shift[keep], np.floor(v[keep] * 0.2).astype(np.int32))


def test_decomposed_line_matches_global_to_truncation():
    """nblk emulated ranks vs one rank: same spline up to the 15-term truncation (<= 4e-9 of max|f|), and the
    periodic variant (a different truncated start of the backward sweep) agrees to the same level."""
    n = 72
    f = np.asfortranarray(RNG.standard_normal((4, n, 3)))
    disp = RNG.uniform(-1.0, 1.0, 4)
    dsel = (1, 1, 0, 1, 4, 1)
    one = orc.spline_dd_advect_axis(f.copy(order="F"), 1, 1, disp, dsel)
    for nblk in (2, 3, 4):
        dec = orc.spline_dd_advect_axis(f.copy(order="F"), 1, nblk, disp, dsel)
        assert 1e-14 < np.abs(dec - one).max() < 4e-8
    exact = orc.advect_axis(f.copy(order="F"), 1, "spline", 4, disp, dsel)
    assert np.abs(one - exact).max() < 4e-8
    for i0 in range(4):
        if disp[i0] >= 0:
            ref = orc.halo_periodic(f[i0, :, 0].copy(), disp[i0])
            assert np.abs(one[i0, :, 0] - ref).max() < 4e-8


def test_too_few_points_rejected():
    f = np.asfortranarray(RNG.standard_normal((2, 30, 2)))
    with pytest.raises(ValueError):
        orc.spline_dd_advect_axis(f, 1, 2, np.zeros(2) + 0.3, (1, 1, 0, 1, 2, 1))
    f = np.asfortranarray(RNG.standard_normal((2, 18, 2)))
    orc.spline_dd_advect_axis(f, 1, 1, np.zeros(2) - 1.7, (1, 1, 0, 1, 2, 1))       # si = -2: 18 >= 16 + 2
    with pytest.raises(ValueError):
        orc.spline_dd_advect_axis(f, 1, 1, np.zeros(2) - 2.7, (1, 1, 0, 1, 2, 1))   # si = -3: out of bounds in the reference


# ---- the CUDA path's per-line functions (sllb_spline15.cuh) compiled for the host, against the oracle -------------
def _oracle_line(line, nblk, si, alpha):
    f = np.asfortranarray(line.reshape(1, -1, 1).copy())
    orc.spline_dd_advect_axis(f, 1, nblk, np.array([si + alpha]), (1, 1, 0, 1, 1, 0), shifts=np.array([si]))
    return f[0, :, 0]


@pytest.mark.parametrize("nblk,si,mode", [(1, 0, 0), (1, -1, 0), (1, 3, 0), (1, -4, 0), (1, 0, 1), (1, -1, 1), (1, 2, 1),
                                          (2, 0, 0), (2, -1, 0), (3, 0, 0), (3, -1, 0), (2, 1, 0), (2, -2, 0), (4, 2, 0)])
def test_device_line_functions_match_oracle(nblk, si, mode):
    """spline15_sums / spline15_line / spline15_prepare (what the kernels run per line) vs the reference sequence
    prepare_exchange -> finish_boundary_conditions -> compute_interpolant -> eval_disp: 1e-12 max|f| (observed ~1e-15)"""
    from host import emu
    n = 40 * nblk
    for alpha in (0.0, 0.3, 0.75, 0.999):
        line = RNG.standard_normal(n)
        hw = max(1, abs(si) + 1)
        got = emu.spline_dd_line(line, nblk, si, alpha, hwl=hw, hwr=hw, mode=mode)
        ref = _oracle_line(line, nblk, si, alpha)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(line).max()


def test_device_line_skip():
    from host import emu
    line = RNG.standard_normal(48)
    assert np.array_equal(emu.spline_dd_line(line, 1, orc.SKIP, 0.0), line)
