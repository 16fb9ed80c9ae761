"""GPU parity tests of K9, the LOCAL cubic spline with halo cells (SURVEY.md section 8(f) rank 1:
sll_m_cubic_spline_halo_1d + sll_t_advection_6d_spline_dd_slim), and of the centred variable-block Lagrange
x-advection (rank 2), through the C ABI against the oracle (oracle/sll_oracle_halo.c).
Tolerance: 1e-12 * max|f| per advection pass (the truncated 15-term series is restated exactly, so the 2.6e-9
truncation itself cancels)."""
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


def _dsel(shape, axis, v_axis):
    if v_axis > axis:
        stride = int(np.prod(shape[axis + 1:v_axis], dtype=np.int64))
        return (stride, shape[v_axis], 1, 1, 1, 0)
    return (1, 1, 0, int(np.prod(shape[:v_axis], dtype=np.int64)), shape[v_axis], 1)


def test_kat_reference_unit_test(sb, orc):
    """test_cubic_spline_halo_1d.F90: n = 64, f = 2(sin x + 2.5 + cos x), alpha = 0.25, si = -2..2: the halo spline
    agrees with the global spline interpolator to 4e-9 (and with the oracle's restatement to 1e-12)."""
    n = 64
    x = np.arange(n + 1) * 2 * np.pi / n
    pdata = 2.0 * (np.sin(x) + 2.5 + np.cos(x))
    F = sb.Field((n,))
    for si in range(-2, 3):
        F.upload(np.asfortranarray(pdata[:n]))
        F.advect_axis_spline_dd(0, np.array([si + 0.25]))
        got = F.download()
        ref = orc.spline_interpolate_array_disp(pdata, 0.0, 2 * np.pi, (si + 0.25) * 2 * np.pi / n)[:n]
        assert np.abs(got - ref).max() <= 4e-9
        exact = orc.spline_dd_advect_axis(np.asfortranarray(pdata[:n].reshape(1, n, 1).copy()), 1, 1, np.array([si + 0.25]),
                                          (1, 1, 0, 1, 1, 0))[0, :, 0]
        assert relerr(got, exact) <= TOL
    F.destroy()


@pytest.mark.parametrize("shape", [(32, 36, 24, 20), (22, 21, 40, 23), (64, 32, 32, 32)])
@pytest.mark.parametrize("staging", [0, 2])
def test_every_axis_vs_oracle(sb, orc, shape, staging):
    """floor(disp) shifts, displacement up to several cells either way, contiguous and strided axes, TMA rows
    (inner % 32 == 0) and the cp.async fallback.  Every axis has >= 16 + |shift| + 1 points: below that the reference
    reads out of bounds (it only asserts np > 15), there is nothing to compare with."""
    rng = np.random.default_rng(SEED)
    f0 = np.asfortranarray(rng.standard_normal(shape))
    F = sb.Field(shape)
    sb.set_staging(staging)
    try:
        for axis in range(4):
            v_axis = (axis + 2) % 4
            disp = rng.uniform(-3.5, 3.5, shape[v_axis])
            disp[0] = 0.0          # no shift table: a zero displacement is advected like any other line
            dsel = _dsel(shape, axis, v_axis)
            ref = orc.spline_dd_advect_axis(f0.copy(order="F"), axis, 1, disp, dsel)
            F.upload(f0)
            F.advect_axis_spline_dd(axis, disp, dsel)
            assert relerr(F.download(), ref) <= TOL, axis
    finally:
        sb.set_staging(0)
    F.destroy()


def test_block_table_and_untouched_lines(sb, orc):
    """make_blocks_spline drives the pass: the v = 0 line is in no block and must stay bit-identical"""
    shape = (32, 20, 18, 24, 2, 2)
    rng = np.random.default_rng(SEED + 1)
    f0 = np.asfortranarray(rng.standard_normal(shape))
    v = -6.0 + 12.0 / shape[3] * np.arange(shape[3])
    disp = -v * 0.31
    shift, alpha, nb = sb.spline_dd_blocks(disp)
    oshift, oalpha, onb = orc.make_blocks_spline(disp)
    assert np.array_equal(shift, oshift) and np.array_equal(alpha, oalpha) and nb == onb
    F = sb.Field(shape)
    for axis in (0, 1, 2):
        dsel = _dsel(shape, axis, 3)
        ref = orc.spline_dd_advect_axis(f0.copy(order="F"), axis, 1, disp, dsel, shifts=oshift)
        F.upload(f0)
        F.advect_axis_spline_dd(axis, disp, dsel, shift=shift)
        got = F.download()
        assert relerr(got, ref) <= TOL
        zero = int(np.where(disp == 0.0)[0][0])
        assert np.array_equal(got[:, :, :, zero], f0[:, :, :, zero])
    F.destroy()


@pytest.mark.parametrize("hw", [(1, 1), (2, 3)])
def test_halo_path_on_one_rank(sb, orc, hw):
    """sllb_dd6d_set_force_halo: K9p (prepare_exchange) + halo pack + the halo-rows kernel on one rank, where the ring
    neighbour is the rank itself -- the reference's own sequence for procs(id) == 1"""
    shape = (32, 4, 2, 24, 20, 36)
    rng = np.random.default_rng(SEED + 2)
    f0 = np.asfortranarray(rng.standard_normal(shape))
    nx3 = shape[0] * shape[1] * shape[2]
    E = rng.uniform(-hw[0], hw[1], nx3) * 0.999
    D = sb.Dd6d(None, shape)
    sb.dd6d_set_force_halo(True)
    try:
        for axis in (3, 4, 5):
            dsel = (1, 1, 0, 1, nx3, 1)
            ref = orc.spline_dd_advect_axis(f0.copy(order="F"), axis, 1, E, dsel)
            D.field().upload(f0)
            D.advect_axis_spline(axis, E, dsel=dsel, hw=hw)
            assert relerr(D.field().download(), ref) <= TOL, axis
    finally:
        sb.dd6d_set_force_halo(False)
    D.destroy()


def test_too_few_points_is_unsupported(sb):
    F = sb.Field((8, 15, 4))
    with pytest.raises(sb.SllbError) as e:
        F.advect_axis_spline_dd(1, np.array([0.3]))
    assert e.value.code == sb.ERR_UNSUPPORTED
    F.destroy()


def test_linearity_and_constants(sb):
    """size-independent properties at a larger size: the pass is linear in f, and a constant field is preserved up to
    the 15-term truncation (2.6e-9 relative, cf. the 4e-9 of the reference's test)"""
    shape = (64, 64, 48, 40)
    rng = np.random.default_rng(SEED + 3)
    a = np.asfortranarray(rng.standard_normal(shape)); b = np.asfortranarray(rng.standard_normal(shape))
    disp = rng.uniform(-1, 1, shape[3])
    F = sb.Field(shape)
    outs = []
    for g in (a, b, 2.0 * a - 3.0 * b, np.full(shape, 1.5, order="F")):
        F.upload(g)
        F.advect_axis_spline_dd(1, disp, _dsel(shape, 1, 3))
        outs.append(F.download())
    assert relerr(outs[2], 2.0 * outs[0] - 3.0 * outs[1]) < 1e-13
    assert 0 < np.abs(outs[3] - 1.5).max() < 1.5 * 8e-9
    F.destroy()


@pytest.mark.parametrize("advector,stencil_x", [(2, 3), (1, 4), (1, 6)])
def test_sim6d_spline_and_centered_vs_oracle(sb, orc, advector, stencil_x):
    """sim_bsl_vp_3d3v_cart_dd_slim with interpolator_type = "spline" / "centered": diagnostics rows and the final
    distribution against the oracle's time loop"""
    n = [20, 20, 20, 18, 18, 18]
    args = (n, 6.0, [4 * np.pi] * 3, stencil_x, 3, 0.05, 0.01, [0.5] * 3)
    S = sb.Sim6d(*args, advector=advector)
    rows = S.run(2)
    f = S.field().download()
    S.destroy()
    orows, of = orc.sim6d(n, 6.0, [4 * np.pi] * 3, stencil_x, 3, 0.05, 2, 0.01, [0.5] * 3, want_f=True, advector=advector)
    assert relerr(f, of) <= 10 * TOL
    big = [1, 2, 3, 4, 5, 6, 7, 11, 12, 13]
    assert np.abs(rows[:, big] / orows[:, big] - 1).max() < 1e-9
    assert np.abs(rows[:, 8:11] - orows[:, 8:11]).max() < 1e-13
