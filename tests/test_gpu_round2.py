"""GPU tests of the round-2 pieces of the 2D2V time loop: the three-kernel dense-DFT Poisson solve against the cuFFT route
and the oracle, the per-step diagnostics computed on the device (fused into the last x4 pass of a step) against the
host-side route, the position-weighted checksum, and the continuation of a run across calls."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 20261018
XMIN, XMAX = [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6]


@pytest.fixture(scope="module")
def sb():
    import selalib_b200 as s
    s.init(0)
    return s


@pytest.fixture(scope="module")
def orc():
    from oracle import orc as o
    return o


def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("n", [(128, 128), (64, 32), (16, 24), (8, 6), (256, 200)])
def test_direct_poisson2d_equals_cufft_route_and_oracle(sb, orc, n):
    """sll_t_poisson_2d_periodic semantics (kx(1,1) := 1, negative Nyquist in x2, FFTW c2r on the non-Hermitian columns):
    dense-DFT kernels == cuFFT + k_poisson2d == oracle."""
    rng = np.random.default_rng(SEED + n[0])
    rho = np.asfortranarray(rng.standard_normal(n))
    xmin, xmax = (0.0, -1.0), (4 * np.pi, 2.5)
    P = sb.Poisson(n, xmin, xmax)
    try:
        sb.set_poisson_direct(True)
        d = P.solve(rho)
        sb.set_poisson_direct(False)
        c = P.solve(rho)
    finally:
        sb.set_poisson_direct(True)
        P.destroy()
    for a, b in zip(d, c):
        assert relmax(a, b) < 1e-13
    ex, ey, phi = orc.poisson_2d(rho, n[0], n[1], xmin[0], xmax[0], xmin[1], xmax[1], want_phi=True)
    for a, b in zip(d, (phi, ex, ey)):
        assert relmax(a, b) < 1e-13


@pytest.mark.parametrize("n", [(128, 128), (48, 20), (512, 512)])
def test_poisson_2d_periodic_par_variant(sb, orc, n):
    """a15: sll_t_poisson_2d_periodic_par (Delta phi = rho, zero mean, the solver of sim_bsl_vp_2d2v_cart) on both routes
    (dense-DFT kernels up to 256, cuFFT + k_poisson2d_par beyond) against the oracle and the reference's own known answer
    (test_poisson_2d_periodic_par.F90: phi = cos x sin y, rho = -2 phi, average error <= 1e-6)."""
    rng = np.random.default_rng(SEED + n[1])
    L = (2 * np.pi, 2 * np.pi)
    P = sb.Poisson(n, (0.0, 0.0), L, par=True)
    rho = rng.standard_normal((n[0] + 1, n[1] + 1)); rho[-1, :] = rho[0, :]; rho[:, -1] = rho[:, 0]
    rho = np.asfortranarray(rho)
    phi = P.solve(rho)
    ref = orc.poisson_2d_par(rho, n[0], n[1], L[0], L[1])
    assert relmax(phi, ref) < 1e-13
    if n[0] <= 256:
        sb.set_poisson_direct(False)
        try:
            assert relmax(P.solve(rho), ref) < 1e-13
        finally:
            sb.set_poisson_direct(True)
    x = np.arange(n[0] + 1) * L[0] / n[0]; y = np.arange(n[1] + 1) * L[1] / n[1]
    phi_an = np.cos(x)[:, None] * np.sin(y)[None, :]
    got = P.solve(np.asfortranarray(-2.0 * phi_an))
    assert np.abs(got - phi_an)[:-1, :-1].sum() / (n[0] * n[1]) < 1e-6      # the reference's threshold
    assert np.abs(got - phi_an).max() < 1e-13
    import ctypes as C
    from selalib_b200.capi import _p
    e = np.zeros_like(rho, order="F")     # the parallel solver has no field outputs: asking for E is refused
    rc = sb.lib().sllb_poisson_solve_host(P.h, _p(rho), (C.c_int * 2)(*rho.shape), _p(e), _p(e), None, None)
    assert rc == sb.ERR_INVALID
    P.destroy()


@pytest.mark.parametrize("split", [0, 1])
def test_device_diagnostics_rows_equal_host_route(sb, split):
    """rows of sllb_sim4d_run (device reductions, moments fused into the last x4 pass for VTV; row sums for TVT) ==
    sllb_sim4d_diagnostics (row sums brought to the host) at the same states."""
    nc = [32, 32, 32, 32]
    S = sb.Sim4d(nc, XMIN, XMAX, 0.5, 0.5, 1e-3, 0.1, split=split)
    got, want = [], []
    for _ in range(3):
        got.append(S.run(1)[0])
        want.append(S.diagnostics())
    S.destroy()
    got, want = np.array(got), np.array(want)
    assert np.abs(got / want - 1).max() < 1e-13, np.abs(got / want - 1).max(axis=0)


def test_rows_independent_of_call_pattern(sb):
    """run(3) and run(1) x 3 give identical rows and f (the fused line moments never leak into the arithmetic of f)."""
    nc = [32, 32, 32, 32]
    S = sb.Sim4d(nc, XMIN, XMAX, 0.5, 0.5, 1e-3, 0.1)
    r3 = S.run(3)
    f3 = S.field().download()
    S.destroy()
    S = sb.Sim4d(nc, XMIN, XMAX, 0.5, 0.5, 1e-3, 0.1)
    r1 = np.vstack([S.run(1), S.run(1), S.run(1)])
    f1 = S.field().download()
    S.destroy()
    assert np.array_equal(f1, f3)
    assert np.abs(r1 / r3 - 1).max() < 1e-14


def test_checksum_detects_misplaced_elements(sb):
    nc = [16, 16, 32, 32]
    S = sb.Sim4d(nc, XMIN, XMAX, 0.5, 0.5, 1e-3, 0.1)
    F = S.field()
    f = F.download()
    c0 = S.checksum()
    i0, i1, i2, i3 = np.meshgrid(*[np.arange(n) for n in nc], indexing="ij")
    w = 1.0 + ((3 * i0 + 5 * i1 + 7 * i2 + 11 * i3) & 63) / 64.0
    assert abs(c0[0] / (w * f).sum() - 1) < 1e-13 and abs(c0[1] / (w * f * f).sum() - 1) < 1e-13
    g = f.copy(order="F")
    g[3, 5, 7, 9], g[3, 5, 15, 16] = f[3, 5, 15, 16], f[3, 5, 7, 9]    # swap two elements: the mass does not notice
    F.upload(g)
    c1 = S.checksum()
    assert abs(g.sum() - f.sum()) < 1e-18 + 1e-15 * abs(f.sum())
    assert abs(c1[0] - c0[0]) > 1e-9 * abs(c0[0]) or abs(c1[1] - c0[1]) > 1e-9 * abs(c0[1])
    S.destroy()


@pytest.mark.parametrize("split", [0, 1, "SLL_ORDER6VPnew1_VTV"])
def test_recorded_time_step_equals_plain_launches(sb, split):
    """the 2D2V step replayed as a CUDA graph (from the third step on) == one launch per kernel, bit for bit: f, the
    diagnostics rows (slot and time come from device-side counters) and the rows of a second run on the same handle"""
    nc = [32, 32, 32, 32]
    out = {}
    for graphs in (False, True):
        sb.set_cuda_graphs(graphs)
        try:
            S = sb.Sim4d(nc, XMIN, XMAX, 0.5, 0.5, 1e-3, 0.1, split=split)
            sb.launch_count_reset()
            r1 = S.run(6)
            n1 = sb.launch_count()
            S.run(2, diagnostics=False)
            r2 = S.run(3)
            if split == 0:   # a run longer than the row buffer: it grows, and the recordings that carry its address go
                r2 = np.vstack([r2, S.run(4100)[[0, 1, 4098, 4099]]])
            out[graphs] = (r1, r2, S.field().download(), n1)
            S.destroy()
        finally:
            sb.set_cuda_graphs(True)
    for a, b in zip(out[False][:3], out[True][:3]):
        assert np.array_equal(a, b)
    assert out[False][3] == out[True][3]          # same kernels, counted per replay
    assert np.allclose(out[True][0][:, 0], 0.1 * np.arange(1, 7)) and np.allclose(out[True][1][:3, 0], 0.1 * np.arange(9, 12))
    assert np.isfinite(out[True][1]).all() and (out[True][1][:, 3] > 1.0).all()     # every row was really written (mass)
