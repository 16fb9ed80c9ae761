"""Ensemble streaming (sllb_sim4d_stream_step): upload of the next state and download of the previous one overlap the
current state's time step; the results must be bit-identical to upload -> run(1) -> download, state by state."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import selalib_b200 as s
    s.init(0)
    return s


def test_stream_step_equals_serial_steps(sb):
    import torch
    nc = [32, 32, 32, 32]
    args = (nc, [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-2, 0.1)
    S = sb.Sim4d(*args)
    f0 = S.field().download()
    rng = np.random.default_rng(20261017)
    members = [np.asfortranarray(f0 * (1.0 + 0.05 * k) + 1e-3 * rng.standard_normal(f0.shape)) for k in range(4)]
    serial = []
    for m in members:
        S.field().upload(m)
        S.run(1, diagnostics=False)
        serial.append(S.field().download())
    n = f0.size
    hin = [torch.from_numpy(np.ascontiguousarray(m.reshape(-1, order="F"))).pin_memory() for m in members]
    hout = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in members]
    # N members take N + 2 calls
    for k in range(len(members) + 2):
        nxt = hin[k].data_ptr() if k < len(members) else None
        prv = hout[k - 2].data_ptr() if k >= 2 else None
        S.stream_step(nxt, prv)
    for k, ref in enumerate(serial):
        got = hout[k].numpy().reshape(f0.shape, order="F")
        assert np.array_equal(got, ref), k
    # the object is still usable as an ordinary simulation afterwards
    S.field().upload(members[0])
    S.run(1, diagnostics=False)
    assert np.array_equal(S.field().download(), serial[0])
    S.destroy()


def test_sim2d_cuda_graph_replay_equals_stream_launches(sb):
    """sllb_sim2d_run replays one recorded time step as a CUDA graph (the 1D1V step is launch-bound): same values, bit for
    bit, as one launch per kernel, with and without per-step diagnostics"""
    out = {}
    for graphs in (1, 0):
        sb.set_cuda_graphs(graphs)
        try:
            S = sb.Sim2d(128, 256, 0.0, 4 * np.pi, -6.0, 6.0, 1, 0.5, 0.01, 0.01, method=sb.METHOD_LAGRANGE_FIXED, order=7)
            rows = S.run(6)
            S.run(20, diagnostics=False)
            rows2 = S.run(3)
            f = S.field().download()
            S.destroy()
        finally:
            sb.set_cuda_graphs(1)
        out[graphs] = (rows, rows2, f)
    for a, b in zip(out[1], out[0]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", [(64, 1024), (40, 512), (1024, 1024), (8, 200, 3)])
def test_lagrange_few_long_lines_vs_oracle(sb, shape):
    """the chunked strided Lagrange kernel (few, long lines: the 1D1V shapes) against the oracle, every stencil"""
    from oracle import orc
    rng = np.random.default_rng(20261017)
    f0 = np.asfortranarray(rng.standard_normal(shape))
    F = sb.Field(shape)
    disp = rng.uniform(-1.0, 1.0, shape[0])
    dsel = (1, 1, 0, 1, shape[0], 1)
    for method, mcode, orders in (("lagrange_fixed", sb.METHOD_LAGRANGE_FIXED, (3, 5, 7, 9, 11)),
                                  ("lagrange_centered", sb.METHOD_LAGRANGE_CENTERED, (4, 6, 8))):
        for order in orders:
            ref = orc.advect_axis(f0.copy(order="F"), 1, method, order, disp, dsel)
            F.upload(f0)
            F.advect_axis(1, mcode, order, disp, 1.0, dsel)
            err = np.abs(F.download() - ref).max() / np.abs(ref).max()
            assert err <= 1e-12, (method, order, err)
    F.destroy()


@pytest.mark.parametrize("shape", [(32, 32, 6, 5, 4), (20, 12, 7, 3), (64, 16, 9)])
def test_lagrange_plane_kernel_equals_two_passes(sb, shape):
    """K2d: the eta1 and eta2 Lagrange passes in one sweep (what the 3D3V x-advection runs) are bit-identical to the two
    separate passes, every stencil, power-of-two and other plane shapes"""
    rng = np.random.default_rng(20261017)
    f0 = np.asfortranarray(rng.standard_normal(shape))
    F = sb.Field(shape)
    v0 = rng.uniform(-1.2, 1.2, shape[2])            # eta1 displacement depends on axis 2
    v1 = rng.uniform(-1.2, 1.2, shape[-1])           # eta2 displacement depends on the last axis
    ds0 = (shape[1], shape[2], 1, 1, 1, 0)
    stride = int(np.prod(shape[2:-1], dtype=np.int64))
    ds1 = (stride, shape[-1], 1, 1, 1, 0)
    for mcode, orders in ((sb.METHOD_LAGRANGE_FIXED, (3, 5, 7, 9, 11)), (sb.METHOD_LAGRANGE_CENTERED, (4, 6, 8))):
        for order in orders:
            F.upload(f0)
            F.advect_axis(0, mcode, order, v0, 1.0, ds0)
            F.advect_axis(1, mcode, order, v1, 1.0, ds1)
            ref = F.download()
            F.upload(f0)
            F.advect_plane(v0, ds0, 1.0, v1, ds1, 1.0, method=mcode, order=order)
            assert np.array_equal(F.download(), ref), (mcode, order)
    F.destroy()


def test_sim2d_reproduces_the_shipped_1d1v_trace(sb):
    """G4 on the GPU: vpsim2d_cartesian_input.nml (Landau damping 32 x 64, cubic splines, Strang VTV, dt = 0.1, 600 steps)
    against the L2-norm and potential-energy columns of the reference's shipped vpsim2d_cartesian_ref.dat
    (tests/golden/vpsim2d_cartesian_ref_l2_epot.dat): L2 to the printed 12 digits, field energy to 1e-8 of its maximum at every step"""
    import os
    gold = np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vpsim2d_cartesian_ref_l2_epot.dat"))
    S = sb.Sim2d(32, 64, 0.0, 4 * np.pi, -6.0, 6.0, 0, 0.5, 0.001, 0.1, method=sb.METHOD_SPLINE, order=4)
    rows = S.run(600)
    S.destroy()
    assert np.abs(rows[:, 0] - gold[:, 0]).max() < 1e-9
    assert np.abs(rows[:, 4] / gold[:, 1] - 1).max() < 1e-11
    # the field energy oscillates through near-zero minima: compare on the scale of its maximum
    assert np.abs(rows[:, 6] - gold[:, 2]).max() < 1e-8 * gold[:, 2].max()


@pytest.mark.parametrize("order", [10, 12, 14, 16, 18])
def test_high_order_periodic_lagrange_vs_oracle(sb, order):
    """sll_p_lagrange of sll_s_periodic_interp takes any even order (the shipped two-stream namelist uses 18): orders
    beyond the closed forms, batched on a contiguous axis, two strided axes (TMA rows and the cp.async fallback) and through
    the line object sll_t_advector_1d_periodic"""
    from oracle import orc
    rng = np.random.default_rng(20261017 + order)
    for shape in ((64, 32, 40), (36, 24, 30)):
        f0 = np.asfortranarray(rng.standard_normal(shape))
        F = sb.Field(shape)
        for axis in range(3):
            v_axis = (axis + 1) % 3
            disp = rng.uniform(-2.5, 2.5, shape[v_axis])
            if v_axis > axis:
                dsel = (1, shape[v_axis], 1, 1, 1, 0)
            else:
                dsel = (1, 1, 0, 1, shape[v_axis], 1)
            ref = orc.advect_axis(f0.copy(order="F"), axis, "fft_lagrange", order, disp, dsel)
            F.upload(f0)
            F.advect_axis(axis, sb.METHOD_LAGRANGE_CENTERED, order, disp, 1.0, dsel)
            err = np.abs(F.download() - ref).max() / np.abs(ref).max()
            assert err <= 1e-12, (shape, axis, err)
        F.destroy()
    nc = 48
    A = sb.Advector1dPeriodic(nc, 0.0, 2.0, kind=sb.ADV_PERIODIC_LAGRANGE, order=order)
    inp = rng.standard_normal(nc + 1); inp[-1] = inp[0]
    out = A.advect_1d_constant(0.37, 0.21, inp)
    ref = orc.advect_1d_periodic_constant("lagrange", nc, 0.0, 2.0, order, 0.37, 0.21, inp)
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(inp).max()
    A.delete()
