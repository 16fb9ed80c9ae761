"""GPU tests of the slim 6D decomposition path (a11/a12) on ONE device: the halo pack kernel (K7) and the
halo-cells Lagrange kernel (K2c) are forced on (sllb_dd6d_set_force_halo) so that the exact sequence the
reference runs -- local periodic halo copy, then sll_s_lagrange_interpolation_1d_fast_disp_fixed_haloc_cells
on left|local|right -- is checked against the oracle.  The NCCL exchange itself is covered by
tests/mgpu/run_mgpu.py (needs >= 2 GPUs; launched by tests/test_multi_gpu.py)."""
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


def test_halo_exchange_single_rank_is_periodic_copy(sb):
    """procs(axis) == 1: halos are the periodic neighbours (sll_m_decomposition.F90:1840-1861,1958-1979)."""
    shape = (6, 5, 4, 7, 6, 5)
    f0 = np.asfortranarray(np.arange(np.prod(shape), dtype=np.float64).reshape(shape, order="F"))
    D = sb.Dd6d(None, shape)
    assert D.procs == (1,) * 6 and D.nw == shape and D.mn == (0,) * 6
    D.field().upload(f0)
    for axis in range(6):
        for hl, hr in ((1, 1), (3, 2), (0, 2), (4, 4)):
            if max(hl, hr) > shape[axis]:
                continue
            D.halo_exchange(axis, hl, hr)
            n = shape[axis]
            assert np.array_equal(D.halo(1), np.take(f0, range(0, hr), axis=axis))
            assert np.array_equal(D.halo(0), np.take(f0, range(n - hl, n), axis=axis))
    D.destroy()


@pytest.mark.parametrize("stencil", [3, 5, 7, 9, 11])
@pytest.mark.parametrize("staging", [0, 2])
def test_halo_cells_kernel_vs_oracle(sb, orc, stencil, staging):
    rng = np.random.default_rng(SEED + stencil)
    shape = (32, 6, 2, 12, 6, 16)          # inner of every axis >= 1 is a multiple of 32: TMA path
    f0 = np.asfortranarray(rng.standard_normal(shape))
    E = rng.uniform(-1.2, 1.2, shape[0] * shape[1] * shape[2])
    nx3 = shape[0] * shape[1] * shape[2]
    D = sb.Dd6d(None, shape)
    sb.dd6d_set_force_halo(True)
    sb.set_staging(staging)
    try:
        for axis in (1, 3, 4, 5):
            if axis >= 3:
                disp, dsel = E, (1, 1, 0, 1, nx3, 1)
            else:
                disp = rng.uniform(-1.0, 1.0, shape[axis + 3])
                stride = int(np.prod(shape[axis + 1:axis + 3]))
                dsel = (stride, shape[axis + 3], 1, 1, 1, 0)
            ref = orc.advect_axis(f0.copy(order="F"), axis, "lagrange_fixed", stencil, disp, dsel)
            D.field().upload(f0)
            D.advect_axis(axis, stencil, disp, 1.0, dsel)
            assert relerr(D.field().download(), ref) < TOL, axis
    finally:
        sb.dd6d_set_force_halo(False)
        sb.set_staging(0)
    D.destroy()


def test_halo_cells_kernel_ragged(sb, orc):
    """inner not a multiple of 32 (cp.async staging) and a block narrower than the tile."""
    rng = np.random.default_rng(SEED + 77)
    shape = (5, 3, 2, 9, 7, 8)
    f0 = np.asfortranarray(rng.standard_normal(shape))
    nx3 = shape[0] * shape[1] * shape[2]
    E = rng.uniform(-0.9, 0.9, nx3)
    D = sb.Dd6d(None, shape)
    sb.dd6d_set_force_halo(True)
    try:
        for axis in (3, 4, 5):
            ref = orc.advect_axis(f0.copy(order="F"), axis, "lagrange_fixed", 7, E, (1, 1, 0, 1, nx3, 1))
            D.field().upload(f0)
            D.advect_axis(axis, 7, E, 1.0, (1, 1, 0, 1, nx3, 1))
            assert relerr(D.field().download(), ref) < TOL, axis
        with pytest.raises(sb.SllbError):
            D.advect_axis(0, 7, E, 1.0, (1, 1, 0, 1, 1, 0))      # split contiguous axis: not implemented
    finally:
        sb.dd6d_set_force_halo(False)
    D.destroy()


def test_sim6d_golden_through_halo_path(sb):
    """G1 golden file with the reference's own sequence (halo copy + halo-cells stencil) on every v axis."""
    gold = np.loadtxt(os.path.join(os.path.dirname(__file__), "golden", "reffile_bsl_vp_3d3v_cart_dd.dat"))
    args = ([16] * 6, 6.0, [12.5663706144] * 3, 3, 3, 0.01, 0.01, [0.499999999998376] * 3)
    S0 = sb.Sim6d(*args)
    r0 = S0.run(2)
    f0 = S0.field().download()
    S0.destroy()
    sb.dd6d_set_force_halo(True)
    try:
        S = sb.Sim6d(*args)
        rows = S.run(2)
        f = S.field().download()
        S.destroy()
    finally:
        sb.dd6d_set_force_halo(False)
    assert np.abs(rows - gold).max() < 5e-7
    assert np.array_equal(rows, r0) and np.array_equal(f, f0)   # identical arithmetic, bit for bit


@pytest.mark.parametrize("in_phase", [False, True])
def test_sim6d_run_in_two_calls_equals_one_call(sb, in_phase):
    """run(a) + run(b) == run(a + b): with time_in_phase the first call ends with the closing half V step and the second
    one opens with the other half (no half step lost at the call boundary); rows written = nsteps (+1 on the first call)."""
    args = ([8, 8, 8, 16, 16, 16], 6.0, [12.5663706144] * 3, 5, 5, 0.01, 0.01, [0.5] * 3)
    S1 = sb.Sim6d(*args, time_in_phase=in_phase)
    r1 = S1.run(3)
    f1 = S1.field().download()
    S1.destroy()
    S2 = sb.Sim6d(*args, time_in_phase=in_phase)
    ra = S2.run(1)
    rb = S2.run(2)
    f2 = S2.field().download()
    S2.destroy()
    assert r1.shape == (4, 14) and ra.shape == (2, 14) and rb.shape == (2, 14)
    r2 = np.vstack([ra, rb])
    if in_phase:
        # two half V steps instead of one whole at the call boundary: the same Strang sequence, but interpolating twice
        # by d/2 is not interpolating once by d -- the two runs differ by that interpolation error (here 5-point
        # Lagrange, displacements ~1e-3 cells), far below what losing the half step would give (~1e-4)
        assert np.abs(r2 - r1).max() <= 1e-8 * np.abs(r1).max()
        assert np.abs(f2 - f1).max() <= 1e-8 * np.abs(f1).max()
        # ... and the half step is NOT lost: a run that skips it is three orders of magnitude further away
        assert np.abs(f2 - f1).max() > 0
    else:          # identical sequence of kernels
        assert np.array_equal(r2, r1) and np.array_equal(f2, f1)


def test_sim6d_clocks_file(sb, tmp_path):
    """sll_clocks.txt (sll_s_finalize_clocks, sll_m_sim_6d_utilities.F90:775-796): the labels the reference's time loop
    uses, wall-clock seconds, only the ones that ran; and the stopwatches do not change the results"""
    args = ([8, 8, 8, 16, 16, 16], 6.0, [12.5663706144] * 3, 5, 5, 0.01, 0.01, [0.5] * 3)
    S0 = sb.Sim6d(*args)
    r0 = S0.run(2)
    S0.destroy()
    S = sb.Sim6d(*args)
    S.set_clocks(True)
    r = S.run(2)
    path = str(tmp_path / "sll_clocks.txt")
    S.write_clocks(path)
    S.destroy()
    assert np.array_equal(r, r0)
    got = {}
    for line in open(path):
        label, val = line.split()
        got[label] = float(val)
    assert set(got) == {"D", "P", "PC", "PF", "V", "X", "X1", "X2", "X3", "X4", "X5", "X6"}      # one GPU: no H4..H6
    assert all(v > 0 for v in got.values())
    assert got["X"] >= got["X1"] + got["X2"] + got["X3"] - 1e-9 and got["P"] >= got["PC"] + got["PF"] - 1e-9
    assert list(got) == sorted(got)                       # table order = ASCII order of the labels


def test_cpp_interface_of_the_reference_simulation(sb, tmp_path):
    """Our restatement of simulations/parallel/bsl_vp_3d3v_cart_dd/test_cpp_interface.cpp: a C++ host drives the
    simulation through the reference's own C symbols (namelist in, <prefix>.dat out) and the result is checked
    against the golden file with the reference's tolerance; passes on the string 'works in cpp' like the CTest."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "selalib_b200", "lib")
    exe = str(tmp_path / "test_cpp_interface_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_cpp_interface_b200.cpp"), "-o", exe,
                           "-L" + libdir, "-lsllb200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, os.path.join(root, "tests", "golden", "param_6d_golden.nml"),
                          os.path.join(root, "tests", "golden", "reffile_bsl_vp_3d3v_cart_dd.dat")],
                         cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "PASSED." in out.stdout and "works in cpp" in out.stdout
    # the file is in the reference's e20.12 layout: 3 rows of 14 numbers, 20 columns each
    lines = open(tmp_path / "vp_3d3v_dd_b200.dat").read().splitlines()
    assert len(lines) == 3 and all(len(ln) == 280 for ln in lines)
    gold = np.loadtxt(os.path.join(root, "tests", "golden", "reffile_bsl_vp_3d3v_cart_dd.dat"))
    assert np.abs(np.loadtxt(tmp_path / "vp_3d3v_dd_b200.dat") - gold).max() < 5e-7
    assert lines[0].startswith("  0.000000000000E+00  0.999999986099E+00")
