"""The 1D1V simulation as a program: namelist of sim_bsl_vp_1d1v_cart in, the reference's files out
(thdiag.dat with the rho^ / f^ mode columns, .bdat streams, .rst restart), checked against the oracle's time loop and
against numpy transforms of the downloaded state."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NML = """
&geometry
  mesh_case_x1 = "SLL_LANDAU_MESH"
  num_cells_x1 = 32
  x1_min = 0.0
  nbox_x1 = 1
  mesh_case_x2 = "SLL_CARTESIAN_MESH"
  num_cells_x2 = 64
  x2_min = -6.0
  x2_max = 6.0
/
&initial_function
  initial_function_case = "SLL_LANDAU"
  kmode = 0.5
  eps = 0.001
  restart_file = "{restart}"
  time_init_from_restart_file = {from_restart}
/
&time_iterations
  dt = 0.1
  number_iterations = {nit}
  freq_diag = 10
  freq_diag_time = 1
  freq_diag_restart = 10
  nb_mode = 3
  time_init = 0.
  split_case = "SLL_STRANG_VTV"
/
&advector
 advector_x1 = "SLL_SPLINES"
 order_x1 = 4
 advector_x2 = "SLL_SPLINES"
 order_x2 = 4
/
&poisson
  poisson_solver = "SLL_FFT"
/
&drive
  drive_type = "SLL_NO_DRIVE"
/
"""


@pytest.fixture(scope="module")
def sb():
    import selalib_b200 as s
    s.init(0)
    return s


@pytest.fixture(scope="module")
def orc():
    from oracle import orc as o
    return o


def read_thdiag(path, ncol):
    rows = []
    for line in open(path):
        assert len(line.rstrip("\n")) == 25 * ncol            # '(8g25.15)' + '(1g25.15)' cells, no separators
        rows.append([float(line[25 * k:25 * (k + 1)]) for k in range(ncol)])
    return np.array(rows)


def test_thdiag_row_modes_against_numpy(sb):
    S = sb.Sim2d(64, 96, 0.0, 4 * np.pi, -6.0, 6.0, 1, 0.5, 0.01, 0.05)
    r8 = S.run(3)
    nb = 7
    row = S.thdiag(nb)
    assert row.shape == (8 + 3 * (nb + 1),)
    assert np.abs(row[:8] / r8[-1] - 1).max() < 1e-13
    f = S.field().download()
    rho, _ = S.fields()
    rh = np.fft.fft(rho) / rho.size
    assert np.abs(row[8:8 + 2 * (nb + 1):2] - rh.real[:nb + 1]).max() < 1e-15
    im = rh.imag[:nb + 1].copy(); im[0] = 0.0
    assert np.abs(row[9:9 + 2 * (nb + 1):2] - im).max() < 1e-15
    fh = (np.abs(np.fft.fft(f, axis=0) / f.shape[0]) ** 2).sum(axis=1) * (12.0 / 96)
    got = row[8 + 2 * (nb + 1):]
    assert np.abs(got - fh[:nb + 1]).max() < 1e-14 * fh[0] and np.abs(got[:3] / fh[:3] - 1).max() < 1e-11
    S.destroy()


def test_run_namelist_files_and_restart(sb, orc, tmp_path):
    a = tmp_path / "a"; b = tmp_path / "b"; a.mkdir(); b.mkdir()
    (a / "in.nml").write_text(NML.format(restart="no_restart_file", from_restart=".false.", nit=20))
    sb.sim2d_run_namelist(str(a / "in"), str(a))          # extension appended like the reference does
    ncol = 8 + 3 * 4
    th = read_thdiag(a / "thdiag.dat", ncol)
    assert th.shape == (20, ncol)
    orows, of, _ = orc.sim2d(32, 64, 0.0, 4 * np.pi, -6.0, 6.0, 0, 0.5, 1e-3, 0.1, 20, method=0, want_f=True)
    for col in (0, 1, 2, 4, 5, 7):                         # time, mass, l1, l2, ekin, etot (written with 15 digits)
        assert np.abs(th[:, col] / orows[:, col] - 1).max() < 1e-10, col
    assert np.abs(th[:, 6] - orows[:, 6]).max() / orows[:, 6].max() < 1e-9       # potential energy
    # streams of doubles: node positions, equilibrium, deltaf at t = 0, 1.0, 2.0, fields at every diagnostic step
    x = np.fromfile(a / "x.bdat"); v = np.fromfile(a / "v.bdat")
    assert np.allclose(x, 4 * np.pi / 32 * np.arange(32)) and np.allclose(v, -6 + 12 / 64 * np.arange(64))
    f0 = np.fromfile(a / "f0.bdat").reshape((32, 64), order="F")
    assert np.abs(f0 - np.exp(-0.5 * v[None, :] ** 2) / np.sqrt(2 * np.pi)).max() < 1e-15
    df = np.fromfile(a / "deltaf.bdat").reshape((32, 64, 3), order="F")
    assert np.abs(df[:, :, 2] + f0 - of[:-1, :-1]).max() < 1e-12
    t = np.fromfile(a / "t.bdat")
    assert t.shape == (21,) and np.allclose(t, 0.1 * np.arange(21))
    ef = np.fromfile(a / "efield.bdat").reshape((32, 21), order="F")
    assert np.abs(0.5 * (ef[:, 1:] ** 2).sum(axis=0) * (4 * np.pi / 32) / th[:, 6] - 1).max() < 1e-12
    assert np.fromfile(a / "rhotot.bdat").size == 32 * 21
    # restart files every 10 steps: time + (N1+1)(N2+1) doubles; iplot advances with freq_diag
    rst = np.fromfile(a / "f_plot_0001_proc_0000.rst")
    assert rst.size == 1 + 33 * 65 and abs(rst[0] - 1.0) < 1e-14
    assert os.path.exists(a / "f_plot_0002_proc_0000.rst")
    g = rst[1:].reshape((33, 65), order="F")
    assert np.array_equal(g[-1, :], g[0, :]) and np.array_equal(g[:, -1], g[:, 0])
    # restart from t = 1.0 and run the remaining 10 steps: the same states as the uninterrupted run
    (b / "in.nml").write_text(NML.format(restart=str(a / "f_plot_0001"), from_restart=".true.", nit=10))
    sb.sim2d_run_namelist(str(b / "in.nml"), str(b))
    th2 = read_thdiag(b / "thdiag.dat", ncol)
    assert th2.shape == (10, ncol)
    assert np.abs(th2[:, 0] - th[10:, 0]).max() < 1e-12
    # The restarted run recomputes E from the stored f (like the reference, :1365-1384) while the uninterrupted one still
    # holds the E of its last T stage; the V stage conserves rho only to rounding, and rho = 1 - int f cancels three
    # digits, so the two E differ by ~1e-13 relative.  The eight integrals column by column (the field energy sees that
    # difference directly), the mode columns (many of them rounding noise) against the largest mode.
    scale = np.abs(th[10:, :8]).max(axis=0)
    scale[3] = scale[1]                                     # momentum ~ 0: measured against the mass
    assert (np.abs(th2[:, :8] - th[10:, :8]) / scale).max() < 1e-11
    assert np.abs(th2[:, 8:] - th[10:, 8:]).max() < 1e-11 * np.abs(th[10:, 8:]).max()


def test_namelist_knobs_and_refusals(sb, tmp_path):
    p = tmp_path / "k.nml"
    txt = NML.format(restart="no_restart_file", from_restart=".false.", nit=2)
    p.write_text(txt.replace('"SLL_STRANG_VTV"', '"SLL_ORDER6VPNEW1_VTV"').replace('advector_x2 = "SLL_SPLINES"', 'advector_x2 = "SLL_LAGRANGE"')
                 .replace("order_x2 = 4", "order_x2 = 6").replace('"SLL_LANDAU"', '"SLL_BUMP_ON_TAIL"'))
    S, nit, fdt, nbm = sb.Sim2d.from_namelist(str(p))
    assert (nit, fdt, nbm) == (2, 1, 3)
    rows = S.run(2)
    assert np.isfinite(rows).all() and abs(rows[-1, 1] / rows[0, 1] - 1) < 1e-10     # mass conserved
    S.destroy()
    for old, new in (('"SLL_STRANG_VTV"', '"SLL_ORDER6VPOT_VTV"'), ('"SLL_CARTESIAN_MESH"', '"SLL_TWO_GRID_MESH"'),
                     ('"SLL_NO_DRIVE"', '"SLL_KEEN_DRIVE"'), ('"SLL_LANDAU"', '"SLL_BEAM"')):
        p.write_text(txt.replace(old, new))
        with pytest.raises(sb.SllbError):
            sb.Sim2d.from_namelist(str(p))
