"""GPU parity of the 2D2V time loop under every family of splitting schedule of sll_m_time_splitting_coeff
(SURVEY.md section 8(f) rank 3), including the dim_split_V = 2 schemes whose V stages also move along the field of the
modified potential (compute_jacobian + a second Poisson solve), against the oracle's time loop.
Tolerances as in test_sim4d_trace (tests/test_gpu_parity.py): the reference carries duplicated velocity end planes
whose trapezoid weight shows up in the field energy at the 1e-6 level."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


CASES = ["SLL_LIE_VT", "SLL_TRIPLE_JUMP_TVT", "SLL_TRIPLE_JUMP_VTV", "SLL_ORDER6_VTV", "SLL_ORDER6_TVT", "SLL_ORDER6VP_TVT",
         "SLL_ORDER6VP_VTV", "SLL_ORDER6VPnew_TVT", "SLL_ORDER6VPnew1_VTV", "SLL_ORDER6VPnew2_VTV", "SLL_ORDER6VP2D_VTV",
         "SLL_ORDER6VPOT_VTV", "SLL_ORDER6VPOTnew1_VTV", "SLL_ORDER6VPOTnew2_VTV", "SLL_ORDER6VPOTnew3_VTV"]


@pytest.mark.parametrize("case", CASES)
def test_sim4d_splitting_case_vs_oracle(sb, orc, case):
    nc = [16, 16, 32, 32]
    xmin, xmax = [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6]
    nsteps = 3
    S = sb.Sim4d(nc, xmin, xmax, 0.5, 0.5, 1e-2, 0.1, split=case)
    th0 = S.thdiag()
    rows = S.run(nsteps)
    th = S.thdiag()
    f = S.field().download([1, 1, 1, 1])
    S.destroy()
    orows, of, othd = orc.sim4d(nc, xmin, xmax, 0.5, 0.5, 1e-2, 0.1, nsteps, split=case, method=0, want_f=True, want_thdiag=True)
    assert relerr(f[:-1, :-1, :-1, :-1], of[:-1, :-1, :-1, :-1]) < 1e-10
    assert np.abs(rows[:, 1] / orows[1:, 1] - 1).max() < 1e-6     # field energy
    assert np.abs(rows[:, 3] / orows[1:, 3] - 1).max() < 1e-8     # mass
    assert np.abs(rows[:, 2] / orows[1:, 2] - 1).max() < 1e-8     # kinetic energy
    # thdiag rows (13 columns) at t = 0 and after the run
    for got, ref in ((th0, othd[0]), (th, othd[nsteps])):
        exact = [0, 3, 4, 10, 11, 12]                              # time and the analytic columns
        assert np.abs(got[exact] - ref[exact]).max() <= 1e-13 * np.abs(ref[exact]).max()
        assert np.abs(got[[2, 7, 8, 9]] / ref[[2, 7, 8, 9]] - 1).max() < 1e-8
        assert abs(got[1] / ref[1] - 1) < 1e-6 and abs(got[5] / ref[5] - 1) < 1e-5
        if ref[6] != 0.0:
            assert abs(got[6] / ref[6] - 1) < 1e-5
        else:
            assert got[6] == 0.0


def test_sim4d_modified_potential_changes_the_result(sb):
    """the second field pair really is used: ORDER6VPOT differs from ORDER6VP2D (same first weights) beyond rounding"""
    nc = [16, 16, 32, 32]
    args = (nc, [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 0.3, 0.25)
    out = []
    for case in ("SLL_ORDER6VP2D_VTV", "SLL_ORDER6VPOT_VTV"):
        S = sb.Sim4d(*args, split=case)
        S.run(2)
        out.append(S.field().download())
        S.destroy()
    assert 1e-9 < relerr(out[1], out[0]) < 1e-2


NML = """
&geometry
  mesh_case_x1="SLL_LANDAU_MESH"
  num_cells_x1 = 16
  x1_min = 0.0
  nbox_x1 = 1
  mesh_case_x2="SLL_LANDAU_MESH"
  num_cells_x2 = 16
  x2_min = 0.0
  nbox_x2 = 1
  mesh_case_x3="SLL_CARTESIAN_MESH"
  num_cells_x3 = 32
  x3_min = -6.
  x3_max = 6.
  mesh_case_x4="SLL_CARTESIAN_MESH"
  num_cells_x4 = 32
  x4_min = -6.
  x4_max = 6.
/

&initial_function
  initial_function_case="SLL_LANDAU"
  kmode_x1 = 0.5
  kmode_x2 = 0.5
  eps = 1e-3
/

&time_iterations
  dt = %(dt)s
  number_iterations = 4
  freq_diag = 20
  freq_diag_time = 2
  !split_case = "SLL_STRANG_VTV"
  split_case = "SLL_ORDER6VPnew1_VTV"
/

&advector
 advector_x1 = "%(adv)s"
 order_x1 = 4
 advector_x2 = "%(adv)s"
 order_x2 = 4
 advector_x3 = "%(adv)s"
 order_x3 = 4
 advector_x4 = "%(adv)s"
 order_x4 = 4
/
&poisson
 stencil_r=-3
 stencil_s=3
/
"""


@pytest.mark.parametrize("adv,method", [("SLL_LAGRANGE", 2), ("SLL_SPLINES", 0)])
def test_namelist_front_end_writes_thdiag(sb, orc, tmp_path, adv, method):
    """the shipped vpsim4d_cartesian_input.nml (simulations/parallel/bsl_vp_2d2v_cart_poisson_serial; Lagrange order 4,
    SLL_ORDER6VPnew1_VTV, stencil -3..3) with a smaller dt drives the GPU build; the thdiag file has the reference's
    format and its rows match the oracle's time loop with the reference-algorithm advectors (FFT sll_p_lagrange /
    sll_p_spline of sll_s_periodic_interp)"""
    nml = tmp_path / "vpsim4d_cartesian_input.nml"
    nml.write_text(NML % {"dt": "0.1", "adv": adv})
    out = tmp_path / "thdiag.dat"
    sb.sim4d_run_namelist(str(tmp_path / "vpsim4d_cartesian_input"), str(out))     # extension appended like the reference
    lines = out.read_text().splitlines()
    assert len(lines) == 3 and all(len(l) == 13 * 20 for l in lines)
    rows = np.array([[float(l[20 * k:20 * k + 20]) for k in range(13)] for l in lines])
    nc = [16, 16, 32, 32]
    xmin, xmax = [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6]
    _, othd = orc.sim4d(nc, xmin, xmax, 0.5, 0.5, 1e-3, 0.1, 4, split="SLL_ORDER6VPnew1_VTV", method=1 if method == 0 else 2,
                        order=4, stencil=(-3, 3), want_thdiag=True)
    ref = othd[[0, 2, 4]]
    assert np.abs(rows[:, 0] - ref[:, 0]).max() < 1e-12
    assert np.abs(rows[:, [2, 7, 8, 9]] / ref[:, [2, 7, 8, 9]] - 1).max() < 1e-8
    assert np.abs(rows[:, 1] / ref[:, 1] - 1).max() < 1e-6 and np.abs(rows[:, 5] / ref[:, 5] - 1).max() < 1e-5
    assert np.abs(rows[:, [3, 4, 10, 11, 12]] / ref[:, [3, 4, 10, 11, 12]] - 1).max() < 1e-11


def test_c_driver_runs_the_namelist(sb, orc, tmp_path):
    """the plain-C counterpart of the reference executable (tests/c/): namelist in, thdiag.dat out, rows as sllb_sim4d_thdiag"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "selalib_b200", "lib")
    exe = str(tmp_path / "sim_bsl_vp_2d2v_cart_poisson_serial_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "c", "sim_bsl_vp_2d2v_cart_poisson_serial_b200.c"), "-o", exe,
                           "-L" + libdir, "-lsllb200", "-Wl,-rpath," + libdir])
    (tmp_path / "in.nml").write_text(NML % {"dt": "0.1", "adv": "SLL_SPLINES"})
    out = subprocess.run([exe, str(tmp_path / "in")], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = (tmp_path / "thdiag.dat").read_text().splitlines()
    assert len(lines) == 3 and all(len(l) == 260 for l in lines)
    rows = np.array([[float(l[20 * k:20 * k + 20]) for k in range(13)] for l in lines])
    S = sb.Sim4d([16, 16, 32, 32], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1, split="SLL_ORDER6VPnew1_VTV",
                 stencil=(-3, 3))
    S.run(4, diagnostics=False)
    ref = S.thdiag()
    S.destroy()
    assert (np.abs(rows[2] - ref) <= 1e-9 * np.abs(ref) + 1e-30).all(), (rows[2], ref)   # 12 printed digits; nrj_jac is 0
