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
