"""The oracle reproduces its committed fixtures (tests/golden/oracle_*.npz, made by tests/golden/make_golden.py):
guards the checker itself against silent changes (compiler flags, edits) on whatever box the tests run."""
import os

import numpy as np

from oracle import orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_matches_committed_advect_fixture():
    d = np.load(os.path.join(G, "oracle_advect4d.npz"))
    f0 = np.asfortranarray(d["f0"])
    n = 0
    for key in d.files:
        if "_axis" not in key:
            continue
        name, axis = key.rsplit("_axis", 1)
        axis = int(axis)
        method = name.rstrip("0123456789")
        order = int(name[len(method):])
        out = orc.advect_axis(f0.copy(order="F"), axis, method, order, d[f"disp{axis}"], tuple(int(v) for v in d[f"dsel{axis}"]))
        assert np.abs(out - d[key]).max() <= 1e-14 * np.abs(d[key]).max(), key
        n += 1
    assert n == 8


def test_oracle_matches_committed_line_and_poisson_fixtures():
    d = np.load(os.path.join(G, "oracle_lines.npz"))
    n = 64
    line = d["line"]
    assert np.allclose(orc.advect_1d_periodic_constant("spline", n, 0.0, 2 * np.pi, 4, 1.3, 0.1, line), d["adv_spline"], rtol=0, atol=1e-14)
    assert np.allclose(orc.advect_1d_periodic_constant("lagrange", n, 0.0, 2 * np.pi, 6, 1.3, 0.1, line), d["adv_lagrange6"], rtol=0, atol=1e-14)
    assert np.allclose(orc.spline_interpolate_array_disp(line, 0.0, 2 * np.pi, -1.2 * (2 * np.pi / n)), d["spline_disp"], rtol=0, atol=1e-14)
    p = np.load(os.path.join(G, "oracle_poisson.npz"))
    e1, e2 = orc.poisson_2d(np.asfortranarray(p["rho2"]), 16, 12, 0.0, 4 * np.pi, 0.0, 2 * np.pi)
    assert np.allclose(e1, p["e1"], rtol=0, atol=1e-13) and np.allclose(e2, p["e2"], rtol=0, atol=1e-13)
    assert np.allclose(orc.poisson_1d(p["rho1"], 0.0, 4 * np.pi), p["e1d"], rtol=0, atol=1e-13)


def test_oracle_matches_committed_traces():
    t = np.load(os.path.join(G, "oracle_traces.npz"))
    rows4 = orc.sim4d([16, 16, 32, 32], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1, 5)
    assert np.allclose(rows4, t["rows4"], rtol=1e-11, atol=1e-18)
    rows2 = orc.sim2d(64, 64, 0.0, 4 * np.pi, -6.0, 6.0, 0, 0.5, 1e-3, 0.1, 20)
    assert np.allclose(rows2, t["rows2"], rtol=1e-11, atol=1e-18)


def test_g4_1d1v_trace_l2_and_field_energy():
    """simulations/parallel/bsl_vp_1d1v_cart/vpsim2d_cartesian_ref.dat (600 steps of vpsim2d_cartesian_input.nml, 12
    printed digits): the L2 norm of f is reproduced over the whole run to the printed precision and the potential energy
    to 1e-8 of its maximum at every step (observed 2.2e-9, the printed precision of the smallest values).  Fixture: tests/golden/vpsim2d_cartesian_ref_l2_epot.dat (make_vpsim2d_fixture.py)."""
    gold = np.loadtxt(os.path.join(G, "vpsim2d_cartesian_ref_l2_epot.dat"))
    rows = orc.sim2d(32, 64, 0.0, 4 * np.pi, -6.0, 6.0, 0, 0.5, 0.001, 0.1, 600, method=0, order=4)
    assert np.abs(rows[:, 0] - gold[:, 0]).max() < 1e-9
    assert np.abs(rows[:, 4] / gold[:, 1] - 1).max() < 1e-11
    # the field energy oscillates through near-zero minima: compare on the scale of its maximum
    assert np.abs(rows[:, 6] - gold[:, 2]).max() < 1e-8 * gold[:, 2].max()
