"""Extracts the columns of the reference's shipped 1D1V trace that are independent of the (stale) quadrature conventions
of the run that produced it -- time, L2 norm, potential energy -- into tests/golden/vpsim2d_cartesian_ref_l2_epot.dat.

Source: simulations/parallel/bsl_vp_1d1v_cart/vpsim2d_cartesian_ref.dat (600 steps of vpsim2d_cartesian_input.nml:
Landau damping 32 x 64, eps = 1e-3, k = 0.5, dt = 0.1, cubic splines, Strang VTV).  Its mass / L1 columns differ from
today's code by the constant weight of the two velocity end points (1.1e-9 relative) and its momentum / kinetic-energy
columns follow an older normalisation, so they are not used as pins.
Run from the repo root in the build container: python tests/golden/make_vpsim2d_fixture.py"""
import os

import numpy as np

SRC = "/root/reference/simulations/parallel/bsl_vp_1d1v_cart/vpsim2d_cartesian_ref.dat"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vpsim2d_cartesian_ref_l2_epot.dat")
ref = np.loadtxt(SRC)
np.savetxt(DST, ref[:, [0, 4, 6]], fmt="%.12e", header="time l2norm potential_energy (columns 1, 5, 7 of vpsim2d_cartesian_ref.dat)")
print(DST, ref.shape)
