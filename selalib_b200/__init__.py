"""selalib_b200 -- B200-native split semi-Lagrangian advection path of SeLaLib.

The product is the C-ABI shared library ``selalib_b200/lib/libsllb200.so`` (CUDA, sm_100a;
headers in ``include/``).  This package is a thin ctypes binding used by the tests and by
bench.py; it never computes anything itself and there is no CPU fallback: importing
:mod:`selalib_b200.capi` raises if the library has not been built.
"""
from .capi import (  # noqa: F401
    SllbError, lib, last_error, init, device_count, synchronize, launch_count, launch_count_reset, set_staging, set_cuda_graphs, set_phase_timers, set_v_overlap, set_poisson_direct, set_spline_split, set_fused_remap, set_remap_rotation, set_plane_kernel, set_plane_variant,
    Advector1dPeriodic, Interpolator1d, Field, Poisson, Comm, Dist4d, Dd6d, dd6d_plan, dd6d_set_force_halo, dd6d_set_halo_p2p, dd6d_set_exchange_timing, dd6d_set_halo_chunks, Sim4d, Sim2d, Sim6d,
    factorize_in_two_powers_of_two, layout4d_boxes, remap4d_plan, set_process_grid, spline_dd_blocks, lagrange_dd_blocks, splitting_case, splitting_coeff, compute_w_hermite, format_g20_12, sim4d_run_namelist, sim2d_run_namelist, format_g,
    BC_PERIODIC, BC_HERMITE, ADVECTOR_FIXED, ADVECTOR_CENTERED, ADVECTOR_SPLINE, SHIFT_SKIP, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE,
    METHOD_SPLINE, METHOD_LAGRANGE_FIXED, METHOD_LAGRANGE_CENTERED,
    ADV_PERIODIC_SPLINE, ADV_PERIODIC_LAGRANGE, ADV_BSL,
    INTERP_CUBIC_SPLINE, INTERP_LAGRANGE_CENTERED, INTERP_LAGRANGE_FIXED, INTERP_PERIODIC_SPLINE,
    INTERP_PERIODIC_LAGRANGE,
)
