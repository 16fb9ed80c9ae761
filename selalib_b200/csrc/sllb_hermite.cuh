// Per-line arithmetic of K10 (cubic spline with Hermite boundary conditions).  Host-compilable like sllb_spline15.cuh
// (tests/host/spline15_host.cpp builds it with g++ for the CPU suite; the product only runs it on the device).
//
// Reference: compute_spline_1D_hermite + compute_spline_1D_hermite_aux (src/splines/splines_basic/
// sll_m_cubic_splines.F90:583-652,692-748, fast algorithm, NUM_TERMS = 27), evaluation by
// sll_s_cubic_spline_1d_eval_disp (:2616-2682, Hermite branch) or, for the in-place interpolator call, clamped
// coordinates + sll_s_cubic_spline_1d_eval_array (sll_m_cubic_spline_interpolator_1d.F90:128-180, :903-954).
// With a = sqrt((2+sqrt3)/6), q = b/a = 2 - sqrt3 the recurrences run on e = a d and G = a^2 c (one FMA per point):
//     e(1)  = sum_{i=1..27} (-q)^(i-1) (f(i) - 2 slope_l delta (i-1))        (:624-632, reflection about x_1)
//     e(i)  = f(i) - q e(i-1),  i = 2..np-1                                   (:634-636)
//     G(np) = (1 + 2/sqrt3) (f(np)/2 - delta slope_r/6 - q e(np-1))           (:618,637,640)
//     G(i)  = e(i) - q G(i+1),  i = np-1..1                                   (:641-643)
//     G(0)  = G(2) - 2 a^2 delta slope_l,  G(np+1) = G(np-1) + 2 a^2 delta slope_r,  G(np+2) = 0   (:644-646)
#pragma once
#include <math.h>
#include "sllb_spline15.cuh"

namespace sllb {

#define SLLB_HERMITE_TERMS 27
SLLB_CONST double c_hq[SLLB_HERMITE_TERMS]; // (-q)^i, i = 0..26

// x0[k*PITCH], k = 0..np-1: data on entry, G(1..np) on exit; returns G(0), G(np+1)
template <int PITCH>
SLLB_DEV void hermite_coeffs_line(double *x0, const int np, const double delta, const int have_slopes, const double sl_in,
                                  const double sr_in, double *g0, double *gnp1) {
    const double q = 0.26794919243112270647;   // 2 - sqrt(3)
    const double a2 = 0.62200846792814621559;  // (2 + sqrt 3)/6
    const double kn = 2.15470053837925152902;  // a^2 * 6/sqrt(3) = 1 + 2/sqrt(3)
    const double rd = 1.0 / delta;
    double sl = sl_in, sr = sr_in;
    if (!have_slopes) { // FORWARD_FD_5PT / BACKWARD_FD_5PT (:176-181)
        sl = rd * (-(25.0 / 12.0) * x0[0] + 4.0 * x0[PITCH] - 3.0 * x0[2 * PITCH] + (4.0 / 3.0) * x0[3 * PITCH] - 0.25 * x0[4 * PITCH]);
        sr = rd * (0.25 * x0[(np - 5) * PITCH] - (4.0 / 3.0) * x0[(np - 4) * PITCH] + 3.0 * x0[(np - 3) * PITCH] -
                   4.0 * x0[(np - 2) * PITCH] + (25.0 / 12.0) * x0[(np - 1) * PITCH]);
    }
    const double fnp = x0[(np - 1) * PITCH] - delta * sr / 3.0;
    double e = x0[0];
    const double tilt = 2.0 * sl * delta;
#pragma unroll
    for (int i = 1; i < SLLB_HERMITE_TERMS; ++i) e = fma(c_hq[i], x0[i * PITCH] - tilt * (double)i, e);
    x0[0] = e;
#pragma unroll 4
    for (int k = 1; k < np - 1; ++k) {
        e = fma(-q, e, x0[k * PITCH]);
        x0[k * PITCH] = e;
    }
    double g = kn * fma(-q, e, 0.5 * fnp); // G(np)
    x0[(np - 1) * PITCH] = g;
    double gm1 = 0.0;                      // G(np-1) after the first step
#pragma unroll 4
    for (int k = np - 2; k >= 0; --k) {
        g = fma(-q, g, x0[k * PITCH]);
        x0[k * PITCH] = g;
        if (k == np - 2) gm1 = g;
    }
    *g0 = x0[PITCH] - 2.0 * a2 * delta * sl;     // G(0) = G(2) - ...
    *gnp1 = gm1 + 2.0 * a2 * delta * sr;         // G(np+1) = G(np-1) + ...
}

// G(k), k = 0..np+2, from the line in shared memory and the two ghost values
template <int PITCH>
SLLB_DEV double hermite_G(const double *x0, const int np, const int k, const double g0, const double gnp1) {
    if (k >= 1 && k <= np) return x0[(k - 1) * PITCH];
    if (k == 0) return g0;
    if (k == np + 1) return gnp1;
    return 0.0; // coeffs(np+2) "not used" (:646)
}

// value at output point i (1-based).  inplace = 1: foot clamped to [xmin, xmax], cell = int(t0) + 1 (eval_array);
// inplace = 0: eval_disp's index ranges, including its evaluation of the ghost cell np at i = np - dcell.
template <int PITCH>
SLLB_DEV double hermite_eval_point(const double *x0, const int np, const int i, const double alpha0, const int inplace,
                                   const double g0, const double gnp1) {
    int cell; double dx;
    if (inplace) {
        double t0 = (double)(i - 1) + alpha0;
        t0 = t0 < 0.0 ? 0.0 : (t0 > (double)(np - 1) ? (double)(np - 1) : t0);
        cell = (int)t0 + 1;
        dx = t0 - (double)(cell - 1);
    } else {
        const double fl = floor(alpha0);
        const int dcell = (int)fl;
        if (i + dcell < 1) { cell = 1; dx = 0.0; }
        else if (i + dcell > np) { cell = np; dx = 0.0; }
        else { cell = i + dcell; dx = (dcell >= 0 && i == np) ? 0.0 : alpha0 - fl; }
    }
    const Weights4 w = spline_weights(dx);
    const double cm = hermite_G<PITCH>(x0, np, cell - 1, g0, gnp1), c0 = hermite_G<PITCH>(x0, np, cell, g0, gnp1);
    const double c1 = hermite_G<PITCH>(x0, np, cell + 1, g0, gnp1), c2 = hermite_G<PITCH>(x0, np, cell + 2, g0, gnp1);
    return fma(w.w3, c2, fma(w.w2, c1, fma(w.w1, c0, w.w0 * cm)));
}

} // namespace sllb
