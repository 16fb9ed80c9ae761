/*
 * sll_oracle_hermite.c -- CPU restatement of SeLaLib's cubic splines with Hermite boundary conditions
 * (SURVEY.md section 8(f) rank 3).  TEST INFRASTRUCTURE ONLY, like sll_oracle.c.
 *
 * Restates src/splines/splines_basic/sll_m_cubic_splines.F90:
 *   FORWARD_FD_5PT / BACKWARD_FD_5PT (:176-181), compute_spline_1D_hermite (:692-748) with the fast algorithm
 *   compute_spline_1D_hermite_aux (:583-652, NUM_TERMS = 27) and the tridiagonal system of the LU path (:340-362,733-743),
 *   sll_s_cubic_spline_1d_eval_array (:903-954), sll_s_cubic_spline_1d_eval_disp (:2616-2682, Hermite branch),
 * and the two entry points of sll_t_cubic_spline_interpolator_1d that the simulations call
 * (src/interpolation/interpolators/sll_m_cubic_spline_interpolator_1d.F90:112-180):
 *   interpolate_array_disp (compute_interpolant + eval_disp) and interpolate_array_disp_inplace (clamped coordinates +
 *   eval_array), as used along v by simulations/parallel/bsl_vp_2d2v_cart/sll_m_sim_bsl_vp_2d2v_cart.F90:470-482,520-545.
 * Pinned by the reference's known-answer test src/splines/splines_basic/testing/test_cubic_splines.F90:57-138
 * (np = 32, f = exp(sin x), exact end slopes: grid values to 1e-14, mid-cell value to 2e-5).
 *
 * coeffs holds spline%coeffs(0:np+2): C[k] = coeffs(k).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define HERMITE_NUM_TERMS 27 /* sll_m_cubic_splines.F90:229 */

/* fast = 1: compute_spline_1D_hermite_aux; fast = 0: the (np+2) x (np+2) system the LU path factorises, solved by
 * Gaussian elimination without pivoting (diagonally dominant), i.e. the same solution to rounding.
 * slopes: have_sl / have_sr = 0 -> 5-point one-sided finite differences of the data (:720-730). */
void orc_spline_hermite_compute_interpolant(const double *f /* 1..np as f[0..np-1] */, int np, double xmin, double xmax,
                                            int fast, int have_sl, double sl, int have_sr, double sr, double *C) {
    const double delta = (xmax - xmin) / (double)(np - 1), r_delta = 1.0 / delta;
#define F(k) f[(k) - 1]
    double slope_l = have_sl ? sl
                             : r_delta * (-(25.0 / 12.0) * F(1) + 4.0 * F(2) - 3.0 * F(3) + (4.0 / 3.0) * F(4) - 0.25 * F(5));
    double slope_r = have_sr ? sr
                             : r_delta * (0.25 * F(np - 4) - (4.0 / 3.0) * F(np - 3) + 3.0 * F(np - 2) - 4.0 * F(np - 1) +
                                          (25.0 / 12.0) * F(np));
    if (np < HERMITE_NUM_TERMS) fast = 0; /* :266-268 */
    if (!fast) {
        /* a(1..3) = (0, 4/6, 2/6); interior (1/6, 4/6, 1/6); last (2/6, 4/6, 0); rhs f_aux (:340-352,733-737) */
        int n = np + 2;
        double *lo = (double *)malloc(sizeof(double) * 4 * n), *di = lo + n, *up = di + n, *rhs = up + n;
        for (int i = 0; i < n; ++i) { lo[i] = 1.0 / 6.0; di[i] = 4.0 / 6.0; up[i] = 1.0 / 6.0; }
        lo[0] = 0.0; up[0] = 2.0 / 6.0; lo[n - 1] = 2.0 / 6.0; up[n - 1] = 0.0;
        for (int i = 1; i <= np; ++i) rhs[i] = F(i);
        rhs[0] = F(1) + (1.0 / 3.0) * delta * slope_l;
        rhs[n - 1] = F(np) - (1.0 / 3.0) * delta * slope_r;
        for (int i = 1; i < n; ++i) { double m = lo[i] / di[i - 1]; di[i] -= m * up[i - 1]; rhs[i] -= m * rhs[i - 1]; }
        C[n - 1] = rhs[n - 1] / di[n - 1];
        for (int i = n - 2; i >= 0; --i) C[i] = (rhs[i] - up[i] * C[i + 1]) / di[i];
        C[np + 2] = 0.0;
        free(lo);
        return;
    }
    /* compute_spline_1D_hermite_aux: the dummy `coeffs` is 1-based there, coeffs_aux(k) = C[k-1] */
    const double a = sqrt((2.0 + sqrt(3.0)) / 6.0), r_a = 1.0 / a, b = sqrt((2.0 - sqrt(3.0)) / 6.0), b_a = b / a;
    const double ralpha = sqrt(6.0 / sqrt(3.0));
    double *d = (double *)malloc(sizeof(double) * (np + 1));
#define CA(k) C[(k) - 1]
    double f1 = F(1), fnp = F(np) - delta * slope_r / 3.0;
    double d1 = f1, coeff_tmp = 1.0;
    for (int i = 2; i <= HERMITE_NUM_TERMS; ++i) {
        coeff_tmp = coeff_tmp * (-b_a);
        d1 = d1 + coeff_tmp * (F(i) - 2.0 * slope_l * delta * (double)(i - 1));
    }
    d[1] = d1 * r_a;
    for (int i = 2; i <= np - 1; ++i) d[i] = r_a * (F(i) - b * d[i - 1]);
    d[np] = ralpha * (0.5 * fnp - b * d[np - 1]);
    CA(np + 1) = ralpha * d[np];
    for (int i = np - 1; i >= 1; --i) CA(i + 1) = r_a * (d[i] - b * CA(i + 2));
    CA(1) = CA(3) - 2.0 * delta * slope_l;
    CA(np + 2) = CA(np) + 2.0 * delta * slope_r;
    CA(np + 3) = 0.0;
#undef CA
#undef F
    free(d);
}

static double cell_dx(const double *C, int cell, double dx) { /* spline_interpolate_from_interpolant_cell_dx :2685-2711 */
    double cdx = 1.0 - dx, cim1 = C[cell - 1], ci = C[cell], cip1 = C[cell + 1], cip2 = C[cell + 2];
    double t1 = 3.0 * ci, t3 = 3.0 * cip1;
    double t2 = cdx * (cdx * (cdx * (cim1 - t1) + t1) + t1) + ci;
    double t4 = dx * (dx * (dx * (cip2 - t3) + t3) + t3) + cip1;
    return (1.0 / 6.0) * (t2 + t4);
}

/* sll_s_cubic_spline_1d_eval_disp, Hermite branch (:2634-2680): num_cells = n_points */
void orc_spline_hermite_eval_disp(const double *C, int np, double xmin, double xmax, double alpha, double *out) {
    const double rdelta = 1.0 / ((xmax - xmin) / (double)(np - 1));
    double alpha0 = alpha * rdelta;
    int dcell = (int)floor(alpha0);
    double alpha1 = alpha0 - (double)dcell;
    int num_cells = np;
    if (dcell >= np || dcell <= -np) return; /* the reference's fill loops would run past the array: nothing to restate */
    int lo = (1 > 1 - dcell) ? 1 : 1 - dcell, hi = (num_cells < num_cells - dcell) ? num_cells : num_cells - dcell;
    for (int i = lo; i <= hi; ++i) out[i - 1] = cell_dx(C, i + dcell, alpha1);
    alpha1 = 0.0;
    if (dcell < 0) {
        out[0] = cell_dx(C, 1, alpha1);
        for (int i = 2; i <= -dcell; ++i) out[i - 1] = out[0];
    } else {
        out[np - 1] = cell_dx(C, np, alpha1);
        for (int i = num_cells - dcell + 1; i <= np - 1; ++i) out[i - 1] = out[np - 1];
    }
}

/* sll_s_cubic_spline_1d_eval_array (:903-954) */
void orc_spline_eval_array(const double *C, int np, double xmin, double xmax, const double *x, int n, double *out) {
    const double rh = 1.0 / ((xmax - xmin) / (double)(np - 1));
    for (int i = 0; i < n; ++i) {
        double t0 = (x[i] - xmin) * rh;
        int cell = (int)t0 + 1;
        double dx = t0 - (double)(cell - 1);
        out[i] = cell_dx(C, cell, dx);
    }
}

/* interpolate_array_disp (:112-126) with sll_p_hermite */
void orc_hermite_interpolate_array_disp(int np, double xmin, double xmax, int fast, int have_sl, double sl, int have_sr,
                                        double sr, const double *data, double alpha, double *out) {
    double *C = (double *)malloc(sizeof(double) * (np + 4));
    orc_spline_hermite_compute_interpolant(data, np, xmin, xmax, fast, have_sl, sl, have_sr, sr, C);
    orc_spline_hermite_eval_disp(C, np, xmin, xmax, alpha, out);
    free(C);
}
/* interpolate_array_disp_inplace (:128-180) with a non-periodic boundary: coordinates clamped to [xmin, xmax] */
void orc_hermite_interpolate_array_disp_inplace(int np, double xmin, double xmax, int fast, int have_sl, double sl,
                                                int have_sr, double sr, double *data, double alpha) {
    double *C = (double *)malloc(sizeof(double) * (2 * np + 8));
    double *coords = C + np + 4;
    orc_spline_hermite_compute_interpolant(data, np, xmin, xmax, fast, have_sl, sl, have_sr, sr, C);
    /* interpolation_points: accumulated, the last one set to xmax (:341-347) */
    const double delta = (xmax - xmin) / (double)(np - 1);
    double p = xmin;
    for (int i = 0; i < np; ++i) {
        if (i > 0) p = p + delta;
        if (i == np - 1) p = xmax;
        coords[i] = (alpha < 0) ? fmax(p + alpha, xmin) : fmin(p + alpha, xmax);
    }
    orc_spline_eval_array(C, np, xmin, xmax, coords, np, data);
    free(C);
}

/* whole-array pass along one axis of f viewed as [outer][n][inner]: interpolate_array_disp_inplace on every line with
 * alpha = dvals[idx] (physical units), slopes from finite differences, as the V stages of bsl_vp_2d2v_cart do */
void orc_hermite_advect_axis(double *f, long outer, int n, long inner, double xmin, double xmax, int fast, int inplace,
                             const double *dvals, long odiv, long omod, long ostr, long idiv, long imodn, long istr) {
#pragma omp parallel
    {
        double *lin = (double *)malloc(sizeof(double) * 2 * (size_t)n);
        double *lout = lin + n;
#pragma omp for schedule(static) collapse(2)
        for (long o = 0; o < outer; ++o)
            for (long in = 0; in < inner; ++in) {
                double *base = f + o * (long)n * inner + in;
                double alpha = dvals[((o / odiv) % omod) * ostr + ((in / idiv) % imodn) * istr];
                for (int i = 0; i < n; ++i) lin[i] = base[(long)i * inner];
                if (inplace) { orc_hermite_interpolate_array_disp_inplace(n, xmin, xmax, fast, 0, 0.0, 0, 0.0, lin, alpha);
                               for (int i = 0; i < n; ++i) base[(long)i * inner] = lin[i]; }
                else { orc_hermite_interpolate_array_disp(n, xmin, xmax, fast, 0, 0.0, 0, 0.0, lin, alpha, lout);
                       for (int i = 0; i < n; ++i) base[(long)i * inner] = lout[i]; }
            }
        free(lin);
    }
}
