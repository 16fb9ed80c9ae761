/*
 * sll_oracle.c -- CPU restatement of SeLaLib's split semi-Lagrangian advection path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 *
 * The reference (Fortran, /root/reference) cannot be compiled in this image
 * (no Fortran compiler, no MPI, no FFTW), so every routine below restates the
 * arithmetic of the cited reference lines in plain C, same operation order,
 * fp64.  Parity pins (tests/test_oracle_*.py):
 *   - G1 golden file reffile_bsl_vp_3d3v_cart_dd.dat (3 rows x 14 columns,
 *     tolerance 5e-7 as in sll_m_sim_6d_utilities.F90:663) via orc_sim6d_run;
 *   - the analytic known-answer thresholds of the reference's own unit tests
 *     (Poisson 1e-14/1e-13, Lagrange fast 1e-8/8e-6/7e-9/3e-7/2e-10, spline 1e-6).
 *
 * All arrays are column-major (Fortran order); indices in comments are the
 * reference's 1-based ones, C code is 0-based.
 * Paths cited are relative to /root/reference.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846264338327950288
#define ORC_TWOPI (2.0 * ORC_PI)
typedef double complex cplx;

static inline int imod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm of bench.py asks for the host's cores explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* FFT helpers (unnormalised DFT, sign = -1 forward / +1 backward).           */
/* Semantics of FFTPACK zfftf/zfftb (external/fftpack) and FFTW c2c: a DFT is  */
/* unique, so any exact algorithm agrees to rounding.                          */
/* ------------------------------------------------------------------------- */
/* forward twiddles exp(-2 pi i k / n), k < n/2, cached per n (FFTPACK's wsave plays this role) */
#define TW_SLOTS 8
static struct { int n; cplx *w; } tw_cache[TW_SLOTS];
static const cplx *twiddles(int n) {
    for (int i = 0; i < TW_SLOTS; ++i) if (tw_cache[i].n == n) return tw_cache[i].w;
    const cplx *res = NULL;
#pragma omp critical(orc_tw)
    {
        int slot = -1;
        for (int i = 0; i < TW_SLOTS; ++i) { if (tw_cache[i].n == n) { res = tw_cache[i].w; break; } if (tw_cache[i].n == 0 && slot < 0) slot = i; }
        if (!res && slot >= 0) {
            cplx *w = (cplx *)malloc(sizeof(cplx) * (n / 2 + 1));
            for (int k = 0; k < n / 2; ++k) { double a = -ORC_TWOPI * k / n; w[k] = cos(a) + I * sin(a); }
            tw_cache[slot].w = w;
#pragma omp flush
            tw_cache[slot].n = n;
            res = w;
        }
    }
    return res;
}

static void fft_inplace(cplx *x, int n, int sign) {
    if (n <= 1) return;
    if ((n & (n - 1)) == 0) {
        /* iterative radix-2 */
        const cplx *tw = twiddles(n);
        for (int i = 1, j = 0; i < n; ++i) {
            int bit = n >> 1;
            for (; j & bit; bit >>= 1) j ^= bit;
            j ^= bit;
            if (i < j) { cplx t = x[i]; x[i] = x[j]; x[j] = t; }
        }
        for (int len = 2; len <= n; len <<= 1) {
            int half = len >> 1, step = n / len;
            for (int i = 0; i < n; i += len) {
                for (int k = 0; k < half; ++k) {
                    cplx w;
                    if (tw) { w = tw[k * step]; if (sign > 0) w = conj(w); }
                    else { double ang = sign * ORC_TWOPI * k / len; w = cos(ang) + I * sin(ang); }
                    cplx u = x[i + k], v = x[i + k + half] * w;
                    x[i + k] = u + v;
                    x[i + k + half] = u - v;
                }
            }
        }
    } else {
        cplx *y = (cplx *)malloc(sizeof(cplx) * n);
        for (int k = 0; k < n; ++k) {
            cplx s = 0;
            for (int j = 0; j < n; ++j) {
                int m = (int)(((long long)j * k) % n);
                double a = sign * ORC_TWOPI * m / n;
                s += x[j] * (cos(a) + I * sin(a));
            }
            y[k] = s;
        }
        memcpy(x, y, sizeof(cplx) * n);
        free(y);
    }
}

/* strided complex FFT of `count` lines */
static void fft_lines(cplx *a, int n, long stride, int sign, cplx *work) {
    for (int i = 0; i < n; ++i) work[i] = a[i * stride];
    fft_inplace(work, n, sign);
    for (int i = 0; i < n; ++i) a[i * stride] = work[i];
}

/* FFTPACK dfftf: real forward, half-complex output r(1)=DC, r(2k)=Re, r(2k+1)=Im,
 * r(n)=Nyquist (n even).  external/fftpack/dfftf.f semantics. */
static void dfftf_(int n, double *r) {
    cplx *x = (cplx *)malloc(sizeof(cplx) * n);
    for (int i = 0; i < n; ++i) x[i] = r[i];
    fft_inplace(x, n, -1);
    r[0] = creal(x[0]);
    for (int k = 1; k <= (n - 1) / 2; ++k) { r[2 * k - 1] = creal(x[k]); r[2 * k] = cimag(x[k]); }
    if (n % 2 == 0) r[n - 1] = creal(x[n / 2]);
    free(x);
}
/* FFTPACK dfftb: unnormalised inverse of dfftf. */
static void dfftb_(int n, double *r) {
    cplx *x = (cplx *)malloc(sizeof(cplx) * n);
    x[0] = r[0];
    for (int k = 1; k <= (n - 1) / 2; ++k) {
        x[k] = r[2 * k - 1] + I * r[2 * k];
        x[n - k] = r[2 * k - 1] - I * r[2 * k];
    }
    if (n % 2 == 0) x[n / 2] = r[n - 1];
    fft_inplace(x, n, +1);
    for (int i = 0; i < n; ++i) r[i] = creal(x[i]);
    free(x);
}

/* ------------------------------------------------------------------------- */
/* a6: uniform periodic cubic splines                                          */
/* ------------------------------------------------------------------------- */
#define NUM_TERMS 27 /* src/splines/splines_basic/sll_m_cubic_splines.F90:229 */

/* compute_spline_1D_periodic_aux, sll_m_cubic_splines.F90:531-581.
 * f: np-1 (or more) values; coeffs: np+3 values = spline%coeffs(0:np+2). */
static void spline_periodic_fast(const double *f, int np, double *d, double *C) {
    const double a = sqrt((2.0 + sqrt(3.0)) / 6.0);
    const double r_a = 1.0 / a;
    const double b = sqrt((2.0 - sqrt(3.0)) / 6.0);
    const double b_a = b / a;
    const int N = np - 1;
    double d1 = f[0], ct = 1.0;
    for (int i = 0; i <= NUM_TERMS - 1; ++i) { ct *= (-b_a); d1 += ct * f[N - 1 - i]; }
    d[0] = d1 * r_a;
    for (int i = 1; i < N; ++i) d[i] = r_a * (f[i] - b * d[i - 1]);
    d1 = d[N - 1]; ct = 1.0;
    for (int i = 1; i <= NUM_TERMS; ++i) { ct *= (-b_a); d1 += ct * d[i - 1]; }
    C[N] = d1 * r_a;
    for (int i = N - 1; i >= 1; --i) C[i] = r_a * (d[i - 1] - b * C[i + 1]);
    C[0] = C[N]; C[N + 1] = C[1]; C[N + 2] = C[2]; C[N + 3] = C[3];
}

/* LU path (num_points < 27 or fast_algorithm=.false.): cyclic tridiagonal system with
 * rows (1/6, 4/6, 1/6) of size N, sll_m_cubic_splines.F90:296-318,669-681, solved by
 * sll_m_tridiagonal.F90:161-594.  The reference LU uses partial pivoting; for this
 * strictly diagonally dominant matrix the pivot test (:231, |s11|<|s21| or |s31|) never
 * fires, so the factorisation reduces to the plain cyclic elimination restated here. */
static void spline_periodic_lu(const double *f, int np, double *C) {
    const int n = np - 1;
    const double lo = 1.0 / 6.0, di = 4.0 / 6.0, up = 1.0 / 6.0;
    double *dd = (double *)malloc(sizeof(double) * n * 4);
    double *l = dd + n, *q = dd + 2 * n, *y = dd + 3 * n;
    double *x = C + 1;
    if (n == 1) { x[0] = f[0]; }
    else if (n == 2) { /* rows: [di, lo+up; lo+up, di] */
        double o = lo + up, det = di * di - o * o;
        x[0] = (di * f[0] - o * f[1]) / det; x[1] = (di * f[1] - o * f[0]) / det;
    } else {
        /* Elimination of the sub-diagonal; q = fill-in of the last column,
         * m (carried in `mrow`) = fill-in of the last row. */
        double mrow = up;     /* A(n,1) */
        double dlast = di;    /* A(n,n) */
        double rhs_last = f[n - 1];
        dd[0] = di; q[0] = lo; /* A(1,n) */
        y[0] = f[0];
        for (int i = 0; i < n - 2; ++i) {
            l[i] = lo / dd[i];
            dd[i + 1] = di - l[i] * up;
            q[i + 1] = -l[i] * q[i];
            y[i + 1] = f[i + 1] - l[i] * y[i];
            double m = mrow / dd[i];
            dlast -= m * q[i];
            rhs_last -= m * y[i];
            mrow = -m * up;
        }
        /* row n-1 (index n-2): its super-diagonal coincides with last column */
        q[n - 2] += up;
        {
            double m = (mrow + lo) / dd[n - 2];
            dlast -= m * q[n - 2];
            rhs_last -= m * y[n - 2];
        }
        x[n - 1] = rhs_last / dlast;
        x[n - 2] = (y[n - 2] - q[n - 2] * x[n - 1]) / dd[n - 2];
        for (int i = n - 3; i >= 0; --i) x[i] = (y[i] - up * x[i + 1] - q[i] * x[n - 1]) / dd[i];
    }
    C[0] = C[n]; C[n + 1] = C[1]; C[n + 2] = C[2]; C[n + 3] = C[3];
    free(dd);
}

/* sll_s_cubic_spline_1d_compute_interpolant (periodic), :498-518,654-690.
 * fast<0: reference default (fast iff num_points >= 27, :267-274). */
void orc_spline_compute_interpolant_periodic(const double *f, int np, int fast, double *coeffs) {
    int use_fast = (np < NUM_TERMS) ? 0 : (fast < 0 ? 1 : fast);
    if (use_fast) {
        double *d = (double *)malloc(sizeof(double) * np);
        spline_periodic_fast(f, np, d, coeffs);
        free(d);
    } else {
        spline_periodic_lu(f, np, coeffs);
    }
}

/* spline_interpolate_from_interpolant_cell_dx, :2685-2711 (cell is 1-based) */
static inline double spline_cell_dx(const double *C, int cell, double dx) {
    const double inv_6 = 1.0 / 6.0;
    double cdx = 1.0 - dx;
    double cim1 = C[cell - 1], ci = C[cell], cip1 = C[cell + 1], cip2 = C[cell + 2];
    double t1 = 3.0 * ci, t3 = 3.0 * cip1;
    double t2 = cdx * (cdx * (cdx * (cim1 - t1) + t1) + t1) + ci;
    double t4 = dx * (dx * (dx * (cip2 - t3) + t3) + t3) + cip1;
    return inv_6 * (t2 + t4);
}

/* sll_s_cubic_spline_1d_eval_disp (periodic), :2616-2682. out has np entries. */
void orc_spline_eval_disp_periodic(const double *C, int np, double xmin, double xmax,
                                   double alpha, double *out) {
    double delta = (xmax - xmin) / (double)(np - 1);
    double rdelta = 1.0 / delta;
    double alpha0 = alpha * rdelta;
    int dcell = (int)floor(alpha0);
    double alpha1 = alpha0 - (double)dcell;
    int N = np - 1;
    for (int i = 1; i <= N; ++i) {
        int cell = imod(i + dcell - 1, N) + 1;
        out[i - 1] = spline_cell_dx(C, cell, alpha1);
    }
    out[np - 1] = out[0];
}

/* sll_s_cubic_spline_1d_eval_array, :903-954 */
static void spline_eval_array(const double *C, int np, double xmin, double xmax,
                              const double *a_in, double *a_out, int n) {
    double delta = (xmax - xmin) / (double)(np - 1);
    double rh = 1.0 / delta;
    for (int i = 0; i < n; ++i) {
        double x = a_in[i];
        double t0 = (x - xmin) * rh;
        int cell = (int)t0 + 1;
        double dx = t0 - (double)(cell - 1);
        double cdx = 1.0 - dx;
        double cim1 = C[cell - 1], ci = C[cell], cip1 = C[cell + 1], cip2 = C[cell + 2];
        double t1 = 3.0 * ci, t3 = 3.0 * cip1;
        double t2 = cdx * (cdx * (cdx * (cim1 - t1) + t1) + t1) + ci;
        double t4 = dx * (dx * (dx * (cip2 - t3) + t3) + t3) + cip1;
        a_out[i] = (1.0 / 6.0) * (t2 + t4);
    }
}

/* a5: sll_t_cubic_spline_interpolator_1d%interpolate_array_disp,
 * src/interpolation/interpolators/sll_m_cubic_spline_interpolator_1d.F90:112-126.
 * data/out: np values (periodic duplicate last). out(i) = S_f(x_i + alpha). */
void orc_cubic_spline_interpolate_array_disp(int np, double xmin, double xmax, int fast,
                                             const double *data, double alpha, double *out) {
    double *C = (double *)malloc(sizeof(double) * (np + 3));
    orc_spline_compute_interpolant_periodic(data, np, fast, C);
    orc_spline_eval_disp_periodic(C, np, xmin, xmax, alpha, out);
    free(C);
}

/* ...%interpolate_array_disp_inplace (periodic), same file :128-180; interpolation
 * points by cumulative addition as in init (:339-346). */
void orc_cubic_spline_interpolate_array_disp_inplace(int np, double xmin, double xmax, int fast,
                                                     double *data, double alpha) {
    double *C = (double *)malloc(sizeof(double) * (np + 3));
    double *pts = (double *)malloc(sizeof(double) * np);
    double *coords = (double *)malloc(sizeof(double) * np);
    orc_spline_compute_interpolant_periodic(data, np, fast, C);
    double delta = (xmax - xmin) / (np - 1);
    pts[0] = xmin;
    for (int i = 1; i < np; ++i) pts[i] = pts[i - 1] + delta;
    pts[np - 1] = xmax;
    double length = pts[np - 1] - pts[0];
    if (alpha == 0.0) {
        for (int i = 0; i < np; ++i) coords[i] = pts[i];
    } else {
        for (int i = 0; i < np; ++i) {
            double a = pts[i] - xmin + alpha;
            double m = a - floor(a / length) * length; /* Fortran modulo(a, length), length>0 */
            coords[i] = xmin + m;
        }
    }
    spline_eval_array(C, np, xmin, xmax, coords, data, np);
    free(C); free(pts); free(coords);
}

/* ------------------------------------------------------------------------- */
/* a3: sll_s_periodic_interp (FFT-diagonalised periodic interpolation)         */
/* ------------------------------------------------------------------------- */
/* sll_s_uniform_bsplines_eval_basis, src/splines/splines_basic/sll_m_low_level_bsplines.F90:880-915 */
static void uniform_bsplines_eval_basis(int degree, double off, double *bspl) {
    bspl[0] = 1.0;
    for (int j = 1; j <= degree; ++j) {
        double xx = -off, j_real = (double)j, inv_j = 1.0 / j_real, saved = 0.0;
        for (int r = 0; r <= j - 1; ++r) {
            xx += 1.0;
            double temp = bspl[r] * inv_j;
            bspl[r] = saved + xx * temp;
            saved = (j_real - xx) * temp;
        }
        bspl[j] = saved;
    }
}

/* sll_p_spline branch, src/interpolation/periodic_interpolation/sll_m_periodic_interp.F90:64-119,202-229.
 * u_out(j) = interpolant(j - alpha), alpha in cells. */
void orc_periodic_interp_spline(int N, int order, const double *u, double alpha, double *u_out) {
    int p = order - 1;
    double biatx0[16], biatx[16];
    cplx *ufft = (cplx *)malloc(sizeof(cplx) * N);
    cplx *modes = (cplx *)malloc(sizeof(cplx) * N);
    uniform_bsplines_eval_basis(p, 0.0, biatx0);
    for (int i = 0; i < N; ++i) ufft[i] = u[i];
    fft_inplace(ufft, N, -1);
    int ishift = (int)floor(-alpha);
    double beta = -ishift - alpha;
    uniform_bsplines_eval_basis(p, beta, biatx);
    for (int i = 1; i <= N; ++i) {
        double a = ORC_TWOPI * (i - 1) / N;
        modes[i - 1] = cos(a) + I * sin(a);
    }
    for (int i = 1; i <= N; ++i) {
        double minv = biatx0[(p + 1) / 2 - 1];
        for (int j = 1; j <= (p + 1) / 2; ++j)
            minv += biatx0[j + (p + 1) / 2 - 1] * 2 * cos(j * ORC_TWOPI * (i - 1) / N);
        minv = 1.0 / minv;
        cplx es = 0;
        for (int j = -(p - 1) / 2; j <= (p + 1) / 2; ++j) {
            int imode = imod((ishift + j) * (i - 1), N);
            es += biatx[j + (p + 1) / 2 - 1] * modes[imode];
        }
        ufft[i - 1] = ufft[i - 1] * es * minv;
    }
    fft_inplace(ufft, N, +1);
    for (int i = 0; i < N; ++i) u_out[i] = creal(ufft[i]) / (double)N;
    free(ufft); free(modes);
}

/* fourier1dperlagodd, sll_m_periodic_interp.F90:290-366 (called with alpha/N, d=order/2-1) */
void orc_periodic_interp_lagrange(int N, int order, const double *u, double alpha_cells, double *E) {
    int d = order / 2 - 1;
    double *buf = (double *)calloc((size_t)N, sizeof(double));
    for (int i = 0; i < N; ++i) E[i] = u[i];
    double x = alpha_cells / (double)N;
    x = x - floor(x);
    x = x * (double)N;
    int ix = (int)floor(x);
    if (ix == N) { x = 0.0; ix = 0; }
    x = x - (double)ix;
    double a = 1.0;
    for (int i = 2; i <= d; ++i) a = a * (x * x - (double)i * (double)i) / ((double)d * (double)d);
    a = a * (x + 1.0) / (double)d;
    a = a * (x - (double)d - 1.0) / (double)d;
    buf[ix] = a * (x - 1.0) / (double)d;
    buf[(ix + 1) % N] = a * x / (double)d;
    a = a * x * (x - 1.0) / ((double)d * (double)d);
    for (int i = -d; i <= -1; ++i) buf[(i + ix + N) % N] = a / ((x - (double)i) / (double)d);
    for (int i = 2; i <= d + 1; ++i) buf[(i + ix + N) % N] = a / ((x - (double)i) / (double)d);
    a = 1.0;
    for (int i = -d; i <= d + 1; ++i) {
        buf[(i + ix + N) % N] *= a;
        a = a * (double)d / (double)(d + i + 1);
    }
    a = 1.0;
    for (int i = d + 1; i >= -d; --i) {
        buf[(i + ix + N) % N] *= a;
        a = a * (double)d / (double)(i - 1 - d - 1);
    }
    dfftf_(N, buf);
    dfftf_(N, E);
    double tmp = 1.0 / (double)N;
    E[0] = E[0] * tmp * buf[0];
    for (int i = 1; i <= (N - 2) / 2; ++i) {
        double rea = E[2 * i - 1], ima = E[2 * i];
        double reb = tmp * buf[2 * i - 1], imb = tmp * buf[2 * i];
        E[2 * i - 1] = rea * reb - ima * imb;
        E[2 * i] = rea * imb + reb * ima;
    }
    if (N % 2 == 0) E[N - 1] = E[N - 1] * tmp * buf[N - 1];
    dfftb_(N, E);
    free(buf);
}

/* a2: periodic_advect_1d_constant, src/semi_lagrangian/advection/sll_m_advection_1d_periodic.F90:101-130.
 * kind: 0 = sll_p_spline, 1 = sll_p_lagrange. n = size(input) (N or N+1). */
void orc_advect_1d_periodic_constant(int kind, int num_cells, double xmin, double xmax, int order,
                                     double A, double dt, const double *input, double *output, int n) {
    double shift = A * dt / (xmax - xmin) * (double)num_cells;
    double *tmp = (double *)malloc(sizeof(double) * num_cells);
    if (kind == 0) orc_periodic_interp_spline(num_cells, order, input, shift, tmp);
    else orc_periodic_interp_lagrange(num_cells, order, input, shift, tmp);
    memcpy(output, tmp, sizeof(double) * num_cells);
    if (n > num_cells) output[num_cells] = output[0];
    free(tmp);
}

/* a4: sll_t_advector_1d_bsl%advect_1d_constant, src/semi_lagrangian/advection/sll_m_advection_1d_BSL.F90:152-164,
 * with the objects the reference's own test wires in (test_advection_1d_BSL.F90): characteristics =
 * sll_t_charac_1d_explicit_euler with sll_p_periodic (feet = eta - dt*A, folded back into [eta_min, eta_max) by
 * sll_f_process_outside_point_periodic when outside or on the border;
 * src/time_integration/characteristics/sll_m_characteristics_1d_explicit_euler.F90:157-189,
 * sll_m_characteristics_1d_base.F90:84-108) and interp = sll_t_cubic_spline_interpolator_1d%interpolate_array
 * (compute_interpolant + eval_array, sll_m_cubic_spline_interpolator_1d.F90:97-109).
 * npts = num_cells + 1 points, eta_coords(i) = eta_min + (i-1)*delta (:109-113). */
void orc_advect_1d_bsl_constant(int npts, double eta_min, double eta_max, int fast, double A, double dt,
                                const double *input, double *output) {
    double *C = (double *)malloc(sizeof(double) * (npts + 3));
    double *feet = (double *)malloc(sizeof(double) * npts);
    double delta_eta = (eta_max - eta_min) / (double)(npts - 1);
    for (int i = 0; i < npts; ++i) {
        double eta = eta_min + (double)i * delta_eta;
        double o = eta - dt * A;
        if (o <= eta_min || o >= eta_max) {
            double e = (o - eta_min) / (eta_max - eta_min);
            e = e - floor(e);
            if (e == 1.0) e = 0.0;
            o = eta_min + e * (eta_max - eta_min);
        }
        feet[i] = o;
    }
    orc_spline_compute_interpolant_periodic(input, npts, fast, C);
    spline_eval_array(C, npts, eta_min, eta_max, feet, output, npts);
    free(C); free(feet);
}

/* ------------------------------------------------------------------------- */
/* a10: sll_m_lagrange_interpolation_1d_fast                                   */
/* ------------------------------------------------------------------------- */
/* lagr_{3,5,7,9,11}pt_coeff :239-246,286-295,341-352,405-419,478-494; even :59-67,110-121,170-182 */
int orc_lagr_coeff(int s, double p, double *pp) {
    const double inv_6 = 1. / 6., inv_12 = 1. / 12., inv_24 = 1. / 24., inv_36 = 1. / 36., inv_48 = 1. / 48.,
                 inv_120 = 1. / 120., inv_144 = 1. / 144., inv_240 = 1. / 240., inv_576 = 1. / 576.,
                 inv_720 = 1. / 720., inv_1440 = 1. / 1440., inv_5040 = 1. / 5040., inv_14400 = 1. / 14400.,
                 inv_17280 = 1. / 17280., inv_30240 = 1. / 30240., inv_40320 = 1. / 40320.,
                 inv_80640 = 1. / 80640., inv_362880 = 1. / 362880., inv_3628800 = 1. / 3628800.;
    double p2 = p * p;
    switch (s) {
    case 3:
        pp[0] = p * (p - 1.) * 0.5; pp[1] = 1. - p * p; pp[2] = p * (p + 1.) * 0.5; return 0;
    case 5:
        pp[0] = (p * p - 1.) * p * (p - 2.) * inv_24;
        pp[1] = -(p - 1.) * p * (p * p - 4.) * inv_6;
        pp[2] = (p * p - 1.) * (p * p - 4.) * 0.25;
        pp[3] = -(p + 1.) * p * (p * p - 4.) * inv_6;
        pp[4] = (p * p - 1.) * p * (p + 2.) * inv_24; return 0;
    case 7:
        pp[0] = p * (p - 3.) * (p2 - 4.) * (p2 - 1.) * inv_720;
        pp[1] = -p * (p - 2.) * (p2 - 9.) * (p2 - 1.) * inv_120;
        pp[2] = p * (p - 1.) * (p2 - 9.) * (p2 - 4.) * inv_48;
        pp[3] = -(p2 - 9.) * (p2 - 4.) * (p2 - 1.) * inv_36;
        pp[4] = (p + 1.) * p * (p2 - 9.) * (p2 - 4.) * inv_48;
        pp[5] = -(p + 2.) * p * (p2 - 9.) * (p2 - 1.) * inv_120;
        pp[6] = (p + 3.) * p * (p2 - 4.) * (p2 - 1.) * inv_720; return 0;
    case 9:
        pp[0] = p * (p - 4.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * inv_40320;
        pp[1] = -p * (p - 3.) * (p2 - 16.) * (p2 - 4.) * (p2 - 1.) * inv_5040;
        pp[2] = p * (p - 2.) * (p2 - 16.) * (p2 - 9.) * (p2 - 1.) * inv_1440;
        pp[3] = -p * (p - 1.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * inv_720;
        pp[4] = (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * inv_576;
        pp[5] = -(p + 1.) * p * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * inv_720;
        pp[6] = (p + 2.) * p * (p2 - 16.) * (p2 - 9.) * (p2 - 1.) * inv_1440;
        pp[7] = -(p + 3.) * p * (p2 - 16.) * (p2 - 4.) * (p2 - 1.) * inv_5040;
        pp[8] = (p + 4.) * p * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * inv_40320; return 0;
    case 11:
        pp[0] = p * (p - 5.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * inv_3628800;
        pp[1] = -p * (p - 4.) * (p2 - 25.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * inv_362880;
        pp[2] = p * (p - 3.) * (p2 - 25.) * (p2 - 16.) * (p2 - 4.) * (p2 - 1.) * inv_80640;
        pp[3] = -p * (p - 2.) * (p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 1.) * inv_30240;
        pp[4] = p * (p - 1.) * (p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * inv_17280;
        pp[5] = -(p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * inv_14400;
        pp[6] = (p + 1.) * p * (p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * inv_17280;
        pp[7] = -(p + 2.) * p * (p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 1.) * inv_30240;
        pp[8] = (p + 3.) * p * (p2 - 25.) * (p2 - 16.) * (p2 - 4.) * (p2 - 1.) * inv_80640;
        pp[9] = -(p + 4.) * p * (p2 - 25.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * inv_362880;
        pp[10] = (p + 5.) * p * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * inv_3628800; return 0;
    case 4:
        pp[0] = -p * (p - 1.) * (p - 2.) * inv_6;
        pp[1] = (p * p - 1.) * (p - 2.) * 0.5;
        pp[2] = -p * (p + 1.) * (p - 2.) * 0.5;
        pp[3] = p * (p * p - 1.0) * inv_6; return 0;
    case 6:
        pp[0] = -p * (p * p - 1.) * (p - 2.) * (p - 3.) * inv_120;
        pp[1] = p * (p - 1.) * (p * p - 4.) * (p - 3.) * inv_24;
        pp[2] = -(p * p - 1.) * (p * p - 4.) * (p - 3.) * inv_12;
        pp[3] = p * (p + 1.) * (p * p - 4.) * (p - 3.) * inv_12;
        pp[4] = -p * (p * p - 1.) * (p + 2.) * (p - 3.0) * inv_24;
        pp[5] = p * (p * p - 1.) * (p * p - 4.) * inv_120; return 0;
    case 8:
        pp[0] = -p * (p - 3) * (p - 4) * (p2 - 4) * (p2 - 1) * inv_5040;
        pp[1] = p * (p - 2) * (p - 4) * (p2 - 9) * (p2 - 1) * inv_720;
        pp[2] = -p * (p - 1) * (p - 4) * (p2 - 9) * (p2 - 4) * inv_240;
        pp[3] = (p - 4) * (p2 - 9) * (p2 - 4) * (p2 - 1) * inv_144;
        pp[4] = -(p + 1) * p * (p - 4) * (p2 - 9) * (p2 - 4) * inv_144;
        pp[5] = (p + 2) * p * (p - 4) * (p2 - 9) * (p2 - 1) * inv_240;
        pp[6] = -(p + 3) * p * (p - 4) * (p2 - 4) * (p2 - 1) * inv_720;
        pp[7] = p * (p2 - 9) * (p2 - 4) * (p2 - 1) * inv_5040; return 0;
    default: return -1;
    }
}

/* left-to-right sum as in lagr_Npt / lagr_Npt_vec (:393-399 for 7 pt) */
static inline double lagr_dot(const double *pp, int s, const double *fi, int i0, int n_wrap) {
    /* i0 = 0-based index of first stencil point; n_wrap>0 wraps periodically */
    double acc;
    if (n_wrap > 0) {
        acc = pp[0] * fi[imod(i0, n_wrap)];
        for (int k = 1; k < s; ++k) acc += pp[k] * fi[imod(i0 + k, n_wrap)];
    } else {
        acc = pp[0] * fi[i0];
        for (int k = 1; k < s; ++k) acc += pp[k] * fi[i0 + k];
    }
    return acc;
}

/* sll_s_lagrange_interpolation_1d_fast_disp_fixed_periodic :609-654 (any odd s in 3..11;
 * the reference implements 3,5,7) */
int orc_lagrange_fixed_periodic(const double *fi, double *fp, int n, double p, int s) {
    double pp[11];
    if (s % 2 == 0 || orc_lagr_coeff(s, p, pp)) return -1;
    int h = (s - 1) / 2;
    for (int i = 0; i < n; ++i) fp[i] = lagr_dot(pp, s, fi, i - h, n);
    return 0;
}
/* ..._fixed_periodicl :663-705: fi, fp have n+1 entries */
int orc_lagrange_fixed_periodicl(const double *fi, double *fp, int np, double p, int s) {
    int n = np - 1;
    if (orc_lagrange_fixed_periodic(fi, fp, n, p, s)) return -1;
    fp[n] = fp[0];
    return 0;
}
/* ..._fixed_haloc_cells :783-808: only i = h+1 .. n-h are written */
int orc_lagrange_fixed_haloc_cells(const double *fi, double *fp, int n, double p, int s) {
    double pp[11];
    if (s % 2 == 0 || orc_lagr_coeff(s, p, pp)) return -1;
    int h = (s - 1) / 2;
    for (int i = h; i < n - h; ++i) fp[i] = lagr_dot(pp, s, fi, i - h, 0);
    return 0;
}
/* ..._fixed_no_bc :562-600 (stencil 3 and 5) */
int orc_lagrange_fixed_no_bc(const double *fi, double *fp, int n, double p, int s) {
    double pp[11];
    if (s != 3 && s != 5) return -1;
    int h = (s - 1) / 2;
    orc_lagr_coeff(s, p, pp);
    for (int i = h; i < n - h; ++i) fp[i] = lagr_dot(pp, s, fi, i - h, 0);
    for (int e = 0; e < h; ++e) {
        /* left: point e uses stencil starting at 0 with offset p-(h-e) */
        orc_lagr_coeff(s, p - (double)(h - e), pp);
        fp[e] = lagr_dot(pp, s, fi, 0, 0);
        orc_lagr_coeff(s, p + (double)(h - e), pp);
        fp[n - 1 - e] = lagr_dot(pp, s, fi, n - s, 0);
    }
    return 0;
}
/* ..._centered_periodicl :710-769 (4, 6; 8 by the same rule) fi/fp have n+1 entries */
int orc_lagrange_centered_periodicl(const double *fi, double *fp, int np, double p, int s) {
    double pp[11];
    int n = np - 1;
    int pi = (int)floor(p);
    double pq = p - (double)pi;
    if (s % 2 != 0 || orc_lagr_coeff(s, pq, pp)) return -1;
    int h = s / 2 - 1;
    for (int i = 0; i < n; ++i) fp[i] = lagr_dot(pp, s, fi, i - h + pi, n);
    fp[n] = fp[0];
    return 0;
}
/* ..._centered_halo_cells :811-837 with lagr_{4,6,8}pt_vec index ranges (:90-107 etc.) */
int orc_lagrange_centered_halo_cells(const double *fi, double *fp, int n, double p, int s) {
    double pp[11];
    int pi = (int)floor(p);
    double pq = p - (double)pi;
    if (s % 2 != 0 || orc_lagr_coeff(s, pq, pp)) return -1;
    int h = s / 2 - 1;
    /* do i = max(h+1-pi,1), min(n-(h+1)-pi, n) (1-based) */
    int lo = (h + 1 - pi > 1) ? h + 1 - pi : 1;
    int hi = (n - (h + 1) - pi < n) ? n - (h + 1) - pi : n;
    for (int i = lo; i <= hi; ++i) fp[i - 1] = lagr_dot(pp, s, fi, i - 1 - h + pi, 0);
    return 0;
}

/* a9: barycentric centred Lagrange, src/interpolation/lagrange_interpolation/sll_m_lagrange_interpolation_1d.F90:72-174
 * called through sll_t_lagrange_interpolator_1d%interpolate_array_disp (centred branch,
 * sll_m_lagrange_interpolator_1d.F90:176-178): out(i) = f(x_i + alpha). periodic only. */
void orc_lagrange_centered_barycentric(const double *fi, double *out, int num_points, double xmin,
                                       double xmax, int d, int periodic_last, double alpha_in) {
    int nb_cell = num_points - 1;
    double deta = (xmax - xmin) / (double)nb_cell;
    double wj[32], wjs[32];
    int table[64];
    for (int i = 1; i <= 2 * d - 1; ++i) { table[i - 1] = 2 * d - 1 - (i - 1); table[i + 2 * d - 2] = i; }
    for (int i = 0; i < 2 * d; ++i) wj[i] = 1.0;
    for (int i = 1; i <= d; ++i) {
        for (int j = 1; j <= 2 * d - 1; ++j) wj[i - 1] *= (double)table[i + j - 2];
        wj[i - 1] = pow(-1.0, (double)(d + i)) * wj[i - 1];
    }
    for (int i = 1; i <= d; ++i) wj[i + d - 1] = -wj[d - i];
    for (int i = 0; i < 2 * d; ++i) wj[i] = 1.0 / wj[i];
    double alpha = alpha_in; /* eval_array(data,-alpha): lagrange%alpha = -(-alpha) */
    double h = deta;
    int index_gap = (int)floor(alpha / h);
    double beta = alpha / h - (double)index_gap;
    if (beta == 1.0) { beta = 0.0; index_gap += 1; }
    int nout = nb_cell + periodic_last;
    if (beta == 0.0) {
        for (int j = 1; j <= nout; ++j) out[j - 1] = fi[imod(index_gap + j - 1, nb_cell)];
    } else {
        double sum2 = 0.0;
        for (int j = 1; j <= 2 * d; ++j) { wjs[j - 1] = wj[j - 1] / (beta + (double)(d - j)); sum2 += wjs[j - 1]; }
        for (int i = 1; i <= nout; ++i) {
            double sum1 = 0.0;
            for (int j = 1; j <= 2 * d; ++j)
                sum1 += wjs[j - 1] * fi[imod(index_gap + (i - 1) + (j - 1) - (d - 1), nb_cell)];
            out[i - 1] = sum1 / sum2;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* a15: periodic Poisson solvers                                               */
/* ------------------------------------------------------------------------- */
/* solve_poisson_1d_periodic, src/field_solvers/poisson_solvers/sll_m_poisson_1d_periodic.F90:125-173.
 * field, rhs: nc+1 values. */
void orc_poisson_1d_periodic_solve(int nc, double xmin, double xmax, const double *rhs, double *field) {
    double *work = (double *)malloc(sizeof(double) * (nc + 1));
    memcpy(work, rhs, sizeof(double) * (nc + 1));
    dfftf_(nc, work);
    for (int i = 0; i <= nc; ++i) work[i] = work[i] / (double)nc;
    double kx0 = 2.0 * ORC_PI / (xmax - xmin);
    field[0] = 0.0;
    for (int ik = 1; ik <= (nc - 2) / 2; ++ik) {
        double kx = (double)ik * kx0, k2 = kx * kx;
        field[2 * ik - 1] = kx / k2 * work[2 * ik];
        field[2 * ik] = -kx / k2 * work[2 * ik - 1];
    }
    field[nc - 1] = 0.0;
    dfftb_(nc, field);
    field[nc] = field[0];
    free(work);
}

/* r2c / c2r 2D with FFTW semantics (halved dimension = first, fastest).  c2r: complex
 * inverse along x2 first, then a 1D c2r along x1 that ignores Im at i=1 and i=N1/2+1
 * (what FFTW's multi-dimensional c2r does; SURVEY.md section 7, "Non-Hermitian spectra"). */
static void r2c_2d(const double *in, int n1, int n2, cplx *out /* (n1/2+1) x n2 */) {
    int nh = n1 / 2 + 1;
    cplx *full = (cplx *)malloc(sizeof(cplx) * n1 * n2);
    cplx *w = (cplx *)malloc(sizeof(cplx) * (n1 > n2 ? n1 : n2));
    for (long i = 0; i < (long)n1 * n2; ++i) full[i] = in[i];
    for (int j = 0; j < n2; ++j) fft_lines(full + (long)j * n1, n1, 1, -1, w);
    for (int i = 0; i < nh; ++i) fft_lines(full + i, n2, n1, -1, w);
    for (int j = 0; j < n2; ++j)
        for (int i = 0; i < nh; ++i) out[i + (long)j * nh] = full[i + (long)j * n1];
    free(full); free(w);
}
static void c2r_2d(const cplx *in, int n1, int n2, double *out) {
    int nh = n1 / 2 + 1;
    cplx *t = (cplx *)malloc(sizeof(cplx) * nh * n2);
    cplx *w = (cplx *)malloc(sizeof(cplx) * (n1 > n2 ? n1 : n2));
    cplx *line = (cplx *)malloc(sizeof(cplx) * n1);
    memcpy(t, in, sizeof(cplx) * nh * n2);
    for (int i = 0; i < nh; ++i) fft_lines(t + i, n2, nh, +1, w);
    for (int j = 0; j < n2; ++j) {
        line[0] = creal(t[(long)j * nh]);
        for (int i = 1; i < nh; ++i) {
            if (2 * i == n1) line[i] = creal(t[i + (long)j * nh]);
            else { line[i] = t[i + (long)j * nh]; line[n1 - i] = conj(t[i + (long)j * nh]); }
        }
        fft_inplace(line, n1, +1);
        for (int i = 0; i < n1; ++i) out[i + (long)j * n1] = creal(line[i]);
    }
    free(t); free(w); free(line);
}

/* initialize + solve_e_fields_poisson_2d_periodic_fft,
 * src/field_solvers/poisson_solvers/sll_m_poisson_2d_periodic.F90:250-310,342-383.
 * rho, ex, ey: (ld1 x ld2) with ld = nc or nc+1 (duplicates filled when nc+1). */
/* sll_s_poisson_2d_periodic_par_solve, src/field_solvers/poisson_solvers_parallel/sll_m_poisson_2d_periodic_par.F90:214-338,
 * on one process (the remaps between the two sequential layouts are the identity): complex transforms along x then y of
 * rho(1:ncx, 1:ncy), phi^ = -rho^ / (((kx/Lx)^2 + (ky/Ly)^2) 4 pi^2) with both mode numbers folded to -n/2 .. n/2-1 and
 * the (1,1) mode set to zero, inverse transforms along y then x, periodic point copied in both directions (:316,329),
 * real part times 1/(ncx ncy).  rho and phi are (ncx+1) x (ncy+1), column-major. */
void orc_poisson_2d_periodic_par_solve(int ncx, int ncy, double Lx, double Ly, const double *rho, double *phi) {
    const int ld = ncx + 1;
    cplx *a = (cplx *)malloc(sizeof(cplx) * (size_t)ncx * ncy);
    cplx *work = (cplx *)malloc(sizeof(cplx) * (size_t)(ncx > ncy ? ncx : ncy));
    for (int j = 0; j < ncy; ++j) for (int i = 0; i < ncx; ++i) a[i + (size_t)ncx * j] = rho[i + (size_t)ld * j];
    for (int j = 0; j < ncy; ++j) fft_lines(a + (size_t)ncx * j, ncx, 1, -1, work);
    for (int i = 0; i < ncx; ++i) fft_lines(a + i, ncy, ncx, -1, work);
    const double r_Lx = 1.0 / Lx, r_Ly = 1.0 / Ly;
    for (int j = 0; j < ncy; ++j) for (int i = 0; i < ncx; ++i) {
        if (i == 0 && j == 0) { a[0] = 0; continue; }
        double kx = (double)i, ky = (double)j;
        if (kx >= ncx / 2) kx = kx - ncx;
        if (ky >= ncy / 2) ky = ky - ncy;
        a[i + (size_t)ncx * j] = -a[i + (size_t)ncx * j] / ((pow(kx * r_Lx, 2) + pow(ky * r_Ly, 2)) * 4.0 * ORC_PI * ORC_PI);
    }
    for (int i = 0; i < ncx; ++i) fft_lines(a + i, ncy, ncx, +1, work);
    for (int j = 0; j < ncy; ++j) fft_lines(a + (size_t)ncx * j, ncx, 1, +1, work);
    const double normalization = 1.0 / ((double)ncx * (double)ncy);
    for (int j = 0; j <= ncy; ++j) for (int i = 0; i <= ncx; ++i)
        phi[i + (size_t)ld * j] = creal(a[(i % ncx) + (size_t)ncx * (j % ncy)]) * normalization;
    free(a); free(work);
}

void orc_poisson_2d_periodic_solve_e(int nc_x, int nc_y, double x_min, double x_max, double y_min,
                                     double y_max, const double *rho, int ld1, int ld2,
                                     double *e_x, double *e_y, double *phi) {
    int nh = nc_x / 2 + 1;
    double *tmp = (double *)malloc(sizeof(double) * nc_x * nc_y);
    cplx *rht = (cplx *)malloc(sizeof(cplx) * nh * nc_y);
    cplx *exy = (cplx *)malloc(sizeof(cplx) * nh * nc_y);
    double *kx = (double *)malloc(sizeof(double) * nh * nc_y);
    double *ky = (double *)malloc(sizeof(double) * nh * nc_y);
    double *k2 = (double *)malloc(sizeof(double) * nh * nc_y);
    double kx0 = 2.0 * ORC_PI / (x_max - x_min), ky0 = 2.0 * ORC_PI / (y_max - y_min);
    for (int ik = 1; ik <= nh; ++ik) {
        double kx1 = (ik - 1) * kx0;
        for (int jk = 1; jk <= nc_y / 2; ++jk) { kx[ik - 1 + (long)(jk - 1) * nh] = kx1; ky[ik - 1 + (long)(jk - 1) * nh] = (jk - 1) * ky0; }
        for (int jk = nc_y / 2 + 1; jk <= nc_y; ++jk) { kx[ik - 1 + (long)(jk - 1) * nh] = kx1; ky[ik - 1 + (long)(jk - 1) * nh] = (jk - 1 - nc_y) * ky0; }
    }
    kx[0] = 1.0;
    for (long i = 0; i < (long)nh * nc_y; ++i) { k2[i] = kx[i] * kx[i] + ky[i] * ky[i]; kx[i] = kx[i] / k2[i]; ky[i] = ky[i] / k2[i]; }
    for (int j = 0; j < nc_y; ++j) for (int i = 0; i < nc_x; ++i) tmp[i + (long)j * nc_x] = rho[i + (long)j * ld1];
    r2c_2d(tmp, nc_x, nc_y, rht);
    double norm = (double)(nc_x * nc_y);
    if (e_x) {
        for (long i = 0; i < (long)nh * nc_y; ++i) exy[i] = -(I * kx[i]) * rht[i];
        c2r_2d(exy, nc_x, nc_y, tmp);
        for (int j = 0; j < nc_y; ++j) for (int i = 0; i < nc_x; ++i) e_x[i + (long)j * ld1] = tmp[i + (long)j * nc_x] / norm;
        for (long i = 0; i < (long)nh * nc_y; ++i) exy[i] = -(I * ky[i]) * rht[i];
        c2r_2d(exy, nc_x, nc_y, tmp);
        for (int j = 0; j < nc_y; ++j) for (int i = 0; i < nc_x; ++i) e_y[i + (long)j * ld1] = tmp[i + (long)j * nc_x] / norm;
        if (ld1 == nc_x + 1) for (int j = 0; j < ld2; ++j) { e_x[nc_x + (long)j * ld1] = e_x[(long)j * ld1]; e_y[nc_x + (long)j * ld1] = e_y[(long)j * ld1]; }
        if (ld2 == nc_y + 1) for (int i = 0; i < ld1; ++i) { e_x[i + (long)nc_y * ld1] = e_x[i]; e_y[i + (long)nc_y * ld1] = e_y[i]; }
    }
    if (phi) { /* solve_potential :314-338 */
        for (long i = 0; i < (long)nh * nc_y; ++i) exy[i] = rht[i] / k2[i];
        c2r_2d(exy, nc_x, nc_y, tmp);
        for (int j = 0; j < nc_y; ++j) for (int i = 0; i < nc_x; ++i) phi[i + (long)j * ld1] = tmp[i + (long)j * nc_x] / norm;
        if (ld1 == nc_x + 1) for (int j = 0; j < ld2; ++j) phi[nc_x + (long)j * ld1] = phi[(long)j * ld1];
        if (ld2 == nc_y + 1) for (int i = 0; i < ld1; ++i) phi[i + (long)nc_y * ld1] = phi[i];
    }
    free(tmp); free(rht); free(exy); free(kx); free(ky); free(k2);
}

/* sll_s_poisson_3d_periodic_par_solve + ..._compute_e_from_phi (single rank),
 * src/field_solvers/poisson_solvers_parallel/sll_m_poisson_3d_periodic_par.F90:297-470,981-1158 */
void orc_poisson_3d_periodic_solve(int nx, int ny, int nz, double Lx, double Ly, double Lz,
                                   const double *rho, double *phi, double *ex, double *ey, double *ez) {
    long ntot = (long)nx * ny * nz;
    int nmax = nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
    cplx *a = (cplx *)malloc(sizeof(cplx) * ntot);
    cplx *w = (cplx *)malloc(sizeof(cplx) * nmax);
    double kx0 = ORC_TWOPI / Lx, ky0 = ORC_TWOPI / Ly, kz0 = ORC_TWOPI / Lz;
    double nxyz_inv = 1.0 / (double)((long)nx * ny * nz);
    for (long i = 0; i < ntot; ++i) a[i] = rho[i];
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) fft_lines(a + ((long)k * ny + j) * nx, nx, 1, -1, w);
    for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) fft_lines(a + (long)k * ny * nx + i, ny, nx, -1, w);
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        fft_lines(a + (long)j * nx + i, nz, (long)nx * ny, -1, w);
        for (int k = 0; k < nz; ++k) a[((long)k * ny + j) * nx + i] *= nxyz_inv;
    }
    for (int gk = 1; gk <= nz; ++gk) for (int gj = 1; gj <= ny; ++gj) for (int gi = 1; gi <= nx; ++gi) {
        long idx = ((long)(gk - 1) * ny + (gj - 1)) * nx + (gi - 1);
        if (gi == 1 && gj == 1 && gk == 1) { a[idx] = 0; continue; }
        double ind_x = (gi <= nx / 2) ? (double)(gi - 1) : (double)(nx - (gi - 1));
        double ind_y = (gj <= ny / 2) ? (double)(gj - 1) : (double)(ny - (gj - 1));
        double ind_z = (gk <= nz / 2) ? (double)(gk - 1) : (double)(nz - (gk - 1));
        double kx = kx0 * ind_x, ky = ky0 * ind_y, kz = kz0 * ind_z;
        a[idx] = a[idx] / (kx * kx + ky * ky + kz * kz);
    }
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) fft_lines(a + (long)j * nx + i, nz, (long)nx * ny, +1, w);
    for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) fft_lines(a + (long)k * ny * nx + i, ny, nx, +1, w);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) fft_lines(a + ((long)k * ny + j) * nx, nx, 1, +1, w);
    for (long i = 0; i < ntot; ++i) phi[i] = creal(a[i]);
    /* E = -grad phi, one spectral derivative per direction (:1079-1158 and siblings) */
    for (int dir = 0; dir < 3; ++dir) {
        double *e = dir == 0 ? ex : (dir == 1 ? ey : ez);
        if (!e) continue;
        int n = dir == 0 ? nx : (dir == 1 ? ny : nz);
        long stride = dir == 0 ? 1 : (dir == 1 ? nx : (long)nx * ny);
        double k0 = dir == 0 ? kx0 : (dir == 1 ? ky0 : kz0);
        double norm_fac = 1.0 / (double)n;
        long nlines = ntot / n;
        for (long l = 0; l < nlines; ++l) {
            long base;
            if (dir == 0) base = l * nx;
            else if (dir == 1) base = (l / nx) * (long)nx * ny + (l % nx);
            else base = l;
            for (int i = 0; i < n; ++i) w[i] = phi[base + i * stride];
            fft_inplace(w, n, -1);
            for (int i = 1; i <= n / 2; ++i) w[i - 1] = w[i - 1] * (I * (-k0 * (double)(i - 1))) * norm_fac;
            for (int i = n / 2 + 1; i <= n; ++i) w[i - 1] = w[i - 1] * (I * (k0 * (double)(n - i + 1))) * norm_fac;
            fft_inplace(w, n, +1);
            for (int i = 0; i < n; ++i) e[base + i * stride] = creal(w[i]);
        }
    }
    free(a); free(w);
}

/* ------------------------------------------------------------------------- */
/* a14: velocity reductions                                                    */
/* ------------------------------------------------------------------------- */
/* sll_s_compute_reduction_4d_to_2d_direction34, src/parallelization/reduction/sll_m_reduction.F90:187-272 */
void orc_reduction_4d_to_2d_direction34(const double *f, int n1, int n2, int n3, int n4,
                                        double delta3, double delta4, double *out) {
#pragma omp parallel for schedule(static)
    for (int i2 = 0; i2 < n2; ++i2)
        for (int i1 = 0; i1 < n1; ++i1) {
            double acc = 0.0;
            for (int i3 = 0; i3 < n3; ++i3) {
#define F4(a, b, c, d) f[(a) + (long)n1 * ((b) + (long)n2 * ((c) + (long)n3 * (d)))]
                double tmp = 0.5 * (F4(i1, i2, i3, 0) + F4(i1, i2, i3, n4 - 1));
                for (int i4 = 1; i4 < n4 - 1; ++i4) tmp += F4(i1, i2, i3, i4);
                tmp *= delta4;
                acc += (i3 == 0 || i3 == n3 - 1) ? 0.5 * tmp : tmp;
            }
            out[i1 + (long)n1 * i2] = acc * delta3;
        }
}
/* sll_s_compute_charge_density_6d_core, simulations/parallel/bsl_vp_3d3v_cart_dd/sll_m_sim_6d_utilities.F90:203-245 */
void orc_charge_density_6d(const double *f, const int n[6], double volume_v, double *rho) {
    long nx = (long)n[0] * n[1] * n[2], nv = (long)n[3] * n[4] * n[5];
#pragma omp parallel for schedule(static)
    for (long jk = 0; jk < (long)n[1] * n[2]; ++jk) {
        double *sm = rho + jk * n[0];
        for (int i = 0; i < n[0]; ++i) sm[i] = 0.0;
        for (long v = 0; v < nv; ++v) {
            const double *row = f + v * nx + jk * n[0];
            for (int i = 0; i < n[0]; ++i) sm[i] -= row[i];
        }
        for (int i = 0; i < n[0]; ++i) sm[i] *= volume_v;
    }
}

/* ------------------------------------------------------------------------- */
/* Batched line loops exactly as the simulations call the advectors (a16):     */
/* copy line -> per-line routine -> copy back.  Used for parity and as the     */
/* timed CPU baseline.                                                         */
/* ------------------------------------------------------------------------- */
/* method: 0 = cubic spline direct (a5/a6: compute_interpolant + eval_disp),
 *         1 = periodic advector sll_p_spline via FFT (a2/a3, `order`),
 *         2 = periodic advector sll_p_lagrange via FFT (a2/a3, `order`),
 *         3 = Lagrange fixed periodic, odd stencil `order` (a10),
 *         4 = Lagrange centred periodic, even stencil `order` (a10).
 * f is [outer][n][inner] (column-major: inner fastest), periodic with n cells (no
 * duplicate point).  disp[] gives, per line, the displacement in CELLS such that
 * out(i) = f(i + disp).  disp index = (o*inner + in) * dstride_flag ... see below. */
static void advect_line(int method, int order, int n, const double *in, double *out, double disp_cells,
                        double *scratch) {
    switch (method) {
    case 0: { /* np = n+1, xmin=0, xmax=n -> delta = 1 */
        double *line = scratch, *C = scratch + (n + 1), *o = C + (n + 4);
        memcpy(line, in, sizeof(double) * n); line[n] = line[0];
        orc_spline_compute_interpolant_periodic(line, n + 1, -1, C);
        orc_spline_eval_disp_periodic(C, n + 1, 0.0, (double)n, disp_cells, o);
        memcpy(out, o, sizeof(double) * n);
    } break;
    case 1: orc_periodic_interp_spline(n, order, in, -disp_cells, out); break;
    case 2: { double *t = scratch; orc_periodic_interp_lagrange(n, order, in, -disp_cells, t); memcpy(out, t, sizeof(double) * n); } break;
    case 3: orc_lagrange_fixed_periodic(in, out, n, disp_cells, order); break;
    case 4: { double *li = scratch, *lo = scratch + n + 1; memcpy(li, in, sizeof(double) * n); li[n] = li[0];
              orc_lagrange_centered_periodicl(li, lo, n + 1, disp_cells, order); memcpy(out, lo, sizeof(double) * n); } break;
    }
}

/* disp_mode: 0 = one displacement per outer index o (disp[o*dmul_o ...]) -- generic form:
 * displacement of line (o, in) = disp[ (o / odiv) % omod * ostr + (in / idiv) % imod * istr ].
 * This covers: x-advection (disp depends on one velocity index) and v-advection
 * (disp depends on the (x1,x2[,x3]) position = low part of the inner index). */
/* outer_count / inner_count restrict the pass to the first lines of each index range (a bounded sample of
 * a pass with full-length lines, used by the timed CPU baseline); the full pass has counts = extents. */
void orc_advect_axis_sub(double *f, long outer, int n, long inner, long outer_count, long inner_count, int method,
                         int order, const double *disp, long odiv, long omod, long ostr, long idiv, long imodn,
                         long istr) {
#pragma omp parallel
    {
        double *lin = (double *)malloc(sizeof(double) * (4 * (size_t)n + 16));
        double *lout = lin + n;
        double *scratch = (double *)malloc(sizeof(double) * (4 * (size_t)n + 32));
#pragma omp for schedule(static) collapse(2)
        for (long o = 0; o < outer_count; ++o)
            for (long in = 0; in < inner_count; ++in) {
                double *base = f + o * (long)n * inner + in;
                double dc = disp[((o / odiv) % omod) * ostr + ((in / idiv) % imodn) * istr];
                for (int i = 0; i < n; ++i) lin[i] = base[(long)i * inner];
                advect_line(method, order, n, lin, lout, dc, scratch);
                for (int i = 0; i < n; ++i) base[(long)i * inner] = lout[i];
            }
        free(lin); free(scratch);
    }
}
void orc_advect_axis(double *f, long outer, int n, long inner, int method, int order,
                     const double *disp, long odiv, long omod, long ostr, long idiv, long imodn, long istr) {
    orc_advect_axis_sub(f, outer, n, inner, outer, inner, method, order, disp, odiv, omod, ostr, idiv, imodn, istr);
}

/* ------------------------------------------------------------------------- */
/* 6D simulation sim_bsl_vp_3d3v_cart_dd_slim, single rank, Lagrange "fixed"   */
/* simulations/parallel/bsl_vp_3d3v_cart_dd/sll_m_sim_bsl_vp_3d3v_cart_dd_slim.F90:278-960 */
/* ------------------------------------------------------------------------- */
/* advect_eta{id} with halo from the local periodic copy (procs(id)==1 branch of
 * sll_s_apply_halo_exchange_slim_6d_real64, sll_m_decomposition.F90:1840-1861,1958-1979)
 * then sll_s_lagrange_interpolation_1d_fast_disp_fixed_haloc_cells on l_halo|f|r_halo
 * (sll_m_advection_6d_lagrange_dd_slim.F90:806-967, 1558-1704). */
static void advect6d_axis(double *f, const int n[6], int id, int s, const double *disp, int disp_is_field) {
    long inner = 1, outer = 1;
    for (int d = 0; d < id; ++d) inner *= n[d];
    for (int d = id + 1; d < 6; ++d) outer *= n[d];
    int nn = n[id], h = (s - 1) / 2;
    long nx3 = (long)n[0] * n[1] * n[2];
#pragma omp parallel
    {
        double *bi = (double *)malloc(sizeof(double) * 2 * (nn + 2 * h));
        double *bo = bi + nn + 2 * h;
#pragma omp for schedule(static) collapse(2)
        for (long o = 0; o < outer; ++o)
            for (long in = 0; in < inner; ++in) {
                double *base = f + o * (long)nn * inner + in;
                double p;
                if (disp_is_field) p = disp[in % nx3];           /* eta4..6: displacement(i,j,k) */
                else {                                            /* eta1..3: displacement(l|m|n) */
                    /* conjugate velocity index: axis id+3 lives in the outer index */
                    long stride_o = 1;
                    for (int d = id + 1; d < id + 3; ++d) stride_o *= n[d];
                    p = disp[(o / stride_o) % n[id + 3]];
                }
                for (int i = 0; i < h; ++i) bi[i] = base[(long)(nn - h + i) * inner];
                for (int i = 0; i < nn; ++i) bi[h + i] = base[(long)i * inner];
                for (int i = 0; i < h; ++i) bi[h + nn + i] = base[(long)i * inner];
                orc_lagrange_fixed_haloc_cells(bi, bo, nn + 2 * h, p, s);
                for (int i = 0; i < nn; ++i) base[(long)i * inner] = bo[h + i];
            }
        free(bi);
    }
}

/* sll_s_time_history_diagnostics, sll_m_sim_6d_utilities.F90:249-644: 13 numbers */
static void diagnostics6d(const double *f, const int n[6], const double vmin[3], const double dv[3],
                          double dV, double dVx, const double *rho, const double *phi, const double *ex,
                          const double *ey, const double *ez, double *out13) {
    long nx3 = (long)n[0] * n[1] * n[2];
    double smm = 0, smsq = 0, p4 = 0, p5 = 0, p6 = 0, k4 = 0, k5 = 0, k6 = 0;
#pragma omp parallel for collapse(3) reduction(+ : smm, smsq, p4, p5, p6, k4, k5, k6) schedule(static)
    for (int nn = 0; nn < n[5]; ++nn)
        for (int m = 0; m < n[4]; ++m)
            for (int l = 0; l < n[3]; ++l) {
                const double *blk = f + (((long)nn * n[4] + m) * n[3] + l) * nx3;
                double sm = 0.0;
                for (long i = 0; i < nx3; ++i) { sm += blk[i]; smsq += blk[i] * blk[i]; }
                double v4 = vmin[0] + dv[0] * l, v5 = vmin[1] + dv[1] * m, v6 = vmin[2] + dv[2] * nn;
                p4 += sm * v4 * dV; p5 += sm * v5 * dV; p6 += sm * v6 * dV;
                k4 += sm * v4 * v4 * dV; k5 += sm * v5 * v5 * dV; k6 += sm * v6 * v6 * dV;
                smm += sm;
            }
    out13[0] = smm * dV; out13[1] = smsq * dV;
    const double *arr[5] = {rho, phi, ex, ey, ez};
    for (int a = 0; a < 5; ++a) { double s = 0; for (long i = 0; i < nx3; ++i) s += arr[a][i] * arr[a][i]; out13[2 + a] = s * dVx; }
    out13[7] = p4; out13[8] = p5; out13[9] = p6; out13[10] = k4; out13[11] = k5; out13[12] = k6;
}

/* Runs init + `nsteps` steps; writes (nsteps+1) rows of 14 numbers (time + 13) to `rows`.
 * landau_prod initial data: sll_m_distribution_function_initializer_6d.F90:547-564,689-702;
 * local grid :312-335.  f (if non-NULL) receives the final distribution. */
/* advector (interpolator_type, sll_m_sim_bsl_vp_3d3v_cart_dd_slim.F90:363-372): 0 "fixed", 1 "centered" (eta1..3 by
 * fadvect_eta1..3 of the Lagrange advector: blocks of make_blocks_lagrange, centered_halo_cells on halo|f|halo with the
 * halo a periodic copy when the axis is not split, i.e. centered_periodicl arithmetic; lines with zero displacement
 * belong to no block), 2 "spline" (sll_t_advection_6d_spline_dd_slim; vblk[d] = number of ring ranks along eta4+d the
 * run is emulated for, the local spline depends on it). */
int orc_spline_dd_advect_axis(double *f, long outer, int n, long inner, int nblk, const double *dvals, const int *shifts,
                              long odiv, long omod, long ostr, long idiv, long imodn, long istr);
int orc_make_blocks_spline(int n, const double *disp_in, int *shift, double *alpha);
int orc_sim6d_run_ex(const int n[6], double v_max, const double xmax[3], int stencil_x, int stencil_v,
                     double delta_t, int nsteps, double alpha, const double kx[3], const double vth[3],
                     int time_in_phase, double *rows, double *f_out, int advector, const int vblk[3]);
int orc_sim6d_run(const int n[6], double v_max, const double xmax[3], int stencil_x, int stencil_v,
                  double delta_t, int nsteps, double alpha, const double kx[3], const double vth[3],
                  int time_in_phase, double *rows, double *f_out) {
    const int one[3] = {1, 1, 1};
    return orc_sim6d_run_ex(n, v_max, xmax, stencil_x, stencil_v, delta_t, nsteps, alpha, kx, vth, time_in_phase, rows,
                            f_out, 0, one);
}
/* x-advection along eta(id+1) of the non-"fixed" advectors, displacement array over the conjugate velocity index */
static void advect6d_x_blocks(double *f, const int n[6], int id, int advector, int stencil_x, const double *disp) {
    long inner = 1, outer = 1, stride_o = 1;
    for (int d = 0; d < id; ++d) inner *= n[d];
    for (int d = id + 1; d < 6; ++d) outer *= n[d];
    for (int d = id + 1; d < id + 3; ++d) stride_o *= n[d];
    int nv = n[id + 3];
    int *shift = (int *)malloc(sizeof(int) * nv);
    double *al = (double *)malloc(sizeof(double) * nv);
    orc_make_blocks_spline(nv, disp, shift, al); /* make_blocks_lagrange walks the array identically (:237-283) */
    if (advector == 2) {
        orc_spline_dd_advect_axis(f, outer, n[id], inner, 1, disp, shift, stride_o, nv, 1, 1, 1, 0);
    } else {
        int nn = n[id];
#pragma omp parallel
        {
            double *lin = (double *)malloc(sizeof(double) * 4 * (nn + 2));
            double *lout = lin + nn + 1, *scratch = lout + nn + 1;
#pragma omp for schedule(static) collapse(2)
            for (long o = 0; o < outer; ++o)
                for (long in = 0; in < inner; ++in) {
                    long l = (o / stride_o) % nv;
                    if (shift[l] == (-2147483647 - 1)) continue;
                    double *base = f + o * (long)nn * inner + in;
                    for (int i = 0; i < nn; ++i) lin[i] = base[(long)i * inner];
                    advect_line(4, stencil_x, nn, lin, lout, disp[l], scratch);
                    for (int i = 0; i < nn; ++i) base[(long)i * inner] = lout[i];
                }
            free(lin);
        }
    }
    free(shift); free(al);
}
int orc_sim6d_run_ex(const int n[6], double v_max, const double xmax[3], int stencil_x, int stencil_v,
                     double delta_t, int nsteps, double alpha, const double kx[3], const double vth[3],
                     int time_in_phase, double *rows, double *f_out, int advector, const int vblk[3]) {
    long ntot = 1; for (int d = 0; d < 6; ++d) ntot *= n[d];
    long nx3 = (long)n[0] * n[1] * n[2];
    double eta_min[6] = {0, 0, 0, -v_max, -v_max, -v_max};
    double eta_max[6] = {xmax[0], xmax[1], xmax[2], v_max, v_max, v_max};
    double de[6];
    for (int d = 0; d < 6; ++d) de[d] = (eta_max[d] - eta_min[d]) / (double)n[d]; /* sll_m_cartesian_meshes.F90:287 */
    double *f = (double *)malloc(sizeof(double) * ntot);
    double *rho = (double *)malloc(sizeof(double) * nx3 * 5);
    double *phi = rho + nx3, *ex = phi + nx3, *ey = ex + nx3, *ez = ey + nx3;
    if (!f || !rho) return -1;
    double *eta[6];
    for (int d = 0; d < 6; ++d) {
        eta[d] = (double *)malloc(sizeof(double) * n[d]);
        for (int k = 1; k <= n[d]; ++k) eta[d][k - 1] = eta_min[d] + de[d] * (double)(1 - 2 + k);
    }
    double factor = 1.0 / (pow(sqrt(ORC_TWOPI), 3) * (vth[0] * vth[1] * vth[2]));
#pragma omp parallel for collapse(3) schedule(static)
    for (int nn = 0; nn < n[5]; ++nn)
        for (int m = 0; m < n[4]; ++m)
            for (int l = 0; l < n[3]; ++l) {
                double ev = exp(-0.5 * (pow(eta[3][l] / vth[0], 2) + pow(eta[4][m] / vth[1], 2) + pow(eta[5][nn] / vth[2], 2)));
                double *blk = f + (((long)nn * n[4] + m) * n[3] + l) * nx3;
                for (int k = 0; k < n[2]; ++k) for (int j = 0; j < n[1]; ++j) for (int i = 0; i < n[0]; ++i)
                    blk[((long)k * n[1] + j) * n[0] + i] =
                        factor * (1.0 + alpha * (cos(kx[0] * eta[0][i]) * cos(kx[1] * eta[1][j]) * cos(kx[2] * eta[2][k]))) * ev;
            }
    double Lx = eta_max[0] - eta_min[0], Ly = eta_max[1] - eta_min[1], Lz = eta_max[2] - eta_min[2];
    double volume = de[0] * de[1] * de[2] * de[3] * de[4] * de[5];
    double vol_x = Lx * Ly * Lz;
    double dV = volume / vol_x, dVx = (de[0] * de[1] * de[2]) / vol_x;
    double vol_v = de[3] * de[4] * de[5];
    double *dispx[3];
    for (int d = 0; d < 3; ++d) {
        dispx[d] = (double *)malloc(sizeof(double) * n[d + 3]);
        for (int l = 0; l < n[d + 3]; ++l) dispx[d][l] = -eta[d + 3][l] * delta_t / de[d];
    }
    double *dfield = (double *)malloc(sizeof(double) * nx3);
    double vmin3[3] = {eta_min[3], eta_min[4], eta_min[5]}, dv3[3] = {de[3], de[4], de[5]};
#define FIELDS() do { orc_charge_density_6d(f, n, vol_v, rho); \
        orc_poisson_3d_periodic_solve(n[0], n[1], n[2], Lx, Ly, Lz, rho, phi, ex, ey, ez); } while (0)
#define ADVECT_V(dtv) do { const double *E3[3] = {ex, ey, ez}; \
        for (int d = 0; d < 3; ++d) { for (long i = 0; i < nx3; ++i) dfield[i] = E3[d][i] * (dtv) / de[3 + d]; \
            if (advector == 2) { long inn = nx3, out = 1; for (int a = 3; a < 3 + d; ++a) inn *= n[a]; \
                for (int a = 4 + d; a < 6; ++a) out *= n[a]; \
                if (orc_spline_dd_advect_axis(f, out, n[3 + d], inn, vblk[d], dfield, NULL, 1, 1, 0, 1, nx3, 1)) return -2; } \
            else advect6d_axis(f, n, 3 + d, stencil_v, dfield, 1); } } while (0)
    FIELDS();
    rows[0] = 0.0;
    diagnostics6d(f, n, vmin3, dv3, dV, dVx, rho, phi, ex, ey, ez, rows + 1);
    ADVECT_V(0.5 * delta_t);
    for (int it = 1; it <= nsteps; ++it) {
        for (int d = 0; d < 3; ++d) {
            if (advector == 0) advect6d_axis(f, n, d, stencil_x, dispx[d], 0);
            else advect6d_x_blocks(f, n, d, advector, stencil_x, dispx[d]);
        }
        FIELDS();
        rows[14 * it] = (double)it * delta_t;
        diagnostics6d(f, n, vmin3, dv3, dV, dVx, rho, phi, ex, ey, ez, rows + 14 * it + 1);
        if (time_in_phase && it == nsteps) ADVECT_V(0.5 * delta_t); else ADVECT_V(delta_t);
    }
    if (f_out) memcpy(f_out, f, sizeof(double) * ntot);
    for (int d = 0; d < 6; ++d) free(eta[d]);
    for (int d = 0; d < 3; ++d) free(dispx[d]);
    free(dfield); free(f); free(rho);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* 2D2V simulation sim_bsl_vp_2d2v_cart_poisson_serial, single rank, Strang,    */
/* periodic advectors (a2), trapezoid rho (a14), serial 2D Poisson (a15).       */
/* simulations/parallel/bsl_vp_2d2v_cart_poisson_serial/sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:650-1364 */
/* Arrays carry the duplicated periodic end point like the reference (nc+1).   */
/* method: 0 direct cubic spline (== sll_p_spline order 4 to ~4e-15),           */
/*         1 = FFT sll_p_spline(order), 2 = FFT sll_p_lagrange(order)           */
/* split: 0 = Strang VTV {.5,1,.5}, 1 = Strang TVT, 2 = Lie TV                  */
/* rows: per diagnostic (nsteps+1 rows) x 6: time, nrj, ekin, int f, int |f|, int f^2 */
/* ------------------------------------------------------------------------- */
static void advect_dup_line(int method, int order, int nc, double xmin, double xmax, double A, double dt,
                            double *line /* nc+1 */, double *scratch) {
    if (method == 0) {
        /* direct spline: out(x) = S(x - A dt) */
        double *C = scratch, *o = scratch + nc + 4;
        orc_spline_compute_interpolant_periodic(line, nc + 1, -1, C);
        orc_spline_eval_disp_periodic(C, nc + 1, xmin, xmax, -A * dt, o);
        memcpy(line, o, sizeof(double) * (nc + 1));
    } else {
        double *o = scratch;
        orc_advect_1d_periodic_constant(method - 1, nc, xmin, xmax, order, A, dt, line, o, nc + 1);
        memcpy(line, o, sizeof(double) * (nc + 1));
    }
}

int orc_splitting_coeff(int split, double dt, double *s, int *nb_split_step, int *split_begin_T, int *dim_split_V);
void orc_compute_jacobian(const double *E1, const double *E2, int nc1, int nc2, double factor, int r, int s, double *jac);
int orc_sim4d_run_ex(const int nc[4], const double xmin[4], const double xmax[4], double kx1, double kx2,
                     double eps, double dt, int nsteps, int split, int method, int order, double *rows,
                     double *f_out, int stencil_r, int stencil_s, double *thdiag);
int orc_sim4d_run(const int nc[4], const double xmin[4], const double xmax[4], double kx1, double kx2,
                  double eps, double dt, int nsteps, int split, int method, int order, double *rows,
                  double *f_out) {
    return orc_sim4d_run_ex(nc, xmin, xmax, kx1, kx2, eps, dt, nsteps, split, method, order, rows, f_out, -2, 2, NULL);
}
/* split: case numbering of sll_oracle_split.c (0 Strang VTV ... 17); stencil_r/s: finite-difference stencil of
 * compute_jacobian (namelist defaults -2, 2, :366-367); thdiag (may be NULL): (nsteps+1) x 13, the rows of the
 * reference's thdiag file (:998-1010 at t = 0 with the analytic mass0 / l20 of SLL_LANDAU :449-452, :1262-1275 later) */
int orc_sim4d_run_ex2(const int nc[4], const double xmin[4], const double xmax[4], double kx1, double kx2,
                      double eps, double dt, int nsteps, int split, int method, int order, double *rows,
                      double *f_out, int stencil_r, int stencil_s, double *thdiag, int cells_only, double *fields_out);
int orc_sim4d_run_ex(const int nc[4], const double xmin[4], const double xmax[4], double kx1, double kx2,
                     double eps, double dt, int nsteps, int split, int method, int order, double *rows,
                     double *f_out, int stencil_r, int stencil_s, double *thdiag) {
    return orc_sim4d_run_ex2(nc, xmin, xmax, kx1, kx2, eps, dt, nsteps, split, method, order, rows, f_out, stencil_r, stencil_s,
                             thdiag, 0, NULL);
}
/* cells_only = 0: the reference as it is.  The arrays carry the duplicated end point nc+1 in every direction; in a T
 * stage the planes v = v_min and v = v_max (the same periodic cell) are moved with OPPOSITE velocities (:1037-1064) and
 * both enter the trapezoid rho with weight 1/2 (sll_m_reduction.F90:229-272); the next V stage overwrites the v_max plane
 * with the v_min one again (out(N+1) = out(1), sll_m_advection_1d_periodic.F90:126-128).
 * cells_only = 1: the SAME code, but after every T stage the v_max planes are reset to copies of the v_min planes, i.e.
 * the state is a function of the periodic cells alone -- what a code that stores N points per direction computes.  The
 * difference between the two modes is that end-plane term and nothing else (tests/test_gpu_baseline_sizes.py). */
int orc_sim4d_run_ex2(const int nc[4], const double xmin[4], const double xmax[4], double kx1, double kx2,
                      double eps, double dt, int nsteps, int split, int method, int order, double *rows,
                      double *f_out, int stencil_r, int stencil_s, double *thdiag, int cells_only, double *fields_out) {
    int np[4]; long ntot = 1;
    double delta[4];
    for (int d = 0; d < 4; ++d) { np[d] = nc[d] + 1; ntot *= np[d]; delta[d] = (xmax[d] - xmin[d]) / (double)nc[d]; }
    double *f = (double *)malloc(sizeof(double) * ntot);
    long n12 = (long)np[0] * np[1];
    double *rho = (double *)malloc(sizeof(double) * n12 * 6);
    double *E1 = rho + n12, *E2 = E1 + n12, *jacE = E2 + n12, *K1 = jacE + n12, *K2 = K1 + n12;
    if (!f || !rho) return -1;
#define F4D(a, b, c, d) f[(a) + (long)np[0] * ((b) + (long)np[1] * ((c) + (long)np[2] * (d)))]
    /* sll_f_landau_mode_initializer_4d, sll_m_common_array_initializers.F90:948-993 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int i4 = 0; i4 < np[3]; ++i4) for (int i3 = 0; i3 < np[2]; ++i3) {
        double vx = xmin[2] + i3 * delta[2], vy = xmin[3] + i4 * delta[3];
        for (int i2 = 0; i2 < np[1]; ++i2) for (int i1 = 0; i1 < np[0]; ++i1) {
            double x = xmin[0] + i1 * delta[0], y = xmin[1] + i2 * delta[1];
            double factor1 = 1.0 + eps * cos(kx1 * x) * cos(kx2 * y);
            F4D(i1, i2, i3, i4) = (1.0 / (2.0 * ORC_PI)) * factor1 * exp(-0.5 * (vx * vx + vy * vy));
        }
    }
    double steps[32]; int nsub; int beginT; int dimV;
    if (orc_splitting_coeff(split, dt, steps, &nsub, &beginT, &dimV)) return -2;
    double nrj = 0, nrj_jac = 0, jac_max = 0;
    /* fields of a V stage (:1110-1126): E from rho, jacobian_E from E, and with dim_split_V == 2 the second field pair
     * from the Poisson solve of jacobian_E */
#define FIELD4() do { orc_reduction_4d_to_2d_direction34(f, np[0], np[1], np[2], np[3], delta[2], delta[3], rho); \
        orc_poisson_2d_periodic_solve_e(nc[0], nc[1], xmin[0], xmax[0], xmin[1], xmax[1], rho, np[0], np[1], E1, E2, NULL); \
        nrj = 0; for (long i = 0; i < n12; ++i) nrj += E1[i] * E1[i] + E2[i] * E2[i]; nrj *= delta[0] * delta[1]; \
        orc_compute_jacobian(E1, E2, nc[0], nc[1], 4.0 / (delta[0] * delta[1]), stencil_r, stencil_s, jacE); \
        jac_max = 0; for (long i = 0; i < n12; ++i) if (fabs(jacE[i]) > jac_max) jac_max = fabs(jacE[i]); \
        nrj_jac = 0; \
        if (dimV == 2) { orc_poisson_2d_periodic_solve_e(nc[0], nc[1], xmin[0], xmax[0], xmin[1], xmax[1], jacE, np[0], np[1], K1, K2, NULL); \
            for (long i = 0; i < n12; ++i) nrj_jac += K1[i] * K1[i] + K2[i] * K2[i]; nrj_jac *= delta[0] * delta[1]; } } while (0)
    FIELD4();
    if (dimV == 2) { /* the t = 0 row sums field_x2 unsquared, twice (:905) */
        nrj_jac = 0; for (long i = 0; i < n12; ++i) nrj_jac += K1[i] * K1[i] + K2[i] + K2[i]; nrj_jac *= delta[0] * delta[1]; }
    /* row 0 (:983-997): time, nrj, ekin(analytic = L1*L2), mass0... we emit the same
     * three integrals as later rows, computed, for a uniform table */
    int maxnp = 0; for (int d = 0; d < 4; ++d) if (np[d] > maxnp) maxnp = np[d];
    for (int it = 0; it <= nsteps; ++it) {
        if (it > 0) {
            int isub = 0, T = beginT;
            for (int ss = 0; ss < nsub; ++ss) {
                if (T) {
                    isub += 1;
                    double st = steps[isub - 1];
#pragma omp parallel
                    {
                        double *line = (double *)malloc(sizeof(double) * (4 * maxnp + 32));
                        double *scr = line + maxnp + 1;
#pragma omp for collapse(2) schedule(static)
                        for (int i4 = 0; i4 < np[3]; ++i4) for (int i3 = 0; i3 < np[2]; ++i3) {
                            double a1 = (xmin[2] + (double)i3 * delta[2]) * st;
                            for (int i2 = 0; i2 < np[1]; ++i2) {
                                for (int i = 0; i < np[0]; ++i) line[i] = F4D(i, i2, i3, i4);
                                advect_dup_line(method, order, nc[0], xmin[0], xmax[0], a1, dt, line, scr);
                                for (int i = 0; i < np[0]; ++i) F4D(i, i2, i3, i4) = line[i];
                            }
                            double a2 = (xmin[3] + (double)i4 * delta[3]) * st;
                            for (int i1 = 0; i1 < np[0]; ++i1) {
                                for (int i = 0; i < np[1]; ++i) line[i] = F4D(i1, i, i3, i4);
                                advect_dup_line(method, order, nc[1], xmin[1], xmax[1], a2, dt, line, scr);
                                for (int i = 0; i < np[1]; ++i) F4D(i1, i, i3, i4) = line[i];
                            }
                        }
                        free(line);
                    }
                    if (cells_only) {
#pragma omp parallel for schedule(static)
                        for (int i4 = 0; i4 < np[3]; ++i4) {
                            for (long k = 0; k < n12; ++k) f[k + n12 * ((long)nc[2] + (long)np[2] * i4)] = f[k + n12 * ((long)np[2] * i4)];
                        }
#pragma omp parallel for schedule(static)
                        for (int i3 = 0; i3 < np[2]; ++i3) {
                            for (long k = 0; k < n12; ++k) f[k + n12 * (i3 + (long)np[2] * nc[3])] = f[k + n12 * i3];
                        }
                    }
                } else {
                    FIELD4();
                    double st = steps[isub], st2 = (dimV == 2) ? steps[isub + 1] : 0.0;
#pragma omp parallel
                    {
                        double *line = (double *)malloc(sizeof(double) * (4 * maxnp + 32));
                        double *scr = line + maxnp + 1;
#pragma omp for collapse(2) schedule(static)
                        for (int i2 = 0; i2 < np[1]; ++i2) for (int i1 = 0; i1 < np[0]; ++i1) {
                            double a3 = 0.0; a3 = a3 + E1[i1 + (long)np[0] * i2] * st;
                            if (dimV == 2) a3 = a3 + K1[i1 + (long)np[0] * i2] * st2;
                            for (int i4 = 0; i4 < np[3]; ++i4) {
                                for (int i = 0; i < np[2]; ++i) line[i] = F4D(i1, i2, i, i4);
                                advect_dup_line(method, order, nc[2], xmin[2], xmax[2], a3, dt, line, scr);
                                for (int i = 0; i < np[2]; ++i) F4D(i1, i2, i, i4) = line[i];
                            }
                            double a4 = 0.0; a4 = a4 + E2[i1 + (long)np[0] * i2] * st;
                            if (dimV == 2) a4 = a4 + K2[i1 + (long)np[0] * i2] * st2;
                            for (int i3 = 0; i3 < np[2]; ++i3) {
                                for (int i = 0; i < np[3]; ++i) line[i] = F4D(i1, i2, i3, i);
                                advect_dup_line(method, order, nc[3], xmin[3], xmax[3], a4, dt, line, scr);
                                for (int i = 0; i < np[3]; ++i) F4D(i1, i2, i3, i) = line[i];
                            }
                        }
                        free(line);
                    }
                    isub += dimV;
                }
                T = !T;
            }
        }
        /* diagnostics (:1181-1275): trapezoid in all four directions
         * (sll_m_reduction.F90:364-476,486-557) */
        double ekin = 0, i0 = 0, i1n = 0, i2n = 0;
        for (int i4 = 0; i4 < np[3]; ++i4) for (int i3 = 0; i3 < np[2]; ++i3) {
            double s0 = 0, s1 = 0, s2 = 0;
            for (int i1 = 0; i1 < np[0]; ++i1) {
                double w1 = (i1 == 0 || i1 == np[0] - 1) ? 0.5 : 1.0;
                double t0 = 0.5 * (F4D(i1, 0, i3, i4) + F4D(i1, np[1] - 1, i3, i4));
                double t1 = 0.5 * (fabs(F4D(i1, 0, i3, i4)) + fabs(F4D(i1, np[1] - 1, i3, i4)));
                double t2 = 0.5 * (F4D(i1, 0, i3, i4) * F4D(i1, 0, i3, i4) + F4D(i1, np[1] - 1, i3, i4) * F4D(i1, np[1] - 1, i3, i4));
                for (int i2 = 1; i2 < np[1] - 1; ++i2) { double v = F4D(i1, i2, i3, i4); t0 += v; t1 += fabs(v); t2 += v * v; }
                s0 += w1 * t0 * delta[1]; s1 += w1 * t1 * delta[1]; s2 += w1 * t2 * delta[1];
            }
            s0 *= delta[0]; s1 *= delta[0]; s2 *= delta[0];
            double w = ((i3 == 0 || i3 == np[2] - 1) ? 0.5 : 1.0) * ((i4 == 0 || i4 == np[3] - 1) ? 0.5 : 1.0) * delta[2] * delta[3];
            double v3 = xmin[2] + i3 * delta[2], v4 = xmin[3] + i4 * delta[3];
            ekin += w * s0 * 0.5 * (v4 * v4 + v3 * v3);
            i0 += w * s0; i1n += w * s1; i2n += w * s2;
        }
        double *r = rows + 6 * it;
        r[0] = it * dt; r[1] = nrj; r[2] = ekin; r[3] = i0; r[4] = i1n; r[5] = i2n;
        if (thdiag) {
            double *t = thdiag + 13 * it;
            const double mass0 = (xmax[0] - xmin[0]) * (xmax[1] - xmin[1]);
            const double nrj0 = (0.5 * eps * ORC_PI) * (0.5 * eps * ORC_PI) / (kx1 * kx2) * (1.0 / (kx1 * kx1) + 1.0 / (kx2 * kx2));
            double l20 = (2.0 * ORC_PI / kx1) * (2.0 * ORC_PI / kx2) * 0.25;
            l20 = l20 * (1.0 + 0.25 * (eps * eps)) / ORC_PI;
            t[0] = it * dt; t[1] = nrj; t[2] = (it == 0) ? mass0 : ekin; t[3] = nrj0; t[4] = mass0; t[5] = jac_max; t[6] = nrj_jac;
            if (it == 0) { t[7] = mass0; t[8] = mass0; t[9] = l20; }
            else { t[7] = i0; t[8] = i1n; t[9] = i2n; }
            t[10] = mass0; t[11] = mass0; t[12] = l20;
        }
    }
    if (f_out) memcpy(f_out, f, sizeof(double) * ntot);
    /* rho, E1, E2 of the last field solve ((nc1+1) x (nc2+1) each, duplicated end points included) */
    if (fields_out) memcpy(fields_out, rho, sizeof(double) * 3 * n12);
    free(f); free(rho);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* 1D1V simulation sim_bsl_vp_1d1v_cart (single rank, no drive), Strang VTV.    */
/* simulations/parallel/bsl_vp_1d1v_cart/sll_m_sim_bsl_vp_1d1v_cart.F90:1066-1907 */
/* method as in orc_sim4d_run, plus 3 = Lagrange fixed periodic-last stencil    */
/* `order` through sll_t_lagrange_interpolator_1d semantics: out(x)=f(x - A dt).*/
/* init: 0 = Landau (:507-540), 1 = two-stream (:566-595)                       */
/* rows: (nsteps) x 8: time, mass, l1, momentum, l2, ekin, epot, etot (:1783)   */
/* ------------------------------------------------------------------------- */
int orc_sim2d_run(int nc_x1, int nc_x2, double x1_min, double x1_max, double x2_min, double x2_max,
                  int init, double kmode, double eps, double dt, int nsteps, int method, int order,
                  double *rows, double *f_out, double *efield_out) {
    int np1 = nc_x1 + 1, np2 = nc_x2 + 1;
    double d1 = (x1_max - x1_min) / nc_x1, d2 = (x2_max - x2_min) / nc_x2;
    double *f = (double *)malloc(sizeof(double) * np1 * np2);
    double *w = (double *)malloc(sizeof(double) * np2);
    double *x2 = (double *)malloc(sizeof(double) * np2);
    double *rho = (double *)malloc(sizeof(double) * np1 * 2);
    double *E = rho + np1;
    int maxnp = np1 > np2 ? np1 : np2;
    double *line = (double *)malloc(sizeof(double) * (5 * maxnp + 32));
    double *scr = line + maxnp + 1;
    for (int j = 0; j < np2; ++j) x2[j] = x2_min + j * d2;
    /* SLL_TRAPEZOID weights :937-943 */
    w[0] = 0.5 * (x2[1] - x2[0]);
    for (int j = 1; j < nc_x2; ++j) w[j] = 0.5 * (x2[j + 1] - x2[j - 1]);
    w[nc_x2] = 0.5 * (x2[nc_x2] - x2[nc_x2 - 1]);
    for (int j = 0; j < np2; ++j) for (int i = 0; i < np1; ++i) {
        double x = x1_min + i * d1, v = x2[j];
        double fac = 1.0 / sqrt(2.0 * ORC_PI);
        f[i + (long)np1 * j] = init == 0 ? fac * (1.0 + eps * cos(kmode * x)) * exp(-0.5 * v * v)
                                         : fac * (1.0 + eps * cos(kmode * x)) * v * v * exp(-0.5 * v * v);
    }
#define FIELD2() do { for (int i = 0; i < np1; ++i) { double s = 0; for (int j = 0; j < np2; ++j) s += f[i + (long)np1 * j] * w[j]; rho[i] = 1.0 - s; } \
        orc_poisson_1d_periodic_solve(nc_x1, x1_min, x1_max, rho, E); } while (0)
    FIELD2();
    double steps[3] = {0.5, 1.0, 0.5};
    for (int it = 1; it <= nsteps; ++it) {
        int T = 0;
        for (int ss = 0; ss < 3; ++ss) {
            if (T) {
                for (int j = 0; j < np2; ++j) {
                    double A = x2[j] * steps[ss];
                    for (int i = 0; i < np1; ++i) line[i] = f[i + (long)np1 * j];
                    if (method == 3) { orc_lagrange_fixed_periodicl(line, scr, np1, -A * dt / d1, order); memcpy(line, scr, sizeof(double) * np1); }
                    else advect_dup_line(method, order, nc_x1, x1_min, x1_max, A, dt, line, scr);
                    for (int i = 0; i < np1; ++i) f[i + (long)np1 * j] = line[i];
                }
                FIELD2();
            } else {
                for (int i = 0; i < np1; ++i) {
                    double A = -E[i] * steps[ss];
                    for (int j = 0; j < np2; ++j) line[j] = f[i + (long)np1 * j];
                    if (method == 3) { orc_lagrange_fixed_periodicl(line, scr, np2, -A * dt / d2, order); memcpy(line, scr, sizeof(double) * np2); }
                    else advect_dup_line(method, order, nc_x2, x2_min, x2_max, A, dt, line, scr);
                    for (int j = 0; j < np2; ++j) f[i + (long)np1 * j] = line[j];
                }
            }
            T = !T;
        }
        double t[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < np1 - 1; ++i) for (int j = 0; j < np2; ++j) {
            double v = f[i + (long)np1 * j];
            t[0] += v * w[j]; t[1] += fabs(v) * w[j]; t[2] += v * v * w[j];
            t[3] += v * x2[j] * w[j]; t[4] += v * x2[j] * x2[j] * w[j];
        }
        double epot = 0; for (int i = 0; i < np1 - 1; ++i) epot += E[i] * E[i];
        epot = 0.5 * epot * d1;
        double *r = rows + 8 * (it - 1);
        r[0] = it * dt; r[1] = t[0] * d1; r[2] = t[1] * d1; r[3] = t[3] * d1; r[4] = t[2] * d1;
        r[5] = 0.5 * t[4] * d1; r[6] = epot; r[7] = r[5] + r[6];
    }
    if (f_out) memcpy(f_out, f, sizeof(double) * np1 * np2);
    if (efield_out) memcpy(efield_out, E, sizeof(double) * np1);
    free(f); free(w); free(x2); free(rho); free(line);
    return 0;
}
