/*
 * sll_oracle_halo.c -- CPU restatement of SeLaLib's LOCAL cubic-spline advection with halo cells
 *  and of the centred variable-block Lagrange x-advection (rank 2).
 *
 * TEST INFRASTRUCTURE ONLY (same rule as sll_oracle.c): loaded by tests/, __graft_entry__.smoke() and
 * bench.py's CPU legs; never by selalib_b200.
 *
 * Restates, routine by routine and in the same operation order:
 *   src/interpolation/interpolators/sll_m_cubic_spline_halo_1d.F90:69-252   (NUM_TERMS = 15, :11)
 *   src/semi_lagrangian/advection/sll_m_advection_6d_spline_dd_slim.F90:202-287 (make_blocks_spline),
 *       :291-515 (fadvect_eta1; eta2, eta3 are the same with another axis), :976-1203 (advect_eta4; eta5, eta6 ditto)
 *   src/parallelization/decomposition/sll_m_decomposition.F90:2260-2289 (bc exchange), :1715-2030 (halo exchange)
 *   src/semi_lagrangian/advection/sll_m_advection_6d_lagrange_dd_slim.F90:173-286 (set_eta123, make_blocks_lagrange),
 *       :291-467 (fadvect_eta1 + core)
 * Pinned by the reference's own known-answer test src/interpolation/interpolators/testing/
 * test_cubic_spline_halo_1d.F90 (n = 64, alpha = 0.25, si = -2..2, tolerance 4e-9), restated in
 * tests/test_oracle_halo.py.
 *
 * A decomposed run is emulated in one process: a line of n points is cut into nblk pieces of np = n/nblk
 * points ("ranks" in a periodic ring along that axis); every piece runs the reference's sequence
 * prepare_exchange -> exchange -> finish_boundary_conditions -> compute_interpolant -> eval_disp with the
 * boundary scalars and halo cells taken from its ring neighbours.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define HALO_NUM_TERMS 15 /* sll_m_cubic_spline_halo_1d.F90:11 */

static double p_a, p_r_a, p_b, p_b_a, p_sqrt3;
static double pba_pow[28];
static int consts_ready = 0;
static void halo_consts(void) {
    if (consts_ready) return;
    p_a = sqrt((2.0 + sqrt(3.0)) / 6.0);   /* :50-55 */
    p_r_a = 1.0 / p_a;
    p_b = sqrt((2.0 - sqrt(3.0)) / 6.0);
    p_b_a = p_b / p_a;
    p_sqrt3 = sqrt(3.0);
    for (int i = 0; i < 28; ++i) pba_pow[i] = pow(-p_b_a, (double)i); /* :59-64 */
    consts_ready = 1;
}

/* sll_s_cubic_spline_halo_1d_prepare_exchange (:70-104).  fdata is 1-based in the reference: F(k) = fdata[k-1]. */
void orc_halo_prepare_exchange(const double *fdata, int si, int np, double *d_0, double *c_np2) {
    halo_consts();
#define F(k) fdata[(k) - 1]
    int i, ind_min;
    if (si > 0) { *d_0 = 0.0; ind_min = si; }
    else { *d_0 = F(np + si); ind_min = 1; }
    for (i = ind_min; i <= HALO_NUM_TERMS; ++i) *d_0 += pba_pow[i] * F(np + si - i);
    if (si < -1) { *c_np2 = 0.0; ind_min = -si - 1; }
    else { *c_np2 = F(2 + si); ind_min = 1; }
    for (i = ind_min; i <= HALO_NUM_TERMS; ++i) *c_np2 += pba_pow[i] * F(2 + si + i);
    for (i = 1; i <= si + 1; ++i) *c_np2 += pba_pow[i] * F(2 + si - i);
#undef F
}

/* sll_s_cubic_spline_halo_1d_finish_boundary_conditions (:108-137) */
void orc_halo_finish_boundary_conditions(const double *fdata, int si, int np, double *d_0, double *c_np2) {
    halo_consts();
#define F(k) fdata[(k) - 1]
    int i, ind_min;
    if (si > 0) *d_0 += F(si);
    for (i = 1; i <= si - 1; ++i) *d_0 += pba_pow[i] * F(si - i);
    *d_0 *= p_r_a;
    if (si < -1) { *c_np2 += F(np + 2 + si); ind_min = 1; }
    else ind_min = si + 2;
    for (i = 1; i <= -si - 2; ++i) *c_np2 += pba_pow[i] * F(np + 2 + si + i);
    for (i = ind_min; i <= HALO_NUM_TERMS; ++i) *c_np2 += pba_pow[i] * F(np + 2 + si - i);
    *c_np2 *= p_sqrt3;
#undef F
}

/* sll_s_cubic_spline_halo_1d_compute_interpolant (:141-172): f(0:np+2), d(0:np+1), coeffs(0:np+2); coeffs may
 * alias f, as at the call sites (sll_m_advection_6d_spline_dd_slim.F90:438-441). */
void orc_halo_compute_interpolant(const double *f, int np, double *d, double *coeffs) {
    halo_consts();
    d[0] = f[0];
    for (int i = 1; i <= np + 1; ++i) d[i] = p_r_a * (f[i] - p_b * d[i - 1]);
    coeffs[np + 2] = f[np + 2];
    for (int i = np + 1; i >= 0; --i) coeffs[i] = p_r_a * (d[i] - p_b * coeffs[i + 1]);
}

/* sll_s_cubic_spline_halo_1d_eval_disp (:176-200): coeffs(0:np+2), alpha in [0,1], fout(1:np) */
void orc_halo_eval_disp(const double *coeffs, double alpha, int np, double *fout) {
    const double calpha = 1.0 - alpha, inv6 = 1.0 / 6.0;
    for (int cell = 1; cell <= np; ++cell) {
        double cim1 = coeffs[cell - 1], ci = coeffs[cell], cip1 = coeffs[cell + 1], cip2 = coeffs[cell + 2];
        double t1 = 3.0 * ci, t3 = 3.0 * cip1;
        double t2 = calpha * (calpha * (calpha * (cim1 - t1) + t1) + t1) + ci;
        double t4 = alpha * (alpha * (alpha * (cip2 - t3) + t3) + t3) + cip1;
        fout[cell - 1] = inv6 * (t2 + t4);
    }
}

/* sll_s_cubic_spline_halo_1d_periodic (:203-252): fin(0:nc+2) holds the nc data values in fin(0:nc-1) on entry
 * and the coefficients on exit; fout(1:nc) doubles as the work array d(0:) */
void orc_halo_periodic(double *fin, double alpha, int nc, double *fout) {
    halo_consts();
    double *d = fout, *fc = fin;
    const int np = nc;
    d[0] = fc[0];
    for (int i = 1; i <= HALO_NUM_TERMS; ++i) d[0] += pba_pow[i] * fc[np - i];
    d[0] *= p_r_a;
    for (int i = 1; i <= np - 1; ++i) d[i] = p_r_a * (fc[i] - p_b * d[i - 1]);
    fc[np] = d[np - 1];
    for (int i = 1; i <= HALO_NUM_TERMS; ++i) fc[np] += d[i - 1] * pba_pow[i];
    fc[np] *= p_r_a;
    for (int i = np - 1; i >= 1; --i) fc[i] = p_r_a * (d[i - 1] - p_b * fc[i + 1]);
    fc[0] = fc[np];
    fc[np + 1] = fc[1]; fc[np + 2] = fc[2];
    /* eval_disp writes fout while reading only fin */
    double *tmp = (double *)malloc(sizeof(double) * nc);
    orc_halo_eval_disp(fin, alpha, nc, tmp);
    memcpy(fout, tmp, sizeof(double) * nc);
    free(tmp);
}

/* make_blocks_spline (sll_m_advection_6d_spline_dd_slim.F90:202-287), literally: disp(1:n) monotonic.
 * Per index j (0-based out arrays): shift[j] = the block's integer displacement, or ORC_SKIP when the index
 * belongs to no block (abs(disp) == 0: the line is left untouched); alpha[j] = disp - floor(disp) (:281-285).
 * Returns the number of blocks. */
#define ORC_SKIP (-2147483647 - 1)
int orc_make_blocks_spline(int n, const double *disp_in, int *shift, double *alpha) {
    /* 1-based like the reference, index_range = [1, n] */
#define D(j) disp_in[(j) - 1]
    int bl, j, box1, box2, blocks, si;
    for (j = 1; j <= n; ++j) shift[j - 1] = ORC_SKIP;
    bl = 1;
    if (fabs(D(bl)) == 0.0) bl = bl + 1;
    box1 = (int)floor(D(bl));
    bl = n;
    if (fabs(D(bl)) == 0.0) bl = bl - 1;
    box2 = (int)floor(D(bl));
    blocks = abs(box2 - box1) + 1;
    if (box1 > box2) {
        j = 1;
        for (bl = 1; bl <= blocks; ++bl) {
            if (j <= n && fabs(D(j)) == 0.0) j = j + 1;
            si = box1 - bl + 1;
            int jstart = j, jend;
            while (j <= n && D(j) > (double)(box1 - bl + 1)) { j = j + 1; if (j > n) break; }
            if (j - 1 >= 1 && fabs(D(j - 1)) == 0.0) jend = j - 2; else jend = j - 1;
            for (int k = jstart; k <= jend; ++k) shift[k - 1] = si;
        }
    } else {
        j = 1;
        for (bl = box1; bl <= box2; ++bl) {
            if (j <= n && fabs(D(j)) == 0.0) j = j + 1;
            int jstart = j, jend;
            while (j <= n && D(j) < (double)(bl + 1)) { j = j + 1; if (j > n) break; }
            if (j - 1 >= 1 && fabs(D(j - 1)) == 0.0) jend = j - 2; else jend = j - 1;
            for (int k = jstart; k <= jend; ++k) shift[k - 1] = bl;
        }
    }
    for (j = 1; j <= n; ++j) alpha[j - 1] = D(j) - floor(D(j));
#undef D
    return blocks;
}

/* One line of n = nblk*np points, emulating nblk ring ranks.  si/alpha as the advector holds them
 * (idisplacement, normalised displacement).  Sequence per rank r (fadvect_eta1, :330-450):
 *   every rank: prepare_exchange on its own data -> (d0, c_np2)                       (:343-346)
 *   bc exchange: my bc_left  (d_0 start)   <- the d0    computed by my LEFT  neighbour,
 *                my bc_right (c_np2 start) <- the c_np2 computed by my RIGHT neighbour
 *     (sll_m_decomposition.F90:2274-2288; the two Sendrecv calls pair neighbours so that this holds for rings of
 *      one or two ranks per axis, which is all sll_f_set_process_grid produces up to 64 ranks; the adjacency used
 *      here is the one the mathematics needs)
 *   halo exchange with widths (-si, si+1): only one side is non-empty                 (:380-386)
 *   finish_boundary_conditions on my data                                             (:404-409)
 *   buf_i = [d_0 | g(si+1 .. si+np+1) | c_np2], compute_interpolant, eval_disp        (:412-448)
 */
static void halo_line(const double *lin, double *lout, int n, int nblk, int si, double alpha, double *work) {
    const int np = n / nblk;
    double *d0s = work, *c2s = work + nblk, *buf = work + 2 * nblk, *d = buf + (np + 3), *o = d + (np + 3);
    for (int r = 0; r < nblk; ++r) orc_halo_prepare_exchange(lin + (long)r * np, si, np, &d0s[r], &c2s[r]);
    for (int r = 0; r < nblk; ++r) {
        const double *mine = lin + (long)r * np;
        double d_0 = d0s[(r + nblk - 1) % nblk], c_np2 = c2s[(r + 1) % nblk];
        orc_halo_finish_boundary_conditions(mine, si, np, &d_0, &c_np2);
        buf[0] = d_0;
        for (int j = 1; j <= np + 1; ++j) { /* window g(si + j), g 1-based local, periodic over the whole line */
            long k = (long)r * np + (si + j - 1);
            k %= n; if (k < 0) k += n;
            buf[j] = lin[k];
        }
        buf[np + 2] = c_np2;
        orc_halo_compute_interpolant(buf, np, d, buf);
        orc_halo_eval_disp(buf, alpha, np, o);
        memcpy(lout + (long)r * np, o, sizeof(double) * np);
    }
}

/* Whole-array pass along one axis of f viewed as [outer][n][inner] (inner fastest), n cut into nblk ring pieces.
 * disp of line (o, in) = dvals[idx], shift = shifts[idx] (NULL: floor(disp), the eta4..6 rule
 * sll_m_advection_6d_spline_dd_slim.F90:1098-1102), idx = ((o/odiv)%omod)*ostr + ((in/idiv)%imod)*istr.
 * Lines whose shift is ORC_SKIP are left untouched (they belong to no block). */
int orc_spline_dd_advect_axis(double *f, long outer, int n, long inner, int nblk, const double *dvals, const int *shifts,
                              long odiv, long omod, long ostr, long idiv, long imodn, long istr) {
    if (nblk < 1 || n % nblk != 0 || n / nblk <= HALO_NUM_TERMS) return -1; /* SLL_ASSERT_ALWAYS(num_points > NUM_TERMS) */
    halo_consts();
    const int np = n / nblk;
    int bad = 0;
#pragma omp parallel
    {
        double *lin = (double *)malloc(sizeof(double) * (2 * (size_t)n + 2 * nblk + 3 * (np + 3) + 8));
        double *lout = lin + n, *work = lout + n;
#pragma omp for schedule(static) collapse(2)
        for (long o = 0; o < outer; ++o)
            for (long in = 0; in < inner; ++in) {
                double *base = f + o * (long)n * inner + in;
                long idx = ((o / odiv) % omod) * ostr + ((in / idiv) % imodn) * istr;
                double dc = dvals[idx];
                int si = shifts ? shifts[idx] : (int)floor(dc);
                if (si == ORC_SKIP) continue;
                /* the neighbour sums must stay inside the neighbour's piece: prepare_exchange reads fdata(np+si-15)
                 * and fdata(2+si+15) (:83-103); the reference only asserts np > 15 and reads out of bounds beyond
                 * that, so such lines are refused here instead of being compared against garbage */
                if (np < 16 - si || np < 17 + si) { bad = 1; continue; }
                double alpha = dc - floor(dc);
                for (int i = 0; i < n; ++i) lin[i] = base[(long)i * inner];
                halo_line(lin, lout, n, nblk, si, alpha, work);
                for (int i = 0; i < n; ++i) base[(long)i * inner] = lout[i];
            }
        free(lin);
    }
    return bad ? -3 : 0;
}
