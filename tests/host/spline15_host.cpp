// Host emulation of K9 (selalib_b200/csrc/sllb_spline15.cuh compiled with g++): TEST INFRASTRUCTURE.
// Emulates, line by line, what k_spline_dd_prepare + k_spline_dd_strided / k_spline_dd_contig do with the shared-memory
// tile (pitch 1 here), for an axis cut into nblk ring pieces, so the per-line arithmetic of the CUDA path can be checked
// against the oracle in the CPU test suite.  The GPU parity tests check the kernels themselves.
#define SLLB_HOST_EMULATION 1
#include "../../selalib_b200/csrc/sllb_spline15.cuh"
#include "../../selalib_b200/csrc/sllb_hermite.cuh"
#define __device__
#define __host__
#define __forceinline__ inline
#include "../../selalib_b200/csrc/sllb_lagrange.cuh"

#include <climits>
#include <cstring>
#include <vector>

using namespace sllb;

static void init_consts() {
    static bool ready = false;
    if (ready) return;
    const double a = sqrt((2.0 + sqrt(3.0)) / 6.0), b = sqrt((2.0 - sqrt(3.0)) / 6.0);
    for (int i = 0; i <= SLLB_HALO_TERMS; ++i) c_hpw[i] = pow(-(b / a), (double)i);
    ready = true;
}

static void init_hermite_consts() {
    static bool ready = false;
    if (ready) return;
    const double a = sqrt((2.0 + sqrt(3.0)) / 6.0), b = sqrt((2.0 - sqrt(3.0)) / 6.0);
    double ct = 1.0;
    for (int i = 0; i < SLLB_HERMITE_TERMS; ++i) { c_hq[i] = ct; ct *= -(b / a); }
    ready = true;
}

// periodic Lagrange line exactly as the kernels form it: lagr_setup<S> (weights + first stencil offset), then the
// left-to-right sum of lagrange_line / k_lagrange_contig
template <int S>
static void lagr_line_t(const double *lin, double *lout, int n, double disp) {
    double pp[S];
    const int off = lagr_setup<S>(disp, n, pp);
    for (int i = 0; i < n; ++i) {
        int j = (i + off) % n;
        double acc = pp[0] * lin[j];
        for (int k = 1; k < S; ++k) { j = (j == n - 1) ? 0 : j + 1; acc = fma(pp[k], lin[j], acc); }
        lout[i] = acc;
    }
}

extern "C" {
// one line through hermite_coeffs_line + hermite_eval_point as k_hermite_strided runs them (pitch 1)
int emu_hermite_line(const double *lin, double *lout, int np, double delta, double alpha, int inplace, int have_slopes,
                     double sl, double sr) {
    init_hermite_consts();
    if (np < SLLB_HERMITE_TERMS) return -1;
    std::vector<double> tile(lin, lin + np);
    double g0, gnp1;
    hermite_coeffs_line<1>(tile.data(), np, delta, have_slopes, sl, sr, &g0, &gnp1);
    for (int i = 1; i <= np; ++i) lout[i - 1] = hermite_eval_point<1>(tile.data(), np, i, alpha / delta, inplace, g0, gnp1);
    return 0;
}
// line of n = nblk*np points; mode 0: strided kernel (TO_GLOBAL), 1: contiguous kernel (parked results + rotation).
// nblk == 1 -> WRAP kernels; nblk > 1 -> prepare + halo rows + non-WRAP kernel per piece.
int emu_spline_dd_line(const double *lin, double *lout, int n, int nblk, int si, double alpha, int hwl, int hwr, int mode) {
    init_consts();
    const int np = n / nblk;
    if (si == INT_MIN) { memcpy(lout, lin, sizeof(double) * n); return 0; }
    if (nblk == 1) {
        std::vector<double> tile(lin, lin + n);
        double sd, sc;
        spline15_sums<1, true>(tile.data(), n, si, 0.0, 0.0, &sd, &sc);
        if (mode == 0) spline15_line<1, true, true>(tile.data(), n, si, alpha, sd, sc, lout, 1);
        else {
            spline15_line<1, true, false>(tile.data(), n, si, alpha, sd, sc, nullptr, 0);
            const int r = wrap_idx(si, n);
            for (int i = 0; i < n; ++i) { int kc = i + r; if (kc >= n) kc -= n; lout[i] = tile[kc]; }
        }
        return 0;
    }
    if (mode != 0) return -1;
    if (si < -hwl || si > hwr - 1) return -2;
    std::vector<double> for_right(nblk), for_left(nblk);
    for (int r = 0; r < nblk; ++r) spline15_prepare(lin + (long)r * np, 1, np, si, &for_right[r], &for_left[r]);
    for (int r = 0; r < nblk; ++r) {
        const int left = (r + nblk - 1) % nblk, right = (r + 1) % nblk;
        std::vector<double> tile(hwl + np + hwr);
        for (int j = 0; j < hwl; ++j) tile[j] = lin[(long)left * np + np - hwl + j];       // last planes of the left neighbour
        for (int j = 0; j < np; ++j) tile[hwl + j] = lin[(long)r * np + j];
        for (int j = 0; j < hwr; ++j) tile[hwl + np + j] = lin[(long)right * np + j];      // first planes of the right neighbour
        double *x0 = tile.data() + hwl, sd, sc;
        spline15_sums<1, false>(x0, np, si, for_right[left], for_left[right], &sd, &sc);
        spline15_line<1, false, true>(x0, np, si, alpha, sd, sc, lout + (long)r * np, 1);
    }
    return 0;
}

int emu_lagrange_line(const double *lin, double *lout, int n, double disp, int stencil) {
    switch (stencil) {
    case 3: lagr_line_t<3>(lin, lout, n, disp); return 0;
    case 5: lagr_line_t<5>(lin, lout, n, disp); return 0;
    case 7: lagr_line_t<7>(lin, lout, n, disp); return 0;
    case 9: lagr_line_t<9>(lin, lout, n, disp); return 0;
    case 11: lagr_line_t<11>(lin, lout, n, disp); return 0;
    case 4: lagr_line_t<4>(lin, lout, n, disp); return 0;
    case 6: lagr_line_t<6>(lin, lout, n, disp); return 0;
    case 8: lagr_line_t<8>(lin, lout, n, disp); return 0;
    case 10: lagr_line_t<10>(lin, lout, n, disp); return 0;
    case 12: lagr_line_t<12>(lin, lout, n, disp); return 0;
    case 14: lagr_line_t<14>(lin, lout, n, disp); return 0;
    case 16: lagr_line_t<16>(lin, lout, n, disp); return 0;
    case 18: lagr_line_t<18>(lin, lout, n, disp); return 0;
    default: return -1;
    }
}

// the two halves of a split-axis pass as separate calls, for the multi-process (gloo) test of the exchange protocol:
// what k_spline_dd_prepare computes for the neighbours from MY piece ...
int emu_spline_dd_prepare(const double *local, int np, int si, double *for_right, double *for_left) {
    init_consts();
    spline15_prepare(local, 1, np, si, for_right, for_left);
    return 0;
}
// ... and what k_spline_dd_strided<false> does with my piece, the received halo cells and the received sums
int emu_spline_dd_piece(const double *local, const double *halo_l, int hwl, const double *halo_r, int hwr, int np, int si,
                        double alpha, double rem_d, double rem_c, double *out) {
    init_consts();
    if (si < -hwl || si > hwr - 1) return -2;
    std::vector<double> tile(hwl + np + hwr);
    for (int j = 0; j < hwl; ++j) tile[j] = halo_l[j];
    for (int j = 0; j < np; ++j) tile[hwl + j] = local[j];
    for (int j = 0; j < hwr; ++j) tile[hwl + np + j] = halo_r[j];
    double sd, sc;
    spline15_sums<1, false>(tile.data() + hwl, np, si, rem_d, rem_c, &sd, &sc);
    spline15_line<1, false, true>(tile.data() + hwl, np, si, alpha, sd, sc, out, 1);
    return 0;
}
}
