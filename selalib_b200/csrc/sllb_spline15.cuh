// Per-line arithmetic of K9 (local cubic spline with halo cells).  Kept in a header of its own, free of CUDA-only
// constructs apart from the macros below, so that tests/host/spline15_host.cpp can compile the very same
// functions with g++ and check them against the reference arithmetic on a machine without a GPU (test infrastructure: the
// product only ever runs them on the device).
#pragma once
#include <math.h>
#ifdef SLLB_HOST_EMULATION
#define SLLB_DEV static inline
#define SLLB_CONST static
#define SLLB_ST(p, v) (*(p) = (v))
#define SLLB_LDG(p) (*(p))
#else
#define SLLB_DEV __device__ __forceinline__
#define SLLB_CONST __constant__
#define SLLB_ST(p, v) __stcs((p), (v))
#define SLLB_LDG(p) __ldg(p)
#endif

namespace sllb {

#define SLLB_HALO_TERMS 15
SLLB_CONST double c_hpw[SLLB_HALO_TERMS + 1]; // (-b/a)^i, i = 0..15

struct Weights4 { double w0, w1, w2, w3; };
SLLB_DEV Weights4 spline_weights(const double dx) {
    const double r2 = 1.60769515458673623883; // 6 (2 - sqrt 3) = 1/a^2
    const double cdx = 1.0 - dx, s6 = r2 * (1.0 / 6.0);
    Weights4 w;
    w.w0 = cdx * cdx * cdx * s6;
    w.w1 = (1.0 + 3.0 * cdx + 3.0 * cdx * cdx - 3.0 * cdx * cdx * cdx) * s6;
    w.w2 = (1.0 + 3.0 * dx + 3.0 * dx * dx - 3.0 * dx * dx * dx) * s6;
    w.w3 = dx * dx * dx * s6;
    return w;
}

SLLB_DEV int wrap_idx(int k, const int n) {
    k %= n;
    return k < 0 ? k + n : k;
}

// Complete boundary sums of one line.  x0 points at local cell 0 of the line (element k at x0[k*PITCH]); without WRAP
// the rows k = -hwl..-1 and np..np+hwr-1 hold halo cells but only LOCAL cells (0 <= k < np) are summed here, the rest
// arrives in rem_d / rem_c.
template <int PITCH, bool WRAP>
SLLB_DEV void spline15_sums(const double *x0, const int np, const int si, const double rem_d,
                                              const double rem_c, double *sum_d, double *sum_c) {
    double sd = WRAP ? 0.0 : rem_d, sc = WRAP ? 0.0 : rem_c;
    if (WRAP) {
        // two / four independent partial sums: the 16- and 31-term chains are latency, not throughput
        double sd1 = 0.0, sc1 = 0.0, sc2 = 0.0, sc3 = 0.0;
        int idx = wrap_idx(si - 1, np);
#pragma unroll
        for (int i = 0; i <= SLLB_HALO_TERMS; ++i) {
            if (i & 1) sd1 = fma(c_hpw[i], x0[idx * PITCH], sd1);
            else sd = fma(c_hpw[i], x0[idx * PITCH], sd);
            idx = (idx == 0) ? np - 1 : idx - 1;
        }
        sd += sd1;
        idx = wrap_idx(si + 1 - SLLB_HALO_TERMS, np); // k = si+np+1+m, m = -15
#pragma unroll
        for (int m = -SLLB_HALO_TERMS; m <= SLLB_HALO_TERMS; ++m) {
            const double t = c_hpw[m < 0 ? -m : m], v = x0[idx * PITCH];
            switch ((m + SLLB_HALO_TERMS) & 3) {
            case 0: sc = fma(t, v, sc); break;
            case 1: sc1 = fma(t, v, sc1); break;
            case 2: sc2 = fma(t, v, sc2); break;
            default: sc3 = fma(t, v, sc3); break;
            }
            idx = (idx == np - 1) ? 0 : idx + 1;
        }
        sc = (sc + sc1) + (sc2 + sc3);
    } else {
#pragma unroll
        for (int i = 0; i <= SLLB_HALO_TERMS; ++i) {
            const int k = si - 1 - i;
            if (k >= 0) sd = fma(c_hpw[i], x0[k * PITCH], sd);
        }
#pragma unroll
        for (int m = -SLLB_HALO_TERMS; m <= SLLB_HALO_TERMS; ++m) {
            const int k = si + np + 1 + m;
            if (k < np) sc = fma(c_hpw[m < 0 ? -m : m], x0[k * PITCH], sc);
        }
    }
    *sum_d = sd; *sum_c = sc;
}

// Recurrences + evaluation of one line.  TO_GLOBAL: out(local j) -> gout[j*gstride] (streaming stores).
// Otherwise out(local j) is parked in the slot of e(j+1), i.e. x0[wrap(si + j)*PITCH] (WRAP only).
template <int PITCH, bool WRAP, bool TO_GLOBAL>
SLLB_DEV void spline15_line(double *x0, const int np, const int si, const double alpha,
                                              const double sum_d, const double sum_c, double *gout,
                                              const long long gstride) {
    const double q = 0.26794919243112270647;  // 2 - sqrt(3) = b/a
    const double kc = 1.07735026918962576451; // a^2 sqrt(3) = (3 + 2 sqrt 3)/6
    const Weights4 w = spline_weights(alpha);
    const int i0 = WRAP ? wrap_idx(si, np) : si; // slot of W(1)
    const double wlast = WRAP ? x0[i0 * PITCH] : x0[(si + np) * PITCH]; // W(np+1)
    // forward: e(0) = sum_d, e(j) = W(j) - q e(j-1), kept in the slot of W(j)
    double e = sum_d;
    int idx = i0;
#pragma unroll 4
    for (int j = 1; j <= np; ++j) {
        e = fma(-q, e, x0[idx * PITCH]);
        x0[idx * PITCH] = e;
        if (WRAP) idx = (idx == np - 1) ? 0 : idx + 1;
        else ++idx;
    }
    const double e_np1 = fma(-q, e, wlast);
    // backward: G(np+2) = a^2 c_np2, G(j) = e(j) - q G(j+1)
    double a3 = kc * sum_c;
    double a2 = fma(-q, a3, e_np1);
    if (WRAP) idx = (idx == 0) ? np - 1 : idx - 1;
    else --idx; // slot of e(np)
    double a1 = fma(-q, a2, x0[idx * PITCH]);
    double *p = TO_GLOBAL ? gout + (long long)(np - 1) * gstride : nullptr;
#pragma unroll 4
    for (int j = np - 1; j >= 0; --j) {
        const int slot_next = idx; // slot of e(j+1): free from now on
        double ej;
        if (j > 0) {
            if (WRAP) idx = (idx == 0) ? np - 1 : idx - 1;
            else --idx;
            ej = x0[idx * PITCH];
        } else ej = sum_d;
        const double a0 = fma(-q, a1, ej);
        const double val = fma(w.w3, a3, fma(w.w2, a2, fma(w.w1, a1, w.w0 * a0))); // cell j+1 = local index j
        if (TO_GLOBAL) { SLLB_ST(p, val); p -= gstride; }
        else x0[slot_next * PITCH] = val;
        a3 = a2; a2 = a1; a1 = a0;
    }
}

// sll_s_cubic_spline_halo_1d_prepare_exchange for one line read in place (element k at base[k*stride]): the parts of
// the neighbours' boundary sums made of MY cells.
//   sd = sum_{i=0..15, si-1-i < 0}   (-q)^i f(np + si-1-i)      -> the right neighbour's d_0   (:78-91)
//   sc = sum_{m=-15..15, si+1+m >= 0} (-q)^|m| f(si+1+m)        -> the left neighbour's c_np2  (:92-103)
SLLB_DEV void spline15_prepare(const double *base, const long long stride, const int np, const int si, double *sd_out,
                               double *sc_out) {
    double sd = 0.0, sc = 0.0;
#pragma unroll
    for (int i = 0; i <= SLLB_HALO_TERMS; ++i) {
        const int k = si - 1 - i;
        if (k < 0) sd = fma(c_hpw[i], SLLB_LDG(base + (long long)(np + k) * stride), sd);
    }
#pragma unroll
    for (int m = -SLLB_HALO_TERMS; m <= SLLB_HALO_TERMS; ++m) {
        const int k = si + 1 + m;
        if (k >= 0) sc = fma(c_hpw[m < 0 ? -m : m], SLLB_LDG(base + (long long)k * stride), sc);
    }
    *sd_out = sd; *sc_out = sc;
}

} // namespace sllb
