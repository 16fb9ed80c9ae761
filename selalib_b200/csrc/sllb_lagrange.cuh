// Lagrange weight polynomials and per-line stencil setup shared by the Lagrange kernels (K2).
#pragma once
#include <math.h>

namespace sllb {

// ------------------------------------------------------------------------------------------------
// Lagrange weights: closed-form polynomials of sll_m_lagrange_interpolation_1d_fast.F90
// (:59-67,110-121,170-182 even; :239-246,286-295,341-352,405-419,478-494 odd)
// ------------------------------------------------------------------------------------------------
// 1 / prod_{j != k} (k - j) of the S-point Lagrange basis on consecutive integers, at compile time
template <int S>
__host__ __device__ constexpr double lagr_inv_den(int k) {
    double den = 1.0;
    for (int j = 0; j < S; ++j)
        if (j != k) den *= (double)(k - j);
    return 1.0 / den;
}

template <int S>
__device__ __forceinline__ void lagr_coeff(double p, double *pp) {
    const double p2 = p * p;
    if constexpr (S == 3) {
        pp[0] = p * (p - 1.) * 0.5; pp[1] = 1. - p * p; pp[2] = p * (p + 1.) * 0.5;
    } else if constexpr (S == 5) {
        pp[0] = (p2 - 1.) * p * (p - 2.) * (1. / 24.);
        pp[1] = -(p - 1.) * p * (p2 - 4.) * (1. / 6.);
        pp[2] = (p2 - 1.) * (p2 - 4.) * 0.25;
        pp[3] = -(p + 1.) * p * (p2 - 4.) * (1. / 6.);
        pp[4] = (p2 - 1.) * p * (p + 2.) * (1. / 24.);
    } else if constexpr (S == 7) {
        pp[0] = p * (p - 3.) * (p2 - 4.) * (p2 - 1.) * (1. / 720.);
        pp[1] = -p * (p - 2.) * (p2 - 9.) * (p2 - 1.) * (1. / 120.);
        pp[2] = p * (p - 1.) * (p2 - 9.) * (p2 - 4.) * (1. / 48.);
        pp[3] = -(p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 36.);
        pp[4] = (p + 1.) * p * (p2 - 9.) * (p2 - 4.) * (1. / 48.);
        pp[5] = -(p + 2.) * p * (p2 - 9.) * (p2 - 1.) * (1. / 120.);
        pp[6] = (p + 3.) * p * (p2 - 4.) * (p2 - 1.) * (1. / 720.);
    } else if constexpr (S == 9) {
        pp[0] = p * (p - 4.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 40320.);
        pp[1] = -p * (p - 3.) * (p2 - 16.) * (p2 - 4.) * (p2 - 1.) * (1. / 5040.);
        pp[2] = p * (p - 2.) * (p2 - 16.) * (p2 - 9.) * (p2 - 1.) * (1. / 1440.);
        pp[3] = -p * (p - 1.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (1. / 720.);
        pp[4] = (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 576.);
        pp[5] = -(p + 1.) * p * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (1. / 720.);
        pp[6] = (p + 2.) * p * (p2 - 16.) * (p2 - 9.) * (p2 - 1.) * (1. / 1440.);
        pp[7] = -(p + 3.) * p * (p2 - 16.) * (p2 - 4.) * (p2 - 1.) * (1. / 5040.);
        pp[8] = (p + 4.) * p * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 40320.);
    } else if constexpr (S == 11) {
        pp[0] = p * (p - 5.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 3628800.);
        pp[1] = -p * (p - 4.) * (p2 - 25.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 362880.);
        pp[2] = p * (p - 3.) * (p2 - 25.) * (p2 - 16.) * (p2 - 4.) * (p2 - 1.) * (1. / 80640.);
        pp[3] = -p * (p - 2.) * (p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 1.) * (1. / 30240.);
        pp[4] = p * (p - 1.) * (p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (1. / 17280.);
        pp[5] = -(p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 14400.);
        pp[6] = (p + 1.) * p * (p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (1. / 17280.);
        pp[7] = -(p + 2.) * p * (p2 - 25.) * (p2 - 16.) * (p2 - 9.) * (p2 - 1.) * (1. / 30240.);
        pp[8] = (p + 3.) * p * (p2 - 25.) * (p2 - 16.) * (p2 - 4.) * (p2 - 1.) * (1. / 80640.);
        pp[9] = -(p + 4.) * p * (p2 - 25.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 362880.);
        pp[10] = (p + 5.) * p * (p2 - 16.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 3628800.);
    } else if constexpr (S == 4) {
        pp[0] = -p * (p - 1.) * (p - 2.) * (1. / 6.);
        pp[1] = (p2 - 1.) * (p - 2.) * 0.5;
        pp[2] = -p * (p + 1.) * (p - 2.) * 0.5;
        pp[3] = p * (p2 - 1.) * (1. / 6.);
    } else if constexpr (S == 6) {
        pp[0] = -p * (p2 - 1.) * (p - 2.) * (p - 3.) * (1. / 120.);
        pp[1] = p * (p - 1.) * (p2 - 4.) * (p - 3.) * (1. / 24.);
        pp[2] = -(p2 - 1.) * (p2 - 4.) * (p - 3.) * (1. / 12.);
        pp[3] = p * (p + 1.) * (p2 - 4.) * (p - 3.) * (1. / 12.);
        pp[4] = -p * (p2 - 1.) * (p + 2.) * (p - 3.) * (1. / 24.);
        pp[5] = p * (p2 - 1.) * (p2 - 4.) * (1. / 120.);
    } else if constexpr (S == 8) {
        pp[0] = -p * (p - 3.) * (p - 4.) * (p2 - 4.) * (p2 - 1.) * (1. / 5040.);
        pp[1] = p * (p - 2.) * (p - 4.) * (p2 - 9.) * (p2 - 1.) * (1. / 720.);
        pp[2] = -p * (p - 1.) * (p - 4.) * (p2 - 9.) * (p2 - 4.) * (1. / 240.);
        pp[3] = (p - 4.) * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 144.);
        pp[4] = -(p + 1.) * p * (p - 4.) * (p2 - 9.) * (p2 - 4.) * (1. / 144.);
        pp[5] = (p + 2.) * p * (p - 4.) * (p2 - 9.) * (p2 - 1.) * (1. / 240.);
        pp[6] = -(p + 3.) * p * (p - 4.) * (p2 - 4.) * (p2 - 1.) * (1. / 720.);
        pp[7] = p * (p2 - 9.) * (p2 - 4.) * (p2 - 1.) * (1. / 5040.);
    } else {
        // even stencils beyond the reference's closed forms (sll_p_lagrange of sll_s_periodic_interp takes any order,
        // sll_m_periodic_interp.F90:290-366): the same basis in product form, nodes -(S/2-1) .. S/2 around the foot cell
        static_assert((S & 1) == 0 && S >= 10 && S <= 18, "Lagrange stencil not implemented");
#pragma unroll
        for (int k = 0; k < S; ++k) {
            double num = 1.0;
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (j != k) num *= p - (double)(j - (S / 2 - 1));
            pp[k] = num * lagr_inv_den<S>(k);
        }
    }
}

// weights and first stencil offset (mod N) for a line.  Odd S: fixed stencil centred on the grid point,
// p = whole displacement (fast_disp_fixed_periodic, :609-654).  Even S: centred on the foot cell,
// pi = floor(p), weights at p - pi (fast_disp_centered_periodicl, :710-769).
template <int S>
__device__ __forceinline__ int lagr_setup(double disp, int N, double *pp) {
    int off;
    if constexpr ((S & 1) != 0) {
        lagr_coeff<S>(disp, pp);
        off = -(S - 1) / 2;
    } else {
        const double fl = floor(disp);
        lagr_coeff<S>(disp - fl, pp);
        off = -(S / 2 - 1) + (int)fl;
    }
    return ((off % N) + N) % N;
}

} // namespace sllb
