// sllb_poisson_direct.cu -- the 2D periodic Poisson solve of the time loops for SMALL grids (N1, N2 <= 256) in three
// kernels, without a library FFT.
//
// Reference: sll_t_poisson_2d_periodic (sll_m_poisson_2d_periodic.F90:250-383: r2c, spectral multipliers with
// kx(1,1) := 1 and the negative Nyquist wavenumber in x2, c2r) and sll_s_poisson_2d_periodic_par_solve
// (sll_m_poisson_2d_periodic_par.F90:214-338: Delta phi = rho, zero mean).
//
// Why: a 128 x 128 solve is 16 K points.  As cuFFT calls it is ~10 launches of a few microseconds each per field solve,
// twice per time step, and on 8 GPUs it does not shrink with the number of ranks (every rank solves the replicated
// problem): 0.15 ms of a 1.4 ms step.  The transform itself is tiny: written as dense DFT sums
//     A(k1, x2)  = sum_x1 rho(x1, x2) w1^(k1 x1)                      K_A  one block per x2 row      (real -> half complex)
//     rho^(k1,k2) = sum_x2 A(k1, x2) w2^(k2 x2);  multipliers;          K_B  one block per k1 column   (forward, multiply,
//     Z_q(k1, x2) = sum_k2 E^_q(k1, k2) w2^(-k2 x2)                          and inverse along x2 never leave the block)
//     E_q(x1, x2) = Re part of the half-complex inverse along k1        K_C  one block per x2 row
// it is 3 x 128^3 complex multiply-adds per field = a few microseconds spread over the SMs, in 3 launches, with the sum
// over the ranks' partial densities fused into K_A and the extraction of the local E tiles fused into K_C.
// The inverse along k1 uses only the real parts of the k1 = 0 and k1 = N1/2 columns, which is exactly what FFTW's c2r
// does with a spectrum that is not Hermitian there (the multipliers -i k/|k|^2 are not); the cuFFT path reaches the same
// values by symmetrising those columns (k_poisson2d).  Twiddles come from sincospi (correctly rounded arguments), the sums
// are plain fp64 FMA chains of <= 256 terms: agreement with the cuFFT path ~1e-15 relative to max|E|.
#include "sllb_internal.h"
#include "sllb_device.cuh"

namespace sllb {

struct Poisson2dDirect {
    int n1 = 0, n2 = 0;
    double L1 = 1, L2 = 1;
    double2 *tw1 = nullptr, *tw2 = nullptr;   // (cos, sin)(2 pi k / N)
    double2 *A = nullptr;                     // [x2][k1] half-complex rows, (N1/2+1) x N2
    double2 *Z = nullptr;                     // 3 x [x2][k1]: phi, E1, E2 after the x2 round trip
};

__global__ void k_twiddles(double2 *tw, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s, c;
    sincospi(2.0 * (double)k / (double)n, &s, &c);
    tw[k] = make_double2(c, s);
}

// K_A: block = row x2; thread k1 <= N1/2.  rho(x1, x2) = scale * sum over `nslots` arrays spaced by slot_stride.
__global__ void __launch_bounds__(256) k_pd_rows_fwd(const double *__restrict__ rho, const int nslots, const long long slot_stride,
                                                     const double scale, const int n1, const double2 *__restrict__ tw1,
                                                     double2 *__restrict__ A, double *__restrict__ rho_sum) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *row = reinterpret_cast<double *>(smem_raw);        // n1
    double2 *tw = reinterpret_cast<double2 *>(row + n1);        // n1
    const int x2 = blockIdx.x, h1 = n1 / 2 + 1;
    for (int x = threadIdx.x; x < n1; x += blockDim.x) {
        double v = rho[(size_t)x2 * n1 + x];
        for (int s = 1; s < nslots; ++s) v += rho[(size_t)s * slot_stride + (size_t)x2 * n1 + x];  // fixed order: identical on every rank
        v *= scale;
        row[x] = v;
        if (rho_sum) rho_sum[(size_t)x2 * n1 + x] = v;
        tw[x] = tw1[x];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < h1; k += blockDim.x) {
        double re = 0.0, im = 0.0;
        int idx = 0;
        for (int x = 0; x < n1; ++x) {
            const double2 w = tw[idx];
            re = fma(row[x], w.x, re);
            im = fma(-row[x], w.y, im);     // e^{-i theta}
            idx += k; if (idx >= n1) idx -= n1;
        }
        A[(size_t)x2 * h1 + k] = make_double2(re, im);
    }
}

// K_B: block = column k1; forward along x2, multipliers, inverse along k2 for up to three outputs.
// mode 0: sll_t_poisson_2d_periodic (phi = rho^/k2, E = -i k/k2 rho^, kx(1,1) := 1, negative Nyquist in x2)
// mode 1: sll_s_poisson_2d_periodic_par_solve (phi^ = -rho^/|k|^2, zero mean; E outputs unused)
__global__ void __launch_bounds__(256) k_pd_cols(const double2 *__restrict__ A, const int n1, const int n2, const double kx0,
                                                 const double ky0, const double2 *__restrict__ tw2, const int mode,
                                                 const int want_phi, const int want_e, double2 *__restrict__ Z) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *col = reinterpret_cast<double2 *>(smem_raw);   // n2: A(k1, :) then reused
    double2 *tw = col + n2;                                 // n2
    double2 *sp = tw + n2;                                  // 3 * n2 spectra (phi, e1, e2)
    const int k1 = blockIdx.x, h1 = n1 / 2 + 1;
    for (int x = threadIdx.x; x < n2; x += blockDim.x) { col[x] = A[(size_t)x * h1 + k1]; tw[x] = tw2[x]; }
    __syncthreads();
    for (int k2 = threadIdx.x; k2 < n2; k2 += blockDim.x) {
        double re = 0.0, im = 0.0;
        int idx = 0;
        for (int x = 0; x < n2; ++x) {
            const double2 w = tw[idx], a = col[x];
            // a * (w.x - i w.y)
            re = fma(a.x, w.x, fma(a.y, w.y, re));
            im = fma(a.y, w.x, fma(-a.x, w.y, im));
            idx += k2; if (idx >= n2) idx -= n2;
        }
        double2 ph = make_double2(0, 0), e1 = ph, e2 = ph;
        if (mode == 0) {
            double kx = (double)k1 * kx0;
            const double ky = (double)((k2 < n2 / 2) ? k2 : k2 - n2) * ky0;
            if (k1 == 0 && k2 == 0) kx = 1.0;
            const double k2n = kx * kx + ky * ky;
            const double kxs = kx / k2n, kys = ky / k2n;
            ph = make_double2(re / k2n, im / k2n);
            e1 = make_double2(kxs * im, -kxs * re);
            e2 = make_double2(kys * im, -kys * re);
        } else if (!(k1 == 0 && k2 == 0)) {
            // wavenumbers as (k / L)^2 4 pi^2 with both indices folded to -n/2 .. n/2-1 (:300-306)
            const double kx = (double)k1 * kx0;   // k1 <= n1/2: the fold only changes the sign at the Nyquist index
            const double ky = (double)((k2 < n2 / 2) ? k2 : k2 - n2) * ky0;
            const double d = -(kx * kx + ky * ky);
            ph = make_double2(re / d, im / d);
        }
        sp[k2] = ph; sp[n2 + k2] = e1; sp[2 * n2 + k2] = e2;
    }
    __syncthreads();
    const int q0 = want_phi ? 0 : 1, q1 = want_e ? 3 : 1;
    for (int x = threadIdx.x; x < n2; x += blockDim.x) {
        for (int q = q0; q < q1; ++q) {
            const double2 *s = sp + (size_t)q * n2;
            double re = 0.0, im = 0.0;
            int idx = 0;
            for (int k2 = 0; k2 < n2; ++k2) {
                const double2 w = tw[idx], a = s[k2];
                // a * (w.x + i w.y)
                re = fma(a.x, w.x, fma(-a.y, w.y, re));
                im = fma(a.y, w.x, fma(a.x, w.y, im));
                idx += x; if (idx >= n2) idx -= n2;
            }
            Z[((size_t)q * n2 + x) * h1 + k1] = make_double2(re, im);
        }
    }
}

// K_C: block = row x2; thread x1.  out(x1) = [Z(0).re + (-1)^x1 Z(N1/2).re + 2 sum_{0<k<N1/2} Re(Z(k) w1^(-k x1))] / (N1 N2).
// Optionally also writes the (x1, x2) sub-box [lo0, lo0+t0) x [lo1, lo1+t1) of E1, E2 into dense tile arrays.
struct PdTile { double *e1, *e2; int lo0, t0, lo1, t1; double *nrj_rows; };
__global__ void __launch_bounds__(256) k_pd_rows_inv(const double2 *__restrict__ Z, const int n1, const int n2,
                                                     const double2 *__restrict__ tw1, double *__restrict__ phi,
                                                     double *__restrict__ e1, double *__restrict__ e2, const PdTile tile) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int h1 = n1 / 2 + 1;
    double2 *z = reinterpret_cast<double2 *>(smem_raw);     // 3 * h1
    double2 *tw = z + 3 * h1;                               // n1
    const int x2 = blockIdx.x;
    double *outs[3] = {phi, e1, e2};
    for (int k = threadIdx.x; k < h1; k += blockDim.x)
        for (int q = 0; q < 3; ++q)
            if (outs[q]) z[q * h1 + k] = Z[((size_t)q * n2 + x2) * h1 + k];
    for (int x = threadIdx.x; x < n1; x += blockDim.x) tw[x] = tw1[x];
    __syncthreads();
    const double nrm = 1.0 / ((double)n1 * (double)n2);
    double esq = 0.0;   // this thread's part of sum_x1 w (E1^2 + E2^2) on the row
    for (int x1 = threadIdx.x; x1 < n1; x1 += blockDim.x) {
        for (int q = 0; q < 3; ++q) {
            if (!outs[q]) continue;
            const double2 *zq = z + q * h1;
            double acc = 0.0;
            int idx = x1; if (idx >= n1) idx -= n1;
            for (int k = 1; k < n1 / 2; ++k) {
                const double2 w = tw[idx], a = zq[k];
                acc = fma(a.x, w.x, fma(-a.y, w.y, acc));   // Re(a * e^{+i theta})
                idx += x1; if (idx >= n1) idx -= n1;
            }
            const double v = (zq[0].x + ((x1 & 1) ? -zq[n1 / 2].x : zq[n1 / 2].x) + 2.0 * acc) * nrm;
            outs[q][(size_t)x2 * n1 + x1] = v;
            if (q > 0) esq = fma((x1 == 0 ? 2.0 : 1.0) * (x2 == 0 ? 2.0 : 1.0) * v, v, esq);
            if (q > 0 && tile.e1) {
                const int a0 = x1 - tile.lo0, a1 = x2 - tile.lo1;
                if (a0 >= 0 && a0 < tile.t0 && a1 >= 0 && a1 < tile.t1) (q == 1 ? tile.e1 : tile.e2)[(size_t)a1 * tile.t0 + a0] = v;
            }
        }
    }
    if (tile.nrj_rows) {   // fixed-shape block sum: warp shuffles, then the warp results in order
        __shared__ double wsum[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) esq += __shfl_xor_sync(0xffffffffu, esq, o);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = esq;
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = wsum[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) a += wsum[w];
            tile.nrj_rows[x2] = a;
        }
    }
}

int poisson2d_direct_supported(int n1, int n2) {
    return n1 >= 4 && n2 >= 4 && n1 <= 256 && n2 <= 256 && n1 % 2 == 0;
}
int poisson2d_direct_create(int n1, int n2, double L1, double L2, Poisson2dDirect **out) {
    if (!poisson2d_direct_supported(n1, n2)) return fail(SLLB_ERR_UNSUPPORTED, "poisson2d_direct: 4 <= N <= 256, N1 even");
    Poisson2dDirect *P = new Poisson2dDirect();
    P->n1 = n1; P->n2 = n2; P->L1 = L1; P->L2 = L2;
    const size_t h1 = (size_t)n1 / 2 + 1;
    cudaError_t e = cudaMalloc(&P->tw1, sizeof(double2) * n1);
    if (e == cudaSuccess) e = cudaMalloc(&P->tw2, sizeof(double2) * n2);
    if (e == cudaSuccess) e = cudaMalloc(&P->A, sizeof(double2) * h1 * n2);
    if (e == cudaSuccess) e = cudaMalloc(&P->Z, sizeof(double2) * 3 * h1 * n2);
    if (e == cudaSuccess) {
        k_twiddles<<<(n1 + 127) / 128, 128, 0, g_stream>>>(P->tw1, n1);
        k_twiddles<<<(n2 + 127) / 128, 128, 0, g_stream>>>(P->tw2, n2);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { poisson2d_direct_destroy(P); return check_cuda(e, "poisson2d_direct_create"); }
    *out = P;
    return SLLB_OK;
}
void poisson2d_direct_destroy(Poisson2dDirect *P) {
    if (!P) return;
    cudaFree(P->tw1); cudaFree(P->tw2); cudaFree(P->A); cudaFree(P->Z);
    delete P;
}
cudaError_t poisson2d_direct_solve(Poisson2dDirect *P, const double *rho, int nslots, long long slot_stride, double scale,
                                   double *rho_sum, int mode, double *phi, double *e1, double *e2, double *nrj_rows,
                                   double *tile_e1, double *tile_e2, const int tile_box[4], cudaStream_t st) {
    const int n1 = P->n1, n2 = P->n2, h1 = n1 / 2 + 1;
    const double tp = 2.0 * 3.14159265358979323846;
    const int ta = h1 < 256 ? ((h1 + 31) / 32) * 32 : 256;
    k_pd_rows_fwd<<<n2, ta, (size_t)n1 * 8 + (size_t)n1 * 16, st>>>(rho, nslots, slot_stride, scale, n1, P->tw1, P->A, rho_sum);
    count_launch();
    const int tb = n2 < 256 ? ((n2 + 31) / 32) * 32 : 256;
    k_pd_cols<<<h1, tb, (size_t)n2 * 16 * 5, st>>>(P->A, n1, n2, tp / P->L1, tp / P->L2, P->tw2, mode, phi != nullptr,
                                                      (e1 != nullptr || e2 != nullptr), P->Z);
    count_launch();
    PdTile tile;
    tile.e1 = tile_e1; tile.e2 = tile_e2;
    tile.lo0 = tile_box ? tile_box[0] : 0; tile.t0 = tile_box ? tile_box[1] : 0;
    tile.lo1 = tile_box ? tile_box[2] : 0; tile.t1 = tile_box ? tile_box[3] : 0;
    tile.nrj_rows = (e1 && e2) ? nrj_rows : nullptr;
    const int tc = n1 < 256 ? ((n1 + 31) / 32) * 32 : 256;
    k_pd_rows_inv<<<n2, tc, (size_t)h1 * 16 * 3 + (size_t)n1 * 16, st>>>(P->Z, n1, n2, P->tw1, phi, e1, e2, tile);
    count_launch();
    return cudaGetLastError();
}

} // namespace sllb
