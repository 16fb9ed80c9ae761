// sllb_diag.cu -- the per-step diagnostics of the 2D2V time loop (sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:1112,
// 1183-1260: field energy, mass, L1/L2 norms, kinetic energy) computed on the device, so that a run with diagnostics
// every step (what the reference does) never waits for the host: every kernel below is a deterministic fixed-shape
// reduction launched on the time loop's stream, and the rows stay in HBM until the run is over.
#include "sllb_internal.h"
#include "sllb_device.cuh"

namespace sllb {

namespace {
constexpr int RT = 1024;
// fixed-order block sum of NV values per thread (warp shuffles, then the 32 warp results by thread 0)
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *out, bool do_max = false) {
    __shared__ double sh[NV][32];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, v[k], o);
            v[k] = do_max ? fmax(v[k], t) : v[k] + t;
        }
    }
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (ln == 0)
        for (int k = 0; k < NV; ++k) sh[k][w] = v[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < NV; ++k) {
            double t = sh[k][0];
            for (int i = 1; i < nw; ++i) t = do_max ? fmax(t, sh[k][i]) : t + sh[k][i];
            out[k] = t;
        }
}
} // namespace

__global__ void __launch_bounds__(RT) k_moments_from_rows(const double *__restrict__ rows3, const int n3, const int n4,
                                                          const double *__restrict__ w3, const double *__restrict__ w4,
                                                          double *__restrict__ out4) {
    double v[4] = {0, 0, 0, 0};
    const long long nv = (long long)n3 * n4;
    for (long long r = threadIdx.x; r < nv; r += RT) {
        const double s0 = rows3[3 * r];
        v[0] += s0; v[1] += rows3[3 * r + 1]; v[2] += rows3[3 * r + 2];
        v[3] = fma(w3[r % n3] + w4[r / n3], s0, v[3]);
    }
    block_sum<4>(v, out4);
}
cudaError_t launch_moments_from_rows(const double *rows3, int n3, int n4, const double *w3, const double *w4, double *out4,
                                     cudaStream_t st) {
    k_moments_from_rows<<<1, RT, 0, st>>>(rows3, n3, n4, w3, w4, out4);
    count_launch();
    return cudaGetLastError();
}

static const int ML_BLOCKS = 148;
size_t moments_from_lines_scratch() { return (size_t)ML_BLOCKS * 4; }
__global__ void __launch_bounds__(RT) k_moments_from_lines1(const double *__restrict__ sum_, const double *__restrict__ l1,
                                                            const double *__restrict__ l2, const double *__restrict__ kin,
                                                            const long long nx, const int n3, const double *__restrict__ w3,
                                                            double *__restrict__ partial) {
    double v[4] = {0, 0, 0, 0};
    const long long n = nx * n3;
    for (long long l = (long long)blockIdx.x * RT + threadIdx.x; l < n; l += (long long)gridDim.x * RT) {
        const double s0 = sum_[l];
        v[0] += s0; v[1] += l1[l]; v[2] += l2[l];
        v[3] += fma(w3[l / nx], s0, kin[l]);
    }
    block_sum<4>(v, partial + 4 * blockIdx.x);
}
__global__ void __launch_bounds__(256) k_moments_from_lines2(const double *__restrict__ partial, const int nb,
                                                             double *__restrict__ out4) {
    double v[4] = {0, 0, 0, 0};
    for (int b = threadIdx.x; b < nb; b += 256)
        for (int k = 0; k < 4; ++k) v[k] += partial[4 * b + k];
    block_sum<4>(v, out4);
}
cudaError_t launch_moments_from_lines_partials(const double *sum_, const double *l1, const double *l2, const double *kin, long long nx,
                                               int n3, const double *w3, double *partial, int *nb_out, cudaStream_t st) {
    k_moments_from_lines1<<<ML_BLOCKS, RT, 0, st>>>(sum_, l1, l2, kin, nx, n3, w3, partial);
    count_launch();
    *nb_out = ML_BLOCKS;
    return cudaGetLastError();
}
cudaError_t launch_moments_from_lines(const double *sum_, const double *l1, const double *l2, const double *kin, long long nx,
                                      int n3, const double *w3, double *scratch, double *out4, cudaStream_t st) {
    k_moments_from_lines1<<<ML_BLOCKS, RT, 0, st>>>(sum_, l1, l2, kin, nx, n3, w3, scratch);
    count_launch();
    k_moments_from_lines2<<<1, 256, 0, st>>>(scratch, ML_BLOCKS, out4);
    count_launch();
    return cudaGetLastError();
}

__global__ void __launch_bounds__(RT) k_dup_energy2d(const double *__restrict__ a, const double *__restrict__ b, const int n1,
                                                     const int n2, const double scale, const int squared,
                                                     double *__restrict__ out1) {
    double v[1] = {0};
    const long long n = (long long)n1 * n2;
    for (long long k = threadIdx.x; k < n; k += RT) {
        const int i = (int)(k % n1), j = (int)(k / n1);
        // node (0, .) is also node (n1, .), node (., 0) also (., n2)
        const double w = (i == 0 ? 2.0 : 1.0) * (j == 0 ? 2.0 : 1.0);
        const double bb = squared ? b[k] * b[k] : b[k] + b[k];
        v[0] = fma(w, fma(a[k], a[k], bb), v[0]);
    }
    block_sum<1>(v, out1);
    if (threadIdx.x == 0) out1[0] *= scale;
}
cudaError_t launch_dup_energy2d(const double *a, const double *b, int n1, int n2, double scale, int squared, double *out1,
                                cudaStream_t st) {
    k_dup_energy2d<<<1, RT, 0, st>>>(a, b, n1, n2, scale, squared, out1);
    count_launch();
    return cudaGetLastError();
}

__global__ void __launch_bounds__(RT) k_absmax(const double *__restrict__ a, const long long n, double *__restrict__ out1) {
    double v[1] = {0};
    for (long long k = threadIdx.x; k < n; k += RT) v[0] = fmax(v[0], fabs(a[k]));
    block_sum<1>(v, out1, true);
}
cudaError_t launch_absmax(const double *a, long long n, double *out1, cudaStream_t st) {
    k_absmax<<<1, RT, 0, st>>>(a, n, out1);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// 3D3V: the charge density and the nine velocity moments of the diagnostics row in ONE sweep over f.
// f is [nv][nx] (x fastest).  Block = 128 consecutive x, blockIdx.y = chunk of v: every thread adds f over the chunk's v
// (first stage of the density reduction, K3) and, on the side, nine accumulators sum_v w_k(v) f(x,v) with
// w = (1, |.|, (.)^2, v4, v5, v6, v4^2, v5^2, v6^2) (sll_s_time_history_diagnostics, sll_m_sim_6d_utilities.F90:249-644,
// asks for exactly these integrals); the block folds them and leaves 9 numbers per block.  The reference sweeps f once
// for rho and once more for the diagnostics; so did round 1 (k_reduce_stage1 + k_row_sums).
// ------------------------------------------------------------------------------------------------
#define SLLB_RM_XPT 4      // x values per thread: the six weights of a velocity index are read once per four elements
#define SLLB_RM_VMAX 64    // most velocity indices a chunk may hold (weights staged in shared memory)
__global__ void __launch_bounds__(128) k_reduce_moments6d(const double *__restrict__ f, const long long nx, const long long nv,
                                                          const int nchunks, const double *__restrict__ wt /* [nv][6] */,
                                                          double *__restrict__ partial, double *__restrict__ mom_part) {
    __shared__ double wsh[SLLB_RM_VMAX * 6];
    __shared__ double sh[9][4];
    const long long x0 = (long long)blockIdx.x * (128 * SLLB_RM_XPT) + threadIdx.x;
    const int c = blockIdx.y;
    const long long v0 = nv * c / nchunks, v1 = nv * (c + 1) / nchunks;
    for (int i = threadIdx.x; i < (int)(v1 - v0) * 6; i += 128) wsh[i] = wt[v0 * 6 + i];
    __syncthreads();
    double a[SLLB_RM_XPT], m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < SLLB_RM_XPT; ++j) a[j] = 0.0;
    for (long long v = v0; v < v1; ++v) {
        double t[SLLB_RM_XPT];
#pragma unroll
        for (int j = 0; j < SLLB_RM_XPT; ++j) {
            const long long x = x0 + 128 * j;
            t[j] = x < nx ? __ldcs(f + x + nx * v) : 0.0;
        }
        const double *w = wsh + 6 * (v - v0);
        double s1 = 0.0, sa = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < SLLB_RM_XPT; ++j) { a[j] += t[j]; s1 += t[j]; sa += fabs(t[j]); s2 = fma(t[j], t[j], s2); }
        m[1] += sa; m[2] += s2;
#pragma unroll
        for (int k = 0; k < 6; ++k) m[3 + k] = fma(w[k], s1, m[3 + k]);   // the weights multiply the sum over this thread's x
    }
#pragma unroll
    for (int j = 0; j < SLLB_RM_XPT; ++j) {
        const long long x = x0 + 128 * j;
        if (x < nx) partial[(long long)c * nx + x] = a[j];
        m[0] += a[j];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m[k] += __shfl_xor_sync(0xffffffffu, m[k], o);
    }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 9; ++k) sh[k][threadIdx.x >> 5] = m[k];
    __syncthreads();
    if (threadIdx.x < 9) {
        const int k = threadIdx.x;
        mom_part[((long long)c * gridDim.x + blockIdx.x) * 9 + k] = (sh[k][0] + sh[k][1]) + (sh[k][2] + sh[k][3]);
    }
}
__global__ void __launch_bounds__(RT) k_moments6d_finish(const double *__restrict__ mom_part, const long long nblocks,
                                                         double *__restrict__ out9) {
    __shared__ double sh[9][32];
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (long long b = threadIdx.x; b < nblocks; b += RT)
        for (int k = 0; k < 9; ++k) v[k] += mom_part[b * 9 + k];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 9; ++k) sh[k][threadIdx.x >> 5] = v[k];
    __syncthreads();
    if (threadIdx.x < 9) {
        double t = 0.0;
        for (int i = 0; i < 32; ++i) t += sh[threadIdx.x][i];
        out9[threadIdx.x] = t;
    }
}
static int rm_chunks(long long nv) {
    long long ch = nv < 128 ? nv : 128;
    while ((nv + ch - 1) / ch > SLLB_RM_VMAX) ch *= 2;   // at most SLLB_RM_VMAX velocity indices per chunk
    return (int)ch;
}
int reduce_moments6d_chunks(long long nv) { return rm_chunks(nv); }
size_t reduce_moments6d_scratch(long long nx, long long nv) {
    return (size_t)(((nx + 128 * SLLB_RM_XPT - 1) / (128 * SLLB_RM_XPT)) * rm_chunks(nv) * 9);
}
// partial: reduce_scratch_doubles(nx, nv) doubles (the density partials, to be folded by launch_sum_partials with *nchunks
// parts); mom_part: reduce_moments6d_scratch doubles; out9: the nine moments
cudaError_t launch_reduce_moments6d(const double *f, long long nx, long long nv, const double *wt, double *partial, int *nchunks_out,
                                    double *mom_part, double *out9, cudaStream_t st) {
    const int nchunks = rm_chunks(nv);
    dim3 grid((unsigned)((nx + 128 * SLLB_RM_XPT - 1) / (128 * SLLB_RM_XPT)), nchunks);
    k_reduce_moments6d<<<grid, 128, 0, st>>>(f, nx, nv, nchunks, wt, partial, mom_part);
    count_launch();
    k_moments6d_finish<<<1, RT, 0, st>>>(mom_part, (long long)grid.x * grid.y, out9);
    count_launch();
    *nchunks_out = nchunks;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// dup_velocity_planes mode of the 2D2V loop: the N4 + N3 + 1 side planes (x3 = N3 with x4 < N4, x4 = N4 with x3 < N3, the
// corner) that the reference's (N+1)-point arrays carry next to the periodic cells.
// ------------------------------------------------------------------------------------------------
// side[m] <- the cell plane it duplicates: (0, m) for m < n4, (m - n4, 0) for n4 <= m < n4 + n3, (0, 0) for the corner
__global__ void __launch_bounds__(256) k_dup_fill(const double *__restrict__ f, const long long n12, const int n3, const int n4,
                                                  double *__restrict__ side) {
    const long long ntot = n12 * (n3 + n4 + 1);
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < ntot; t += (long long)gridDim.x * 256) {
        const long long m = t / n12, x = t - m * n12;
        long long plane = 0;
        if (m < n4) plane = (long long)n3 * m;
        else if (m < n4 + n3) plane = m - n4;
        side[t] = f[plane * n12 + x];
    }
}
cudaError_t launch_dup_fill(const double *f, long long n12, int n3, int n4, double *side, cudaStream_t st) {
    k_dup_fill<<<148 * 4, 256, 0, st>>>(f, n12, n3, n4, side);
    count_launch();
    return cudaGetLastError();
}
// rho(x) += scale * [ sum_m cd(m) side[m](x) - 1/2 sum_{i4} c(i4) f(x,0,i4) ... ], i.e. the trapezoid rule over the
// (N3+1)(N4+1) nodes written as the plain sum over the cells plus a correction: the cells of the x3 = 0 and x4 = 0 planes
// lose half their weight (the corner cell three quarters), the side planes come in with weight 1/2 (the corner 1/4).
__global__ void __launch_bounds__(128) k_dup_rho_corr(const double *__restrict__ f, const double *__restrict__ side, const long long n12,
                                                      const int n3, const int n4, const double scale, double *__restrict__ rho) {
    const long long x = (long long)blockIdx.x * 128 + threadIdx.x;
    if (x >= n12) return;
    double a = 0.0;
    for (int m = 0; m < n4; ++m)       // x3 = N3 side (weight 1/2, 1/4 at x4 = 0) against the x3 = 0 cells they duplicate
        a += (m == 0 ? 0.25 : 0.5) * (side[(long long)m * n12 + x] - f[(long long)n3 * m * n12 + x]);
    for (int i3 = 0; i3 < n3; ++i3)    // x4 = N4 side against the x4 = 0 cells
        a += (i3 == 0 ? 0.25 : 0.5) * (side[(long long)(n4 + i3) * n12 + x] - f[(long long)i3 * n12 + x]);
    // corner node (N3, N4): weight 1/4; the cell (0,0) keeps 1/4 of its weight: 1 - 1/4 (above, twice) - 1/4 here
    a += 0.25 * (side[(long long)(n3 + n4) * n12 + x] - f[x]);
    rho[x] += scale * a;
}
cudaError_t launch_dup_rho_corr(const double *f, const double *side, long long n12, int n3, int n4, double scale, double *rho,
                                cudaStream_t st) {
    k_dup_rho_corr<<<(unsigned)((n12 + 127) / 128), 128, 0, st>>>(f, side, n12, n3, n4, scale, rho);
    count_launch();
    return cudaGetLastError();
}

// position-weighted checksums of a local box of the 4D field: out2 = (sum w f, sum w f^2) with a weight that depends on the
// GLOBAL index of the point, w = 1 + ((3 g0 + 5 g1 + 7 g2 + 11 g3) mod 64) / 64, so that any misplaced element shows
__global__ void __launch_bounds__(RT) k_checksum4d1(const double *__restrict__ f, const int n0, const int n1, const int n2,
                                                    const int n3, const int lo0, const int lo1, const int lo2, const int lo3,
                                                    double *__restrict__ partial) {
    double v[2] = {0, 0};
    const long long n = (long long)n0 * n1 * n2 * n3;
    for (long long t = (long long)blockIdx.x * RT + threadIdx.x; t < n; t += (long long)gridDim.x * RT) {
        long long r = t;
        const int i0 = (int)(r % n0); r /= n0;
        const int i1 = (int)(r % n1); r /= n1;
        const int i2 = (int)(r % n2); r /= n2;
        const int i3 = (int)r;
        const int h = (3 * (i0 + lo0) + 5 * (i1 + lo1) + 7 * (i2 + lo2) + 11 * (i3 + lo3)) & 63;
        const double w = 1.0 + (double)h * (1.0 / 64.0), x = f[t];
        v[0] = fma(w, x, v[0]);
        v[1] = fma(w * x, x, v[1]);
    }
    block_sum<2>(v, partial + 2 * blockIdx.x);
}
__global__ void __launch_bounds__(256) k_checksum4d2(const double *__restrict__ partial, const int nb, double *__restrict__ out2) {
    double v[2] = {0, 0};
    for (int b = threadIdx.x; b < nb; b += 256) { v[0] += partial[2 * b]; v[1] += partial[2 * b + 1]; }
    block_sum<2>(v, out2);
}
cudaError_t launch_checksum4d(const double *f, const int ext[4], const int lo[4], double *scratch, double *out2, cudaStream_t st) {
    k_checksum4d1<<<ML_BLOCKS, RT, 0, st>>>(f, ext[0], ext[1], ext[2], ext[3], lo[0], lo[1], lo[2], lo[3], scratch);
    count_launch();
    k_checksum4d2<<<1, 256, 0, st>>>(scratch, ML_BLOCKS, out2);
    count_launch();
    return cudaGetLastError();
}

// |f^_k(v)|^2 for the first modes k < nmodes of every row f(:, v), f^_k = (1/n1) sum_x f(x, v) e^{-2 pi i k x / n1}
// (the normalised r2r transform + sll_f_fft_get_mode_r2c_1d of sll_m_sim_bsl_vp_1d1v_cart.F90:1745-1754)
__global__ void __launch_bounds__(128) k_row_modes(const double *__restrict__ f, const int n1, const int nmodes,
                                                   double *__restrict__ part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *row = reinterpret_cast<double *>(smem_raw);
    double2 *tw = reinterpret_cast<double2 *>(row + n1);
    const long long v = blockIdx.x;
    for (int x = threadIdx.x; x < n1; x += blockDim.x) {
        row[x] = f[v * n1 + x];
        double sn, cs;
        sincospi(2.0 * (double)x / (double)n1, &sn, &cs);
        tw[x] = make_double2(cs, sn);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nmodes; k += blockDim.x) {
        double re = 0.0, im = 0.0;
        int idx = 0;
        const int kk = k % n1;
        for (int x = 0; x < n1; ++x) {
            re = fma(row[x], tw[idx].x, re);
            im = fma(-row[x], tw[idx].y, im);
            idx += kk; if (idx >= n1) idx -= n1;
        }
        re /= (double)n1; im /= (double)n1;
        part[v * nmodes + k] = re * re + im * im;
    }
}
// out[k] = sum_v w * part[v][k], fixed order
__global__ void k_modes_reduce(const double *__restrict__ part, const long long nv, const int nmodes, const double w,
                               double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nmodes) return;
    double a = 0.0;
    for (long long v = 0; v < nv; ++v) a = fma(part[v * nmodes + k], w, a);
    out[k] = a;
}
cudaError_t launch_row_modes(const double *f, int n1, long long nv, int nmodes, double w, double *part, double *out,
                             cudaStream_t st) {
    k_row_modes<<<(unsigned)nv, 128, (size_t)n1 * 24, st>>>(f, n1, nmodes, part);
    count_launch();
    k_modes_reduce<<<(nmodes + 63) / 64, 64, 0, st>>>(part, nv, nmodes, w, out);
    count_launch();
    return cudaGetLastError();
}

// The diagnostics row of a time step: folds the last stage of the two reductions that feed it (the nb partial quadruples of
// the moments, the nn partial sums of the field energy) in a fixed order.  counters == NULL: time and destination as
// given.  Otherwise (time loop, possibly a CUDA-graph replay whose arguments are frozen): counters[0] = index of the row
// inside rows_base, counters[1] = number of the time step that has just been completed minus one; both are advanced
// here, so every launch writes the next row.
__global__ void k_sim4d_row(const double *__restrict__ m4p, const int nb, const double *__restrict__ nrjp, const int nn,
                            const double nrj_scale, const double time, const double dt, const double vol, const int root,
                            double *__restrict__ row6, double *__restrict__ rows_base, double *__restrict__ counters) {
    __shared__ double sh[5];
    const int t = threadIdx.x;
    if (t < 4) {
        double a = 0.0;
        for (int b = 0; b < nb; ++b) a += m4p[4 * b + t];
        sh[t] = a;
    } else if (t == 4) {
        double a = 0.0;
        for (int i = 0; i < nn; ++i) a += nrjp[i];
        sh[4] = a * nrj_scale;
    }
    __syncthreads();
    if (t != 0) return;
    double tm = time;
    if (counters) {
        row6 = rows_base + 6 * (long long)counters[0];
        tm = (counters[1] + 1.0) * dt;
        counters[0] += 1.0; counters[1] += 1.0;
    }
    row6[0] = root ? tm : 0.0; row6[1] = root ? sh[4] : 0.0;
    row6[2] = 0.5 * sh[3] * vol;
    row6[3] = sh[0] * vol; row6[4] = sh[1] * vol; row6[5] = sh[2] * vol;
}
cudaError_t launch_sim4d_row(const double *m4_partials, int nb, const double *nrj_parts, int nn, double nrj_scale, double time,
                             double dt, double vol, int root, double *row6, double *rows_base, double *counters, cudaStream_t st) {
    k_sim4d_row<<<1, 32, 0, st>>>(m4_partials, nb, nrj_parts, nn, nrj_scale, time, dt, vol, root, row6, rows_base, counters);
    count_launch();
    return cudaGetLastError();
}

} // namespace sllb
