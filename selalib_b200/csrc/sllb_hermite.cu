// sllb_hermite.cu -- K10: cubic-spline interpolation with Hermite boundary conditions along a NON-periodic axis
// (velocity in simulations/parallel/bsl_vp_2d2v_cart/sll_m_sim_bsl_vp_2d2v_cart.F90:470-482,520-545): every line of np
// grid points is replaced by S(x_i + alpha), the spline's end slopes taken from 5-point one-sided differences of the data
// (or given), feet beyond the interval taking the boundary value.  SURVEY.md section 8(f) rank 3.
// Arithmetic: sllb_hermite.cuh.  Tiles as K9: one warp = 32 adjacent lines of a strided axis (256-byte TMA rows, thread
// per line), or 32 consecutive lines of the contiguous axis transposed into shared memory.  16 B/point.
#include <cstring>

#include "sllb_device.cuh"
#include "sllb_kernels.cuh"
#include "sllb_hermite.cuh"

namespace sllb {

struct HermiteArgs {
    double delta;        // cell size of the axis, (xmax - xmin)/(np - 1)
    int inplace;         // 1: interpolate_array_disp_inplace semantics (clamped feet), 0: interpolate_array_disp
    int have_slopes;
    double sl, sr;
};

__device__ __forceinline__ long long hdisp_index(const DispDesc &d, long long o, long long in) {
    return ((o / d.odiv) % d.omod) * d.ostr + ((in / d.idiv) % d.imod) * d.istr;
}

__global__ void __launch_bounds__(32) k_hermite_strided(double *__restrict__ f, const long long nlines, const int np,
                                                        const long long inner, const DispDesc dd, const HermiteArgs ha,
                                                        const int use_tma) {
    constexpr int BW = 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *s = reinterpret_cast<double *>(smem_raw + 128);
    const int tid = threadIdx.x;
    const long long l = (long long)blockIdx.x * BW + tid;
    const bool active = l < nlines;
    const long long o = active ? l / inner : 0, in = active ? l - o * inner : 0;
    double *base = f + o * (long long)np * inner + in;
    if (use_tma) {
        if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(np * BW * 8));
        for (int j = tid; j < np; j += BW) bulk_g2s(s + (size_t)j * BW, base - tid + (long long)j * inner, BW * 8, bar);
        mbar_wait(bar, 0);
    } else {
        if (active)
            for (int j = 0; j < np; ++j) cp_async8(s + (size_t)j * BW + tid, base + (long long)j * inner);
        cp_async_wait_all();
    }
    if (!active) return;
    const double alpha0 = dd.scale * dd.v[hdisp_index(dd, o, in)] / ha.delta; // displacement in cells
    double *x0 = s + tid;
    double g0, gnp1;
    hermite_coeffs_line<BW>(x0, np, ha.delta, ha.have_slopes, ha.sl, ha.sr, &g0, &gnp1);
#pragma unroll 2
    for (int i = 1; i <= np; ++i)
        st_stream(base + (long long)(i - 1) * inner, hermite_eval_point<BW>(x0, np, i, alpha0, ha.inplace, g0, gnp1));
}

template <int BW>
__global__ void __launch_bounds__(BW) k_hermite_contig(double *__restrict__ f, const long long nlines, const int np,
                                                       const DispDesc dd, const HermiteArgs ha) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int P = BW + 1;
    double *s = reinterpret_cast<double *>(smem_raw);   // coefficients [np][P]
    double *r = s + (size_t)np * P;                      // results      [np][P]
    const int tid = threadIdx.x;
    const long long l0 = (long long)blockIdx.x * BW;
    const int nl = (int)((nlines - l0 < BW) ? (nlines - l0) : BW);
    double *tile = f + l0 * (long long)np;
    for (int ln = 0; ln < nl; ++ln)
        for (int j = tid; j < np; j += BW) cp_async8(s + (size_t)j * P + ln, tile + (long long)ln * np + j);
    cp_async_wait_all();
    __syncthreads();
    if (tid < nl) {
        const double alpha0 = dd.scale * dd.v[hdisp_index(dd, l0 + tid, 0)] / ha.delta;
        double *x0 = s + tid;
        double g0, gnp1;
        hermite_coeffs_line<P>(x0, np, ha.delta, ha.have_slopes, ha.sl, ha.sr, &g0, &gnp1);
        for (int i = 1; i <= np; ++i) r[(size_t)(i - 1) * P + tid] = hermite_eval_point<P>(x0, np, i, alpha0, ha.inplace, g0, gnp1);
    }
    __syncthreads();
    for (int ln = 0; ln < nl; ++ln)
        for (int i = tid; i < np; i += BW) st_stream(tile + (long long)ln * np + i, r[(size_t)i * P + ln]);
}

static bool g_hq_ready = false;
static cudaError_t ensure_hermite_constants() {
    if (g_hq_ready) return cudaSuccess;
    double pw[SLLB_HERMITE_TERMS];
    const double a = sqrt((2.0 + sqrt(3.0)) / 6.0), b = sqrt((2.0 - sqrt(3.0)) / 6.0);
    double ct = 1.0;
    for (int i = 0; i < SLLB_HERMITE_TERMS; ++i) { pw[i] = ct; ct *= -(b / a); } // coeff_tmp*(-b_a) (:630)
    cudaError_t e = cudaMemcpyToSymbol(c_hq, pw, sizeof(pw));
    if (e == cudaSuccess) g_hq_ready = true;
    return e;
}

// every line of f viewed as [outer][np][inner]; disp in PHYSICAL units (alpha of interpolate_array_disp), delta = cell size
cudaError_t launch_hermite(double *f, long long outer, int np, long long inner, const DispDesc &dd, double delta, int inplace,
                           int have_slopes, double sl, double sr, int staging, cudaStream_t st) {
    if (np < SLLB_HERMITE_TERMS || outer < 1 || inner < 1 || !(delta > 0.0)) return cudaErrorInvalidValue; // fast algorithm only (:266)
    cudaError_t e = ensure_hermite_constants();
    if (e != cudaSuccess) return e;
    HermiteArgs ha;
    ha.delta = delta; ha.inplace = inplace; ha.have_slopes = have_slopes; ha.sl = sl; ha.sr = sr;
    const long long nlines = outer * inner;
    const size_t SMAX = 227 * 1024;
    if (inner == 1) {
#define SLLB_HERM_CONTIG(BWV)                                                                 \
    do {                                                                                      \
        const size_t smem = 2 * (size_t)np * (BWV + 1) * 8;                                   \
        if (smem <= SMAX) {                                                                   \
            auto kern = k_hermite_contig<BWV>;                                                \
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                   \
            kern<<<(unsigned)((nlines + BWV - 1) / BWV), BWV, smem, st>>>(f, nlines, np, dd, ha); \
            count_launch();                                                                   \
            return cudaGetLastError();                                                        \
        }                                                                                     \
    } while (0)
        SLLB_HERM_CONTIG(32);
        SLLB_HERM_CONTIG(16);
        SLLB_HERM_CONTIG(8);
        return cudaErrorInvalidValue;
    }
    const size_t smem = 128 + (size_t)np * 32 * 8;
    if (smem > SMAX) return cudaErrorInvalidValue;
    const bool tma_ok = (inner % 32 == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0);
    const int use_tma = (staging == STAGING_CPASYNC) ? 0 : (tma_ok ? 1 : 0);
    e = cudaFuncSetAttribute(k_hermite_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_hermite_strided<<<(unsigned)((nlines + 31) / 32), 32, smem, st>>>(f, nlines, np, inner, dd, ha, use_tma);
    count_launch();
    return cudaGetLastError();
}

} // namespace sllb
