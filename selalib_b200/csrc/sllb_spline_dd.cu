// sllb_spline_dd.cu -- K9: LOCAL cubic-spline advection with halo cells, the arithmetic of
// sll_m_cubic_spline_halo_1d (src/interpolation/interpolators/sll_m_cubic_spline_halo_1d.F90:69-200) as driven by
// sll_t_advection_6d_spline_dd_slim (src/semi_lagrangian/advection/sll_m_advection_6d_spline_dd_slim.F90:291-515,
// 976-1203): SURVEY.md section 8(f) rank 1.
//
// A line piece of np local points g(1..np) with integer shift si and fractional shift alpha in [0,1) is advected as
//     window      W(j) = g(si+j), j = 1..np+1                      (local cells + halo cells of ONE side)
//     forward     d(0) = d_0,  d(j) = (W(j) - b d(j-1))/a          (compute_interpolant :158-162)
//     backward    c(np+2) = c_np2,  c(j) = (d(j) - b c(j+1))/a     (:163-167)
//     evaluate    out(cell) = B-spline combination of c(cell-1..cell+2) at alpha, cell = 1..np   (eval_disp :176-200)
// where the two start values are series truncated after NUM_TERMS = 15 terms (:11):
//     d_0   = (1/a)  sum_{i=0..15} (-b/a)^i g(si-i)                            (prepare :83-91  + finish :118-122)
//     c_np2 = sqrt3 [g(J) + sum_{i=1..15} (-b/a)^i (g(J+i) + g(J-i))], J = si+np+2   (prepare :92-103 + finish :123-136)
// The terms that live on the ring neighbours are summed THERE (K9p, the reference's prepare_exchange) and travel
// as two scalars per line; the terms on this rank are added here (finish_boundary_conditions).  With a = sqrt((2+sqrt3)/6),
// q = b/a = 2-sqrt3 the recurrences are run on e = a d and G = a^2 c (one FMA per point each), 1/a^2 and the 1/6 of
// the evaluation folded into the four per-line weights, as in K1.
//
// WRAP = the axis is not split (procs(axis) == 1): the ring neighbour is the rank itself, so halo cells and remote sums
// are the periodic images of the line in shared memory and the whole pass is ONE kernel, 16 B/point.
// Split axis: K9p (remote sums, reads the edge planes) -> exchange of sums + halo planes -> K9 with halo rows staged
// next to the local rows.
#include <climits>
#include <cstring>

#include "sllb_device.cuh"
#include "sllb_kernels.cuh"
#include "sllb_spline15.cuh"

namespace sllb {

__device__ __forceinline__ long long disp_index(const DispDesc &d, long long o, long long in) {
    return ((o / d.odiv) % d.omod) * d.ostr + ((in / d.idiv) % d.imod) * d.istr;
}

// ------------------------------------------------------------------------------------------------
// K9a: strided axis (inner > 1).  Block = one warp = 32 adjacent lines; tile rows are 256-byte bulk TMA copies
// (left halo rows | local rows | right halo rows), one thread per line.
// ------------------------------------------------------------------------------------------------
template <bool WRAP>
__global__ void __launch_bounds__(32) k_spline_dd_strided(double *__restrict__ f, const double *__restrict__ hl,
                                                          const double *__restrict__ hr, const double *__restrict__ bcl,
                                                          const double *__restrict__ bcr, const int *__restrict__ shift,
                                                          const long long nlines, const int np, const long long inner,
                                                          const int hwl, const int hwr, const DispDesc dd,
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
    const int R = np + hwl + hwr;
    const double *lb = WRAP ? nullptr : hl + o * (long long)hwl * inner + in;
    const double *rb = WRAP ? nullptr : hr + o * (long long)hwr * inner + in;
    if (use_tma) {
        if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(R * BW * 8));
        for (int j = tid; j < R; j += BW) {
            const double *src = (j < hwl) ? lb - tid + (long long)j * inner
                              : (j < hwl + np) ? base - tid + (long long)(j - hwl) * inner
                                               : rb - tid + (long long)(j - hwl - np) * inner;
            bulk_g2s(s + (size_t)j * BW, src, BW * 8, bar);
        }
        mbar_wait(bar, 0);
    } else {
        if (active) {
            for (int j = 0; j < hwl; ++j) cp_async8(s + (size_t)j * BW + tid, lb + (long long)j * inner);
            for (int j = 0; j < np; ++j) cp_async8(s + (size_t)(j + hwl) * BW + tid, base + (long long)j * inner);
            for (int j = 0; j < hwr; ++j) cp_async8(s + (size_t)(j + hwl + np) * BW + tid, rb + (long long)j * inner);
        }
        cp_async_wait_all();
    }
    if (!active) return;
    const long long di = disp_index(dd, o, in);
    const double disp = dd.scale * dd.v[di];
    const double fl = floor(disp);
    int si = shift ? shift[di] : (int)fl;
    if (si == INT_MIN) return; // the line belongs to no displacement block: left untouched (make_blocks_spline :219-262)
    if (!WRAP) si = max(-hwl, min(si, hwr - 1)); // the halo covers si in [-hwl, hwr-1] (the reference assumes it, :1104-1110)
    double *x0 = s + (size_t)hwl * BW + tid;
    double sum_d, sum_c;
    spline15_sums<BW, WRAP>(x0, np, si, WRAP ? 0.0 : bcl[l], WRAP ? 0.0 : bcr[l], &sum_d, &sum_c);
    spline15_line<BW, WRAP, true>(x0, np, si, disp - fl, sum_d, sum_c, base, inner);
}

// ------------------------------------------------------------------------------------------------
// K9b: contiguous axis (inner == 1, never split: sll_f_set_process_grid splits eta1 last).  Block = BW consecutive
// lines = one contiguous chunk, transposed on the way in (pitch BW+1); results leave through coalesced stores that
// undo the slot rotation of spline15_line.
// ------------------------------------------------------------------------------------------------
template <int BW>
__global__ void __launch_bounds__(BW) k_spline_dd_contig(double *__restrict__ f, const int *__restrict__ shift,
                                                         const long long nlines, const int N, const DispDesc dd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int *rot = reinterpret_cast<int *>(smem_raw); // BW ints: slot of output 0 per line, -1 = untouched line
    double *s = reinterpret_cast<double *>(smem_raw + 128 + ((BW * 4) / 128) * 128);
    constexpr int P = BW + 1;
    const int tid = threadIdx.x;
    const long long l0 = (long long)blockIdx.x * BW;
    const int nl = (int)((nlines - l0 < BW) ? (nlines - l0) : BW);
    double *tile = f + l0 * (long long)N;
    for (int ln = 0; ln < nl; ++ln)
        for (int j = tid; j < N; j += BW) cp_async8(s + (size_t)j * P + ln, tile + (long long)ln * N + j);
    cp_async_wait_all();
    __syncthreads();
    if (tid < nl) {
        const long long di = disp_index(dd, l0 + tid, 0);
        const double disp = dd.scale * dd.v[di];
        const double fl = floor(disp);
        const int si = shift ? shift[di] : (int)fl;
        if (si == INT_MIN) rot[tid] = -1;
        else {
            rot[tid] = wrap_idx(si, N);
            double sum_d, sum_c;
            spline15_sums<P, true>(s + tid, N, si, 0.0, 0.0, &sum_d, &sum_c);
            spline15_line<P, true, false>(s + tid, N, si, disp - fl, sum_d, sum_c, nullptr, 0);
        }
    }
    __syncthreads();
    for (int ln = 0; ln < nl; ++ln) {
        const int r = rot[ln];
        if (r < 0) continue;
        for (int i = tid; i < N; i += BW) {
            int kc = i + r;
            if (kc >= N) kc -= N;
            st_stream(tile + (long long)ln * N + i, s[(size_t)kc * P + ln]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K9p: sll_s_cubic_spline_halo_1d_prepare_exchange for every line of a split axis: the parts of the neighbours'
// boundary sums that consist of MY cells.
//   for_right[line] = sum_{i=0..15, si-1-i < 0}  (-q)^i f(np + si-1-i)      -> the right neighbour's d_0   (:78-91)
//   for_left[line]  = sum_{m=-15..15, si+1+m >= 0} (-q)^|m| f(si+1+m)       -> the left neighbour's c_np2  (:92-103)
// Reads the top 16+ and bottom 17+ planes straight from global memory (coalesced across the 32 lines of a warp).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_spline_dd_prepare(const double *__restrict__ f, const int *__restrict__ shift,
                                                           const long long nlines, const int np, const long long inner,
                                                           const int hwl, const int hwr, const DispDesc dd,
                                                           double *__restrict__ for_right, double *__restrict__ for_left) {
    const long long l = (long long)blockIdx.x * 128 + threadIdx.x;
    if (l >= nlines) return;
    const long long o = l / inner, in = l - o * inner;
    const double *base = f + o * (long long)np * inner + in;
    const long long di = disp_index(dd, o, in);
    const double disp = dd.scale * dd.v[di];
    int si = shift ? shift[di] : (int)floor(disp);
    if (si == INT_MIN) si = 0;
    si = max(-hwl, min(si, hwr - 1));
    double sd, sc;
    spline15_prepare(base, inner, np, si, &sd, &sc);
    for_right[l] = sd;
    for_left[l] = sc;
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static bool g_hpw_ready = false;
static cudaError_t ensure_halo_constants() {
    if (g_hpw_ready) return cudaSuccess;
    double pw[SLLB_HALO_TERMS + 1];
    const double a = sqrt((2.0 + sqrt(3.0)) / 6.0), b = sqrt((2.0 - sqrt(3.0)) / 6.0);
    for (int i = 0; i <= SLLB_HALO_TERMS; ++i) pw[i] = pow(-(b / a), (double)i); // pba_pow (:59-64)
    cudaError_t e = cudaMemcpyToSymbol(c_hpw, pw, sizeof(pw));
    if (e == cudaSuccess) g_hpw_ready = true;
    return e;
}
static const size_t SMEM_MAX_DD = 227 * 1024;

int spline_dd_min_points(int hwl, int hwr) {
    // every neighbour sum must stay inside the neighbour's block: top cells np+si-16 >= 0 and bottom cells si+16 <= np-1
    const int h = hwl > hwr ? hwl : hwr;
    return SLLB_HALO_TERMS + 2 + h;
}

cudaError_t launch_spline_dd(double *f, long long outer, int np, long long inner, const DispDesc &dd, const int *d_shift,
                             const double *halo_l, int hwl, const double *halo_r, int hwr, const double *bc_l,
                             const double *bc_r, int staging, cudaStream_t st) {
    if (np <= SLLB_HALO_TERMS || outer < 1 || inner < 1) return cudaErrorInvalidValue; // SLL_ASSERT_ALWAYS(num_points > NUM_TERMS)
    cudaError_t e = ensure_halo_constants();
    if (e != cudaSuccess) return e;
    const long long nlines = outer * inner;
    const bool wrap = halo_l == nullptr;
    if (wrap) { hwl = 0; hwr = 0; }
    else if (!halo_r || !bc_l || !bc_r || hwl < 0 || hwr < 0 || hwl + hwr < 1 || np < spline_dd_min_points(hwl, hwr)) return cudaErrorInvalidValue;
    if (inner == 1) {
        if (!wrap) return cudaErrorNotSupported;
#define SLLB_DD_CONTIG(BWV)                                                                  \
    do {                                                                                     \
        const size_t smem = 128 + ((BWV * 4) / 128) * 128 + (size_t)np * (BWV + 1) * 8;      \
        if (smem <= SMEM_MAX_DD) {                                                           \
            auto kern = k_spline_dd_contig<BWV>;                                             \
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                  \
            kern<<<(unsigned)((nlines + BWV - 1) / BWV), BWV, smem, st>>>(f, d_shift, nlines, np, dd); \
            count_launch();                                                                  \
            return cudaGetLastError();                                                       \
        }                                                                                    \
    } while (0)
        SLLB_DD_CONTIG(32);
        SLLB_DD_CONTIG(16);
        SLLB_DD_CONTIG(8);
        return cudaErrorInvalidValue;
    }
    const int R = np + hwl + hwr;
    const size_t smem = 128 + (size_t)R * 32 * 8;
    if (smem > SMEM_MAX_DD) return cudaErrorInvalidValue;
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool tma_ok = (inner % 32 == 0) && al16(f) && (wrap || (al16(halo_l) && al16(halo_r)));
    const int use_tma = (staging == STAGING_CPASYNC) ? 0 : (tma_ok ? 1 : 0);
    const long long nblk = (nlines + 31) / 32;
    if (wrap) {
        auto kern = k_spline_dd_strided<true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)nblk, 32, smem, st>>>(f, nullptr, nullptr, nullptr, nullptr, d_shift, nlines, np, inner, 0, 0, dd, use_tma);
    } else {
        auto kern = k_spline_dd_strided<false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)nblk, 32, smem, st>>>(f, halo_l, halo_r, bc_l, bc_r, d_shift, nlines, np, inner, hwl, hwr, dd, use_tma);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_spline_dd_prepare(const double *f, long long outer, int np, long long inner, const DispDesc &dd,
                                     const int *d_shift, int hwl, int hwr, double *for_right, double *for_left,
                                     cudaStream_t st) {
    if (np < spline_dd_min_points(hwl, hwr) || outer < 1 || inner < 1) return cudaErrorInvalidValue;
    cudaError_t e = ensure_halo_constants();
    if (e != cudaSuccess) return e;
    const long long nlines = outer * inner;
    k_spline_dd_prepare<<<(unsigned)((nlines + 127) / 128), 128, 0, st>>>(f, d_shift, nlines, np, inner, hwl, hwr, dd, for_right,
                                                                          for_left);
    count_launch();
    return cudaGetLastError();
}

} // namespace sllb
