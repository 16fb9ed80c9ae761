// sllb_lagrange_plane.cu -- K2d: the eta1 AND the eta2 pass of a Lagrange x-advection in ONE sweep over f.
//
// sll_s_advection_6d_lagrange_dd_slim_advect_eta1 / _eta2 (src/semi_lagrangian/advection/
// sll_m_advection_6d_lagrange_dd_slim.F90:806-967,1104-1267) run one after the other over the whole 6D array; both act
// inside the contiguous (eta1, eta2) planes and their displacements are constant over a plane (they depend on the
// conjugate velocity indices only), so a plane is read once, both stencils are applied in shared memory and the plane is
// written once: 16 B/point for two passes, and the contiguous-axis pass -- the slow one on 32-point lines -- disappears.
// Same weights, same left-to-right sums as the separate passes (bit-identical results).
//
// Block: 256 threads, planes handed out grid-stride; every plane arrives by ONE bulk TMA copy (it is contiguous) into one
// of two input buffers, the copy of the next plane overlapping the work on the current one (mbarrier per buffer).
#include <cstring>

#include "sllb_device.cuh"
#include "sllb_kernels.cuh"
#include "sllb_lagrange.cuh"

namespace sllb {

// Work split inside a block (256 threads), both passes with a sliding register window (one shared-memory load, S FMAs and
// one store per point instead of S loads with wrap arithmetic each):
//   pass A (along eta1): unit = (row j, segment of eta1); lanes run over rows, so the input rows sit at a pitch of n0 + 2
//           doubles (16-byte aligned for the row-wise TMA copies, at most 2-way bank conflicts), results go to `mid`;
//   pass B (along eta2): unit = (column i, segment of eta2); lanes run over columns: conflict-free reads of `mid` (pitch
//           n0 + 1) and coalesced 256-byte stores to global memory.
template <int S>
__global__ void __launch_bounds__(256) k_lagrange_plane(double *__restrict__ f, const int n0, const int n1,
                                                        const long long nplanes, const DispDesc dd0, const DispDesc dd1) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw); // bar[0], bar[1]
    const int np = n0 * n1, PI = n0 + 2, PM = n0 + 1;
    double *in0 = reinterpret_cast<double *>(smem_raw + 128);
    double *in1 = in0 + (size_t)n1 * PI;
    double *mid = in1 + (size_t)n1 * PI;
    double *wts = mid + (size_t)n1 * PM + 1;                  // [2][2][S]: per buffer parity, pass A / pass B weights
    int *offs = reinterpret_cast<int *>(wts + 4 * S);         // [2][2]
    const int tid = threadIdx.x;
    const uint32_t bytes = (uint32_t)((size_t)np * 8), rowbytes = (uint32_t)(n0 * 8);
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    __syncthreads();
    const long long G = gridDim.x;
    long long p = blockIdx.x;
    auto load_plane = [&](long long q, double *dst, uint64_t *b) { // row-wise bulk copies into the padded rows
        if (tid == 0) mbar_arrive_expect_tx(b, bytes);
        for (int j = tid; j < n1; j += 256) bulk_g2s(dst + (size_t)j * PI, f + q * (long long)np + (long long)j * n0, rowbytes, b);
    };
    if (p < nplanes) load_plane(p, in0, &bar[0]);
    if (p + G < nplanes) load_plane(p + G, in1, &bar[1]);
    // per-plane weights, computed ONCE per plane by one thread (the coefficient polynomials are ~100 fp64 operations:
    // recomputing them in every thread makes the kernel fp64-bound): the displacement of an eta1 line of plane q is that
    // of its first line (o = q*n1, in = 0), the displacement of an eta2 line that of (o = q, in = 0); both are constant
    // over the plane (checked by the launcher).  Plane q + G is prepared while plane q is in pass B.
    auto prepare = [&](long long q, int buf) {
        double pa_[S], pb_[S];
        offs[buf * 2 + 0] = lagr_setup<S>(disp_of(dd0, q * (long long)n1, 0), n0, pa_);
        offs[buf * 2 + 1] = lagr_setup<S>(disp_of(dd1, q, 0), n1, pb_);
#pragma unroll
        for (int k = 0; k < S; ++k) { wts[(buf * 2 + 0) * S + k] = pa_[k]; wts[(buf * 2 + 1) * S + k] = pb_[k]; }
    };
    if (tid == 255 && p < nplanes) prepare(p, 0);
    // segments: as many as give every thread a unit
    const int sega = (256 / n1 > 1) ? ((256 / n1 < n0) ? 256 / n1 : n0) : 1, la = (n0 + sega - 1) / sega;
    const int segb = (256 / n0 > 1) ? ((256 / n0 < n1) ? 256 / n0 : n1) : 1, lb = (n1 + segb - 1) / segb;
    __syncthreads();
    uint32_t phase[2] = {0, 0};
    int cur = 0;
    for (; p < nplanes; p += G, cur ^= 1) {
        double *in = cur ? in1 : in0;
        double pw[S], w[S];
#pragma unroll
        for (int k = 0; k < S; ++k) pw[k] = wts[(cur * 2 + 0) * S + k];
        const int oa = offs[cur * 2 + 0], ob = offs[cur * 2 + 1];
        mbar_wait(&bar[cur], phase[cur]);
        phase[cur] ^= 1;
        // pass A: along eta1
        for (int u = tid; u < n1 * sega; u += 256) {
            const int seg = u / n1, j = u - seg * n1;
            const int i0 = seg * la, i1 = (i0 + la < n0) ? i0 + la : n0;
            const double *row = in + (size_t)j * PI;
            double *dst = mid + (size_t)j * PM;
            int idx = i0 + oa; if (idx >= n0) idx -= n0;
#pragma unroll
            for (int k = 1; k < S; ++k) { w[k] = row[idx]; idx = (idx == n0 - 1) ? 0 : idx + 1; }
            for (int i = i0; i < i1; ++i) {
#pragma unroll
                for (int k = 0; k < S - 1; ++k) w[k] = w[k + 1];
                w[S - 1] = row[idx];
                idx = (idx == n0 - 1) ? 0 : idx + 1;
                double acc = pw[0] * w[0];
#pragma unroll
                for (int k = 1; k < S; ++k) acc = fma(pw[k], w[k], acc);
                dst[i] = acc;
            }
        }
        __syncthreads(); // mid complete, `in` free
        if (p + 2 * G < nplanes) load_plane(p + 2 * G, in, &bar[cur]);
        if (tid == 255 && p + G < nplanes) prepare(p + G, cur ^ 1);
        // pass B: along eta2, results straight to global memory (coalesced along eta1)
#pragma unroll
        for (int k = 0; k < S; ++k) pw[k] = wts[(cur * 2 + 1) * S + k];
        double *out = f + p * (long long)np;
        for (int u = tid; u < n0 * segb; u += 256) {
            const int seg = u / n0, i = u - seg * n0;
            const int j0 = seg * lb, j1 = (j0 + lb < n1) ? j0 + lb : n1;
            const double *col = mid + i;
            int idx = j0 + ob; if (idx >= n1) idx -= n1;
#pragma unroll
            for (int k = 1; k < S; ++k) { w[k] = col[(size_t)idx * PM]; idx = (idx == n1 - 1) ? 0 : idx + 1; }
            for (int j = j0; j < j1; ++j) {
#pragma unroll
                for (int k = 0; k < S - 1; ++k) w[k] = w[k + 1];
                w[S - 1] = col[(size_t)idx * PM];
                idx = (idx == n1 - 1) ? 0 : idx + 1;
                double acc = pw[0] * w[0];
#pragma unroll
                for (int k = 1; k < S; ++k) acc = fma(pw[k], w[k], acc);
                st_stream(out + i + (size_t)j * n0, acc);
            }
        }
        __syncthreads(); // mid and the weights of this parity are free for the next plane
    }
}

template <int S>
static cudaError_t launch_lagrange_plane_t(double *f, int n0, int n1, long long nplanes, const DispDesc &dd0,
                                           const DispDesc &dd1, cudaStream_t st) {
    const size_t smem = 128 + (2 * (size_t)n1 * (n0 + 2) + (size_t)n1 * (n0 + 1) + 1 + 4 * S) * 8 + 16;
    if (smem > 227 * 1024 || (size_t)n0 * n1 * 8 >= (1u << 20) || n0 % 2 != 0) return cudaErrorNotSupported; // 16-byte rows
    auto kern = k_lagrange_plane<S>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    long long grid = 148LL * per_sm;
    if (grid > nplanes) grid = nplanes;
    kern<<<(unsigned)grid, 256, smem, st>>>(f, n0, n1, nplanes, dd0, dd1);
    count_launch();
    return cudaGetLastError();
}

// cudaErrorNotSupported: plane too large / misaligned / displacements not constant over a plane / unknown stencil
cudaError_t launch_lagrange_plane(double *f, int n0, int n1, long long nplanes, int method, int order, const DispDesc &dd0,
                                  const DispDesc &dd1, cudaStream_t st) {
    if (n0 < 8 || n1 < 8 || nplanes < 1) return cudaErrorNotSupported;
    if ((reinterpret_cast<uintptr_t>(f) & 15) != 0 || ((size_t)n0 * n1) % 2 != 0) return cudaErrorNotSupported;
    // both displacements must be constant over a plane: dd0 may depend on (o / n1) only, dd1 on o only
    const bool c0 = (dd0.istr == 0 || dd0.imod == 1) && (dd0.ostr == 0 || dd0.omod == 1 || dd0.odiv % n1 == 0);
    const bool c1 = (dd1.istr == 0 || dd1.imod == 1);
    if (!c0 || !c1) return cudaErrorNotSupported;
    const bool odd = method == METHOD_LAGRANGE_FIXED, even = method == METHOD_LAGRANGE_CENTERED;
    if (!odd && !even) return cudaErrorNotSupported;
#define LP_CASE(SS) case SS: return launch_lagrange_plane_t<SS>(f, n0, n1, nplanes, dd0, dd1, st);
    if (odd) switch (order) { LP_CASE(3) LP_CASE(5) LP_CASE(7) LP_CASE(9) LP_CASE(11) default: return cudaErrorNotSupported; }
    switch (order) { LP_CASE(4) LP_CASE(6) LP_CASE(8) default: return cudaErrorNotSupported; }
}

} // namespace sllb
