// sllb_kernels.cu -- hand-written sm_100a kernels of the split semi-Lagrangian advection path.
//
// K1  periodic cubic-spline shift     (replaces compute_spline_1D_periodic_aux + eval_disp,
//                                      sll_m_cubic_splines.F90:531-581,2616-2711, per line)
// K2  Lagrange fixed/centred shift    (sll_m_lagrange_interpolation_1d_fast.F90:609-837)
// K3  velocity reduction              (sll_m_reduction.F90:187-272, sll_m_sim_6d_utilities.F90:203-245)
// K4  spectral Poisson multipliers    (sll_m_poisson_{1d,2d}_periodic.F90, sll_m_poisson_3d_periodic_par.F90)
// K6  pack / unpack for remaps        (apply_remap_4D_double, sll_m_remapper.F90:3386-3395,3444-3453)
// K8  diagnostics row sums
//
// Every advection kernel stages whole lines of f in shared memory (TMA bulk copies completing on an
// mbarrier, or cp.async when the tile is not 16-byte tileable), solves / evaluates in the block and
// writes each point once: 16 B of HBM traffic per point per pass.  All arithmetic is fp64.
#include "sllb_kernels.cuh"
#include "sllb_device.cuh"
#include "sllb_lagrange.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace sllb {

static long long g_launches = 0;
long long launch_count() { return g_launches; }
void launch_count_reset() { g_launches = 0; }
void count_launch() { ++g_launches; }
void count_launches(long long n) { g_launches += n; }
#define COUNT_LAUNCH() (++g_launches)

// ------------------------------------------------------------------------------------------------
// Where a pass writes its output.  In place: point iout of line (o, in) goes to f[(o*N + iout)*inner + in].
// Fused remap (RemapDst.on): the pass that ends a splitting stage writes every point straight into the
// OTHER layout's array on whichever rank owns it there (peer-mapped pointers over NVLink), i.e. the
// reference's pack + MPI_Alltoall + unpack (apply_remap_4D_double, sll_m_remapper.F90:3308-3456) happens in
// the store of the advection kernel.  Uniform boxes only (every axis divisible by its process count).
// The cursor walks the advected-axis index up or down with periodic wrap; `li` is the index inside the
// destination rank's block, `pc` that rank's coordinate along the advected axis.
// ------------------------------------------------------------------------------------------------
struct OutMap {
    const RemapDst *rd;
    long long off0, os;   // element offset of (line, li = 0) in the destination array; stride per li
    int r0, rs, nl, npc;  // rank of the line's fixed coordinates, rank stride per pc, block width, blocks along the axis
    int pc, li;
    double *p;
    __device__ __forceinline__ void seek(int iout) {
        pc = iout / nl; li = iout - pc * nl;
        p = rd->base[r0 + pc * rs] + off0 + (long long)li * os;
    }
    __device__ __forceinline__ void prev() {
        if (li > 0) { --li; p -= os; }
        else { pc = (pc == 0) ? npc - 1 : pc - 1; li = nl - 1; p = rd->base[r0 + pc * rs] + off0 + (long long)li * os; }
    }
    __device__ __forceinline__ void next() {
        if (li < nl - 1) { ++li; p += os; }
        else { pc = (pc == npc - 1) ? 0 : pc + 1; li = 0; p = rd->base[r0 + pc * rs] + off0; }
    }
};
__device__ __forceinline__ OutMap make_outmap(const RemapDst &rd, const long long o, const long long in, const int N,
                                              const long long inner) {
    OutMap m;
    m.rd = &rd;
    if (!rd.on) {
        m.off0 = o * (long long)N * inner + in; m.os = inner;
        m.r0 = 0; m.rs = 0; m.nl = N; m.npc = 1;
    } else {
        const int a = rd.axis;
        long long ri = in, ro = o, off = 0, ostr = 1;
        int rank = 0, rstr = 1;
        m.os = 1; m.rs = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (d == a) { m.os = ostr; m.rs = rstr; }
            else {
                int c;
                if (d < a) { c = (int)(ri % rd.se[d]); ri /= rd.se[d]; }
                else { c = (int)(ro % rd.se[d]); ro /= rd.se[d]; }
                c += rd.slo[d];
                const int pcd = c / rd.te[d];
                off += (long long)(c - pcd * rd.te[d]) * ostr;
                rank += pcd * rstr;
            }
            ostr *= rd.te[d]; rstr *= rd.tp[d];
        }
        m.off0 = off; m.r0 = rank; m.nl = rd.te[a]; m.npc = rd.tp[a];
    }
    m.pc = 0; m.li = 0; m.p = nullptr;
    return m;
}

// ------------------------------------------------------------------------------------------------
// Periodic cubic spline on one line held in shared memory (elements at sc[k*PITCH]).
//
// Reference recurrences (sll_m_cubic_splines.F90:560-580), a = sqrt((2+sqrt3)/6), b = sqrt((2-sqrt3)/6):
//     d(i) = (f(i) - b d(i-1))/a,   c(i) = (d(i) - b c(i+1))/a,   wrap terms = 27-term series in -b/a.
// Scaled by 1/a they become e(i) = f(i) - q e(i-1), g(i) = e(i) - q g(i+1) with q = b/a = 2 - sqrt3 and
// c = g/a^2; the factor 1/a^2 = 6(2-sqrt3) and the 1/6 of the B-spline evaluation are folded into the four
// per-line weights, so the whole pass costs 6 FMA per point.
// ------------------------------------------------------------------------------------------------
#define SLLB_NUM_TERMS 27
__constant__ double c_pw[SLLB_NUM_TERMS]; // (-q)^(i+1)

template <int PITCH, bool TO_GLOBAL>
__device__ __forceinline__ void spline_line(double *sc, const int N, const double disp, OutMap *om) {
    const double q = 0.26794919243112270647; // 2 - sqrt(3)
    const double r2 = 1.60769515458673623883; // 6 (2 - sqrt 3) = 1/a^2
    const double fl = floor(disp);
    const int dcell = (int)fl;
    const double dx = disp - fl, cdx = 1.0 - dx;
    const double s6 = r2 * (1.0 / 6.0);
    const double w0 = cdx * cdx * cdx * s6;
    const double w1 = (1.0 + 3.0 * cdx + 3.0 * cdx * cdx - 3.0 * cdx * cdx * cdx) * s6;
    const double w2 = (1.0 + 3.0 * dx + 3.0 * dx * dx - 3.0 * dx * dx * dx) * s6;
    const double w3 = dx * dx * dx * s6;

    // forward sweep: e_1 = f_1 + sum_{i=0..26} (-q)^{i+1} f_{N-i}
    double e = sc[0];
    {
        int idx = N - 1;
#pragma unroll
        for (int i = 0; i < SLLB_NUM_TERMS; ++i) {
            e = fma(c_pw[i], sc[idx * PITCH], e);
            idx = (idx == 0) ? N - 1 : idx - 1;
        }
    }
    sc[0] = e;
#pragma unroll 8
    for (int k = 1; k < N; ++k) {
        e = fma(-q, e, sc[k * PITCH]);
        sc[k * PITCH] = e;
    }
    // backward sweep: g_N = e_N + sum_{i=1..27} (-q)^i e_i
    double g = e;
    {
        int idx = 0;
#pragma unroll
        for (int i = 0; i < SLLB_NUM_TERMS; ++i) {
            g = fma(c_pw[i], sc[idx * PITCH], g);
            idx = (idx == N - 1) ? 0 : idx + 1;
        }
    }
    const double gt0 = g;                                   // g[N-1]
    const double gt1 = fma(-q, gt0, sc[(N - 2) * PITCH]);   // g[N-2]
    const double gt2 = fma(-q, gt1, sc[(N - 3) * PITCH]);   // g[N-3]
    double a3 = gt0, a2 = gt1, a1 = gt2;
    // output index of cell kc is (kc - dcell) mod N; cells are emitted kc = N-3, N-4, ..., 1
    if (TO_GLOBAL) om->seek(((N - 3 - dcell) % N + N) % N);
#pragma unroll 8
    for (int k = N - 4; k >= 0; --k) {
        const double a0 = fma(-q, a1, sc[k * PITCH]);
        const double val = fma(w3, a3, fma(w2, a2, fma(w1, a1, w0 * a0))); // cell k+1
        if (TO_GLOBAL) { st_stream(om->p, val); om->prev(); }
        else sc[(k + 1) * PITCH] = val; // slot k+1 already consumed
        a3 = a2; a2 = a1; a1 = a0;
    }
    // now a1 = g[0], a2 = g[1], a3 = g[2]
    const double v0 = fma(w3, a3, fma(w2, a2, fma(w1, a1, w0 * gt0)));   // cell 0
    const double vm1 = fma(w3, a2, fma(w2, a1, fma(w1, gt0, w0 * gt1))); // cell N-1
    const double vm2 = fma(w3, a1, fma(w2, gt0, fma(w1, gt1, w0 * gt2))); // cell N-2
    if (TO_GLOBAL) {
        st_stream(om->p, v0); om->prev();
        st_stream(om->p, vm1); om->prev();
        st_stream(om->p, vm2);
    } else {
        sc[0] = v0;
        sc[(N - 1) * PITCH] = vm1;
        sc[(N - 2) * PITCH] = vm2;
    }
}

// one line from shared memory (sc[k*PITCH]) to global memory with a register window
template <int S, int PITCH>
__device__ __forceinline__ void lagrange_line(const double *sc, const int N, const double disp, OutMap *om) {
    double pp[S], w[S];
    int idx = lagr_setup<S>(disp, N, pp);
#pragma unroll
    for (int k = 1; k < S; ++k) {
        w[k] = sc[idx * PITCH];
        idx = (idx == N - 1) ? 0 : idx + 1;
    }
    om->seek(0);
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int k = 0; k < S - 1; ++k) w[k] = w[k + 1];
        w[S - 1] = sc[idx * PITCH];
        idx = (idx == N - 1) ? 0 : idx + 1;
        double acc = pp[0] * w[0]; // left-to-right sum like lagr_Npt_vec (:393-399)
#pragma unroll
        for (int k = 1; k < S; ++k) acc = fma(pp[k], w[k], acc);
        st_stream(om->p, acc);
        om->next();
    }
}

// ------------------------------------------------------------------------------------------------
// K1d: periodic spline of order 6 / 8 (degree 5 / 7) on one line held in shared memory: sll_s_periodic_interp with
// sll_p_spline (sll_m_periodic_interp.F90:108-119,202-229).  The reference divides the spectrum by the eigenvalues of the
// circulant collocation matrix M (B-spline values at the nodes: (1,26,66,26,1)/120, (1,120,1191,2416,1191,120,1)/5040)
// and multiplies by those of the shifted evaluation.  Here M^-1 is applied in place as the exact factorisation of M into
// first-order recursive filters, one causal + one anticausal sweep per root z of its symbol inside the unit circle
//     1/m(z) = p! prod_i  -z_i / ((1 - z_i/z)(1 - z_i z)),
// each started from the periodic sum of the line (|z|^72 < 3e-20: truncated for long lines, closed with 1/(1 - z^N) for
// short ones) -- the same scheme as the cubic kernels, where the single root is sqrt(3) - 2.  Then
// out(k) = sum_j b_j(beta) c(k + floor(disp) + j), j = -(p-1)/2 .. (p+1)/2, b = sll_s_uniform_bsplines_eval_basis(p, beta)
// (sll_m_low_level_bsplines.F90:880-915), as a register window sliding along the line.
// ------------------------------------------------------------------------------------------------
template <int ORDER>
__device__ __forceinline__ void bspline_basis(const double off, double *bspl) {
    bspl[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= ORDER - 1; ++j) {
        double xx = -off;
        const double j_real = (double)j, inv_j = 1.0 / j_real;
        double saved = 0.0;
#pragma unroll
        for (int r = 0; r <= j - 1; ++r) {
            xx += 1.0;
            const double temp = bspl[r] * inv_j;
            bspl[r] = saved + xx * temp;
            saved = (j_real - xx) * temp;
        }
        bspl[j] = saved;
    }
}
template <int ORDER, int PITCH>
__device__ __forceinline__ void bspline_line(double *sc, const int N, const double disp, OutMap *om) {
    static_assert(ORDER == 6 || ORDER == 8, "periodic splines of order 6 and 8");
    constexpr int NP = (ORDER - 2) / 2;
    // roots of z^2 + 26 z + 66 + 26/z + 1/z^2 and of z^3 + 120 z^2 + 1191 z + 2416 + ... inside the unit circle
    const double poles[3] = {ORDER == 6 ? -0.43057534709997379185 : -0.53528043079643816554,
                             ORDER == 6 ? -0.043096288203264654247 : -0.12255461519232669052, -0.0091486948096082769286};
    const double gain = ORDER == 6 ? 120.0 : 5040.0;
    const int K = N < 72 ? N : 72;
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) {
        const double z = poles[ip];
        double zp = 1.0, acc = 0.0;
        int idx = 0;
        for (int m = 0; m < K; ++m) {          // c+(0) = sum_m z^m s(-m)
            acc = fma(zp, sc[idx * PITCH], acc);
            zp *= z;
            idx = (idx == 0) ? N - 1 : idx - 1;
        }
        const double closing = (K == N) ? 1.0 / (1.0 - zp) : 1.0;   // zp = z^N when the whole line was summed
        double e = acc * closing;
        sc[0] = e;
        for (int k = 1; k < N; ++k) {
            e = fma(z, e, sc[k * PITCH]);
            sc[k * PITCH] = e;
        }
        zp = z; acc = 0.0; idx = N - 1;
        for (int m = 0; m < K; ++m) {          // c(N-1) = -sum_m z^(m+1) c+(N-1+m)
            acc = fma(zp, sc[idx * PITCH], acc);
            zp *= z;
            idx = (idx == N - 1) ? 0 : idx + 1;
        }
        double c = -acc * closing;
        sc[(N - 1) * PITCH] = c;
        for (int k = N - 2; k >= 0; --k) {
            c = z * (c - sc[k * PITCH]);
            sc[k * PITCH] = c;
        }
    }
    const double fl = floor(disp);
    double pp[ORDER], w[ORDER];
    bspline_basis<ORDER>(disp - fl, pp);
#pragma unroll
    for (int k = 0; k < ORDER; ++k) pp[k] *= gain;
    int idx = (int)(((long long)fl - (ORDER - 2) / 2) % N);
    if (idx < 0) idx += N;
#pragma unroll
    for (int k = 1; k < ORDER; ++k) {
        w[k] = sc[idx * PITCH];
        idx = (idx == N - 1) ? 0 : idx + 1;
    }
    om->seek(0);
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int k = 0; k < ORDER - 1; ++k) w[k] = w[k + 1];
        w[ORDER - 1] = sc[idx * PITCH];
        idx = (idx == N - 1) ? 0 : idx + 1;
        double a = pp[0] * w[0];
#pragma unroll
        for (int k = 1; k < ORDER; ++k) a = fma(pp[k], w[k], a);
        st_stream(om->p, a);
        om->next();
    }
}

// ------------------------------------------------------------------------------------------------
// K1a/K2a: strided axis (inner > 1).  Block = BW adjacent lines (adjacent in the flattened faster
// axes, i.e. contiguous in memory row by row); shared tile s[k*BW + t] = f(line t, point k).
// METHOD 0 = spline, 1 = Lagrange with stencil S.
// ------------------------------------------------------------------------------------------------
template <int BW, int METHOD, int S>
__global__ void __launch_bounds__(BW) k_advect_strided(double *__restrict__ f, const long long nlines, const int N,
                                                        const long long inner, const DispDesc dd, const int use_tma,
                                                        const __grid_constant__ RemapDst rd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *s = reinterpret_cast<double *>(smem_raw + 128);
    const int tid = threadIdx.x;
    const long long l = (long long)blockIdx.x * BW + tid;
    const bool active = l < nlines;
    const long long o = active ? l / inner : 0, in = active ? l - o * inner : 0;
    double *base = f + o * (long long)N * inner + in;

    if (use_tma) { // whole tile inside one `o` and 16-byte tileable (checked by the launcher)
        if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(N * BW * 8));
        const double *src0 = base - tid; // line of thread 0
        for (int j = tid; j < N; j += BW) bulk_g2s(s + (size_t)j * BW, src0 + (long long)j * inner, BW * 8, bar);
        mbar_wait(bar, 0);
    } else {
        if (active)
            for (int j = 0; j < N; ++j) cp_async8(s + (size_t)j * BW + tid, base + (long long)j * inner);
        cp_async_wait_all();
    }
    if (!active) return;
    const double disp = disp_of(dd, o, in);
    OutMap om = make_outmap(rd, o, in, N, inner);
    if constexpr (METHOD == 0) spline_line<BW, true>(s + tid, N, disp, &om);
    else if constexpr (METHOD == 2) bspline_line<S, BW>(s + tid, N, disp, &om);
    else lagrange_line<S, BW>(s + tid, N, disp, &om);
}

// ------------------------------------------------------------------------------------------------
// K2a (split): Lagrange on a strided axis with every line cut into P chunks, one group of BW lanes per chunk
// (block = BW lines x P chunks).  Same tile and arithmetic as k_advect_strided; used when one thread per line would
// leave the GPU mostly idle (few lines, long lines: the 1D1V problems) -- the stencil needs no sweep, so the chunks
// are independent once the whole lines sit in shared memory (in-place update: nobody reads global memory after that).
// ------------------------------------------------------------------------------------------------
template <int BW, int S>
__global__ void __launch_bounds__(1024) k_lagrange_strided_split(double *__restrict__ f, const long long nlines, const int N,
                                                                  const long long inner, const DispDesc dd, const int use_tma) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *s = reinterpret_cast<double *>(smem_raw + 128);
    const int tid = threadIdx.x, lane = tid % BW, chunk = tid / BW, P = blockDim.x / BW;
    const long long l = (long long)blockIdx.x * BW + lane;
    const bool active = l < nlines;
    const long long o = active ? l / inner : 0, in = active ? l - o * inner : 0;
    double *base = f + o * (long long)N * inner + in;
    if (use_tma) {
        if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(N * BW * 8));
        const double *src0 = base - lane;
        for (int j = tid; j < N; j += blockDim.x) bulk_g2s(s + (size_t)j * BW, src0 + (long long)j * inner, BW * 8, bar);
        mbar_wait(bar, 0);
    } else {
        if (active)
            for (int j = chunk; j < N; j += P) cp_async8(s + (size_t)j * BW + lane, base + (long long)j * inner);
        cp_async_wait_all();
        __syncthreads();
    }
    if (!active) return;
    const int C = (N + P - 1) / P, i0 = chunk * C, i1 = (i0 + C < N) ? i0 + C : N;
    if (i0 >= i1) return;
    double pp[S], w[S];
    int idx = lagr_setup<S>(disp_of(dd, o, in), N, pp) + i0;
    idx %= N;
    const double *sc = s + lane;
#pragma unroll
    for (int k = 1; k < S; ++k) {
        w[k] = sc[idx * BW];
        idx = (idx == N - 1) ? 0 : idx + 1;
    }
    double *p = base + (long long)i0 * inner;
#pragma unroll 4
    for (int i = i0; i < i1; ++i) {
#pragma unroll
        for (int k = 0; k < S - 1; ++k) w[k] = w[k + 1];
        w[S - 1] = sc[idx * BW];
        idx = (idx == N - 1) ? 0 : idx + 1;
        double acc = pp[0] * w[0];
#pragma unroll
        for (int k = 1; k < S; ++k) acc = fma(pp[k], w[k], acc);
        st_stream(p, acc);
        p += inner;
    }
}

// ------------------------------------------------------------------------------------------------
// K1a (split): spline on a strided axis with every line cut into P chunks of C = N/P points, one warp
// per chunk (block = P warps = 32 lines).  The two first-order recurrences are restarted at every chunk
// boundary with the same 27-term geometric series the reference uses for the periodic wrap
// (sll_m_cubic_splines.F90:548-556,567-571; truncation (2-sqrt3)^27 = 3.6e-16), so the serial chain per
// tile is C instead of N steps and P times more warps are resident to hide latency.
// Chunk [k0,k1) computes g[k] for k = k1+2 .. k0 and emits cells k0+1 .. k1 (mod N).
// ------------------------------------------------------------------------------------------------
// NC > 0: line length known at compile time and a power of two (128, 64: the BASELINE sizes) -- wraps become masks and
// the chunk loops have constant trip counts; matters for the DIAG variant, whose extra arithmetic makes it issue-bound.
template <int P, bool REMAP, bool DIAG = false, int NC = 0>
__global__ void __launch_bounds__(32 * P) k_spline_strided_split(double *__restrict__ f, const int Nr,
                                                                  const long long inner, const DispDesc dd,
                                                                  const int use_tma, const long long nlines,
                                                                  double *__restrict__ linesum,
                                                                  const __grid_constant__ RemapDst rd,
                                                                  const LineDiag dg, const LineSub sub) {
    static_assert(NC == 0 || ((NC & (NC - 1)) == 0 && NC % P == 0), "power-of-two line length");
    const int N = NC > 0 ? NC : Nr;
    constexpr int BW = 32;
    constexpr int NPART = DIAG ? 4 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *part = reinterpret_cast<double *>(smem_raw + 128);          // [NPART][P][32] chunk sums (linesum only)
    double *s = reinterpret_cast<double *>(smem_raw + 128 + NPART * P * 32 * 8);
    const int tid = threadIdx.x, lane = tid & 31, chunk = tid >> 5;
    unsigned bid = blockIdx.x;
    if constexpr (REMAP) { bid += (unsigned)rd.block_rot; if (bid >= gridDim.x) bid -= gridDim.x; }
    long long l = (long long)bid * BW + lane;
    const bool active = l < nlines;
    long long o = active ? l / inner : 0, in = active ? l - o * inner : 0;
    if (sub.icount > 0 && active) {   // the l-th line of a subset (see LineSub); l becomes the line's number in the whole pass
        const long long q = l / sub.icount, r = l - q * sub.icount;
        o = (long long)sub.o_mul * q; in = sub.i0 + r + sub.in_pitch * q;
        l = o * inner + in;
    }
    double *base = f + o * (long long)N * inner + in;
    const int C = N / P, k0 = chunk * C, k1 = k0 + C;
    // DIAG: the kinetic-energy weights of the axis sit in the (until the end unused) chunk-sum area when they fit
    // (N <= 128 P): one shared-memory read per point -- a broadcast where neighbouring lines share the cell shift --
    // instead of a global load
    const bool w2_shared = DIAG && N <= NPART * P * 32;
    if constexpr (DIAG)
        if (w2_shared)
            for (int j = tid; j < N; j += 32 * P) part[j] = __ldg(dg.w2 + j);

    if (use_tma) {
        if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(N * BW * 8));
        const double *src0 = base - lane;
        for (int j = tid; j < N; j += 32 * P) bulk_g2s(s + (size_t)j * BW, src0 + (long long)j * inner, BW * 8, bar);
        mbar_wait(bar, 0);
    } else {
        if (active)
            for (int j = k0; j < k1; ++j) cp_async8(s + (size_t)j * BW + lane, base + (long long)j * inner);
        cp_async_wait_all();
        __syncthreads();
    }
    const double q = 0.26794919243112270647;
    double *sc = s + lane;
    // forward start value from pristine f: e[k0] = f[k0] + sum_i (-q)^{i+1} f[k0-1-i]
    double e = sc[k0 * BW];
    {
        int idx = NC > 0 ? ((k0 - 1) & (NC - 1)) : ((k0 == 0) ? N - 1 : k0 - 1);
#pragma unroll
        for (int i = 0; i < SLLB_NUM_TERMS; ++i) {
            e = fma(c_pw[i], sc[idx * BW], e);
            idx = NC > 0 ? ((idx - 1) & (NC - 1)) : ((idx == 0) ? N - 1 : idx - 1);
        }
    }
    __syncthreads(); // every chunk has read its warm-up inputs before anybody overwrites f with e
    sc[k0 * BW] = e;
#pragma unroll 8
    for (int k = k0 + 1; k < k1; ++k) {
        e = fma(-q, e, sc[k * BW]);
        sc[k * BW] = e;
    }
    __syncthreads();
    double total = 0.0, t1 = 0.0, t2 = 0.0, tk = 0.0;
    if (active) {
        // per-line weights (see spline_line)
        const double disp = disp_of(dd, o, in);
        const double r2 = 1.60769515458673623883;
        const double fl = floor(disp);
        const int dcell = (int)fl;
        const double dx = disp - fl, cdx = 1.0 - dx;
        const double s6 = r2 * (1.0 / 6.0);
        const double w0 = cdx * cdx * cdx * s6;
        const double w1 = (1.0 + 3.0 * cdx + 3.0 * cdx * cdx - 3.0 * cdx * cdx * cdx) * s6;
        const double w2 = (1.0 + 3.0 * dx + 3.0 * dx * dx - 3.0 * dx * dx * dx) * s6;
        const double w3 = dx * dx * dx * s6;
        // backward start: g[k1+2] = e[k1+2] + sum_i (-q)^{i+1} e[k1+3+i]   (indices mod N)
        int i2 = k1 + 2; if (i2 >= N) i2 -= N;
        int i1 = k1 + 1; if (i1 >= N) i1 -= N;
        int i0 = k1;     if (i0 >= N) i0 -= N;
        double g = sc[i2 * BW];
        {
            int idx = NC > 0 ? ((i2 + 1) & (NC - 1)) : ((i2 == N - 1) ? 0 : i2 + 1);
#pragma unroll
            for (int i = 0; i < SLLB_NUM_TERMS; ++i) {
                g = fma(c_pw[i], sc[idx * BW], g);
                idx = NC > 0 ? ((idx + 1) & (NC - 1)) : ((idx == N - 1) ? 0 : idx + 1);
            }
        }
        double a3 = g;                                  // g[k1+2]
        double a2 = fma(-q, a3, sc[i1 * BW]);           // g[k1+1]
        double a1 = fma(-q, a2, sc[i0 * BW]);           // g[k1]
        const int iout0 = NC > 0 ? ((k1 - dcell) & (NC - 1)) : ((k1 - dcell) % N + N) % N;   // output index of cell k1 (mod N)
        if constexpr (REMAP) {
            OutMap om = make_outmap(rd, o, in, N, inner);
            om.seek(iout0);
#pragma unroll 8
            for (int k = k1 - 1; k >= k0; --k) {
                const double a0 = fma(-q, a1, sc[k * BW]);
                const double val = fma(w3, a3, fma(w2, a2, fma(w1, a1, w0 * a0))); // cell k+1
                st_stream(om.p, val);
                total += val;
                om.prev();
                a3 = a2; a2 = a1; a1 = a0;
            }
        } else {
            double *p = base + (long long)iout0 * inner;
            double *const ptop = base + (long long)(N - 1) * inner;
            int iout = iout0;
            const double *wk = w2_shared ? part : dg.w2;
#pragma unroll 8
            for (int k = k1 - 1; k >= k0; --k) {
                const double a0 = fma(-q, a1, sc[k * BW]);
                const double val = fma(w3, a3, fma(w2, a2, fma(w1, a1, w0 * a0))); // cell k+1
                st_stream(p, val);
                total += val;
                if constexpr (DIAG) { t1 += fabs(val); t2 = fma(val, val, t2); tk = fma(wk[iout], val, tk); }
                p = (iout == 0) ? ptop : p - inner;
                iout = NC > 0 ? ((iout - 1) & (NC - 1)) : ((iout == 0) ? N - 1 : iout - 1);
                a3 = a2; a2 = a1; a1 = a0;
            }
        }
    }
    // optional: sum of every advected line (charge-density reduction fused into the pass, see K3b)
    if (linesum != nullptr) {
        if constexpr (DIAG) __syncthreads();   // the weights in `part` are dead only when every warp has left its loop
        part[chunk * 32 + lane] = total;
        if constexpr (DIAG) {
            part[(P + chunk) * 32 + lane] = t1;
            part[(2 * P + chunk) * 32 + lane] = t2;
            part[(3 * P + chunk) * 32 + lane] = tk;
        }
        __syncthreads();
        if (active)
            for (int m = chunk; m < NPART; m += P) { // warp `chunk` folds moments chunk, chunk + P, ... of the block's 32 lines
                const double *pm = part + (size_t)m * P * 32;
                double t = pm[lane];
#pragma unroll
                for (int c = 1; c < P; ++c) t += pm[c * 32 + lane];
                double *dst = linesum;
                if constexpr (DIAG) dst = (m == 0) ? linesum : (m == 1 ? dg.l1 : (m == 2 ? dg.l2 : dg.kin));
                dst[l] = t;
            }
    }
}

// ------------------------------------------------------------------------------------------------
// K1b: spline, contiguous axis (inner == 1).  Block = BW consecutive lines = one contiguous chunk of
// BW*N doubles.  The tile is transposed on the way in (pitch BW+1: conflict-free for both the
// coalesced staging and the thread-per-line sweeps) and on the way out, where the integer part of the
// shift is applied.
// ------------------------------------------------------------------------------------------------
template <int BW>
__global__ void __launch_bounds__(BW) k_spline_contig(double *__restrict__ f, const long long nlines, const int N,
                                                       const DispDesc dd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int *dcm = reinterpret_cast<int *>(smem_raw); // BW ints
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
        const double disp = disp_of(dd, l0 + tid, 0);
        const int dcell = (int)floor(disp);
        dcm[tid] = ((dcell % N) + N) % N;
        spline_line<P, false>(s + tid, N, disp, nullptr);
    }
    __syncthreads();
    for (int ln = 0; ln < nl; ++ln) {
        const int d = dcm[ln];
        for (int i = tid; i < N; i += BW) {
            int kc = i + d;
            if (kc >= N) kc -= N;
            st_stream(tile + (long long)ln * N + i, s[(size_t)kc * P + ln]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1b (split): spline, contiguous axis.  Block = 32 consecutive lines = one contiguous chunk of 32*N
// doubles, brought in by ONE bulk TMA copy in its natural layout s[line*N + k].  Thread (lane = line,
// warp = chunk) runs the two recurrences over C = N/P points of its line, restarted with the reference's
// 27-term series like K1a.  Because the line is periodic the chunk boundaries may sit anywhere: line l
// starts its chunks at k = l mod 16, so the 32 lanes of a warp touch 16 different bank pairs at every step
// (conflict-free for 8-byte words) although the pitch N is a multiple of 16.  The result of cell k+1 is
// parked in slot k (the slot just consumed), and the tile leaves through coalesced stores that apply the
// integer part of the shift.
// ------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(32 * P) k_spline_contig_split(double *__restrict__ f, const long long nlines, const int N,
                                                                 const DispDesc dd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    int *rot = reinterpret_cast<int *>(smem_raw + 16); // 32 ints: slot of output point 0 per line
    double *s = reinterpret_cast<double *>(smem_raw + 256);
    const int tid = threadIdx.x, lane = tid & 31, chunk = tid >> 5;
    const long long l0 = (long long)blockIdx.x * 32;
    const int nl = (int)((nlines - l0 < 32) ? (nlines - l0) : 32);
    double *tile = f + l0 * (long long)N;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, (uint32_t)((size_t)nl * N * 8));
        bulk_g2s(s, tile, (uint32_t)((size_t)nl * N * 8), bar);
    }
    const bool active = lane < nl;
    const int C = N / P;
    int k0 = chunk * C + (lane & 15);
    if (k0 >= N) k0 -= N;
    double *sc = s + (size_t)(active ? lane : 0) * N;
    const double q = 0.26794919243112270647;
    double disp = 0.0;
    if (active) disp = disp_of(dd, l0 + lane, 0);
    mbar_wait(bar, 0);
    // forward restart from pristine f: e[k0] = f[k0] + sum_i (-q)^{i+1} f[k0-1-i]
    double e = sc[k0];
    {
        int idx = (k0 == 0) ? N - 1 : k0 - 1;
#pragma unroll
        for (int i = 0; i < SLLB_NUM_TERMS; ++i) {
            e = fma(c_pw[i], sc[idx], e);
            idx = (idx == 0) ? N - 1 : idx - 1;
        }
    }
    __syncthreads();
    if (active) {
        sc[k0] = e;
        int k = k0;
#pragma unroll 8
        for (int t = 1; t < C; ++t) {
            k = (k == N - 1) ? 0 : k + 1;
            e = fma(-q, e, sc[k]);
            sc[k] = e;
        }
    }
    __syncthreads();
    // backward restart: with k1 = k0 + C, g[k1+2] = e[k1+2] + sum_i (-q)^{i+1} e[k1+3+i]   (indices mod N)
    int i0 = k0 + C; if (i0 >= N) i0 -= N;
    int i1 = (i0 == N - 1) ? 0 : i0 + 1;
    int i2 = (i1 == N - 1) ? 0 : i1 + 1;
    double g = sc[i2];
    {
        int idx = (i2 == N - 1) ? 0 : i2 + 1;
#pragma unroll
        for (int i = 0; i < SLLB_NUM_TERMS; ++i) {
            g = fma(c_pw[i], sc[idx], g);
            idx = (idx == N - 1) ? 0 : idx + 1;
        }
    }
    double a3 = g;                             // g[k1+2]
    double a2 = fma(-q, a3, sc[i1]);           // g[k1+1]
    double a1 = fma(-q, a2, sc[i0]);           // g[k1]
    __syncthreads(); // every chunk holds its start values before anybody overwrites e with results
    if (active) {
        const double r2 = 1.60769515458673623883;
        const double fl = floor(disp);
        const double dx = disp - fl, cdx = 1.0 - dx;
        const double s6 = r2 * (1.0 / 6.0);
        const double w0 = cdx * cdx * cdx * s6;
        const double w1 = (1.0 + 3.0 * cdx + 3.0 * cdx * cdx - 3.0 * cdx * cdx * cdx) * s6;
        const double w2 = (1.0 + 3.0 * dx + 3.0 * dx * dx - 3.0 * dx * dx * dx) * s6;
        const double w3 = dx * dx * dx * s6;
        if (chunk == 0) {
            // output point i takes cell (i + dcell) mod N, which is parked in slot (i + dcell - 1) mod N
            const int dcell = (int)fl;
            rot[lane] = (((dcell - 1) % N) + N) % N;
        }
        int k = i0;
#pragma unroll 8
        for (int t = 0; t < C; ++t) {
            k = (k == 0) ? N - 1 : k - 1;
            const double a0 = fma(-q, a1, sc[k]);
            sc[k] = fma(w3, a3, fma(w2, a2, fma(w1, a1, w0 * a0))); // cell k+1
            a3 = a2; a2 = a1; a1 = a0;
        }
    }
    __syncthreads();
    for (int ln = chunk; ln < nl; ln += P) {
        const int r = rot[ln];
        const double *row = s + (size_t)ln * N;
        double *orow = tile + (long long)ln * N;
        for (int i = lane; i < N; i += 32) {
            int kc = i + r;
            if (kc >= N) kc -= N;
            st_stream(orow + i, row[kc]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1c: the whole T stage in one sweep.  The x1 pass and the x2 pass of a splitting stage act on the same
// contiguous (x1,x2) plane of f and their displacements are constant over that plane (alpha = v3*step,
// alpha = v4*step: sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:1037-1064), so a persistent CTA stages one
// N1 x N2 plane (one bulk TMA copy), runs both periodic spline solves + shift-evaluates in shared memory
// and writes the plane once: two advection passes for 16 B of HBM traffic per point.  Optionally the CTA
// also accumulates its planes into a per-CTA partial charge density (the reduction over x3,x4 that follows
// the T stage, sll_m_reduction.F90:187-272), so rho needs no extra sweep over f.
// Threads = N1*N2/16: every line is cut into chunks of 16 points (27-term restarts as in K1a).  Pass A runs
// along x1 (rows, skewed chunk starts for conflict-free banks), pass B along x2 (columns).  Each pass parks
// cell k+1 in slot k, so after both passes out(i1,i2) sits at slot (i1+r1, i2+r2) with r = dcell-1 (mod N).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void plane_pass(double *sc, const int stride, const int N, const int k0, const int C,
                                           const double disp) {
    const double q = 0.26794919243112270647;
    // restart values: the 27-term series of the reference in Horner form, i.e. the recurrence itself started
    // 27 points upstream from zero (no coefficient table: it would be hoisted into 54 registers here)
    double e;
    {
        int idx = k0 - SLLB_NUM_TERMS;
        while (idx < 0) idx += N;
        e = sc[idx * stride];
#pragma unroll 9
        for (int i = 0; i < SLLB_NUM_TERMS; ++i) {
            idx = (idx == N - 1) ? 0 : idx + 1;
            e = fma(-q, e, sc[idx * stride]);
        }
    }
    __syncthreads();
    sc[k0 * stride] = e;
    int k = k0;
#pragma unroll 4
    for (int t = 1; t < C; ++t) {
        k = (k == N - 1) ? 0 : k + 1;
        e = fma(-q, e, sc[k * stride]);
        sc[k * stride] = e;
    }
    __syncthreads();
    int i0 = k0 + C; if (i0 >= N) i0 -= N;
    const int i1 = (i0 == N - 1) ? 0 : i0 + 1;
    const int i2 = (i1 == N - 1) ? 0 : i1 + 1;
    double g;
    {
        int idx = i2 + SLLB_NUM_TERMS;
        while (idx >= N) idx -= N;
        g = sc[idx * stride];
#pragma unroll 9
        for (int i = 0; i < SLLB_NUM_TERMS; ++i) {
            idx = (idx == 0) ? N - 1 : idx - 1;
            g = fma(-q, g, sc[idx * stride]);
        }
    }
    double a3 = g;
    double a2 = fma(-q, a3, sc[i1 * stride]);
    double a1 = fma(-q, a2, sc[i0 * stride]);
    __syncthreads();
    const double r2 = 1.60769515458673623883;
    const double fl = floor(disp);
    const double dx = disp - fl, cdx = 1.0 - dx;
    const double s6 = r2 * (1.0 / 6.0);
    const double w0 = cdx * cdx * cdx * s6;
    const double w1 = (1.0 + 3.0 * cdx + 3.0 * cdx * cdx - 3.0 * cdx * cdx * cdx) * s6;
    const double w2 = (1.0 + 3.0 * dx + 3.0 * dx * dx - 3.0 * dx * dx * dx) * s6;
    const double w3 = dx * dx * dx * s6;
    k = i0;
#pragma unroll 4
    for (int t = 0; t < C; ++t) {
        k = (k == 0) ? N - 1 : k - 1;
        const double a0 = fma(-q, a1, sc[k * stride]);
        sc[k * stride] = fma(w3, a3, fma(w2, a2, fma(w1, a1, w0 * a0))); // cell k+1 parked in slot k
        a3 = a2; a2 = a1; a1 = a0;
    }
    __syncthreads();
}

// EPT = plane points per thread = chunk length; RHO: keep the per-thread partial sums (EPT registers)
template <int EPT, bool RHO>
__global__ void __launch_bounds__(16384 / EPT, 1) k_spline_plane(double *__restrict__ f, const int N1, const int N2,
                                                                 const long long nplanes, const DispDesc dd1,
                                                                 const DispDesc dd2, double *__restrict__ rho_partial) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *s = reinterpret_cast<double *>(smem_raw + 128);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, T = blockDim.x;
    const int npl = N1 * N2;
    constexpr int C = EPT;
    const int PA = N1 / C, PB = N2 / C;              // chunks per line in pass A / B
    // pass A: line = x2 index (row of N1 contiguous points); pass B: line = x1 index (column, stride N1)
    const int lineA = (w / PA) * 32 + lane, chA = w % PA;
    int k0A = chA * C + (lane & 15); if (k0A >= N1) k0A -= N1;
    const int lineB = (w / PB) * 32 + lane, chB = w % PB;
    const int k0B = chB * C;
    double acc[RHO ? EPT : 1];
#pragma unroll
    for (int j = 0; j < (RHO ? EPT : 1); ++j) acc[j] = 0.0;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t phase = 0;
    for (long long pl = blockIdx.x; pl < nplanes; pl += gridDim.x) {
        double *gp = f + pl * (long long)npl;
        if (tid == 0) {
            mbar_arrive_expect_tx(bar, (uint32_t)(npl * 8));
            bulk_g2s(s, gp, (uint32_t)(npl * 8), bar);
        }
        const double d1 = disp_of(dd1, pl * N2, 0); // constant over the plane (checked by the launcher)
        const double d2 = disp_of(dd2, pl, 0);
        mbar_wait(bar, phase);
        phase ^= 1;
        plane_pass(s + (size_t)lineA * N1, 1, N1, k0A, C, d1);
        plane_pass(s + lineB, N1, N2, k0B, C, d2);
        const int r1 = ((((int)floor(d1) - 1) % N1) + N1) % N1;
        const int r2 = ((((int)floor(d2) - 1) % N2) + N2) % N2;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int idx = tid + j * T;
            const int i2 = idx / N1, i1 = idx - i2 * N1;
            int c1 = i1 + r1; if (c1 >= N1) c1 -= N1;
            int c2 = i2 + r2; if (c2 >= N2) c2 -= N2;
            const double v = s[(size_t)c2 * N1 + c1];
            st_stream(gp + idx, v);
            if constexpr (RHO) acc[j] += v;
        }
        __syncthreads(); // the plane has left shared memory before the next bulk copy lands
    }
    if constexpr (RHO) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) rho_partial[(long long)blockIdx.x * npl + tid + j * T] = acc[j];
    }
}

// ------------------------------------------------------------------------------------------------
// K1c, register-resident variant.  Same job as k_spline_plane, organised around the shared-memory pipe,
// which is what bounds the in-place variant (5.7 shared accesses per point per pass):
//   * a thread keeps its chunk of 32 points in registers through both sweeps of a pass: one shared read and
//     one shared write per point per pass;
//   * the sweeps start from zero inside every chunk; the exact start values arrive afterwards from the
//     neighbouring chunk as one scalar each (end value of the previous chunk for the forward sweep, first
//     value of the next one for the backward sweep) and enter through the same 27-term geometric series the
//     reference truncates at (sll_m_cubic_splines.F90:548-571): e[j] += (-q)^(j+1) e_prev_end, j < 27;
//   * pass A scatters its results to their final x1 positions, pass B shifts its chunk grid by the integer
//     part of the x2 displacement, so every thread owns fixed output points (x1, 32*chunk + j): the results go
//     from registers straight to global memory (coalesced rows) and into fixed charge-density accumulators
//     (12 in registers, 20 in shared memory);
//   * the plane is dead in shared memory as soon as pass B has loaded it, so the bulk TMA copy of the next
//     plane is issued there and overlaps the arithmetic and the stores of pass B.
// ------------------------------------------------------------------------------------------------
#define SLLB_PR_C 32     // chunk length = points per thread
#define SLLB_PR_ACCR 12  // charge-density accumulators kept in registers (the other 20 live in shared memory: 227 KB budget)

// two sweeps + evaluation on a register-resident chunk g[0..31]; exch: [0..nchunks*nlines) end values,
// then 3 arrays of first values.  `line` / `nlines` / `ch` / `nch` identify the chunk for the exchange.
__device__ __forceinline__ void chunk_solve(double (&g)[SLLB_PR_C], double *exch, const int line, const int nlines,
                                            const int ch, const int nch, const double dx) {
    constexpr int C = SLLB_PR_C;
    const double q = 0.26794919243112270647;
    double *e_end = exch, *g0 = exch + (size_t)nlines * nch, *g1 = g0 + (size_t)nlines * nch, *g2 = g1 + (size_t)nlines * nch;
#pragma unroll
    for (int j = 1; j < C; ++j) g[j] = fma(-q, g[j - 1], g[j]);
    e_end[ch * nlines + line] = g[C - 1];
    __syncthreads();
    {
        const int cp = (ch == 0) ? nch - 1 : ch - 1;
        const double ep = e_end[cp * nlines + line];
#pragma unroll
        for (int j = 0; j < SLLB_NUM_TERMS; ++j) g[j] = fma(c_pw[j], ep, g[j]);
    }
#pragma unroll
    for (int j = C - 2; j >= 0; --j) g[j] = fma(-q, g[j + 1], g[j]);
    g0[ch * nlines + line] = g[0];
    g1[ch * nlines + line] = g[1];
    g2[ch * nlines + line] = g[2];
    __syncthreads();
    const int cn = (ch == nch - 1) ? 0 : ch + 1;
    const double n0 = g0[cn * nlines + line], n1 = g1[cn * nlines + line], n2 = g2[cn * nlines + line];
#pragma unroll
    for (int j = 0; j < SLLB_NUM_TERMS; ++j) g[C - 1 - j] = fma(c_pw[j], n0, g[C - 1 - j]);
    const double r2 = 1.60769515458673623883;
    const double cdx = 1.0 - dx, s6 = r2 * (1.0 / 6.0);
    const double w0 = cdx * cdx * cdx * s6;
    const double w1 = (1.0 + 3.0 * cdx + 3.0 * cdx * cdx - 3.0 * cdx * cdx * cdx) * s6;
    const double w2 = (1.0 + 3.0 * dx + 3.0 * dx * dx - 3.0 * dx * dx * dx) * s6;
    const double w3 = dx * dx * dx * s6;
    // g[j] <- value of cell (k0 + j + 1): w0 g[j] + w1 g[j+1] + w2 g[j+2] + w3 g[j+3]
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const double b1 = (j + 1 < C) ? g[j + 1] : n0;
        const double b2 = (j + 2 < C) ? g[j + 2] : ((j + 2 == C) ? n0 : n1);
        const double b3 = (j + 3 < C) ? g[j + 3] : ((j + 3 == C) ? n0 : ((j + 3 == C + 1) ? n1 : n2));
        g[j] = fma(w3, b3, fma(w2, b2, fma(w1, b1, w0 * g[j])));
    }
}

// TACC (with RHO): the 32 charge-density accumulators of every thread live in tensor memory (warp w: lanes of quadrant
// w mod 4, columns 64 (w / 4) .. + 63) instead of 12 registers + 20 shared-memory slots: no register spills, 80 KB of shared
// memory and 40 shared-memory accesses per thread and plane less; the TMEM loads of a group of 8 are in flight while the 8
// values are evaluated.
// NC > 0: square planes of NC x NC points with NC a power of two known at compile time (128^4, 64^4: the BASELINE sizes) --
// the periodic wraps become masks, the chunk and row arithmetic constants, the strides immediates.
template <int NC>
__device__ __forceinline__ int wrap_next(const int k, const int n) {
    if constexpr (NC > 0) return (k + 1) & (NC - 1);
    else return (k == n - 1) ? 0 : k + 1;
}
template <bool RHO, bool REMAP, bool TACC = false, int NC = 0>
__global__ void __launch_bounds__(512, 1) k_spline_plane_r(double *__restrict__ f, const int N1r, const int N2r,
                                                           const long long nplanes, const DispDesc dd1,
                                                           const DispDesc dd2, double *__restrict__ rho_partial,
                                                           const __grid_constant__ RemapDst rd, const int l2_prefetch) {
    constexpr int C = SLLB_PR_C;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *s = reinterpret_cast<double *>(smem_raw + 128);
    static_assert(NC == 0 || ((NC & (NC - 1)) == 0 && NC >= 32 && NC * NC / SLLB_PR_C <= 512), "square power-of-two planes");
    const int N1 = NC > 0 ? NC : N1r, N2 = NC > 0 ? NC : N2r;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, T = NC > 0 ? NC * NC / SLLB_PR_C : (int)blockDim.x;
    const int npl = N1 * N2;
    double *exch = s + npl;                    // 4 * T doubles
    double *accs = exch + 4 * (size_t)T;       // RHO: (C - ACCR) * T doubles
    const int PA = N1 / C, PB = N2 / C;
    const int rowA = (w / PA) * 32 + lane, chA = w % PA; // pass A: row = x2 index, chunk along x1
    const int colB = (w / PB) * 32 + lane, chB = w % PB; // pass B: column = x1 index, chunk along x2
    double acc[(RHO && TACC) ? 1 : SLLB_PR_ACCR];
#pragma unroll
    for (int j = 0; j < ((RHO && TACC) ? 1 : SLLB_PR_ACCR); ++j) acc[j] = 0.0;
    if (RHO && !TACC)
        for (int j = 0; j < C - SLLB_PR_ACCR; ++j) accs[(size_t)j * T + tid] = 0.0;
    uint32_t tbase = 0, tcols = 0, talloc = 0;
    if constexpr (RHO && TACC) {
        uint32_t *tslot = reinterpret_cast<uint32_t *>(smem_raw + 64);
        tcols = 64u * (uint32_t)((T / 32 + 3) / 4);
        tcols = tcols <= 64 ? 64 : (tcols <= 128 ? 128 : 256);
        if (w == 0) tmem_alloc(tslot, tcols);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        talloc = *tslot;
        tbase = talloc + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)((w >> 2) * 64);
        uint32_t z[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) z[k] = 0u;
#pragma unroll
        for (int jb = 0; jb < C; jb += 8) tmem_st16(tbase + 2 * jb, z);
        tmem_wait_st();
    }
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t phase = 0;
    long long pl = blockIdx.x;
    if (tid == 0 && pl < nplanes) {
        mbar_arrive_expect_tx(bar, (uint32_t)(npl * 8));
        bulk_g2s(s, f + pl * (long long)npl, (uint32_t)(npl * 8), bar);
    }
    for (; pl < nplanes; pl += gridDim.x) {
        double *gp = f + pl * (long long)npl;
        // The shared plane is busy until pass B has taken it into registers, so the bulk copy of the next plane can
        // only be issued halfway through this iteration: during pass A nothing of this CTA is in flight from HBM.
        // Asking for the next plane in L2 now keeps the DRAM reads going through pass A; the bulk copy then hits L2.
        if (l2_prefetch && tid == 0 && pl + gridDim.x < nplanes) bulk_prefetch_l2(f + (pl + gridDim.x) * (long long)npl, (uint32_t)(npl * 8));
        const double d1 = disp_of(dd1, pl * N2, 0); // constant over the plane (checked by the launcher)
        const double d2 = disp_of(dd2, pl, 0);
        const double fl1 = floor(d1), fl2 = floor(d2);
        double g[C];
        // ---- pass A: rows ----
        int k0 = chA * C + (lane & 15);        // skewed chunk start: conflict-free banks with pitch N1
        if (k0 >= N1) k0 -= N1;
        mbar_wait(bar, phase);
        phase ^= 1;
        {
            const double *row = s + (size_t)rowA * N1;
            int k = k0;
#pragma unroll
            for (int j = 0; j < C; ++j) { g[j] = row[k]; k = wrap_next<NC>(k, N1); }
        }
        chunk_solve(g, exch, rowA, N2, chA, PA, d1 - fl1);
        {
            // cell k0+j+1 is output point (k0 + j + 1 - dcell1) mod N1
            double *row = s + (size_t)rowA * N1;
            int i1 = (int)(((long long)k0 + 1 - (long long)fl1) % N1);
            if (i1 < 0) i1 += N1;
#pragma unroll
            for (int j = 0; j < C; ++j) { row[i1] = g[j]; i1 = wrap_next<NC>(i1, N1); }
        }
        __syncthreads();
        // ---- pass B: columns; the chunk grid is shifted so that this thread's cells are points 32*chB + j ----
        {
            int k = (int)(((long long)chB * C - 1 + (long long)fl2) % N2);
            if (k < 0) k += N2;
            const double *col = s + colB;
#pragma unroll
            for (int j = 0; j < C; ++j) { g[j] = col[(size_t)k * N1]; k = wrap_next<NC>(k, N2); }
        }
        // (the first barrier inside chunk_solve also says: every thread has read the plane)
        {
            constexpr double q = 0.26794919243112270647;
#pragma unroll
            for (int j = 1; j < C; ++j) g[j] = fma(-q, g[j - 1], g[j]);
            exch[chB * N1 + colB] = g[C - 1];
            __syncthreads();
            const long long nxt = pl + gridDim.x;
            if (tid == 0 && nxt < nplanes) { // the plane is dead in shared memory: fetch the next one
                mbar_arrive_expect_tx(bar, (uint32_t)(npl * 8));
                bulk_g2s(s, f + nxt * (long long)npl, (uint32_t)(npl * 8), bar);
            }
            double *e_end = exch, *g0 = exch + (size_t)N1 * PB, *g1 = g0 + (size_t)N1 * PB, *g2 = g1 + (size_t)N1 * PB;
            {
                const int cp = (chB == 0) ? PB - 1 : chB - 1;
                const double ep = e_end[cp * N1 + colB];
#pragma unroll
                for (int j = 0; j < SLLB_NUM_TERMS; ++j) g[j] = fma(c_pw[j], ep, g[j]);
            }
#pragma unroll
            for (int j = C - 2; j >= 0; --j) g[j] = fma(-q, g[j + 1], g[j]);
            g0[chB * N1 + colB] = g[0];
            g1[chB * N1 + colB] = g[1];
            g2[chB * N1 + colB] = g[2];
            __syncthreads();
            const int cn = (chB == PB - 1) ? 0 : chB + 1;
            const double n0 = g0[cn * N1 + colB], n1 = g1[cn * N1 + colB], n2 = g2[cn * N1 + colB];
#pragma unroll
            for (int j = 0; j < SLLB_NUM_TERMS; ++j) g[C - 1 - j] = fma(c_pw[j], n0, g[C - 1 - j]);
            const double dx = d2 - fl2, cdx = 1.0 - dx, s6 = 1.60769515458673623883 * (1.0 / 6.0);
            const double w0 = cdx * cdx * cdx * s6;
            const double w1 = (1.0 + 3.0 * cdx + 3.0 * cdx * cdx - 3.0 * cdx * cdx * cdx) * s6;
            const double w2 = (1.0 + 3.0 * dx + 3.0 * dx * dx - 3.0 * dx * dx * dx) * s6;
            const double w3 = dx * dx * dx * s6;
            double *out = gp + (size_t)(chB * C) * N1 + colB;
            OutMap om;
            if constexpr (REMAP) {
                // this thread's points are the x2 line (o = plane, in = x1) of the source layout: same map as K1a
                om = make_outmap(rd, pl, colB, N2, N1);
                om.seek(chB * C);
            }
#pragma unroll
            for (int jb = 0; jb < C; jb += 4) {
                uint32_t r[8];
                if constexpr (RHO && TACC) tmem_ld8(tbase + 2 * jb, r);   // in flight while the 4 values are evaluated
                double vv[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = jb + jj;
                    const double b1 = (j + 1 < C) ? g[j + 1] : n0;
                    const double b2 = (j + 2 < C) ? g[j + 2] : ((j + 2 == C) ? n0 : n1);
                    const double b3 = (j + 3 < C) ? g[j + 3] : ((j + 3 == C) ? n0 : ((j + 3 == C + 1) ? n1 : n2));
                    const double v = fma(w3, b3, fma(w2, b2, fma(w1, b1, w0 * g[j])));
                    vv[jj] = v;
                    if constexpr (REMAP) { st_stream(om.p, v); om.next(); }
                    else st_stream(out + (size_t)j * N1, v);
                    if constexpr (RHO && !TACC) {
                        if (j < SLLB_PR_ACCR) acc[j] += v;
                        else accs[(size_t)(j - SLLB_PR_ACCR) * T + tid] += v;
                    }
                }
                if constexpr (RHO && TACC) {
                    tmem_wait_ld();
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const double a = __hiloint2double((int)r[2 * jj + 1], (int)r[2 * jj]) + vv[jj];
                        r[2 * jj] = (uint32_t)__double2loint(a); r[2 * jj + 1] = (uint32_t)__double2hiint(a);
                    }
                    tmem_st8(tbase + 2 * jb, r);
                }
            }
            if constexpr (RHO && TACC) tmem_wait_st();   // the next plane loads these columns again
        }
        __syncthreads(); // exchange arrays are reused by the next plane
    }
    if constexpr (RHO && TACC) {
        double *rp = rho_partial + (long long)blockIdx.x * npl + (size_t)(chB * C) * N1 + colB;
#pragma unroll
        for (int jb = 0; jb < C; jb += 8) {
            uint32_t r[16];
            tmem_ld16(tbase + 2 * jb, r);
            tmem_wait_ld();
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) rp[(size_t)(jb + jj) * N1] = __hiloint2double((int)r[2 * jj + 1], (int)r[2 * jj]);
        }
        tmem_fence_before_sync();
        __syncthreads();
        if (w == 0) tmem_dealloc(talloc, tcols);
    } else if constexpr (RHO) {
        double *rp = rho_partial + (long long)blockIdx.x * npl + (size_t)(chB * C) * N1 + colB;
#pragma unroll
        for (int j = 0; j < C; ++j) rp[(size_t)j * N1] = (j < SLLB_PR_ACCR) ? acc[j] : accs[(size_t)(j - SLLB_PR_ACCR) * T + tid];
    }
}

// ------------------------------------------------------------------------------------------------
// K2b: Lagrange, contiguous axis.  Tile = LT whole lines in natural layout (one bulk TMA copy),
// thread per output point, per-line weights in shared memory.
// ------------------------------------------------------------------------------------------------
// POW2: N is a power of two (the usual case): line / point indices and the periodic wrap are shifts and masks instead
// of an integer division and seven compare-selects per point (the kernel is issue-bound on 32-point lines otherwise).
template <int S, bool POW2>
__global__ void __launch_bounds__(256) k_lagrange_contig(double *__restrict__ f, const long long nlines, const int N,
                                                         const DispDesc dd, const int LT, const int use_tma) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *tile = reinterpret_cast<double *>(smem_raw + 128);
    double *wts = tile + (size_t)LT * N;
    int *offs = reinterpret_cast<int *>(wts + (size_t)LT * S);
    const int tid = threadIdx.x;
    const long long l0 = (long long)blockIdx.x * LT;
    const int nl = (int)((nlines - l0 < LT) ? (nlines - l0) : LT);
    double *g = f + l0 * (long long)N;
    const int npts = nl * N;
    if (use_tma) {
        if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(bar, (uint32_t)(npts * 8));
            bulk_g2s(tile, g, (uint32_t)(npts * 8), bar);
        }
    } else {
        for (int i = tid; i < npts; i += 256) cp_async8(tile + i, g + i);
    }
    for (int ln = tid; ln < nl; ln += 256) {
        double pp[S];
        offs[ln] = lagr_setup<S>(disp_of(dd, l0 + ln, 0), N, pp);
#pragma unroll
        for (int k = 0; k < S; ++k) wts[ln * S + k] = pp[k];
    }
    if (use_tma) mbar_wait(bar, 0);
    else cp_async_wait_all();
    __syncthreads();
    if constexpr (POW2) {
        const int mask = N - 1, sh = 31 - __clz(N);
        if constexpr (S <= 11) {
            // two consecutive points per thread: their S-point windows overlap in S - 1 points, so (S + 3) / 2 aligned
            // 16-byte shared loads + S broadcast weight loads serve 2 points instead of 2 S + 2 S 8-byte loads -- the
            // kernel was bound by shared-memory wavefronts on 32-point lines (S = 7: 0.69 per point, now 0.44).  Same
            // left-to-right sums, bit-identical results.
            if ((reinterpret_cast<uintptr_t>(f) & 15) == 0) {
                constexpr int NL = (S + 3) / 2;
                for (int idx = 2 * tid; idx < npts; idx += 512) {
                    const int ln = idx >> sh, i = idx & mask;
                    const int j0 = i + offs[ln];
                    const int a = j0 & ~1, odd = j0 & 1;
                    const double *row = tile + ((size_t)ln << sh);
                    double v[2 * NL];
#pragma unroll
                    for (int m = 0; m < NL; ++m) {
                        const double2 t = *reinterpret_cast<const double2 *>(row + ((a + 2 * m) & mask));
                        v[2 * m] = t.x; v[2 * m + 1] = t.y;
                    }
                    double u[S + 1];
#pragma unroll
                    for (int m = 0; m <= S; ++m) u[m] = odd ? v[m + 1] : v[m];
                    const double *w = wts + ln * S;
                    const double w0 = w[0];
                    double acc0 = w0 * u[0], acc1 = w0 * u[1];
#pragma unroll
                    for (int k = 1; k < S; ++k) {
                        const double wk = w[k];
                        acc0 = fma(wk, u[k], acc0);
                        acc1 = fma(wk, u[k + 1], acc1);
                    }
                    st_stream2(g + idx, acc0, acc1);
                }
                return;
            }
        }
        for (int idx = tid; idx < npts; idx += 256) {
            const int ln = idx >> sh, i = idx & mask;
            const int j0 = i + offs[ln];
            const double *row = tile + ((size_t)ln << sh);
            const double *w = wts + ln * S;
            double acc = w[0] * row[j0 & mask];
#pragma unroll
            for (int k = 1; k < S; ++k) acc = fma(w[k], row[(j0 + k) & mask], acc);
            st_stream(g + idx, acc);
        }
    } else {
        for (int idx = tid; idx < npts; idx += 256) {
            const int ln = idx / N, i = idx - ln * N;
            int j = i + offs[ln];
            if (j >= N) j -= N;
            const double *row = tile + (size_t)ln * N;
            const double *w = wts + ln * S;
            double acc = w[0] * row[j];
#pragma unroll
            for (int k = 1; k < S; ++k) {
                j = (j == N - 1) ? 0 : j + 1;
                acc = fma(w[k], row[j], acc);
            }
            st_stream(g + idx, acc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2c: fixed odd Lagrange on a strided axis whose lines are split across ranks: the local piece of N
// points is extended by H = (S-1)/2 halo planes on each side (left | local | right), exactly the
// buf_i the reference assembles before sll_s_lagrange_interpolation_1d_fast_disp_fixed_haloc_cells
// (sll_m_advection_6d_lagrange_dd_slim.F90:1636-1660).  Halo buffers are [outer][H][inner].
// ------------------------------------------------------------------------------------------------
template <int BW, int S>
__global__ void __launch_bounds__(BW) k_lagrange_halo(double *__restrict__ f, const double *__restrict__ hl,
                                                       const double *__restrict__ hr, const long long nlines,
                                                       const int N, const long long inner, const DispDesc dd,
                                                       const int use_tma, const LineBox lbx) {
    constexpr int H = (S - 1) / 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *s = reinterpret_cast<double *>(smem_raw + 128);
    const int tid = threadIdx.x;
    // the lines of the sub-box lbx (all of them when it spans the block): o in [o0, o0+ocount), in in [i0, i0+icount)
    const long long l = (long long)blockIdx.x * BW + tid;
    const bool active = l < nlines;
    const long long osub = active ? l / lbx.icount : 0;
    const long long o = lbx.o0 + osub, in = lbx.i0 + (active ? l - osub * lbx.icount : 0);
    double *base = f + o * (long long)N * inner + in;
    const double *lb = hl + o * (long long)H * inner + in;
    const double *rb = hr + o * (long long)H * inner + in;
    const int R = N + 2 * H;
    if (use_tma) {
        if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(R * BW * 8));
        for (int j = tid; j < R; j += BW) {
            const double *src = (j < H) ? lb - tid + (long long)j * inner
                              : (j < H + N) ? base - tid + (long long)(j - H) * inner
                                            : rb - tid + (long long)(j - H - N) * inner;
            bulk_g2s(s + (size_t)j * BW, src, BW * 8, bar);
        }
        mbar_wait(bar, 0);
    } else {
        if (active) {
            for (int j = 0; j < H; ++j) cp_async8(s + (size_t)j * BW + tid, lb + (long long)j * inner);
            for (int j = 0; j < N; ++j) cp_async8(s + (size_t)(j + H) * BW + tid, base + (long long)j * inner);
            for (int j = 0; j < H; ++j) cp_async8(s + (size_t)(j + H + N) * BW + tid, rb + (long long)j * inner);
        }
        cp_async_wait_all();
    }
    if (!active) return;
    double pp[S], w[S];
    lagr_coeff<S>(disp_of(dd, o, in), pp);
    const double *sc = s + tid;
#pragma unroll
    for (int k = 1; k < S; ++k) w[k] = sc[(k - 1) * BW];
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int k = 0; k < S - 1; ++k) w[k] = w[k + 1];
        w[S - 1] = sc[(i + S - 1) * BW];
        double acc = pp[0] * w[0];
#pragma unroll
        for (int k = 1; k < S; ++k) acc = fma(pp[k], w[k], acc);
        st_stream(base + (long long)i * inner, acc);
    }
}

// K7: halo pack, buf[o][j][in] = f[o][j0 + j][in] for j < hw; f viewed as [outer][n][inner]
// VEC: 16-byte loads/stores (rows 16-byte aligned and of even length), four in flight per thread: the stores may cross
// NVLink (peer halo buffers), where latency, not issue rate, limits a thread
template <bool VEC>
__global__ void __launch_bounds__(256) k_halo_pack(const double *__restrict__ f, const int n, const long long inner,
                                                   const int j0, const int hw, double *__restrict__ buf, const LineBox lbx) {
    const long long stride = (long long)gridDim.x * 256, t0 = (long long)blockIdx.x * 256 + threadIdx.x;
    for (long long o = lbx.o0 + blockIdx.y; o < lbx.o0 + lbx.ocount; o += gridDim.y)
        for (int j = 0; j < hw; ++j) {
            const double *src = f + (o * n + j0 + j) * inner + lbx.i0;
            double *dst = buf + (o * hw + j) * inner + lbx.i0;
            if (VEC) {
                const double2 *s2 = reinterpret_cast<const double2 *>(src);
                double2 *d2 = reinterpret_cast<double2 *>(dst);
                const long long n2 = lbx.icount / 2;
                long long t = t0;
                for (; t + 3 * stride < n2; t += 4 * stride) {
                    const double2 a = __ldcs(s2 + t), b = __ldcs(s2 + t + stride), c = __ldcs(s2 + t + 2 * stride),
                                  d = __ldcs(s2 + t + 3 * stride);
                    d2[t] = a; d2[t + stride] = b; d2[t + 2 * stride] = c; d2[t + 3 * stride] = d;
                }
                for (; t < n2; t += stride) d2[t] = __ldcs(s2 + t);
            } else {
                for (long long t = t0; t < lbx.icount; t += stride) dst[t] = __ldcs(src + t);
            }
        }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static bool g_pw_ready = false;
static cudaError_t ensure_constants() {
    if (g_pw_ready) return cudaSuccess;
    double pw[SLLB_NUM_TERMS];
    // b/a evaluated like the reference (sqrt of the two constants), then powers by repeated product
    const double a = sqrt((2.0 + sqrt(3.0)) / 6.0), b = sqrt((2.0 - sqrt(3.0)) / 6.0);
    double ct = 1.0;
    for (int i = 0; i < SLLB_NUM_TERMS; ++i) { ct *= -(b / a); pw[i] = ct; }
    cudaError_t e = cudaMemcpyToSymbol(c_pw, pw, sizeof(pw));
    if (e == cudaSuccess) g_pw_ready = true;
    return e;
}

template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

static const size_t SMEM_MAX = 227 * 1024;

template <int BW, int METHOD, int S>
static cudaError_t launch_strided_t(double *f, long long nlines, int N, long long inner, const DispDesc &dd,
                                    int staging, cudaStream_t st, const RemapDst &rd) {
    size_t smem = 128 + (size_t)N * BW * 8;
    auto kern = k_advect_strided<BW, METHOD, S>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    bool tma_ok = (inner % BW == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0) && ((size_t)N * BW * 8 < (1u << 20));
    int use_tma = (staging == STAGING_CPASYNC) ? 0 : (tma_ok ? 1 : 0);
    long long nblk = (nlines + BW - 1) / BW;
    kern<<<(unsigned)nblk, BW, smem, st>>>(f, nlines, N, inner, dd, use_tma, rd);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

int g_spline_split = [] { const char *e = getenv("SLLB_SPLINE_SPLIT"); return e ? atoi(e) : -1; }(); // -1 auto, else forced P in {1,2,4,8}
int g_remap_rotation = 1; // fused remap: 1 = ranks start their sweep at different destination ranks
// A fused pass walks its tiles with the slowest non-advected axis outermost.  When the destination layout splits
// that axis, every rank would store to the same group of receivers at the same time (all senders into
// 1/tp of the receivers: their NVLink ingress is the bottleneck, the other receivers idle).  Rotating the start
// tile by rank/tp of the sweep spreads the senders evenly over the receivers at every moment.
static int remap_block_rotation(const RemapDst &rd, long long nblk) {
    if (!rd.on || !g_remap_rotation) return 0;
    int slow = 3;
    if (slow == rd.axis) slow = 2;
    const int tp = rd.tp[slow];
    if (tp <= 1 || nblk % tp != 0) return 0;
    return (int)((rd.rank % tp) * (nblk / tp));
}
// strided spline pass, instantiations with a compile-time line length (N = 32 P: 128 with P = 4, 64 with P = 2; 28 % fewer
// instructions): 0 none, 1 the variant with the fused diagnostics, 2 (default) the plain variant too (SLLB_SPLIT_CONST_LEN).
// 128^4 step 3.54 / 3.48 / 3.43 ms, 64^4 step 0.313 / 0.307 / 0.301 ms (profiles/r02_split_const_ab.log).
int g_split_const_len = [] { const char *e = getenv("SLLB_SPLIT_CONST_LEN"); return e ? atoi(e) : 2; }();
template <int P>
static cudaError_t launch_spline_split_t(double *f, long long nlines, int N, long long inner, const DispDesc &dd,
                                         int staging, cudaStream_t st, const RemapDst &rd, double *linesum,
                                         const LineDiag *diag, const LineSub *subp) {
    LineSub sub = {0, 0, 0, 0, 0};
    if (subp && subp->icount > 0) {
        if (subp->icount % 32 != 0 || subp->i0 % 32 != 0 || subp->nlines % 32 != 0) return cudaErrorInvalidValue;
        sub = *subp;
        nlines = sub.nlines;
    }
    const bool with_diag = diag != nullptr && linesum != nullptr && !rd.on;
    if (diag != nullptr && !with_diag) return cudaErrorNotSupported;
    LineDiag dg = {nullptr, nullptr, nullptr, nullptr};
    if (with_diag) dg = *diag;
    size_t smem = 128 + (size_t)(with_diag ? 4 : 1) * P * 32 * 8 + (size_t)N * 32 * 8;
    bool tma_ok = (inner % 32 == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0) && ((size_t)N * 32 * 8 < (1u << 20));
    int use_tma = (staging == STAGING_CPASYNC) ? 0 : (tma_ok ? 1 : 0);
    long long nblk = (nlines + 31) / 32;
    cudaError_t e;
    RemapDst rr = rd;
    if (rd.on) rr.block_rot = remap_block_rotation(rd, nblk);
#define SLLB_SPLIT_LAUNCH(...)                                                                                  \
    do {                                                                                                        \
        auto kern = k_spline_strided_split<__VA_ARGS__>;                                                        \
        e = set_smem(kern, smem);                                                                               \
        if (e != cudaSuccess) return e;                                                                         \
        kern<<<(unsigned)nblk, 32 * P, smem, st>>>(f, N, inner, dd, use_tma, nlines, linesum, rr, dg, sub);     \
    } while (0)
    // compile-time line length: 128 or 64 points, at least 16 per chunk
    const int nc = ((N == 128 || N == 64) && N / P >= 16 && g_split_const_len >= (with_diag ? 1 : 2)) ? N : 0;
    if (rd.on) {
        if (nc == 128) SLLB_SPLIT_LAUNCH(P, true, false, 128);
        else if (nc == 64) SLLB_SPLIT_LAUNCH(P, true, false, 64);
        else SLLB_SPLIT_LAUNCH(P, true);
    } else if (with_diag) {
        if (nc == 128) SLLB_SPLIT_LAUNCH(P, false, true, 128);
        else if (nc == 64) SLLB_SPLIT_LAUNCH(P, false, true, 64);
        else SLLB_SPLIT_LAUNCH(P, false, true);
    } else {
        if (nc == 128) SLLB_SPLIT_LAUNCH(P, false, false, 128);
        else if (nc == 64) SLLB_SPLIT_LAUNCH(P, false, false, 64);
        else SLLB_SPLIT_LAUNCH(P, false);
    }
#undef SLLB_SPLIT_LAUNCH
    COUNT_LAUNCH();
    return cudaGetLastError();
}

template <int BW, int S>
static cudaError_t launch_lagrange_split_t(double *f, long long nlines, int N, long long inner, const DispDesc &dd,
                                           int staging, cudaStream_t st, int P) {
    const size_t smem = 128 + (size_t)N * BW * 8;
    auto kern = k_lagrange_strided_split<BW, S>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    const bool tma_ok = (inner % BW == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0) && ((size_t)N * BW * 8 < (1u << 20));
    const int use_tma = (staging == STAGING_CPASYNC) ? 0 : (tma_ok ? 1 : 0);
    const long long nblk = (nlines + BW - 1) / BW;
    kern<<<(unsigned)nblk, BW * P, smem, st>>>(f, nlines, N, inner, dd, use_tma);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

template <int METHOD, int S>
static cudaError_t launch_strided(double *f, long long nlines, int N, long long inner, const DispDesc &dd, int staging,
                                  cudaStream_t st, const RemapDst &rd) {
    if constexpr (METHOD == 1 && S <= 11) { // the chunked kernel is bounded to 1024 threads (64 registers): short stencils only
        // few, long lines (the 1D1V problems): one thread per line leaves most SMs idle -- cut the lines into chunks
        const long long blocks32 = (nlines + 31) / 32;
        if (!rd.on && blocks32 < 2 * 148 && N >= 128) {
            int P = 1;
            while (P < 32 && N / (2 * P) >= 16 && blocks32 * P < 8 * 148) P *= 2;
            if (P > 1) {
                // narrowest group of lanes that still gives every SM a block, widest that fits otherwise
                int bw = 32;
                while (bw > 8 && ((size_t)N * bw * 8 + 128 > SMEM_MAX || (nlines + bw - 1) / bw < 148)) bw /= 2;
                if ((size_t)N * bw * 8 + 128 <= SMEM_MAX && bw * P <= 1024) {
                    if (bw == 32) return launch_lagrange_split_t<32, S>(f, nlines, N, inner, dd, staging, st, P);
                    if (bw == 16) return launch_lagrange_split_t<16, S>(f, nlines, N, inner, dd, staging, st, P);
                    return launch_lagrange_split_t<8, S>(f, nlines, N, inner, dd, staging, st, P);
                }
            }
        }
    }
    // widest tile that fits: BW lanes x N points x 8 B (+128 B header)
    if ((size_t)N * 32 * 8 + 128 <= SMEM_MAX && (inner >= 32 || nlines >= 32))
        return launch_strided_t<32, METHOD, S>(f, nlines, N, inner, dd, staging, st, rd);
    if ((size_t)N * 16 * 8 + 128 <= SMEM_MAX) return launch_strided_t<16, METHOD, S>(f, nlines, N, inner, dd, staging, st, rd);
    if ((size_t)N * 8 * 8 + 128 <= SMEM_MAX) return launch_strided_t<8, METHOD, S>(f, nlines, N, inner, dd, staging, st, rd);
    return cudaErrorInvalidValue;
}

template <int S>
static cudaError_t launch_lagrange_contig(double *f, long long nlines, int N, const DispDesc &dd, int staging,
                                          cudaStream_t st) {
    // lines per tile: aim at ~32 KB tiles, at least 1 line
    int LT = (int)(4096 / N);
    if (LT < 1) LT = 1;
    if (LT > nlines) LT = (int)nlines;
    size_t smem = 128 + (size_t)LT * N * 8 + (size_t)LT * S * 8 + (size_t)LT * 4 + 16;
    if (smem > SMEM_MAX) return cudaErrorInvalidValue;
    const bool pow2 = (N & (N - 1)) == 0;
    auto kern = pow2 ? k_lagrange_contig<S, true> : k_lagrange_contig<S, false>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    bool tma_ok = (((long long)LT * N) % 2 == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0) && (nlines % LT == 0) &&
                  ((size_t)LT * N * 8 < (1u << 20));
    int use_tma = (staging == STAGING_CPASYNC) ? 0 : (tma_ok ? 1 : 0);
    long long nblk = (nlines + LT - 1) / LT;
    kern<<<(unsigned)nblk, 256, smem, st>>>(f, nlines, N, dd, LT, use_tma);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

template <int BW>
static cudaError_t launch_spline_contig_t(double *f, long long nlines, int N, const DispDesc &dd, cudaStream_t st) {
    size_t smem = 128 + ((BW * 4) / 128) * 128 + (size_t)N * (BW + 1) * 8;
    auto kern = k_spline_contig<BW>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    long long nblk = (nlines + BW - 1) / BW;
    kern<<<(unsigned)nblk, BW, smem, st>>>(f, nlines, N, dd);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

template <int P>
static cudaError_t launch_spline_contig_split_t(double *f, long long nlines, int N, const DispDesc &dd, cudaStream_t st) {
    size_t smem = 256 + (size_t)N * 32 * 8;
    auto kern = k_spline_contig_split<P>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    long long nblk = (nlines + 31) / 32;
    kern<<<(unsigned)nblk, 32 * P, smem, st>>>(f, nlines, N, dd);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

// K1c launcher.  Returns cudaErrorNotSupported when the plane does not fit the kernel's assumptions (the
// caller then runs the two passes separately).
// L2 prefetch of the next plane (cp.async.bulk.prefetch.L2, SASS UBLKPF) at the top of every iteration of the plane
// kernel.  Measured on 128^4 (profiles/r02_plane_ab_s18.log): without the fused charge density 0.851 -> 0.788 ms, with it
// 0.998 -> 1.059 ms for the generic-extent kernel; the compile-time-extent instantiations (no spills any more) gain in both
// forms (profiles/r02_plane_const_ab.log: 0.712 -> 0.694 and 0.803 -> 0.777 ms).  1 (default): the variant without the
// charge density and the compile-time-extent instantiations prefetch; 2: all; 0: none.
int g_plane_l2_prefetch = [] { const char *e = getenv("SLLB_PLANE_L2_PREFETCH"); return e ? atoi(e) : 1; }();
int g_plane_ept = 0; // tuning knob: 0 auto (register-resident variant), 16 or 32: in-place variant with that many points per thread
static int plane_ept(int n1, int n2, bool rho) {
    if (g_plane_ept == 16 && !rho) return 16;
    return 32;
}
// The charge-density accumulators of the plane kernel: 12 registers + 20 shared-memory slots per thread, or all 32 in tensor
// memory (tcgen05.alloc / ld / st used as a scratchpad).  With run-time extents the TMEM form is the slower one (1.027 vs
// 0.975 ms on 128^4, profiles/r02_plane_ab_s19_tmem.log: the kernel spills either way); with compile-time extents neither
// form spills and TMEM wins clearly: 0.663 vs 0.777 ms (profiles/r02_plane_const_ab2.log), faster than the kernel without
// the fused density.  SLLB_PLANE_TMEM: 1 / 0 force, unset = TMEM for the 128 x 128 compile-time-extent instantiation only.
int g_plane_tmem = [] { const char *e = getenv("SLLB_PLANE_TMEM"); return e ? atoi(e) : -1; }();
// 1: square 128 x 128 / 64 x 64 planes run the instantiation with compile-time extents (SLLB_PLANE_CONST_DIMS=0: the generic one)
int g_plane_const_dims = [] { const char *e = getenv("SLLB_PLANE_CONST_DIMS"); return e ? atoi(e) : 1; }();
static int plane_const_extent(int n1, int n2) {
    return (g_plane_const_dims && g_plane_ept == 0 && n1 == n2 && (n1 == 128 || n1 == 64)) ? n1 : 0;
}
static bool plane_tmem(int n1, int n2, bool rho) {
    if (!rho || g_plane_ept != 0) return false;
    if (g_plane_tmem >= 0) return g_plane_tmem != 0;
    return plane_const_extent(n1, n2) == 128;   // 64 x 64 (128 threads, 4 CTAs per SM): 0.069 vs 0.064 ms on 64^4, stays off
}
static size_t plane_smem(int n1, int n2, bool rho) {
    const size_t T = (size_t)n1 * n2 / SLLB_PR_C;
    if (g_plane_ept != 0) return 128 + (size_t)n1 * n2 * 8;
    return 128 + (size_t)n1 * n2 * 8 + 4 * T * 8 + ((rho && !plane_tmem(n1, n2, rho)) ? (SLLB_PR_C - SLLB_PR_ACCR) * T * 8 : 0);
}
int plane_grid(int n1, int n2, long long nplanes) {
    const size_t smem = plane_smem(n1, n2, true);
    const int threads = n1 * n2 / 32;
    int per_sm = (int)(SMEM_MAX / (smem + 1024));
    if (per_sm * threads > 2048) per_sm = 2048 / threads;
    if (g_plane_ept == 0 && per_sm * threads * 128 > 65536) per_sm = 65536 / (threads * 128);
    if (per_sm < 1) per_sm = 1;
    long long g = 148LL * per_sm;
    return (int)(nplanes < g ? nplanes : g);
}
cudaError_t launch_spline_plane(double *f, int n1, int n2, long long nplanes, const DispDesc &dd1, const DispDesc &dd2,
                                double *rho_partial, cudaStream_t st, const RemapDst *remap) {
    if (n1 % 32 != 0 || n2 % 32 != 0 || n1 < 32 || n2 < 32) return cudaErrorNotSupported;
    RemapDst rd;
    if (remap && remap->on) {
        if (g_plane_ept != 0 || remap->axis != 1) return cudaErrorNotSupported;
        rd = *remap;
    } else {
        memset(&rd, 0, sizeof(rd));
        rd.base[0] = f;
    }
    const bool rho = rho_partial != nullptr;
    const int ept = plane_ept(n1, n2, rho);
    const int threads = n1 * n2 / ept;
    if (threads > 16384 / ept || threads < 32) return cudaErrorNotSupported;
    const size_t smem = plane_smem(n1, n2, rho);
    if (smem > SMEM_MAX || (size_t)n1 * n2 * 8 >= (1u << 20)) return cudaErrorNotSupported;
    if ((reinterpret_cast<uintptr_t>(f) & 15) != 0) return cudaErrorNotSupported;
    // both displacements must be constant over a plane
    const bool c1 = (dd1.istr == 0 || dd1.imod == 1) && (dd1.ostr == 0 || dd1.omod == 1 || dd1.odiv % n2 == 0);
    const bool c2 = (dd2.istr == 0 || dd2.imod == 1);
    if (!c1 || !c2) return cudaErrorNotSupported;
    cudaError_t e = ensure_constants();
    if (e != cudaSuccess) return e;
    const int grid = plane_grid(n1, n2, nplanes);
#define SLLB_PLANE_LAUNCH_R(KERN)                                                            \
    do {                                                                                     \
        e = set_smem(KERN, smem);                                                            \
        if (e != cudaSuccess) return e;                                                      \
        KERN<<<grid, threads, smem, st>>>(f, n1, n2, nplanes, dd1, dd2, rho_partial, rd, (g_plane_l2_prefetch >= 2 || (g_plane_l2_prefetch == 1 && (!rho || nc == 128 || nc == 64))) ? 1 : 0);    \
    } while (0)
#define SLLB_PLANE_LAUNCH(KERN)                                                              \
    do {                                                                                     \
        e = set_smem(KERN, smem);                                                            \
        if (e != cudaSuccess) return e;                                                      \
        KERN<<<grid, threads, smem, st>>>(f, n1, n2, nplanes, dd1, dd2, rho_partial);        \
    } while (0)
    const int nc = plane_const_extent(n1, n2);
    const bool tm = plane_tmem(n1, n2, rho);
    if (nc == 128) {
        if (rd.on) { if (tm) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, true, true, 128>)); else if (rho) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, true, false, 128>)); else SLLB_PLANE_LAUNCH_R((k_spline_plane_r<false, true, false, 128>)); }
        else { if (tm) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, false, true, 128>)); else if (rho) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, false, false, 128>)); else SLLB_PLANE_LAUNCH_R((k_spline_plane_r<false, false, false, 128>)); }
    } else if (nc == 64) {
        if (rd.on) { if (tm) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, true, true, 64>)); else if (rho) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, true, false, 64>)); else SLLB_PLANE_LAUNCH_R((k_spline_plane_r<false, true, false, 64>)); }
        else { if (tm) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, false, true, 64>)); else if (rho) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, false, false, 64>)); else SLLB_PLANE_LAUNCH_R((k_spline_plane_r<false, false, false, 64>)); }
    } else if (g_plane_ept == 0) {
        if (rd.on) {
            if (tm) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, true, true>));
            else if (rho) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, true>));
            else SLLB_PLANE_LAUNCH_R((k_spline_plane_r<false, true>));
        } else {
            if (tm) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, false, true>));
            else if (rho) SLLB_PLANE_LAUNCH_R((k_spline_plane_r<true, false>));
            else SLLB_PLANE_LAUNCH_R((k_spline_plane_r<false, false>));
        }
    } else if (rho) SLLB_PLANE_LAUNCH((k_spline_plane<32, true>));
    else if (ept == 16) SLLB_PLANE_LAUNCH((k_spline_plane<16, false>));
    else SLLB_PLANE_LAUNCH((k_spline_plane<32, false>));
    COUNT_LAUNCH();
    return cudaGetLastError();
}
__global__ void k_reduce_stage2(const double *__restrict__ partial, long long nx, int nchunks, double scale,
                                double *__restrict__ rho);   // defined with K3 below
// rho[x] = scale * sum_b partial[b][x]
cudaError_t launch_sum_partials(const double *partial, long long nx, int nparts, double scale, double *rho, cudaStream_t st) {
    k_reduce_stage2<<<(unsigned)((nx + 31) / 32), 256, 0, st>>>(partial, nx, nparts, scale, rho);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t launch_advect(double *f, long long outer, int n, long long inner, int method, int order,
                          const DispDesc &dd, int staging, cudaStream_t st, const RemapDst *remap, double *linesum,
                          const LineDiag *diag, const LineSub *sub) {
    if (n < 8 || outer < 1 || inner < 1) return cudaErrorInvalidValue;
    RemapDst rd;
    if (remap && remap->on) {
        if (inner == 1) return cudaErrorInvalidValue; // fused remap is implemented for the strided kernels
        rd = *remap;
    } else {
        memset(&rd, 0, sizeof(rd));
        rd.base[0] = f;
    }
    cudaError_t e = ensure_constants();
    if (e != cudaSuccess) return e;
    const long long nlines = outer * inner;
    if (nlines > 0x7fffffffLL * 8) return cudaErrorInvalidValue;
    // line subsets: chunked strided spline kernel only
    if (sub && sub->icount > 0 && !(method == METHOD_SPLINE && order == 4 && inner > 1 && (size_t)n * 32 * 8 + 128 + 8 * 32 * 8 <= SMEM_MAX &&
                                    ((n % 4 == 0 && n / 4 >= 8) || (n % 2 == 0 && n / 2 >= 8))))
        return cudaErrorNotSupported;
    if (method == METHOD_SPLINE && (order == 6 || order == 8)) {
        // quintic / septic periodic splines (sll_p_spline of order 6 / 8): thread per line on the strided tiles, for the
        // contiguous axis too (there the rows of the tile are gathered element by element)
        if (linesum || diag) return cudaErrorNotSupported;
        return order == 6 ? launch_strided<2, 6>(f, nlines, n, inner, dd, staging, st, rd)
                          : launch_strided<2, 8>(f, nlines, n, inner, dd, staging, st, rd);
    }
    if (method == METHOD_SPLINE) {
        if (order != 4) return cudaErrorInvalidValue;
        if (inner == 1 && (linesum || diag)) return cudaErrorNotSupported;
        if (inner == 1) {
            // split kernel: one bulk TMA copy per 32-line tile (16-byte granularity) and P chunks per line
            const bool tile_ok = (size_t)n * 32 * 8 + 256 <= SMEM_MAX && (size_t)n * 32 * 8 < (1u << 20) &&
                                 (reinterpret_cast<uintptr_t>(f) & 15) == 0 && staging != STAGING_CPASYNC &&
                                 (n % 2 == 0 || (nlines % 32 == 0)) && ((nlines % 32) * n) % 2 == 0;
            if (tile_ok && g_spline_split != 1) {
                int P = g_spline_split;
                if (P < 0) P = (n % 4 == 0 && n / 4 >= 32) ? 4 : ((n % 2 == 0 && n / 2 >= 32) ? 2 : 0);
                if (P == 8 && n % 8 == 0 && n / 8 >= 30) return launch_spline_contig_split_t<8>(f, nlines, n, dd, st);
                if (P == 4 && n % 4 == 0 && n / 4 >= 30) return launch_spline_contig_split_t<4>(f, nlines, n, dd, st);
                if (P == 2 && n % 2 == 0 && n / 2 >= 30) return launch_spline_contig_split_t<2>(f, nlines, n, dd, st);
            }
            if ((size_t)n * 33 * 8 + 256 <= SMEM_MAX) return launch_spline_contig_t<32>(f, nlines, n, dd, st);
            if ((size_t)n * 17 * 8 + 256 <= SMEM_MAX) return launch_spline_contig_t<16>(f, nlines, n, dd, st);
            if ((size_t)n * 9 * 8 + 256 <= SMEM_MAX) return launch_spline_contig_t<8>(f, nlines, n, dd, st);
            return cudaErrorInvalidValue;
        }
        if ((size_t)n * 32 * 8 + 128 + 8 * 32 * 8 <= SMEM_MAX) {
            int P = g_spline_split;
            if (P < 0) P = (n % 4 == 0 && n / 4 >= 32) ? 4 : ((n % 2 == 0 && n / 2 >= 32) ? 2 : 1);
            if (P == 4 && n % 4 == 0 && n / 4 >= 8) return launch_spline_split_t<4>(f, nlines, n, inner, dd, staging, st, rd, linesum, diag, sub);
            if (P == 2 && n % 2 == 0 && n / 2 >= 8) return launch_spline_split_t<2>(f, nlines, n, inner, dd, staging, st, rd, linesum, diag, sub);
            if (P == 8 && n % 8 == 0 && n / 8 >= 8) return launch_spline_split_t<8>(f, nlines, n, inner, dd, staging, st, rd, linesum, diag, sub);
        }
        if (linesum || diag || (sub && sub->icount > 0)) return cudaErrorNotSupported; // line sums / subsets: chunked kernel only
        return launch_strided<0, 0>(f, nlines, n, inner, dd, staging, st, rd);
    }
    if (linesum || diag) return cudaErrorNotSupported;
#define LAGR_CASE(SS)                                                                            \
    case SS:                                                                                     \
        if (inner == 1) return launch_lagrange_contig<SS>(f, nlines, n, dd, staging, st);       \
        return launch_strided<1, SS>(f, nlines, n, inner, dd, staging, st, rd);
    if (method == METHOD_LAGRANGE_FIXED) {
        switch (order) {
            LAGR_CASE(3) LAGR_CASE(5) LAGR_CASE(7) LAGR_CASE(9) LAGR_CASE(11)
        default: return cudaErrorInvalidValue;
        }
    }
    if (method == METHOD_LAGRANGE_CENTERED) {
        switch (order) {
            LAGR_CASE(4) LAGR_CASE(6) LAGR_CASE(8) LAGR_CASE(10) LAGR_CASE(12) LAGR_CASE(14) LAGR_CASE(16) LAGR_CASE(18)
        default: return cudaErrorInvalidValue;
        }
    }
    return cudaErrorInvalidValue;
}

template <int S>
static cudaError_t launch_lagrange_halo_t(double *f, const double *hl, const double *hr, long long nlines, int n,
                                          long long inner, const DispDesc &dd, int staging, cudaStream_t st, const LineBox &lbx) {
    constexpr int BW = 32;
    const int R = n + (S - 1);
    size_t smem = 128 + (size_t)R * BW * 8;
    if (smem > SMEM_MAX) return cudaErrorInvalidValue;
    auto kern = k_lagrange_halo<BW, S>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    bool tma_ok = (inner % BW == 0) && (lbx.i0 % BW == 0) && (lbx.icount % BW == 0) && al16(f) && al16(hl) && al16(hr);
    int use_tma = (staging == STAGING_CPASYNC) ? 0 : (tma_ok ? 1 : 0);
    long long nblk = (nlines + BW - 1) / BW;
    kern<<<(unsigned)nblk, BW, smem, st>>>(f, hl, hr, nlines, n, inner, dd, use_tma, lbx);
    COUNT_LAUNCH();
    return cudaGetLastError();
}
cudaError_t launch_lagrange_halo(double *f, const double *halo_left, const double *halo_right, long long outer, int n,
                                 long long inner, int order, const DispDesc &dd, int staging, cudaStream_t st,
                                 const LineBox *box) {
    if (n < 1 || outer < 1 || inner < 2) return cudaErrorInvalidValue;
    LineBox lbx;
    if (box) lbx = *box;
    else { lbx.o0 = 0; lbx.ocount = outer; lbx.i0 = 0; lbx.icount = inner; }
    if (lbx.o0 < 0 || lbx.i0 < 0 || lbx.ocount < 1 || lbx.icount < 1 || lbx.o0 + lbx.ocount > outer || lbx.i0 + lbx.icount > inner)
        return cudaErrorInvalidValue;
    const long long nlines = lbx.ocount * lbx.icount;
    switch (order) {
    case 3: return launch_lagrange_halo_t<3>(f, halo_left, halo_right, nlines, n, inner, dd, staging, st, lbx);
    case 5: return launch_lagrange_halo_t<5>(f, halo_left, halo_right, nlines, n, inner, dd, staging, st, lbx);
    case 7: return launch_lagrange_halo_t<7>(f, halo_left, halo_right, nlines, n, inner, dd, staging, st, lbx);
    case 9: return launch_lagrange_halo_t<9>(f, halo_left, halo_right, nlines, n, inner, dd, staging, st, lbx);
    case 11: return launch_lagrange_halo_t<11>(f, halo_left, halo_right, nlines, n, inner, dd, staging, st, lbx);
    default: return cudaErrorInvalidValue;
    }
}
// max_blocks > 0 caps the grid (pipelined exchange: the pack kernel must leave room on every SM for the stencil kernel
// that runs beside it)
cudaError_t launch_halo_pack(const double *f, long long outer, int n, long long inner, int j0, int hw, double *buf,
                             cudaStream_t st, const LineBox *box, int max_blocks) {
    if (hw <= 0) return cudaSuccess;
    LineBox lbx;
    if (box) lbx = *box;
    else { lbx.o0 = 0; lbx.ocount = outer; lbx.i0 = 0; lbx.icount = inner; }
    const bool vec = inner % 2 == 0 && lbx.i0 % 2 == 0 && lbx.icount % 2 == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(buf) & 15) == 0;
    const long long cap = max_blocks > 0 ? max_blocks : 148LL * 64;
    long long gx = (lbx.icount / (vec ? 2 : 1) + 1023) / 1024;
    if (gx > 148 * 8) gx = 148 * 8;
    if (gx < 1) gx = 1;
    long long gy = lbx.ocount < 65535 ? lbx.ocount : 65535;
    while (gx * gy > cap && gx > 1) gx = (gx + 1) / 2;
    while (gx * gy > cap && gy > 1) gy = (gy + 1) / 2;
    if (vec) k_halo_pack<true><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, st>>>(f, n, inner, j0, hw, buf, lbx);
    else k_halo_pack<false><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, st>>>(f, n, inner, j0, hw, buf, lbx);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K3: velocity reduction, deterministic two-stage sum.  f is [nv][nx] (x fastest).
// ------------------------------------------------------------------------------------------------
static const int RED_CHUNKS = 128;
size_t reduce_scratch_doubles(long long nx, long long nv) {
    long long ch = nv < RED_CHUNKS ? nv : RED_CHUNKS;
    return (size_t)(nx * ch);
}
__global__ void __launch_bounds__(128) k_reduce_stage1(const double *__restrict__ f, long long nx, long long nv,
                                                        int nchunks, double *__restrict__ partial) {
    const long long x = (long long)blockIdx.x * 128 + threadIdx.x;
    if (x >= nx) return;
    const int c = blockIdx.y;
    const long long v0 = nv * c / nchunks, v1 = nv * (c + 1) / nchunks;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    long long v = v0;
    for (; v + 3 < v1; v += 4) {
        a0 += __ldcs(f + x + nx * v);
        a1 += __ldcs(f + x + nx * (v + 1));
        a2 += __ldcs(f + x + nx * (v + 2));
        a3 += __ldcs(f + x + nx * (v + 3));
    }
    for (; v < v1; ++v) a0 += __ldcs(f + x + nx * v);
    partial[(long long)c * nx + x] = (a0 + a1) + (a2 + a3);
}
// rho[x] = scale * sum_c partial[c][x].  Block = 32 consecutive x (one 256-byte row per part) times 8 groups of parts:
// thread (x, g) adds the parts c = g, g + 8, ... in order, the eight group sums are then added in order -- a fixed
// summation tree, eight times the parallelism of one thread per x (the plane kernel leaves up to 592 parts behind).
__global__ void __launch_bounds__(256) k_reduce_stage2(const double *__restrict__ partial, long long nx, int nchunks, double scale,
                                                       double *__restrict__ rho) {
    __shared__ double sh[8][33];
    const int xl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long long x = (long long)blockIdx.x * 32 + xl;
    double a0 = 0, a1 = 0;
    if (x < nx) {
        int c = g;
        for (; c + 8 < nchunks; c += 16) {   // two independent chains per thread: loads of both are in flight together
            a0 += partial[(long long)c * nx + x];
            a1 += partial[(long long)(c + 8) * nx + x];
        }
        if (c < nchunks) a0 += partial[(long long)c * nx + x];
    }
    sh[g][xl] = a0 + a1;
    __syncthreads();
    if (g == 0 && x < nx) {
        double a = sh[0][xl];
#pragma unroll
        for (int k = 1; k < 8; ++k) a += sh[k][xl];
        rho[x] = a * scale;
    }
}
cudaError_t launch_reduce_velocity_partials(const double *f, long long nx, long long nv, double *scratch, int *nchunks_out,
                                            cudaStream_t st) {
    int nchunks = (int)(nv < RED_CHUNKS ? nv : RED_CHUNKS);
    dim3 grid((unsigned)((nx + 127) / 128), nchunks);
    k_reduce_stage1<<<grid, 128, 0, st>>>(f, nx, nv, nchunks, scratch);
    COUNT_LAUNCH();
    *nchunks_out = nchunks;
    return cudaGetLastError();
}
cudaError_t launch_reduce_velocity(const double *f, long long nx, long long nv, double scale, double *rho,
                                   double *scratch, cudaStream_t st) {
    int nchunks = (int)(nv < RED_CHUNKS ? nv : RED_CHUNKS);
    dim3 grid((unsigned)((nx + 127) / 128), nchunks);
    k_reduce_stage1<<<grid, 128, 0, st>>>(f, nx, nv, nchunks, scratch);
    COUNT_LAUNCH();
    k_reduce_stage2<<<(unsigned)((nx + 31) / 32), 256, 0, st>>>(scratch, nx, nchunks, scale, rho);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K8: row sums for diagnostics: for each v: (sum_x f, sum_x |f|, sum_x f^2)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_sums(const double *__restrict__ f, long long nx, double *__restrict__ out3) {
    const long long v = blockIdx.x;
    const double *row = f + v * nx;
    double s0 = 0, s1 = 0, s2 = 0;
    for (long long x = threadIdx.x; x < nx; x += 256) {
        const double t = row[x];
        s0 += t; s1 += fabs(t); s2 += t * t;
    }
    __shared__ double sh[3][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    if (ln == 0) { sh[0][w] = s0; sh[1][w] = s1; sh[2][w] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0;
        for (int i = 0; i < 8; ++i) a += sh[threadIdx.x][i];
        out3[v * 3 + threadIdx.x] = a;
    }
}
cudaError_t launch_row_sums(const double *f, long long nx, long long nv, double *out3, cudaStream_t st) {
    k_row_sums<<<(unsigned)nv, 256, 0, st>>>(f, nx, out3);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K4: spectral multipliers (unnormalised cuFFT: the 1/N factors are folded in here)
// ------------------------------------------------------------------------------------------------
// 1D: E_hat(k) = -i rho_hat(k) / (k kx0) / N for k = 1..N/2-1, zero at DC and Nyquist
// (sll_m_poisson_1d_periodic.F90:148-164)
__global__ void k_poisson1d(const cufftDoubleComplex *__restrict__ r, int nc, double kx0, cufftDoubleComplex *__restrict__ e) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nc / 2) return;
    cufftDoubleComplex out = {0.0, 0.0};
    if (k >= 1 && k <= (nc - 2) / 2) {
        const double kx = (double)k * kx0;
        const double sc = (kx / (kx * kx)) / (double)nc;
        out.x = sc * r[k].y;
        out.y = -sc * r[k].x;
    }
    e[k] = out;
}
cudaError_t launch_poisson1d_mult(const cufftDoubleComplex *rho_hat, int nc, double L, cufftDoubleComplex *e_hat,
                                  cudaStream_t st) {
    k_poisson1d<<<(nc / 2 + 1 + 127) / 128, 128, 0, st>>>(rho_hat, nc, 2.0 * 3.14159265358979323846 / L, e_hat);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

// 2D (sll_m_poisson_2d_periodic.F90:285-310,355-366): kx(1,1) := 1; E1_hat = -i kx/k2 rho_hat,
// E2_hat = -i ky/k2 rho_hat, phi_hat = rho_hat/k2; ky uses the NEGATIVE Nyquist wavenumber (:299-302).
// FFTW's c2r drops Im of the i=1 and i=N1/2+1 planes after the complex transform along x2; the same
// result is obtained from a Hermitian-consistent spectrum: those two columns are replaced by their
// Hermitian part in x2, X(j) <- (X(j) + conj X(-j))/2.
__device__ __forceinline__ void p2d_vals(const cufftDoubleComplex *r, int nh, int n2, int i, int j, double kx0,
                                         double ky0, double &phr, double &phi_, double &e1r, double &e1i, double &e2r,
                                         double &e2i) {
    double kx = (double)i * kx0;
    const double ky = (double)((j < n2 / 2) ? j : j - n2) * ky0;
    if (i == 0 && j == 0) kx = 1.0;
    const double k2 = kx * kx + ky * ky;
    const double kxs = kx / k2, kys = ky / k2;
    const cufftDoubleComplex v = r[i + (size_t)j * nh];
    phr = v.x / k2; phi_ = v.y / k2;
    e1r = kxs * v.y; e1i = -kxs * v.x; // -i kxs (a + ib) = kxs b - i kxs a
    e2r = kys * v.y; e2i = -kys * v.x;
}
__global__ void k_poisson2d(const cufftDoubleComplex *__restrict__ r, int n1, int n2, double kx0, double ky0,
                            cufftDoubleComplex *__restrict__ ph, cufftDoubleComplex *__restrict__ e1,
                            cufftDoubleComplex *__restrict__ e2) {
    const int nh = n1 / 2 + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nh) return;
    double a[6];
    p2d_vals(r, nh, n2, i, j, kx0, ky0, a[0], a[1], a[2], a[3], a[4], a[5]);
    if (i == 0 || 2 * i == n1) {
        double b[6];
        const int jm = (j == 0) ? 0 : n2 - j;
        p2d_vals(r, nh, n2, i, jm, kx0, ky0, b[0], b[1], b[2], b[3], b[4], b[5]);
#pragma unroll
        for (int c = 0; c < 6; c += 2) {
            a[c] = 0.5 * (a[c] + b[c]);
            a[c + 1] = 0.5 * (a[c + 1] - b[c + 1]);
        }
    }
    const double nrm = 1.0 / ((double)n1 * (double)n2);
    const size_t idx = i + (size_t)j * nh;
    if (ph) ph[idx] = make_cuDoubleComplex(a[0] * nrm, a[1] * nrm);
    if (e1) e1[idx] = make_cuDoubleComplex(a[2] * nrm, a[3] * nrm);
    if (e2) e2[idx] = make_cuDoubleComplex(a[4] * nrm, a[5] * nrm);
}
cudaError_t launch_poisson2d_mult(const cufftDoubleComplex *rho_hat, int n1, int n2, double L1, double L2,
                                  cufftDoubleComplex *phi_hat, cufftDoubleComplex *e1_hat, cufftDoubleComplex *e2_hat,
                                  cudaStream_t st) {
    const double tp = 2.0 * 3.14159265358979323846;
    dim3 grid((n1 / 2 + 1 + 63) / 64, n2);
    k_poisson2d<<<grid, 64, 0, st>>>(rho_hat, n1, n2, tp / L1, tp / L2, phi_hat, e1_hat, e2_hat);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

// 2D, parallel variant (sll_s_poisson_2d_periodic_par_solve, sll_m_poisson_2d_periodic_par.F90:214-338): Delta phi = rho,
// phi^ = -rho^ / (4 pi^2 ((kx/Lx)^2 + (ky/Ly)^2)) with both indices folded to -n/2 .. n/2-1 (:300-306), phi^(0,0) = 0.
// The reference transforms complex-to-complex and keeps the real part; the multiplier is real and even, so the half
// spectrum of a real-to-complex transform gives the same field.
__global__ void k_poisson2d_par(const cufftDoubleComplex *__restrict__ r, int n1, int n2, double kx0, double ky0,
                                cufftDoubleComplex *__restrict__ ph) {
    const int nh = n1 / 2 + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nh) return;
    const size_t idx = i + (size_t)j * nh;
    cufftDoubleComplex out = {0.0, 0.0};
    if (i + j != 0) {
        const double kx = (double)i * kx0, ky = (double)((j < n2 / 2) ? j : j - n2) * ky0;
        const double d = -(kx * kx + ky * ky) * ((double)n1 * (double)n2);
        out.x = r[idx].x / d; out.y = r[idx].y / d;
    }
    ph[idx] = out;
}
cudaError_t launch_poisson2d_par_mult(const cufftDoubleComplex *rho_hat, int n1, int n2, double L1, double L2,
                                      cufftDoubleComplex *phi_hat, cudaStream_t st) {
    const double tp = 2.0 * 3.14159265358979323846;
    dim3 grid((n1 / 2 + 1 + 63) / 64, n2);
    k_poisson2d_par<<<grid, 64, 0, st>>>(rho_hat, n1, n2, tp / L1, tp / L2, phi_hat);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

// 3D (sll_m_poisson_3d_periodic_par.F90:403-441,1079-1158): phi_hat = rho_hat/(N^3 |k|^2), zero mean;
// E_d_hat = -i k_d phi_hat with the signed wavenumber and the Nyquist mode of direction d dropped
// (the reference multiplies it by a purely imaginary factor and keeps the real part).
__global__ void k_poisson3d(const cufftDoubleComplex *__restrict__ r, int n1, int n2, int n3, double k10, double k20,
                            double k30, cufftDoubleComplex *__restrict__ ph, cufftDoubleComplex *__restrict__ e1,
                            cufftDoubleComplex *__restrict__ e2, cufftDoubleComplex *__restrict__ e3) {
    const int nh = n1 / 2 + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= nh) return;
    const size_t idx = i + (size_t)nh * (j + (size_t)n2 * k);
    const double f1 = (double)i, f2 = (double)((j < n2 / 2) ? j : n2 - j), f3 = (double)((k < n3 / 2) ? k : n3 - k);
    const double kx = k10 * f1, ky = k20 * f2, kz = k30 * f3;
    cufftDoubleComplex p = {0.0, 0.0};
    if (i + j + k != 0) {
        const double sc = 1.0 / ((double)n1 * (double)n2 * (double)n3) / (kx * kx + ky * ky + kz * kz);
        p.x = r[idx].x * sc; p.y = r[idx].y * sc;
    }
    if (ph) ph[idx] = p;
    const double s1 = (2 * i == n1) ? 0.0 : k10 * (double)i;
    const double s2 = (2 * j == n2) ? 0.0 : k20 * (double)((j < n2 / 2) ? j : j - n2);
    const double s3 = (2 * k == n3) ? 0.0 : k30 * (double)((k < n3 / 2) ? k : k - n3);
    if (e1) e1[idx] = make_cuDoubleComplex(s1 * p.y, -s1 * p.x);
    if (e2) e2[idx] = make_cuDoubleComplex(s2 * p.y, -s2 * p.x);
    if (e3) e3[idx] = make_cuDoubleComplex(s3 * p.y, -s3 * p.x);
}
cudaError_t launch_poisson3d_mult(const cufftDoubleComplex *rho_hat, int n1, int n2, int n3, double L1, double L2,
                                  double L3, cufftDoubleComplex *phi_hat, cufftDoubleComplex *e1_hat,
                                  cufftDoubleComplex *e2_hat, cufftDoubleComplex *e3_hat, cudaStream_t st) {
    const double tp = 2.0 * 3.14159265358979323846;
    dim3 grid((n1 / 2 + 1 + 31) / 32, n2, n3);
    k_poisson3d<<<grid, 32, 0, st>>>(rho_hat, n1, n2, n3, tp / L1, tp / L2, tp / L3, phi_hat, e1_hat, e2_hat, e3_hat);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void k_affine(double *out, long long n, double a0, double a1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a0 + a1 * (double)i;
}
cudaError_t launch_affine(double *out, long long n, double a0, double a1, cudaStream_t st) {
    k_affine<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, n, a0, a1);
    COUNT_LAUNCH();
    return cudaGetLastError();
}
__global__ void k_rho_1d1v(double *rho, long long n, double c0, double c1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rho[i] = c0 - c1 * rho[i];
}
cudaError_t launch_rho_1d1v(double *rho, long long n, double c0, double c1, cudaStream_t st) {
    k_rho_1d1v<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rho, n, c0, c1);
    COUNT_LAUNCH();
    return cudaGetLastError();
}
// single block deterministic sum of squares (fields are small: <= 128^2 or 32^3 values)
__global__ void __launch_bounds__(1024) k_sum_squares(const double *__restrict__ a, long long n, double *out1) {
    double s = 0;
    for (long long i = threadIdx.x; i < n; i += 1024) s += a[i] * a[i];
    __shared__ double sh[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < 32; ++i) t += sh[i];
        out1[0] = t;
    }
}
cudaError_t launch_sum_squares(const double *a, long long n, double *out1, cudaStream_t st) {
    k_sum_squares<<<1, 1024, 0, st>>>(a, n, out1);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

// K6: box <-> contiguous buffer (column-major order inside the box, like the reference's pack loops)
__global__ void __launch_bounds__(256) k_pack4d(const double *__restrict__ src, int e0, int e1, int e2, Box4 b,
                                                double *__restrict__ buf, int unpack, double *__restrict__ dst) {
    const long long n = (long long)b.n[0] * b.n[1] * b.n[2] * b.n[3];
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
        long long r = t;
        const int i0 = (int)(r % b.n[0]); r /= b.n[0];
        const int i1 = (int)(r % b.n[1]); r /= b.n[1];
        const int i2 = (int)(r % b.n[2]); r /= b.n[2];
        const int i3 = (int)r;
        const long long g = (b.lo[0] + i0) + (long long)e0 * ((b.lo[1] + i1) + (long long)e1 * ((b.lo[2] + i2) + (long long)e2 * (b.lo[3] + i3)));
        if (unpack) dst[g] = buf[t];
        else buf[t] = src[g];
    }
}
cudaError_t launch_pack4d(const double *src, const int ext[4], Box4 box, double *buf, cudaStream_t st) {
    long long n = (long long)box.n[0] * box.n[1] * box.n[2] * box.n[3];
    if (n <= 0) return cudaSuccess;
    unsigned nb = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    k_pack4d<<<nb, 256, 0, st>>>(src, ext[0], ext[1], ext[2], box, buf, 0, nullptr);
    COUNT_LAUNCH();
    return cudaGetLastError();
}
cudaError_t launch_unpack4d(double *dst, const int ext[4], Box4 box, const double *buf, cudaStream_t st) {
    long long n = (long long)box.n[0] * box.n[1] * box.n[2] * box.n[3];
    if (n <= 0) return cudaSuccess;
    unsigned nb = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    k_pack4d<<<nb, 256, 0, st>>>(nullptr, ext[0], ext[1], ext[2], box, const_cast<double *>(buf), 1, dst);
    COUNT_LAUNCH();
    return cudaGetLastError();
}

} // namespace sllb
