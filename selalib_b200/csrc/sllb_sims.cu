// sllb_sims.cu -- layouts / remap (a13), NCCL communicator, and the time loops of the three
// simulations the hot path serves, running entirely on the device.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sllb_internal.h"

using namespace sllb;

namespace {
// split_array_indices, src/parallelization/remap/sll_m_remapper.F90:1860-1939 (0-based, inclusive)
void split_aux(std::vector<int> &lo, std::vector<int> &hi, int start, int seglen, int mn, int mx) {
    if (seglen == 1) { lo[start] = mn; hi[start] = mx; return; }
    const int num = mx - mn + 1;
    int max1 = (num % 2 == 0) ? mn + num / 2 - 1 : mn + num / 2;
    split_aux(lo, hi, start, seglen / 2, mn, max1);
    split_aux(lo, hi, start + seglen / 2, seglen / 2, max1 + 1, mx);
}
bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
} // namespace

extern "C" {

/* sll_s_factorize_in_two_powers_of_two, sll_m_remapper.F90:6452-6476 */
int sllb_factorize_in_two_powers_of_two(int num_procs, int *f1, int *f2) {
    if (!f1 || !f2 || !is_pow2(num_procs)) return fail(SLLB_ERR_INVALID, "factorize: num_procs must be a power of two");
    int exponent = 0;
    while ((1 << exponent) < num_procs) ++exponent;
    if (exponent > 0 && exponent % 2 == 0) { *f1 = 1 << (exponent / 2); *f2 = 1 << (exponent / 2); }
    else if (exponent == 0) { *f1 = 1; *f2 = 1; }
    else { *f1 = 1 << ((exponent - 1) / 2); *f2 = 1 << ((exponent + 1) / 2); }
    return SLLB_OK;
}

/* initialize_layout_with_distributed_4D_array, sll_m_remapper.F90:1345-1487; rank order :1056-1066 */
int sllb_layout4d_boxes(const int global[4], const int procs[4], int nranks, int *boxes) {
    if (!global || !procs || !boxes) return fail(SLLB_ERR_INVALID, "layout4d_boxes: null");
    if ((long long)procs[0] * procs[1] * procs[2] * procs[3] != nranks)
        return fail(SLLB_ERR_INVALID, "layout4d_boxes: process mesh does not match the number of ranks");
    std::vector<int> lo[4], hi[4];
    for (int d = 0; d < 4; ++d) {
        if (!is_pow2(procs[d])) return fail(SLLB_ERR_INVALID, "layout4d_boxes: process mesh must be powers of two (sll_m_remapper.F90:1389-1397)");
        if (global[d] < procs[d]) return fail(SLLB_ERR_INVALID, "layout4d_boxes: fewer points than processes along an axis");
        lo[d].resize(procs[d]); hi[d].resize(procs[d]);
        split_aux(lo[d], hi[d], 0, procs[d], 0, global[d] - 1);
    }
    for (int l = 0; l < procs[3]; ++l) for (int k = 0; k < procs[2]; ++k) for (int j = 0; j < procs[1]; ++j) for (int i = 0; i < procs[0]; ++i) {
        const int r = i + procs[0] * (j + procs[1] * (k + procs[2] * l));
        const int c[4] = {i, j, k, l};
        for (int d = 0; d < 4; ++d) { boxes[r * 8 + 2 * d] = lo[d][c[d]]; boxes[r * 8 + 2 * d + 1] = hi[d][c[d]]; }
    }
    return SLLB_OK;
}

/* box intersections = the remap plan (sll_m_remapper.F90:2073-2157) */
int sllb_remap4d_plan(const int global[4], const int procs_from[4], const int procs_to[4], int nranks, int rank,
                      int *send_boxes, int *recv_boxes) {
    if (rank < 0 || rank >= nranks || !send_boxes || !recv_boxes) return fail(SLLB_ERR_INVALID, "remap4d_plan: bad arguments");
    std::vector<int> from((size_t)nranks * 8), to((size_t)nranks * 8);
    SLLB_TRY(sllb_layout4d_boxes(global, procs_from, nranks, from.data()));
    SLLB_TRY(sllb_layout4d_boxes(global, procs_to, nranks, to.data()));
    for (int r = 0; r < nranks; ++r)
        for (int d = 0; d < 4; ++d) {
            // what I (in the source layout) send to r (its target box); what I (target) receive from r (its source box)
            send_boxes[r * 8 + 2 * d] = std::max(from[rank * 8 + 2 * d], to[r * 8 + 2 * d]);
            send_boxes[r * 8 + 2 * d + 1] = std::min(from[rank * 8 + 2 * d + 1], to[r * 8 + 2 * d + 1]);
            recv_boxes[r * 8 + 2 * d] = std::max(to[rank * 8 + 2 * d], from[r * 8 + 2 * d]);
            recv_boxes[r * 8 + 2 * d + 1] = std::min(to[rank * 8 + 2 * d + 1], from[r * 8 + 2 * d + 1]);
        }
    return SLLB_OK;
}

/* sll_f_set_process_grid, src/parallelization/decomposition/sll_m_decomposition.F90:2473-2543:
 * powers of two are distributed starting from the LAST (velocity) dimensions. */
int sllb_set_process_grid(int nranks, int grid[6]) {
    if (!grid || !is_pow2(nranks)) return fail(SLLB_ERR_INVALID, "set_process_grid: number of ranks must be a power of two");
    for (int d = 0; d < 6; ++d) grid[d] = 1;
    int rem = nranks, d = 5;
    while (rem > 1) { grid[d] *= 2; rem /= 2; d = (d == 0) ? 5 : d - 1; }
    return SLLB_OK;
}

} // extern "C"

/* ------------------------------------------------------------------------------------------ */
/* communicator                                                                                */
/* ------------------------------------------------------------------------------------------ */

extern "C" {
int sllb_comm_unique_id(void *id128) {
    if (!id128) return fail(SLLB_ERR_INVALID, "comm_unique_id: null");
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    SLLB_NCCL(ncclGetUniqueId(&id));
    memcpy(id128, &id, 128);
    return SLLB_OK;
}
int sllb_comm_create(const void *id128, int nranks, int rank, sllb_comm_t *c) {
    if (!id128 || !c || nranks < 1 || rank < 0 || rank >= nranks) return fail(SLLB_ERR_INVALID, "comm_create: bad arguments");
    SLLB_TRY(require_device());
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    sllb_comm *cc = new sllb_comm();
    cc->nranks = nranks; cc->rank = rank;
    int rc = check_nccl(ncclCommInitRank(&cc->comm, nranks, id, rank), "ncclCommInitRank");
    if (rc) { delete cc; return rc; }
    *c = cc;
    return SLLB_OK;
}
int sllb_comm_destroy(sllb_comm_t c) {
    if (!c) return SLLB_OK;
    if (c->comm) ncclCommDestroy(c->comm);
    delete c;
    return SLLB_OK;
}
int sllb_comm_allreduce_sum(sllb_comm_t c, double *d_buf, int64_t count) {
    if (!c || !d_buf) return fail(SLLB_ERR_INVALID, "allreduce: null");
    SLLB_NCCL(ncclAllReduce(d_buf, d_buf, (size_t)count, ncclDouble, ncclSum, c->comm, g_stream));
    return SLLB_OK;
}
int sllb_comm_allgather(sllb_comm_t c, const double *d_send, double *d_recv, int64_t count_per_rank) {
    if (!c || !d_send || !d_recv) return fail(SLLB_ERR_INVALID, "allgather: null");
    SLLB_NCCL(ncclAllGather(d_send, d_recv, (size_t)count_per_rank, ncclDouble, c->comm, g_stream));
    return SLLB_OK;
}
} // extern "C"

/* ------------------------------------------------------------------------------------------ */
/* phase timers (CUDA events on the launch stream)                                              */
/* ------------------------------------------------------------------------------------------ */
namespace {
// Opt-in (sllb_set_phase_timers): a mark is one cudaEventRecord on the launch stream; the events come from a pool that
// lives as long as the simulation, so a long run neither creates nor destroys events per stage.
struct PhaseTimer {
    std::vector<cudaEvent_t> ev;   // pool
    std::vector<int> tag;
    size_t used = 0;
    bool on = false;
    void begin(bool enable) { used = 0; tag.clear(); on = enable; }
    void mark(int phase_just_finished) {
        if (!on) return;
        if (used == ev.size()) {
            if (ev.size() >= 4096) { // bounded: fold what has been recorded so far, keep the last event as the new origin
                fold();
                std::swap(ev[0], ev[used - 1]);
                used = 1; tag.assign(1, -1);
            } else {
                cudaEvent_t e;
                if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
                ev.push_back(e);
            }
        }
        cudaEventRecord(ev[used], g_stream);
        ++used; tag.push_back(phase_just_finished);
    }
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    void fold() {
        if (used < 2) return;
        cudaEventSynchronize(ev[used - 1]);
        for (size_t i = 1; i < used; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            if (tag[i] >= 0 && tag[i] < 8) acc[tag[i]] += ms;
        }
    }
    void collect(double out[8]) {
        fold();
        for (int k = 0; k < 8; ++k) { out[k] = acc[k]; acc[k] = 0; }
        used = 0; tag.clear(); on = false;
    }
    ~PhaseTimer() { for (auto e : ev) cudaEventDestroy(e); }
};
} // namespace
static int g_phase_timers = 0;
static void phase_mark(PhaseTimer *t, int tag) { if (t) t->mark(tag); }

/* ------------------------------------------------------------------------------------------ */
/* distributed 4D field + remap                                                                 */
/* ------------------------------------------------------------------------------------------ */
struct sllb_dist4d {
    sllb_comm *comm = nullptr;
    int nranks = 1, rank = 0;
    int global[4];
    int procs[2][4];              // [0] x-sequential (axes 2,3 split), [1] v-sequential (axes 0,1 split)
    std::vector<int> boxes[2];    // all ranks' boxes per layout
    sllb_field *F[2] = {nullptr, nullptr};
    DevBuf sendbuf, recvbuf;
    std::vector<int> sb[2], rb[2]; // plans per direction
    // fused remap: every rank's two layout arrays mapped into this process (CUDA IPC over NVLink)
    bool p2p = false;
    double *peer[2][8];
    std::vector<void *> ipc_opened;
    DevBuf flag;                   // 1 double: all-reduce used as the cross-rank barrier (NCCL fallback of the flag barrier)
    // flag barrier + density exchange over peer memory (no NCCL call in the time loop):
    //   sig[q]  = last epoch rank q has published to me (64-bit counters, written by the peers with st.release.sys)
    //   xbuf    = 8 slots of N1*N2 doubles (slot q: rank q's partial charge density) + slot 8: the density gathered from
    //             the ranks' (x1,x2) tiles
    DevBuf sigbuf, xbuf, errflag;
    unsigned long long *peer_sig[8];
    double *peer_x[8];
    unsigned long long epoch = 0;
    bool flag_barrier = false;
    long long n12 = 0;
    // chunked V stage: x3 pass of chunk c+1 (s_local, high priority) under the x4 + remap pass of chunk c (s_remap)
    cudaStream_t s_local = nullptr, s_remap = nullptr;
    cudaEvent_t ev_chunk[8] = {};
};

int g_fused_remap = 1; // 1: advect + remap in one kernel over peer memory when possible, 0: pack + NCCL + unpack

// Map every rank's F[0] and F[1] into this process (CUDA IPC, see peer_map_buffers).
static int dist4d_setup_p2p(sllb_dist4d *D) {
    D->p2p = false;
    if (D->nranks < 2 || D->nranks > 8) return SLLB_OK;
    for (int w = 0; w < 2; ++w)
        for (int d = 0; d < 4; ++d)
            if (D->global[d] % D->procs[w][d] != 0) return SLLB_OK; // non-uniform boxes: NCCL path
    const char *env = getenv("SLLB_FUSED_REMAP");
    if (env && env[0] == '0') return SLLB_OK;
    D->n12 = (long long)D->global[0] * D->global[1];
    SLLB_TRY(D->sigbuf.ensure(16));
    SLLB_TRY(D->xbuf.ensure((size_t)9 * D->n12));
    SLLB_TRY(D->errflag.ensure(1));
    SLLB_CUDA(cudaMemset(D->sigbuf.p, 0, 16 * sizeof(double)));
    SLLB_CUDA(cudaMemset(D->errflag.p, 0, sizeof(double)));
    SLLB_CUDA(cudaMemset(D->xbuf.p, 0, (size_t)9 * D->n12 * sizeof(double)));
    void *mine[4] = {D->F[0]->d, D->F[1]->d, D->sigbuf.p, D->xbuf.p};
    std::vector<void *> peers;
    bool ok = false;
    SLLB_TRY(D->flag.ensure(2));
    SLLB_TRY(peer_map_buffers(D->comm, mine, 4, peers, D->ipc_opened, &ok));
    if (ok)
        for (int r = 0; r < D->nranks; ++r) {
            for (int w = 0; w < 2; ++w) D->peer[w][r] = static_cast<double *>(peers[(size_t)r * 4 + w]);
            D->peer_sig[r] = static_cast<unsigned long long *>(peers[(size_t)r * 4 + 2]);
            D->peer_x[r] = static_cast<double *>(peers[(size_t)r * 4 + 3]);
        }
    D->p2p = ok;
    const char *eb = getenv("SLLB_FLAG_BARRIER");
    D->flag_barrier = ok && !(eb && eb[0] == '0');
    return SLLB_OK;
}

// ------------------------------------------------------------------------------------------------
// Cross-rank barrier through flags in peer-mapped memory.  Stream order puts this kernel after the kernel whose peer
// stores it guards; thread q fences (system scope), publishes the new epoch into slot [me] of rank q's flag array with a
// release store and spins with acquire loads on slot [q] of its own array.  When the kernel ends every rank has finished
// the guarded kernel, and what it stored into this GPU's memory is visible to the kernels that follow on this stream.
// One launch of a few microseconds instead of an ncclAllReduce used as a barrier (~50 us on 8 GPUs).
// A peer that never arrives (it failed) trips the time-out: the error flag is raised and checked at the end of the run.
// ------------------------------------------------------------------------------------------------
struct PeerSig { unsigned long long *p[8]; };
__global__ void __launch_bounds__(32) k_flag_barrier(const PeerSig sig, const int nranks, const int rank,
                                                     const unsigned long long epoch, double *err) {
    const int q = threadIdx.x;
    if (q < nranks && q != rank) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sig.p[q] + rank), "l"(epoch) : "memory");
        const unsigned long long *mine = sig.p[rank] + q;
        const long long t0 = clock64();
        unsigned long long seen = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
            if (seen >= epoch) break;
            if (clock64() - t0 > (20LL << 30)) { *err = 1.0; break; }   // ~10 s
        }
    }
    __syncthreads();
}
namespace sllb {
cudaError_t launch_flag_barrier(unsigned long long *const sig[8], int nranks, int rank, unsigned long long epoch, double *err,
                                cudaStream_t st) {
    PeerSig s;
    for (int r = 0; r < 8; ++r) s.p[r] = r < nranks ? sig[r] : nullptr;
    k_flag_barrier<<<1, 32, 0, st>>>(s, nranks, rank, epoch, err);
    count_launch();
    return cudaGetLastError();
}
} // namespace sllb
// all ranks call this at the same point of their streams
static int dist4d_barrier(sllb_dist4d *D) {
    if (D->nranks < 2) return SLLB_OK;
    if (!D->flag_barrier) {
        SLLB_NCCL(ncclAllReduce(D->flag.p, D->flag.p, 1, ncclDouble, ncclSum, D->comm->comm, g_stream));
        return SLLB_OK;
    }
    PeerSig sig;
    for (int r = 0; r < 8; ++r) sig.p[r] = r < D->nranks ? D->peer_sig[r] : nullptr;
    D->epoch += 1;
    k_flag_barrier<<<1, 32, 0, g_stream>>>(sig, D->nranks, D->rank, D->epoch, D->errflag.p);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_flag_barrier");
}
// my array `src` (n doubles) into slot `slot` of every OTHER rank's exchange buffer (peer stores); my own slot is written
// by the producer directly
struct PeerX { double *p[8]; };
__global__ void __launch_bounds__(256) k_bcast_slot(const double *__restrict__ src, const long long n, const PeerX dst,
                                                    const long long off, const int nranks, const int rank) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double v = src[i];
    for (int r = 0; r < nranks; ++r)
        if (r != rank) dst.p[r][off + i] = v;
}
// a dense t0 x t1 tile into the (lo0, lo1) corner of the N1 x N2 array in slot 8 of EVERY rank's exchange buffer
__global__ void __launch_bounds__(256) k_bcast_tile(const double *__restrict__ tile, const int t0, const int t1, const int lo0,
                                                    const int lo1, const int n1, const PeerX dst, const long long off,
                                                    const int nranks) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= t0 * t1) return;
    const int a0 = i % t0, a1 = i / t0;
    const double v = tile[i];
    const long long g = off + (lo0 + a0) + (long long)n1 * (lo1 + a1);
    for (int r = 0; r < nranks; ++r) dst.p[r][g] = v;
}

static void box_of(const std::vector<int> &boxes, int r, int lo[4], int n[4]) {
    for (int d = 0; d < 4; ++d) { lo[d] = boxes[r * 8 + 2 * d]; n[d] = boxes[r * 8 + 2 * d + 1] - lo[d] + 1; }
}

static void dist4d_remap_dst(sllb_dist4d *D, int from, int axis, RemapDst *rdp) {
    const int to = 1 - from;
    RemapDst &rd = *rdp;
    memset(&rd, 0, sizeof(rd));
    for (int r = 0; r < D->nranks; ++r) rd.base[r] = D->peer[to][r];
    rd.on = 1; rd.axis = axis; rd.rank = D->rank; rd.block_rot = 0;
    for (int d = 0; d < 4; ++d) {
        rd.se[d] = D->F[from]->ext[d];
        rd.slo[d] = D->boxes[from][D->rank * 8 + 2 * d];
        rd.tp[d] = D->procs[to][d];
        rd.te[d] = D->global[d] / D->procs[to][d];
    }
}
static int dist4d_advect_remap_dev(sllb_dist4d *D, int from, int axis, int method, int order, const DispDesc &dd,
                                   PhaseTimer *timer = nullptr) {
    if (!D->p2p || !g_fused_remap) return fail(SLLB_ERR_UNSUPPORTED, "dist4d_advect_remap: peer mapping not available (use advect + sllb_dist4d_remap)");
    if (D->procs[from][axis] != 1) return fail(SLLB_ERR_INVALID, "dist4d_advect_remap: the advected axis must be whole in the source layout");
    RemapDst rd;
    dist4d_remap_dst(D, from, axis, &rd);
    phase_mark(timer, 0);
    SLLB_TRY(advect_axis_dev(D->F[from], axis, method, order, dd, &rd));
    phase_mark(timer, 4);
    // all ranks' stores into my destination array are complete once every rank's kernel has finished
    SLLB_TRY(dist4d_barrier(D));
    phase_mark(timer, 5);
    return SLLB_OK;
}

extern "C" {
int sllb_dist4d_create(sllb_comm_t c, const int global[4], sllb_dist4d_t *D) {
    if (!global || !D) return fail(SLLB_ERR_INVALID, "dist4d_create: null");
    SLLB_TRY(require_device());
    sllb_dist4d *dd = new sllb_dist4d();
    dd->comm = c;
    dd->nranks = c ? c->nranks : 1;
    dd->rank = c ? c->rank : 0;
    int f1 = 1, f2 = 1;
    int rc = sllb_factorize_in_two_powers_of_two(dd->nranks, &f1, &f2);
    if (rc) { delete dd; return rc; }
    for (int d = 0; d < 4; ++d) dd->global[d] = global[d];
    const int px[4] = {1, 1, f1, f2}, pv[4] = {f1, f2, 1, 1};
    memcpy(dd->procs[0], px, sizeof(px));
    memcpy(dd->procs[1], pv, sizeof(pv));
    for (int w = 0; w < 2; ++w) {
        dd->boxes[w].resize((size_t)dd->nranks * 8);
        rc = sllb_layout4d_boxes(global, dd->procs[w], dd->nranks, dd->boxes[w].data());
        if (rc) { delete dd; return rc; }
    }
    for (int dir = 0; dir < 2; ++dir) {
        dd->sb[dir].resize((size_t)dd->nranks * 8); dd->rb[dir].resize((size_t)dd->nranks * 8);
        rc = sllb_remap4d_plan(global, dd->procs[dir], dd->procs[1 - dir], dd->nranks, dd->rank, dd->sb[dir].data(), dd->rb[dir].data());
        if (rc) { delete dd; return rc; }
    }
    int lo[4], n[4];
    box_of(dd->boxes[0], dd->rank, lo, n);
    rc = field_alloc(4, n, &dd->F[0]);
    if (rc) { delete dd; return rc; }
    if (dd->nranks == 1) {
        rc = field_wrap(4, n, dd->F[0]->d, &dd->F[1]); // both layouts coincide: no copy, no exchange
    } else {
        box_of(dd->boxes[1], dd->rank, lo, n);
        rc = field_alloc(4, n, &dd->F[1]);
        if (!rc) rc = dd->sendbuf.ensure((size_t)dd->F[0]->total > (size_t)dd->F[1]->total ? dd->F[0]->total : dd->F[1]->total);
        if (!rc) rc = dd->recvbuf.ensure(dd->sendbuf.n);
    }
    if (!rc && dd->nranks > 1) rc = dist4d_setup_p2p(dd);
    if (rc) { sllb_dist4d_destroy(dd); return rc; }
    *D = dd;
    return SLLB_OK;
}
int sllb_dist4d_p2p(sllb_dist4d_t D, int *enabled) {
    if (!D || !enabled) return fail(SLLB_ERR_INVALID, "dist4d_p2p: null");
    *enabled = (D->p2p && g_fused_remap) ? 1 : 0;
    return SLLB_OK;
}
int sllb_set_fused_remap(int on) {
    g_fused_remap = on ? 1 : 0;
    return SLLB_OK;
}
/* One advection pass along `axis` of the layout `from` whose stores land in the OTHER layout on the owning
 * ranks (peer memory over NVLink), followed by a cross-rank barrier: advect_1d_constant on every line +
 * apply_remap_4D_double (sll_m_remapper.F90:3308-3456) in one kernel. */
int sllb_dist4d_advect_remap(sllb_dist4d_t D, int from, int axis, int method, int order, const sllb_disp_t *disp) {
    if (!D || !disp || !disp->values || from < 0 || from > 1 || axis < 0 || axis > 3) return fail(SLLB_ERR_INVALID, "dist4d_advect_remap: bad arguments");
    if (!disp->values_on_device) return fail(SLLB_ERR_INVALID, "dist4d_advect_remap: displacement values must be on the device");
    DispDesc dd;
    dd.v = disp->values; dd.scale = disp->scale;
    dd.odiv = disp->odiv > 0 ? disp->odiv : 1; dd.omod = disp->omod > 0 ? disp->omod : 1; dd.ostr = disp->ostr;
    dd.idiv = disp->idiv > 0 ? disp->idiv : 1; dd.imod = disp->imod > 0 ? disp->imod : 1; dd.istr = disp->istr;
    return dist4d_advect_remap_dev(D, from, axis, method, order, dd);
}
int sllb_dist4d_destroy(sllb_dist4d_t D) {
    if (!D) return SLLB_OK;
    if (D->s_local) {
        cudaStreamDestroy(D->s_local); cudaStreamDestroy(D->s_remap);
        for (cudaEvent_t e : D->ev_chunk) if (e) cudaEventDestroy(e);
    }
    for (void *ptr : D->ipc_opened) cudaIpcCloseMemHandle(ptr);
    sllb_field_destroy(D->F[1]);
    sllb_field_destroy(D->F[0]);
    delete D;
    return SLLB_OK;
}
int sllb_dist4d_field(sllb_dist4d_t D, int which, sllb_field_t *F) {
    if (!D || !F || which < 0 || which > 1) return fail(SLLB_ERR_INVALID, "dist4d_field: bad arguments");
    *F = D->F[which];
    return SLLB_OK;
}
int sllb_dist4d_box(sllb_dist4d_t D, int which, int box[8]) {
    if (!D || !box || which < 0 || which > 1) return fail(SLLB_ERR_INVALID, "dist4d_box: bad arguments");
    for (int k = 0; k < 8; ++k) box[k] = D->boxes[which][D->rank * 8 + k];
    return SLLB_OK;
}
/* apply_remap_4D_double (sll_m_remapper.F90:3308-3456): pack per destination, exchange, unpack. */
int sllb_dist4d_remap(sllb_dist4d_t D, int direction) {
    if (!D || direction < 0 || direction > 1) return fail(SLLB_ERR_INVALID, "dist4d_remap: bad arguments");
    if (D->nranks == 1) return SLLB_OK;
    sllb_field *src = D->F[direction], *dst = D->F[1 - direction];
    int slo[4], sn[4], dlo[4], dn[4];
    box_of(D->boxes[direction], D->rank, slo, sn);
    box_of(D->boxes[1 - direction], D->rank, dlo, dn);
    const std::vector<int> &sb = D->sb[direction], &rb = D->rb[direction];
    std::vector<long long> soff(D->nranks + 1, 0), roff(D->nranks + 1, 0);
    for (int r = 0; r < D->nranks; ++r) {
        long long cs = 1, cr = 1;
        for (int d = 0; d < 4; ++d) {
            cs *= std::max(0, sb[r * 8 + 2 * d + 1] - sb[r * 8 + 2 * d] + 1);
            cr *= std::max(0, rb[r * 8 + 2 * d + 1] - rb[r * 8 + 2 * d] + 1);
        }
        soff[r + 1] = soff[r] + cs; roff[r + 1] = roff[r] + cr;
        if (cs > 0) {
            Box4 b;
            for (int d = 0; d < 4; ++d) { b.lo[d] = sb[r * 8 + 2 * d] - slo[d]; b.n[d] = sb[r * 8 + 2 * d + 1] - sb[r * 8 + 2 * d] + 1; }
            SLLB_CUDA(launch_pack4d(src->d, sn, b, D->sendbuf.p + soff[r], g_stream));
        }
    }
    SLLB_NCCL(ncclGroupStart());
    for (int r = 0; r < D->nranks; ++r) {
        const long long cs = soff[r + 1] - soff[r], cr = roff[r + 1] - roff[r];
        if (cs > 0) SLLB_NCCL(ncclSend(D->sendbuf.p + soff[r], (size_t)cs, ncclDouble, r, D->comm->comm, g_stream));
        if (cr > 0) SLLB_NCCL(ncclRecv(D->recvbuf.p + roff[r], (size_t)cr, ncclDouble, r, D->comm->comm, g_stream));
    }
    SLLB_NCCL(ncclGroupEnd());
    for (int r = 0; r < D->nranks; ++r) {
        if (roff[r + 1] - roff[r] <= 0) continue;
        Box4 b;
        for (int d = 0; d < 4; ++d) { b.lo[d] = rb[r * 8 + 2 * d] - dlo[d]; b.n[d] = rb[r * 8 + 2 * d + 1] - rb[r * 8 + 2 * d] + 1; }
        SLLB_CUDA(launch_unpack4d(dst->d, dn, b, D->recvbuf.p + roff[r], g_stream));
    }
    return SLLB_OK;
}
} // extern "C"


/* ------------------------------------------------------------------------------------------ */
/* 2D2V: sim_bsl_vp_2d2v_cart_poisson_serial                                                    */
/* simulations/parallel/bsl_vp_2d2v_cart_poisson_serial/sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:650-1364 */
/* ------------------------------------------------------------------------------------------ */
// 1: on one GPU the last stage of the density reductions (sum of the per-chunk / per-CTA partial sums) is folded into the
// first kernel of the direct Poisson solve instead of running as a kernel of its own (SLLB_FOLD_SUMS)
static int g_fold_sums = [] { const char *e = getenv("SLLB_FOLD_SUMS"); return (e && e[0] == '1') ? 1 : 0; }();
static int g_cuda_graphs = 1;   // sllb_set_cuda_graphs: replay one recorded time step as a CUDA graph (1D1V and 2D2V loops)
struct sllb_sim4d {
    sllb_sim4d_params_t p;
    sllb_comm *comm = nullptr;
    sllb_dist4d *D = nullptr;
    sllb_poisson *poisson = nullptr;
    double delta[4];
    int bx[8], bv[8];  // my boxes in the two layouts
    DevBuf rho_tile, rho_gather, rho_full, E1, E2, E1loc, E2loc, small, linesum;
    // splitting schedule (sll_t_splitting_coeff) and, for dim_split_V = 2, the modified-potential fields
    double steps[SLLB_SPLIT_MAX_STEPS];
    int nb_split_step = 3, dim_split_V = 1;
    bool begin_T = false;
    int stencil_r = -2, stencil_s = 2;
    DevBuf jacE, K1, K2, C1, C2, fdw;
    double jac_max = 0.0, nrj_jac = 0.0;
    // ensemble streaming (sllb_sim4d_stream_step): two more copies of f and two copy streams, so that the upload of the
    // next member and the download of the previous one overlap this member's step (PCIe is full duplex)
    DevBuf ring[2];
    cudaStream_t s_up = nullptr, s_down = nullptr;
    bool have_incoming = false, have_outgoing = false;
    int m[4], o[4];   // interpolation method / order per axis (advector_x1..x4, order_x1..x4 of the namelist)
    // where the charge density of the current f can be had without another sweep over f:
    // 0 nothing (reduce f), 1 rho_full already holds it (T stage plane kernel), 2 line sums of the last x4 pass
    // 3: the ranks' partial densities sit in the slots of the exchange buffer (summed inside the Poisson solve)
    int rho_state = 0;
    bool eloc_valid = false; // E1loc / E2loc were filled by the last field solve (tile extraction fused into it)
    // 4: the per-CTA partial densities of the T-stage plane kernel are still unsummed (single GPU: the Poisson solve sums them)
    const double *rho_parts = nullptr; int rho_nparts = 0;
    // dup_velocity_planes: the N4 + N3 + 1 side planes of the reference's (N+1)-point velocity axes, their velocities
    sllb_field *Fdup = nullptr;
    DevBuf vdup[2];
    bool dup_valid = false;  // the side planes differ from the cells they duplicate (a T stage has moved them apart)
    DevBuf vtab[2];          // velocities of my x3 / x4 indices in the x-sequential layout (constant: no launch per stage)
    DevBuf nrj_rows;         // per-row parts of the field energy written by the direct Poisson solve
    bool nrj_rows_valid = false;
    double nrj = 0.0;
    int istep = 0;
    int layout = 0;    // which copy of f is current: 0 x-sequential, 1 v-sequential
    // per-step diagnostics on the device (sllb_diag.cu): second-moment weights of the two velocity axes per layout,
    // the per-line moments the last x4 pass of a step leaves behind, the four global moments, the field energy, the rows
    DevBuf w2dev[2], dl1, dl2, dkin, dscratch, m4, nrjd, rows_dev;
    // one time step recorded as a CUDA graph (single GPU; the step is ~20 kernels, several of them a few microseconds
    // long: on grids that fit in L2 the gaps between launches are a fifth of the step).  Two variants: with / without the
    // diagnostics row.  A recording is replayed only from the same entry state it was recorded from.
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        int entry_rho_state = -1, entry_layout = -1; bool entry_line_diag = false;
        const double *fd = nullptr;   // the array of f the recording works on (ensemble streaming rotates three of them)
        int exit_rho_state = 0, exit_layout = 0; bool exit_line_diag = false, exit_eloc = false;
        long long launches = 0;
    } graph[2];
    cudaStream_t gstream = nullptr;
    int eager_steps = 0;          // steps run with plain launches since f was last touched from outside: a step is recorded
                                  // only after two of them, when every scratch buffer of the steady state exists
    DevBuf dstep;                 // device-side counters read by the row kernel: [0] row index in this run, [1] time step
    bool want_line_diag = false;  // set by the time loop for the last stage of a step when diagnostics are on
    bool line_diag_valid = false; // dl1/dl2/dkin/linesum describe the current f
    PhaseTimer timer;
    // local passes, rho+poisson, NCCL remap, diagnostics, fused V-stage pass, barrier after it, fused T-stage plane
    // kernel, all-reduce (rho + barrier) after it
    double phase_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

__global__ void k_landau4d(double *f, int n0, int n1, int n2, int n3, int lo2, int lo3, double x0min, double x1min,
                           double x2min, double x3min, double d0, double d1, double d2, double d3, double kx1, double kx2,
                           double eps) {
    const long long ntot = (long long)n0 * n1 * n2 * n3;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ntot; t += (long long)gridDim.x * blockDim.x) {
        long long r = t;
        const int i0 = (int)(r % n0); r /= n0;
        const int i1 = (int)(r % n1); r /= n1;
        const int i2 = (int)(r % n2); r /= n2;
        const int i3 = (int)r;
        const double x = x0min + (double)i0 * d0, y = x1min + (double)i1 * d1;
        const double vx = x2min + (double)(i2 + lo2) * d2, vy = x3min + (double)(i3 + lo3) * d3;
        const double factor1 = 1.0 + eps * cos(kx1 * x) * cos(kx2 * y);
        f[t] = (1.0 / (2.0 * 3.14159265358979323846)) * factor1 * exp(-0.5 * (vx * vx + vy * vy));
    }
}

// Bring f into the wanted layout.  The reference remaps back to the x-sequential layout after every V
// stage (:1171) and forth again before the next one (:1069); when two V stages follow each other (the end of
// one Strang step and the start of the next) that round trip moves every element back to where it was, so it
// is skipped here: same values, two remaps per step instead of four.
static int sim4d_to_layout(sllb_sim4d *S, int want) {
    if (S->layout == want) return SLLB_OK;
    SLLB_TRY(sllb_dist4d_remap(S->D, S->layout));
    S->layout = want;
    return SLLB_OK;
}

static int sim4d_fields(sllb_sim4d *S) {
    // rho(i1,i2) = delta3*delta4 * sum over (x3,x4) [trapezoid over the duplicated end points == plain sum],
    // computed in the v-sequential layout; tiles gathered to every rank (split_to_full, :1366-1401).
    sllb_dist4d *D = S->D;
    sllb_field *Fv = D->F[1];
    const int N1 = S->p.nc[0], N2 = S->p.nc[1];
    const int P = D->nranks;
    const long long n12 = (long long)N1 * N2;
    const double scale = S->delta[2] * S->delta[3];
    const long long tile = (long long)Fv->ext[0] * Fv->ext[1];
    const bool xchg = P > 1 && D->p2p && D->flag_barrier;   // density exchange by peer stores + flag barrier
    const double *rho_in = S->rho_full.p;   // what the Poisson solve reads: in_scale * sum of nslots arrays n12 apart
    int nslots = 1;
    double in_scale = 1.0;
    const bool direct = S->poisson->direct && g_poisson_direct;
    if (S->rho_state == 3) {
        // every rank's partial sum over ITS planes sits in slot [rank] of my exchange buffer (T stage plane kernel)
        rho_in = D->xbuf.p; nslots = P;
    } else if (S->rho_state == 4) {
        // the plane kernel's per-CTA partial sums, unscaled: summed (in order) and scaled inside the Poisson solve
        rho_in = S->rho_parts; nslots = S->rho_nparts; in_scale = scale;
    } else if (S->rho_state == 1) {
        // rho_full was accumulated by the plane kernel during the T stage (and all-reduced over the ranks)
    } else if (P == 1 && direct && g_fold_sums) {
        // one GPU: only the first stage of the velocity reduction runs as a kernel; its per-chunk partial sums are folded
        // (in order) into the first kernel of the Poisson solve
        const double *src = Fv->d;
        long long nv = (long long)Fv->ext[2] * Fv->ext[3];
        if (S->rho_state == 2) { src = S->linesum.p; nv = Fv->ext[2]; }   // sum over x4 came out of the last x4 pass
        SLLB_TRY(Fv->red_scratch.ensure(reduce_scratch_doubles(tile, nv)));
        SLLB_CUDA(launch_reduce_velocity_partials(src, tile, nv, Fv->red_scratch.p, &nslots, g_stream));
        rho_in = Fv->red_scratch.p; in_scale = scale;
    } else {
        double *rho_local = (P == 1) ? S->rho_full.p : S->rho_tile.p;
        if (S->rho_state == 2) {
            // sum over x4 came out of the last x4 pass; finish the sum over x3 (K3 on a [x1 x2][x3] array)
            SLLB_TRY(Fv->red_scratch.ensure(reduce_scratch_doubles(tile, Fv->ext[2])));
            SLLB_CUDA(launch_reduce_velocity(S->linesum.p, tile, Fv->ext[2], scale, rho_local, Fv->red_scratch.p, g_stream));
        } else {
            SLLB_TRY(sllb_reduce_velocity(Fv, 2, scale, rho_local));
        }
        if (P > 1 && xchg) {
            // my tile goes straight into slot 8 of every rank's exchange buffer; the flag barrier completes the gather
            PeerX px;
            for (int r = 0; r < 8; ++r) px.p[r] = r < P ? D->peer_x[r] : nullptr;
            k_bcast_tile<<<(unsigned)((tile + 255) / 256), 256, 0, g_stream>>>(S->rho_tile.p, Fv->ext[0], Fv->ext[1], S->bv[0], S->bv[2], N1,
                                                                        px, 8 * n12, P);
            count_launch();
            SLLB_CUDA(cudaGetLastError());
            SLLB_TRY(dist4d_barrier(D));
            rho_in = D->xbuf.p + 8 * n12;
        } else if (P > 1) {
            SLLB_TRY(sllb_comm_allgather(S->comm, S->rho_tile.p, S->rho_gather.p, tile));
            const int ext[4] = {N1, N2, 1, 1};
            for (int r = 0; r < P; ++r) {
                Box4 b;
                const int *bb = &D->boxes[1][r * 8];
                b.lo[0] = bb[0]; b.n[0] = bb[1] - bb[0] + 1; b.lo[1] = bb[2]; b.n[1] = bb[3] - bb[2] + 1;
                b.lo[2] = b.lo[3] = 0; b.n[2] = b.n[3] = 1;
                SLLB_CUDA(launch_unpack4d(S->rho_full.p, ext, b, S->rho_gather.p + (long long)r * tile, g_stream));
            }
        }
    }
    if (S->Fdup && S->dup_valid) {
        // trapezoid rule over the (N3+1)(N4+1) nodes = plain sum over the cells + the end-plane correction
        if (rho_in != S->rho_full.p || nslots != 1) return fail(SLLB_ERR_UNSUPPORTED, "sim4d: dup_velocity_planes needs the summed density (SLLB_FOLD_SUMS=0)");
        SLLB_CUDA(launch_dup_rho_corr(Fv->d, S->Fdup->d, n12, Fv->ext[2], Fv->ext[3], scale, S->rho_full.p, g_stream));
    }
    S->rho_state = 0;
    S->eloc_valid = false;
    S->nrj_rows_valid = false;
    if (direct) {
        // three kernels: sum of the slots fused into the first; extraction of my E tiles and the per-row parts of the
        // field energy fused into the last
        const int tbox[4] = {S->bv[0], Fv->ext[0], S->bv[2], Fv->ext[1]};
        SLLB_TRY(S->nrj_rows.ensure((size_t)N2));
        SLLB_CUDA(poisson2d_direct_solve(S->poisson->direct, rho_in, nslots, n12, in_scale, rho_in != S->rho_full.p ? S->rho_full.p : nullptr,
                                         0, nullptr, S->E1.p, S->E2.p, S->nrj_rows.p, P > 1 ? S->E1loc.p : nullptr,
                                         P > 1 ? S->E2loc.p : nullptr, P > 1 ? tbox : nullptr, g_stream));
        S->eloc_valid = P > 1;
        S->nrj_rows_valid = true;
    } else {
        if (nslots > 1) SLLB_CUDA(launch_sum_partials(rho_in, n12, nslots, in_scale, S->rho_full.p, g_stream));
        else if (rho_in != S->rho_full.p)
            SLLB_CUDA(cudaMemcpyAsync(S->rho_full.p, rho_in, (size_t)n12 * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
        SLLB_TRY(sllb_poisson_solve(S->poisson, S->rho_full.p, nullptr, S->E1.p, S->E2.p, nullptr));
    }
    if (S->dim_split_V == 2) {
        // field_x{1,2}(:,:,2) = E of the Poisson problem with jacobian_E as right-hand side (:1113-1126)
        SLLB_CUDA(launch_jacobian2d(S->E1.p, S->E2.p, N1, N2, S->stencil_r, S->stencil_s, S->fdw.p,
                                    4.0 / (S->delta[0] * S->delta[1]), S->jacE.p, g_stream));
        SLLB_TRY(sllb_poisson_solve(S->poisson, S->jacE.p, nullptr, S->K1.p, S->K2.p, nullptr));
    }
    return SLLB_OK;
}

// sum over the (N1+1)(N2+1) nodes INCLUDING the periodic duplicates of a(i,j)^2 + (squared ? b^2 : 2 b)
static double dup_sum(const std::vector<double> &a, const std::vector<double> &b, int N1, int N2, bool squared) {
    double s = 0;
    for (int j = 0; j <= N2; ++j) for (int i = 0; i <= N1; ++i) {
        const size_t k = (size_t)(i % N1) + (size_t)N1 * (j % N2);
        s += a[k] * a[k] + (squared ? b[k] * b[k] : b[k] + b[k]);
    }
    return s;
}
// thdiag columns 6 and 7: max|jacobian_E| (always computed by the reference, :1113) and nrj_jac (:1124-1126; the t = 0
// row sums field_x2 unsquared, :905)
static int sim4d_jac_diag(sllb_sim4d *S) {
    const int N1 = S->p.nc[0], N2 = S->p.nc[1];
    const size_t n12 = (size_t)N1 * N2;
    if (S->dim_split_V != 2)
        SLLB_CUDA(launch_jacobian2d(S->E1.p, S->E2.p, N1, N2, S->stencil_r, S->stencil_s, S->fdw.p,
                                    4.0 / (S->delta[0] * S->delta[1]), S->jacE.p, g_stream));
    std::vector<double> j(n12);
    SLLB_CUDA(cudaMemcpy(j.data(), S->jacE.p, n12 * 8, cudaMemcpyDeviceToHost));
    double m = 0;
    for (double v : j) if (fabs(v) > m) m = fabs(v);
    S->jac_max = m;
    S->nrj_jac = 0.0;
    if (S->dim_split_V == 2) {
        std::vector<double> k1(n12), k2(n12);
        SLLB_CUDA(cudaMemcpy(k1.data(), S->K1.p, n12 * 8, cudaMemcpyDeviceToHost));
        SLLB_CUDA(cudaMemcpy(k2.data(), S->K2.p, n12 * 8, cudaMemcpyDeviceToHost));
        S->nrj_jac = dup_sum(k1, k2, N1, N2, S->istep > 0) * S->delta[0] * S->delta[1];
    }
    return SLLB_OK;
}

// nrj = sum(E1^2 + E2^2) delta1 delta2 over the (N1+1)(N2+1) nodes INCLUDING the periodic duplicates (:1112)
static int sim4d_nrj(sllb_sim4d *S) {
    const int N1 = S->p.nc[0], N2 = S->p.nc[1];
    std::vector<double> e1((size_t)N1 * N2), e2((size_t)N1 * N2);
    SLLB_CUDA(cudaMemcpy(e1.data(), S->E1.p, e1.size() * 8, cudaMemcpyDeviceToHost));
    SLLB_CUDA(cudaMemcpy(e2.data(), S->E2.p, e2.size() * 8, cudaMemcpyDeviceToHost));
    double s = 0;
    for (int j = 0; j <= N2; ++j) for (int i = 0; i <= N1; ++i) {
        const size_t k = (size_t)(i % N1) + (size_t)N1 * (j % N2);
        s += e1[k] * e1[k] + e2[k] * e2[k];
    }
    S->nrj = s * S->delta[0] * S->delta[1];
    return SLLB_OK;
}

// second-moment weights of the local velocity indices of a layout; the trapezoid rule over the duplicated end points
// averages v^2 at both ends, which for the periodic pair (vmin, vmax) is 0.5 (vmin^2 + vmax^2)
static int sim4d_w2_tables(sllb_sim4d *S) {
    const sllb_sim4d_params_t &p = S->p;
    for (int lay = 0; lay < 2; ++lay) {
        sllb_field *F = S->D->F[lay];
        const int *bx = lay == 0 ? S->bx : S->bv;
        std::vector<double> w2;
        for (int a = 2; a < 4; ++a)
            for (int i = 0; i < F->ext[a]; ++i) {
                const int ig = i + bx[2 * a];
                const double v = p.xmin[a] + ig * S->delta[a];
                w2.push_back(ig == 0 ? 0.5 * (p.xmin[a] * p.xmin[a] + p.xmax[a] * p.xmax[a]) : v * v);
            }
        SLLB_TRY(S->w2dev[lay].ensure(w2.size()));
        SLLB_CUDA(cudaMemcpy(S->w2dev[lay].p, w2.data(), w2.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    return SLLB_OK;
}
// One diagnostics row (time, field energy, kinetic energy, mass, L1, L2: :1112,1183-1260) written to DEVICE memory: no
// host synchronisation, so a run with diagnostics every step keeps the GPU busy.
static int sim4d_diag_device(sllb_sim4d *S, double *d_row6) {
    sllb_field *F = S->D->F[S->layout];
    const sllb_sim4d_params_t &p = S->p;
    SLLB_TRY(S->m4.ensure(4));
    SLLB_TRY(S->nrjd.ensure(1));
    // field energy: the per-row parts the direct Poisson solve left behind, or one single-block reduction
    const double *nrj_parts = S->nrjd.p;
    int nn = 1;
    double nrj_scale = 1.0;
    if (S->nrj_rows_valid) { nrj_parts = S->nrj_rows.p; nn = p.nc[1]; nrj_scale = S->delta[0] * S->delta[1]; }
    else SLLB_CUDA(launch_dup_energy2d(S->E1.p, S->E2.p, p.nc[0], p.nc[1], S->delta[0] * S->delta[1], 1, S->nrjd.p, g_stream));
    const double *w3 = S->w2dev[S->layout].p, *w4 = w3 + F->ext[2];
    const double *m4p = S->m4.p;
    int nb = 1;
    if (S->line_diag_valid && S->layout == 1) {
        SLLB_TRY(S->dscratch.ensure(moments_from_lines_scratch()));
        SLLB_CUDA(launch_moments_from_lines_partials(S->linesum.p, S->dl1.p, S->dl2.p, S->dkin.p, (long long)F->ext[0] * F->ext[1],
                                                     F->ext[2], w3, S->dscratch.p, &nb, g_stream));
        m4p = S->dscratch.p;
    } else {
        const long long nx = (long long)F->ext[0] * F->ext[1], nv = (long long)F->ext[2] * F->ext[3];
        SLLB_TRY(F->rows.ensure((size_t)nv * 3));
        SLLB_CUDA(launch_row_sums(F->d, nx, nv, F->rows.p, g_stream));
        SLLB_CUDA(launch_moments_from_rows(F->rows.p, F->ext[2], F->ext[3], w3, w4, S->m4.p, g_stream));
    }
    // several ranks: the row holds MY part of the four integrals (time and field energy on rank 0 only); the rows of the
    // whole run are summed over the ranks by ONE all-reduce after the loop instead of one per step.
    // d_row6 == nullptr: the time loop's rows -- slot and time come from the device-side counters (graph replays)
    SLLB_CUDA(launch_sim4d_row(m4p, nb, nrj_parts, nn, nrj_scale, S->istep * p.dt, p.dt,
                               S->delta[0] * S->delta[1] * S->delta[2] * S->delta[3], S->D->rank == 0 ? 1 : 0, d_row6,
                               d_row6 ? nullptr : S->rows_dev.p, d_row6 ? nullptr : S->dstep.p, g_stream));
    return SLLB_OK;
}

// dup_velocity_planes: the side planes take part in a T stage with THEIR velocities (x3 = N3 / x4 = N4 sit at +v_max, the
// cells they duplicate at v_min = -v_max: :1037-1064 loops over all N+1 points of the velocity axes)
static int sim4d_T_side_planes(sllb_sim4d *S, double step) {
    sllb_field *Fx = S->D->F[0], *G = S->Fdup;
    const sllb_sim4d_params_t &p = S->p;
    const int M = G->ext[2];
    if (!S->dup_valid) SLLB_CUDA(launch_dup_fill(Fx->d, (long long)Fx->ext[0] * Fx->ext[1], Fx->ext[2], Fx->ext[3], G->d, g_stream));
    DispDesc d0, d1;
    d0.v = S->vdup[0].p; d0.scale = -step * p.dt / S->delta[0];
    d0.odiv = G->ext[1]; d0.omod = M; d0.ostr = 1; d0.idiv = d0.imod = 1; d0.istr = 0;
    d1.v = S->vdup[1].p; d1.scale = -step * p.dt / S->delta[1];
    d1.odiv = 1; d1.omod = M; d1.ostr = 1; d1.idiv = d1.imod = 1; d1.istr = 0;
    SLLB_TRY(advect_axis_dev(G, 0, S->m[0], S->o[0], d0));
    SLLB_TRY(advect_axis_dev(G, 1, S->m[1], S->o[1], d1));
    S->dup_valid = true;
    return SLLB_OK;
}
static int sim4d_T_cells(sllb_sim4d *S, double step, bool fuse);
// `fuse`: the stage's last pass writes straight into the other layout (advect + remap in one kernel)
static int sim4d_T(sllb_sim4d *S, double step, bool fuse) {
    if (S->Fdup) SLLB_TRY(sim4d_T_side_planes(S, step));   // before the cells move: a fresh copy is taken from them
    return sim4d_T_cells(S, step, fuse);
}
static int sim4d_T_cells(sllb_sim4d *S, double step, bool fuse) {
    sllb_field *Fx = S->D->F[0];
    const sllb_sim4d_params_t &p = S->p;
    S->rho_state = 0;
    S->line_diag_valid = false;
    // cubic splines: both passes and the charge density in one sweep (K1c); on several GPUs the same kernel also
    // stores into the v-sequential layout of the owning ranks when the next stage is a V stage (`fuse`)
    if (S->m[0] == SLLB_METHOD_SPLINE && S->m[1] == SLLB_METHOD_SPLINE && S->o[0] == 4 && S->o[1] == 4 && g_plane_kernel) {
        // displacement = -v * step * dt / delta_x with the velocity tables of my x3 / x4 indices (built once)
        DispDesc d0, d1;
        d0.v = S->vtab[0].p; d0.scale = -step * p.dt / S->delta[0];
        d0.odiv = Fx->ext[1]; d0.omod = Fx->ext[2]; d0.ostr = 1; d0.idiv = d0.imod = 1; d0.istr = 0;
        d1.v = S->vtab[1].p; d1.scale = -step * p.dt / S->delta[1];
        d1.odiv = Fx->ext[2]; d1.omod = Fx->ext[3]; d1.ostr = 1; d1.idiv = d1.imod = 1; d1.istr = 0;
        RemapDst rd;
        if (fuse) dist4d_remap_dst(S->D, 0, 1, &rd);
        S->timer.mark(0);
        sllb_dist4d *D = S->D;
        const long long n12 = (long long)p.nc[0] * p.nc[1];
        const bool xchg = D->nranks > 1 && D->p2p && D->flag_barrier;
        // several ranks: my partial density (the sum over MY planes) goes to slot [rank] of the exchange buffers
        double *rho_dst = xchg ? D->xbuf.p + (long long)D->rank * n12 : S->rho_full.p;
        // one GPU with the direct Poisson solve: the per-CTA partial densities stay unsummed, the solve folds them
        const bool leave_parts = D->nranks == 1 && S->poisson->direct && g_poisson_direct && g_fold_sums;
        int rc = advect_plane_dev(Fx, d0, d1, S->delta[2] * S->delta[3], rho_dst, fuse ? &rd : nullptr,
                                  leave_parts ? &S->rho_parts : nullptr, leave_parts ? &S->rho_nparts : nullptr);
        if (rc == SLLB_OK && leave_parts) { S->rho_state = 4; return SLLB_OK; }
        if (rc == SLLB_OK) {
            if (fuse) S->timer.mark(6);
            if (xchg) {
                PeerX px;
                for (int r = 0; r < 8; ++r) px.p[r] = r < D->nranks ? D->peer_x[r] : nullptr;
                k_bcast_slot<<<(unsigned)((n12 + 255) / 256), 256, 0, g_stream>>>(rho_dst, n12, px, (long long)D->rank * n12, D->nranks, D->rank);
                count_launch();
                SLLB_CUDA(cudaGetLastError());
                // ONE barrier covers both the remap stores of the plane kernel and the density slots
                SLLB_TRY(dist4d_barrier(D));
                S->rho_state = 3;
            } else {
                // rho_full holds the sum over MY planes: the all-reduce completes it and is the barrier of the remap
                if (D->nranks > 1) SLLB_TRY(sllb_comm_allreduce_sum(S->comm, S->rho_full.p, (int64_t)n12));
                S->rho_state = 1;
            }
            if (fuse) { S->timer.mark(7); S->layout = 1; }
            return SLLB_OK;
        }
        if (rc != SLLB_ERR_UNSUPPORTED) return rc;
    }
    // out(x) = in(x - v*step*dt): displacement in cells = -v*step*dt/delta_x  (:1037-1064)
    SLLB_TRY(sllb_advect_axis_affine(Fx, 0, S->m[0], S->o[0], 2, p.xmin[2] + S->bx[4] * S->delta[2], S->delta[2],
                                     -step * p.dt / S->delta[0]));
    DispDesc dd;
    SLLB_TRY(make_affine_disp(Fx, 1, 3, p.xmin[3] + S->bx[6] * S->delta[3], S->delta[3], -step * p.dt / S->delta[1], &dd));
    if (fuse) {
        SLLB_TRY(dist4d_advect_remap_dev(S->D, 0, 1, S->m[1], S->o[1], dd, &S->timer));
        S->layout = 1;
    } else {
        SLLB_TRY(advect_axis_dev(Fx, 1, S->m[1], S->o[1], dd));
    }
    return SLLB_OK;
}
// Chunked V stage on several GPUs.  The x4 + remap pass is bound by NVLink (its stores go to the peers), the x3 pass by
// HBM: the (x1,x2) tile of this rank is cut into chunks, the x3 pass of chunk c+1 runs on a high-priority stream while the
// x4 + remap pass of chunk c runs on a second one.  The two streams are ordinary (blocking) streams: what was launched
// before on the default stream (the field solve) precedes them, the flag barrier launched after them waits for both.
// Same kernels, same arithmetic per line: values are bit-identical to the two whole passes.
// SLLB_ERR_UNSUPPORTED (nothing launched) when the shape does not fit; the caller then runs the whole passes.
static int g_v_chunks = [] { const char *e = getenv("SLLB_V_CHUNKS"); const int v = e ? atoi(e) : 4; return v >= 2 && v <= 8 ? v : 4; }();
static int g_v_overlap = [] { const char *e = getenv("SLLB_V_OVERLAP"); return (e && e[0] == '1') ? 1 : 0; }();
static int sim4d_V_chunked(sllb_sim4d *S, const double *e1, const double *e2, double step) {
    sllb_dist4d *D = S->D;
    sllb_field *Fv = D->F[1];
    const sllb_sim4d_params_t &p = S->p;
    if (!(S->m[2] == SLLB_METHOD_SPLINE && S->o[2] == 4 && S->m[3] == SLLB_METHOD_SPLINE && S->o[3] == 4)) return SLLB_ERR_UNSUPPORTED;
    if (S->timer.on) return SLLB_ERR_UNSUPPORTED;   // the phase timers time whole passes on the default stream
    const long long n12l = (long long)Fv->ext[0] * Fv->ext[1];
    int nch = g_v_chunks;
    while (nch > 1 && (n12l % (32LL * nch) != 0)) --nch;
    if (nch < 2) return SLLB_ERR_UNSUPPORTED;
    for (int a = 2; a < 4; ++a) {   // the conditions under which launch_advect takes the chunked strided kernel
        const int n = Fv->ext[a];
        if (n % 2 != 0 || n / 2 < 32 || (size_t)n * 32 * 8 + 128 + 8 * 32 * 8 > 227 * 1024 || g_spline_split >= 0) return SLLB_ERR_UNSUPPORTED;
    }
    if (!D->s_local) {
        int lo = 0, hi = 0;
        SLLB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        SLLB_CUDA(cudaStreamCreateWithPriority(&D->s_local, cudaStreamDefault, hi));
        SLLB_CUDA(cudaStreamCreateWithPriority(&D->s_remap, cudaStreamDefault, lo));
        for (cudaEvent_t &e : D->ev_chunk) SLLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    DispDesc d3, d4;
    SLLB_TRY(make_field_disp(Fv, 2, e1, 2, -step * p.dt / S->delta[2], &d3));
    SLLB_TRY(make_field_disp(Fv, 3, e2, 2, -step * p.dt / S->delta[3], &d4));
    RemapDst rd;
    dist4d_remap_dst(D, 1, 3, &rd);
    const long long cnt = n12l / nch;
    const cudaStream_t keep = g_stream;
    int rc = SLLB_OK;
    for (int c = 0; c < nch && !rc; ++c) {
        // x3 pass (outer = x4 planes, inner = the (x1,x2) tile): lines (x4 = q, in = i0 + r)
        LineSub s3 = {cnt, c * cnt, 0, cnt * Fv->ext[3], 1};
        g_stream = D->s_local;
        rc = advect_axis_dev(Fv, 2, S->m[2], S->o[2], d3, nullptr, nullptr, nullptr, &s3);
        if (!rc) rc = check_cuda(cudaEventRecord(D->ev_chunk[c], D->s_local), "event record");
        if (!rc) rc = check_cuda(cudaStreamWaitEvent(D->s_remap, D->ev_chunk[c], 0), "stream wait");
        // x4 + remap pass (outer = 1, inner = tile * N3): lines in = i0 + r + n12l * x3
        LineSub s4 = {cnt, c * cnt, n12l, cnt * Fv->ext[2], 0};
        g_stream = D->s_remap;
        if (!rc) rc = advect_axis_dev(Fv, 3, S->m[3], S->o[3], d4, &rd, nullptr, nullptr, &s4);
    }
    g_stream = keep;
    if (rc) return fail(SLLB_ERR_CUDA, "sim4d: chunked V stage failed after its first launch: " + std::string(sllb_last_error()));
    SLLB_TRY(dist4d_barrier(D));
    return SLLB_OK;
}
static int sim4d_V(sllb_sim4d *S, double step, double step2, bool fuse) {
    sllb_field *Fv = S->D->F[1];
    S->line_diag_valid = false;
    S->dup_valid = false;   // out(N+1) = out(1): every V pass rewrites the duplicated planes from the cells
    const sllb_sim4d_params_t &p = S->p;
    const double *e1 = S->E1.p, *e2 = S->E2.p;
    if (S->dim_split_V == 2) {
        // alpha = field(:,:,1) step(k+1) + field(:,:,2) step(k+2) (:1137-1141,1153-1157): one combined field, unit step
        const long long n12 = (long long)p.nc[0] * p.nc[1];
        SLLB_CUDA(launch_lincomb2(S->E1.p, S->K1.p, step, step2, n12, S->C1.p, g_stream));
        SLLB_CUDA(launch_lincomb2(S->E2.p, S->K2.p, step, step2, n12, S->C2.p, g_stream));
        e1 = S->C1.p; e2 = S->C2.p;
        step = 1.0;
    }
    if (S->D->nranks > 1 && S->eloc_valid && S->dim_split_V != 2) {
        e1 = S->E1loc.p; e2 = S->E2loc.p;   // my tiles came out of the field solve
    } else if (S->D->nranks > 1) { // my (x1,x2) tile of the replicated field
        const int ext[4] = {p.nc[0], p.nc[1], 1, 1};
        Box4 b;
        b.lo[0] = S->bv[0]; b.n[0] = S->bv[1] - S->bv[0] + 1; b.lo[1] = S->bv[2]; b.n[1] = S->bv[3] - S->bv[2] + 1;
        b.lo[2] = b.lo[3] = 0; b.n[2] = b.n[3] = 1;
        SLLB_CUDA(launch_pack4d(e1, ext, b, S->E1loc.p, g_stream));
        SLLB_CUDA(launch_pack4d(e2, ext, b, S->E2loc.p, g_stream));
        e1 = S->E1loc.p; e2 = S->E2loc.p;
    }
    // out(v) = in(v - E*step*dt)  (:1137-1166), displacement computed from E inside the kernel (K5)
    S->rho_state = 0;
    if (fuse && g_v_overlap) {
        const int rcc = sim4d_V_chunked(S, e1, e2, step);
        if (rcc == SLLB_OK) { S->layout = 0; return SLLB_OK; }
        if (rcc != SLLB_ERR_UNSUPPORTED) return rcc;
    }
    SLLB_TRY(sllb_advect_axis_field(Fv, 2, S->m[2], S->o[2], e1, 2, -step * p.dt / S->delta[2]));
    DispDesc dd;
    SLLB_TRY(make_field_disp(Fv, 3, e2, 2, -step * p.dt / S->delta[3], &dd));
    if (fuse) {
        SLLB_TRY(dist4d_advect_remap_dev(S->D, 1, 3, S->m[3], S->o[3], dd, &S->timer));
        S->layout = 0;
    } else {
        // f stays in this layout, so the next thing that happens to it is another V stage: hand its charge
        // density over as line sums of this pass (sum over x4 per (x1,x2,x3) line) instead of re-reading f
        int rc = SLLB_ERR_UNSUPPORTED;
        if (S->m[3] == SLLB_METHOD_SPLINE && S->o[3] == 4 && g_plane_kernel) {
            const size_t nl = (size_t)Fv->ext[0] * Fv->ext[1] * Fv->ext[2];
            rc = S->linesum.ensure(nl);
            if (!rc && S->want_line_diag) {
                // the moments of the diagnostics row come out of this pass too (no extra sweep over f)
                rc = S->dl1.ensure(nl);
                if (!rc) rc = S->dl2.ensure(nl);
                if (!rc) rc = S->dkin.ensure(nl);
                if (!rc) {
                    LineDiag dg = {S->dl1.p, S->dl2.p, S->dkin.p, S->w2dev[1].p + Fv->ext[2]};
                    rc = advect_axis_dev(Fv, 3, S->m[3], S->o[3], dd, nullptr, S->linesum.p, &dg);
                    if (rc == SLLB_OK) S->line_diag_valid = true;
                }
            }
            if (!rc && !S->line_diag_valid) rc = advect_axis_dev(Fv, 3, S->m[3], S->o[3], dd, nullptr, S->linesum.p);
            if (rc == SLLB_OK) S->rho_state = 2;
        }
        if (rc == SLLB_ERR_UNSUPPORTED) rc = advect_axis_dev(Fv, 3, S->m[3], S->o[3], dd);
        if (rc) return rc;
    }
    return SLLB_OK;
}

extern "C" {

int sllb_sim4d_create(const sllb_sim4d_params_t *p, sllb_comm_t comm, sllb_sim4d_t *Sout) {
    if (!p || !Sout) return fail(SLLB_ERR_INVALID, "sim4d_create: null");
    SLLB_TRY(require_device());
    sllb_sim4d *S = new sllb_sim4d();
    S->p = *p;
    S->comm = comm;
    {
        int bt = 0;
        int rcs = sllb_splitting_coeff(p->split, p->dt, S->steps, nullptr, &S->nb_split_step, &bt, &S->dim_split_V);
        if (rcs) { delete S; return rcs; }
        S->begin_T = bt != 0;
        for (int d = 0; d < 4; ++d) {
            const bool per_axis = p->order_axis[d] != 0;
            S->m[d] = per_axis ? p->method_axis[d] : p->method;
            S->o[d] = per_axis ? p->order_axis[d] : p->order;
        }
        if (p->stencil_r != 0 || p->stencil_s != 0) { S->stencil_r = p->stencil_r; S->stencil_s = p->stencil_s; }
        if (S->stencil_r >= 0 || S->stencil_s <= 0 || S->stencil_s - S->stencil_r > 16) { delete S; return fail(SLLB_ERR_INVALID, "sim4d_create: need stencil_r < 0 < stencil_s"); }
    }
    for (int d = 0; d < 4; ++d) S->delta[d] = (p->xmax[d] - p->xmin[d]) / (double)p->nc[d];
    int rc = sllb_dist4d_create(comm, p->nc, &S->D);
    if (!rc) rc = sllb_poisson2d_create(p->nc[0], p->nc[1], p->xmin[0], p->xmax[0], p->xmin[1], p->xmax[1], &S->poisson);
    if (rc) { sllb_sim4d_destroy(S); return rc; }
    sllb_dist4d_box(S->D, 0, S->bx);
    sllb_dist4d_box(S->D, 1, S->bv);
    rc = sim4d_w2_tables(S);
    for (int a = 0; a < 2 && !rc; ++a) {
        const int n = S->D->F[0]->ext[2 + a];
        rc = S->vtab[a].ensure((size_t)n);
        if (!rc) rc = check_cuda(launch_affine(S->vtab[a].p, n, p->xmin[2 + a] + S->bx[4 + 2 * a] * S->delta[2 + a], S->delta[2 + a], g_stream), "k_affine");
    }
    if (!rc && p->dup_velocity_planes) {
        // a step must end with a V stage (it re-synchronises the duplicated planes before the diagnostics are taken)
        const bool ends_with_T = (S->begin_T && S->nb_split_step % 2 == 1) || (!S->begin_T && S->nb_split_step % 2 == 0);
        if (S->D->nranks != 1) rc = fail(SLLB_ERR_UNSUPPORTED, "sim4d_create: dup_velocity_planes is a single-GPU mode");
        else if (ends_with_T) rc = fail(SLLB_ERR_UNSUPPORTED, "sim4d_create: dup_velocity_planes needs a splitting scheme whose step ends with a V stage");
        if (!rc) {
            const int n3 = p->nc[2], n4 = p->nc[3], M = n3 + n4 + 1;
            const int ext3[3] = {p->nc[0], p->nc[1], M};
            rc = field_alloc(3, ext3, &S->Fdup);
            std::vector<double> v3(M), v4(M);
            for (int m = 0; m < M; ++m) {
                // side plane m: (x3 = N3, x4 = m) | (x3 = m - N4, x4 = N4) | the corner; the duplicated node sits at v_max
                v3[m] = (m < n4 || m == n3 + n4) ? p->xmax[2] : p->xmin[2] + (m - n4) * S->delta[2];
                v4[m] = (m < n4) ? p->xmin[3] + m * S->delta[3] : p->xmax[3];
            }
            for (int a = 0; a < 2 && !rc; ++a) {
                rc = S->vdup[a].ensure((size_t)M);
                if (!rc) rc = check_cuda(cudaMemcpy(S->vdup[a].p, a == 0 ? v3.data() : v4.data(), sizeof(double) * M, cudaMemcpyHostToDevice), "dup velocities");
            }
        }
    }
    if (rc) { sllb_sim4d_destroy(S); return rc; }
    const size_t n12 = (size_t)p->nc[0] * p->nc[1];
    sllb_field *Fv = S->D->F[1];
    const size_t tile = (size_t)Fv->ext[0] * Fv->ext[1];
    rc = S->rho_full.ensure(n12);
    if (!rc) rc = S->E1.ensure(n12);
    if (!rc) rc = S->E2.ensure(n12);
    if (!rc) rc = S->rho_tile.ensure(tile);
    if (!rc) rc = S->rho_gather.ensure(tile * S->D->nranks);
    if (!rc) rc = S->E1loc.ensure(tile);
    if (!rc) rc = S->E2loc.ensure(tile);
    if (!rc) rc = S->small.ensure(16);
    if (!rc) rc = S->jacE.ensure(n12);
    if (!rc) rc = S->fdw.ensure(32);
    if (!rc && S->dim_split_V == 2) {
        rc = S->K1.ensure(n12);
        if (!rc) rc = S->K2.ensure(n12);
        if (!rc) rc = S->C1.ensure(n12);
        if (!rc) rc = S->C2.ensure(n12);
    }
    if (!rc) {
        double w[32];
        rc = sllb_compute_w_hermite(S->stencil_r, S->stencil_s, w);
        if (!rc) rc = check_cuda(cudaMemcpy(S->fdw.p, w, sizeof(double) * (S->stencil_s - S->stencil_r + 1), cudaMemcpyHostToDevice), "fd weights");
    }
    if (rc) { sllb_sim4d_destroy(S); return rc; }
    // initial data in the x-sequential layout (sll_f_landau_mode_initializer_4d)
    sllb_field *Fx = S->D->F[0];
    k_landau4d<<<148 * 8, 256>>>(Fx->d, Fx->ext[0], Fx->ext[1], Fx->ext[2], Fx->ext[3], S->bx[4], S->bx[6], p->xmin[0],
                                 p->xmin[1], p->xmin[2], p->xmin[3], S->delta[0], S->delta[1], S->delta[2], S->delta[3],
                                 p->kx1, p->kx2, p->eps);
    rc = check_cuda(cudaGetLastError(), "k_landau4d");
    // E at t = 0 (:851-887)
    if (!rc) rc = sim4d_to_layout(S, 1);
    if (!rc) rc = sim4d_fields(S);
    if (!rc) rc = sim4d_nrj(S);
    if (rc) { sllb_sim4d_destroy(S); return rc; }
    *Sout = S;
    return SLLB_OK;
}
int sllb_sim4d_destroy(sllb_sim4d_t S) {
    if (!S) return SLLB_OK;
    sllb_field_destroy(S->Fdup);
    for (auto &g : S->graph) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (S->gstream) cudaStreamDestroy(S->gstream);
    if (S->s_up) cudaStreamDestroy(S->s_up);
    if (S->s_down) cudaStreamDestroy(S->s_down);
    sllb_poisson_destroy(S->poisson);
    sllb_dist4d_destroy(S->D);
    delete S;
    return SLLB_OK;
}
int sllb_sim4d_field(sllb_sim4d_t S, sllb_field_t *F) {
    if (!S || !F) return fail(SLLB_ERR_INVALID, "sim4d_field: null");
    S->rho_state = 0; // the caller may overwrite f through the handle
    S->line_diag_valid = false;
    S->eager_steps = 0;
    SLLB_TRY(sim4d_to_layout(S, 0));
    *F = S->D->F[0];
    return SLLB_OK;
}
/* Ensemble streaming on one GPU: every call (1) starts the download of the member stepped by the PREVIOUS call to
 * host_prev_out, (2) starts the upload of host_next_in (both pinned host arrays of the periodic cells, either may be
 * NULL), (3) advances the member uploaded by the previous call by one time step while those copies run, (4) waits for
 * all three and rotates the three device copies of f.  N members take N + 2 calls: the first call only uploads, the
 * last only downloads.  Every member is an independent state: its fields are recomputed from its f. */
int sllb_sim4d_stream_step(sllb_sim4d_t S, const double *host_next_in, double *host_prev_out) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim4d_stream_step: null");
    if (S->D->nranks != 1) return fail(SLLB_ERR_UNSUPPORTED, "sim4d_stream_step: single GPU only");
    sllb_field *F = S->D->F[0];
    const size_t n = (size_t)F->total;
    if (!S->s_up) {
        SLLB_TRY(S->ring[0].ensure(n));
        SLLB_TRY(S->ring[1].ensure(n));
        SLLB_CUDA(cudaStreamCreateWithFlags(&S->s_up, cudaStreamNonBlocking));
        SLLB_CUDA(cudaStreamCreateWithFlags(&S->s_down, cudaStreamNonBlocking));
    }
    auto swap_cur = [&](DevBuf &other) { // exchange the compute copy with a ring copy (both layouts alias it on one GPU)
        double *t = F->d; F->d = other.p; other.p = t;
        S->D->F[1]->d = F->d;
    };
    DevBuf &rin = S->ring[0], &rout = S->ring[1];
    SLLB_CUDA(cudaDeviceSynchronize());
    if (S->have_outgoing && host_prev_out)
        SLLB_CUDA(cudaMemcpyAsync(host_prev_out, rout.p, n * sizeof(double), cudaMemcpyDeviceToHost, S->s_down));
    const bool step_now = S->have_incoming;
    if (step_now) swap_cur(rin); // the member uploaded by the previous call becomes the compute copy; rin is free again
    if (host_next_in)
        SLLB_CUDA(cudaMemcpyAsync(rin.p, host_next_in, n * sizeof(double), cudaMemcpyHostToDevice, S->s_up));
    int rc = SLLB_OK;
    if (step_now) {
        S->rho_state = 0; S->layout = 0; // a fresh, independent state: fields are recomputed from this f
        S->line_diag_valid = false;
        S->eager_steps = 0;              // every call works on another of the three arrays: nothing to record
        rc = sllb_sim4d_run(S, 1, 0, nullptr);
    }
    SLLB_CUDA(cudaStreamSynchronize(S->s_up));
    SLLB_CUDA(cudaStreamSynchronize(S->s_down));
    if (rc) return rc;
    if (step_now) swap_cur(rout); // the stepped member waits in rout for the next call's download
    S->have_outgoing = step_now;
    S->have_incoming = host_next_in != nullptr;
    return SLLB_OK;
}
/* Position-weighted checksums of the global f, (sum w f, sum w f^2) with w a function of the GLOBAL index of every
 * point: the same numbers (to rounding) on any number of ranks and in either layout, unlike the mass they change when an
 * element lands in the wrong place. */
int sllb_sim4d_checksum(sllb_sim4d_t S, double out[2]) {
    if (!S || !out) return fail(SLLB_ERR_INVALID, "sim4d_checksum: null");
    sllb_field *F = S->D->F[S->layout];
    const int *b = S->layout == 0 ? S->bx : S->bv;
    const int lo[4] = {b[0], b[2], b[4], b[6]};
    SLLB_TRY(S->dscratch.ensure(moments_from_lines_scratch()));
    SLLB_TRY(S->m4.ensure(4));
    SLLB_CUDA(launch_checksum4d(F->d, F->ext, lo, S->dscratch.p, S->m4.p, g_stream));
    if (S->D->nranks > 1) SLLB_TRY(sllb_comm_allreduce_sum(S->comm, S->m4.p, 2));
    SLLB_CUDA(cudaMemcpy(out, S->m4.p, 2 * sizeof(double), cudaMemcpyDeviceToHost));
    return SLLB_OK;
}
/* rho, E1, E2 of the last field solve (N1 x N2 periodic cells each, replicated on every rank); any pointer may be NULL */
int sllb_sim4d_fields_host(sllb_sim4d_t S, double *rho, double *e1, double *e2) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim4d_fields_host: null");
    const size_t bytes = (size_t)S->p.nc[0] * S->p.nc[1] * sizeof(double);
    if (rho) SLLB_CUDA(cudaMemcpy(rho, S->rho_full.p, bytes, cudaMemcpyDeviceToHost));
    if (e1) SLLB_CUDA(cudaMemcpy(e1, S->E1.p, bytes, cudaMemcpyDeviceToHost));
    if (e2) SLLB_CUDA(cudaMemcpy(e2, S->E2.p, bytes, cudaMemcpyDeviceToHost));
    return SLLB_OK;
}
int sllb_sim4d_box(sllb_sim4d_t S, int which, int box[8]) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim4d_box: null");
    return sllb_dist4d_box(S->D, which, box);
}
int sllb_sim4d_diagnostics(sllb_sim4d_t S, double *row6) {
    if (!S || !row6) return fail(SLLB_ERR_INVALID, "sim4d_diagnostics: null");
    sllb_field *Fx = S->D->F[S->layout];
    const int *bx = S->layout == 0 ? S->bx : S->bv;
    const sllb_sim4d_params_t &p = S->p;
    // second velocity moments; the trapezoid rule over the duplicated end points averages v^2 at both
    // ends, which for the periodic pair (vmin, vmax) is 0.5 (vmin^2 + vmax^2)
    std::vector<double> w2;
    for (int a = 2; a < 4; ++a)
        for (int i = 0; i < Fx->ext[a]; ++i) {
            const int ig = i + bx[2 * a];
            double v = p.xmin[a] + ig * S->delta[a];
            double vv = v * v;
            if (ig == 0) vv = 0.5 * (p.xmin[a] * p.xmin[a] + p.xmax[a] * p.xmax[a]);
            w2.push_back(vv);
        }
    double m[7];
    SLLB_TRY(moments_local(Fx, 2, nullptr, w2.data(), m));
    double loc[4] = {m[0], m[1], m[2], 0.5 * (m[5] + m[6])};
    if (S->D->nranks > 1) {
        SLLB_CUDA(cudaMemcpy(S->small.p, loc, sizeof(loc), cudaMemcpyHostToDevice));
        SLLB_TRY(sllb_comm_allreduce_sum(S->comm, S->small.p, 4));
        SLLB_CUDA(cudaMemcpy(loc, S->small.p, sizeof(loc), cudaMemcpyDeviceToHost));
    }
    const double vol = S->delta[0] * S->delta[1] * S->delta[2] * S->delta[3];
    row6[0] = S->istep * p.dt;
    row6[1] = S->nrj;
    row6[2] = loc[3] * vol;
    row6[3] = loc[0] * vol; row6[4] = loc[1] * vol; row6[5] = loc[2] * vol;
    return SLLB_OK;
}
int sllb_sim4d_thdiag(sllb_sim4d_t S, double *row13) {
    if (!S || !row13) return fail(SLLB_ERR_INVALID, "sim4d_thdiag: null");
    const sllb_sim4d_params_t &p = S->p;
    double r6[6];
    SLLB_TRY(sllb_sim4d_diagnostics(S, r6));
    SLLB_TRY(sim4d_nrj(S));
    SLLB_TRY(sim4d_jac_diag(S));
    const double pi = 3.14159265358979323846;
    const double mass0 = (p.xmax[0] - p.xmin[0]) * (p.xmax[1] - p.xmin[1]); // analytic (:913-918)
    const double nrj0 = (0.5 * p.eps * pi) * (0.5 * p.eps * pi) / (p.kx1 * p.kx2) * (1.0 / (p.kx1 * p.kx1) + 1.0 / (p.kx2 * p.kx2)); // :449-450
    double l20 = (2.0 * pi / p.kx1) * (2.0 * pi / p.kx2) * 0.25;                                                                       // :451-452
    l20 = l20 * (1.0 + 0.25 * (p.eps * p.eps)) / pi;
    const bool t0 = S->istep == 0;
    row13[0] = S->istep * p.dt; row13[1] = S->nrj; row13[2] = t0 ? mass0 : r6[2]; row13[3] = nrj0; row13[4] = mass0;
    row13[5] = S->jac_max; row13[6] = S->nrj_jac;
    row13[7] = t0 ? mass0 : r6[3]; row13[8] = t0 ? mass0 : r6[4]; row13[9] = t0 ? l20 : r6[5];
    row13[10] = mass0; row13[11] = mass0; row13[12] = l20;
    return SLLB_OK;
}
// one time step: the stages of the splitting scheme + the diagnostics row (all launches go to g_stream)
static int sim4d_one_step(sllb_sim4d *S, bool with_diagnostics, bool last_step, bool can_fuse) {
    const double *steps = S->steps;
    const int nsub = S->nb_split_step, dimV = S->dim_split_V;
    const bool beginT = S->begin_T;
    int isub = 0; bool T = beginT;
    for (int ss = 0; ss < nsub; ++ss) {
        // the stage after this one (possibly the first stage of the next step) is of the other kind
        // <=> f is needed in the other layout next: fuse the remap into this stage's last pass
        const bool last_stage = (last_step && ss == nsub - 1);
        const bool nextT = (ss == nsub - 1) ? beginT : !T;
        const bool fuse = can_fuse && !last_stage && (nextT != T);
        S->want_line_diag = with_diagnostics && ss == nsub - 1;
        if (T) {
            isub += 1;
            SLLB_TRY(sim4d_to_layout(S, 0));
            S->timer.mark(2);
            SLLB_TRY(sim4d_T(S, steps[isub - 1], fuse));
            S->timer.mark(0);
        } else {
            SLLB_TRY(sim4d_to_layout(S, 1));
            S->timer.mark(2);
            SLLB_TRY(sim4d_fields(S));
            S->timer.mark(1);
            SLLB_TRY(sim4d_V(S, steps[isub], dimV == 2 ? steps[isub + 1] : 0.0, fuse));
            S->timer.mark(0);
            isub += dimV;
        }
        T = !T;
    }
    S->want_line_diag = false;
    if (with_diagnostics) {
        SLLB_TRY(sim4d_diag_device(S, nullptr));
        S->timer.mark(3);
    }
    return SLLB_OK;
}
int sllb_sim4d_run(sllb_sim4d_t S, int nsteps, int with_diagnostics, double *rows) {
    if (!S || nsteps < 0) return fail(SLLB_ERR_INVALID, "sim4d_run: bad arguments");
    S->timer.begin(g_phase_timers != 0);
    S->timer.mark(-1);
    const bool can_fuse = S->D->p2p && g_fused_remap && S->D->nranks > 1;
    const bool diag = with_diagnostics != 0;
    if (diag && nsteps > 0) {
        // the recorded steps carry the address of the row buffer: if it has to grow, the recordings go
        const double *before = S->rows_dev.p;
        SLLB_TRY(S->rows_dev.ensure((size_t)6 * (nsteps > 4096 ? nsteps : 4096)));
        if (before && before != S->rows_dev.p)
            for (auto &g : S->graph) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    }
    SLLB_TRY(S->dstep.ensure(2));
    {
        const double init[2] = {0.0, (double)S->istep};
        SLLB_CUDA(cudaMemcpy(S->dstep.p, init, sizeof(init), cudaMemcpyHostToDevice));
    }
    // graphs: one GPU, no phase timers (their events would be recorded into the graph)
    const bool graphs = g_cuda_graphs && S->D->nranks == 1 && !S->timer.on;
    sllb_sim4d::StepGraph &G = S->graph[diag ? 1 : 0];
    bool on_gstream = false;   // replays run on their own stream: the default stream must be idle when they start
    for (int it = 0; it < nsteps; ++it) {
        const bool matches = G.exec && G.fd == S->D->F[0]->d && G.entry_rho_state == S->rho_state && G.entry_layout == S->layout &&
                             G.entry_line_diag == S->line_diag_valid && !S->dup_valid;
        if (graphs && !matches && S->eager_steps >= 2 && !G.exec) {
            // every buffer the steady-state step needs exists after two plain steps (no allocation may happen while
            // recording): record the next one (it is executed by the launch below, not while recording)
            if (!S->gstream) SLLB_CUDA(cudaStreamCreateWithFlags(&S->gstream, cudaStreamNonBlocking));
            SLLB_CUDA(cudaDeviceSynchronize());
            cudaGraph_t graph = nullptr;
            const long long l0 = launch_count();
            G.entry_rho_state = S->rho_state; G.entry_layout = S->layout; G.entry_line_diag = S->line_diag_valid;
            G.fd = S->D->F[0]->d;
            const int keep_rho = S->rho_state, keep_layout = S->layout; const bool keep_ld = S->line_diag_valid, keep_el = S->eloc_valid;
            g_stream = S->gstream;
            cudaError_t ce = cudaStreamBeginCapture(S->gstream, cudaStreamCaptureModeThreadLocal);
            int rc = ce == cudaSuccess ? sim4d_one_step(S, diag, false, can_fuse) : SLLB_ERR_CUDA;
            cudaError_t ee = cudaStreamEndCapture(S->gstream, &graph);
            g_stream = 0;
            G.launches = launch_count() - l0;
            count_launches(-G.launches); // recorded, not executed
            G.exit_rho_state = S->rho_state; G.exit_layout = S->layout; G.exit_line_diag = S->line_diag_valid; G.exit_eloc = S->eloc_valid;
            S->rho_state = keep_rho; S->layout = keep_layout; S->line_diag_valid = keep_ld; S->eloc_valid = keep_el;
            if (rc == SLLB_OK && ee == cudaSuccess && graph) ee = cudaGraphInstantiate(&G.exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (rc != SLLB_OK || ce != cudaSuccess || ee != cudaSuccess) { // fall back to plain launches, loudly recorded
                cudaGetLastError();
                G.exec = nullptr;
                g_cuda_graphs = 0;
                set_error("sim4d_run: CUDA graph capture failed, continuing with stream launches");
            }
        }
        const bool replay = graphs && G.exec && G.fd == S->D->F[0]->d && G.entry_rho_state == S->rho_state &&
                            G.entry_layout == S->layout && G.entry_line_diag == S->line_diag_valid && !S->dup_valid;
        if (replay) {
            if (!on_gstream) { SLLB_CUDA(cudaDeviceSynchronize()); on_gstream = true; }
            SLLB_CUDA(cudaGraphLaunch(G.exec, S->gstream));
            count_launches(G.launches);
            S->rho_state = G.exit_rho_state; S->layout = G.exit_layout; S->line_diag_valid = G.exit_line_diag; S->eloc_valid = G.exit_eloc;
        } else {
            if (on_gstream) { SLLB_CUDA(cudaStreamSynchronize(S->gstream)); on_gstream = false; }
            SLLB_TRY(sim4d_one_step(S, diag, it == nsteps - 1, can_fuse));
            S->eager_steps += 1;
        }
        S->istep += 1;
    }
    SLLB_CUDA(cudaDeviceSynchronize());
    if (S->D->nranks > 1 && S->D->flag_barrier) {
        double err = 0.0;
        SLLB_CUDA(cudaMemcpy(&err, S->D->errflag.p, sizeof(double), cudaMemcpyDeviceToHost));
        if (err != 0.0) return fail(SLLB_ERR_CUDA, "sim4d_run: a rank did not reach the flag barrier within the time-out");
    }
    if (diag && nsteps > 0 && S->D->nranks > 1) SLLB_TRY(sllb_comm_allreduce_sum(S->comm, S->rows_dev.p, (int64_t)6 * nsteps));
    if (diag && nsteps > 0) {
        std::vector<double> h((size_t)6 * nsteps);
        SLLB_CUDA(cudaMemcpy(h.data(), S->rows_dev.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
        S->nrj = h[6 * (size_t)(nsteps - 1) + 1];
        if (rows) memcpy(rows, h.data(), h.size() * sizeof(double));
    }
    S->timer.collect(S->phase_ms);
    return SLLB_OK;
}
/* 1: on several GPUs the V stage that ends with a remap runs chunked, the x3 pass of one chunk under the NVLink-bound
 * x4 + remap pass of the previous one; 0 (default): two whole passes.  Same values.  Measured on 2 GPUs (128^4): 2.76 ms per
 * step whole, 2.80 / 2.86 / 2.89 ms with 2 / 4 / 8 chunks -- the two kernels compete for the same SM slots and the remap
 * pass loses more store bandwidth than the x3 pass hides (profiles/r02_v_overlap_n2_s10.json), hence opt-in. */
int sllb_set_v_overlap(int on) {
    g_v_overlap = on ? 1 : 0;
    if (on >= 2 && on <= 8) g_v_chunks = on;   // 2..8: that many chunks
    return SLLB_OK;
}
/* 1: sllb_sim4d_run records a CUDA event per phase (sllb_sim4d_phase_ms*); 0 (default): no events on the hot path */
int sllb_set_phase_timers(int on) {
    g_phase_timers = on ? 1 : 0;
    return SLLB_OK;
}
int sllb_sim4d_phase_ms6(sllb_sim4d_t S, double out[6]) {
    if (!S || !out) return fail(SLLB_ERR_INVALID, "sim4d_phase_ms6: null");
    for (int k = 0; k < 6; ++k) out[k] = S->phase_ms[k];
    out[4] += S->phase_ms[6]; out[5] += S->phase_ms[7];
    return SLLB_OK;
}
int sllb_sim4d_phase_ms8(sllb_sim4d_t S, double out[8]) {
    if (!S || !out) return fail(SLLB_ERR_INVALID, "sim4d_phase_ms8: null");
    for (int k = 0; k < 8; ++k) out[k] = S->phase_ms[k];
    return SLLB_OK;
}
int sllb_sim4d_phase_ms(sllb_sim4d_t S, double out[4]) {
    if (!S || !out) return fail(SLLB_ERR_INVALID, "sim4d_phase_ms: null");
    for (int k = 0; k < 4; ++k) out[k] = S->phase_ms[k];
    out[0] += S->phase_ms[4] + S->phase_ms[6]; out[2] += S->phase_ms[5] + S->phase_ms[7];
    return SLLB_OK;
}

} // extern "C"

/* ------------------------------------------------------------------------------------------ */
/* 1D1V: sim_bsl_vp_1d1v_cart (no drive), Strang VTV                                            */
/* simulations/parallel/bsl_vp_1d1v_cart/sll_m_sim_bsl_vp_1d1v_cart.F90:1066-1907               */
/* ------------------------------------------------------------------------------------------ */
struct sllb_sim2d {
    int nc[2]; double xmin[2], xmax[2], delta[2];
    int init; double kmode, eps, dt; int method, order;
    int m[2], o[2];                       // advector_x1 / advector_x2 of the namelist (method, order per axis)
    double steps[SLLB_SPLIT_MAX_STEPS];   // split_step of sll_t_splitting_coeff (dim_split_V = 1 cases)
    int nsub = 3; bool beginT = false;
    double time_init = 0.0;
    DevBuf modes_part, modes_out;
    sllb_field *F = nullptr;
    sllb_poisson *poisson = nullptr;
    DevBuf rho, E;
    int istep = 0;
    // one time step recorded as a CUDA graph (the 1D1V problems are a few MB: the step is bound by its ~10 kernel
    // launches, not by memory traffic)
    cudaStream_t gstream = nullptr;
    cudaGraphExec_t gexec = nullptr;
    long long glaunches = 0;   // kernels in the recorded step (launch bookkeeping of the replays)
};
__global__ void k_init2d(double *f, int n0, int n1, double x0min, double x1min, double d0, double d1, int init,
                         double kmode, double eps) {
    const long long ntot = (long long)n0 * n1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ntot; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % n0), j = (int)(t / n0);
        const double x = x0min + i * d0, v = x1min + j * d1;
        const double fac = 1.0 / sqrt(2.0 * 3.14159265358979323846);
        // sll_f_landau_initializer_2d, sll_f_two_stream_instability_initializer_2d, sll_f_bump_on_tail_initializer_2d
        // (sll_m_common_array_initializers.F90:507-595)
        if (init == 0) f[t] = fac * (1.0 + eps * cos(kmode * x)) * exp(-0.5 * v * v);
        else if (init == 1) f[t] = fac * (1.0 + eps * cos(kmode * x)) * v * v * exp(-0.5 * v * v);
        else f[t] = fac * (1.0 + eps * cos(kmode * x)) * (0.9 * exp(-0.5 * v * v) + 0.2 * exp(-0.5 * ((v - 4.5) * (v - 4.5)) / (0.5 * 0.5)));
    }
}
static int sim2d_field(sllb_sim2d *S) {
    // rho = 1 - sum_j f w_j, trapezoid weights over the duplicated v end points == delta_v * plain sum (:937-943,1590-1604)
    SLLB_TRY(sllb_reduce_velocity(S->F, 1, S->delta[1], S->rho.p));
    SLLB_CUDA(launch_rho_1d1v(S->rho.p, S->nc[0], 1.0, 1.0, g_stream));
    SLLB_TRY(sllb_poisson_solve(S->poisson, S->rho.p, nullptr, S->E.p, nullptr, nullptr));
    return SLLB_OK;
}
extern "C" {
int sllb_sim2d_create(int nc_x1, int nc_x2, double x1_min, double x1_max, double x2_min, double x2_max, int init,
                      double kmode, double eps, double dt, int method, int order, sllb_sim2d_t *Sout) {
    if (!Sout || nc_x1 < 8 || nc_x2 < 8) return fail(SLLB_ERR_INVALID, "sim2d_create: bad arguments");
    SLLB_TRY(require_device());
    sllb_sim2d *S = new sllb_sim2d();
    S->nc[0] = nc_x1; S->nc[1] = nc_x2; S->xmin[0] = x1_min; S->xmax[0] = x1_max; S->xmin[1] = x2_min; S->xmax[1] = x2_max;
    S->delta[0] = (x1_max - x1_min) / nc_x1; S->delta[1] = (x2_max - x2_min) / nc_x2;
    S->init = init; S->kmode = kmode; S->eps = eps; S->dt = dt; S->method = method; S->order = order;
    if (init < 0 || init > 2) { delete S; return fail(SLLB_ERR_UNSUPPORTED, "sim2d_create: initial function 0 Landau, 1 two-stream, 2 bump-on-tail"); }
    S->m[0] = S->m[1] = method; S->o[0] = S->o[1] = order;
    S->steps[0] = 0.5; S->steps[1] = 1.0; S->steps[2] = 0.5;   // SLL_STRANG_VTV
    int rc = field_alloc(2, S->nc, &S->F);
    if (!rc) rc = sllb_poisson1d_create(nc_x1, x1_min, x1_max, &S->poisson);
    if (!rc) rc = S->rho.ensure(nc_x1);
    if (!rc) rc = S->E.ensure(nc_x1);
    if (!rc) {
        k_init2d<<<148, 256>>>(S->F->d, nc_x1, nc_x2, x1_min, x2_min, S->delta[0], S->delta[1], init, kmode, eps);
        rc = check_cuda(cudaGetLastError(), "k_init2d");
    }
    if (!rc) rc = sim2d_field(S);
    if (rc) { sllb_sim2d_destroy(S); return rc; }
    *Sout = S;
    return SLLB_OK;
}
int sllb_sim2d_destroy(sllb_sim2d_t S) {
    if (!S) return SLLB_OK;
    if (S->gexec) cudaGraphExecDestroy(S->gexec);
    if (S->gstream) cudaStreamDestroy(S->gstream);
    sllb_poisson_destroy(S->poisson);
    sllb_field_destroy(S->F);
    delete S;
    return SLLB_OK;
}
int sllb_sim2d_field(sllb_sim2d_t S, sllb_field_t *F) {
    if (!S || !F) return fail(SLLB_ERR_INVALID, "sim2d_field: null");
    *F = S->F;
    return SLLB_OK;
}
int sllb_sim2d_run(sllb_sim2d_t S, int nsteps, double *rows) {
    if (!S || nsteps < 0) return fail(SLLB_ERR_INVALID, "sim2d_run: bad arguments");
    const double *steps = S->steps;
    // velocity weights for the moments: trapezoid over duplicated end points (see sim4d diagnostics)
    std::vector<double> w1(S->nc[1]), w2(S->nc[1]);
    for (int j = 0; j < S->nc[1]; ++j) {
        const double v = S->xmin[1] + j * S->delta[1];
        w1[j] = v; w2[j] = v * v;
    }
    w1[0] = 0.5 * (S->xmin[1] + S->xmax[1]);
    w2[0] = 0.5 * (S->xmin[1] * S->xmin[1] + S->xmax[1] * S->xmax[1]);
    std::vector<double> hE(S->nc[0]);
    if (S->gexec) SLLB_CUDA(cudaDeviceSynchronize()); // the replays run on their own stream: nothing of the caller's may be in flight
    auto one_step = [&]() -> int {
        bool T = S->beginT;
        for (int ss = 0; ss < S->nsub; ++ss) {
            if (T) { // out(x) = in(x - v*step*dt)  (:1570-1585)
                SLLB_TRY(sllb_advect_axis_affine(S->F, 0, S->m[0], S->o[0], 1, S->xmin[1], S->delta[1],
                                                 -steps[ss] * S->dt / S->delta[0]));
                SLLB_TRY(sim2d_field(S));
            } else { // alpha = -E*step, out(v) = in(v + E*step*dt)  (:1656-1686)
                SLLB_TRY(sllb_advect_axis_field(S->F, 1, S->m[1], S->o[1], S->E.p, 1, steps[ss] * S->dt / S->delta[1]));
            }
            T = !T;
        }
        return SLLB_OK;
    };
    for (int it = 0; it < nsteps; ++it) {
        if (g_cuda_graphs && !S->gexec && S->istep >= 1) {
            // every buffer the step needs exists after the first step: record the next one (it is executed by the
            // launch below, not while recording)
            if (!S->gstream) SLLB_CUDA(cudaStreamCreateWithFlags(&S->gstream, cudaStreamNonBlocking));
            SLLB_CUDA(cudaDeviceSynchronize());
            cudaGraph_t graph = nullptr;
            const long long l0 = launch_count();
            g_stream = S->gstream;
            cudaError_t ce = cudaStreamBeginCapture(S->gstream, cudaStreamCaptureModeThreadLocal);
            int rc = ce == cudaSuccess ? one_step() : SLLB_ERR_CUDA;
            cudaError_t ee = cudaStreamEndCapture(S->gstream, &graph);
            g_stream = 0;
            S->glaunches = launch_count() - l0;
            count_launches(-S->glaunches); // recorded, not executed
            if (rc == SLLB_OK && ee == cudaSuccess && graph) ee = cudaGraphInstantiate(&S->gexec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (rc != SLLB_OK || ce != cudaSuccess || ee != cudaSuccess) { // fall back to plain launches, loudly recorded
                cudaGetLastError();
                S->gexec = nullptr;
                g_cuda_graphs = 0;
                set_error("sim2d_run: CUDA graph capture failed, continuing with stream launches");
            }
        }
        if (S->gexec) {
            SLLB_CUDA(cudaGraphLaunch(S->gexec, S->gstream));
            count_launches(S->glaunches);
            if (rows) SLLB_CUDA(cudaStreamSynchronize(S->gstream)); // the diagnostics below read f and E on the default stream
        } else {
            SLLB_TRY(one_step());
        }
        S->istep += 1;
        if (rows) {
            double m[5];
            SLLB_TRY(moments_local(S->F, 1, w1.data(), w2.data(), m));
            SLLB_CUDA(cudaMemcpy(hE.data(), S->E.p, hE.size() * 8, cudaMemcpyDeviceToHost));
            double epot = 0;
            for (int i = 0; i < S->nc[0]; ++i) epot += hE[i] * hE[i];
            epot = 0.5 * epot * S->delta[0];
            const double dv = S->delta[1], dx = S->delta[0];
            double *r = rows + 8 * it;
            r[0] = S->time_init + S->istep * S->dt; r[1] = m[0] * dv * dx; r[2] = m[1] * dv * dx; r[3] = m[3] * dv * dx;
            r[4] = m[2] * dv * dx; r[5] = 0.5 * m[4] * dv * dx; r[6] = epot; r[7] = r[5] + r[6];
        }
    }
    SLLB_CUDA(cudaDeviceSynchronize());
    return SLLB_OK;
}
static void sim2d_drop_graph(sllb_sim2d *S) {
    if (S->gexec) { cudaDeviceSynchronize(); cudaGraphExecDestroy(S->gexec); S->gexec = nullptr; }
}
/* split_case of the namelist (sll_m_sim_bsl_vp_1d1v_cart.F90:816-846): the 1D1V loop walks split_step(1..nb_split_step)
 * alternating T and V (:1462-1696), so the schemes whose V stage takes two coefficients (the *VP* cases) do not apply */
int sllb_sim2d_set_splitting(sllb_sim2d_t S, int split_case) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim2d_set_splitting: null");
    double st[SLLB_SPLIT_MAX_STEPS];
    int nsub = 0, bt = 0, dimv = 1;
    SLLB_TRY(sllb_splitting_coeff(split_case, S->dt, st, nullptr, &nsub, &bt, &dimv));
    if (dimv != 1) return fail(SLLB_ERR_UNSUPPORTED, "sim2d_set_splitting: splitting schemes with dim_split_V = 2 belong to the 2D2V simulation");
    sim2d_drop_graph(S);
    memcpy(S->steps, st, sizeof(st));
    S->nsub = nsub; S->beginT = bt != 0;
    return SLLB_OK;
}
int sllb_sim2d_set_advectors(sllb_sim2d_t S, int method_x1, int order_x1, int method_x2, int order_x2) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim2d_set_advectors: null");
    sim2d_drop_graph(S);
    S->m[0] = method_x1; S->o[0] = order_x1; S->m[1] = method_x2; S->o[1] = order_x2;
    return SLLB_OK;
}
/* lim = {x1_min, x1_max, x2_min, x2_max} */
int sllb_sim2d_geometry(sllb_sim2d_t S, double lim[4]) {
    if (!S || !lim) return fail(SLLB_ERR_INVALID, "sim2d_geometry: null");
    lim[0] = S->xmin[0]; lim[1] = S->xmax[0]; lim[2] = S->xmin[1]; lim[3] = S->xmax[1];
    return SLLB_OK;
}
int sllb_sim2d_set_time(sllb_sim2d_t S, double time_init) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim2d_set_time: null");
    S->time_init = time_init - S->istep * S->dt;
    return SLLB_OK;
}
/* rho and E of the current state (N1 periodic cells each; NULL = skip) */
int sllb_sim2d_fields_host(sllb_sim2d_t S, double *rho, double *efield) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim2d_fields_host: null");
    SLLB_CUDA(cudaDeviceSynchronize());
    if (rho) SLLB_CUDA(cudaMemcpy(rho, S->rho.p, (size_t)S->nc[0] * sizeof(double), cudaMemcpyDeviceToHost));
    if (efield) SLLB_CUDA(cudaMemcpy(efield, S->E.p, (size_t)S->nc[0] * sizeof(double), cudaMemcpyDeviceToHost));
    return SLLB_OK;
}
/* One row of the reference's thdiag.dat (:1703-1801): time, mass, l1norm, momentum, l2norm, kinetic_energy,
 * potential_energy, their sum, then Re/Im of rho^_k for k = 0..nb_mode and f_hat_x2(k) = sum_v w_v |f^_k(v)|^2 for
 * k = 0..nb_mode, with the normalised transform (1/N) sum_j u_j e^{-2 pi i j k / N}: 8 + 3 (nb_mode + 1) numbers. */
int sllb_sim2d_thdiag(sllb_sim2d_t S, int nb_mode, double *row) {
    if (!S || !row || nb_mode < 0) return fail(SLLB_ERR_INVALID, "sim2d_thdiag: bad arguments (nb_mode >= 0, :811-814)");
    const int N1 = S->nc[0], N2 = S->nc[1], nm = nb_mode + 1;
    SLLB_CUDA(cudaDeviceSynchronize());
    std::vector<double> w1(N2), w2(N2);
    for (int j = 0; j < N2; ++j) { const double v = S->xmin[1] + j * S->delta[1]; w1[j] = v; w2[j] = v * v; }
    w1[0] = 0.5 * (S->xmin[1] + S->xmax[1]);
    w2[0] = 0.5 * (S->xmin[1] * S->xmin[1] + S->xmax[1] * S->xmax[1]);
    double m[5];
    SLLB_TRY(moments_local(S->F, 1, w1.data(), w2.data(), m));
    std::vector<double> hE(N1), hr(N1);
    SLLB_TRY(sllb_sim2d_fields_host(S, hr.data(), hE.data()));
    double epot = 0;
    for (int i = 0; i < N1; ++i) epot += hE[i] * hE[i];
    epot = 0.5 * epot * S->delta[0];
    const double dv = S->delta[1], dx = S->delta[0];
    row[0] = S->time_init + S->istep * S->dt; row[1] = m[0] * dv * dx; row[2] = m[1] * dv * dx; row[3] = m[3] * dv * dx;
    row[4] = m[2] * dv * dx; row[5] = 0.5 * m[4] * dv * dx; row[6] = epot; row[7] = row[5] + row[6];
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < nm; ++k) {
        double re = 0, im = 0;
        for (int j = 0; j < N1; ++j) {
            const double a = 2.0 * pi * (double)(((long long)j * k) % N1) / (double)N1;
            re += hr[j] * cos(a); im -= hr[j] * sin(a);
        }
        // sll_f_fft_get_mode_r2c_1d: purely real at k = 0 and k = N/2 (interfaces/fft/sll_m_fft_fftw3.F90:320-323)
        if (k % N1 == 0 || 2 * (k % N1) == N1) im = 0.0;
        row[8 + 2 * k] = re / N1; row[9 + 2 * k] = im / N1;
    }
    SLLB_TRY(S->modes_part.ensure((size_t)N2 * nm));
    SLLB_TRY(S->modes_out.ensure((size_t)nm));
    SLLB_CUDA(launch_row_modes(S->F->d, N1, N2, nm, dv, S->modes_part.p, S->modes_out.p, g_stream));
    SLLB_CUDA(cudaMemcpy(row + 8 + 2 * nm, S->modes_out.p, (size_t)nm * sizeof(double), cudaMemcpyDeviceToHost));
    return SLLB_OK;
}
/* <restart_file>_proc_0000.rst of the reference (:1281-1307 read, :1762-1770 write): a raw stream of the time followed by
 * f_x1 with the duplicated end points, (N1+1) x (N2+1) doubles, column-major */
int sllb_sim2d_write_restart(sllb_sim2d_t S, const char *path) {
    if (!S || !path) return fail(SLLB_ERR_INVALID, "sim2d_write_restart: null");
    const int N1 = S->nc[0], N2 = S->nc[1];
    std::vector<double> f((size_t)(N1 + 1) * (N2 + 1));
    const int dup[2] = {1, 1};
    SLLB_TRY(sllb_field_download(S->F, f.data(), dup));
    FILE *fp = fopen(path, "wb");
    if (!fp) return fail(SLLB_ERR_INVALID, std::string("sim2d_write_restart: cannot create ") + path);
    const double t = S->time_init + S->istep * S->dt;
    const bool ok = fwrite(&t, sizeof(double), 1, fp) == 1 && fwrite(f.data(), sizeof(double), f.size(), fp) == f.size();
    fclose(fp);
    return ok ? SLLB_OK : fail(SLLB_ERR_INVALID, "sim2d_write_restart: short write");
}
int sllb_sim2d_read_restart(sllb_sim2d_t S, const char *path, double *time) {
    if (!S || !path) return fail(SLLB_ERR_INVALID, "sim2d_read_restart: null");
    const int N1 = S->nc[0], N2 = S->nc[1];
    std::vector<double> f((size_t)(N1 + 1) * (N2 + 1));
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail(SLLB_ERR_INVALID, std::string("#file ") + path + " does not exist");
    double t = 0.0;
    const bool ok = fread(&t, sizeof(double), 1, fp) == 1 && fread(f.data(), sizeof(double), f.size(), fp) == f.size();
    fclose(fp);
    if (!ok) return fail(SLLB_ERR_INVALID, std::string("sim2d_read_restart: ") + path + " is shorter than 1 + (N1+1)(N2+1) doubles");
    const int dup[2] = {1, 1};
    SLLB_TRY(sllb_field_upload(S->F, f.data(), dup));
    if (time) *time = t;
    SLLB_TRY(sim2d_field(S));   // E of the restored state
    return SLLB_OK;
}
/* 1 (default): sllb_sim2d_run replays one recorded time step as a CUDA graph; 0: one launch per kernel */
int sllb_set_cuda_graphs(int on) {
    g_cuda_graphs = on ? 1 : 0;
    return SLLB_OK;
}
} // extern "C"

