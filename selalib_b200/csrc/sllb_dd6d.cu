// sllb_dd6d.cu -- 6D slim domain decomposition, halo exchange (a12) and the 3D3V simulation on 1..P GPUs.
//
// Reference: sll_t_cartesian_topology_6d / sll_t_decomposition_slim_6d and
// sll_s_apply_halo_exchange_slim_6d_real64 (src/parallelization/decomposition/sll_m_decomposition.F90:124-141,
// 209-227,379-555,835-869,1715-2030), the fixed-stencil whole-array advector
// (src/semi_lagrangian/advection/sll_m_advection_6d_lagrange_dd_slim.F90:806-2001) and the time loop of
// simulations/parallel/bsl_vp_3d3v_cart_dd/sll_m_sim_bsl_vp_3d3v_cart_dd_slim.F90:278-960.
// MPI_Sendrecv with the ring neighbours becomes a grouped ncclSend/ncclRecv pair per side over NVLink.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sllb_internal.h"

using namespace sllb;

struct sllb_dd6d {
    sllb_comm *comm = nullptr;
    int nranks = 1, rank = 0;
    int global[6], procs[6], coords[6];
    int mn[6], nw[6];          // global 0-based offset and width of the local block
    int left[6], right[6];     // ring neighbours per axis (ranks)
    sllb_field *F = nullptr;
    DevBuf halo_l, halo_r, send_lo, send_hi;
    double *cur_l = nullptr, *cur_r = nullptr; // halo planes of the last exchange
    int hw_l = 0, hw_r = 0, halo_axis = -1;
    // peer path: double-buffered halo arrays of every rank mapped into this process; the pack kernel stores the
    // edge planes straight into the neighbour's halo buffer over NVLink (no send buffer, no NCCL copy)
    bool p2p = false;
    DevBuf pbuf[8];                       // [parity*2 + side], side 0 = left halo, 1 = right halo; 4 + same: the
                                          // local spline's boundary sums per line (bc_left, bc_right)
    size_t bcap = 0;                      // capacity of each boundary-sum buffer in doubles
    DevBuf bc_l, bc_r, bc_send_r, bc_send_l; // NCCL / single-rank path of the boundary sums
    size_t pcap = 0;                      // capacity of each in doubles
    std::vector<void *> peers, ipc_opened;
    DevBuf flag;
    // barrier through flags in peer-mapped memory instead of an all-reduce (one launch of a few microseconds)
    DevBuf sigbuf, errflag;
    unsigned long long *peer_sig[8] = {};
    unsigned long long epoch = 0;
    bool flag_barrier = false;
    int parity = 0;
    double exch_ms = 0.0;      // device time of the last halo exchange (pack + send/recv)
    // pipelined split-axis pass: the lines are cut into chunks, the exchange of chunk c+1 (peer stores + barrier, on
    // s_comm) overlaps the stencil kernel of chunk c (on s_comp)
    cudaStream_t s_comm = nullptr, s_comp = nullptr;
    cudaEvent_t ev_start = nullptr, ev_end = nullptr, ev_comm0 = nullptr, ev_comm1 = nullptr, ev_chunk[16] = {};
    bool exch_pending = false; // exch_ms of the last pipelined pass still to be read from ev_comm0/1
    int pre_axis = -1, pre_h = 0, pre_chunks = 0;   // a pipelined exchange issued ahead of its pass (dd6d_halo_prefetch)
    cudaEvent_t ev_sten[16] = {};                   // stencil kernel of chunk c of the last pipelined pass has finished
    int sten_axis = -1, sten_chunks = 0;            // ... which pass that was
    cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr; // around the last plain exchange (read lazily by sllb_dd6d_exchange_ms)
    bool xch_pending = false;
};
static int g_halo_chunks = -1; // -1: SLLB_HALO_CHUNKS or 4; 1 = exchange everything, then one kernel

static int g_force_halo = 0; // 1: take the halo-exchange + halo-cells kernel path even when procs(axis) == 1
static int g_halo_p2p = 1;   // 1: peer stores when available, 0: pack + ncclSend/ncclRecv
static const int HALO_P2P_HW_MAX = 5; // widest halo (stencil 11) the peer buffers are sized for

namespace {
// MPI_Cart_create ordering (row-major: the LAST dimension varies fastest), sll_m_decomposition.F90:379-555
int cart_rank(const int procs[6], const int c[6]) {
    int r = 0;
    for (int d = 0; d < 6; ++d) r = r * procs[d] + c[d];
    return r;
}
void cart_coords(const int procs[6], int rank, int c[6]) {
    for (int d = 5; d >= 0; --d) { c[d] = rank % procs[d]; rank /= procs[d]; }
}
long long outer_of(const sllb_dd6d *D, int axis) { long long o = 1; for (int d = axis + 1; d < 6; ++d) o *= D->nw[d]; return o; }
long long inner_of(const sllb_dd6d *D, int axis) { long long i = 1; for (int d = 0; d < axis; ++d) i *= D->nw[d]; return i; }
} // namespace

// Edge planes j0 .. j0+hw-1 of `axis` (lines of the sub-box `box`, all lines when NULL) into a halo buffer laid out
// [outer][hw][inner], possibly in a peer's memory.  Default: the COPY ENGINES (cudaMemcpy2DAsync, device to device over
// NVLink): no SM is spent on the transfer, so the stencil kernel of the previous chunk keeps the whole GPU, and the DMA
// path moves 750+ GB/s where the capped pack kernel running beside the stencil kernel reached ~510.  SLLB_HALO_DMA=0: the
// pack kernel (K7).
static int g_halo_dma = [] { const char *e = getenv("SLLB_HALO_DMA"); return (e && e[0] == '0') ? 0 : 1; }();
static int halo_copy(sllb_dd6d *D, int axis, int j0, int hw, double *dst, cudaStream_t st, const LineBox *box, int max_blocks) {
    if (hw <= 0) return SLLB_OK;
    const int n = D->nw[axis];
    long long outer = 1, inner = 1;
    for (int d = axis + 1; d < 6; ++d) outer *= D->nw[d];
    for (int d = 0; d < axis; ++d) inner *= D->nw[d];
    LineBox b;
    if (box) b = *box;
    else { b.o0 = 0; b.ocount = outer; b.i0 = 0; b.icount = inner; }
    const bool whole_rows = (b.i0 == 0 && b.icount == inner);
    if (!g_halo_dma || (!whole_rows && b.ocount != 1)) {
        SLLB_CUDA(launch_halo_pack(D->F->d, outer, n, inner, j0, hw, dst, st, box, max_blocks));
        return SLLB_OK;
    }
    const double *src = D->F->d;
    if (whole_rows) {
        // per o one contiguous block of hw * inner doubles: rows of a 2D copy
        SLLB_CUDA(cudaMemcpy2DAsync(dst + (size_t)b.o0 * hw * inner, (size_t)hw * inner * sizeof(double),
                                    src + ((size_t)b.o0 * n + j0) * inner, (size_t)n * inner * sizeof(double),
                                    (size_t)hw * inner * sizeof(double), (size_t)b.ocount, cudaMemcpyDeviceToDevice, st));
    } else {
        // one o, a range of `in`: the hw planes are the rows
        SLLB_CUDA(cudaMemcpy2DAsync(dst + (size_t)b.o0 * hw * inner + b.i0, (size_t)inner * sizeof(double),
                                    src + ((size_t)b.o0 * n + j0) * inner + b.i0, (size_t)inner * sizeof(double),
                                    (size_t)b.icount * sizeof(double), (size_t)hw, cudaMemcpyDeviceToDevice, st));
    }
    return SLLB_OK;
}
// cross-rank barrier after peer stores: flags in peer-mapped memory, or the all-reduce it replaces
static int dd6d_barrier(sllb_dd6d *D, cudaStream_t st) {
    if (D->flag_barrier) {
        D->epoch += 1;
        return check_cuda(launch_flag_barrier(D->peer_sig, D->nranks, D->rank, D->epoch, D->errflag.p, st), "k_flag_barrier");
    }
    SLLB_NCCL(ncclAllReduce(D->flag.p, D->flag.p, 1, ncclDouble, ncclSum, D->comm->comm, st));
    return SLLB_OK;
}

extern "C" {

/* host-only: the decomposition of `rank` (no device needed) */
int sllb_dd6d_plan(int nranks, int rank, const int global[6], const int procs_in[6], int procs[6], int coords[6], int mn[6],
                   int nw[6], int left[6], int right[6]) {
    if (!global || nranks < 1 || rank < 0 || rank >= nranks) return fail(SLLB_ERR_INVALID, "dd6d_plan: bad arguments");
    int pr[6], co[6];
    bool given = false;
    if (procs_in) for (int d = 0; d < 6; ++d) if (procs_in[d] > 0) given = true;
    if (given) {
        long long prod = 1;
        for (int d = 0; d < 6; ++d) { pr[d] = procs_in[d] > 0 ? procs_in[d] : 1; prod *= pr[d]; }
        if (prod != nranks) return fail(SLLB_ERR_INVALID, "dd6d: process grid does not match the number of ranks");
    } else {
        SLLB_TRY(sllb_set_process_grid(nranks, pr));
    }
    cart_coords(pr, rank, co);
    for (int d = 0; d < 6; ++d) {
        // slim decomposition requires n % procs == 0 (sll_m_decomposition.F90:846)
        if (global[d] < 1 || global[d] % pr[d] != 0)
            return fail(SLLB_ERR_INVALID, "dd6d: number of cells must be divisible by the number of processes along every axis");
        int cl[6], cr[6];
        memcpy(cl, co, sizeof(cl)); memcpy(cr, co, sizeof(cr));
        cl[d] = (co[d] + pr[d] - 1) % pr[d];
        cr[d] = (co[d] + 1) % pr[d];
        if (procs) procs[d] = pr[d];
        if (coords) coords[d] = co[d];
        if (nw) nw[d] = global[d] / pr[d];
        if (mn) mn[d] = co[d] * (global[d] / pr[d]);
        if (left) left[d] = cart_rank(pr, cl);
        if (right) right[d] = cart_rank(pr, cr);
    }
    return SLLB_OK;
}

int sllb_dd6d_create(sllb_comm_t c, const int global[6], const int procs_in[6], sllb_dd6d_t *Dout) {
    if (!global || !Dout) return fail(SLLB_ERR_INVALID, "dd6d_create: null");
    SLLB_TRY(require_device());
    sllb_dd6d *D = new sllb_dd6d();
    D->comm = c;
    D->nranks = c ? c->nranks : 1;
    D->rank = c ? c->rank : 0;
    for (int d = 0; d < 6; ++d) D->global[d] = global[d];
    int rc = sllb_dd6d_plan(D->nranks, D->rank, global, procs_in, D->procs, D->coords, D->mn, D->nw, D->left, D->right);
    if (!rc) rc = field_alloc(6, D->nw, &D->F);
    if (!rc && D->nranks > 1 && D->nranks <= 8) {
        const char *env = getenv("SLLB_HALO_P2P");
        if (!(env && env[0] == '0')) {
            // largest halo over the split axes: HW_MAX planes
            size_t cap = 0;
            for (int d = 1; d < 6; ++d)
                if (D->procs[d] > 1) {
                    const int hw = HALO_P2P_HW_MAX < D->nw[d] ? HALO_P2P_HW_MAX : D->nw[d];
                    const size_t c = (size_t)(D->F->total / D->nw[d]) * hw;
                    if (c > cap) cap = c;
                }
            if (cap > 0) {
                size_t bcap = 0; // one value per line of the longest split axis
                for (int d = 1; d < 6; ++d)
                    if (D->procs[d] > 1 && (size_t)(D->F->total / D->nw[d]) > bcap) bcap = (size_t)(D->F->total / D->nw[d]);
                void *mine[9];
                for (int k = 0; k < 4 && !rc; ++k) { rc = D->pbuf[k].ensure(cap); mine[k] = D->pbuf[k].p; }
                for (int k = 4; k < 8 && !rc; ++k) { rc = D->pbuf[k].ensure(bcap); mine[k] = D->pbuf[k].p; }
                if (!rc) rc = D->flag.ensure(2);
                if (!rc) rc = D->sigbuf.ensure(16);
                if (!rc) rc = D->errflag.ensure(1);
                if (!rc) rc = check_cuda(cudaMemset(D->sigbuf.p, 0, 16 * sizeof(double)), "memset");
                if (!rc) rc = check_cuda(cudaMemset(D->errflag.p, 0, sizeof(double)), "memset");
                mine[8] = D->sigbuf.p;
                bool ok = false;
                std::vector<void *> all;
                if (!rc) rc = peer_map_buffers(D->comm, mine, 9, all, D->ipc_opened, &ok);
                if (!rc && ok) {
                    // the halo / boundary-sum buffers keep their [rank][8] indexing; the ninth buffer is the flag array
                    D->peers.assign((size_t)D->nranks * 8, nullptr);
                    for (int r = 0; r < D->nranks; ++r) {
                        for (int k = 0; k < 8; ++k) D->peers[(size_t)r * 8 + k] = all[(size_t)r * 9 + k];
                        D->peer_sig[r] = static_cast<unsigned long long *>(all[(size_t)r * 9 + 8]);
                    }
                    const char *eb = getenv("SLLB_FLAG_BARRIER");
                    D->flag_barrier = !(eb && eb[0] == '0');
                }
                D->p2p = ok;
                D->pcap = cap;
                D->bcap = bcap;
            }
        }
    }
    if (rc) { sllb_dd6d_destroy(D); return rc; }
    *Dout = D;
    return SLLB_OK;
}
int sllb_dd6d_set_halo_p2p(int on) {
    g_halo_p2p = on ? 1 : 0;
    return SLLB_OK;
}
int sllb_dd6d_p2p(sllb_dd6d_t D, int *enabled) {
    if (!D || !enabled) return fail(SLLB_ERR_INVALID, "dd6d_p2p: null");
    *enabled = (D->p2p && g_halo_p2p) ? 1 : 0;
    return SLLB_OK;
}
int sllb_dd6d_set_force_halo(int on) {
    g_force_halo = on ? 1 : 0;
    return SLLB_OK;
}
int sllb_dd6d_set_halo_chunks(int chunks) {
    if (chunks < 1 || chunks > 16) return fail(SLLB_ERR_INVALID, "dd6d_set_halo_chunks: 1..16");
    g_halo_chunks = chunks;
    return SLLB_OK;
}
int sllb_dd6d_destroy(sllb_dd6d_t D) {
    if (!D) return SLLB_OK;
    if (D->s_comm) {
        cudaStreamDestroy(D->s_comm); cudaStreamDestroy(D->s_comp);
        cudaEventDestroy(D->ev_start); cudaEventDestroy(D->ev_end); cudaEventDestroy(D->ev_comm0); cudaEventDestroy(D->ev_comm1);
        for (cudaEvent_t e : D->ev_chunk) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : D->ev_sten) if (e) cudaEventDestroy(e);
    }
    if (D->ev_x0) { cudaEventDestroy(D->ev_x0); cudaEventDestroy(D->ev_x1); }
    for (void *ptr : D->ipc_opened) cudaIpcCloseMemHandle(ptr);
    sllb_field_destroy(D->F);
    delete D;
    return SLLB_OK;
}
int sllb_dd6d_field(sllb_dd6d_t D, sllb_field_t *F) {
    if (!D || !F) return fail(SLLB_ERR_INVALID, "dd6d_field: null");
    *F = D->F;
    return SLLB_OK;
}
int sllb_dd6d_layout(sllb_dd6d_t D, int procs[6], int coords[6], int mn[6], int nw[6], int left[6], int right[6]) {
    if (!D) return fail(SLLB_ERR_INVALID, "dd6d_layout: null");
    for (int d = 0; d < 6; ++d) {
        if (procs) procs[d] = D->procs[d];
        if (coords) coords[d] = D->coords[d];
        if (mn) mn[d] = D->mn[d];
        if (nw) nw[d] = D->nw[d];
        if (left) left[d] = D->left[d];
        if (right) right[d] = D->right[d];
    }
    return SLLB_OK;
}

/* sll_s_apply_halo_exchange_slim_6d_real64 (sll_m_decomposition.F90:1715-2030):
 * right halo (hw_right planes) <- the first planes of the right neighbour, i.e. every rank sends its
 * first hw_right planes to its LEFT neighbour (:1839-1897); left halo (hw_left planes) <- the last planes
 * of the left neighbour (:1958-2015).  procs(axis) == 1: local periodic copy (:1840-1861). */
int sllb_dd6d_halo_exchange(sllb_dd6d_t D, int axis, int hw_left, int hw_right) {
    if (!D || axis < 0 || axis > 5 || hw_left < 0 || hw_right < 0) return fail(SLLB_ERR_INVALID, "dd6d_halo_exchange: bad arguments");
    const int n = D->nw[axis];
    if (hw_left > n || hw_right > n) return fail(SLLB_ERR_INVALID, "dd6d_halo_exchange: halo wider than the local block");
    const long long outer = outer_of(D, axis), inner = inner_of(D, axis);
    const size_t cl = (size_t)(outer * hw_left * inner), cr = (size_t)(outer * hw_right * inner);
    const bool peer_path = D->p2p && g_halo_p2p && D->procs[axis] > 1 && cl <= D->pcap && cr <= D->pcap;
    if (!peer_path) {
        SLLB_TRY(D->halo_l.ensure(cl > 0 ? cl : 1));
        SLLB_TRY(D->halo_r.ensure(cr > 0 ? cr : 1));
        D->cur_l = D->halo_l.p; D->cur_r = D->halo_r.p;
    }
    // two events owned by the decomposition, recorded around the exchange and read only if somebody asks for the time:
    // no event creation, no host synchronisation per exchange
    if (!D->ev_x0) { SLLB_CUDA(cudaEventCreate(&D->ev_x0)); SLLB_CUDA(cudaEventCreate(&D->ev_x1)); }
    SLLB_CUDA(cudaEventRecord(D->ev_x0, 0));
    D->xch_pending = false;
    if (D->procs[axis] == 1) {
        SLLB_CUDA(launch_halo_pack(D->F->d, outer, n, inner, 0, hw_right, D->halo_r.p, 0));
        SLLB_CUDA(launch_halo_pack(D->F->d, outer, n, inner, n - hw_left, hw_left, D->halo_l.p, 0));
    } else if (peer_path) {
        // my first planes are the RIGHT halo of my left neighbour, my last planes the LEFT halo of my right
        // neighbour: store them there directly.  Buffers alternate between exchanges, so a neighbour that is
        // still reading the previous halo is not disturbed; the all-reduce is the cross-rank barrier.
        const int par = D->parity;
        double *dst_r = static_cast<double *>(D->peers[(size_t)D->left[axis] * 8 + par * 2 + 1]);
        double *dst_l = static_cast<double *>(D->peers[(size_t)D->right[axis] * 8 + par * 2 + 0]);
        SLLB_TRY(halo_copy(D, axis, 0, hw_right, dst_r, 0, nullptr, 0));
        SLLB_TRY(halo_copy(D, axis, n - hw_left, hw_left, dst_l, 0, nullptr, 0));
        SLLB_TRY(dd6d_barrier(D, 0));
        D->cur_l = D->pbuf[par * 2 + 0].p; D->cur_r = D->pbuf[par * 2 + 1].p;
        D->parity ^= 1;
    } else {
        SLLB_TRY(D->send_lo.ensure(cr > 0 ? cr : 1));
        SLLB_TRY(D->send_hi.ensure(cl > 0 ? cl : 1));
        SLLB_CUDA(launch_halo_pack(D->F->d, outer, n, inner, 0, hw_right, D->send_lo.p, 0));
        SLLB_CUDA(launch_halo_pack(D->F->d, outer, n, inner, n - hw_left, hw_left, D->send_hi.p, 0));
        ncclComm_t cm = D->comm->comm;
        SLLB_NCCL(ncclGroupStart());
        if (cr > 0) {
            SLLB_NCCL(ncclSend(D->send_lo.p, cr, ncclDouble, D->left[axis], cm, 0));
            SLLB_NCCL(ncclRecv(D->halo_r.p, cr, ncclDouble, D->right[axis], cm, 0));
        }
        if (cl > 0) {
            SLLB_NCCL(ncclSend(D->send_hi.p, cl, ncclDouble, D->right[axis], cm, 0));
            SLLB_NCCL(ncclRecv(D->halo_l.p, cl, ncclDouble, D->left[axis], cm, 0));
        }
        SLLB_NCCL(ncclGroupEnd());
    }
    SLLB_CUDA(cudaEventRecord(D->ev_x1, 0));
    D->xch_pending = true;
    D->exch_pending = false;
    D->hw_l = hw_left; D->hw_r = hw_right; D->halo_axis = axis;
    return SLLB_OK;
}
int sllb_dd6d_halo_download(sllb_dd6d_t D, int side, double *host) {
    if (!D || !host || D->halo_axis < 0) return fail(SLLB_ERR_INVALID, "dd6d_halo_download: no halo present");
    const int hw = side == 0 ? D->hw_l : D->hw_r;
    const size_t cnt = (size_t)(outer_of(D, D->halo_axis) * hw * inner_of(D, D->halo_axis));
    if (cnt) SLLB_CUDA(cudaMemcpy(host, side == 0 ? D->cur_l : D->cur_r, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    return SLLB_OK;
}
int sllb_dd6d_exchange_ms(sllb_dd6d_t D, double *ms) {
    if (!D || !ms) return fail(SLLB_ERR_INVALID, "dd6d_exchange_ms: null");
    if (D->exch_pending) { // pipelined pass: first pack to last barrier on the communication stream (overlapped with compute)
        float t = 0;
        cudaEventSynchronize(D->ev_comm1);
        cudaEventElapsedTime(&t, D->ev_comm0, D->ev_comm1);
        D->exch_ms = t;
        D->exch_pending = false;
    } else if (D->xch_pending) {
        float t = 0;
        cudaEventSynchronize(D->ev_x1);
        cudaEventElapsedTime(&t, D->ev_x0, D->ev_x1);
        D->exch_ms = t;
        D->xch_pending = false;
    }
    *ms = D->exch_ms;
    return SLLB_OK;
}

} // extern "C"
// The lines of a split-axis pass cut into at most `nchunks` pieces for the pipelined exchange: whole `o` slabs when there
// are enough of them, else ranges of `in` in multiples of 32 lines (the TMA rows of the stencil kernel); one piece when
// neither works.  The pieces tile the line space exactly.
static std::vector<LineBox> chunk_boxes(long long outer, long long inner, int nchunks) {
    std::vector<LineBox> boxes;
    if (outer >= nchunks) {
        for (int c = 0; c < nchunks; ++c) {
            const long long a = outer * c / nchunks, b = outer * (c + 1) / nchunks;
            if (b > a) boxes.push_back(LineBox{a, b - a, 0, inner});
        }
    } else {
        const long long units = inner / 32;
        if (outer != 1 || inner % 32 != 0 || units < nchunks) boxes.push_back(LineBox{0, outer, 0, inner});
        else
            for (int c = 0; c < nchunks; ++c) {
                const long long a = units * c / nchunks * 32, b = units * (c + 1) / nchunks * 32;
                if (b > a) boxes.push_back(LineBox{0, 1, a, b - a});
            }
    }
    return boxes;
}
extern "C" int sllb_dd6d_chunk_boxes(long long outer, long long inner, int nchunks, long long *boxes4, int *nboxes) {
    if (outer < 1 || inner < 1 || nchunks < 1 || nchunks > 16 || !boxes4 || !nboxes) return fail(SLLB_ERR_INVALID, "dd6d_chunk_boxes: bad arguments");
    const std::vector<LineBox> b = chunk_boxes(outer, inner, nchunks);
    for (size_t k = 0; k < b.size(); ++k) { boxes4[4 * k] = b[k].o0; boxes4[4 * k + 1] = b[k].ocount; boxes4[4 * k + 2] = b[k].i0; boxes4[4 * k + 3] = b[k].icount; }
    *nboxes = (int)b.size();
    return SLLB_OK;
}
static int dd6d_pipeline_setup(sllb_dd6d *D) {
    if (D->s_comm) return SLLB_OK;
    // the stencil kernel fills every SM up to the resident-block limit with its one-warp blocks; the communication
    // stream gets the higher priority so that the few pack blocks of the next chunk are placed as soon as slots free up
    int prio_lo = 0, prio_hi = 0;
    SLLB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    SLLB_CUDA(cudaStreamCreateWithPriority(&D->s_comm, cudaStreamNonBlocking, prio_hi));
    SLLB_CUDA(cudaStreamCreateWithPriority(&D->s_comp, cudaStreamNonBlocking, prio_lo));
    SLLB_CUDA(cudaEventCreateWithFlags(&D->ev_start, cudaEventDisableTiming));
    SLLB_CUDA(cudaEventCreateWithFlags(&D->ev_end, cudaEventDisableTiming));
    SLLB_CUDA(cudaEventCreate(&D->ev_comm0));
    SLLB_CUDA(cudaEventCreate(&D->ev_comm1));
    for (cudaEvent_t &e : D->ev_chunk) SLLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t &e : D->ev_sten) SLLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return SLLB_OK;
}
// The exchange half of a pipelined split-axis pass: chunk by chunk, the edge planes go into the neighbours' halo buffers and a
// barrier follows, all on the communication stream, which starts after what has been launched on the default stream so
// far.  Event c fires when the halo of chunk c is complete on every rank.  The halo planes depend on f only (not on the
// displacement), so this half may be issued EARLY: sllb_sim6d_run starts the exchange of the first split velocity axis
// right after the x passes, and it runs under the charge density / Poisson / diagnostics work that separates the x passes
// from the v passes (dd6d_halo_prefetch); the pass itself then finds its halo already on its way.
// chain_after_axis >= 0: chunk c of this exchange waits only for the stencil kernel of chunk c of the pipelined pass along
// that axis (both passes cut the SAME slowest axis into the same ranges, so the edge planes of chunk c are final as soon
// as that kernel is) instead of for the whole previous pass.
static int dd6d_pipeline_exchange(sllb_dd6d *D, int axis, int h, int nchunks, int chain_after_axis = -1) {
    const int n = D->nw[axis];
    const long long outer = outer_of(D, axis), inner = inner_of(D, axis);
    SLLB_TRY(dd6d_pipeline_setup(D));
    const std::vector<LineBox> boxes = chunk_boxes(outer, inner, nchunks);
    // two pack blocks per SM (pack-kernel route): enough stores in flight for NVLink, and the stencil kernel of the
    // previous chunk keeps most of every SM
    static const int pack_blocks = [] { const char *e = getenv("SLLB_PACK_BLOCKS"); return (e && atoi(e) > 0) ? atoi(e) : 148 * 2; }();
    const int par = D->parity;
    double *dst_r = static_cast<double *>(D->peers[(size_t)D->left[axis] * 8 + par * 2 + 1]);
    double *dst_l = static_cast<double *>(D->peers[(size_t)D->right[axis] * 8 + par * 2 + 0]);
    const bool chained = chain_after_axis >= 0 && D->sten_axis == chain_after_axis && D->sten_chunks == (int)boxes.size();
    if (!chained) {
        SLLB_CUDA(cudaEventRecord(D->ev_start, 0));
        SLLB_CUDA(cudaStreamWaitEvent(D->s_comm, D->ev_start, 0));
    }
    SLLB_CUDA(cudaEventRecord(D->ev_comm0, D->s_comm));
    for (size_t c = 0; c < boxes.size(); ++c) {
        if (chained) SLLB_CUDA(cudaStreamWaitEvent(D->s_comm, D->ev_sten[c], 0));
        SLLB_TRY(halo_copy(D, axis, 0, h, dst_r, D->s_comm, &boxes[c], pack_blocks));
        SLLB_TRY(halo_copy(D, axis, n - h, h, dst_l, D->s_comm, &boxes[c], pack_blocks));
        SLLB_TRY(dd6d_barrier(D, D->s_comm));
        SLLB_CUDA(cudaEventRecord(D->ev_chunk[c], D->s_comm));
    }
    SLLB_CUDA(cudaEventRecord(D->ev_comm1, D->s_comm));
    D->pre_axis = axis; D->pre_h = h; D->pre_chunks = nchunks;
    return SLLB_OK;
}
/* Split-axis pass with the exchange pipelined against the stencil: chunk c of the lines is exchanged (edge planes stored
 * straight into the neighbours' halo buffers, barrier) on the communication stream while the halo-cells kernel works on
 * chunk c-1 on the compute stream.  Same kernels, same values as exchange-then-advect. */
static int dd6d_advect_axis_pipelined(sllb_dd6d *D, int axis, int stencil, const DispDesc &dd, int nchunks) {
    const int h = (stencil - 1) / 2, n = D->nw[axis];
    const long long outer = outer_of(D, axis), inner = inner_of(D, axis);
    // the exchange half: already under way if it was prefetched for exactly this pass
    if (!(D->pre_axis == axis && D->pre_h == h && D->pre_chunks == nchunks)) SLLB_TRY(dd6d_pipeline_exchange(D, axis, h, nchunks));
    D->pre_axis = -1;
    const std::vector<LineBox> boxes = chunk_boxes(outer, inner, nchunks);
    const int par = D->parity;
    D->cur_l = D->pbuf[par * 2 + 0].p; D->cur_r = D->pbuf[par * 2 + 1].p;
    // the stencil half starts after everything launched on the default stream so far (the displacement field)
    SLLB_CUDA(cudaEventRecord(D->ev_end, 0));
    SLLB_CUDA(cudaStreamWaitEvent(D->s_comp, D->ev_end, 0));
    for (size_t c = 0; c < boxes.size(); ++c) {
        SLLB_CUDA(cudaStreamWaitEvent(D->s_comp, D->ev_chunk[c], 0));
        cudaError_t e = launch_lagrange_halo(D->F->d, D->cur_l, D->cur_r, outer, n, inner, stencil, dd, g_staging, D->s_comp, &boxes[c]);
        if (e == cudaErrorInvalidValue) { cudaGetLastError(); return fail(SLLB_ERR_UNSUPPORTED, "dd6d_advect_axis: stencil / block size not implemented"); }
        SLLB_TRY(check_cuda(e, "k_lagrange_halo launch"));
        SLLB_CUDA(cudaEventRecord(D->ev_sten[c], D->s_comp));
    }
    D->sten_axis = axis; D->sten_chunks = (int)boxes.size();
    SLLB_CUDA(cudaEventRecord(D->ev_end, D->s_comp));
    SLLB_CUDA(cudaStreamWaitEvent(0, D->ev_end, 0));
    D->parity ^= 1;
    D->hw_l = h; D->hw_r = h; D->halo_axis = axis;
    D->exch_pending = true;
    D->xch_pending = false;
    return SLLB_OK;
}
// conditions under which sllb_dd6d_advect_axis takes the pipelined route (device-resident displacement assumed)
static bool dd6d_pipelined_ok(sllb_dd6d *D, int axis, int h) {
    if (g_halo_chunks < 0) { const char *e = getenv("SLLB_HALO_CHUNKS"); g_halo_chunks = (e && atoi(e) >= 1 && atoi(e) <= 16) ? atoi(e) : 4; }
    if (axis < 1 || axis > 5 || D->procs[axis] < 2) return false;
    const size_t cl = (size_t)(outer_of(D, axis) * h * inner_of(D, axis));
    return D->p2p && g_halo_p2p && cl <= D->pcap && g_halo_chunks > 1 && h <= D->nw[axis];
}
// 1: sllb_sim6d_advect_v reads the exchange time of every split pass (sllb_sim6d_halo_ms) -- a host synchronisation per
// pass, which also keeps the next exchange from being issued early; 0 (default): no timing, no synchronisation
static int g_exchange_timing = 0;
static int g_halo_prefetch = [] { const char *e = getenv("SLLB_HALO_PREFETCH"); return (e && e[0] == '0') ? 0 : 1; }();
// Start the halo exchange of a coming fixed-stencil pass along `axis` now (f is final for it, its displacement is not
// known yet).  No-op when the pass would not be pipelined.
static int dd6d_halo_prefetch(sllb_dd6d *D, int axis, int stencil) {
    const int h = (stencil - 1) / 2;
    if (!g_halo_prefetch || stencil < 3 || stencil > 11 || stencil % 2 == 0 || !dd6d_pipelined_ok(D, axis, h)) return SLLB_OK;
    return dd6d_pipeline_exchange(D, axis, h, g_halo_chunks);
}
// The exchange of the pass along `axis`, chunk by chunk behind the stencil kernels of the pipelined pass along `prev` that
// has just been issued: legal when both passes cut the slowest axis (eta6) into the same ranges -- both have at least
// `chunks` outer indices and eta6 divides evenly -- and `prev` is not the slowest axis itself.
static int dd6d_halo_prefetch_chained(sllb_dd6d *D, int axis, int prev, int stencil) {
    const int h = (stencil - 1) / 2;
    if (!g_halo_prefetch || stencil < 3 || stencil > 11 || stencil % 2 == 0 || !dd6d_pipelined_ok(D, axis, h)) return SLLB_OK;
    if (D->sten_axis != prev || prev >= axis || axis >= 5 || D->sten_chunks != g_halo_chunks) return SLLB_OK;
    if (outer_of(D, prev) < g_halo_chunks || outer_of(D, axis) < g_halo_chunks || D->nw[5] % g_halo_chunks != 0) return SLLB_OK;
    return dd6d_pipeline_exchange(D, axis, h, g_halo_chunks, prev);
}
extern "C" {
/* halo exchange + sll_s_advection_6d_lagrange_dd_slim_advect_eta{axis+1}: fixed odd stencil, in place.
 * procs(axis) == 1 uses the periodic kernel directly (same arithmetic as the reference's local periodic
 * halo copy followed by the halo-cells stencil). */
int sllb_dd6d_advect_axis(sllb_dd6d_t D, int axis, int stencil, const sllb_disp_t *disp) {
    if (!D || !disp || !disp->values) return fail(SLLB_ERR_INVALID, "dd6d_advect_axis: null");
    if (axis < 0 || axis > 5) return fail(SLLB_ERR_INVALID, "dd6d_advect_axis: bad axis");
    if (D->procs[axis] == 1 && !g_force_halo) return sllb_advect_axis(D->F, axis, SLLB_METHOD_LAGRANGE_FIXED, stencil, disp);
    if (stencil < 3 || stencil > 11 || stencil % 2 == 0) return fail(SLLB_ERR_UNSUPPORTED, "dd6d_advect_axis: fixed Lagrange stencils 3,5,7,9,11");
    if (axis == 0) return fail(SLLB_ERR_UNSUPPORTED, "dd6d_advect_axis: a split contiguous axis (eta1) is not implemented "
                                                     "(sll_f_set_process_grid splits eta1 only from 64 ranks on)");
    const int h = (stencil - 1) / 2;
    if (disp->values_on_device && dd6d_pipelined_ok(D, axis, h)) {
        DispDesc ddp;
        ddp.v = disp->values; ddp.scale = disp->scale;
        ddp.odiv = disp->odiv > 0 ? disp->odiv : 1; ddp.omod = disp->omod > 0 ? disp->omod : 1; ddp.ostr = disp->ostr;
        ddp.idiv = disp->idiv > 0 ? disp->idiv : 1; ddp.imod = disp->imod > 0 ? disp->imod : 1; ddp.istr = disp->istr;
        return dd6d_advect_axis_pipelined(D, axis, stencil, ddp, g_halo_chunks);
    }
    if (D->pre_axis >= 0) return fail(SLLB_ERR_INVALID, "dd6d_advect_axis: a prefetched halo exchange is pending for another pass");
    D->exch_pending = false;
    SLLB_TRY(sllb_dd6d_halo_exchange(D, axis, h, h));
    DispDesc dd;
    if (disp->values_on_device) dd.v = disp->values;
    else {
        if (disp->nvalues < 1) return fail(SLLB_ERR_INVALID, "dd6d_advect_axis: nvalues < 1");
        SLLB_TRY(D->F->disp_scratch.ensure((size_t)disp->nvalues));
        SLLB_CUDA(cudaMemcpyAsync(D->F->disp_scratch.p, disp->values, (size_t)disp->nvalues * sizeof(double), cudaMemcpyHostToDevice, 0));
        dd.v = D->F->disp_scratch.p;
    }
    dd.scale = disp->scale;
    dd.odiv = disp->odiv > 0 ? disp->odiv : 1; dd.omod = disp->omod > 0 ? disp->omod : 1; dd.ostr = disp->ostr;
    dd.idiv = disp->idiv > 0 ? disp->idiv : 1; dd.imod = disp->imod > 0 ? disp->imod : 1; dd.istr = disp->istr;
    cudaError_t e = launch_lagrange_halo(D->F->d, D->cur_l, D->cur_r, outer_of(D, axis), D->nw[axis], inner_of(D, axis),
                                         stencil, dd, g_staging, 0);
    if (e == cudaErrorInvalidValue) { cudaGetLastError(); return fail(SLLB_ERR_UNSUPPORTED, "dd6d_advect_axis: stencil / block size not implemented"); }
    return check_cuda(e, "k_lagrange_halo launch");
}

/* sll_s_advection_6d_spline_dd_slim_[f]advect_eta{axis+1} (sll_m_advection_6d_spline_dd_slim.F90:291-515,976-1203):
 *   prepare_exchange on every line (K9p)  ->  bc exchange + halo exchange  ->  finish_boundary_conditions,
 *   compute_interpolant, eval_disp (K9).  The boundary sums ride on the same barrier as the halo planes. */
int sllb_dd6d_advect_axis_spline(sllb_dd6d_t D, int axis, const sllb_disp_t *disp, const int32_t *shift, int hw_left,
                                 int hw_right) {
    if (!D || !disp || !disp->values) return fail(SLLB_ERR_INVALID, "dd6d_advect_axis_spline: null");
    if (axis < 0 || axis > 5) return fail(SLLB_ERR_INVALID, "dd6d_advect_axis_spline: bad axis");
    if (D->procs[axis] == 1 && !g_force_halo) return sllb_advect_axis_spline_dd(D->F, axis, disp, shift);
    if (axis == 0) return fail(SLLB_ERR_UNSUPPORTED, "dd6d_advect_axis_spline: a split contiguous axis (eta1) is not implemented");
    if (hw_left < 0 || hw_right < 0 || hw_left + hw_right < 1) return fail(SLLB_ERR_INVALID, "dd6d_advect_axis_spline: halo widths must cover shifts -hw_left..hw_right-1");
    const int np = D->nw[axis];
    if (np < spline_dd_min_points(hw_left, hw_right))
        return fail(SLLB_ERR_UNSUPPORTED, "dd6d_advect_axis_spline: too few local points for the 15-term boundary series "
                                          "(SLL_ASSERT_ALWAYS(num_points > NUM_TERMS), sll_m_cubic_spline_halo_1d.F90:79)");
    const long long outer = outer_of(D, axis), inner = inner_of(D, axis);
    const size_t nlines = (size_t)(outer * inner);
    DispDesc dd;
    SLLB_TRY(to_dispdesc(disp, D->F->disp_scratch, &dd));
    const int *d_shift = nullptr;
    SLLB_TRY(upload_shift(D->F, shift, disp->nvalues, &d_shift));
    const size_t cl = (size_t)(outer * hw_left * inner), cr = (size_t)(outer * hw_right * inner);
    const bool peer_path = D->p2p && g_halo_p2p && D->procs[axis] > 1 && cl <= D->pcap && cr <= D->pcap && nlines <= D->bcap;
    const double *bc_l = nullptr, *bc_r = nullptr;
    if (peer_path) {
        // my top cells feed the d_0 of my RIGHT neighbour (its bc_left), my bottom cells the c_np2 of my LEFT
        // neighbour (its bc_right): K9p stores them there; the halo exchange below ends with the barrier
        const int par = D->parity;
        double *dst_for_right = static_cast<double *>(D->peers[(size_t)D->right[axis] * 8 + 4 + par * 2 + 0]);
        double *dst_for_left = static_cast<double *>(D->peers[(size_t)D->left[axis] * 8 + 4 + par * 2 + 1]);
        SLLB_CUDA(launch_spline_dd_prepare(D->F->d, outer, np, inner, dd, d_shift, hw_left, hw_right, dst_for_right, dst_for_left, 0));
        bc_l = D->pbuf[4 + par * 2 + 0].p; bc_r = D->pbuf[4 + par * 2 + 1].p;
        SLLB_TRY(sllb_dd6d_halo_exchange(D, axis, hw_left, hw_right));
    } else if (D->procs[axis] == 1) {
        SLLB_TRY(D->bc_l.ensure(nlines)); SLLB_TRY(D->bc_r.ensure(nlines));
        SLLB_CUDA(launch_spline_dd_prepare(D->F->d, outer, np, inner, dd, d_shift, hw_left, hw_right, D->bc_l.p, D->bc_r.p, 0));
        bc_l = D->bc_l.p; bc_r = D->bc_r.p;
        SLLB_TRY(sllb_dd6d_halo_exchange(D, axis, hw_left, hw_right));
    } else {
        SLLB_TRY(D->bc_l.ensure(nlines)); SLLB_TRY(D->bc_r.ensure(nlines));
        SLLB_TRY(D->bc_send_r.ensure(nlines)); SLLB_TRY(D->bc_send_l.ensure(nlines));
        SLLB_CUDA(launch_spline_dd_prepare(D->F->d, outer, np, inner, dd, d_shift, hw_left, hw_right, D->bc_send_r.p, D->bc_send_l.p, 0));
        ncclComm_t cm = D->comm->comm;
        SLLB_NCCL(ncclGroupStart());
        SLLB_NCCL(ncclSend(D->bc_send_r.p, nlines, ncclDouble, D->right[axis], cm, 0));
        SLLB_NCCL(ncclRecv(D->bc_l.p, nlines, ncclDouble, D->left[axis], cm, 0));
        SLLB_NCCL(ncclSend(D->bc_send_l.p, nlines, ncclDouble, D->left[axis], cm, 0));
        SLLB_NCCL(ncclRecv(D->bc_r.p, nlines, ncclDouble, D->right[axis], cm, 0));
        SLLB_NCCL(ncclGroupEnd());
        bc_l = D->bc_l.p; bc_r = D->bc_r.p;
        SLLB_TRY(sllb_dd6d_halo_exchange(D, axis, hw_left, hw_right));
    }
    // hw_left may be 0 (all shifts >= 0): the kernel still wants a valid pointer
    const double *hl = D->cur_l ? D->cur_l : D->cur_r;
    cudaError_t e = launch_spline_dd(D->F->d, outer, np, inner, dd, d_shift, hl, hw_left, D->cur_r, hw_right, bc_l, bc_r, g_staging, 0);
    if (e == cudaErrorInvalidValue) { cudaGetLastError(); return fail(SLLB_ERR_UNSUPPORTED, "dd6d_advect_axis_spline: block size not implemented"); }
    return check_cuda(e, "k_spline_dd_strided launch");
}

} // extern "C"

/* ------------------------------------------------------------------------------------------ */
/* 3D3V: sim_bsl_vp_3d3v_cart_dd_slim, Lagrange fixed stencils, 1..P ranks                      */
/* ------------------------------------------------------------------------------------------ */
struct sllb_sim6d {
    sllb_sim6d_params_t p;
    double emin[6], emax[6], de[6];
    sllb_comm *comm = nullptr;
    sllb_dd6d *D = nullptr;
    sllb_field *F = nullptr;
    sllb_poisson *poisson = nullptr;
    DevBuf rho, phi, ex, ey, ez, small;
    DevBuf wt6, mom_part, mom9;   // fused density + moments sweep (sllb_diag.cu): weights per local velocity index, partials
    bool mom_valid = false;       // mom9 holds the nine moments of the current f
    bool want_moments = false;    // the time loop writes diagnostics rows: sllb_sim6d_fields takes the fused sweep
    // sll_t_clocks of the reference (sll_m_sim_6d_utilities.F90:132-140,765-826): wall-clock seconds per labelled phase,
    // slot [first char - '/'][second char - '/'] ('/' = one-character label).  Opt-in: every phase boundary waits for the
    // default stream.
    bool clocks_on = false;
    double clocks[44][44] = {};
    bool started = false;
    bool half_kick_pending = false; // the previous sllb_sim6d_run ended with the closing half kick of time_in_phase
    int itime = 0;
    double halo_ms = 0.0, advect_ms = 0.0;
    // spline / centred advectors: per x axis the displacement -v dt/dx of the local velocity indices, as the reference
    // stores it (sll_m_sim_bsl_vp_3d3v_cart_dd_slim.F90:590-592), and the block table of make_blocks_spline
    std::vector<double> disp_x[3];
    std::vector<int32_t> shift_x[3];
    int hw_x[3][2] = {{0, 1}, {0, 1}, {0, 1}}; // halo widths that cover every block's shift: max(-si), max(si+1)
};
__global__ void k_landau6d(double *f, Ext6 n, Ext6 lo, double d0, double d1, double d2, double d3, double d4, double d5,
                           double vmax, double factor, double alpha, double k0, double k1, double k2, double t0, double t1,
                           double t2) {
    long long ntot = 1;
    for (int d = 0; d < 6; ++d) ntot *= n.e[d];
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ntot; t += (long long)gridDim.x * blockDim.x) {
        long long r = t;
        int i[6];
        for (int d = 0; d < 6; ++d) { i[d] = (int)(r % n.e[d]) + lo.e[d]; r /= n.e[d]; }
        const double x0 = d0 * i[0], x1 = d1 * i[1], x2 = d2 * i[2];
        const double v0 = -vmax + d3 * i[3], v1 = -vmax + d4 * i[4], v2 = -vmax + d5 * i[5];
        const double a0 = v0 / t0, a1 = v1 / t1, a2 = v2 / t2;
        f[t] = factor * (1.0 + alpha * (cos(k0 * x0) * cos(k1 * x1) * cos(k2 * x2))) * exp(-0.5 * (a0 * a0 + a1 * a1 + a2 * a2));
    }
}

extern "C" {

/* rho = -dV_v sum f (sll_m_sim_6d_utilities.F90:203-245), summed over the velocity communicator (:195),
 * then Poisson and E (sll_m_sim_bsl_vp_3d3v_cart_dd_slim.F90:605-616) */
} // extern "C"
#include <chrono>
namespace {
// a running phase of the reference's stopwatch table; stop() waits for the default stream first so that the interval
// covers the device work issued inside it
struct Clock6d {
    sllb_sim6d *S; const char *label; std::chrono::steady_clock::time_point t0;
    Clock6d(sllb_sim6d *s, const char *l) : S(s), label(l) {
        if (S->clocks_on) { cudaStreamSynchronize(0); t0 = std::chrono::steady_clock::now(); }
    }
    void stop() {
        if (!S->clocks_on || !label) return;
        cudaStreamSynchronize(0);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const int i = label[0] - '/', j = label[1] ? label[1] - '/' : 0;
        if (i >= 0 && i < 44 && j >= 0 && j < 44) S->clocks[i][j] += dt;
        label = nullptr;
    }
    ~Clock6d() { stop(); }
};
} // namespace
extern "C" {
int sllb_sim6d_fields(sllb_sim6d_t S) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim6d_fields: null");
    Clock6d ck_p(S, "P"), ck_pc(S, "PC");
    S->mom_valid = false;
    if (S->want_moments) {
        // one sweep over f gives the density partials AND the nine velocity moments of the diagnostics row
        sllb_field *F = S->F;
        const long long nx = (long long)F->ext[0] * F->ext[1] * F->ext[2], nv = (long long)F->ext[3] * F->ext[4] * F->ext[5];
        if (!S->wt6.p) {
            std::vector<double> w((size_t)nv * 6);
            for (long long v = 0; v < nv; ++v) {
                const int idx[3] = {(int)(v % F->ext[3]), (int)((v / F->ext[3]) % F->ext[4]), (int)(v / ((long long)F->ext[3] * F->ext[4]))};
                for (int a = 0; a < 3; ++a) {
                    const double vel = S->emin[3 + a] + S->de[3 + a] * (double)(idx[a] + S->D->mn[3 + a]);
                    w[(size_t)v * 6 + a] = vel; w[(size_t)v * 6 + 3 + a] = vel * vel;
                }
            }
            SLLB_TRY(S->wt6.ensure(w.size()));
            SLLB_CUDA(cudaMemcpy(S->wt6.p, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
            SLLB_TRY(S->mom_part.ensure(reduce_moments6d_scratch(nx, nv)));
            SLLB_TRY(S->mom9.ensure(9));
        }
        SLLB_TRY(F->red_scratch.ensure((size_t)nx * reduce_moments6d_chunks(nv)));
        int nparts = 0;
        SLLB_CUDA(launch_reduce_moments6d(F->d, nx, nv, S->wt6.p, F->red_scratch.p, &nparts, S->mom_part.p, S->mom9.p, g_stream));
        SLLB_CUDA(launch_sum_partials(F->red_scratch.p, nx, nparts, -(S->de[3] * S->de[4] * S->de[5]), S->rho.p, g_stream));
        S->mom_valid = true;
    } else
    SLLB_TRY(sllb_reduce_velocity(S->F, 3, -(S->de[3] * S->de[4] * S->de[5]), S->rho.p));
    if (S->D->nranks > 1) SLLB_TRY(sllb_comm_allreduce_sum(S->comm, S->rho.p, (int64_t)S->p.n[0] * S->p.n[1] * S->p.n[2]));
    ck_pc.stop();
    Clock6d ck_pf(S, "PF");
    SLLB_TRY(sllb_poisson_solve(S->poisson, S->rho.p, S->phi.p, S->ex.p, S->ey.p, S->ez.p));
    return SLLB_OK;
}
/* sll_t_clocks: 1 = accumulate wall-clock seconds under the reference's labels (P, PC, PF: charge density + Poisson; D:
 * diagnostics; X, X1..X3: x advections; V, X4..X6, H4..H6: v advections and their halo exchanges), 0 (default) = off */
int sllb_sim6d_set_clocks(sllb_sim6d_t S, int on) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim6d_set_clocks: null");
    S->clocks_on = on != 0;
    g_exchange_timing = on != 0;
    return SLLB_OK;
}
/* sll_s_finalize_clocks (:775-796): one line per label with a positive time, in the order of the table */
int sllb_sim6d_write_clocks(sllb_sim6d_t S, const char *path) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim6d_write_clocks: null");
    FILE *fp = fopen(path ? path : "sll_clocks.txt", "w");
    if (!fp) return fail(SLLB_ERR_INVALID, "sim6d_write_clocks: cannot create the file");
    for (int i = 0; i < 44; ++i)
        for (int j = 0; j < 44; ++j)
            if (S->clocks[i][j] > 0.0) {
                char label[3] = {(char)('/' + i), j == 0 ? '\0' : (char)('/' + j), '\0'};
                fprintf(fp, " %s  %.16G     \n", label, S->clocks[i][j]);   // list-directed: blank, label, blanks, real
            }
    fclose(fp);
    return SLLB_OK;
}
/* sll_s_time_history_diagnostics (sll_m_sim_6d_utilities.F90:249-644): time + 13 numbers */
int sllb_sim6d_diagnostics(sllb_sim6d_t S, double time, double *row14) {
    if (!S || !row14) return fail(SLLB_ERR_INVALID, "sim6d_diagnostics: null");
    Clock6d ck_d(S, "D");
    const int *n = S->p.n;
    const double Lx = S->emax[0] - S->emin[0], Ly = S->emax[1] - S->emin[1], Lz = S->emax[2] - S->emin[2];
    const double vol_x = Lx * Ly * Lz;
    double volume = 1.0;
    for (int d = 0; d < 6; ++d) volume *= S->de[d];
    const double dV = volume / vol_x, dVx = (S->de[0] * S->de[1] * S->de[2]) / vol_x;
    std::vector<double> w1, w2;
    for (int a = 3; a < 6; ++a)
        for (int i = 0; i < S->D->nw[a]; ++i) {
            const double v = S->emin[a] + S->de[a] * (double)(i + S->D->mn[a]);
            w1.push_back(v); w2.push_back(v * v);
        }
    double m[9];
    if (S->mom_valid) SLLB_CUDA(cudaMemcpy(m, S->mom9.p, sizeof(m), cudaMemcpyDeviceToHost));   // came out of the density sweep
    else SLLB_TRY(moments_local(S->F, 3, w1.data(), w2.data(), m));
    if (S->D->nranks > 1) {
        SLLB_CUDA(cudaMemcpy(S->small.p + 5, m, sizeof(m), cudaMemcpyHostToDevice));
        SLLB_TRY(sllb_comm_allreduce_sum(S->comm, S->small.p + 5, 9));
        SLLB_CUDA(cudaMemcpy(m, S->small.p + 5, sizeof(m), cudaMemcpyDeviceToHost));
    }
    const long long nx3 = (long long)n[0] * n[1] * n[2];
    const double *arr[5] = {S->rho.p, S->phi.p, S->ex.p, S->ey.p, S->ez.p};
    for (int a = 0; a < 5; ++a) SLLB_CUDA(launch_sum_squares(arr[a], nx3, S->small.p + a, 0));
    double ss[5];
    SLLB_CUDA(cudaMemcpy(ss, S->small.p, sizeof(ss), cudaMemcpyDeviceToHost));
    row14[0] = time;
    row14[1] = m[0] * dV; row14[2] = m[2] * dV;
    for (int a = 0; a < 5; ++a) row14[3 + a] = ss[a] * dVx;
    for (int a = 0; a < 3; ++a) { row14[8 + a] = m[3 + a] * dV; row14[11 + a] = m[6 + a] * dV; }
    return SLLB_OK;
}

int sllb_sim6d_create_dist(const sllb_sim6d_params_t *p, sllb_comm_t comm, const int process_grid[6], sllb_sim6d_t *Sout) {
    if (!p || !Sout) return fail(SLLB_ERR_INVALID, "sim6d_create: null");
    SLLB_TRY(require_device());
    sllb_sim6d *S = new sllb_sim6d();
    S->p = *p;
    S->comm = comm;
    for (int d = 0; d < 3; ++d) { S->emin[d] = 0.0; S->emax[d] = p->x_max[d]; S->emin[d + 3] = -p->v_max; S->emax[d + 3] = p->v_max; }
    for (int d = 0; d < 6; ++d) S->de[d] = (S->emax[d] - S->emin[d]) / (double)p->n[d];
    const size_t nx3 = (size_t)p->n[0] * p->n[1] * p->n[2];
    int rc = sllb_dd6d_create(comm, p->n, process_grid, &S->D);
    if (!rc) {
        S->F = S->D->F;
        for (int d = 0; d < 3; ++d)
            if (S->D->procs[d] != 1)
                rc = fail(SLLB_ERR_UNSUPPORTED, "sim6d_create: process grids that split eta1..3 are not implemented (the reference's "
                                                "table splits velocity axes only up to 8 ranks, sll_m_decomposition.F90:2489-2498)");
    }
    if (!rc) rc = sllb_poisson3d_create(p->n[0], p->n[1], p->n[2], S->emax[0], S->emax[1], S->emax[2], &S->poisson);
    if (!rc) rc = S->rho.ensure(nx3);
    if (!rc) rc = S->phi.ensure(nx3);
    if (!rc) rc = S->ex.ensure(nx3);
    if (!rc) rc = S->ey.ensure(nx3);
    if (!rc) rc = S->ez.ensure(nx3);
    if (!rc) rc = S->small.ensure(16);
    if (!rc) {
        Ext6 n, lo;
        for (int d = 0; d < 6; ++d) { n.e[d] = S->D->nw[d]; lo.e[d] = S->D->mn[d]; }
        const double twopi = 2.0 * 3.14159265358979323846;
        const double factor = 1.0 / (pow(sqrt(twopi), 3) * (p->v_thermal[0] * p->v_thermal[1] * p->v_thermal[2]));
        k_landau6d<<<148 * 8, 256>>>(S->F->d, n, lo, S->de[0], S->de[1], S->de[2], S->de[3], S->de[4], S->de[5], p->v_max, factor,
                                     p->alpha, p->kx[0], p->kx[1], p->kx[2], p->v_thermal[0], p->v_thermal[1], p->v_thermal[2]);
        rc = check_cuda(cudaGetLastError(), "k_landau6d");
    }
    if (!rc && p->advector != SLLB_ADVECTOR_FIXED && p->advector != SLLB_ADVECTOR_CENTERED && p->advector != SLLB_ADVECTOR_SPLINE)
        rc = fail(SLLB_ERR_UNSUPPORTED, "sim6d_create: Interpolator type not implemented.");
    if (!rc && p->advector != SLLB_ADVECTOR_FIXED)
        for (int d = 0; d < 3 && !rc; ++d) {
            const int nv = S->D->nw[d + 3];
            S->disp_x[d].resize(nv); S->shift_x[d].resize(nv);
            for (int l = 0; l < nv; ++l) {
                const double v = S->emin[d + 3] + S->de[d + 3] * (double)(l + S->D->mn[d + 3]); // sll_s_set_local_grid
                S->disp_x[d][l] = -v * p->delta_t / S->de[d];
            }
            if (p->advector == SLLB_ADVECTOR_SPLINE) rc = sllb_spline_dd_blocks(nv, S->disp_x[d].data(), S->shift_x[d].data(), nullptr, nullptr);
            else rc = sllb_lagrange_dd_blocks(nv, p->stencil_x, S->disp_x[d].data(), S->shift_x[d].data(), nullptr, nullptr);
            int hl = 0, hr = 0;
            for (int l = 0; l < nv; ++l) {
                const int si = S->shift_x[d][l];
                if (si == SLLB_SHIFT_SKIP) continue;
                if (-si > hl) hl = -si;
                if (si + 1 > hr) hr = si + 1;
            }
            if (hl + hr < 1) hr = 1;
            S->hw_x[d][0] = hl; S->hw_x[d][1] = hr;
        }
    if (!rc) rc = sllb_sim6d_fields(S);
    if (rc) { sllb_sim6d_destroy(S); return rc; }
    *Sout = S;
    return SLLB_OK;
}
int sllb_sim6d_create(const sllb_sim6d_params_t *p, sllb_sim6d_t *Sout) { return sllb_sim6d_create_dist(p, nullptr, nullptr, Sout); }
int sllb_sim6d_destroy(sllb_sim6d_t S) {
    if (!S) return SLLB_OK;
    sllb_poisson_destroy(S->poisson);
    sllb_dd6d_destroy(S->D);
    delete S;
    return SLLB_OK;
}
int sllb_sim6d_field(sllb_sim6d_t S, sllb_field_t *F) {
    if (!S || !F) return fail(SLLB_ERR_INVALID, "sim6d_field: null");
    *F = S->F;
    return SLLB_OK;
}
int sllb_sim6d_decomposition(sllb_sim6d_t S, sllb_dd6d_t *D) {
    if (!S || !D) return fail(SLLB_ERR_INVALID, "sim6d_decomposition: null");
    *D = S->D;
    return SLLB_OK;
}
/* advect_x (:817-865): eta1..3 with disp_eta = -v*dt/dx (:590-592); x is not split, local periodic wrap */
int sllb_sim6d_advect_x(sllb_sim6d_t S) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim6d_advect_x: null");
    Clock6d ck_x(S, "X");
    static const char *const xl[3] = {"X1", "X2", "X3"};
    if (S->p.advector != SLLB_ADVECTOR_FIXED) {
        // fadvect_eta1..3 (:866-880): the displacement array of the conjugate velocity axis, block by block in the
        // reference, line by line here (the block only fixes the integer part, carried by the shift table)
        for (int d = 0; d < 3; ++d) {
            sllb_disp_t ds;
            memset(&ds, 0, sizeof(ds));
            ds.values = S->disp_x[d].data(); ds.nvalues = (int64_t)S->disp_x[d].size(); ds.values_on_device = 0; ds.scale = 1.0;
            long long stride = 1;
            for (int a = d + 1; a < d + 3; ++a) stride *= S->D->nw[a];
            ds.odiv = stride; ds.omod = S->D->nw[d + 3]; ds.ostr = 1; ds.idiv = 1; ds.imod = 1; ds.istr = 0;
            Clock6d ck_a(S, xl[d]);
            if (S->p.advector == SLLB_ADVECTOR_SPLINE) SLLB_TRY(sllb_dd6d_advect_axis_spline(S->D, d, &ds, S->shift_x[d].data(), S->hw_x[d][0], S->hw_x[d][1]));
            else SLLB_TRY(sllb_advect_axis(S->F, d, SLLB_METHOD_LAGRANGE_CENTERED, S->p.stencil_x, &ds));
        }
        return SLLB_OK;
    }
    // eta1 and eta2 in one sweep over f (K2d), then eta3.  Opt-in (SLLB_LAGRANGE_PLANE=1): on 32 x 32 planes the fused
    // kernel is bound by instruction issue and block barriers, 1.6-1.7 ms against 0.76 + 0.62 ms for the two separate
    // passes on a 262 M-point block (profiles/r01_ncu_lagrange_plane_s8.txt), so the separate passes stay the default.
    int first = 0;
    static const bool use_plane = [] { const char *e = getenv("SLLB_LAGRANGE_PLANE"); return e && e[0] == '1'; }();
    if (use_plane) {
        DispDesc d0, d1;
        sllb_field *F = S->F;
        SLLB_TRY(F->disp_scratch2.ensure((size_t)F->ext[4]));
        SLLB_TRY(make_affine_disp(F, 0, 3, S->emin[3] + S->D->mn[3] * S->de[3], S->de[3], -S->p.delta_t / S->de[0], &d0));
        SLLB_CUDA(launch_affine(F->disp_scratch2.p, F->ext[4], S->emin[4] + S->D->mn[4] * S->de[4], S->de[4], g_stream));
        d1.v = F->disp_scratch2.p; d1.scale = -S->p.delta_t / S->de[1];
        d1.odiv = (long long)F->ext[2] * F->ext[3]; d1.omod = F->ext[4]; d1.ostr = 1; d1.idiv = d1.imod = 1; d1.istr = 0;
        const int rc = advect_lagrange_plane_dev(F, SLLB_METHOD_LAGRANGE_FIXED, S->p.stencil_x, d0, d1);
        if (rc == SLLB_OK) first = 2;
        else if (rc != SLLB_ERR_UNSUPPORTED) return rc;
    }
    for (int d = first; d < 3; ++d) {
        Clock6d ck_a(S, xl[d]);
        SLLB_TRY(sllb_advect_axis_affine(S->F, d, SLLB_METHOD_LAGRANGE_FIXED, S->p.stencil_x, d + 3,
                                         S->emin[d + 3] + S->D->mn[d + 3] * S->de[d + 3], S->de[d + 3], -S->p.delta_t / S->de[d]));
    }
    return SLLB_OK;
}
/* advect_v (:889-958): per axis halo exchange, then eta4..6 with displacement E*dt/dv as a 3D field */
int sllb_sim6d_advect_v(sllb_sim6d_t S, double dt) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim6d_advect_v: null");
    const double *E[3] = {S->ex.p, S->ey.p, S->ez.p};
    const long long nx3 = (long long)S->p.n[0] * S->p.n[1] * S->p.n[2];
    Clock6d ck_v(S, "V");
    static const char *const vl[3] = {"X4", "X5", "X6"};
    for (int d = 0; d < 3; ++d) {
        Clock6d ck_a(S, vl[d]);
        const double halo_before = S->halo_ms;
        struct HaloClock {   // H4..H6: the exchange part of this pass, from the device-side exchange timer
            sllb_sim6d *S; int d; double before;
            ~HaloClock() { if (S->clocks_on) S->clocks['H' - '/']['4' + d - '/'] += (S->halo_ms - before) * 1e-3; }
        } hck{S, d, halo_before};
        sllb_disp_t ds;
        memset(&ds, 0, sizeof(ds));
        ds.values = E[d]; ds.nvalues = nx3; ds.values_on_device = 1; ds.scale = dt / S->de[3 + d];
        ds.odiv = ds.omod = 1; ds.ostr = 0; ds.idiv = 1; ds.imod = nx3; ds.istr = 1;
        // splines: halo of one plane on both sides, shifts 0 and -1 (:1082-1087); else fixed Lagrange
        if (S->p.advector == SLLB_ADVECTOR_SPLINE) SLLB_TRY(sllb_dd6d_advect_axis_spline(S->D, 3 + d, &ds, nullptr, 1, 1));
        else {
            S->D->sten_axis = -1;
            SLLB_TRY(sllb_dd6d_advect_axis(S->D, 3 + d, S->p.stencil_v, &ds));
            if (g_exchange_timing && S->D->procs[3 + d] > 1) { double ms = 0; SLLB_TRY(sllb_dd6d_exchange_ms(S->D, &ms)); S->halo_ms += ms; }
            // the halo of the next v pass follows this pass chunk by chunk instead of waiting for all of it
            if (d < 2 && S->D->sten_axis == 3 + d) SLLB_TRY(dd6d_halo_prefetch_chained(S->D, 4 + d, 3 + d, S->p.stencil_v));
            continue;
        }
        if (g_exchange_timing && S->D->procs[3 + d] > 1) { double ms = 0; SLLB_TRY(sllb_dd6d_exchange_ms(S->D, &ms)); S->halo_ms += ms; }
    }
    return SLLB_OK;
}
/* For hosts that keep their own time loop: call between sllb_sim6d_advect_x and sllb_sim6d_fields.  f is final for the
 * first v pass then, so the halo exchange of the first split velocity axis starts now and runs under the field solve. */
int sllb_sim6d_prefetch_v_halo(sllb_sim6d_t S) {
    if (!S) return fail(SLLB_ERR_INVALID, "sim6d_prefetch_v_halo: null");
    if (S->p.advector == SLLB_ADVECTOR_SPLINE) return SLLB_OK;
    return dd6d_halo_prefetch(S->D, 3, S->p.stencil_v);
}
int sllb_dd6d_set_exchange_timing(int on) {
    g_exchange_timing = on ? 1 : 0;
    return SLLB_OK;
}
/* Rows sllb_sim6d_run(S, nsteps, rows) will write: the first call on a handle also writes the t = 0 row. */
int sllb_sim6d_run_rows(sllb_sim6d_t S, int nsteps, int *nrows) {
    if (!S || !nrows || nsteps < 0) return fail(SLLB_ERR_INVALID, "sim6d_run_rows: bad arguments");
    *nrows = nsteps + (S->started ? 0 : 1);
    return SLLB_OK;
}
/* Time loop (:643-760).  With time_in_phase the reference closes its one loop with a half V step so that x and v are
 * known at the same time; here a handle may be run in several calls, so a call that follows such an ending opens with
 * the other half of that V step (E has not changed in between: a V step leaves rho untouched), i.e.
 * run(a) + run(b) advances f exactly as far as run(a + b); the results differ by the interpolation error of doing that
 * V step in two halves, not by a lost half step. */
int sllb_sim6d_run(sllb_sim6d_t S, int nsteps, double *rows) {
    if (!S || nsteps < 0) return fail(SLLB_ERR_INVALID, "sim6d_run: bad arguments");
    int row = 0;
    S->want_moments = rows != nullptr;   // the field solve of every step is followed by a diagnostics row
    if (!S->started) {
        if (rows) SLLB_TRY(sllb_sim6d_diagnostics(S, 0.0, rows));
        row = 1;
        SLLB_TRY(sllb_sim6d_advect_v(S, 0.5 * S->p.delta_t));
        S->started = true;
    } else if (S->half_kick_pending && nsteps > 0) {
        SLLB_TRY(sllb_sim6d_advect_v(S, 0.5 * S->p.delta_t));
        S->half_kick_pending = false;
    }
    for (int it = 1; it <= nsteps; ++it) {
        SLLB_TRY(sllb_sim6d_advect_x(S));
        // f is final for the first v pass: its halo leaves now, under the field solve and the diagnostics
        if (S->p.advector != SLLB_ADVECTOR_SPLINE) SLLB_TRY(dd6d_halo_prefetch(S->D, 3, S->p.stencil_v));
        SLLB_TRY(sllb_sim6d_fields(S));
        S->itime += 1;
        if (rows) SLLB_TRY(sllb_sim6d_diagnostics(S, (double)S->itime * S->p.delta_t, rows + 14 * (row++)));
        if (S->p.time_in_phase && it == nsteps) {
            SLLB_TRY(sllb_sim6d_advect_v(S, 0.5 * S->p.delta_t));
            S->half_kick_pending = true;
        } else SLLB_TRY(sllb_sim6d_advect_v(S, S->p.delta_t));
        S->mom_valid = false;
    }
    S->want_moments = false;
    SLLB_CUDA(cudaDeviceSynchronize());
    if (S->D->flag_barrier) {
        double err = 0.0;
        SLLB_CUDA(cudaMemcpy(&err, S->D->errflag.p, sizeof(double), cudaMemcpyDeviceToHost));
        if (err != 0.0) return fail(SLLB_ERR_CUDA, "sim6d_run: a rank did not reach the flag barrier within the time-out");
    }
    return SLLB_OK;
}
int sllb_sim6d_halo_ms(sllb_sim6d_t S, double *ms, int reset) {
    if (!S || !ms) return fail(SLLB_ERR_INVALID, "sim6d_halo_ms: null");
    *ms = S->halo_ms;
    if (reset) S->halo_ms = 0.0;
    return SLLB_OK;
}
} // extern "C"
