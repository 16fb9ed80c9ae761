// sllb_capi.cu -- C ABI (include/sll_b200.h): handles, device-resident fields, batched axis
// advection, velocity reduction, moments and the cuFFT Poisson solvers.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "sllb_internal.h"

namespace sllb {

static thread_local std::string t_error;
int g_staging = STAGING_AUTO;
// the stream every launch of this thread goes to: the legacy default stream, except while sllb_sim2d_run records one
// time step into a CUDA graph (stream capture needs a stream of its own)
thread_local cudaStream_t g_stream = 0;
static int g_device = 0;
static bool g_device_ok = false;

void set_error(const std::string &msg) { t_error = msg; }
int fail(int code, const std::string &msg) {
    t_error = msg;
    return code;
}
int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return SLLB_OK;
    if (e == cudaErrorInvalidValue && what && strstr(what, "launch_advect"))
        return fail(SLLB_ERR_UNSUPPORTED, std::string("advection method / stencil / size not implemented: ") + what);
    return fail(SLLB_ERR_CUDA, std::string(what ? what : "cuda") + ": " + cudaGetErrorString(e));
}
int check_cufft(cufftResult r, const char *what) {
    if (r == CUFFT_SUCCESS) return SLLB_OK;
    return fail(SLLB_ERR_CUDA, std::string(what ? what : "cufft") + ": cufft error " + std::to_string((int)r));
}
int check_nccl(ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return SLLB_OK;
    return fail(SLLB_ERR_CUDA, std::string(what) + ": " + ncclGetErrorString(r));
}
int require_device() {
    if (g_device_ok) return SLLB_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(SLLB_ERR_NO_DEVICE, "no usable CUDA device (there is no CPU fallback)");
    }
    if (g_device >= n) return fail(SLLB_ERR_NO_DEVICE, "requested device index out of range");
    e = cudaSetDevice(g_device);
    if (e != cudaSuccess) return fail(SLLB_ERR_NO_DEVICE, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    g_device_ok = true;
    return SLLB_OK;
}

// Peer mapping over CUDA IPC: every rank contributes `count` device buffers; on return peers[r*count + k] is
// rank r's k-th buffer mapped into this process (my own pointers for r == me).  cudaIpcGetMemHandle names a
// whole allocation, so the offset of each buffer inside its allocation travels with the handle.  *ok is false
// (on every rank) when any rank could not export or open a handle; callers then keep to NCCL.
int peer_map_buffers(sllb_comm *comm, void *const *mine, int count, std::vector<void *> &peers,
                     std::vector<void *> &opened, bool *ok) {
    *ok = false;
    if (!comm || comm->nranks < 2) return SLLB_OK;
    typedef int (*getrange_t)(unsigned long long *, size_t *, unsigned long long);
    getrange_t get_range = nullptr;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn)
            get_range = reinterpret_cast<getrange_t>(fn);
        else cudaGetLastError();
    }
    struct Slot { cudaIpcMemHandle_t h; long long offset; long long ok; };
    static_assert(sizeof(Slot) % 8 == 0, "slot size");
    const int P = comm->nranks;
    std::vector<Slot> my((size_t)count), all((size_t)count * P);
    int good = 1;
    for (int k = 0; k < count; ++k) {
        unsigned long long base = 0; size_t size = 0;
        memset(&my[k], 0, sizeof(Slot));
        if (!get_range || get_range(&base, &size, (unsigned long long)(uintptr_t)mine[k]) != 0) { good = 0; continue; }
        if (cudaIpcGetMemHandle(&my[k].h, reinterpret_cast<void *>((uintptr_t)base)) != cudaSuccess) { cudaGetLastError(); good = 0; continue; }
        my[k].offset = (long long)((unsigned long long)(uintptr_t)mine[k] - base);
    }
    for (int k = 0; k < count; ++k) my[k].ok = good;
    char *dsend = nullptr, *drecv = nullptr;
    SLLB_CUDA(cudaMalloc(&dsend, (size_t)count * sizeof(Slot)));
    SLLB_CUDA(cudaMalloc(&drecv, (size_t)count * P * sizeof(Slot)));
    SLLB_CUDA(cudaMemcpy(dsend, my.data(), (size_t)count * sizeof(Slot), cudaMemcpyHostToDevice));
    SLLB_NCCL(ncclAllGather(dsend, drecv, (size_t)count * sizeof(Slot), ncclChar, comm->comm, 0));
    SLLB_CUDA(cudaMemcpy(all.data(), drecv, (size_t)count * P * sizeof(Slot), cudaMemcpyDeviceToHost));
    for (int r = 0; r < P; ++r) if (!all[(size_t)r * count].ok) good = 0;
    peers.assign((size_t)count * P, nullptr);
    if (good) {
        // several buffers of one rank may live in the same allocation: open each distinct handle once
        for (int r = 0; r < P && good; ++r)
            for (int k = 0; k < count; ++k) {
                if (r == comm->rank) { peers[(size_t)r * count + k] = mine[k]; continue; }
                void *ptr = nullptr;
                for (int j = 0; j < k; ++j)
                    if (memcmp(&all[(size_t)r * count + j].h, &all[(size_t)r * count + k].h, sizeof(cudaIpcMemHandle_t)) == 0)
                        ptr = static_cast<char *>(peers[(size_t)r * count + j]) - all[(size_t)r * count + j].offset;
                if (!ptr) {
                    if (cudaIpcOpenMemHandle(&ptr, all[(size_t)r * count + k].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = 0; break; }
                    opened.push_back(ptr);
                }
                peers[(size_t)r * count + k] = static_cast<char *>(ptr) + all[(size_t)r * count + k].offset;
            }
    }
    // everybody must agree, otherwise nobody uses peer stores
    double h = good ? 0.0 : 1.0, *dflag = reinterpret_cast<double *>(dsend);
    SLLB_CUDA(cudaMemcpy(dflag, &h, sizeof(double), cudaMemcpyHostToDevice));
    SLLB_NCCL(ncclAllReduce(dflag, dflag, 1, ncclDouble, ncclSum, comm->comm, 0));
    SLLB_CUDA(cudaMemcpy(&h, dflag, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dsend); cudaFree(drecv);
    *ok = (h == 0.0);
    return SLLB_OK;
}

int DevBuf::ensure(size_t count) {
    if (count <= n && p) return SLLB_OK;
    release();
    SLLB_CUDA(cudaMalloc(&p, count * sizeof(double)));
    n = count;
    return SLLB_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
}

// generic <=6D copy with periodic wrap of the source index: dst[i] = src[i mod ext_src]
__global__ void __launch_bounds__(256) k_copy_wrap(const double *__restrict__ src, Ext6 es, double *__restrict__ dst,
                                                   Ext6 ed, long long ntot) {
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < ntot; t += (long long)gridDim.x * 256) {
        long long r = t, sidx = 0, sstride = 1;
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            int i = (int)(r % ed.e[d]);
            r /= ed.e[d];
            if (i >= es.e[d]) i -= es.e[d];
            sidx += (long long)i * sstride;
            sstride *= es.e[d];
        }
        dst[t] = src[sidx];
    }
}

int field_alloc(int ndim, const int *ext, sllb_field **F) {
    if (ndim < 1 || ndim > 6 || !ext || !F) return fail(SLLB_ERR_INVALID, "field_create: bad arguments");
    SLLB_TRY(require_device());
    sllb_field *f = new sllb_field();
    f->ndim = ndim;
    f->total = 1;
    for (int d = 0; d < ndim; ++d) {
        if (ext[d] < 1) { delete f; return fail(SLLB_ERR_INVALID, "field_create: extent < 1"); }
        f->ext[d] = ext[d];
        f->total *= ext[d];
    }
    cudaError_t e = cudaMalloc(&f->d, (size_t)f->total * sizeof(double));
    if (e != cudaSuccess) { delete f; return check_cuda(e, "cudaMalloc(field)"); }
    *F = f;
    return SLLB_OK;
}
int field_wrap(int ndim, const int *ext, double *d, sllb_field **F) {
    sllb_field *f = new sllb_field();
    f->ndim = ndim;
    f->total = 1;
    for (int k = 0; k < ndim; ++k) { f->ext[k] = ext[k]; f->total *= ext[k]; }
    f->d = d;
    f->owns = false;
    *F = f;
    return SLLB_OK;
}

int advect_axis_dev(sllb_field *F, int axis, int method, int order, const DispDesc &dd, const RemapDst *remap,
                    double *linesum, const LineDiag *diag, const LineSub *sub) {
    if (!F || axis < 0 || axis >= F->ndim) return fail(SLLB_ERR_INVALID, "advect_axis: bad field/axis");
    long long inner = 1, outer = 1;
    for (int d = 0; d < axis; ++d) inner *= F->ext[d];
    for (int d = axis + 1; d < F->ndim; ++d) outer *= F->ext[d];
    cudaError_t e = launch_advect(F->d, outer, F->ext[axis], inner, method, order, dd, g_staging, g_stream, remap, linesum, diag, sub);
    if (e == cudaErrorNotSupported) { cudaGetLastError(); return fail(SLLB_ERR_UNSUPPORTED, "advect_axis: line sums are produced by the chunked strided spline kernel only"); }
    if (e == cudaErrorInvalidValue) {
        cudaGetLastError();
        return fail(SLLB_ERR_UNSUPPORTED, "advect_axis: method/order/line length not implemented (spline: orders 4, 6, 8; "
                                          "Lagrange fixed: 3,5,7,9,11; centred: even 4..18; 8 <= n, line must fit shared memory)");
    }
    return check_cuda(e, "advect kernel launch");
}

// K1c: spline passes along axes 0 and 1 on every plane in one sweep (+ optional rho = scale * sum over the
// other axes of the result).  SLLB_ERR_UNSUPPORTED when the shape / displacement pattern does not fit.
int g_plane_kernel = 1;
int g_poisson_direct = [] { const char *e = getenv("SLLB_POISSON_DIRECT"); return (e && e[0] == '0') ? 0 : 1; }();
int advect_plane_dev(sllb_field *F, const DispDesc &dd0, const DispDesc &dd1, double rho_scale, double *d_rho,
                     const RemapDst *remap, const double **partials_out, int *nparts_out) {
    if (!F || F->ndim < 2) return fail(SLLB_ERR_INVALID, "advect_plane: bad field");
    if (!g_plane_kernel) return fail(SLLB_ERR_UNSUPPORTED, "advect_plane: disabled (sllb_set_plane_kernel)");
    const int n1 = F->ext[0], n2 = F->ext[1];
    long long nplanes = 1;
    for (int d = 2; d < F->ndim; ++d) nplanes *= F->ext[d];
    double *partial = nullptr;
    int nparts = 0;
    if (d_rho || partials_out) {
        nparts = plane_grid(n1, n2, nplanes);
        SLLB_TRY(F->red_scratch.ensure((size_t)nparts * n1 * n2));
        partial = F->red_scratch.p;
    }
    cudaError_t e = launch_spline_plane(F->d, n1, n2, nplanes, dd0, dd1, partial, g_stream, remap);
    if (e == cudaErrorNotSupported) { cudaGetLastError(); return fail(SLLB_ERR_UNSUPPORTED, "advect_plane: plane shape or displacement pattern not supported"); }
    SLLB_TRY(check_cuda(e, "k_spline_plane launch"));
    if (partials_out) { *partials_out = partial; *nparts_out = nparts; }   // the caller sums (and scales) them
    else if (d_rho) SLLB_CUDA(launch_sum_partials(partial, (long long)n1 * n2, nparts, rho_scale, d_rho, g_stream));
    return SLLB_OK;
}

// K2d: Lagrange passes along axes 0 and 1 on every plane in one sweep
int advect_lagrange_plane_dev(sllb_field *F, int method, int order, const DispDesc &dd0, const DispDesc &dd1) {
    if (!F || F->ndim < 2) return fail(SLLB_ERR_INVALID, "advect_plane: bad field");
    if (!g_plane_kernel) return fail(SLLB_ERR_UNSUPPORTED, "advect_plane: disabled (sllb_set_plane_kernel)");
    long long nplanes = 1;
    for (int d = 2; d < F->ndim; ++d) nplanes *= F->ext[d];
    cudaError_t e = launch_lagrange_plane(F->d, F->ext[0], F->ext[1], nplanes, method, order, dd0, dd1, g_stream);
    if (e == cudaErrorNotSupported) { cudaGetLastError(); return fail(SLLB_ERR_UNSUPPORTED, "advect_plane: plane shape, stencil or displacement pattern not supported"); }
    return check_cuda(e, "k_lagrange_plane launch");
}

int moments_local(sllb_field *F, int nv, const double *w1, const double *w2, double *out) {
    if (!F || nv < 0 || nv >= F->ndim || !out) return fail(SLLB_ERR_INVALID, "moments: bad arguments");
    long long nx = 1, nvt = 1;
    for (int d = 0; d < F->ndim - nv; ++d) nx *= F->ext[d];
    for (int d = F->ndim - nv; d < F->ndim; ++d) nvt *= F->ext[d];
    SLLB_TRY(F->rows.ensure((size_t)nvt * 3));
    SLLB_CUDA(launch_row_sums(F->d, nx, nvt, F->rows.p, g_stream));
    std::vector<double> h((size_t)nvt * 3);
    SLLB_CUDA(cudaMemcpy(h.data(), F->rows.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3 + 2 * nv; ++k) out[k] = 0.0;
    int idx[6] = {0, 0, 0, 0, 0, 0};
    const int *ve = F->ext + (F->ndim - nv);
    std::vector<long long> woff(nv + 1, 0);
    for (int a = 0; a < nv; ++a) woff[a + 1] = woff[a] + ve[a];
    for (long long v = 0; v < nvt; ++v) {
        const double s0 = h[3 * v], s1 = h[3 * v + 1], s2 = h[3 * v + 2];
        out[0] += s0; out[1] += s1; out[2] += s2;
        for (int a = 0; a < nv; ++a) {
            if (w1) out[3 + a] += s0 * w1[woff[a] + idx[a]];
            if (w2) out[3 + nv + a] += s0 * w2[woff[a] + idx[a]];
        }
        for (int a = 0; a < nv; ++a) { // increment mixed-radix index, first velocity axis fastest
            if (++idx[a] < ve[a]) break;
            idx[a] = 0;
        }
    }
    return SLLB_OK;
}

} // namespace sllb

using namespace sllb;

extern "C" {

const char *sllb_last_error(void) { return t_error.c_str(); }
int sllb_version(void) { return 100; }
int sllb_init(int device) {
    g_device = device;
    g_device_ok = false;
    return require_device();
}
int sllb_device_count(int *count) {
    if (!count) return fail(SLLB_ERR_INVALID, "device_count: null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return SLLB_OK;
}
int sllb_synchronize(void) {
    SLLB_TRY(require_device());
    SLLB_CUDA(cudaDeviceSynchronize());
    return SLLB_OK;
}
int64_t sllb_launch_count(void) { return launch_count(); }
void sllb_launch_count_reset(void) { launch_count_reset(); }
int sllb_set_staging(int mode) {
    if (mode < 0 || mode > 2) return fail(SLLB_ERR_INVALID, "set_staging: mode must be 0,1,2");
    g_staging = mode;
    return SLLB_OK;
}

int sllb_set_plane_kernel(int on, int points_per_thread) {
    if (points_per_thread != 0 && points_per_thread != 16 && points_per_thread != 32)
        return fail(SLLB_ERR_INVALID, "set_plane_kernel: points_per_thread must be 0 (auto), 16 or 32");
    g_plane_kernel = on ? 1 : 0;
    g_plane_ept = points_per_thread;
    return SLLB_OK;
}

int sllb_set_plane_variant(int tmem_accumulators, int const_extents) {
    if (tmem_accumulators < -1 || tmem_accumulators > 1 || const_extents < 0 || const_extents > 1)
        return fail(SLLB_ERR_INVALID, "set_plane_variant: tmem_accumulators in {-1, 0, 1}, const_extents in {0, 1}");
    g_plane_tmem = tmem_accumulators;
    g_plane_const_dims = const_extents;
    return SLLB_OK;
}

int sllb_set_remap_rotation(int on) {
    g_remap_rotation = on ? 1 : 0;
    return SLLB_OK;
}

int sllb_set_spline_split(int chunks) {
    if (chunks != -1 && chunks != 1 && chunks != 2 && chunks != 4 && chunks != 8)
        return fail(SLLB_ERR_INVALID, "set_spline_split: chunks must be -1, 1, 2, 4 or 8");
    g_spline_split = chunks;
    return SLLB_OK;
}

/* ---------------- fields ---------------- */
int sllb_field_create(int ndim, const int *extents, sllb_field_t *F) { return field_alloc(ndim, extents, F); }
int sllb_field_destroy(sllb_field_t F) {
    if (!F) return SLLB_OK;
    if (F->owns && F->d) cudaFree(F->d);
    delete F;
    return SLLB_OK;
}
int sllb_field_device_ptr(sllb_field_t F, double **dptr) {
    if (!F || !dptr) return fail(SLLB_ERR_INVALID, "field_device_ptr: null");
    *dptr = F->d;
    return SLLB_OK;
}
int sllb_field_extents(sllb_field_t F, int *ndim, int *extents) {
    if (!F) return fail(SLLB_ERR_INVALID, "field_extents: null");
    if (ndim) *ndim = F->ndim;
    if (extents) for (int d = 0; d < F->ndim; ++d) extents[d] = F->ext[d];
    return SLLB_OK;
}
static bool any_dup(const sllb_field *F, const int *dup) {
    if (!dup) return false;
    for (int d = 0; d < F->ndim; ++d) if (dup[d]) return true;
    return false;
}
int sllb_field_upload(sllb_field_t F, const double *host, const int *dup_last) {
    if (!F || !host) return fail(SLLB_ERR_INVALID, "field_upload: null");
    SLLB_TRY(require_device());
    if (!any_dup(F, dup_last)) {
        SLLB_CUDA(cudaMemcpy(F->d, host, (size_t)F->total * sizeof(double), cudaMemcpyHostToDevice));
        return SLLB_OK;
    }
    Ext6 es, ed;
    long long ntot_src = 1;
    for (int d = 0; d < 6; ++d) {
        ed.e[d] = F->ext[d];
        es.e[d] = F->ext[d] + ((d < F->ndim && dup_last[d]) ? 1 : 0);
        ntot_src *= es.e[d];
    }
    SLLB_TRY(F->stage.ensure((size_t)ntot_src));
    SLLB_CUDA(cudaMemcpy(F->stage.p, host, (size_t)ntot_src * sizeof(double), cudaMemcpyHostToDevice));
    k_copy_wrap<<<148 * 8, 256>>>(F->stage.p, es, F->d, ed, F->total);
    SLLB_CUDA(cudaGetLastError());
    SLLB_CUDA(cudaDeviceSynchronize());
    F->stage.release();
    return SLLB_OK;
}
int sllb_field_download(sllb_field_t F, double *host, const int *dup_last) {
    if (!F || !host) return fail(SLLB_ERR_INVALID, "field_download: null");
    SLLB_TRY(require_device());
    if (!any_dup(F, dup_last)) {
        SLLB_CUDA(cudaMemcpy(host, F->d, (size_t)F->total * sizeof(double), cudaMemcpyDeviceToHost));
        return SLLB_OK;
    }
    Ext6 es, ed;
    long long ntot_dst = 1;
    for (int d = 0; d < 6; ++d) {
        es.e[d] = F->ext[d];
        ed.e[d] = F->ext[d] + ((d < F->ndim && dup_last[d]) ? 1 : 0);
        ntot_dst *= ed.e[d];
    }
    SLLB_TRY(F->stage.ensure((size_t)ntot_dst));
    k_copy_wrap<<<148 * 8, 256>>>(F->d, es, F->stage.p, ed, ntot_dst);
    SLLB_CUDA(cudaGetLastError());
    SLLB_CUDA(cudaMemcpy(host, F->stage.p, (size_t)ntot_dst * sizeof(double), cudaMemcpyDeviceToHost));
    F->stage.release();
    return SLLB_OK;
}

/* ---------------- batched advection ---------------- */
int sllb_advect_axis(sllb_field_t F, int axis, int method, int order, const sllb_disp_t *disp) {
    if (!F || !disp || !disp->values) return fail(SLLB_ERR_INVALID, "advect_axis: null argument");
    SLLB_TRY(require_device());
    DispDesc dd;
    if (disp->values_on_device) dd.v = disp->values;
    else {
        if (disp->nvalues < 1) return fail(SLLB_ERR_INVALID, "advect_axis: nvalues < 1");
        SLLB_TRY(F->disp_scratch.ensure((size_t)disp->nvalues));
        SLLB_CUDA(cudaMemcpyAsync(F->disp_scratch.p, disp->values, (size_t)disp->nvalues * sizeof(double),
                                  cudaMemcpyHostToDevice, g_stream));
        dd.v = F->disp_scratch.p;
    }
    dd.scale = disp->scale;
    dd.odiv = disp->odiv > 0 ? disp->odiv : 1; dd.omod = disp->omod > 0 ? disp->omod : 1; dd.ostr = disp->ostr;
    dd.idiv = disp->idiv > 0 ? disp->idiv : 1; dd.imod = disp->imod > 0 ? disp->imod : 1; dd.istr = disp->istr;
    return advect_axis_dev(F, axis, method, order, dd);
}
} // extern "C"
namespace sllb {
int make_affine_disp(sllb_field *F, int axis, int v_axis, double vmin, double dv, double scale, DispDesc *dd) {
    if (!F || v_axis < 0 || v_axis >= F->ndim || v_axis == axis)
        return fail(SLLB_ERR_INVALID, "advect_axis_affine: bad v_axis");
    SLLB_TRY(require_device());
    const int nvv = F->ext[v_axis];
    SLLB_TRY(F->disp_scratch.ensure((size_t)nvv));
    SLLB_CUDA(launch_affine(F->disp_scratch.p, nvv, vmin, dv, g_stream));
    dd->v = F->disp_scratch.p; dd->scale = scale;
    dd->odiv = dd->omod = dd->idiv = dd->imod = 1; dd->ostr = dd->istr = 0;
    long long stride = 1;
    if (v_axis > axis) {
        for (int d = axis + 1; d < v_axis; ++d) stride *= F->ext[d];
        dd->odiv = stride; dd->omod = nvv; dd->ostr = 1;
    } else {
        for (int d = 0; d < v_axis; ++d) stride *= F->ext[d];
        dd->idiv = stride; dd->imod = nvv; dd->istr = 1;
    }
    return SLLB_OK;
}
int make_field_disp(sllb_field *F, int axis, const double *d_field, int nfield_axes, double scale, DispDesc *dd) {
    if (!F || !d_field || nfield_axes < 1 || nfield_axes > axis)
        return fail(SLLB_ERR_INVALID, "advect_axis_field: field axes must be faster than the advected axis");
    SLLB_TRY(require_device());
    long long nf = 1;
    for (int d = 0; d < nfield_axes; ++d) nf *= F->ext[d];
    dd->v = d_field; dd->scale = scale;
    dd->odiv = dd->omod = 1; dd->ostr = 0;
    dd->idiv = 1; dd->imod = nf; dd->istr = 1;
    return SLLB_OK;
}
} // namespace sllb
extern "C" {
int sllb_advect_axis_affine(sllb_field_t F, int axis, int method, int order, int v_axis, double vmin, double dv,
                            double scale) {
    DispDesc dd;
    SLLB_TRY(make_affine_disp(F, axis, v_axis, vmin, dv, scale, &dd));
    return advect_axis_dev(F, axis, method, order, dd);
}
int sllb_advect_axis_field(sllb_field_t F, int axis, int method, int order, const double *d_field, int nfield_axes,
                           double scale) {
    DispDesc dd;
    SLLB_TRY(make_field_disp(F, axis, d_field, nfield_axes, scale, &dd));
    return advect_axis_dev(F, axis, method, order, dd);
}

} // extern "C"
namespace sllb {
int to_dispdesc(const sllb_disp_t *disp, DevBuf &scratch, DispDesc *dd) {
    if (!disp || !disp->values) return fail(SLLB_ERR_INVALID, "displacement: null");
    if (disp->values_on_device) dd->v = disp->values;
    else {
        if (disp->nvalues < 1) return fail(SLLB_ERR_INVALID, "displacement: nvalues < 1");
        SLLB_TRY(scratch.ensure((size_t)disp->nvalues));
        SLLB_CUDA(cudaMemcpyAsync(scratch.p, disp->values, (size_t)disp->nvalues * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        dd->v = scratch.p;
    }
    dd->scale = disp->scale;
    dd->odiv = disp->odiv > 0 ? disp->odiv : 1; dd->omod = disp->omod > 0 ? disp->omod : 1; dd->ostr = disp->ostr;
    dd->idiv = disp->idiv > 0 ? disp->idiv : 1; dd->imod = disp->imod > 0 ? disp->imod : 1; dd->istr = disp->istr;
    return SLLB_OK;
}
// integer shift table of the local-spline advector: HOST int32[n] -> device (kept in F's scratch); NULL stays NULL
int upload_shift(sllb_field *F, const int32_t *shift, long long n, const int **d_shift) {
    *d_shift = nullptr;
    if (!shift) return SLLB_OK;
    if (n < 1) return fail(SLLB_ERR_INVALID, "shift table: nvalues < 1");
    SLLB_TRY(F->shift_scratch.ensure((size_t)(n + 1) / 2));
    SLLB_CUDA(cudaMemcpyAsync(F->shift_scratch.p, shift, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, g_stream));
    *d_shift = reinterpret_cast<const int *>(F->shift_scratch.p);
    return SLLB_OK;
}
} // namespace sllb
extern "C" {
/* make_blocks_spline (sll_m_advection_6d_spline_dd_slim.F90:202-287), host only: the reference walks the monotonic
 * displacement array once and cuts it into blocks of equal integer part; indices with abs(disp) == 0 belong to no
 * block and stay untouched.  The walk is restated literally (including which block an exactly integer displacement
 * lands in) so that the per-index shift table drives the kernels exactly like the reference's block loop. */
int sllb_spline_dd_blocks(int n, const double *disp, int32_t *shift, double *alpha, int *nblocks) {
    if (n < 1 || !disp || !shift) return fail(SLLB_ERR_INVALID, "spline_dd_blocks: bad arguments");
    for (int j = 0; j < n; ++j) shift[j] = SLLB_SHIFT_SKIP;
    auto D = [&](int j) { return disp[j - 1]; }; // 1-based like the reference
    int bl = 1;
    if (n > 1 && fabs(D(bl)) == 0.0) bl = bl + 1;
    const int box1 = (int)floor(D(bl));
    bl = n;
    if (n > 1 && fabs(D(bl)) == 0.0) bl = bl - 1;
    const int box2 = (int)floor(D(bl));
    const int blocks = abs(box2 - box1) + 1;
    int j = 1;
    for (int b = 1; b <= blocks; ++b) {
        if (j <= n && fabs(D(j)) == 0.0) j = j + 1;
        const int si = (box1 > box2) ? box1 - b + 1 : box1 + b - 1;
        const int jstart = j;
        if (box1 > box2) { while (j <= n && D(j) > (double)si) ++j; }
        else { while (j <= n && D(j) < (double)(si + 1)) ++j; }
        const int jend = (j - 1 >= 1 && fabs(D(j - 1)) == 0.0) ? j - 2 : j - 1;
        for (int k = jstart; k <= jend; ++k) shift[k - 1] = si;
    }
    if (alpha) for (int k = 0; k < n; ++k) alpha[k] = disp[k] - floor(disp[k]);
    if (nblocks) *nblocks = blocks;
    return SLLB_OK;
}
int sllb_lagrange_dd_blocks(int n, int stencil, const double *disp, int32_t *box, int *nblocks, int *halo_width) {
    if (n < 1 || !disp || !box || stencil < 2 || stencil % 2 != 0) return fail(SLLB_ERR_INVALID, "lagrange_dd_blocks: bad arguments");
    int nb = 0;
    SLLB_TRY(sllb_spline_dd_blocks(n, disp, box, nullptr, &nb)); // same walk (:237-283 == spline :219-277)
    int first = 1, last = n;
    if (n > 1 && fabs(disp[0]) == 0.0) first = 2;
    if (n > 1 && fabs(disp[n - 1]) == 0.0) last = n - 1;
    const int box1 = (int)floor(disp[first - 1]), box2 = (int)floor(disp[last - 1]);
    for (int b : {box1, box2})
        if (b < -stencil / 2 || b >= stencil / 2)
            return fail(SLLB_ERR_INVALID, "lagrange_dd_blocks: displacement too large for the stencil (box outside [-stencil/2, stencil/2))");
    if (halo_width)
        for (int b = 0; b < nb; ++b) {
            const int bx = (box1 > box2) ? box1 - b : box1 + b;
            halo_width[2 * b] = stencil / 2 - bx - 1;
            halo_width[2 * b + 1] = stencil / 2 + bx;
        }
    if (nblocks) *nblocks = nb;
    return SLLB_OK;
}
/* local cubic spline (NUM_TERMS = 15) along an axis that is NOT split: the ring neighbour is the line itself */
int sllb_advect_axis_spline_dd(sllb_field_t F, int axis, const sllb_disp_t *disp, const int32_t *shift) {
    if (!F || axis < 0 || axis >= F->ndim) return fail(SLLB_ERR_INVALID, "advect_axis_spline_dd: bad arguments");
    SLLB_TRY(require_device());
    DispDesc dd;
    SLLB_TRY(to_dispdesc(disp, F->disp_scratch, &dd));
    const int *d_shift = nullptr;
    SLLB_TRY(upload_shift(F, shift, disp->nvalues, &d_shift));
    long long inner = 1, outer = 1;
    for (int d = 0; d < axis; ++d) inner *= F->ext[d];
    for (int d = axis + 1; d < F->ndim; ++d) outer *= F->ext[d];
    cudaError_t e = launch_spline_dd(F->d, outer, F->ext[axis], inner, dd, d_shift, nullptr, 0, nullptr, 0, nullptr, nullptr,
                                     g_staging, g_stream);
    if (e == cudaErrorInvalidValue) {
        cudaGetLastError();
        return fail(SLLB_ERR_UNSUPPORTED, "advect_axis_spline_dd: the local spline needs more than 15 points per line "
                                          "(SLL_ASSERT_ALWAYS(num_points > NUM_TERMS), sll_m_cubic_spline_halo_1d.F90:79)");
    }
    return check_cuda(e, "k_spline_dd launch");
}
int sllb_advect_plane(sllb_field_t F, int method, int order, const sllb_disp_t *disp0, const sllb_disp_t *disp1,
                      double rho_scale, double *d_rho) {
    if (!F) return fail(SLLB_ERR_INVALID, "advect_plane: null field");
    SLLB_TRY(require_device());
    DispDesc d0, d1;
    SLLB_TRY(to_dispdesc(disp0, F->disp_scratch, &d0));
    SLLB_TRY(to_dispdesc(disp1, F->disp_scratch2, &d1));
    if (method == SLLB_METHOD_LAGRANGE_FIXED || method == SLLB_METHOD_LAGRANGE_CENTERED) {
        if (d_rho) return fail(SLLB_ERR_UNSUPPORTED, "advect_plane: the charge density comes with the spline plane kernel only");
        return advect_lagrange_plane_dev(F, method, order, d0, d1);
    }
    if (method != SLLB_METHOD_SPLINE || order != 4) return fail(SLLB_ERR_UNSUPPORTED, "advect_plane: cubic splines (order 4) or Lagrange stencils");
    return advect_plane_dev(F, d0, d1, rho_scale, d_rho);
}

/* ---------------- reductions ---------------- */
int sllb_reduce_velocity(sllb_field_t F, int nx_axes, double scale, double *d_rho) {
    if (!F || !d_rho || nx_axes < 1 || nx_axes >= F->ndim) return fail(SLLB_ERR_INVALID, "reduce_velocity: bad arguments");
    SLLB_TRY(require_device());
    long long nx = 1, nv = 1;
    for (int d = 0; d < nx_axes; ++d) nx *= F->ext[d];
    for (int d = nx_axes; d < F->ndim; ++d) nv *= F->ext[d];
    SLLB_TRY(F->red_scratch.ensure(reduce_scratch_doubles(nx, nv)));
    SLLB_CUDA(launch_reduce_velocity(F->d, nx, nv, scale, d_rho, F->red_scratch.p, g_stream));
    return SLLB_OK;
}
int sllb_reduce_velocity_host(sllb_field_t F, int nx_axes, double scale, double *h_rho) {
    if (!F || !h_rho || nx_axes < 1 || nx_axes >= F->ndim) return fail(SLLB_ERR_INVALID, "reduce_velocity: bad arguments");
    long long nx = 1;
    for (int d = 0; d < nx_axes; ++d) nx *= F->ext[d];
    SLLB_TRY(require_device());
    SLLB_TRY(F->rho_scratch.ensure((size_t)nx));
    SLLB_TRY(sllb_reduce_velocity(F, nx_axes, scale, F->rho_scratch.p));
    SLLB_CUDA(cudaMemcpy(h_rho, F->rho_scratch.p, (size_t)nx * sizeof(double), cudaMemcpyDeviceToHost));
    return SLLB_OK;
}
int sllb_moments(sllb_field_t F, int nv, const double *w1, const double *w2, double *out) {
    SLLB_TRY(require_device());
    return moments_local(F, nv, w1, w2, out);
}

/* ---------------- Poisson ---------------- */
int sllb_set_poisson_direct(int on) {
    g_poisson_direct = on ? 1 : 0;
    return SLLB_OK;
}
static int poisson_common(sllb_poisson *p) {
    p->nreal = (long long)p->n[0] * p->n[1] * p->n[2];
    p->ncplx = (long long)(p->n[0] / 2 + 1) * p->n[1] * p->n[2];
    SLLB_CUDA(cudaMalloc(&p->rho_hat, (size_t)p->ncplx * sizeof(cufftDoubleComplex)));
    for (int k = 0; k < 4; ++k) SLLB_CUDA(cudaMalloc(&p->spec[k], (size_t)p->ncplx * sizeof(cufftDoubleComplex)));
    if (p->dim == 1) {
        SLLB_CUFFT(cufftPlan1d(&p->fwd, p->n[0], CUFFT_D2Z, 1));
        SLLB_CUFFT(cufftPlan1d(&p->bwd, p->n[0], CUFFT_Z2D, 1));
    } else if (p->dim == 2) {
        SLLB_CUFFT(cufftPlan2d(&p->fwd, p->n[1], p->n[0], CUFFT_D2Z));
        SLLB_CUFFT(cufftPlan2d(&p->bwd, p->n[1], p->n[0], CUFFT_Z2D));
    } else {
        SLLB_CUFFT(cufftPlan3d(&p->fwd, p->n[2], p->n[1], p->n[0], CUFFT_D2Z));
        SLLB_CUFFT(cufftPlan3d(&p->bwd, p->n[2], p->n[1], p->n[0], CUFFT_Z2D));
    }
    p->plans = true;
    SLLB_CUFFT(cufftSetStream(p->fwd, 0));
    SLLB_CUFFT(cufftSetStream(p->bwd, 0));
    return SLLB_OK;
}
int sllb_poisson1d_create(int nc, double xmin, double xmax, sllb_poisson_t *P) {
    if (!P || nc < 4 || !(xmax > xmin)) return fail(SLLB_ERR_INVALID, "poisson1d_create: bad arguments");
    SLLB_TRY(require_device());
    sllb_poisson *p = new sllb_poisson();
    p->dim = 1; p->n[0] = nc; p->L[0] = xmax - xmin; p->xmin[0] = xmin;
    int rc = poisson_common(p);
    if (rc) { sllb_poisson_destroy(p); return rc; }
    *P = p;
    return SLLB_OK;
}
int sllb_poisson2d_create(int nc_x, int nc_y, double x_min, double x_max, double y_min, double y_max,
                          sllb_poisson_t *P) {
    if (!P || nc_x < 4 || nc_y < 4 || !(x_max > x_min) || !(y_max > y_min))
        return fail(SLLB_ERR_INVALID, "poisson2d_create: bad arguments");
    SLLB_TRY(require_device());
    sllb_poisson *p = new sllb_poisson();
    p->dim = 2; p->n[0] = nc_x; p->n[1] = nc_y; p->L[0] = x_max - x_min; p->L[1] = y_max - y_min;
    int rc = poisson_common(p);
    // small grids: the three-kernel dense-DFT solve (a 128 x 128 problem is launch-bound as library FFT calls)
    if (!rc && poisson2d_direct_supported(nc_x, nc_y)) rc = poisson2d_direct_create(nc_x, nc_y, p->L[0], p->L[1], &p->direct);
    if (rc) { sllb_poisson_destroy(p); return rc; }
    *P = p;
    return SLLB_OK;
}
/* sll_t_poisson_2d_periodic_par (sll_m_poisson_2d_periodic_par.F90:120-338): solves Delta phi = rho (the caller gives the
 * source its sign) on [0,Lx] x [0,Ly], zero-mean phi, no field outputs; replicated on one device */
int sllb_poisson2d_par_create(int ncx, int ncy, double Lx, double Ly, sllb_poisson_t *P) {
    SLLB_TRY(sllb_poisson2d_create(ncx, ncy, 0.0, Lx, 0.0, Ly, P));
    (*P)->par_variant = 1;
    return SLLB_OK;
}
int sllb_poisson3d_create(int nx, int ny, int nz, double Lx, double Ly, double Lz, sllb_poisson_t *P) {
    if (!P || nx < 4 || ny < 4 || nz < 4 || !(Lx > 0) || !(Ly > 0) || !(Lz > 0))
        return fail(SLLB_ERR_INVALID, "poisson3d_create: bad arguments");
    SLLB_TRY(require_device());
    sllb_poisson *p = new sllb_poisson();
    p->dim = 3; p->n[0] = nx; p->n[1] = ny; p->n[2] = nz; p->L[0] = Lx; p->L[1] = Ly; p->L[2] = Lz;
    int rc = poisson_common(p);
    if (rc) { sllb_poisson_destroy(p); return rc; }
    *P = p;
    return SLLB_OK;
}
int sllb_poisson_destroy(sllb_poisson_t P) {
    if (!P) return SLLB_OK;
    if (P->plans) { cufftDestroy(P->fwd); cufftDestroy(P->bwd); }
    poisson2d_direct_destroy(P->direct);
    if (P->rho_hat) cudaFree(P->rho_hat);
    for (int k = 0; k < 4; ++k) if (P->spec[k]) cudaFree(P->spec[k]);
    delete P;
    return SLLB_OK;
}
int sllb_poisson_solve(sllb_poisson_t P, const double *d_rho, double *d_phi, double *d_e1, double *d_e2, double *d_e3) {
    if (!P || !d_rho) return fail(SLLB_ERR_INVALID, "poisson_solve: null");
    SLLB_TRY(require_device());
    if (P->dim == 2 && P->par_variant && (d_e1 || d_e2 || !d_phi))
        return fail(SLLB_ERR_INVALID, "poisson_solve: sll_t_poisson_2d_periodic_par returns the potential only (phi must be given, e1/e2 NULL)");
    if (P->dim == 2 && P->direct && g_poisson_direct) {
        SLLB_CUDA(poisson2d_direct_solve(P->direct, d_rho, 1, 0, 1.0, nullptr, P->par_variant, d_phi, d_e1, d_e2, nullptr, nullptr,
                                         nullptr, nullptr, g_stream));
        return SLLB_OK;
    }
    if (P->stream != g_stream) {
        SLLB_CUFFT(cufftSetStream(P->fwd, g_stream));
        SLLB_CUFFT(cufftSetStream(P->bwd, g_stream));
        P->stream = g_stream;
    }
    SLLB_CUFFT(cufftExecD2Z(P->fwd, const_cast<double *>(d_rho), P->rho_hat));
    cufftDoubleComplex *ph = d_phi ? P->spec[0] : nullptr, *e1 = d_e1 ? P->spec[1] : nullptr;
    cufftDoubleComplex *e2 = d_e2 ? P->spec[2] : nullptr, *e3 = d_e3 ? P->spec[3] : nullptr;
    if (P->dim == 1) {
        if (d_phi) return fail(SLLB_ERR_UNSUPPORTED, "poisson1d: potential output not implemented (the reference returns E only)");
        if (e1) SLLB_CUDA(launch_poisson1d_mult(P->rho_hat, P->n[0], P->L[0], e1, g_stream));
        e2 = e3 = nullptr;
    } else if (P->dim == 2 && P->par_variant) {
        SLLB_CUDA(launch_poisson2d_par_mult(P->rho_hat, P->n[0], P->n[1], P->L[0], P->L[1], ph, g_stream));
        e1 = e2 = e3 = nullptr;
    } else if (P->dim == 2) {
        SLLB_CUDA(launch_poisson2d_mult(P->rho_hat, P->n[0], P->n[1], P->L[0], P->L[1], ph, e1, e2, g_stream));
        e3 = nullptr;
    } else {
        SLLB_CUDA(launch_poisson3d_mult(P->rho_hat, P->n[0], P->n[1], P->n[2], P->L[0], P->L[1], P->L[2], ph, e1, e2, e3, g_stream));
    }
    if (ph) SLLB_CUFFT(cufftExecZ2D(P->bwd, ph, d_phi));
    if (e1) SLLB_CUFFT(cufftExecZ2D(P->bwd, e1, d_e1));
    if (e2) SLLB_CUFFT(cufftExecZ2D(P->bwd, e2, d_e2));
    if (e3) SLLB_CUFFT(cufftExecZ2D(P->bwd, e3, d_e3));
    return SLLB_OK;
}
int sllb_poisson_solve_host(sllb_poisson_t P, const double *rho, const int *ld, double *phi, double *e1, double *e2,
                            double *e3) {
    if (!P || !rho || !ld) return fail(SLLB_ERR_INVALID, "poisson_solve_host: null");
    SLLB_TRY(require_device());
    int l[3] = {1, 1, 1};
    for (int d = 0; d < P->dim; ++d) {
        l[d] = ld[d];
        if (l[d] != P->n[d] && l[d] != P->n[d] + 1) return fail(SLLB_ERR_INVALID, "poisson_solve_host: ld must be nc or nc+1");
    }
    std::vector<double> h((size_t)P->nreal);
    for (int k = 0; k < P->n[2]; ++k) for (int j = 0; j < P->n[1]; ++j) for (int i = 0; i < P->n[0]; ++i)
        h[i + (size_t)P->n[0] * (j + (size_t)P->n[1] * k)] = rho[i + (size_t)l[0] * (j + (size_t)l[1] * k)];
    SLLB_TRY(P->rho_in.ensure((size_t)P->nreal));
    SLLB_CUDA(cudaMemcpy(P->rho_in.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    double *outs[4] = {phi, e1, P->dim >= 2 ? e2 : nullptr, P->dim >= 3 ? e3 : nullptr};
    double *douts[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; ++k) if (outs[k]) { SLLB_TRY(P->out[k].ensure((size_t)P->nreal)); douts[k] = P->out[k].p; }
    SLLB_TRY(sllb_poisson_solve(P, P->rho_in.p, douts[0], douts[1], douts[2], douts[3]));
    for (int q = 0; q < 4; ++q) {
        if (!outs[q]) continue;
        SLLB_CUDA(cudaMemcpy(h.data(), douts[q], h.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int k = 0; k < l[2]; ++k) for (int j = 0; j < l[1]; ++j) for (int i = 0; i < l[0]; ++i)
            outs[q][i + (size_t)l[0] * (j + (size_t)l[1] * k)] =
                h[(i % P->n[0]) + (size_t)P->n[0] * ((j % P->n[1]) + (size_t)P->n[1] * (k % P->n[2]))];
    }
    return SLLB_OK;
}

/* ---------------- line-granular drop-in handles ---------------- */
} // extern "C"

struct sllb_adv1d {
    int kind, num_cells, order;
    double xmin, xmax;
    int method, stencil;
    sllb_field *line = nullptr;
};
struct sllb_interp1d {
    int kind, num_points, num_cells, periodic_last;
    double xmin, xmax, delta;
    int method, stencil;
    sllb_field *line = nullptr;
    int bc = SLLB_BC_PERIODIC;
    int have_slopes = 0;
    double slope_left = 0.0, slope_right = 0.0;
};
static int hermite_line(sllb_interp1d *h, const double *in, double *out, double alpha, int inplace) {
    const int N = h->num_points;
    SLLB_CUDA(cudaMemcpy(h->line->d, in, (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
    SLLB_TRY(h->line->disp_scratch.ensure(1));
    SLLB_CUDA(cudaMemcpy(h->line->disp_scratch.p, &alpha, sizeof(double), cudaMemcpyHostToDevice));
    DispDesc dd;
    dd.v = h->line->disp_scratch.p; dd.scale = 1.0;
    dd.odiv = dd.omod = dd.idiv = dd.imod = 1; dd.ostr = dd.istr = 0;
    SLLB_CUDA(launch_hermite(h->line->d, 1, N, 1, dd, h->delta, inplace, h->have_slopes, h->slope_left, h->slope_right, g_staging, 0));
    SLLB_CUDA(cudaMemcpy(out, h->line->d, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
    return SLLB_OK;
}

static int line_shift(sllb_field *line, int method, int stencil, double disp_cells, const double *in, double *out, int n) {
    const int N = line->ext[0];
    SLLB_CUDA(cudaMemcpy(line->d, in, (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
    sllb_disp_t d;
    memset(&d, 0, sizeof(d));
    d.values = &disp_cells; d.nvalues = 1; d.values_on_device = 0; d.scale = 1.0;
    d.odiv = d.omod = d.idiv = d.imod = 1;
    SLLB_TRY(sllb_advect_axis(line, 0, method, stencil, &d));
    SLLB_CUDA(cudaMemcpy(out, line->d, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
    if (n > N) out[N] = out[0];
    return SLLB_OK;
}

extern "C" {

int sllb_adv1d_create(int kind, int num_cells, double xmin, double xmax, int order, sllb_adv1d_t *h) {
    if (!h || num_cells < 8 || !(xmax > xmin)) return fail(SLLB_ERR_INVALID, "adv1d_create: bad arguments");
    int method, stencil;
    if (kind == SLLB_ADV_PERIODIC_SPLINE) {
        if (order != 4 && order != 6 && order != 8) return fail(SLLB_ERR_UNSUPPORTED, "adv1d_create: sll_p_spline implemented for orders 4, 6 and 8");
        method = SLLB_METHOD_SPLINE; stencil = order;
    } else if (kind == SLLB_ADV_PERIODIC_LAGRANGE) {
        if (order < 4 || order > 18 || order % 2 != 0)
            return fail(SLLB_ERR_UNSUPPORTED, "adv1d_create: sll_p_lagrange implemented for even orders 4 .. 18");
        method = SLLB_METHOD_LAGRANGE_CENTERED; stencil = order;
    } else if (kind == SLLB_ADV_BSL) {
        method = SLLB_METHOD_SPLINE; stencil = 4;
    } else return fail(SLLB_ERR_UNSUPPORTED, "adv1d_create: advector kind not implemented");
    SLLB_TRY(require_device());
    sllb_adv1d *a = new sllb_adv1d();
    a->kind = kind; a->num_cells = num_cells; a->order = order; a->xmin = xmin; a->xmax = xmax;
    a->method = method; a->stencil = stencil;
    int rc = field_alloc(1, &num_cells, &a->line);
    if (rc) { delete a; return rc; }
    *h = a;
    return SLLB_OK;
}
int sllb_adv1d_advect_constant(sllb_adv1d_t h, double A, double dt, const double *in, double *out, int n) {
    if (!h || !in || !out) return fail(SLLB_ERR_INVALID, "advect_constant: null");
    if (n != h->num_cells && n != h->num_cells + 1) return fail(SLLB_ERR_INVALID, "advect_constant: n must be num_cells or num_cells+1");
    SLLB_TRY(require_device());
    /* shift = A*dt/(xmax-xmin)*num_cells (sll_m_advection_1d_periodic.F90:117); out(j) = interp(j - shift) */
    if (h->kind == SLLB_ADV_BSL) {
        /* feet = eta_i - dt*A folded into [eta_min, eta_max) (explicit Euler, periodic), then the interpolant is
         * evaluated there: out(i) = S(x_i - A*dt), displacement in cells = -A*dt/delta */
        const double delta = (h->xmax - h->xmin) / (double)h->num_cells;
        return line_shift(h->line, h->method, h->stencil, -(A * dt) / delta, in, out, n);
    }
    const double shift = A * dt / (h->xmax - h->xmin) * (double)h->num_cells;
    return line_shift(h->line, h->method, h->stencil, -shift, in, out, n);
}
int sllb_adv1d_delete(sllb_adv1d_t h) {
    if (!h) return SLLB_OK;
    sllb_field_destroy(h->line);
    delete h;
    return SLLB_OK;
}

int sllb_interp1d_create(int kind, int num_points, double xmin, double xmax, int bc, int d_or_order, int periodic_last,
                         int fast_algorithm, sllb_interp1d_t *h) {
    if (!h || !(xmax > xmin)) return fail(SLLB_ERR_INVALID, "interp1d_create: bad arguments");
    if (bc == SLLB_BC_HERMITE) {
        // sll_t_cubic_spline_interpolator_1d with sll_p_hermite: num_points grid points, no periodic duplicate
        if (kind != SLLB_INTERP_CUBIC_SPLINE) return fail(SLLB_ERR_UNSUPPORTED, "interp1d_create: sll_p_hermite is a cubic-spline boundary condition");
        if (num_points < 27 || !fast_algorithm)
            return fail(SLLB_ERR_UNSUPPORTED, "interp1d_create: Hermite splines are implemented with the fast algorithm only "
                                              "(num_points >= 27, sll_m_cubic_splines.F90:266-272)");
        SLLB_TRY(require_device());
        sllb_interp1d *p = new sllb_interp1d();
        p->kind = kind; p->num_points = num_points; p->num_cells = num_points - 1; p->periodic_last = 0;
        p->xmin = xmin; p->xmax = xmax; p->delta = (xmax - xmin) / (double)(num_points - 1);
        p->method = -1; p->stencil = 4; p->bc = bc;
        int rc = field_alloc(1, &num_points, &p->line);
        if (rc) { delete p; return rc; }
        *h = p;
        return SLLB_OK;
    }
    if (bc != SLLB_BC_PERIODIC) return fail(SLLB_ERR_UNSUPPORTED, "interp1d_create: boundary condition not implemented (sll_p_periodic, sll_p_hermite)");
    int method, stencil, ncell;
    switch (kind) {
    case SLLB_INTERP_CUBIC_SPLINE: method = SLLB_METHOD_SPLINE; stencil = 4; periodic_last = 1; break;
    case SLLB_INTERP_PERIODIC_SPLINE:
        if (d_or_order != 4 && d_or_order != 6 && d_or_order != 8) return fail(SLLB_ERR_UNSUPPORTED, "interp1d_create: periodic spline implemented for orders 4, 6 and 8");
        method = SLLB_METHOD_SPLINE; stencil = d_or_order; periodic_last = 1; break;
    case SLLB_INTERP_PERIODIC_LAGRANGE: method = SLLB_METHOD_LAGRANGE_CENTERED; stencil = d_or_order; periodic_last = 1; break;
    case SLLB_INTERP_LAGRANGE_CENTERED: method = SLLB_METHOD_LAGRANGE_CENTERED; stencil = 2 * d_or_order; break;
    case SLLB_INTERP_LAGRANGE_FIXED: method = SLLB_METHOD_LAGRANGE_FIXED; stencil = 2 * d_or_order + 1; break;
    default: return fail(SLLB_ERR_UNSUPPORTED, "interp1d_create: interpolator kind not implemented");
    }
    if (method == SLLB_METHOD_LAGRANGE_CENTERED && (stencil < 4 || stencil > 18 || stencil % 2 != 0))
        return fail(SLLB_ERR_UNSUPPORTED, "interp1d_create: centred Lagrange implemented for even stencils 4 .. 18");
    if (method == SLLB_METHOD_LAGRANGE_FIXED && (stencil < 3 || stencil > 11))
        return fail(SLLB_ERR_UNSUPPORTED, "interp1d_create: fixed Lagrange implemented for stencils 3..11");
    /* both interpolators define the cell size from num_points-1 cells
     * (sll_m_cubic_splines.F90:262, sll_m_lagrange_interpolation_1d.F90:58-60) */
    ncell = num_points - 1;
    if (ncell < 8) return fail(SLLB_ERR_INVALID, "interp1d_create: too few points");
    SLLB_TRY(require_device());
    sllb_interp1d *p = new sllb_interp1d();
    p->kind = kind; p->num_points = num_points; p->num_cells = ncell; p->periodic_last = periodic_last ? 1 : 0;
    p->xmin = xmin; p->xmax = xmax; p->delta = (xmax - xmin) / (double)ncell;
    p->method = method; p->stencil = stencil;
    int rc = field_alloc(1, &ncell, &p->line);
    if (rc) { delete p; return rc; }
    *h = p;
    return SLLB_OK;
}
int sllb_interp1d_array_disp(sllb_interp1d_t h, int n, const double *data, double alpha, double *out) {
    if (!h || !data || !out) return fail(SLLB_ERR_INVALID, "interpolate_array_disp: null");
    if (h->bc == SLLB_BC_HERMITE) {
        if (n != h->num_points) return fail(SLLB_ERR_INVALID, "interpolate_array_disp: bad num_pts");
        SLLB_TRY(require_device());
        return hermite_line(h, data, out, alpha, 0);
    }
    if (n != h->num_cells && n != h->num_cells + 1) return fail(SLLB_ERR_INVALID, "interpolate_array_disp: bad num_pts");
    SLLB_TRY(require_device());
    return line_shift(h->line, h->method, h->stencil, alpha / h->delta, data, out, n);
}
int sllb_interp1d_array_disp_inplace(sllb_interp1d_t h, int n, double *data, double alpha) {
    if (h && data && h->bc == SLLB_BC_HERMITE) { // clamped feet + eval_array (sll_m_cubic_spline_interpolator_1d.F90:165-178)
        if (n != h->num_points) return fail(SLLB_ERR_INVALID, "interpolate_array_disp_inplace: bad num_pts");
        SLLB_TRY(require_device());
        return hermite_line(h, data, data, alpha, 1);
    }
    return sllb_interp1d_array_disp(h, n, data, alpha, data);
}
/* the optional slope_left / slope_right of sll_t_cubic_spline_interpolator_1d%init (:323-366); without them the
 * slopes come from 5-point one-sided differences of the data (sll_m_cubic_splines.F90:720-730) */
int sllb_interp1d_set_slopes(sllb_interp1d_t h, double slope_left, double slope_right) {
    if (!h || h->bc != SLLB_BC_HERMITE) return fail(SLLB_ERR_INVALID, "interp1d_set_slopes: not a Hermite spline interpolator");
    h->have_slopes = 1; h->slope_left = slope_left; h->slope_right = slope_right;
    return SLLB_OK;
}
/* batched: every line of F along `axis` (np = extents[axis] grid points on [xmin, xmax], NOT periodic) becomes
 * S(x_i + alpha), alpha = the displacement of the line in physical units */
int sllb_advect_axis_hermite(sllb_field_t F, int axis, double xmin, double xmax, const sllb_disp_t *disp, int inplace_semantics) {
    if (!F || axis < 0 || axis >= F->ndim || !(xmax > xmin)) return fail(SLLB_ERR_INVALID, "advect_axis_hermite: bad arguments");
    SLLB_TRY(require_device());
    DispDesc dd;
    SLLB_TRY(to_dispdesc(disp, F->disp_scratch, &dd));
    long long inner = 1, outer = 1;
    for (int d = 0; d < axis; ++d) inner *= F->ext[d];
    for (int d = axis + 1; d < F->ndim; ++d) outer *= F->ext[d];
    const int np = F->ext[axis];
    cudaError_t e = launch_hermite(F->d, outer, np, inner, dd, (xmax - xmin) / (double)(np - 1), inplace_semantics ? 1 : 0, 0, 0.0, 0.0,
                                   g_staging, g_stream);
    if (e == cudaErrorInvalidValue) {
        cudaGetLastError();
        return fail(SLLB_ERR_UNSUPPORTED, "advect_axis_hermite: Hermite splines need >= 27 points per line (fast algorithm)");
    }
    return check_cuda(e, "k_hermite launch");
}
int sllb_interp1d_delete(sllb_interp1d_t h) {
    if (!h) return SLLB_OK;
    sllb_field_destroy(h->line);
    delete h;
    return SLLB_OK;
}

} // extern "C"
