// Internal structures shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <nccl.h>
#include <string>
#include <vector>

#include "../../include/sll_b200.h"
#include "sllb_kernels.cuh"

namespace sllb {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
int check_cuda(cudaError_t e, const char *what);
int check_cufft(cufftResult r, const char *what);
int require_device();
extern int g_staging;
extern thread_local cudaStream_t g_stream;

#define SLLB_CUDA(call)                                  \
    do {                                                 \
        int _rc = sllb::check_cuda((call), #call);       \
        if (_rc) return _rc;                             \
    } while (0)
#define SLLB_CUFFT(call)                                 \
    do {                                                 \
        int _rc = sllb::check_cufft((call), #call);      \
        if (_rc) return _rc;                             \
    } while (0)
#define SLLB_TRY(call)                                   \
    do {                                                 \
        int _rc = (call);                                \
        if (_rc) return _rc;                             \
    } while (0)

int check_nccl(ncclResult_t r, const char *what);
#define SLLB_NCCL(call)                                  \
    do {                                                 \
        int _rc = sllb::check_nccl((call), #call);       \
        if (_rc) return _rc;                             \
    } while (0)

struct Ext6 { int e[6]; };

// simple owning device buffer
struct DevBuf {
    double *p = nullptr;
    size_t n = 0;
    int ensure(size_t count);
    void release();
    ~DevBuf() { release(); }
};

} // namespace sllb

struct sllb_field {
    int ndim = 0;
    int ext[6] = {1, 1, 1, 1, 1, 1};
    long long total = 0;
    double *d = nullptr;
    bool owns = true;
    sllb::DevBuf disp_scratch;   // uploaded displacement values
    sllb::DevBuf disp_scratch2;  // second displacement (plane kernel)
    sllb::DevBuf rho_scratch;    // device rho of the host-returning reduction
    sllb::DevBuf red_scratch;    // reduction partials
    sllb::DevBuf stage;          // upload/download staging with duplicates
    sllb::DevBuf rows;           // diagnostics row sums
    sllb::DevBuf shift_scratch;  // uploaded integer shift table (local spline)
};

struct sllb_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
};

namespace sllb { struct Poisson2dDirect; }
struct sllb_poisson {
    int par_variant = 0;                       // 2D: 1 = sll_t_poisson_2d_periodic_par (Delta phi = rho, phi only)
    sllb::Poisson2dDirect *direct = nullptr;   // 2D, small grids: three-kernel dense-DFT solve (sllb_poisson_direct.cu)
    int dim = 0;
    int n[3] = {1, 1, 1};
    double L[3] = {1, 1, 1};
    double xmin[3] = {0, 0, 0};
    cufftHandle fwd = 0, bwd = 0;
    bool plans = false;
    cufftDoubleComplex *rho_hat = nullptr, *spec[4] = {nullptr, nullptr, nullptr, nullptr};
    sllb::DevBuf rho_in, out[4];
    long long nreal = 0, ncplx = 0;
    cudaStream_t stream = 0;   // the stream the two plans are currently bound to
};

namespace sllb {
// sllb_poisson_direct.cu: 2D periodic Poisson for N1, N2 <= 256 as three dense-DFT kernels (no library FFT).
// rho = scale * sum of `nslots` arrays spaced by slot_stride (rho_sum != NULL: the summed density is also written there);
// mode 0: sll_t_poisson_2d_periodic (phi, E1, E2), mode 1: sll_s_poisson_2d_periodic_par_solve (Delta phi = rho, phi only);
// tile_box = {lo0, n0, lo1, n1}: E1/E2 of that sub-box additionally go to the dense arrays tile_e1 / tile_e2.
int poisson2d_direct_supported(int n1, int n2);
int poisson2d_direct_create(int n1, int n2, double L1, double L2, Poisson2dDirect **out);
void poisson2d_direct_destroy(Poisson2dDirect *P);
// nrj_rows != NULL (needs e1 and e2): nrj_rows[x2] = sum over x1 of w (E1^2 + E2^2), w = 2 on the x1 = 0 / x2 = 0 lines
// (each counts twice in the reference's sum over the (N1+1)(N2+1) nodes with the periodic duplicates)
cudaError_t poisson2d_direct_solve(Poisson2dDirect *P, const double *rho, int nslots, long long slot_stride, double scale,
                                   double *rho_sum, int mode, double *phi, double *e1, double *e2, double *nrj_rows,
                                   double *tile_e1, double *tile_e2, const int tile_box[4], cudaStream_t st);
// cross-rank barrier through 64-bit flags in peer-mapped memory (sllb_sims.cu): sig[q] = rank q's flag array (8 slots,
// slot r = last epoch rank r published); epochs grow by one per barrier; *err is raised when a peer never arrives
cudaError_t launch_flag_barrier(unsigned long long *const sig[8], int nranks, int rank, unsigned long long epoch, double *err,
                                cudaStream_t st);
extern int g_poisson_direct;   // 1 (default): 2D solves on small grids take the direct path; 0: always cuFFT
int peer_map_buffers(sllb_comm *comm, void *const *mine, int count, std::vector<void *> &peers,
                     std::vector<void *> &opened, bool *ok);
// internal (device-pointer) entry points used by the simulations
int advect_axis_dev(sllb_field *F, int axis, int method, int order, const DispDesc &dd, const RemapDst *remap = nullptr,
                    double *linesum = nullptr, const LineDiag *diag = nullptr, const LineSub *sub = nullptr);
// partials_out != NULL: the per-CTA partial densities are left unsummed ([nparts][n1*n2], unscaled) for the caller
int advect_plane_dev(sllb_field *F, const DispDesc &dd0, const DispDesc &dd1, double rho_scale, double *d_rho,
                     const RemapDst *remap = nullptr, const double **partials_out = nullptr, int *nparts_out = nullptr);
int advect_lagrange_plane_dev(sllb_field *F, int method, int order, const DispDesc &dd0, const DispDesc &dd1);
extern int g_plane_kernel;
int make_affine_disp(sllb_field *F, int axis, int v_axis, double vmin, double dv, double scale, DispDesc *dd);
int make_field_disp(sllb_field *F, int axis, const double *d_field, int nfield_axes, double scale, DispDesc *dd);
int field_alloc(int ndim, const int *ext, sllb_field **F);
int field_wrap(int ndim, const int *ext, double *d, sllb_field **F);
int to_dispdesc(const sllb_disp_t *disp, DevBuf &scratch, DispDesc *dd);
int upload_shift(sllb_field *F, const int32_t *shift, long long n, const int **d_shift);
cudaError_t launch_jacobian2d(const double *e1, const double *e2, int n1, int n2, int r, int s, const double *d_w,
                              double factor, double *jac, cudaStream_t st);
cudaError_t launch_lincomb2(const double *x, const double *y, double a, double b, long long n, double *out, cudaStream_t st);
int moments_local(sllb_field *F, int nv, const double *w1, const double *w2, double *out);
// sllb_diag.cu: time-loop diagnostics of the 2D2V simulation without leaving the device
// out4 = (sum f, sum |f|, sum f^2, sum (w3[i3] + w4[i4]) f) from the per-row sums of K8 (rows3 = [n3*n4][3])
cudaError_t launch_moments_from_rows(const double *rows3, int n3, int n4, const double *w3, const double *w4, double *out4,
                                     cudaStream_t st);
// the same four numbers from the per-line arrays the last x4 pass of a step leaves behind (lines = [nx][n3], x fastest):
// sum f = sum sum_, ..., kinetic = sum_l (w3[i3(l)] sum_[l] + kin[l])
cudaError_t launch_moments_from_lines(const double *sum_, const double *l1, const double *l2, const double *kin, long long nx,
                                      int n3, const double *w3, double *scratch, double *out4, cudaStream_t st);
size_t moments_from_lines_scratch();
// nrj = scale * sum over the (n1+1)(n2+1) nodes including the periodic duplicates of a^2 + (squared ? b^2 : 2 b)
cudaError_t launch_dup_energy2d(const double *a, const double *b, int n1, int n2, double scale, int squared, double *out1,
                                cudaStream_t st);
// 3D3V: density partials + the nine velocity moments of the diagnostics row in one sweep over f ([nv][nx]); wt = [nv][6]
// weights (v4, v5, v6, v4^2, v5^2, v6^2 of every local velocity index)
size_t reduce_moments6d_scratch(long long nx, long long nv);
int reduce_moments6d_chunks(long long nv);   // the density partials are [chunks][nx]
cudaError_t launch_reduce_moments6d(const double *f, long long nx, long long nv, const double *wt, double *partial, int *nchunks_out,
                                    double *mom_part, double *out9, cudaStream_t st);
// dup_velocity_planes mode: side planes <- the cell planes they duplicate; rho += scale * (trapezoid - plain sum)
cudaError_t launch_dup_fill(const double *f, long long n12, int n3, int n4, double *side, cudaStream_t st);
cudaError_t launch_dup_rho_corr(const double *f, const double *side, long long n12, int n3, int n4, double scale, double *rho,
                                cudaStream_t st);
// (sum w f, sum w f^2), w from the global index of every point; scratch: moments_from_lines_scratch() doubles
cudaError_t launch_checksum4d(const double *f, const int ext[4], const int lo[4], double *scratch, double *out2, cudaStream_t st);
// out[k] = w * sum_v |f^_k(v)|^2, k < nmodes (f is [nv][n1], x fastest); part: scratch of nv * nmodes doubles
cudaError_t launch_row_modes(const double *f, int n1, long long nv, int nmodes, double w, double *part, double *out,
                             cudaStream_t st);
cudaError_t launch_absmax(const double *a, long long n, double *out1, cudaStream_t st);
// row6 = (time, nrj[0], 0.5 vol m4[3], vol m4[0], vol m4[1], vol m4[2]); root = 0: time and nrj are written as 0 (the
// rows of several ranks are summed afterwards)
// m4 = sum of nb partial quadruples (fixed order); nrj = nrj_scale * sum of nn values (fixed order);
// counters != NULL: the row goes to rows_base + 6 * counters[0] with time (counters[1] + 1) dt, and both counters advance
cudaError_t launch_sim4d_row(const double *m4_partials, int nb, const double *nrj_parts, int nn, double nrj_scale, double time,
                             double dt, double vol, int root, double *row6, double *rows_base, double *counters, cudaStream_t st);
// first stage of launch_moments_from_lines only: partial[148][4] (moments_from_lines_scratch() doubles)
cudaError_t launch_moments_from_lines_partials(const double *sum_, const double *l1, const double *l2, const double *kin, long long nx,
                                               int n3, const double *w3, double *partial, int *nb_out, cudaStream_t st);
} // namespace sllb
