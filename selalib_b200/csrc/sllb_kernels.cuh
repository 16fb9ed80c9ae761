// Internal interface between the CUDA kernels (sllb_kernels.cu) and the C ABI (sllb_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

namespace sllb {

// Displacement of line (o, in), in cells: scale * v[((o/odiv)%omod)*ostr + ((in/idiv)%imod)*istr]
struct DispDesc {
    const double *v;
    double scale;
    long long odiv, omod, ostr, idiv, imod, istr;
};

enum { METHOD_SPLINE = 0, METHOD_LAGRANGE_FIXED = 1, METHOD_LAGRANGE_CENTERED = 2 };
enum { STAGING_AUTO = 0, STAGING_TMA = 1, STAGING_CPASYNC = 2 };

// Fused remap: destination of a pass that writes into the other layout (see OutMap in sllb_kernels.cu).
struct RemapDst {
    double *base[8];            // destination array on every rank (peer-mapped); in place: base[0] = f
    int on, axis;               // on = 0: in place, everything below ignored
    int se[4], slo[4];          // source layout: local extents and global offset of my box
    int te[4], tp[4];           // destination layout: local extents (uniform boxes) and process mesh
    int block_rot;              // tiles are visited starting from this one (set by the launcher from `rank`): ranks
                                // start at different destination ranks so no receiver is hit by everybody at once
    int rank;                   // my rank (seed of the rotation)
};

// K1/K2: one 1D periodic advection on every line of f viewed as [outer][n][inner], in place.
// Returns cudaSuccess, cudaErrorInvalidValue (bad n / order) or a launch error.
// Per-line moments of the advected line next to its sum (time-loop diagnostics fused into the last pass of a step):
// l1[line] = sum |out|, l2[line] = sum out^2, kin[line] = sum_i w2[i] out(i) with w2 indexed by the OUTPUT point.
struct LineDiag { double *l1, *l2, *kin; const double *w2; };
// A subset of the lines of a pass (chunked V stage: the x3 pass of one chunk of (x1,x2) overlaps the x4 + remap pass of
// the previous one).  The n-th line of the subset is (o, in) = (o_mul * q, i0 + r + in_pitch * q), q = n / icount,
// r = n % icount; icount = 0: all lines.  i0 and icount multiples of 32 keep the 256-byte TMA rows.
struct LineSub { long long icount, i0, in_pitch, nlines; int o_mul; };
cudaError_t launch_advect(double *f, long long outer, int n, long long inner, int method, int order,
                          const DispDesc &dd, int staging, cudaStream_t st, const RemapDst *remap = nullptr,
                          double *linesum = nullptr, const LineDiag *diag = nullptr, const LineSub *sub = nullptr);
// linesum (strided spline passes only, else cudaErrorNotSupported): linesum[line] = sum of the advected line;
// diag needs linesum

// K1c: both passes of a T stage (axes 0 and 1) on every contiguous n1 x n2 plane in one sweep, optionally
// accumulating per-CTA partial sums over the planes: rho_partial[plane_grid(...)][n1*n2].
// cudaErrorNotSupported when the plane shape / displacement pattern does not fit.
int plane_grid(int n1, int n2, long long nplanes);
cudaError_t launch_spline_plane(double *f, int n1, int n2, long long nplanes, const DispDesc &dd1, const DispDesc &dd2,
                                double *rho_partial, cudaStream_t st, const RemapDst *remap = nullptr);
cudaError_t launch_sum_partials(const double *partial, long long nx, int nparts, double scale, double *rho, cudaStream_t st);

// K2d: the axis-0 and the axis-1 Lagrange pass on every contiguous n0 x n1 plane in one sweep (displacements constant
// over a plane).  cudaErrorNotSupported when the plane / displacement pattern / stencil does not fit.
cudaError_t launch_lagrange_plane(double *f, int n0, int n1, long long nplanes, int method, int order, const DispDesc &dd0,
                                  const DispDesc &dd1, cudaStream_t st);
// a sub-box of the lines of an axis pass: o in [o0, o0+ocount), in in [i0, i0+icount) (pipelined halo exchange)
struct LineBox { long long o0, ocount, i0, icount; };
// K2c: fixed odd Lagrange with halo planes (domain-decomposed axis): halo_left/right are [outer][(order-1)/2][inner];
// box != nullptr: only the lines of that sub-box
cudaError_t launch_lagrange_halo(double *f, const double *halo_left, const double *halo_right, long long outer, int n,
                                 long long inner, int order, const DispDesc &dd, int staging, cudaStream_t st,
                                 const LineBox *box = nullptr);
// K9: local cubic spline with halo cells (sll_m_cubic_spline_halo_1d, NUM_TERMS = 15) on every line of f viewed as
// [outer][np][inner], in place.  halo_l == nullptr: the axis is not split, the ring neighbour is the line itself
// (periodic wrap, one kernel).  Otherwise halo_l / halo_r are [outer][hwl|hwr][inner] planes of the neighbours and
// bc_l / bc_r [outer][inner] the neighbours' parts of the two boundary sums (launch_spline_dd_prepare on their side).
// d_shift: optional DEVICE table of integer shifts indexed like dd.v (INT_MIN = leave the line untouched);
// nullptr = floor(displacement).  cudaErrorInvalidValue: too few points (np <= 15, or a neighbour sum would leave
// the neighbour's block), cudaErrorNotSupported: split contiguous axis.
int spline_dd_min_points(int hwl, int hwr);
cudaError_t launch_spline_dd(double *f, long long outer, int np, long long inner, const DispDesc &dd, const int *d_shift,
                             const double *halo_l, int hwl, const double *halo_r, int hwr, const double *bc_l,
                             const double *bc_r, int staging, cudaStream_t st);
// K9p: for_right[line] / for_left[line] = the parts of the right neighbour's d_0 sum / the left neighbour's c_np2 sum
// made of MY cells (sll_s_cubic_spline_halo_1d_prepare_exchange)
cudaError_t launch_spline_dd_prepare(const double *f, long long outer, int np, long long inner, const DispDesc &dd,
                                     const int *d_shift, int hwl, int hwr, double *for_right, double *for_left,
                                     cudaStream_t st);
// K10: cubic spline with Hermite boundary conditions (fast algorithm, np >= 27) on every line of f viewed as
// [outer][np][inner]; displacement in PHYSICAL units (the alpha of interpolate_array_disp[_inplace]), delta = cell size.
cudaError_t launch_hermite(double *f, long long outer, int np, long long inner, const DispDesc &dd, double delta, int inplace,
                           int have_slopes, double sl, double sr, int staging, cudaStream_t st);
// K7: buf[o][j][in] = f[o][j0+j][in], j < hw
cudaError_t launch_halo_pack(const double *f, long long outer, int n, long long inner, int j0, int hw, double *buf,
                             cudaStream_t st, const LineBox *box = nullptr, int max_blocks = 0);

// K3: rho[x] = scale * sum_v f[x + nx*v]; partial = scratch of reduce_scratch_doubles(nx, nv)
size_t reduce_scratch_doubles(long long nx, long long nv);
cudaError_t launch_reduce_velocity(const double *f, long long nx, long long nv, double scale, double *rho,
                                   double *scratch, cudaStream_t st);

cudaError_t launch_reduce_velocity_partials(const double *f, long long nx, long long nv, double *scratch, int *nchunks_out,
                                            cudaStream_t st);

// K8: per velocity index v: s[v] = (sum_x f, sum_x |f|, sum_x f^2)   -> out[3*nv]
cudaError_t launch_row_sums(const double *f, long long nx, long long nv, double *out3, cudaStream_t st);

// K4: spectral multipliers around cuFFT
cudaError_t launch_poisson1d_mult(const cufftDoubleComplex *rho_hat, int nc, double L, cufftDoubleComplex *e_hat,
                                  cudaStream_t st);
cudaError_t launch_poisson2d_mult(const cufftDoubleComplex *rho_hat, int n1, int n2, double L1, double L2,
                                  cufftDoubleComplex *phi_hat, cufftDoubleComplex *e1_hat, cufftDoubleComplex *e2_hat,
                                  cudaStream_t st);
cudaError_t launch_poisson2d_par_mult(const cufftDoubleComplex *rho_hat, int n1, int n2, double L1, double L2,
                                      cufftDoubleComplex *phi_hat, cudaStream_t st);
cudaError_t launch_poisson3d_mult(const cufftDoubleComplex *rho_hat, int n1, int n2, int n3, double L1, double L2,
                                  double L3, cufftDoubleComplex *phi_hat, cufftDoubleComplex *e1_hat,
                                  cufftDoubleComplex *e2_hat, cufftDoubleComplex *e3_hat, cudaStream_t st);

// small helpers
cudaError_t launch_affine(double *out, long long n, double a0, double a1, cudaStream_t st); // out[i]=a0+a1*i
cudaError_t launch_sum_squares(const double *a, long long n, double *out1, cudaStream_t st);
cudaError_t launch_rho_1d1v(double *rho, long long n, double c0, double c1, cudaStream_t st); // rho = c0 - c1*rho
// K6 pack/unpack between a local box array and a contiguous buffer
struct Box4 { int lo[4]; int n[4]; };   // sub-box origin (local coords) and extents
cudaError_t launch_pack4d(const double *src, const int ext[4], Box4 box, double *buf, cudaStream_t st);
cudaError_t launch_unpack4d(double *dst, const int ext[4], Box4 box, const double *buf, cudaStream_t st);

extern int g_plane_ept;    // plane kernel: 0 auto, 16 or 32 points per thread
extern int g_plane_tmem;   // plane kernel, charge-density accumulators in tensor memory: -1 auto, 0, 1
extern int g_plane_const_dims; // plane kernel: instantiations with compile-time extents for 128 x 128 / 64 x 64 planes
extern int g_remap_rotation; // fused remap: rank-dependent start tile
extern int g_spline_split; // -1 auto, else lines are cut into this many chunks (1,2,4,8)
long long launch_count();
void count_launch();
void count_launches(long long n);
void launch_count_reset();

} // namespace sllb
