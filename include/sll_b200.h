/*
 * sll_b200.h -- C ABI of the B200-native split semi-Lagrangian advection path.
 *
 * Drop-in boundary for SeLaLib's sim_bsl_vp_* simulations.
 * Every entry point is extern "C", takes plain pointers and sizes, and returns an
 * int status (0 = ok, non-zero = error; text via sllb_last_error()).  A Fortran
 * caller binds these with iso_c_binding (see fortran/ and INTEGRATION.md) and turns a
 * non-zero status into SLL_ERROR, which is how the reference reports failures
 * (src/low_level_utilities/errors/sll_errors.h:4).
 *
 * Conventions
 *  - all arrays are column-major (Fortran order), fp64, indices int32 like sll_int32;
 *  - line-granular calls (sllb_adv1d_*, sllb_interp1d_*) take HOST pointers and mirror
 *    the reference objects call by call (parity / drop-in; each call copies the line to
 *    the GPU and back);
 *  - batched calls operate on a device-resident field (sllb_field_t) that stores the
 *    periodic cells only (N per axis).  Upload/download add or strip the duplicated
 *    periodic end point the 1D1V/2D2V simulations carry (N+1);
 *  - one handle per host thread (same rule as the reference: objects own scratch);
 *  - there is no CPU fallback: every compute entry point fails with SLLB_ERR_NO_DEVICE
 *    when no CUDA device is usable.
 * Paths cited below are relative to the reference tree.
 */
#ifndef SLL_B200_H
#define SLL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------- */
#define SLLB_OK 0
#define SLLB_ERR_INVALID 1     /* bad argument */
#define SLLB_ERR_UNSUPPORTED 2 /* method/stencil not implemented (reference: SLL_ERROR 'not implemented') */
#define SLLB_ERR_CUDA 3        /* CUDA / cuFFT / NCCL runtime failure */
#define SLLB_ERR_NO_DEVICE 4   /* no usable CUDA device */

const char *sllb_last_error(void);
int sllb_version(void);
/* Select the device this thread's handles live on (default 0).  */
int sllb_init(int device);
int sllb_device_count(int *count);
int sllb_synchronize(void);
/* number of kernels this library launched since the last reset (bench bookkeeping) */
int64_t sllb_launch_count(void);
void sllb_launch_count_reset(void);

/* ---- interpolation / advection methods ---------------------------------- */
/* advector kinds: sll_t_advector_1d_periodic with sll_p_spline / sll_p_lagrange
 * (src/semi_lagrangian/advection/sll_m_advection_1d_periodic.F90:41-54,
 *  src/interpolation/periodic_interpolation/sll_m_periodic_interp.F90:38-41) */
#define SLLB_ADV_PERIODIC_SPLINE 0
#define SLLB_ADV_PERIODIC_LAGRANGE 1
/* sll_t_advector_1d_bsl (src/semi_lagrangian/advection/sll_m_advection_1d_BSL.F90:40-54,152-164) wired as in the
 * reference's test_advection_1d_bsl.F90: explicit-Euler characteristics with sll_p_periodic
 * (sll_m_characteristics_1d_explicit_euler.F90:157-189) + sll_t_cubic_spline_interpolator_1d%interpolate_array.
 * For a constant advection field the feet are x_i - A*dt folded into the period, i.e. one shift per line:
 * the same fused kernel serves it.  `order` is ignored (cubic). */
#define SLLB_ADV_BSL 2

/* interpolator kinds (sll_c_interpolator_1d implementations) */
#define SLLB_INTERP_CUBIC_SPLINE 0      /* sll_t_cubic_spline_interpolator_1d */
#define SLLB_INTERP_LAGRANGE_CENTERED 1 /* sll_t_lagrange_interpolator_1d, sll_p_lagrange_centered */
#define SLLB_INTERP_LAGRANGE_FIXED 2    /* sll_t_lagrange_interpolator_1d, sll_p_lagrange_fixed */
#define SLLB_INTERP_PERIODIC_SPLINE 3   /* sll_t_periodic_interpolator_1d, sll_p_spline */
#define SLLB_INTERP_PERIODIC_LAGRANGE 4 /* sll_t_periodic_interpolator_1d, sll_p_lagrange */

#define SLLB_BC_PERIODIC 0 /* sll_p_periodic */
#define SLLB_BC_HERMITE 1  /* sll_p_hermite (cubic-spline interpolator only, fast algorithm, num_points >= 27);
                              other boundary types -> SLLB_ERR_UNSUPPORTED */

/* batched per-axis methods */
#define SLLB_METHOD_SPLINE 0            /* periodic spline: order 4 (cubic), 6 (quintic), 8 (septic) */
#define SLLB_METHOD_LAGRANGE_FIXED 1    /* odd stencil 3,5,7,9,11 centred on the grid point */
#define SLLB_METHOD_LAGRANGE_CENTERED 2 /* even stencil 4..18 centred on the foot cell (closed forms up to 8 as in the
                                           reference's fast module, product form beyond) */

/* ---- a1/a2: sll_c_advector_1d%advect_1d_constant -------------------------
 * replaces sll_f_new_periodic_1d_advector / periodic_advect_1d_constant
 * (sll_m_advection_1d_periodic.F90:57-130) and the abstract interface
 * (sll_m_advection_1d_base.F90:53-68).  out(x_i) = in(x_i - A*dt); `in` may alias
 * `out`; n = num_cells or num_cells+1 (the duplicate is filled when n > num_cells).
 * kind PERIODIC_SPLINE supports orders 4, 6 and 8; PERIODIC_LAGRANGE supports even orders 4..18 (the shipped two-stream
 * namelist of the 1D1V simulation uses 18); BSL: n = num_cells+1 points as the
 * reference object is built on npts grid points (n = num_cells also accepted). */
typedef struct sllb_adv1d *sllb_adv1d_t;
int sllb_adv1d_create(int kind, int num_cells, double xmin, double xmax, int order, sllb_adv1d_t *h);
int sllb_adv1d_advect_constant(sllb_adv1d_t h, double A, double dt, const double *in, double *out, int n);
int sllb_adv1d_delete(sllb_adv1d_t h);

/* ---- a5/a8: sll_c_interpolator_1d%interpolate_array_disp[_inplace] --------
 * replaces spline_interpolate1d_disp[_inplace]
 * (src/interpolation/interpolators/sll_m_cubic_spline_interpolator_1d.F90:112-180),
 * interpolate_array_disp_li1d (sll_m_lagrange_interpolator_1d.F90:169-199) and
 * per_interpolate1d_disp (sll_m_periodic_interpolator_1d.F90).
 * out(i) = f(x_i + alpha).  num_points counts the duplicated periodic point when
 * periodic_last = 1 (always for the spline interpolators).  d_or_order: Lagrange d
 * (stencil 2d centred / 2d+1 fixed) or the periodic interpolator's order.
 * fast_algorithm is accepted for API parity; both reference algorithms agree to 1e-15
 * and the device solve is exact to 3.6e-16 (27-term series as in the reference). */
typedef struct sllb_interp1d *sllb_interp1d_t;
int sllb_interp1d_create(int kind, int num_points, double xmin, double xmax, int bc, int d_or_order,
                         int periodic_last, int fast_algorithm, sllb_interp1d_t *h);
int sllb_interp1d_array_disp(sllb_interp1d_t h, int n, const double *data, double alpha, double *out);
int sllb_interp1d_array_disp_inplace(sllb_interp1d_t h, int n, double *data, double alpha);
int sllb_interp1d_delete(sllb_interp1d_t h);
/* (f)3: sll_p_hermite (src/splines/splines_basic/sll_m_cubic_splines.F90:325-369,583-652,692-748).  num_points grid
 * points on [xmin, xmax], no periodic duplicate; end slopes from 5-point one-sided differences of the data (:176-181)
 * unless set here (the optional slope_left / slope_right of init).  interpolate_array_disp follows
 * sll_s_cubic_spline_1d_eval_disp (:2616-2682), the in-place call clamps the feet to [xmin, xmax] and evaluates like
 * sll_s_cubic_spline_1d_eval_array (sll_m_cubic_spline_interpolator_1d.F90:165-178). */
int sllb_interp1d_set_slopes(sllb_interp1d_t h, double slope_left, double slope_right);

/* ---- device-resident distribution function ------------------------------ */
typedef struct sllb_field *sllb_field_t;
/* extents[d] = number of periodic cells along axis d (1 <= ndim <= 6) */
int sllb_field_create(int ndim, const int *extents, sllb_field_t *F);
int sllb_field_destroy(sllb_field_t F);
/* dup_last[d] = 1: the host array has extents[d]+1 points along d (last = first).
 * dup_last may be NULL (= all 0).  On download the duplicate is filled from index 1. */
int sllb_field_upload(sllb_field_t F, const double *host, const int *dup_last);
int sllb_field_download(sllb_field_t F, double *host, const int *dup_last);
int sllb_field_device_ptr(sllb_field_t F, double **dptr);
int sllb_field_extents(sllb_field_t F, int *ndim, int *extents);

/* Displacement of line (o, in) of an axis pass, in cells, out(i) = f(i + disp):
 *   disp = scale * values[ ((o / odiv) % omod) * ostr + ((in / idiv) % imod) * istr ]
 * where `in` is the flattened index of the axes faster than the advected one and `o` of
 * the slower ones.  Covers a16's formulas: alpha = v*step (one velocity index) and
 * alpha = E(i1,i2[,i3])*step (the field, fused into the kernel = K5).
 * values_on_device: 0 = host array of `nvalues` doubles (copied), 1 = device pointer. */
typedef struct {
    const double *values;
    int64_t nvalues;
    int values_on_device;
    double scale;
    int64_t odiv, omod, ostr, idiv, imod, istr;
} sllb_disp_t;

/* a2/a5/a10/a11 batched: one 1D advection applied to every line of F along `axis`,
 * in place.  Replaces the per-line loops of
 * sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:1037-1064,1143-1166,
 * sll_m_sim_bsl_vp_1d1v_cart.F90:1570-1585,1656-1686 and
 * sll_s_advection_6d_lagrange_dd_slim_advect_eta1..6
 * (sll_m_advection_6d_lagrange_dd_slim.F90:806-2001). order: 4 for SPLINE, stencil
 * width for Lagrange. */
int sllb_advect_axis(sllb_field_t F, int axis, int method, int order, const sllb_disp_t *disp);
/* convenience: disp = (vmin + i_vaxis*dv) * scale, the x-advection of a16 */
int sllb_advect_axis_affine(sllb_field_t F, int axis, int method, int order, int v_axis, double vmin,
                            double dv, double scale);
/* convenience: disp = field[x] * scale with field a DEVICE array over the first
 * `nfield_axes` axes of F (v-advection; K5) */
int sllb_advect_axis_field(sllb_field_t F, int axis, int method, int order, const double *d_field,
                           int nfield_axes, double scale);
/* K10, batched Hermite-BC spline: every line of F along `axis` (extents[axis] grid points on [xmin, xmax], NOT
 * periodic) is replaced by S(x_i + alpha), alpha = displacement of the line in PHYSICAL units, as the velocity
 * passes of simulations/parallel/bsl_vp_2d2v_cart/sll_m_sim_bsl_vp_2d2v_cart.F90:520-545 do line by line.
 * inplace_semantics: 1 = interpolate_array_disp_inplace (what that simulation calls), 0 = interpolate_array_disp. */
int sllb_advect_axis_hermite(sllb_field_t F, int axis, double xmin, double xmax, const sllb_disp_t *disp, int inplace_semantics);
/* K1c: the x1 and the x2 pass of a T stage in ONE sweep over f: every contiguous (axis 0, axis 1) plane is
 * staged in shared memory, both periodic cubic-spline advections are applied there and the plane is written
 * once (the two displacements must be constant over a plane, as alpha = v3*step, alpha = v4*step of
 * sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:1037-1064 are).  d_rho != NULL (DEVICE, extents[0]*extents[1]):
 * also returns rho = rho_scale * sum over the remaining axes of the result, i.e. the reduction that follows
 * the T stage (sll_m_reduction.F90:187-272), without another sweep.  Same values as two sllb_advect_axis
 * calls to rounding (restart series evaluated in Horner form).  SLLB_ERR_UNSUPPORTED when the plane does not
 * fit (extents multiples of 32, <= 16384 points) -- call sllb_advect_axis twice instead. */
int sllb_advect_plane(sllb_field_t F, int method, int order, const sllb_disp_t *disp0, const sllb_disp_t *disp1,
                      double rho_scale, double *d_rho);
/* tuning knob: on = 0 disables K1c inside the simulations; points_per_thread 0 (auto), 16 or 32 */
int sllb_set_plane_kernel(int on, int points_per_thread);
/* tuning knob of the register-resident plane kernel: tmem_accumulators = 1 / 0: the 32 charge-density accumulators of a
 * thread in tensor memory / in 12 registers + 20 shared-memory slots, -1 (default): tensor memory where the extents are
 * compile-time constants; const_extents = 1 (default): 128 x 128 and 64 x 64 planes run instantiations with compile-time
 * extents (no register spills, wraps as masks), 0: the run-time-extent kernel for every shape.  Same results either way. */
int sllb_set_plane_variant(int tmem_accumulators, int const_extents);
/* tuning knob: 0 = auto, 1 = TMA bulk staging, 2 = cp.async staging (strided kernels) */
int sllb_set_staging(int mode);
/* tuning knob: chunks per line of the strided spline kernel: -1 = auto, 1, 2, 4, 8 */
int sllb_set_spline_split(int chunks);

/* ---- (f)1: LOCAL cubic spline with halo cells ------------------------------
 * sll_t_advection_6d_spline_dd_slim (src/semi_lagrangian/advection/sll_m_advection_6d_spline_dd_slim.F90) on top of
 * sll_m_cubic_spline_halo_1d (src/interpolation/interpolators/sll_m_cubic_spline_halo_1d.F90:69-200): every rank
 * interpolates its piece of a line with a spline whose two start values are series truncated after NUM_TERMS = 15
 * terms (:11; truncation 2.6e-9 relative, the reference's own test tolerance is 4e-9), so results depend on the
 * decomposition exactly as in the reference.  out(i) = S(x_i + disp), disp = shift + alpha, alpha in [0,1).
 * The reference cuts the monotonic displacement array of eta1..3 into blocks of equal integer part
 * (make_blocks_spline :202-287); indices with disp == 0 are in no block and the line stays untouched. */
#define SLLB_SHIFT_SKIP INT32_MIN
/* host only: shift[j] = integer displacement of index j's block or SLLB_SHIFT_SKIP, alpha[j] = disp - floor(disp)
 * (alpha / nblocks may be NULL) */
int sllb_spline_dd_blocks(int n, const double *disp, int32_t *shift, double *alpha, int *nblocks);
/* one pass along `axis` of a field whose axis is whole on this rank (procs(axis) == 1: the halo is the periodic image).
 * shift: HOST table indexed like disp->values (from sllb_spline_dd_blocks), or NULL = floor(disp) per line
 * (the eta4..6 rule, :1098-1102).  SLLB_ERR_UNSUPPORTED when the axis has <= 15 points (:79). */
int sllb_advect_axis_spline_dd(sllb_field_t F, int axis, const sllb_disp_t *disp, const int32_t *shift);

/* ---- a14: velocity reduction -> charge density ----------------------------
 * rho[x] = scale * sum over the last (ndim - nx_axes) axes of F.  With periodic cells
 * only, the trapezoid rule over the duplicated end points
 * (src/parallelization/reduction/sll_m_reduction.F90:187-272) IS the plain sum;
 * 6D: scale = -dV_v (sll_m_sim_6d_utilities.F90:203-245). d_rho: DEVICE, prod(extents[0:nx_axes]). */
int sllb_reduce_velocity(sllb_field_t F, int nx_axes, double scale, double *d_rho);
/* same, result copied to a HOST array */
int sllb_reduce_velocity_host(sllb_field_t F, int nx_axes, double scale, double *h_rho);
/* K8: weighted moments over the whole field (diagnostics), HOST out[3 + 2*nv]:
 *   [sum f, sum |f|, sum f^2, sum f*w1_a[i_a] (a = 1..nv), sum f*w2_a[i_a] (a = 1..nv)]
 * for the trailing `nv` axes; w1 / w2 are HOST arrays holding the per-axis weights of those axes
 * back to back (first and second velocity moments: v and v^2, with whatever end-point rule the
 * caller's quadrature uses). */
int sllb_moments(sllb_field_t F, int nv, const double *w1, const double *w2, double *out);

/* ---- a15: periodic Poisson solvers (cuFFT) -------------------------------- */
typedef struct sllb_poisson *sllb_poisson_t;
/* sll_t_poisson_1d_periodic (sll_m_poisson_1d_periodic.F90:102-173) */
int sllb_poisson1d_create(int nc, double xmin, double xmax, sllb_poisson_t *P);
/* sll_t_poisson_2d_periodic_fft (sll_m_poisson_2d_periodic.F90:250-383) */
int sllb_poisson2d_create(int nc_x, int nc_y, double x_min, double x_max, double y_min, double y_max,
                          sllb_poisson_t *P);
/* sll_t_poisson_2d_periodic_par (sll_m_poisson_2d_periodic_par.F90:120-338, the solver of sim_bsl_vp_2d2v_cart): solves
 * Delta phi = rho on [0,Lx] x [0,Ly] (the caller gives the source its sign), zero-mean potential, potential only
 * (sllb_poisson_solve[_host] with e1 = e2 = NULL); the reference distributes the two 1D transforms over the ranks of a
 * 2D layout, here the nc_x x nc_y problem is solved on the device that holds rho */
int sllb_poisson2d_par_create(int ncx, int ncy, double Lx, double Ly, sllb_poisson_t *P);
/* sll_t_poisson_3d_periodic_par (sll_m_poisson_3d_periodic_par.F90:297-470,981-1158), replicated */
int sllb_poisson3d_create(int nx, int ny, int nz, double Lx, double Ly, double Lz, sllb_poisson_t *P);
int sllb_poisson_destroy(sllb_poisson_t P);
/* DEVICE arrays of the periodic cells only (nc_x*nc_y..., no duplicates).  Any of the
 * outputs may be NULL.  1D: e2,e3 ignored. 2D: e3 ignored.  E = -grad phi, -lap phi = rho. */
int sllb_poisson_solve(sllb_poisson_t P, const double *d_rho, double *d_phi, double *d_e1, double *d_e2,
                       double *d_e3);
/* 2D solves on grids up to 256 x 256 run as three dense-DFT kernels written for this path (sllb_poisson_direct.cu:
 * a 128 x 128 problem is bound by the ~10 launches of the library FFT route); 0 = always cuFFT.  Same values to ~1e-15. */
int sllb_set_poisson_direct(int on);
/* HOST arrays with leading dimensions ld (= nc or nc+1, duplicates filled like
 * sll_m_poisson_2d_periodic.F90:370-374 / sll_m_poisson_1d_periodic.F90:170) */
int sllb_poisson_solve_host(sllb_poisson_t P, const double *rho, const int *ld, double *phi, double *e1,
                            double *e2, double *e3);

/* ---- a13: layouts and remap (x <-> v transposes), multi-GPU ----------------
 * Layout = box per rank of a global 4D array (sll_t_layout_4d,
 * src/parallelization/remap/sll_m_remapper.F90:1345-1487); process meshes are powers of two
 * from sll_s_factorize_in_two_powers_of_two (:6452-6476); rank = i+P1*(j+P2*(k+P3*l)) (:1056-1066).
 * Host-only helpers (no device needed): */
int sllb_factorize_in_two_powers_of_two(int num_procs, int *f1, int *f2);
/* boxes[rank][axis][0/1] = inclusive 0-based min/max, sll split rule (:1860-1939) */
int sllb_layout4d_boxes(const int global[4], const int procs[4], int nranks, int *boxes /* nranks*8 */);
/* send/recv plan between two layouts for `rank`: for every peer the intersection box
 * (global 0-based, inclusive) or an empty box (min>max). out: nranks*8 ints each. */
int sllb_remap4d_plan(const int global[4], const int procs_from[4], const int procs_to[4], int nranks, int rank,
                      int *send_boxes, int *recv_boxes);

/* communicator: one process per GPU; id = 128-byte ncclUniqueId from rank 0 */
typedef struct sllb_comm *sllb_comm_t;
int sllb_comm_unique_id(void *id128);
int sllb_comm_create(const void *id128, int nranks, int rank, sllb_comm_t *c);
int sllb_comm_destroy(sllb_comm_t c);
int sllb_comm_allreduce_sum(sllb_comm_t c, double *d_buf, int64_t count);
int sllb_comm_allgather(sllb_comm_t c, const double *d_send, double *d_recv, int64_t count_per_rank);

/* distributed 4D field: two layouts (x-sequential: axes 0,1 whole, axes 2,3 split;
 * v-sequential: axes 2,3 whole, axes 0,1 split) and the remap between them
 * (apply_remap_4D_double, sll_m_remapper.F90:3308-3456) = pack kernel + NCCL
 * send/recv group (all-to-all) + unpack kernel. */
typedef struct sllb_dist4d *sllb_dist4d_t;
int sllb_dist4d_create(sllb_comm_t c, const int global[4], sllb_dist4d_t *D);
int sllb_dist4d_destroy(sllb_dist4d_t D);
/* local fields (owned by D); which = 0: x-sequential layout, 1: v-sequential layout */
int sllb_dist4d_field(sllb_dist4d_t D, int which, sllb_field_t *F);
int sllb_dist4d_box(sllb_dist4d_t D, int which, int box[8]);
/* direction 0: x-seq -> v-seq, 1: v-seq -> x-seq */
int sllb_dist4d_remap(sllb_dist4d_t D, int direction);
/* Fused pass: advect every line of layout `from` along `axis` (whole in that layout) and store the result
 * straight into the OTHER layout on the ranks that own it there (peer-mapped arrays over NVLink, CUDA IPC),
 * then a cross-rank barrier.  = advect_1d_constant on every line + apply_remap_4D_double in one kernel.
 * Needs uniform boxes and peer access (sllb_dist4d_p2p); disp values must be a device pointer. */
int sllb_dist4d_advect_remap(sllb_dist4d_t D, int from, int axis, int method, int order, const sllb_disp_t *disp);
int sllb_dist4d_p2p(sllb_dist4d_t D, int *enabled);
/* 1 (default): the simulations use the fused pass when available; 0: pack + NCCL send/recv + unpack */
int sllb_set_fused_remap(int on);
/* 1 (default): a fused pass starts its sweep at a rank-dependent tile so that at every moment the senders are
 * spread evenly over the receivers (no NVLink ingress hot spot); 0: every rank sweeps in the same order */
int sllb_set_remap_rotation(int on);

/* ---- a12: 6D slim domain decomposition + halo exchange ----------------------
 * sll_f_set_process_grid (src/parallelization/decomposition/sll_m_decomposition.F90:2473-2543) */
int sllb_set_process_grid(int nranks, int grid[6]);
/* sll_t_cartesian_topology_6d + sll_t_decomposition_slim_6d (:124-141,209-227,379-555,835-869): block
 * decomposition of a global 6D array, rank order of MPI_Cart_create (last axis fastest), periodic ring
 * neighbours per axis.  procs = NULL or all zeros: sll_f_set_process_grid.  n % procs must be 0 (:846). */
typedef struct sllb_dd6d *sllb_dd6d_t;
/* host-only (no device): the block, coordinates and neighbours of `rank`; outputs may be NULL */
int sllb_dd6d_plan(int nranks, int rank, const int global[6], const int procs_in[6], int procs[6], int coords[6],
                   int mn[6], int nw[6], int left[6], int right[6]);
int sllb_dd6d_create(sllb_comm_t c /* NULL = one rank */, const int global[6], const int procs[6], sllb_dd6d_t *D);
int sllb_dd6d_destroy(sllb_dd6d_t D);
int sllb_dd6d_field(sllb_dd6d_t D, sllb_field_t *F); /* local block, owned by D */
/* any output may be NULL: process grid, my coordinates, global offset (0-based) and width of my block,
 * left / right neighbour ranks per axis */
int sllb_dd6d_layout(sllb_dd6d_t D, int procs[6], int coords[6], int mn[6], int nw[6], int left[6], int right[6]);
/* sll_s_apply_halo_exchange_slim_6d_real64 (:1715-2030): fills the left halo (hw_left planes, the last
 * planes of the left neighbour) and the right halo (hw_right planes, the first planes of the right
 * neighbour) along `axis`; pack kernel + ncclSend/ncclRecv pair per side, or a local periodic copy when
 * procs(axis) == 1. Halo buffers are [outer][hw][inner] like the reference's 6D halo arrays. */
int sllb_dd6d_halo_exchange(sllb_dd6d_t D, int axis, int hw_left, int hw_right);
int sllb_dd6d_halo_download(sllb_dd6d_t D, int side /* 0 left, 1 right */, double *host);
int sllb_dd6d_exchange_ms(sllb_dd6d_t D, double *ms); /* device time of the last exchange */
/* halo exchange + sll_s_advection_6d_lagrange_dd_slim_advect_eta{axis+1}
 * (src/semi_lagrangian/advection/sll_m_advection_6d_lagrange_dd_slim.F90:806-2001), fixed odd stencil,
 * in place on the local block; `disp` indexes the LOCAL block like sllb_advect_axis. */
int sllb_dd6d_advect_axis(sllb_dd6d_t D, int axis, int stencil, const sllb_disp_t *disp);
/* (f)1 on a decomposed block: sll_s_advection_6d_spline_dd_slim_[f]advect_eta{axis+1}.  Split axis: K9p computes the
 * neighbours' parts of the boundary sums (prepare_exchange), they travel with hw_left / hw_right halo planes
 * (sll_s_apply_bc_exchange_slim_6d_real64 + halo exchange, peer stores over NVLink or ncclSend/ncclRecv), then the
 * local spline runs on halo | block | halo.  Every line's shift must lie in [-hw_left, hw_right-1] (the reference's
 * eta4..6 use 1, 1: shifts 0 and -1, :1082-1087).  Unsplit axis: same as sllb_advect_axis_spline_dd. */
int sllb_dd6d_advect_axis_spline(sllb_dd6d_t D, int axis, const sllb_disp_t *disp, const int32_t *shift, int hw_left,
                                 int hw_right);
/* make_blocks_lagrange (sll_m_advection_6d_lagrange_dd_slim.F90:202-286), host only: the blocks of equal integer
 * displacement of the centred Lagrange x-advection and the halo widths each block exchanges,
 * halo_width[2*b] = stencil/2 - box - 1 (left), halo_width[2*b+1] = stencil/2 + box (right).  box[j] = block's integer
 * displacement or SLLB_SHIFT_SKIP (disp == 0: untouched line).  SLLB_ERR_INVALID when a displacement leaves
 * [-stencil/2, stencil/2) (the reference's SLL_ASSERT :232-235).  halo_width holds up to `stencil` blocks. */
int sllb_lagrange_dd_blocks(int n, int stencil, const double *disp, int32_t *box, int *nblocks, int *halo_width);
/* tuning / test knob: 1 = also take the halo path (local periodic halo copy + halo-cells kernel, exactly
 * the reference's sequence) when procs(axis) == 1; default 0 = periodic kernel, same arithmetic */
int sllb_dd6d_set_force_halo(int on);
/* 1 (default): when peer mapping is available (CUDA IPC, <= 8 ranks, halo <= 5 planes) the pack kernel stores the
 * edge planes straight into the neighbour's halo buffer over NVLink; 0: pack + ncclSend/ncclRecv */
int sllb_dd6d_set_halo_p2p(int on);
int sllb_dd6d_p2p(sllb_dd6d_t D, int *enabled);
/* a split-axis pass over peer memory is pipelined: the lines are cut into `chunks` pieces, the exchange of piece c+1
 * (edge planes stored into the neighbours' halo buffers + barrier) overlaps the stencil kernel of piece c on a second
 * stream.  Default 4 (or SLLB_HALO_CHUNKS); 1 = exchange everything, then advect.  Same values either way. */
int sllb_dd6d_set_halo_chunks(int chunks);
/* host only: the pieces (o0, ocount, i0, icount) x *nboxes the lines [outer][inner] of a pass are cut into; boxes4 holds
 * up to 4 * 16 values */
int sllb_dd6d_chunk_boxes(long long outer, long long inner, int nchunks, long long *boxes4, int *nboxes);

/* ---- a17 / (f)3: operator-splitting schedules ------------------------------
 * sll_f_new_time_splitting_coeff (src/time_integration/splitting_methods/sll_m_time_splitting_coeff.F90:86-594):
 * every split_case the 2D2V simulation's namelist accepts (sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:514-554). */
#define SLLB_SPLIT_STRANG_VTV 0
#define SLLB_SPLIT_STRANG_TVT 1
#define SLLB_SPLIT_LIE_TV 2
#define SLLB_SPLIT_LIE_VT 3
#define SLLB_SPLIT_TRIPLE_JUMP_TVT 4
#define SLLB_SPLIT_TRIPLE_JUMP_VTV 5
#define SLLB_SPLIT_ORDER6_VTV 6
#define SLLB_SPLIT_ORDER6_TVT 7
#define SLLB_SPLIT_ORDER6VP_TVT 8
#define SLLB_SPLIT_ORDER6VP_VTV 9
#define SLLB_SPLIT_ORDER6VPNEW_TVT 10
#define SLLB_SPLIT_ORDER6VPNEW1_VTV 11 /* the simulation's default (:342) */
#define SLLB_SPLIT_ORDER6VPNEW2_VTV 12
#define SLLB_SPLIT_ORDER6VP2D_VTV 13
#define SLLB_SPLIT_ORDER6VPOT_VTV 14     /* 14..17: dim_split_V = 2, V stages also move along the modified potential */
#define SLLB_SPLIT_ORDER6VPOTNEW1_VTV 15
#define SLLB_SPLIT_ORDER6VPOTNEW2_VTV 16
#define SLLB_SPLIT_ORDER6VPOTNEW3_VTV 17
#define SLLB_SPLIT_MAX_STEPS 32
/* host only.  name = the namelist string, e.g. "SLL_ORDER6VPnew1_VTV" */
int sllb_splitting_case_from_name(const char *name, int *split_case);
const char *sllb_splitting_case_name(int split_case);
/* steps[SLLB_SPLIT_MAX_STEPS] = split_step(:) as sll_t_splitting_coeff holds it (one entry per T stage, dim_split_V
 * entries per V stage; *nsteps of them); dt enters the Vlasov-Poisson order-6 weights.  Outputs may be NULL. */
int sllb_splitting_coeff(int split_case, double dt, double *steps, int *nsteps, int *nb_split_step, int *split_begin_T,
                         int *dim_split_V);
/* sll_s_compute_w_hermite (src/semi_lagrangian/fcisl/sll_m_fcisl.F90:413-486): first-derivative weights on the
 * stencil r..s (r < 0 < s), w[k - r]; used by compute_jacobian (...poisson_serial.F90:1403-1436) */
int sllb_compute_w_hermite(int r, int s, double *w);

/* ---- simulations (time loops of SURVEY.md section 3) --------------------- */
/* 2D2V sim_bsl_vp_2d2v_cart_poisson_serial on 1..P GPUs.
 * split: SLLB_SPLIT_* (0 Strang VTV, 1 Strang TVT, 2 Lie TV, ...). method/order as SLLB_METHOD_*. */
typedef struct sllb_sim4d *sllb_sim4d_t;
typedef struct {
    int nc[4];
    double xmin[4], xmax[4];
    double kx1, kx2, eps; /* sll_f_landau_mode_initializer_4d */
    double dt;
    int split;
    int method, order;
    int stencil_r, stencil_s; /* finite-difference stencil of compute_jacobian; 0, 0 = the namelist default -2, 2 (:366-367) */
    /* advector_x1..x4 / order_x1..x4 of the namelist (:556-624): per-axis method and order; order_axis[d] = 0 means
     * "use method / order above" */
    int method_axis[4], order_axis[4];
    /* 0 (default): f lives on the periodic cells (N points per direction).  1: additionally carry the duplicated v_max
     * planes the reference's (N+1)-point arrays hold: a T stage moves them with +v_max while the v_min planes (the same
     * periodic cell) move with -v_max (:1037-1064), both enter the trapezoid rho with weight 1/2
     * (sll_m_reduction.F90:229-272), the next V stage overwrites them with the v_min planes again.  That end-plane term is
     * the whole difference between the two modes (~1e-9 .. 1e-6 of the field energy); with 1 the traces follow the
     * reference to rounding.  One GPU, splitting schemes whose step ends with a V stage. */
    int dup_velocity_planes;
} sllb_sim4d_params_t;
int sllb_sim4d_create(const sllb_sim4d_params_t *p, sllb_comm_t comm /* NULL = single GPU */, sllb_sim4d_t *S);
int sllb_sim4d_destroy(sllb_sim4d_t S);
/* advance nsteps; rows (HOST, may be NULL): nsteps x 6 = time, nrj, ekin, int f, int|f|, int f^2
 * (thdiag columns 1,2,3,8,9,10 of sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:1262-1275) */
int sllb_sim4d_run(sllb_sim4d_t S, int nsteps, int with_diagnostics, double *rows);
/* row for the current state (time 0 row before any step) */
int sllb_sim4d_diagnostics(sllb_sim4d_t S, double *row6);
/* Ensemble streaming on ONE GPU (parameter scans: many independent states through the same time step).  Every call
 * (1) starts the download of the state stepped by the PREVIOUS call into host_prev_out, (2) starts the upload of
 * host_next_in, (3) advances the state uploaded by the previous call by one time step while both copies run (PCIe
 * is full duplex; three device copies of f rotate), (4) waits for all three.  Both host arrays hold the periodic
 * cells (column-major), should be pinned, and may be NULL: N states take N + 2 calls, the first only uploads, the
 * last only downloads.  The states are independent: each one's fields are recomputed from its own f. */
int sllb_sim4d_stream_step(sllb_sim4d_t S, const double *host_next_in, double *host_prev_out);
/* Namelist front-end: reads the file sim_bsl_vp_2d2v_cart_poisson_serial takes (&geometry, &initial_function,
 * &time_iterations, &advector, &poisson; defaults and mesh cases as sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:300-440;
 * `filename` with or without the ".nml" the reference appends, :375), builds the simulation and returns
 * number_iterations / freq_diag_time.  SLL_SPLINES -> periodic splines of order 4, 6 or 8, SLL_LAGRANGE -> centred Lagrange of the
 * given order; SLL_LANDAU initial function.  Anything else: SLLB_ERR_UNSUPPORTED with the reference's message. */
int sllb_sim4d_create_from_namelist(const char *filename, sllb_comm_t comm, sllb_sim4d_t *S, int *number_iterations,
                                    int *freq_diag_time);
/* the whole program: create from the namelist, run number_iterations steps and write `thdiag_path` (rank 0) in the
 * reference's format, one '(13g20.12)' row at t = 0 and after every freq_diag_time steps (:998-1010,1262-1275) */
int sllb_sim4d_run_namelist(const char *filename, sllb_comm_t comm, const char *thdiag_path);
/* Fortran G20.12 edit descriptor (host helper of the writer): buf receives exactly 20 characters + NUL */
int sllb_format_g20_12(double x, char *buf21);
int sllb_format_g(double x, int w, int d, char *buf /* w + 1 chars */); /* Fortran Gw.d */
/* the 13 columns of the reference's thdiag file for the current state (:998-1010 at t = 0, :1262-1275 later):
 * time, nrj, ekin, nrj0, ekin0, max|jacobian_E|, nrj_jac, int f, int |f|, int f^2, mass0, mass0, l20 */
int sllb_sim4d_thdiag(sllb_sim4d_t S, double *row13);
int sllb_sim4d_field(sllb_sim4d_t S, sllb_field_t *F); /* local x-sequential field (remaps into it if needed) */
int sllb_sim4d_box(sllb_sim4d_t S, int which, int box[8]); /* my box: which = 0 x-sequential, 1 v-sequential */
/* per-phase device time of the last run() in ms: [advect, reduce+poisson, remap, diag] */
int sllb_sim4d_phase_ms(sllb_sim4d_t S, double out[4]);
/* finer: [local passes, reduce+poisson, NCCL remap, diag, fused advect+remap kernels, barriers after them] */
int sllb_sim4d_phase_ms6(sllb_sim4d_t S, double out[6]);
/* finest: [local passes, reduce+poisson, NCCL remap, diag, fused V-stage pass (x4 + remap), barrier after it,
 *          fused T-stage plane kernel (x1 + x2 + rho + remap), all-reduce (rho + barrier) after it] */
int sllb_sim4d_phase_ms8(sllb_sim4d_t S, double out[8]);
/* rho, E1, E2 of the last field solve of the time loop (N1 x N2 periodic cells, column-major); NULL = skip */
int sllb_sim4d_fields_host(sllb_sim4d_t S, double *rho, double *e1, double *e2);
/* (sum w f, sum w f^2) over the GLOBAL field with a weight that depends on the global index of every point: equal (to
 * rounding) on any number of ranks, sensitive to misplaced elements (cross-rank exactness signal of bench.py) */
int sllb_sim4d_checksum(sllb_sim4d_t S, double out[2]);
/* several GPUs: 1 = the V stage that ends with a remap is cut into chunks of the local (x1,x2) tile, the HBM-bound x3 pass
 * of chunk c+1 running under the NVLink-bound x4 + remap pass of chunk c (two streams); 0 (default) = two whole passes:
 * measured slower on 2 GPUs (2.80-2.89 vs 2.76 ms per 128^4 step), kept as an opt-in.  Bit-identical values.  on = 2..8 also sets the number of chunks (default 4). */
int sllb_set_v_overlap(int on);
/* phase timing is opt-in: 1 = sllb_sim4d_run records one CUDA event per phase (pooled, reused), 0 (default) = none */
int sllb_set_phase_timers(int on);

/* 1D1V sim_bsl_vp_1d1v_cart (single GPU). init 0 Landau, 1 two-stream. rows: nsteps x 8
 * (time, mass, l1, momentum, l2, ekin, epot, etot; sll_m_sim_bsl_vp_1d1v_cart.F90:1783) */
typedef struct sllb_sim2d *sllb_sim2d_t;
int sllb_sim2d_create(int nc_x1, int nc_x2, double x1_min, double x1_max, double x2_min, double x2_max,
                      int init, double kmode, double eps, double dt, int method, int order, sllb_sim2d_t *S);
int sllb_sim2d_run(sllb_sim2d_t S, int nsteps, double *rows);
/* the 1D1V step is a dozen kernels on a few MB (launch-bound): by default sllb_sim2d_run records one time step as a CUDA
 * graph after the first step and replays it; 0 = one launch per kernel (same values) */
int sllb_set_cuda_graphs(int on);
int sllb_sim2d_field(sllb_sim2d_t S, sllb_field_t *F);
int sllb_sim2d_destroy(sllb_sim2d_t S);
/* init also takes 2 = SLL_BUMP_ON_TAIL.  The knobs the namelist of sim_bsl_vp_1d1v_cart sets
 * (sll_m_sim_bsl_vp_1d1v_cart.F90:431-494): split_case (numbering of sllb_splitting_case_from_name; the 1D1V loop walks
 * split_step(1..nb_split_step) alternating T and V, :1462-1696, so only schemes with dim_split_V = 1 apply),
 * advector_x1 / advector_x2 with their orders, time_init. */
int sllb_sim2d_set_splitting(sllb_sim2d_t S, int split_case);
int sllb_sim2d_set_advectors(sllb_sim2d_t S, int method_x1, int order_x1, int method_x2, int order_x2);
int sllb_sim2d_set_time(sllb_sim2d_t S, double time_init);
int sllb_sim2d_geometry(sllb_sim2d_t S, double lim[4]); /* x1_min, x1_max, x2_min, x2_max */
int sllb_sim2d_fields_host(sllb_sim2d_t S, double *rho, double *efield); /* N1 cells each, NULL = skip */
/* one row of the reference's thdiag.dat (:1703-1801): time, mass, l1norm, momentum, l2norm, kinetic_energy,
 * potential_energy, total, Re/Im rho^_k (k = 0..nb_mode), f_hat_x2(k) = sum_v w_v |f^_k(v)|^2 (k = 0..nb_mode):
 * 8 + 3 (nb_mode + 1) doubles; transforms normalised by 1/N like the reference's r2r plan (:1195) */
int sllb_sim2d_thdiag(sllb_sim2d_t S, int nb_mode, double *row);
/* the reference's restart stream (:1281-1307 read, :1762-1770 write): time, then f with the duplicated end points,
 * (N1+1) x (N2+1) doubles column-major; a file written by either code is read by the other */
int sllb_sim2d_write_restart(sllb_sim2d_t S, const char *path);
int sllb_sim2d_read_restart(sllb_sim2d_t S, const char *path, double *time);
/* namelist front-end (:431-565 with the reference's defaults and error messages) and the whole program: thdiag.dat in
 * '(8g25.15)' + modes, x/v/f0/deltaf/rhotot/efield/t .bdat, f_plot_<iplot>_proc_0000.rst every freq_diag_restart steps,
 * restart_file / time_init_from_restart_file on the way in; files go to `outdir` (NULL = current directory) */
int sllb_sim2d_create_from_namelist(const char *filename, sllb_sim2d_t *S, int *number_iterations, int *freq_diag_time,
                                    int *nb_mode);
int sllb_sim2d_run_namelist(const char *filename, const char *outdir);

/* 3D3V sim_bsl_vp_3d3v_cart_dd_slim (fixed / centred Lagrange or local splines) on 1..P GPUs (velocity axes split,
 * halo exchange per v-advection, rho all-reduced). rows: R x 14 as the reference's <prefix>.dat
 * (sll_m_sim_6d_utilities.F90:357-364,632-633), R = nsteps + 1 for the FIRST call on a handle (it also writes the
 * t = 0 row) and nsteps for every later call (sllb_sim6d_run_rows tells which); every rank gets the global row.
 * time_in_phase: the reference ends its single loop with a half V step (:735-741).  A handle may be run in several
 * calls: a call that follows such an ending first applies the other half of that V step, so that run(a) followed by
 * run(b) advances f exactly as far as one run(a+b).  The two differ only by the interpolation error of applying that V
 * step as two halves (not by a lost half step); with time_in_phase = 0 they are bit-identical. */
typedef struct sllb_sim6d *sllb_sim6d_t;
typedef struct {
    int n[6];
    double v_max, x_max[3];
    int stencil_x, stencil_v;
    double delta_t;
    double alpha, kx[3], v_thermal[3]; /* landau_prod */
    int time_in_phase;
    /* interpolator_type of the namelist (sll_m_sim_bsl_vp_3d3v_cart_dd_slim.F90:363-372):
     * SLLB_ADVECTOR_FIXED    "fixed":    Lagrange, odd stencils stencil_x / stencil_v in all six directions;
     * SLLB_ADVECTOR_CENTERED "centered": eta1..3 by the variable-block centred Lagrange advector (even stencil_x,
     *                        fadvect_eta1..3, sll_m_advection_6d_lagrange_dd_slim.F90:173-322), eta4..6 fixed stencil_v;
     * SLLB_ADVECTOR_SPLINE   "spline":   local cubic splines (sll_t_advection_6d_spline_dd_slim), stencils ignored */
    int advector;
} sllb_sim6d_params_t;
#define SLLB_ADVECTOR_FIXED 0
#define SLLB_ADVECTOR_CENTERED 1
#define SLLB_ADVECTOR_SPLINE 2
int sllb_sim6d_create(const sllb_sim6d_params_t *p, sllb_sim6d_t *S);
/* process_grid: NULL / zeros = sll_f_set_process_grid(nranks) */
int sllb_sim6d_create_dist(const sllb_sim6d_params_t *p, sllb_comm_t comm, const int process_grid[6], sllb_sim6d_t *S);
int sllb_sim6d_run(sllb_sim6d_t S, int nsteps, double *rows);
int sllb_sim6d_run_rows(sllb_sim6d_t S, int nsteps, int *nrows); /* rows the next sllb_sim6d_run(S, nsteps, rows) writes */
int sllb_sim6d_field(sllb_sim6d_t S, sllb_field_t *F); /* local block */
int sllb_sim6d_decomposition(sllb_sim6d_t S, sllb_dd6d_t *D);
int sllb_sim6d_advect_x(sllb_sim6d_t S);
int sllb_sim6d_advect_v(sllb_sim6d_t S, double dt);
int sllb_sim6d_fields(sllb_sim6d_t S);                                   /* rho, Poisson, E */
int sllb_sim6d_diagnostics(sllb_sim6d_t S, double time, double *row14);  /* one row of <prefix>.dat */
int sllb_sim6d_halo_ms(sllb_sim6d_t S, double *ms, int reset);           /* accumulated halo-exchange device time */
/* the accumulation above costs one host synchronisation per split pass (and keeps the exchange of the next pass from being
 * issued early), so it is opt-in: 1 = time every exchange, 0 (default) = none */
int sllb_dd6d_set_exchange_timing(int on);
/* sll_t_clocks of the 6D simulation (sll_m_sim_6d_utilities.F90:132-140,765-826; labels of
 * sll_m_sim_bsl_vp_3d3v_cart_dd_slim.F90:684-957): wall-clock seconds under P, PC, PF, D, X, X1..X3, V, X4..X6, H4..H6;
 * sllb_sim6d_write_clocks writes them like sll_s_finalize_clocks does ("sll_clocks.txt" when path is NULL).  Opt-in:
 * every phase boundary waits for the default stream. */
int sllb_sim6d_set_clocks(sllb_sim6d_t S, int on);
int sllb_sim6d_write_clocks(sllb_sim6d_t S, const char *path);
/* hosts with their own time loop: between sllb_sim6d_advect_x and sllb_sim6d_fields -- the halo of the first split
 * velocity axis leaves now and travels under the field solve (sllb_sim6d_run does this itself) */
int sllb_sim6d_prefetch_v_halo(sllb_sim6d_t S);
int sllb_sim6d_destroy(sllb_sim6d_t S);

#ifdef __cplusplus
}
#endif
#endif /* SLL_B200_H */
