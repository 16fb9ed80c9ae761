/*
 * sll_b200_sim6d_compat.h -- the C interface SeLaLib's 6D simulation already exports, served by libsllb200.so.
 *
 * Same symbol names and by-reference calling convention as the bind(C) procedures of
 * simulations/parallel/bsl_vp_3d3v_cart_dd/sll_m_sim_bsl_vp_3d3v_cart_dd_slim_interface.F90:63-283
 * (C++ prototypes: test_cpp_interface.cpp:12-30).  `sim` is the address of an opaque handle that init fills.
 * Like the originals these return nothing; on error they print the message and stop the program, which is what
 * SLL_ERROR does (src/low_level_utilities/errors/sll_errors.h:4).
 *
 * init reads the reference's namelist file (groups sim_params, grid_dims, domain_dims, advect_params, output,
 * parallel_params, landau_params; sll_m_sim_bsl_vp_3d3v_cart_dd_slim.F90:322-358) and writes the t = 0 row of
 * <file_prefix>.dat; run does advect_v(dt/2) followed by n_iterations steps, one row per step, in the
 * reference's e20.12 format (sll_m_sim_6d_utilities.F90:632-633).
 * Supported: test_case "landau_prod", interpolator_type "fixed" (odd stencils 3..11), bc_type sll_p_periodic.
 */
#ifndef SLL_B200_SIM6D_COMPAT_H
#define SLL_B200_SIM6D_COMPAT_H

#include <stdint.h>

#include "sll_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

void sim_bsl_vp_3d3v_cart_dd_slim_init(void **sim, const char *filename);              /* interface.F90:63-84 */
void sim_bsl_vp_3d3v_cart_dd_slim_run(void **sim);                                     /* :95-103 */
void sim_bsl_vp_3d3v_cart_dd_slim_delete(void **sim);                                  /* :105-112 */
/* hands out a HOST mirror of the local block (column-major, extents from get_local_size); the caller may
 * write into it, the next compute call uploads it again (the reference hands out the live array, :114-122) */
void sim_bsl_vp_3d3v_cart_dd_slim_get_distribution(void **sim, double **f);
/* the Fortran dummy is type(c_ptr), VALUE (:124-138): the pointer itself, not its address */
void sim_bsl_vp_3d3v_cart_dd_slim_set_distribution(void **sim, double *f);
void sim_bsl_vp_3d3v_cart_dd_slim_get_local_size(void **sim, int32_t *n6);            /* :140-146 */
void sim_bsl_vp_3d3v_cart_dd_slim_advect_v(void **sim, double *delta_t);              /* :269-275 */
void sim_bsl_vp_3d3v_cart_dd_slim_advect_x(void **sim);                               /* :277-282 */
void sim_bsl_vp_3d3v_cart_dd_slim_print_etas(void **sim);                             /* :149-178 */
void sim_bsl_vp_3d3v_cart_dd_slim_write_diagnostics_init(void **sim);                 /* :181-226 */
void sim_bsl_vp_3d3v_cart_dd_slim_write_diagnostics(void **sim, int32_t *timeStepNumber); /* :228-267 */

/* MPI hand-over of the reference (src/parallelization/collective/sll_m_collective.F90:433-460): no-ops here.
 * A multi-GPU host passes its NCCL communicator with sllb_sim6d_compat_set_comm before init. */
void sll_s_allocate_collective(void);
void sll_s_set_communicator_collective(int *mpi_comm_f);
void sll_s_halt_collective(void);
int sllb_sim6d_compat_set_comm(sllb_comm_t comm);

/* sll_s_check_diagnostics (sll_m_sim_6d_utilities.F90:648-687): 3 x 14 numbers, max abs difference < 5e-7;
 * prints "PASSED." / "FAILED." like the reference and returns 0 when passed */
int sllb_sim6d_compat_check(const char *reffile, const char *simfile);

#ifdef __cplusplus
}
#endif
#endif
