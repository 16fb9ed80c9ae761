!> @brief iso_c_binding interfaces to libsllb200.so (include/sll_b200.h).
!> @details
!> SOURCE ONLY: there is no Fortran compiler in the build image, so this file has never
!> been compiled.  It is the binding a SeLaLib maintainer adds to the tree (see
!> INTEGRATION.md); every interface below mirrors one prototype of include/sll_b200.h.
!> Handles are opaque type(c_ptr); all entry points return an integer(c_int) status,
!> 0 = ok; sll_s_b200_check turns a non-zero status into SLL_ERROR like the rest of the
!> library reports failures (src/low_level_utilities/errors/sll_errors.h).
module sll_m_b200_bindings
#include "sll_working_precision.h"
#include "sll_errors.h"
   use, intrinsic :: iso_c_binding
   implicit none

   public

   integer(c_int), parameter :: sllb_adv_periodic_spline = 0
   integer(c_int), parameter :: sllb_adv_periodic_lagrange = 1
   integer(c_int), parameter :: sllb_interp_cubic_spline = 0
   integer(c_int), parameter :: sllb_interp_lagrange_centered = 1
   integer(c_int), parameter :: sllb_interp_lagrange_fixed = 2
   integer(c_int), parameter :: sllb_interp_periodic_spline = 3
   integer(c_int), parameter :: sllb_interp_periodic_lagrange = 4
   integer(c_int), parameter :: sllb_bc_periodic = 0
   integer(c_int), parameter :: sllb_bc_hermite = 1   ! sll_p_hermite, cubic-spline interpolator only
   integer(c_int), parameter :: sllb_method_spline = 0
   integer(c_int), parameter :: sllb_method_lagrange_fixed = 1
   integer(c_int), parameter :: sllb_method_lagrange_centered = 2

   interface
      function sllb_last_error() bind(C, name="sllb_last_error") result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function
      function sllb_init(device) bind(C, name="sllb_init") result(ierr)
         import :: c_int
         integer(c_int), value :: device
         integer(c_int) :: ierr
      end function

      ! ---- sll_c_advector_1d drop-in (line granular, host arrays) ----
      function sllb_adv1d_create(kind, num_cells, xmin, xmax, order, h) &
         bind(C, name="sllb_adv1d_create") result(ierr)
         import :: c_int, c_double, c_ptr
         integer(c_int), value :: kind, num_cells, order
         real(c_double), value :: xmin, xmax
         type(c_ptr), intent(out) :: h
         integer(c_int) :: ierr
      end function
      function sllb_adv1d_advect_constant(h, a, dt, input, output, n) &
         bind(C, name="sllb_adv1d_advect_constant") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: h
         real(c_double), value :: a, dt
         real(c_double), intent(in) :: input(*)
         real(c_double), intent(inout) :: output(*)
         integer(c_int), value :: n
         integer(c_int) :: ierr
      end function
      function sllb_adv1d_delete(h) bind(C, name="sllb_adv1d_delete") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         integer(c_int) :: ierr
      end function

      ! ---- sll_c_interpolator_1d drop-in ----
      function sllb_interp1d_create(kind, num_points, xmin, xmax, bc, d_or_order, periodic_last, &
                                    fast_algorithm, h) bind(C, name="sllb_interp1d_create") result(ierr)
         import :: c_int, c_double, c_ptr
         integer(c_int), value :: kind, num_points, bc, d_or_order, periodic_last, fast_algorithm
         real(c_double), value :: xmin, xmax
         type(c_ptr), intent(out) :: h
         integer(c_int) :: ierr
      end function
      function sllb_interp1d_array_disp(h, n, data, alpha, output) &
         bind(C, name="sllb_interp1d_array_disp") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: h
         integer(c_int), value :: n
         real(c_double), intent(in) :: data(*)
         real(c_double), value :: alpha
         real(c_double), intent(inout) :: output(*)
         integer(c_int) :: ierr
      end function
      function sllb_interp1d_array_disp_inplace(h, n, data, alpha) &
         bind(C, name="sllb_interp1d_array_disp_inplace") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: h
         integer(c_int), value :: n
         real(c_double), intent(inout) :: data(*)
         real(c_double), value :: alpha
         integer(c_int) :: ierr
      end function
      function sllb_interp1d_delete(h) bind(C, name="sllb_interp1d_delete") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         integer(c_int) :: ierr
      end function

      ! ---- device-resident field + batched passes (the performance surface) ----
      function sllb_field_create(ndim, extents, f) bind(C, name="sllb_field_create") result(ierr)
         import :: c_int, c_ptr
         integer(c_int), value :: ndim
         integer(c_int), intent(in) :: extents(*)
         type(c_ptr), intent(out) :: f
         integer(c_int) :: ierr
      end function
      function sllb_field_destroy(f) bind(C, name="sllb_field_destroy") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: f
         integer(c_int) :: ierr
      end function
      function sllb_field_upload(f, host, dup_last) bind(C, name="sllb_field_upload") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: f
         real(c_double), intent(in) :: host(*)
         integer(c_int), intent(in) :: dup_last(*)
         integer(c_int) :: ierr
      end function
      function sllb_field_download(f, host, dup_last) bind(C, name="sllb_field_download") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: f
         real(c_double), intent(inout) :: host(*)
         integer(c_int), intent(in) :: dup_last(*)
         integer(c_int) :: ierr
      end function
      function sllb_advect_axis_affine(f, axis, method, order, v_axis, vmin, dv, scale) &
         bind(C, name="sllb_advect_axis_affine") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: f
         integer(c_int), value :: axis, method, order, v_axis
         real(c_double), value :: vmin, dv, scale
         integer(c_int) :: ierr
      end function
      function sllb_advect_axis_field(f, axis, method, order, d_field, nfield_axes, scale) &
         bind(C, name="sllb_advect_axis_field") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: f, d_field
         integer(c_int), value :: axis, method, order, nfield_axes
         real(c_double), value :: scale
         integer(c_int) :: ierr
      end function
      function sllb_reduce_velocity_host(f, nx_axes, scale, rho) &
         bind(C, name="sllb_reduce_velocity_host") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: f
         integer(c_int), value :: nx_axes
         real(c_double), value :: scale
         real(c_double), intent(inout) :: rho(*)
         integer(c_int) :: ierr
      end function
      function sllb_interp1d_set_slopes(h, slope_left, slope_right) bind(C, name="sllb_interp1d_set_slopes") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: h
         real(c_double), value :: slope_left, slope_right
         integer(c_int) :: ierr
      end function
      !> Hermite-BC spline on every line of a device field along `axis` (disp: type(sllb_disp_t) by reference)
      function sllb_advect_axis_hermite(f, axis, xmin, xmax, disp, inplace_semantics) &
         bind(C, name="sllb_advect_axis_hermite") result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: f, disp
         integer(c_int), value :: axis, inplace_semantics
         real(c_double), value :: xmin, xmax
         integer(c_int) :: ierr
      end function
      ! ---- local cubic splines with halo cells (sll_t_advection_6d_spline_dd_slim) ----
      function sllb_spline_dd_blocks(n, disp, shift, alpha, nblocks) bind(C, name="sllb_spline_dd_blocks") result(ierr)
         import :: c_int, c_int32_t, c_double
         integer(c_int), value :: n
         real(c_double), intent(in) :: disp(*)
         integer(c_int32_t), intent(out) :: shift(*)
         real(c_double), intent(out) :: alpha(*)
         integer(c_int), intent(out) :: nblocks
         integer(c_int) :: ierr
      end function
      function sllb_dd6d_create(comm, global, procs, d) bind(C, name="sllb_dd6d_create") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: comm
         integer(c_int), intent(in) :: global(6), procs(6)
         type(c_ptr), intent(out) :: d
         integer(c_int) :: ierr
      end function
      function sllb_dd6d_field(d, f) bind(C, name="sllb_dd6d_field") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: d
         type(c_ptr), intent(out) :: f
         integer(c_int) :: ierr
      end function
      !> disp: type(sllb_disp_t) by reference; shift: the table of sllb_spline_dd_blocks or c_null_ptr
      function sllb_dd6d_advect_axis_spline(d, axis, disp, shift, hw_left, hw_right) &
         bind(C, name="sllb_dd6d_advect_axis_spline") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: d, disp, shift
         integer(c_int), value :: axis, hw_left, hw_right
         integer(c_int) :: ierr
      end function
      ! ---- splitting schedules (sll_f_new_time_splitting_coeff) ----
      function sllb_splitting_case_from_name(name, split_case) bind(C, name="sllb_splitting_case_from_name") result(ierr)
         import :: c_int, c_char
         character(kind=c_char), intent(in) :: name(*)
         integer(c_int), intent(out) :: split_case
         integer(c_int) :: ierr
      end function
      function sllb_splitting_coeff(split_case, dt, steps, nsteps, nb_split_step, split_begin_t, dim_split_v) &
         bind(C, name="sllb_splitting_coeff") result(ierr)
         import :: c_int, c_double
         integer(c_int), value :: split_case
         real(c_double), value :: dt
         real(c_double), intent(out) :: steps(32)
         integer(c_int), intent(out) :: nsteps, nb_split_step, split_begin_t, dim_split_v
         integer(c_int) :: ierr
      end function
   end interface

contains

   !> Non-zero status -> SLL_ERROR with the library's message (program stops, as in the reference).
   subroutine sll_s_b200_check(ierr, fun)
      integer(c_int), intent(in) :: ierr
      character(len=*), intent(in) :: fun
      character(kind=c_char), pointer :: cmsg(:)
      character(len=512) :: msg
      type(c_ptr) :: p
      integer :: i
      if (ierr == 0) return
      msg = ' '
      p = sllb_last_error()
      if (c_associated(p)) then
         call c_f_pointer(p, cmsg, [512])
         do i = 1, 512
            if (cmsg(i) == c_null_char) exit
            msg(i:i) = cmsg(i)
         end do
      end if
      SLL_ERROR(fun, trim(msg))
   end subroutine sll_s_b200_check

end module sll_m_b200_bindings
