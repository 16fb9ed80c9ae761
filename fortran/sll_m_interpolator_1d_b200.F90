!> @ingroup interpolators
!> @brief B200 drop-in for the constant-displacement procedures of sll_c_interpolator_1d.
!> @details
!> SOURCE ONLY (never compiled: no Fortran compiler in the build image).
!> Replaces sll_t_cubic_spline_interpolator_1d (sll_m_cubic_spline_interpolator_1d.F90:62-180),
!> sll_t_lagrange_interpolator_1d (sll_m_lagrange_interpolator_1d.F90:43-199) and
!> sll_t_periodic_interpolator_1d for the two procedures the split semi-Lagrangian time loops call:
!> interpolate_array_disp and interpolate_array_disp_inplace, output(i) = f(x_i + alpha).
!> The seven other deferred procedures are not part of the accelerated path; like the reference's
!> Lagrange interpolator (sll_m_lagrange_interpolator_1d.F90:270-299) they print and stop.
module sll_m_interpolator_1d_b200
#include "sll_working_precision.h"
#include "sll_errors.h"
   use, intrinsic :: iso_c_binding
   use sll_m_interpolators_1d_base, only: sll_c_interpolator_1d
   use sll_m_b200_bindings
   implicit none

   public :: sll_t_interpolator_1d_b200
   private

   type, extends(sll_c_interpolator_1d) :: sll_t_interpolator_1d_b200
      type(c_ptr) :: handle = c_null_ptr
      sll_int32   :: num_points
   contains
      !> kind = sllb_interp_* ; d_or_order = Lagrange d or periodic order; periodic_last as in
      !> sll_m_lagrange_interpolator_1d.F90:89-130
      procedure, pass(interpolator) :: init => b200_interp_init
      procedure :: compute_interpolants => b200_compute_interpolants
      procedure :: set_coefficients => b200_set_coefficients
      procedure :: get_coefficients => b200_get_coefficients
      procedure :: interpolate_from_interpolant_value => b200_value
      procedure :: interpolate_from_interpolant_derivative_eta1 => b200_deriv
      procedure :: interpolate_from_interpolant_array => b200_array_interpolant
      procedure :: interpolate_array => b200_array
      procedure :: interpolate_array_disp => b200_array_disp
      procedure :: interpolate_array_disp_inplace => b200_array_disp_inplace
      procedure, pass(interpolator) :: free => b200_interp_free
   end type sll_t_interpolator_1d_b200

contains

   !> bc_type = sllb_bc_hermite with optional slope_left / slope_right mirrors
   !> sll_t_cubic_spline_interpolator_1d%init(num_points, xmin, xmax, sll_p_hermite [, slope_left, slope_right])
   !> (sll_m_cubic_spline_interpolator_1d.F90:317-381)
   subroutine b200_interp_init(interpolator, kind, num_points, xmin, xmax, d_or_order, periodic_last, fast_algorithm, &
                               bc_type, slope_left, slope_right)
      class(sll_t_interpolator_1d_b200), intent(inout) :: interpolator
      sll_int32, intent(in) :: kind, num_points
      sll_real64, intent(in) :: xmin, xmax
      sll_int32, intent(in), optional :: d_or_order, periodic_last, bc_type
      logical, intent(in), optional :: fast_algorithm
      sll_real64, intent(in), optional :: slope_left, slope_right
      integer(c_int) :: d, pl, fa, bc
      d = 4; pl = 1; fa = 1; bc = sllb_bc_periodic
      if (present(bc_type)) bc = int(bc_type, c_int)
      if (present(d_or_order)) d = int(d_or_order, c_int)
      if (present(periodic_last)) pl = int(periodic_last, c_int)
      if (present(fast_algorithm)) fa = merge(1_c_int, 0_c_int, fast_algorithm)
      interpolator%num_points = num_points
      call sll_s_b200_check(sllb_interp1d_create(int(kind, c_int), int(num_points, c_int), real(xmin, c_double), &
                                                 real(xmax, c_double), bc, d, pl, fa, &
                                                 interpolator%handle), 'b200_interp_init')
      if (present(slope_left) .and. present(slope_right)) then
         call sll_s_b200_check(sllb_interp1d_set_slopes(interpolator%handle, real(slope_left, c_double), &
                                                        real(slope_right, c_double)), 'b200_interp_init')
      end if
   end subroutine b200_interp_init

   subroutine b200_array_disp(this, num_pts, data, alpha, output_array)
      class(sll_t_interpolator_1d_b200), intent(inout) :: this
      sll_int32, intent(in) :: num_pts
      sll_real64, intent(in) :: data(:)
      sll_real64, intent(in) :: alpha
      sll_real64, intent(out) :: output_array(num_pts)
      sll_real64, allocatable :: tmp(:)
      allocate (tmp(num_pts))
      tmp = data(1:num_pts)
      call sll_s_b200_check(sllb_interp1d_array_disp(this%handle, int(num_pts, c_int), tmp, real(alpha, c_double), &
                                                     output_array), 'b200_array_disp')
      deallocate (tmp)
   end subroutine b200_array_disp

   subroutine b200_array_disp_inplace(this, num_pts, data, alpha)
      class(sll_t_interpolator_1d_b200), intent(inout) :: this
      sll_int32, intent(in) :: num_pts
      sll_real64, intent(inout) :: data(num_pts)
      sll_real64, intent(in) :: alpha
      call sll_s_b200_check(sllb_interp1d_array_disp_inplace(this%handle, int(num_pts, c_int), data, &
                                                             real(alpha, c_double)), 'b200_array_disp_inplace')
   end subroutine b200_array_disp_inplace

   subroutine b200_interp_free(interpolator)
      class(sll_t_interpolator_1d_b200), intent(inout) :: interpolator
      if (c_associated(interpolator%handle)) then
         call sll_s_b200_check(sllb_interp1d_delete(interpolator%handle), 'b200_interp_free')
         interpolator%handle = c_null_ptr
      end if
   end subroutine b200_interp_free

   ! ---- procedures outside the accelerated path: print and stop, like the reference's Lagrange interpolator ----
   subroutine b200_compute_interpolants(interpolator, data_array, eta_coords, size_eta_coords)
      class(sll_t_interpolator_1d_b200), intent(inout) :: interpolator
      sll_real64, intent(in) :: data_array(:)
      sll_real64, intent(in), optional :: eta_coords(:)
      sll_int32, intent(in), optional :: size_eta_coords
      print *, 'b200 interpolator: compute_interpolants not implemented (constant displacement only)'
      print *, interpolator%num_points, maxval(data_array), present(eta_coords), present(size_eta_coords)
      stop
   end subroutine b200_compute_interpolants

   subroutine b200_set_coefficients(interpolator, coeffs)
      class(sll_t_interpolator_1d_b200), intent(inout) :: interpolator
      sll_real64, dimension(:), intent(in), optional :: coeffs
      print *, 'b200 interpolator: set_coefficients not implemented'
      print *, interpolator%num_points, present(coeffs)
      stop
   end subroutine b200_set_coefficients

   function b200_get_coefficients(interpolator)
      class(sll_t_interpolator_1d_b200), intent(in) :: interpolator
      sll_real64, dimension(:), pointer :: b200_get_coefficients
      print *, 'b200 interpolator: get_coefficients not implemented'
      print *, interpolator%num_points
      b200_get_coefficients => null()
      stop
   end function b200_get_coefficients

   function b200_value(interpolator, eta1) result(val)
      class(sll_t_interpolator_1d_b200), intent(in) :: interpolator
      sll_real64 :: val
      sll_real64, intent(in) :: eta1
      print *, 'b200 interpolator: interpolate_from_interpolant_value not implemented'
      print *, interpolator%num_points, eta1
      val = 0._f64
      stop
   end function b200_value

   function b200_deriv(interpolator, eta1) result(val)
      class(sll_t_interpolator_1d_b200), intent(in) :: interpolator
      sll_real64 :: val
      sll_real64, intent(in) :: eta1
      print *, 'b200 interpolator: interpolate_from_interpolant_derivative_eta1 not implemented'
      print *, interpolator%num_points, eta1
      val = 0._f64
      stop
   end function b200_deriv

   subroutine b200_array_interpolant(interpolator, num_pts, vals_to_interpolate, output_array)
      class(sll_t_interpolator_1d_b200), intent(inout) :: interpolator
      sll_int32, intent(in) :: num_pts
      sll_real64, intent(in) :: vals_to_interpolate(num_pts)
      sll_real64, intent(out) :: output_array(num_pts)
      print *, 'b200 interpolator: interpolate_from_interpolant_array not implemented'
      print *, interpolator%num_points, maxval(vals_to_interpolate)
      output_array = 0._f64
      stop
   end subroutine b200_array_interpolant

   subroutine b200_array(this, num_pts, data, coordinates, output_array)
      class(sll_t_interpolator_1d_b200), intent(inout) :: this
      sll_int32, intent(in) :: num_pts
      sll_real64, intent(in) :: data(:)
      sll_real64, intent(in) :: coordinates(num_pts)
      sll_real64, intent(out) :: output_array(num_pts)
      print *, 'b200 interpolator: interpolate_array (arbitrary feet) not implemented; use interpolate_array_disp'
      print *, this%num_points, maxval(data), maxval(coordinates)
      output_array = 0._f64
      stop
   end subroutine b200_array

end module sll_m_interpolator_1d_b200
