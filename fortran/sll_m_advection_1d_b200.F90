!> @ingroup advection
!> @brief B200 drop-in for sll_t_advector_1d_periodic (sll_m_advection_1d_periodic.F90:41-130).
!> @details
!> SOURCE ONLY (never compiled: no Fortran compiler in the build image).
!> Same constructor arguments and the same deferred procedures as the reference type, so a
!> simulation switches by replacing
!>     sll_f_new_periodic_1d_advector(...)   ->   sll_f_new_b200_1d_advector(...)
!> advect_1d_constant computes output(x_i) = input(x_i - a*dt) on the periodic line; input and
!> output may alias (the 2D2V simulation calls it with the same array twice,
!> sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:1045-1049).  advect_1d (non-constant A) is not
!> part of the accelerated path and stops like the reference (sll_m_advection_1d_periodic.F90:132-147).
module sll_m_advection_1d_b200
#include "sll_working_precision.h"
#include "sll_errors.h"
   use, intrinsic :: iso_c_binding
   use sll_m_advection_1d_base, only: sll_c_advector_1d
   use sll_m_periodic_interp, only: sll_p_spline, sll_p_lagrange
   use sll_m_b200_bindings
   implicit none

   public :: sll_t_advector_1d_b200, sll_f_new_b200_1d_advector
   private

   type, extends(sll_c_advector_1d) :: sll_t_advector_1d_b200
      type(c_ptr) :: handle = c_null_ptr
      sll_int32   :: num_cells
   contains
      procedure, pass(adv) :: init => b200_advect_1d_init
      procedure, pass(adv) :: advect_1d_constant => b200_advect_1d_constant
      procedure, pass(adv) :: advect_1d => b200_advect_1d
      procedure, pass(adv) :: delete => b200_advect_1d_delete
   end type sll_t_advector_1d_b200

contains

   function sll_f_new_b200_1d_advector(num_cells, xmin, xmax, type, order) result(adv)
      type(sll_t_advector_1d_b200), pointer :: adv
      sll_int32, intent(in) :: num_cells
      sll_real64, intent(in) :: xmin, xmax
      sll_int32, intent(in) :: type   !< sll_p_spline or sll_p_lagrange
      sll_int32, intent(in) :: order
      sll_int32 :: ierr
      allocate (adv, stat=ierr)
      if (ierr /= 0) then
         SLL_ERROR('sll_f_new_b200_1d_advector', 'allocation failed')
      end if
      call adv%init(num_cells, xmin, xmax, type, order)
   end function sll_f_new_b200_1d_advector

   subroutine b200_advect_1d_init(adv, num_cells, xmin, xmax, type, order)
      class(sll_t_advector_1d_b200), intent(inout) :: adv
      sll_int32, intent(in) :: num_cells
      sll_real64, intent(in) :: xmin, xmax
      sll_int32, intent(in) :: type, order
      integer(c_int) :: kind
      if (type == sll_p_spline) then
         kind = sllb_adv_periodic_spline
      else if (type == sll_p_lagrange) then
         kind = sllb_adv_periodic_lagrange
      else
         SLL_ERROR('b200_advect_1d_init', 'interpolation type not available on the B200 path')
      end if
      adv%num_cells = num_cells
      call sll_s_b200_check(sllb_adv1d_create(kind, int(num_cells, c_int), real(xmin, c_double), &
                                              real(xmax, c_double), int(order, c_int), adv%handle), &
                            'b200_advect_1d_init')
   end subroutine b200_advect_1d_init

   subroutine b200_advect_1d_constant(adv, a, dt, input, output)
      class(sll_t_advector_1d_b200) :: adv
      sll_real64, intent(in) :: a
      sll_real64, intent(in) :: dt
      sll_real64, dimension(:), intent(in) :: input
      sll_real64, dimension(:), intent(out) :: output
      sll_real64, allocatable :: tmp(:)
      integer(c_int) :: n
      n = int(size(input), c_int)
      ! assumed-shape dummies may be strided sections: hand contiguous storage to C
      allocate (tmp(n))
      tmp = input
      call sll_s_b200_check(sllb_adv1d_advect_constant(adv%handle, real(a, c_double), real(dt, c_double), &
                                                       tmp, tmp, n), 'b200_advect_1d_constant')
      output(1:n) = tmp
      deallocate (tmp)
   end subroutine b200_advect_1d_constant

   subroutine b200_advect_1d(adv, a, dt, input, output)
      class(sll_t_advector_1d_b200) :: adv
      sll_real64, dimension(:), intent(in) :: a
      sll_real64, intent(in) :: dt
      sll_real64, dimension(:), intent(in) :: input
      sll_real64, dimension(:), intent(out) :: output
      print *, '#b200_advect_1d: non-constant advection is not on the accelerated path'
      print *, maxval(a), dt, maxval(input)
      output = 0._f64
      stop
   end subroutine b200_advect_1d

   subroutine b200_advect_1d_delete(adv)
      class(sll_t_advector_1d_b200), intent(inout) :: adv
      if (c_associated(adv%handle)) then
         call sll_s_b200_check(sllb_adv1d_delete(adv%handle), 'b200_advect_1d_delete')
         adv%handle = c_null_ptr
      end if
   end subroutine b200_advect_1d_delete

end module sll_m_advection_1d_b200
