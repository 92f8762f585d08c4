      module mod_pipe
!
! --- Stub of HYCOM's debugging interface (mod_pipe.F90) for the standalone
! --- reference driver (fortran/ref_driver.F90): mod_tsadvc only needs lpipe and
! --- the three comparison entries below, which are no-ops when lpipe is false
! --- (mod_tsadvc.F90:193-195, :2049-2058, :2088).  The real module drags in
! --- mod_tides, mod_stokes and the whole comparall machinery.
!
      use mod_xc  ! HYCOM communication interface
      implicit none
      logical, save, public :: lpipe = .false.
!
      contains
!
      subroutine pipe_compare_sym1(field,mask,what)
      real,    dimension (1-nbdy:idm+nbdy,1-nbdy:jdm+nbdy), &
               intent(in) :: field
      integer, dimension (1-nbdy:idm+nbdy,1-nbdy:jdm+nbdy), &
               intent(in) :: mask
      character*12, intent(in) :: what
      return
      end subroutine pipe_compare_sym1
!
      subroutine pipe_compare_sym2(field_u,mask_u,what_u, &
                                   field_v,mask_v,what_v)
      real,    dimension (1-nbdy:idm+nbdy,1-nbdy:jdm+nbdy), &
               intent(in) :: field_u,field_v
      integer, dimension (1-nbdy:idm+nbdy,1-nbdy:jdm+nbdy), &
               intent(in) :: mask_u,mask_v
      character*12, intent(in) :: what_u,what_v
      return
      end subroutine pipe_compare_sym2
!
      subroutine pipe_comparall(m,n, cinfo)
      integer, intent(in) :: m,n
      character*12, intent(in) :: cinfo
      return
      end subroutine pipe_comparall
!
      end module mod_pipe
