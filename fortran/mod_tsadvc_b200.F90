      module mod_tsadvc
!
! --- Drop-in replacement of HYCOM-src mod_tsadvc.F90: same module name, same
! --- public name, same call `call tsadvc(m,n)` from HYCOM_Run
! --- (mod_hycom.F90:2535-2537), same build dependency line (Makefile:135-136).
! --- The advection itself runs on a B200 behind the C ABI of
! --- include/hycom_tsadvc_b200.h (libhycom_tsadvc_b200.so).
! --- Free-form source like the reference (its Makefile compiles .F90 as free form).
!
! --- The C entries are declared with ASSUMED-SIZE array dummies, so the
! --- mod_cb_arrays fields are passed as ordinary actual arguments (address of
! --- their first element, no copy: they are contiguous allocatables): no
! --- c_loc, hence no TARGET attribute to add to mod_cb_arrays.
!
! --- Multi-tile (MPI) builds: the library owns the halo exchange.  On the
! --- first call tile 1 draws a 128-byte id, mod_xc's communicator broadcasts
! --- it and every tile attaches (hycom_tsadvc_comm_init); after that
! --- hycom_tsadvc_step performs xctilr (:1829-1836, :2140-2151, :1186-1187),
! --- xcminr and xcmaxr (:2093-2094) on the device mirrors over NVLink.
! --- The host arrays only need what the reference needs on entry:
! --- "dp halo is up to date".
!
! --- Not compiled in this repository's CI (the build image has no Fortran
! --- compiler); fortran/build_ref.sh compiles it against stub modules where
! --- gfortran exists.  The same entry points are exercised through ctypes
! --- (hycom-src_b200/cabi.py) by the tests.
!
      use iso_c_binding
      use mod_xc         ! HYCOM communication interface (xcstop, mnproc, ...)
      implicit none
      private
      public :: tsadvc
!
      integer, parameter :: mxtrcr_c = 16       ! HYCOM_TSADVC_MXTRCR
!
      type, bind(c) :: tsadvc_dims              ! hycom_tsadvc_dims
        integer(c_int32_t) :: idm,jdm,kdm,nbdy, ii,jj, i0,j0, itdm,jtdm, &
                              nreg, ipr,jpr, mproc,nproc, ntracr, device
      end type
      type, bind(c) :: tsadvc_params            ! hycom_tsadvc_params
        integer(c_int32_t) :: advtyp,advflg,btrmas,nhybrd,hybrid, &
                              isopyc,mxlmy, nstep,diagno
        integer(c_int32_t) :: trcflg(mxtrcr_c)
        integer(c_int32_t) :: sigver
        real(c_double)     :: delt1,temdf2,temdfc,thbase,onemm
      end type
!
      interface
        integer(c_int) function hycom_tsadvc_create(dims,h) &
                 bind(c,name='hycom_tsadvc_create')
          import
          type(tsadvc_dims), intent(in) :: dims
          type(c_ptr), intent(out)      :: h
        end function
        integer(c_int) function hycom_tsadvc_set_static(h, &
                 scp2,scp2i,scuy,scvx,aspux,aspvy,ip,iu,iv) &
                 bind(c,name='hycom_tsadvc_set_static')
          import
          type(c_ptr), value :: h
          real(c_double),     intent(in) :: scp2(*),scp2i(*),scuy(*), &
                                            scvx(*),aspux(*),aspvy(*)
          integer(c_int32_t), intent(in) :: ip(*),iu(*),iv(*)
        end function
        integer(c_int) function hycom_tsadvc_step(h,m,n,prm, &
                 temp,saln,th3d,tracer,dp,uflx,vflx,oneta,xmin,xmax) &
                 bind(c,name='hycom_tsadvc_step')
          import
          type(c_ptr), value        :: h
          integer(c_int32_t), value :: m,n
          type(tsadvc_params), intent(in) :: prm
          real(c_double) :: temp(*),saln(*),th3d(*),tracer(*)
          real(c_double), intent(in) :: dp(*),uflx(*),vflx(*),oneta(*)
          real(c_double) :: xmin(*),xmax(*)
        end function
        integer(c_int) function hycom_tsadvc_upload(h,field,ktr, &
                 tlev,k0,nk,host) bind(c,name='hycom_tsadvc_upload')
          import
          type(c_ptr), value        :: h
          integer(c_int32_t), value :: field,ktr,tlev,k0,nk
          real(c_double), intent(in) :: host(*)
        end function
        integer(c_int) function hycom_tsadvc_download(h,field,ktr, &
                 tlev,k0,nk,host) bind(c,name='hycom_tsadvc_download')
          import
          type(c_ptr), value        :: h
          integer(c_int32_t), value :: field,ktr,tlev,k0,nk
          real(c_double) :: host(*)
        end function
        integer(c_int) function hycom_tsadvc_comm_unique_id(id) &
                 bind(c,name='hycom_tsadvc_comm_unique_id')
          import
          character(kind=c_char) :: id(128)
        end function
        integer(c_int) function hycom_tsadvc_comm_init(h,id) &
                 bind(c,name='hycom_tsadvc_comm_init')
          import
          type(c_ptr), value :: h
          character(kind=c_char), intent(in) :: id(128)
        end function
      end interface
!
      type(c_ptr), save :: handle = c_null_ptr
      real, save, allocatable :: xmin(:),xmax(:)
      real, save :: tr0(1)   ! stands for tracer when ntracr=0
!
      contains
!
      subroutine tsadvc(m,n)
      use mod_cb_arrays  ! HYCOM saved arrays
      implicit none
      integer m,n
!
! --- same meaning as the reference: (:,:,:,n) holds t-1 on entry and t+1
! --- on exit, (:,:,:,m) holds t.  Collective over all tiles.
!
      type(tsadvc_dims)   :: d
      type(tsadvc_params) :: p
      integer rc,ktr,t,i,j,k
      real    sminn,smaxx
      character(kind=c_char) :: id(128)
      real    idr(16)        ! the 128 id bytes as 16 reals (-fdefault-real-8) for xcastr
!
      include 'stmt_fns.h'   ! for sigver: the EOS family compiled in
!
      if     (.not.c_associated(handle)) then
! ---   first call: device mirrors + scratch (the analogue of the lazy
! ---   allocation of the module scratch in the reference advem)
        d%idm=idm; d%jdm=jdm; d%kdm=kdm; d%nbdy=nbdy
        d%ii=ii; d%jj=jj; d%i0=i0; d%j0=j0
        d%itdm=itdm; d%jtdm=jtdm; d%nreg=nreg
        d%ipr=ipr; d%jpr=jpr; d%mproc=mproc; d%nproc=nproc
        d%ntracr=ntracr
        d%device=0          ! one rank per GPU: CUDA_VISIBLE_DEVICES
        rc = hycom_tsadvc_create(d,handle)
        if (rc.ne.0) call b200_stop(rc)
        rc = hycom_tsadvc_set_static(handle, &
               scp2,scp2i,scuy,scvx,aspux,aspvy,ip,iu,iv)
        if (rc.ne.0) call b200_stop(rc)
        allocate( xmin(kdm),xmax(kdm) )
        if     (ipr*jpr.gt.1) then
! ---     the library's communicator: mnproc = mproc + ipr*(nproc-1) is
! ---     the rank it expects (mod_xc_mp.h:2830 idproc).  mpi_comm_hycom
! ---     is private to mod_xc, so the id travels through mod_xc's own
! ---     broadcast, xcastr (mod_xc_mp.h:831), as 16 reals.
          idr(:) = 0.0
          if     (mnproc.eq.1) then
            rc = hycom_tsadvc_comm_unique_id(id)
            if (rc.ne.0) call b200_stop(rc)
            idr = transfer(id,idr)
          endif
          call xcastr(idr, 1)
          id  = transfer(idr,id)
          rc = hycom_tsadvc_comm_init(handle,id)
          if (rc.ne.0) call b200_stop(rc)
        endif
! ---   theta (isopycnic target densities) is constant in time and only
! ---   read by the diffusion part in exactly-isopycnal layers (k>nhybrd)
        if     (temdf2.gt.0.0 .and. nhybrd.lt.kdm) then
          rc = hycom_tsadvc_upload(handle,8,0,1,1,kdm,theta)
          if (rc.ne.0) call b200_stop(rc)
        endif
      endif
!
      p%advtyp=advtyp; p%advflg=advflg
      p%btrmas=merge(1,0,btrmas); p%nhybrd=nhybrd
      p%hybrid=merge(1,0,hybrid); p%isopyc=merge(1,0,isopyc)
      p%mxlmy=merge(1,0,mxlmy); p%nstep=nstep
      p%diagno=merge(1,0,diagno)
      p%trcflg(:)=0
      do ktr= 1,ntracr
        p%trcflg(ktr)=trcflg(ktr)
      enddo
      p%delt1=delt1; p%temdf2=temdf2; p%temdfc=temdfc
      p%thbase=thbase; p%onemm=onemm
      p%sigver=sigver     ! stmt_fns.h: the EOS this executable was built with
!
! --- mxlmy: q2,q2l (0:kk+1, both slots) are not in the argument list of
! --- hycom_tsadvc_step; fill their mirrors (field ids 9, 10) around the call
      if     (mxlmy) then
        do t= 1,2
          rc = hycom_tsadvc_upload(handle, 9,0,t,1,kdm+2, &
                 q2( 1-nbdy,1-nbdy,0,t))
          if (rc.ne.0) call b200_stop(rc)
          rc = hycom_tsadvc_upload(handle,10,0,t,1,kdm+2, &
                 q2l(1-nbdy,1-nbdy,0,t))
          if (rc.ne.0) call b200_stop(rc)
        enddo
      endif
!
! --- the call: copies in, exchanges halos, advects (diffuses), copies
! --- (:,:,:,n) back on 1:ii,1:jj; xmin/xmax are the GLOBAL salinity range
! --- per layer (xcminr/xcmaxr done on the device) when mod(nstep,3).eq.0
! --- or diagno
      if     (ntracr.gt.0) then
        rc = hycom_tsadvc_step(handle,m,n,p, &
               temp,saln,th3d,tracer, dp,uflx,vflx,oneta, xmin,xmax)
      else
        rc = hycom_tsadvc_step(handle,m,n,p, &
               temp,saln,th3d,tr0,    dp,uflx,vflx,oneta, xmin,xmax)
      endif
      if (rc.ne.0) call b200_stop(rc)
      if     (mxlmy) then
! ---   the device copy returns whole slabs: the halo of q2,q2l(:,:,:,n)
! ---   comes back as the exchange left it (valid to width mbdy)
        rc = hycom_tsadvc_download(handle, 9,0,n,1,kdm+2, &
               q2( 1-nbdy,1-nbdy,0,n))
        if (rc.ne.0) call b200_stop(rc)
        rc = hycom_tsadvc_download(handle,10,0,n,1,kdm+2, &
               q2l(1-nbdy,1-nbdy,0,n))
        if (rc.ne.0) call b200_stop(rc)
      endif
!
! --- check for negative scalar fields (mod_tsadvc.F90:2090-2132)
!
      if     (mod(nstep,3).eq.0 .or. diagno) then
        do k= 1,kk
          sminn=xmin(k)
          smaxx=xmax(k)
!
          if (sminn.lt.0.0) then
            do j=1,jj
              do i=1,ii
                if (ip(i,j).ne.0) then
                  if (saln(i,j,k,n).eq.sminn) then
                    write (lp,'(i9,a,2i6,i4,a,f10.2)')  &
                      nstep,' i,j,k =',i+i0,j+j0,k, &
                      ' neg. saln after advem call ', &
                      saln(i,j,k,n)
                  endif !sminn
                endif !ip
              enddo !i
            enddo !j
            call xcsync(flush_lp)
          endif
!
          if (diagno) then
            if     (mnproc.eq.1) then
            if     (sminn.le.smaxx) then
              write (lp,'(i9,i4, a,2f9.3, a,1pe9.2,a)') &
                nstep,k, &
                ' min/max of s after advection:',sminn,smaxx, &
                '   (range:',smaxx-sminn,')'
            else
              write (lp,'(i9,i4, a,a)') &
                nstep,k, &
                ' min/max of s after advection:',' N/A (thin layer)'
            endif !normal:thin layer
            call flush(lp)
            endif
          endif
        enddo !k
      endif !every 3 time steps or diagno
      return
      end subroutine tsadvc
!
      subroutine b200_stop(rc)
      integer rc
! --- the reference prints on mnproc.eq.1 and calls xcstop('tsadvc')
! --- (:1817-1825); rc 4 = nbdy too small, 5 = bad advtyp (xcstop('advem'))
      if     (mnproc.eq.1) then
        write(lp,'(/ a,i3 /)') 'error - hycom_tsadvc_b200 rc =',rc
        call flush(lp)
      endif
      if     (rc.eq.5) then
        call xcstop('advem')
               stop 'advem'
      endif
      call xcstop('tsadvc')
             stop 'tsadvc'
      end subroutine b200_stop
!
      end module mod_tsadvc
