      module mod_tsadvc
c
c --- Drop-in replacement of HYCOM-src mod_tsadvc.F90 (module name, public
c --- name and call signature unchanged: `call tsadvc(m,n)` from HYCOM_Run,
c --- mod_hycom.F90:2535-2537; build dependency line Makefile:135-136).
c --- The advection itself runs on a B200 behind the C ABI of
c --- include/hycom_tsadvc_b200.h (libhycom_tsadvc_b200.so).
c
c --- NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran
c --- compiler.  It is the binding a maintainer adds; the same entry points
c --- are exercised through ctypes (hycom-src_b200/cabi.py) by the tests.
c
      use iso_c_binding
      use mod_xc         ! HYCOM communication interface (xcstop, mnproc, ...)
      implicit none
      private
      public :: tsadvc
c
      integer, parameter :: mxtrcr_c = 16       ! HYCOM_TSADVC_MXTRCR
c
      type, bind(c) :: tsadvc_dims              ! hycom_tsadvc_dims
        integer(c_int32_t) :: idm,jdm,kdm,nbdy, ii,jj, i0,j0, itdm,jtdm,
     &                        nreg, ipr,jpr, mproc,nproc, ntracr, device
      end type
      type, bind(c) :: tsadvc_params            ! hycom_tsadvc_params
        integer(c_int32_t) :: advtyp,advflg,btrmas,nhybrd,hybrid,
     &                        isopyc,mxlmy, nstep,diagno
        integer(c_int32_t) :: trcflg(mxtrcr_c)
        integer(c_int32_t) :: sigver
        real(c_double)     :: delt1,temdf2,temdfc,thbase,onemm
      end type
c
      interface
        integer(c_int) function hycom_tsadvc_create(dims,h)
     &           bind(c,name='hycom_tsadvc_create')
          import
          type(tsadvc_dims), intent(in) :: dims
          type(c_ptr), intent(out)      :: h
        end function
        integer(c_int) function hycom_tsadvc_set_static(h,
     &           scp2,scp2i,scuy,scvx,aspux,aspvy,ip,iu,iv)
     &           bind(c,name='hycom_tsadvc_set_static')
          import
          type(c_ptr), value :: h, scp2,scp2i,scuy,scvx,aspux,aspvy,
     &                          ip,iu,iv
        end function
        integer(c_int) function hycom_tsadvc_step(h,m,n,prm,
     &           temp,saln,th3d,tracer,dp,uflx,vflx,oneta,xmin,xmax)
     &           bind(c,name='hycom_tsadvc_step')
          import
          type(c_ptr), value        :: h
          integer(c_int32_t), value :: m,n
          type(tsadvc_params), intent(in) :: prm
          type(c_ptr), value :: temp,saln,th3d,tracer,dp,uflx,vflx,
     &                          oneta,xmin,xmax
        end function
        integer(c_int) function hycom_tsadvc_upload(h,field,ktr,
     &           tlev,k0,nk,host) bind(c,name='hycom_tsadvc_upload')
          import
          type(c_ptr), value        :: h, host
          integer(c_int32_t), value :: field,ktr,tlev,k0,nk
        end function
        integer(c_int) function hycom_tsadvc_download(h,field,ktr,
     &           tlev,k0,nk,host) bind(c,name='hycom_tsadvc_download')
          import
          type(c_ptr), value        :: h, host
          integer(c_int32_t), value :: field,ktr,tlev,k0,nk
        end function
        function hycom_tsadvc_last_error(h)
     &           bind(c,name='hycom_tsadvc_last_error')
          import
          type(c_ptr), value :: h
          type(c_ptr)        :: hycom_tsadvc_last_error
        end function
      end interface
c
      type(c_ptr), save :: handle = c_null_ptr
c
      contains
c
      subroutine tsadvc(m,n)
      use mod_cb_arrays  ! HYCOM saved arrays
      implicit none
      integer m,n
c
c --- same meaning as the reference: (:,:,:,n) holds t-1 on entry and t+1
c --- on exit, (:,:,:,m) holds t.  Collective over all tiles.
c
      type(tsadvc_dims)   :: d
      type(tsadvc_params) :: p
      real, save, allocatable, target :: xmin(:),xmax(:)
      type(c_ptr) :: ptrc
      integer rc,ktr,t
c
      include 'stmt_fns.h'   ! for sigver: the EOS family compiled in
c
      if     (.not.c_associated(handle)) then
c ---   first call: device mirrors + scratch (the analogue of the lazy
c ---   allocation of the module scratch in the reference advem)
        d%idm=idm; d%jdm=jdm; d%kdm=kdm; d%nbdy=nbdy
        d%ii=ii; d%jj=jj; d%i0=i0; d%j0=j0
        d%itdm=itdm; d%jtdm=jtdm; d%nreg=nreg
        d%ipr=ipr; d%jpr=jpr; d%mproc=mproc; d%nproc=nproc
        d%ntracr=ntracr
        d%device=0          ! one rank per GPU: CUDA_VISIBLE_DEVICES
        rc = hycom_tsadvc_create(d,handle)
        if (rc.ne.0) call b200_stop(rc)
        rc = hycom_tsadvc_set_static(handle,
     &         c_loc(scp2),c_loc(scp2i),c_loc(scuy),c_loc(scvx),
     &         c_loc(aspux),c_loc(aspvy),c_loc(ip),c_loc(iu),c_loc(iv))
        if (rc.ne.0) call b200_stop(rc)
        allocate( xmin(kdm),xmax(kdm) )
c ---   theta (isopycnic target densities) is constant in time and only
c ---   read by the diffusion part in exactly-isopycnal layers (k>nhybrd)
        if     (temdf2.gt.0.0 .and. nhybrd.lt.kdm) then
          rc = hycom_tsadvc_upload(handle,8,0,1,1,kdm,c_loc(theta))
          if (rc.ne.0) call b200_stop(rc)
        endif
      endif
c
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
c
c --- multi-tile host-array mode: keep the reference's xctilr calls on the
c --- HOST arrays here (halo width 5 of temp,saln,tracer both slots, uflx,
c --- vflx) so that the arrays handed over have valid halos; the device
c --- resident mode exchanges on the device instead (INTEGRATION.md).
c
c --- mxlmy: q2,q2l (0:kk+1, both slots) are not in the argument list of
c --- hycom_tsadvc_step; fill their mirrors (field ids 9, 10) around the call
      if     (mxlmy) then
        do t= 1,2
          rc = hycom_tsadvc_upload(handle, 9,0,t,1,kdm+2,
     &           c_loc(q2( 1-nbdy,1-nbdy,0,t)))
          if (rc.ne.0) call b200_stop(rc)
          rc = hycom_tsadvc_upload(handle,10,0,t,1,kdm+2,
     &           c_loc(q2l(1-nbdy,1-nbdy,0,t)))
          if (rc.ne.0) call b200_stop(rc)
        enddo
      endif
c
      ptrc = c_null_ptr
      if (ntracr.gt.0) ptrc = c_loc(tracer)
      rc = hycom_tsadvc_step(handle,m,n,p,
     &       c_loc(temp),c_loc(saln),c_loc(th3d),ptrc,
     &       c_loc(dp),c_loc(uflx),c_loc(vflx),c_loc(oneta),
     &       c_loc(xmin),c_loc(xmax))
      if (rc.ne.0) call b200_stop(rc)
      if     (mxlmy) then
c ---   the device copy returns whole slabs: the halo of q2,q2l(:,:,:,n)
c ---   comes back as the exchange left it (valid to width mbdy)
        rc = hycom_tsadvc_download(handle, 9,0,n,1,kdm+2,
     &         c_loc(q2( 1-nbdy,1-nbdy,0,n)))
        if (rc.ne.0) call b200_stop(rc)
        rc = hycom_tsadvc_download(handle,10,0,n,1,kdm+2,
     &         c_loc(q2l(1-nbdy,1-nbdy,0,n)))
        if (rc.ne.0) call b200_stop(rc)
      endif
c
c --- xmin/xmax now hold this tile's salinity range per layer when
c --- mod(nstep,3).eq.0 or diagno: xcminr/xcmaxr and the negative-salinity
c --- report of the reference follow unchanged.
      return
      end subroutine tsadvc
c
      subroutine b200_stop(rc)
      integer rc
c --- the reference prints on mnproc.eq.1 and calls xcstop('tsadvc')
      if     (mnproc.eq.1) then
        write(lp,'(/ a,i3 /)') 'error - hycom_tsadvc_b200 rc =',rc
        call flush(lp)
      endif
      call xcstop('tsadvc')
             stop 'tsadvc'
      end subroutine b200_stop
c
      end module mod_tsadvc
