      program ref_driver
!
! --- Standalone driver of the REFERENCE tsadvc(m,n) (HYCOM-src mod_tsadvc.F90,
! --- unmodified) on one tile: the pin of the CPU oracle (SURVEY.md section 8c,
! --- BASELINE.md "CPU baseline plan" item 3).
! ---
! --- Inputs are raw little-endian real*8 / integer*4 files written by
! --- fortran/ref_case.py from the SAME synthetic generator the tests use (a
! --- counter-based RNG keyed on the global (i,j,k): re-implementing it in
! --- Fortran would only add a second thing to trust).  Outputs are the
! --- (:,:,:,n) slabs of temp, saln, th3d and the tracers, which ref_case.py
! --- digests into tests/golden/from_reference.json in the format of
! --- tests/golden/tsadvc_golden.json.
! ---
! --- Serial RELO build (TYPE=one): mod_dimensions, mod_xc (mod_xc_sm.h),
! --- mod_cb_arrays, bigrid and mod_tsadvc are the reference's own files, compiled
! --- where they lie; mod_pipe is the stub next to this file.  No blkdat.input:
! --- the public scalars are set directly, as SURVEY.md 8c describes.
! ---
! --- NOT compiled in this repository's CI (no Fortran compiler in the image).
! --- usage: ref_driver <case directory>     (see fortran/build_ref.sh)
!
      use mod_xc         ! HYCOM communication interface
      use mod_cb_arrays  ! HYCOM saved arrays
      use mod_tsadvc     ! the reference module
      implicit none
!
      character*240 cdir
      integer       m,n,mapflg,nreg_in,i
      integer       ihdr(16)
      real          rhdr(8)
      real          t0,t1
      real*8        wtime
      external      wtime
!
      call getarg(1,cdir)
      call xcspmd                     ! lp=6, one tile, itdm=jtdm=-1 (RELO)
!
! --- case header: integers then reals, one per line (ref_case.py)
      open(unit=11,file=trim(cdir)//'/case.txt',form='formatted', &
           status='old',action='read')
      do i= 1,16
        read(11,*) ihdr(i)
      enddo
      do i= 1,8
        read(11,*) rhdr(i)
      enddo
      close(11)
      itdm   = ihdr(1);  jtdm = ihdr(2);  kdm = ihdr(3)
      idm    = itdm;     jdm  = jtdm;     kk  = kdm
      ii     = itdm;     jj   = jtdm;     i0  = 0;  j0 = 0
      nreg_in= ihdr(4)                ! what bigrid should find (checked below)
      ntracr = ihdr(5)
      advtyp = ihdr(6);  advflg = ihdr(7)
      btrmas = ihdr(8).ne.0
      nhybrd = ihdr(9);  hybrid = ihdr(10).ne.0
      isopyc = ihdr(11).ne.0
      mxlmy  = ihdr(12).ne.0
      nstep  = ihdr(13); diagno = ihdr(14).ne.0
      m      = ihdr(15); n      = ihdr(16)
      delt1  = rhdr(1);  temdf2 = rhdr(2);  temdfc = rhdr(3)
      thbase = rhdr(4)
!
! --- what cb_allocate needs besides the dimensions (blkdat.F90:645-647,
! --- :820-822, :1560, :2600-2604)
      mtracr = 0;  mstrcr = 0
      natm   = 2
      kknest = 1;  kkwall = 1
      if (ntracr.gt.0) kkwall = kdm
      kkmy25 = -1
      if (mxlmy) kkmy25 = kk
      flxflg = 0
      relaxt = .false.
      itest  = -99; jtest = -99
      call cb_allocate
      call rd_i4(trim(cdir)//'/trcflg.bin', trcflg, max(ntracr,1))
!
! --- topography -> masks, loop bounds, region type (bigrid.F90)
      call rd_r8(trim(cdir)//'/depths.bin', depths, (idm+2*nbdy)*(jdm+2*nbdy))
      mapflg = 0
      call bigrid(depths, mapflg, util1,util2,util3)
      if     (nreg.ne.nreg_in) then
        write(lp,*) 'error - bigrid found nreg =',nreg,' expected',nreg_in
        call xcstop('(ref_driver)')
      endif
!
! --- metrics (geopar.F90:311-340), halos valid
      call rd_r8(trim(cdir)//'/scp2.bin',  scp2,  (idm+2*nbdy)*(jdm+2*nbdy))
      call rd_r8(trim(cdir)//'/scp2i.bin', scp2i, (idm+2*nbdy)*(jdm+2*nbdy))
      call rd_r8(trim(cdir)//'/scuy.bin',  scuy,  (idm+2*nbdy)*(jdm+2*nbdy))
      call rd_r8(trim(cdir)//'/scvx.bin',  scvx,  (idm+2*nbdy)*(jdm+2*nbdy))
      call rd_r8(trim(cdir)//'/aspux.bin', aspux, (idm+2*nbdy)*(jdm+2*nbdy))
      call rd_r8(trim(cdir)//'/aspvy.bin', aspvy, (idm+2*nbdy)*(jdm+2*nbdy))
!
! --- state: both leapfrog slots of the advected fields, dp, mass fluxes
      call rd_r8(trim(cdir)//'/temp.bin', temp, (idm+2*nbdy)*(jdm+2*nbdy)*kdm*2)
      call rd_r8(trim(cdir)//'/saln.bin', saln, (idm+2*nbdy)*(jdm+2*nbdy)*kdm*2)
      call rd_r8(trim(cdir)//'/th3d.bin', th3d, (idm+2*nbdy)*(jdm+2*nbdy)*kdm*2)
      call rd_r8(trim(cdir)//'/dp.bin',   dp,   (idm+2*nbdy)*(jdm+2*nbdy)*kdm*2)
      call rd_r8(trim(cdir)//'/uflx.bin', uflx, (idm+2*nbdy)*(jdm+2*nbdy)*kdm)
      call rd_r8(trim(cdir)//'/vflx.bin', vflx, (idm+2*nbdy)*(jdm+2*nbdy)*kdm)
      call rd_r8(trim(cdir)//'/oneta.bin',oneta,(idm+2*nbdy)*(jdm+2*nbdy)*2)
      if     (ntracr.gt.0) then
        call rd_r8(trim(cdir)//'/tracer.bin', tracer, &
                   (idm+2*nbdy)*(jdm+2*nbdy)*kdm*2*ntracr)
      endif
      if     (temdf2.gt.0.0) then
        call rd_r8(trim(cdir)//'/theta.bin', theta, (idm+2*nbdy)*(jdm+2*nbdy)*kdm)
      endif
      if     (mxlmy) then
        call rd_r8(trim(cdir)//'/q2.bin',  q2,  (idm+2*nbdy)*(jdm+2*nbdy)*(kdm+2)*2)
        call rd_r8(trim(cdir)//'/q2l.bin', q2l, (idm+2*nbdy)*(jdm+2*nbdy)*(kdm+2)*2)
      endif
!
      t0 = wtime()
      call tsadvc(m,n)
      t1 = wtime()
      write(lp,'(a,f12.6,a,i10,a)') 'tsadvc: ',t1-t0,' s for ', &
        itdm*jtdm*kdm,' layer-cells'
!
      call wr_r8(trim(cdir)//'/out_temp.bin', temp, (idm+2*nbdy)*(jdm+2*nbdy)*kdm*2)
      call wr_r8(trim(cdir)//'/out_saln.bin', saln, (idm+2*nbdy)*(jdm+2*nbdy)*kdm*2)
      call wr_r8(trim(cdir)//'/out_th3d.bin', th3d, (idm+2*nbdy)*(jdm+2*nbdy)*kdm*2)
      if     (ntracr.gt.0) then
        call wr_r8(trim(cdir)//'/out_tracer.bin', tracer, &
                   (idm+2*nbdy)*(jdm+2*nbdy)*kdm*2*ntracr)
      endif
      call xcstop(' ')
!
      contains
!
      subroutine rd_r8(cfile, a, nw)
      character*(*), intent(in) :: cfile
      integer,       intent(in) :: nw
      real                      :: a(nw)   ! real*8 under -fdefault-real-8
      open(unit=12,file=cfile,form='unformatted',access='stream', &
           status='old',action='read')
      read(12) a
      close(12)
      end subroutine rd_r8
!
      subroutine rd_i4(cfile, ia, nw)
      character*(*), intent(in) :: cfile
      integer,       intent(in) :: nw
      integer                   :: ia(nw)
      open(unit=12,file=cfile,form='unformatted',access='stream', &
           status='old',action='read')
      read(12) ia
      close(12)
      end subroutine rd_i4
!
      subroutine wr_r8(cfile, a, nw)
      character*(*), intent(in) :: cfile
      integer,       intent(in) :: nw
      real                      :: a(nw)
      open(unit=12,file=cfile,form='unformatted',access='stream', &
           status='replace',action='write')
      write(12) a
      close(12)
      end subroutine wr_r8
!
      end program ref_driver
