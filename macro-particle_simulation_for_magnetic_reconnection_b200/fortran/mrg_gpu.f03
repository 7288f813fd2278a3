!***********************************************************************
!*  mrg_gpu.f03 -- ISO_C_BINDING shim that replaces subroutine fulmov   *
!*  of @mrg37-080A.f03 (F:1044-1390) by the B200 CUDA path.             *
!*                                                                     *
!*  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no        *
!*  Fortran compiler.  The logic it calls (upload-once particle         *
!*  residency, /fields/ upload, COMMON side effects) lives in           *
!*  csrc/mrg_host.cpp and is tested there through the same C entry      *
!*  points (tests/test_gpu_host_mirror.py).                             *
!*                                                                     *
!*  Usage in the reference (see INTEGRATION.md):                        *
!*   1. delete (or rename) the reference's own fulmov, F:1044-1390;     *
!*   2. compile this file with the same param_080A.h;                   *
!*   3. link  -lmrg_host -lmrg_fulmov  (plus the CUDA runtime);         *
!*   4. for size > 1, broadcast the NCCL id once before trans:          *
!*        if (rank == 0) ierr = mrg_host_unique_id(id)                  *
!*        call mpi_bcast(id,128,mpi_byte,0,mpi_comm_world,ierror)       *
!*        ierr = mrg_host_set_unique_id(id)                             *
!*   5. call mrg_pull_particles before restrt(iresrt=2) / diag1 and     *
!*      mrg_host_particles_changed after restrt(iresrt=1).              *
!*   6. (optional, saves 3/4 of the host->device field traffic) mark    *
!*      the three places where trans changes COMMON /fields/:           *
!*        call mrg_host_set_auto_fields(0)         ! once, before trans *
!*        call mrg_host_prefld_done()              ! after prefld F:759 *
!*          (the entry is repeated on the device; ..._mask(56) uploads) *
!*        call mrg_host_emfild_done()              ! after emfild F:771 *
!*          (only ex,ey,ez are uploaded; ..._mask(63) uploads all six)  *
!*        call mrg_host_fields_renewed()           ! after F:796-807    *
!***********************************************************************
      module mrg_gpu
      use, intrinsic :: iso_c_binding
      implicit none
!
!  Mirror of struct mrg_common_view (csrc/mrg_host.h): pointers into the
!  caller's COMMON blocks.
      type, bind(C) :: mrg_common_view
        integer(C_INT32_T) :: mx,my,mz
        type(C_PTR) :: ex,ey,ez,bx,by,bz,ex0,ey0,ez0,bx0,by0,bz0
        type(C_PTR) :: qix,qiy,qiz,qex,qey,qez,qi,qe
        type(C_PTR) :: it,ldec,ifilx,ifily,ifilz,nha
        type(C_PTR) :: xmax,ymax,zmax,dt,aimpl,adt,hdt,bxc,byc,bzc
        type(C_PTR) :: edec
        type(C_PTR) :: wkix,wkih
        type(C_PTR) :: zcent,ycent1,ycent2,Ez00
        type(C_PTR) :: ranfb
        type(C_PTR) :: io_pe
      end type mrg_common_view
!
      interface
        function mrg_host_bind (view,device) bind(C,name='mrg_host_bind')
          import :: C_INT, C_INT32_T, mrg_common_view
          integer(C_INT) :: mrg_host_bind
          type(mrg_common_view),intent(in) :: view
          integer(C_INT32_T),value :: device
        end function mrg_host_bind
!
!  The C++ drop-in with the argument list of F:1044 (all by reference).
        subroutine fulmov_gpu (x,y,z,vx,vy,vz,qmult,wmult,npr,ipc,ksp, &
                               ipar,size) bind(C,name='mrg_host_fulmov')
          import :: C_DOUBLE, C_INT32_T
          real(C_DOUBLE) :: x(*),y(*),z(*),vx(*),vy(*),vz(*)
          real(C_DOUBLE) :: qmult,wmult
          integer(C_INT32_T) :: npr,ipc,ksp,ipar,size
        end subroutine fulmov_gpu
!
        function mrg_host_pull_particles (ksp,x,y,z,vx,vy,vz,npr,ipar, &
                     size) bind(C,name='mrg_host_pull_particles')
          import :: C_INT, C_INT32_T, C_DOUBLE
          integer(C_INT) :: mrg_host_pull_particles
          integer(C_INT32_T),value :: ksp,npr,ipar,size
          real(C_DOUBLE) :: x(*),y(*),z(*),vx(*),vy(*),vz(*)
        end function mrg_host_pull_particles
!
        subroutine mrg_host_particles_changed (ksp) &
                     bind(C,name='mrg_host_particles_changed')
          import :: C_INT32_T
          integer(C_INT32_T),value :: ksp
        end subroutine mrg_host_particles_changed
!
        subroutine mrg_host_fields_changed () &
                     bind(C,name='mrg_host_fields_changed')
        end subroutine mrg_host_fields_changed
!
        subroutine mrg_host_set_auto_fields (on) &
                     bind(C,name='mrg_host_set_auto_fields')
          import :: C_INT32_T
          integer(C_INT32_T),value :: on
        end subroutine mrg_host_set_auto_fields
!
!  bit i-1 of mask = i-th member of common/fields/ was rewritten on the host
        subroutine mrg_host_fields_changed_mask (mask) &
                     bind(C,name='mrg_host_fields_changed_mask')
          import :: C_INT32_T
          integer(C_INT32_T),value :: mask
        end subroutine mrg_host_fields_changed_mask
!
!  the host has run the renewal loop ex0 <- ex (F:796-807)
        subroutine mrg_host_prefld_done () &
                     bind(C,name='mrg_host_prefld_done')
        end subroutine mrg_host_prefld_done
!
        subroutine mrg_host_emfild_done () &
                     bind(C,name='mrg_host_emfild_done')
        end subroutine mrg_host_emfild_done
!
        subroutine mrg_host_fields_renewed () &
                     bind(C,name='mrg_host_fields_renewed')
        end subroutine mrg_host_fields_renewed
!
        subroutine mrg_host_set_sort_interval (n) &
                     bind(C,name='mrg_host_set_sort_interval')
          import :: C_INT32_T
          integer(C_INT32_T),value :: n
        end subroutine mrg_host_set_sort_interval
!
        function mrg_host_unique_id (id) bind(C,name='mrg_host_unique_id')
          import :: C_INT, C_CHAR
          integer(C_INT) :: mrg_host_unique_id
          character(kind=C_CHAR) :: id(128)
        end function mrg_host_unique_id
!
        function mrg_host_set_unique_id (id) &
                     bind(C,name='mrg_host_set_unique_id')
          import :: C_INT, C_CHAR
          integer(C_INT) :: mrg_host_set_unique_id
          character(kind=C_CHAR) :: id(128)
        end function mrg_host_set_unique_id
!
!   A rank that stops alone leaves its peers inside NCCL / MPI: register a
!   routine that calls mpi_abort; the shim calls it instead of exit.
!     call mrg_host_set_abort (c_funloc(my_abort))   ! subroutine my_abort(code) bind(C); integer(C_INT),value :: code
        subroutine mrg_host_set_abort (fn) &
                     bind(C,name='mrg_host_set_abort')
          import :: C_FUNPTR
          type(C_FUNPTR),value :: fn
        end subroutine mrg_host_set_abort
      end interface
!
      end module mrg_gpu
!
!
!-----------------------------------------------------------------------
      subroutine fulmov (x,y,z,vx,vy,vz,qmult,wmult,npr,ipc,ksp,ipar,size)
!-----------------------------------------------------------------------
!  Same name, argument list and COMMON blocks as F:1044-1120.  The first
!  call binds the COMMON storage; every call forwards to the C++ host
!  mirror, which keeps the particles on the GPU and writes qix..qe,
!  wkix/wkih, edec(ldec,5..8) and the ranfp state where the reference does.
!
      use, intrinsic :: iso_c_binding
      use mrg_gpu
      implicit none
!
      include 'param_080A.h'
!
      real(C_DOUBLE),dimension(np0) :: x,y,z,vx,vy,vz
      real(C_DOUBLE) qmult,wmult
      integer(C_INT) npr,ipc,ksp,ipar,size
!
      real(C_DOUBLE),dimension(-2:mx+1,-1:my+1,-2:mz+1),target :: &
                                             ex,ey,ez,bx,by,bz,        &
                                             ex0,ey0,ez0,bx0,by0,bz0,  &
                                             qix,qiy,qiz,qex,qey,qez,  &
                                             emx,emy,emz,qi,qe
      common/fields/ ex,ey,ez,bx,by,bz,ex0,ey0,ez0,bx0,by0,bz0
      common/srimp7/ qix,qiy,qiz,qex,qey,qez,emx,emy,emz,qi,qe
!
      integer(C_INT),target :: it,ldec,iaver,ifilx,ifily,ifilz,iloadp, &
                     itermx,iterfx,itersx,nspec,nfwrt,npwrt,           &
                     nha,nplot,nhist
      common/parm1/  it,ldec,iaver,ifilx,ifily,ifilz,iloadp,         &
                     itermx,iterfx,itersx,nspec(4),nfwrt,npwrt,      &
                     nha,nplot,nhist
!
      real(C_DOUBLE),target :: xmax,ymax,zmax,hxi,hyi,hzi,           &
                     xmaxe,ymaxe,zmaxe,                              &
                     qspec,wspec,veth,te_by_ti,wce_by_wpe,thb,       &
                     rwd,pi,ait,t,dt,aimpl,adt,hdt,ahdt2,adtsq,      &
                     q0,qi0,qe0,aqi0,aqe0,epsln1,qwi,qwe,aqwi,aqwe,  &
                     qqwi,qqwe,vthx,vthz,vdr,vbeam,                  &
                     efe,efb,etot0,bxc,byc,bzc,vlima,vlimb,bmin,emin,&
                     edec
      common/parm2/  xmax,ymax,zmax,hxi,hyi,hzi,xmaxe,ymaxe,zmaxe,   &
                     qspec(4),wspec(4),veth,te_by_ti,wce_by_wpe,thb, &
                     rwd,pi,ait,t,dt,aimpl,adt,hdt,ahdt2,adtsq,      &
                     q0,qi0,qe0,aqi0,aqe0,epsln1,qwi,qwe,aqwi,aqwe,  &
                     qqwi,qqwe,vthx(4),vthz(4),vdr(4),vbeam(4),      &
                     efe,efb,etot0,bxc,byc,bzc,vlima,vlimb,bmin,emin,&
                     edec(3000,12)
!
      real(C_DOUBLE),target :: wkix,wkih,wkex,wkeh
      common/wkinel/ wkix,wkih,wkex,wkeh
!
      real(C_DOUBLE),target :: arb,zcent,ycent1,ycent2,Ez00
      common/profl/  arb,zcent,ycent1,ycent2,Ez00
!
      integer(C_INT),target :: ir_ranfp
      common/ranfb/  ir_ranfp
!
      integer(C_INT),target :: io_pe
      common/iope66/ io_pe
!
      type(mrg_common_view) :: v
      logical,save :: bound = .false.
      integer(C_INT) :: ierr,device
!
      if(.not.bound) then
        v%mx= mx
        v%my= my
        v%mz= mz
        v%ex = c_loc(ex)  ; v%ey = c_loc(ey)  ; v%ez = c_loc(ez)
        v%bx = c_loc(bx)  ; v%by = c_loc(by)  ; v%bz = c_loc(bz)
        v%ex0= c_loc(ex0) ; v%ey0= c_loc(ey0) ; v%ez0= c_loc(ez0)
        v%bx0= c_loc(bx0) ; v%by0= c_loc(by0) ; v%bz0= c_loc(bz0)
        v%qix= c_loc(qix) ; v%qiy= c_loc(qiy) ; v%qiz= c_loc(qiz)
        v%qex= c_loc(qex) ; v%qey= c_loc(qey) ; v%qez= c_loc(qez)
        v%qi = c_loc(qi)  ; v%qe = c_loc(qe)
        v%it = c_loc(it)  ; v%ldec= c_loc(ldec)
        v%ifilx= c_loc(ifilx) ; v%ifily= c_loc(ifily)
        v%ifilz= c_loc(ifilz) ; v%nha= c_loc(nha)
        v%xmax= c_loc(xmax) ; v%ymax= c_loc(ymax) ; v%zmax= c_loc(zmax)
        v%dt= c_loc(dt) ; v%aimpl= c_loc(aimpl)
        v%adt= c_loc(adt) ; v%hdt= c_loc(hdt)
        v%bxc= c_loc(bxc) ; v%byc= c_loc(byc) ; v%bzc= c_loc(bzc)
        v%edec= c_loc(edec)
        v%wkix= c_loc(wkix) ; v%wkih= c_loc(wkih)
        v%zcent= c_loc(zcent) ; v%ycent1= c_loc(ycent1)
        v%ycent2= c_loc(ycent2) ; v%Ez00= c_loc(Ez00)
        v%ranfb= c_loc(ir_ranfp)
        v%io_pe= c_loc(io_pe)
!
!  One GPU per rank of a node: ranks 0..7 -> devices 0..7.
        device= mod(ipar-1,8)
        ierr= mrg_host_bind (v,device)
        if(ierr.ne.0) stop 'mrg_host_bind failed'
        bound= .true.
      end if
!
      call fulmov_gpu (x,y,z,vx,vy,vz,qmult,wmult,npr,ipc,ksp,ipar,size)
!
      return
      end subroutine fulmov
!
!
!-----------------------------------------------------------------------
      subroutine mrg_pull_particles (x,y,z,vx,vy,vz,npr,ksp,ipar,size)
!-----------------------------------------------------------------------
!  Copy the rank's particles back into the host arrays (original l order)
!  before host code reads them: restrt(iresrt=2) F:9622-9668, diag1 F:7879.
!
      use, intrinsic :: iso_c_binding
      use mrg_gpu
      implicit none
      include 'param_080A.h'
!
      real(C_DOUBLE),dimension(np0) :: x,y,z,vx,vy,vz
      integer(C_INT) npr,ksp,ipar,size,ierr
!
      ierr= mrg_host_pull_particles (ksp,x,y,z,vx,vy,vz,npr,ipar,size)
      if(ierr.ne.0) stop 'mrg_host_pull_particles failed'
!
      return
      end subroutine mrg_pull_particles
