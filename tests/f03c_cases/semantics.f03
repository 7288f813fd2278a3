!  Small units that exercise the Fortran semantics oracle/f03c.py has to keep (tests/test_f03c.py).
!  Written for the tests; nothing here comes from the reference.
!
      subroutine arith (iout,rout)
      use, intrinsic :: iso_c_binding
      implicit none
      include 'case_sizes.h'
      integer(C_INT) iout(10),i,j
      real(C_DOUBLE) rout(10),x,y
      real(C_float)  s
!
      i= 7
      j= -7
      iout(1)= i/2            ! integer division truncates toward zero
      iout(2)= j/2
      iout(3)= mod(j,3)       ! sign of the first argument
      iout(4)= 2**10
      iout(5)= int(-2.7d0)
      iout(6)= nint(2.5d0) +nint(-2.5d0)
      iout(7)= i/2*2          ! left to right
      iout(8)= nn
      x= 1.d0/3
      rout(1)= x
      rout(2)= 0.1            ! default-real literal: single precision value widened
      s= 0.1
      rout(3)= s*3            ! single-precision product, then widened
      rout(4)= 0.1d0*3
      y= 1.1d0
      rout(5)= y**3           ! y*y*y
      rout(6)= y**(-2)        ! 1/(y*y)
      rout(7)= 2.d0**0.5d0
      rout(8)= sign(3.d0,-0.d0) +abs(-2.5d0) +max(1.d0,2.d0,-3.d0) +min(4,2)
      rout(9)= float(i)/2     ! real*4 division
      rout(10)= (x +1.d20) -1.d20     ! parentheses kept: 0
      return
      end subroutine arith
!
!
      subroutine loops (iout)
      use, intrinsic :: iso_c_binding
      implicit none
      integer(C_INT) iout(10),i,k,n
!
      n= 0
      do i= 5,1            ! zero-trip
      n= n +1
      end do
      iout(1)= n
      iout(2)= i           ! the DO variable keeps its initial value
      n= 0
      do i= 10,1,-3        ! 10,7,4,1
      n= n +i
      end do
      iout(3)= n
      iout(4)= i           ! -2
      n= 0
      do 100 i= 1,3
      do 100 k= 1,2        ! shared terminal label
      n= n +i*k
  100 continue
      iout(5)= n
      n= 0
      i= 0
      do while (i.lt.4)
        i= i +1
        if(i.eq.2) cycle
        if(i.eq.4) exit
        n= n +i
      end do
      iout(6)= n
      n= 0
      i= 0
  200 i= i +1
      if(i.gt.3) go to 300
      n= n +10
      go to 200
  300 continue
      iout(7)= n
      if(n.gt.5 .and. .not.(n.eq.7)) iout(8)= 1
      return
      end subroutine loops
!
!
      subroutine blocks_a
      use, intrinsic :: iso_c_binding
      implicit none
      include 'case_sizes.h'
      real(C_DOUBLE) a(-2:mx+1,0:my),b(3)
      integer(C_INT) kk
      common/blk/ a,b,kk
      integer(C_INT) i,j
!
      do j= 0,my
      do i= -2,mx+1
      a(i,j)= 100*j +i
      end do
      end do
      b(1)= 1
      b(2)= 2
      b(3)= 3
      kk= 42
      return
      end subroutine blocks_a
!
!
      subroutine blocks_b (sout)
!  another view of the same storage sequence: one long vector + the tail
      use, intrinsic :: iso_c_binding
      implicit none
      include 'case_sizes.h'
      real(C_DOUBLE) v((mx+4)*(my+1)),c1,c2,c3,sout(4)
      integer(C_INT) kk
      common/blk/ v,c1,c2,c3,kk
!
      sout(1)= v(1)                  ! a(-2,0)
      sout(2)= v((mx+4)*my +3)       ! a(0,my)
      sout(3)= c1 +10*c2 +100*c3
      sout(4)= kk
      return
      end subroutine blocks_b
!
!
      subroutine byref (x,n,arr)
      use, intrinsic :: iso_c_binding
      implicit none
      real(C_DOUBLE) x,arr(3)
      integer(C_INT) n
      x= x +1
      n= n*2
      arr(2)= -arr(2)
      return
      end subroutine byref
!
      function twice (x)
      use, intrinsic :: iso_c_binding
      implicit none
      real(C_DOUBLE) twice,x
      twice= 2*x
      return
      end function twice
!
      subroutine caller (rout)
      use, intrinsic :: iso_c_binding
      implicit none
      real(C_DOUBLE) rout(8),x,w(5),twice
      integer(C_INT) n
      integer(C_INT),save :: ncall
      data ncall/0/
!
      x= 1
      n= 3
      w(1)= 1
      w(2)= 2
      w(3)= 3
      w(4)= 4
      w(5)= 5
      call byref (x,n,w(2))         ! array element actual: arr(1) is w(2)
      rout(1)= x
      rout(2)= n
      rout(3)= w(3)
      call byref (x+1,n,w)          ! expression actual: a temporary
      rout(4)= x
      rout(5)= n
      rout(6)= w(2)
      rout(7)= twice(x) +twice(1.5d0)
      ncall= ncall +1               ! SAVE: survives between calls
      rout(8)= ncall
      return
      end subroutine caller
!
!
      subroutine overlay (rout)
      use, intrinsic :: iso_c_binding
      implicit none
      real(C_DOUBLE) rout(3),w0(6),w1(2,3)
      equivalence (w0(1),w1(1,1))
      integer(C_INT) i
      do i= 1,6
      w0(i)= i
      end do
      rout(1)= w1(2,1)      ! w0(2)
      rout(2)= w1(1,3)      ! w0(5)
      w1(2,2)= -4
      rout(3)= w0(4)
      return
      end subroutine overlay
!
!
      subroutine twodoors (n,rout)
      use, intrinsic :: iso_c_binding
      implicit none
      real(C_DOUBLE) rout(2),acc
      integer(C_INT) n
      common/doors/ acc
      acc= acc +n
      rout(1)= acc
      return
!
      entry sidedoor
      acc= acc +1000
      return
      end subroutine twodoors
!
!
      subroutine ring (rank,rout)
!  every rank sends its number upward round the ring and adds what it receives; then a rank-ordered sum
      use, intrinsic :: iso_c_binding
      implicit none
      include 'mpif.h'
      include 'case_sizes.h'
      integer(C_INT) rank,up,dn,ierror
      real(C_DOUBLE) rout(3),sbuf(2),rbuf(2),part(1),tot(1)
      integer(kind=4),dimension(MPI_STATUS_SIZE) :: st1,st2
      integer(kind=4),dimension(1) :: rq1,rq2
!
      up= rank +1
      if(up.eq.npc) up= 0
      dn= rank -1
      if(dn.lt.0) dn= npc -1
      sbuf(1)= 10*(rank+1)
      sbuf(2)= -rank
      call mpi_irecv (rbuf,2,mpi_real8,dn,mpi_any_tag,mpi_comm_world,rq2,ierror)
      call mpi_isend (sbuf,2,mpi_real8,up,0,mpi_comm_world,rq1,ierror)
      call mpi_wait (rq1,st1,ierror)
      call mpi_wait (rq2,st2,ierror)
      rout(1)= rbuf(1)
      rout(2)= rbuf(2)
      part(1)= 0.1d0*(rank+1)
      call mpi_allreduce (part,tot,1,mpi_real8,mpi_sum,mpi_comm_world,ierror)
      rout(3)= tot(1)
      return
      end subroutine ring
!
!
      subroutine dump (iwhat)
!  iwhat = 2 writes two unformatted records, iwhat = 1 reads them back into other variables
      use, intrinsic :: iso_c_binding
      implicit none
      integer(C_INT) iwhat,i,j,n,kk(4),ll(4),nback
      real(C_DOUBLE) a(2,3),b(2,3),s,sback
      real(C_float)  r4(3),q4(3)
      common/dumpc/ b,sback,q4,ll,nback
!
      if(iwhat.eq.1) go to 100
      n= 7
      s= 2.5d0
      do i= 1,4
      kk(i)= 10*i
      end do
      do j= 1,3
      do i= 1,2
      a(i,j)= i +10*j
      end do
      r4(j)= 0.5*j
      end do
      open (unit=31,file='ignored'//'.bin',status='replace',form='unformatted')
      write(31) n,(kk(i),i=1,4),s
      write(31) ((a(i,j),i=1,2),j=1,3),r4
      close(31)
      return
!
  100 continue
      open (unit=31,file='ignored'//'.bin',form='unformatted')
      read(31) nback,(ll(i),i=1,4),sback
      read(31) b,q4
      close(31)
      return
      end subroutine dump
