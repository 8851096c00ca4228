! -
! cans_b200 -- ISO_C_BINDING shim that plugs the B200-native Poisson/Helmholtz solver into an UNCHANGED CaNS host.
!
! Drop these modules in place of src/fft.f90 (mod_fft), src/solver.f90 / src/solver_gpu.f90 (mod_solver / mod_solver_gpu)
! and src/workspaces.f90 (mod_workspaces); main.f90, rk.f90, mom.f90, initsolver.f90, solve_helmholtz.f90, bound.f90 and the
! input.nml logic stay as they are.  Every public name, argument list and `arrplan` type below is the reference's
! (file:line cited at each routine).  Build: add this file to the source list instead of the four it replaces and link
!   -L$(CANS_B200)/cans_b200/lib -lcans_b200 -lcudart            (configs/libs.mk)
! The C side is include/cans_b200.h.  NOTE: the image this repo is built in has no Fortran compiler, so this file is
! written against the standard, not compiled here; the very same C entry points, with the same argument meaning, are
! exercised through ctypes by tests/test_gpu_boundary.py (test_reference_call_sequence et al.).
! -
module mod_cansb200
  use, intrinsic :: iso_c_binding
  use mod_types, only: rp
  implicit none
  public
  type(c_ptr), save :: ctx = c_null_ptr      ! one context per rank, created by cansb200_setup (called from initmpi's place)
#if defined(_OPENACC)
  integer(c_int), parameter :: MEM_KIND = 1  ! device pointers (host_data use_device), zero copies
  integer(c_int), parameter :: LAMBDA_ORDER = 1 ! initsolver's _OPENACC eigenvalue order (src/initsolver.f90:98-117)
#else
  integer(c_int), parameter :: MEM_KIND = 0  ! host arrays: the library does H2D / D2H itself
  integer(c_int), parameter :: LAMBDA_ORDER = 0
#endif
  interface
    integer(c_int) function cansb200_init(ctx,ng,dims,ipencil_axis,rank,nranks,nccl_id,is_fp32) bind(C,name='cansb200_init')
      import; type(c_ptr) :: ctx; integer(c_int) :: ng(3),dims(2)
      integer(c_int), value :: ipencil_axis,rank,nranks,is_fp32; type(c_ptr), value :: nccl_id
    end function
    integer(c_int) function cansb200_finalize(ctx) bind(C,name='cansb200_finalize')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function cansb200_ctx_set(ctx,what,val) bind(C,name='cansb200_ctx_set')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: what,val
    end function
    integer(c_int) function cansb200_get_extents(ctx,n,lo,n_z,lo_z) bind(C,name='cansb200_get_extents')
      import; type(c_ptr), value :: ctx; integer(c_int) :: n(3),lo(3),n_z(3),lo_z(3)
    end function
    integer(c_int) function cansb200_dist_blob_size() bind(C,name='cansb200_dist_blob_size')
      import
    end function
    integer(c_int) function cansb200_dist_export(ctx,blob) bind(C,name='cansb200_dist_export')
      import; type(c_ptr), value :: ctx,blob
    end function
    integer(c_int) function cansb200_dist_connect(ctx,blobs) bind(C,name='cansb200_dist_connect')
      import; type(c_ptr), value :: ctx,blobs
    end function
    integer(c_int) function cansb200_fftini(ctx,bcxy,c_or_f_xy,normfft,id) bind(C,name='cansb200_fftini')
      import; type(c_ptr), value :: ctx; character(kind=c_char) :: bcxy(4),c_or_f_xy(2); real(c_double) :: normfft; integer(c_int) :: id
    end function
    integer(c_int) function cansb200_fftend(ctx,id) bind(C,name='cansb200_fftend')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: id
    end function
    integer(c_int) function cansb200_solver(ctx,id,bc,c_or_f,p,n,nhalo,normfft,lambdaxy,a,b,c,lambda_order,mem_kind,stream) &
                   bind(C,name='cansb200_solver')
      import; type(c_ptr), value :: ctx,p,lambdaxy,a,b,c,stream; integer(c_int), value :: id,nhalo,lambda_order,mem_kind
      character(kind=c_char) :: bc(6),c_or_f(3); integer(c_int) :: n(3); real(c_double), value :: normfft
    end function
    integer(c_int) function cansb200_solver_fillps(ctx,id,bc,c_or_f,p,n,nhalo,normfft,lambdaxy,a,b,c,lambda_order,dli,dzfi,dti,u,v,w, &
                                                   is_bound,have,rhsb,stream) bind(C,name='cansb200_solver_fillps')
      import; type(c_ptr), value :: ctx,p,lambdaxy,a,b,c,dzfi,u,v,w,stream; integer(c_int), value :: id,nhalo,lambda_order
      character(kind=c_char) :: bc(6),c_or_f(3); integer(c_int) :: n(3),is_bound(6),have(3)
      real(c_double), value :: normfft,dti; real(c_double) :: dli(3),rhsb(6)
    end function
    integer(c_int) function cansb200_solve_z_bc(ctx,bcz,c_or_f_z,p,n,nhalo,norm,a,b,c,mem_kind,stream) bind(C,name='cansb200_solve_z_bc')
      import; type(c_ptr), value :: ctx,p,a,b,c,stream; character(kind=c_char) :: bcz(2); character(kind=c_char), value :: c_or_f_z
      integer(c_int) :: n(3); integer(c_int), value :: nhalo,mem_kind; real(c_double), value :: norm
    end function
    integer(c_int) function cansb200_get_work(ctx,which,ptr,nelem) bind(C,name='cansb200_get_work')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: which; type(c_ptr) :: ptr; integer(c_size_t) :: nelem
    end function
  end interface
contains
  ! called where initmpi sets up cuDecomp / 2DECOMP (src/initmpi.f90:84-146); dims = [1,nranks] (z slabs)
  subroutine cansb200_setup(ng,dims,ipencil_axis,myid,nranks,is_poisson_dtdma)
    use mpi
    integer, intent(in) :: ng(3),dims(2),ipencil_axis,myid,nranks
    logical, intent(in) :: is_poisson_dtdma
    integer :: nb,ierr
    character(kind=c_char), allocatable, target :: mine(:),everyone(:)
    if(cansb200_init(ctx,ng,dims,ipencil_axis,myid,nranks,c_null_ptr,merge(1,0,rp == kind(1.0))) /= 0) error stop 'cansb200_init'
    if(is_poisson_dtdma) then
      if(cansb200_ctx_set(ctx,9,1) /= 0) error stop 'cansb200_ctx_set(DTDMA)'   ! CANSB200_CTX_DTDMA, before any plan
    end if
#if !defined(_OPENACC)
    if(cansb200_ctx_set(ctx,7,1) /= 0) error stop 'cansb200_ctx_set(PIN_HOST)'   ! page-lock the host's p on first use
#endif
    if(nranks > 1) then   ! the exchange runs over CUDA-IPC peer mappings: allgather one small blob per rank
      nb = cansb200_dist_blob_size()
      allocate(mine(nb),everyone(nb*nranks))
      if(cansb200_dist_export(ctx,c_loc(mine)) /= 0) error stop 'cansb200_dist_export'
      call MPI_ALLGATHER(mine,nb,MPI_BYTE,everyone,nb,MPI_BYTE,MPI_COMM_WORLD,ierr)
      if(cansb200_dist_connect(ctx,c_loc(everyone)) /= 0) error stop 'cansb200_dist_connect'
    end if
  end subroutine
  function cuda_stream_of_queue_1() result(s)   ! every kernel of the reference runs on OpenACC queue 1 (src/workspaces.f90:101-106)
#if defined(_OPENACC)
    use openacc
#endif
    type(c_ptr) :: s
    s = c_null_ptr
#if defined(_OPENACC)
    s = transfer(acc_get_cuda_stream(1),s)
#endif
  end function
end module mod_cansb200
!
module mod_fft   ! replaces src/fft.f90: fftini (:25-209), fftend (:211-245); `fft` itself is only called from inside solver
  use, intrinsic :: iso_c_binding
  use mod_types, only: rp
  use mod_cansb200
  implicit none
  private
  public fftini,fftend
contains
  subroutine fftini(ng,n_x,n_y,bcxy,c_or_f,arrplan,normfft)
    integer , intent(in), dimension(3) :: ng,n_x,n_y
    character(len=1), intent(in), dimension(0:1,2) :: bcxy
    character(len=1), intent(in), dimension(2) :: c_or_f
#if !defined(_OPENACC) || defined(_USE_HIP)
    type(C_PTR), intent(out), dimension(2,2) :: arrplan
#else
    integer    , intent(out), dimension(2,2) :: arrplan
#endif
    real(rp), intent(out) :: normfft
    character(kind=c_char) :: b4(4),cf(2)
    real(c_double) :: nf
    integer(c_int) :: id
    b4 = [bcxy(0,1),bcxy(1,1),bcxy(0,2),bcxy(1,2)]; cf = c_or_f
    if(cansb200_fftini(ctx,b4,cf,nf,id) /= 0) error stop 'cansb200_fftini'
    normfft = real(nf,rp)
#if !defined(_OPENACC) || defined(_USE_HIP)
    arrplan(:,:) = c_null_ptr; arrplan(1,1) = transfer(int(id,c_intptr_t),arrplan(1,1))   ! the id travels in the C_PTR
#else
    arrplan(:,:) = 0; arrplan(1,1) = id
#endif
  end subroutine
  subroutine fftend(arrplan)
#if !defined(_OPENACC) || defined(_USE_HIP)
    type(C_PTR), intent(in), dimension(:,:) :: arrplan
    if(cansb200_fftend(ctx,int(transfer(arrplan(1,1),0_c_intptr_t),c_int)) /= 0) error stop 'cansb200_fftend'
#else
    integer    , intent(in), dimension(:,:) :: arrplan
    if(cansb200_fftend(ctx,arrplan(1,1)) /= 0) error stop 'cansb200_fftend'
#endif
  end subroutine
end module mod_fft
!
#if defined(_OPENACC)
module mod_solver_gpu   ! replaces src/solver_gpu.f90: solver_gpu (:34-276), solver_gaussel_z_gpu (:956-1105)
#else
module mod_solver       ! replaces src/solver.f90: solver (:17-112), solver_gaussel_z (:547-616)
#endif
  use, intrinsic :: iso_c_binding
  use mod_types, only: rp
  use mod_cansb200
  implicit none
  private
#if defined(_OPENACC)
  public solver_gpu,solver_gaussel_z_gpu,fillps_solver_gpu
#else
  public solver,solver_gaussel_z
#endif
contains
#if defined(_OPENACC)
  subroutine solver_gpu(n,ng,arrplan,normfft,lambdaxy,a,b,c,bc,c_or_f,p,is_dtdma_update,aa_z,cc_z)
#else
  subroutine solver(n,ng,arrplan,normfft,lambdaxy,a,b,c,bc,c_or_f,p,is_dtdma_update,aa_z,cc_z)
#endif
    integer , intent(in), dimension(3) :: n,ng
#if !defined(_OPENACC) || defined(_USE_HIP)
    type(C_PTR), intent(in), dimension(2,2) :: arrplan
#else
    integer    , intent(in), dimension(2,2) :: arrplan
#endif
    real(rp), intent(in) :: normfft
    real(rp), intent(in), target, dimension(:,:) :: lambdaxy
    real(rp), intent(in), target, dimension(:) :: a,b,c
    character(len=1), intent(in), dimension(0:1,3) :: bc
    character(len=1), intent(in), dimension(3) :: c_or_f
    real(rp), intent(inout), target, contiguous, dimension(0:,0:,0:) :: p
    logical , intent(inout), optional :: is_dtdma_update               ! the library caches by content: nothing to do here
    real(rp), intent(inout), dimension(:,:,:), optional :: aa_z,cc_z   ! idem
    character(kind=c_char) :: b6(6),cf(3)
    integer(c_int) :: istat,id
    b6 = [bc(0,1),bc(1,1),bc(0,2),bc(1,2),bc(0,3),bc(1,3)]; cf = c_or_f
#if !defined(_OPENACC) || defined(_USE_HIP)
    id = int(transfer(arrplan(1,1),0_c_intptr_t),c_int)
#else
    id = arrplan(1,1)
#endif
    !$acc host_data use_device(p,lambdaxy,a,b,c)
    istat = cansb200_solver(ctx,id,b6,cf,c_loc(p),n,1,real(normfft,c_double),c_loc(lambdaxy),c_loc(a),c_loc(b),c_loc(c), &
                            LAMBDA_ORDER,MEM_KIND,cuda_stream_of_queue_1())
    !$acc end host_data
    if(istat /= 0) error stop 'cansb200_solver'   ! the reference aborts too (src/fft.f90:525,696)
  end subroutine
#if defined(_OPENACC)
  ! OPTIONAL (device builds): the three calls of the pressure-correction step (src/main.f90:465-467)
  !     call fillps(n,dli,dzfi,dtrki,u,v,w,pp)
  !     call updt_rhs_b(['c','c','c'],cbcpre,n,is_bound,rhsbp%x,rhsbp%y,rhsbp%z,pp)
  !     call solver(n,ng,arrplanp,normfftp,lambdaxyp,ap,bp,cp,cbcpre,['c','c','c'],pp)
  ! as ONE call -- the only edit to main.f90 this file ever asks for, and only if the host wants the saving: the forward x
  ! transform evaluates fillps and the wall terms at load time (pp is not written and read back: -16 B/point of HBM traffic;
  ! C3 on one B200: 6.35 -> 4.77 ms for the step).  Arguments = solver's, then fillps's, then updt_rhs_b's; rhsbp%x/y/z are
  ! uniform per wall (src/initsolver.f90:189-232), so their first element per side is passed.
  subroutine fillps_solver_gpu(n,ng,arrplan,normfft,lambdaxy,a,b,c,bc,c_or_f,dli,dzfi,dti,u,v,w,is_bound,rhsbx,rhsby,rhsbz,p)
    integer , intent(in), dimension(3) :: n,ng
#if defined(_USE_HIP)
    type(C_PTR), intent(in), dimension(2,2) :: arrplan
#else
    integer    , intent(in), dimension(2,2) :: arrplan
#endif
    real(rp), intent(in) :: normfft,dti
    real(rp), intent(in), target, dimension(:,:) :: lambdaxy
    real(rp), intent(in), target, dimension(:) :: a,b,c
    character(len=1), intent(in), dimension(0:1,3) :: bc
    character(len=1), intent(in), dimension(3) :: c_or_f
    real(rp), intent(in), dimension(3) :: dli
    real(rp), intent(in), target, dimension(0:) :: dzfi
    real(rp), intent(in), target, contiguous, dimension(0:,0:,0:) :: u,v,w
    logical , intent(in), dimension(0:1,3) :: is_bound
    real(rp), intent(in), dimension(:,:,0:) :: rhsbx,rhsby,rhsbz
    real(rp), intent(inout), target, contiguous, dimension(0:,0:,0:) :: p
    character(kind=c_char) :: b6(6),cf(3)
    integer(c_int) :: istat,id,isb(6),have(3)
    real(c_double) :: rh(6),dl3(3)
    b6 = [bc(0,1),bc(1,1),bc(0,2),bc(1,2),bc(0,3),bc(1,3)]; cf = c_or_f
#if defined(_USE_HIP)
    id = int(transfer(arrplan(1,1),0_c_intptr_t),c_int)
#else
    id = arrplan(1,1)
#endif
    isb = merge(1,0,[is_bound(0,1),is_bound(1,1),is_bound(0,2),is_bound(1,2),is_bound(0,3),is_bound(1,3)]); have = 1
    ! rhsb* are filled on the host by initsolver and only copied to the device afterwards: the host copies are current
    rh = real([rhsbx(1,1,0),rhsbx(1,1,1),rhsby(1,1,0),rhsby(1,1,1),rhsbz(1,1,0),rhsbz(1,1,1)],c_double); dl3 = real(dli,c_double)
    !$acc host_data use_device(p,lambdaxy,a,b,c,dzfi,u,v,w)
    istat = cansb200_solver_fillps(ctx,id,b6,cf,c_loc(p),n,1,real(normfft,c_double),c_loc(lambdaxy),c_loc(a),c_loc(b),c_loc(c), &
                                   LAMBDA_ORDER,dl3,c_loc(dzfi),real(dti,c_double),c_loc(u),c_loc(v),c_loc(w),isb,have,rh, &
                                   cuda_stream_of_queue_1())
    !$acc end host_data
    if(istat /= 0) error stop 'cansb200_solver_fillps'
  end subroutine
  subroutine solver_gaussel_z_gpu(n,ng,hi,a,b,c,bcz,c_or_f,norm,p)
#else
  subroutine solver_gaussel_z(n,ng,hi,a,b,c,bcz,c_or_f,norm,p)
#endif
    integer , intent(in), dimension(3) :: n,ng,hi
    real(rp), intent(in), target, dimension(:) :: a,b,c
    character(len=1), intent(in), dimension(0:1) :: bcz
    character(len=1), intent(in), dimension(3) :: c_or_f
    real(rp), intent(in) :: norm
    real(rp), intent(inout), target, contiguous, dimension(0:,0:,0:) :: p
    character(kind=c_char) :: b2(2)
    integer(c_int) :: istat
    b2 = [bcz(0),bcz(1)]
    !$acc host_data use_device(p,a,b,c)
    istat = cansb200_solve_z_bc(ctx,b2,c_or_f(3),c_loc(p),n,1,real(norm,c_double),c_loc(a),c_loc(b),c_loc(c),MEM_KIND, &
                                cuda_stream_of_queue_1())
    !$acc end host_data
    if(istat /= 0) error stop 'cansb200_solve_z_bc'
  end subroutine
#if defined(_OPENACC)
end module mod_solver_gpu
#else
end module mod_solver
#endif
!
#if defined(_OPENACC)
module mod_workspaces   ! replaces src/workspaces.f90 (public init_wspace_arrays,set_cufft_wspace,cudecomp_finalize, :14)
  use, intrinsic :: iso_c_binding
  use mod_types, only: rp
  use mod_cansb200
  use mod_common_cudecomp, only: work,solver_buf_0,solver_buf_1   ! the three scratch pencils rk.f90 aliases (src/rk.f90:27-29)
  use openacc
  implicit none
  private
  public init_wspace_arrays,set_cufft_wspace,cudecomp_finalize
contains
  subroutine init_wspace_arrays
    ! the reference allocates host arrays and maps them onto cuDecomp's device buffers with acc_map_data
    ! (src/workspaces.f90:58-63); here they are mapped onto the library's own scratch pencils
    type(c_ptr) :: dptr
    integer(c_size_t) :: nel
    if(cansb200_get_work(ctx,0,dptr,nel) /= 0) error stop 'cansb200_get_work(0)'
    allocate(work(nel));         call acc_map_data(work,dptr,nel*c_sizeof(1._rp))
    if(cansb200_get_work(ctx,1,dptr,nel) /= 0) error stop 'cansb200_get_work(1)'
    allocate(solver_buf_0(nel)); call acc_map_data(solver_buf_0,dptr,nel*c_sizeof(1._rp))
    if(cansb200_get_work(ctx,2,dptr,nel) /= 0) error stop 'cansb200_get_work(2)'
    allocate(solver_buf_1(nel)); call acc_map_data(solver_buf_1,dptr,nel*c_sizeof(1._rp))
  end subroutine
  subroutine set_cufft_wspace(arrplan,istream)   ! src/workspaces.f90:112-130: nothing to attach, the plans own their tables
    integer, intent(in), dimension(:) :: arrplan
    integer(acc_handle_kind), target, intent(in), optional :: istream
  end subroutine
  subroutine cudecomp_finalize                    ! src/workspaces.f90:132-150
    call acc_unmap_data(work); call acc_unmap_data(solver_buf_0); call acc_unmap_data(solver_buf_1)
    if(cansb200_finalize(ctx) /= 0) error stop 'cansb200_finalize'
    ctx = c_null_ptr
  end subroutine
end module mod_workspaces
#endif
