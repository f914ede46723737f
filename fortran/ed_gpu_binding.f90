!> ED_GPU_BINDING -- ISO_C_BINDING interface of libedgpu.so (include/edgpu.h) for dmft-ed.
!>
!> Thin layer: one `bind(C)` interface per C entry point plus the three wrapper routines that replace the bodies of
!> the hot-path call sites of the reference:
!>   gpu_lanc_eigh     <- sp_lanc_eigh / sp_eigh          ED_DIAG.f90:149-181
!>   gpu_lanc_gf_chain <- seed loop + sp_lanc_tridiag     ED_GF_NORMAL.f90:159-192 , 212-245
!>   gpu_observables   <- observables core                 ED_OBSERVABLES.f90:127-158
!> It cannot be compiled in the development image (no Fortran compiler, SURVEY F3); INTEGRATION.md shows where each
!> routine is called from.  All handles are opaque type(c_ptr); every C function returns 0 on success.
MODULE ED_GPU_BINDING
  USE, INTRINSIC :: ISO_C_BINDING
  USE ED_INPUT_VARS, only: Norb,Nbath,Nspin,hfmode,Uloc,Ust,Jh,Jx,Jp,xmu
  USE ED_VARS_GLOBAL, only: Ns,impHloc
  implicit none
  private

  type,bind(C) :: edgpu_params
     integer(c_int32_t) :: norb,nbath,nspin,hfmode,layout,hxv_kernel
     integer(c_int32_t) :: reserved(8)
  end type edgpu_params

  type(c_ptr),save :: gpu_ctx = c_null_ptr

  interface
     integer(c_int) function edgpu_init(p,device,stream,ctx) bind(C,name="edgpu_init")
       import :: edgpu_params,c_int,c_ptr
       type(edgpu_params),intent(in) :: p
       integer(c_int),value          :: device
       type(c_ptr),value             :: stream
       type(c_ptr),intent(out)       :: ctx
     end function edgpu_init
     integer(c_int) function edgpu_finalize(ctx) bind(C,name="edgpu_finalize")
       import :: c_int,c_ptr
       type(c_ptr),value :: ctx
     end function edgpu_finalize
     type(c_ptr) function edgpu_last_error(ctx) bind(C,name="edgpu_last_error")
       import :: c_ptr
       type(c_ptr),value :: ctx
     end function edgpu_last_error
     integer(c_int) function edgpu_set_hamiltonian(ctx,bath,bath_len,hloc,uloc,ust,jh,jx,jp,xmu) bind(C,name="edgpu_set_hamiltonian")
       import :: c_int,c_ptr,c_double,c_int32_t,c_double_complex
       type(c_ptr),value                 :: ctx
       real(c_double),intent(in)         :: bath(*)
       integer(c_int32_t),value          :: bath_len
       complex(c_double_complex),intent(in) :: hloc(*)
       real(c_double),intent(in)         :: uloc(*)
       real(c_double),value              :: ust,jh,jx,jp,xmu
     end function edgpu_set_hamiltonian
     integer(c_int) function edgpu_sector_build(ctx,nup,ndw,s) bind(C,name="edgpu_sector_build")
       import :: c_int,c_ptr,c_int32_t
       type(c_ptr),value        :: ctx
       integer(c_int32_t),value :: nup,ndw
       type(c_ptr),intent(out)  :: s
     end function edgpu_sector_build
     integer(c_int) function edgpu_sector_free(s) bind(C,name="edgpu_sector_free")
       import :: c_int,c_ptr
       type(c_ptr),value :: s
     end function edgpu_sector_free
     integer(c_int) function edgpu_sector_build_csr(s) bind(C,name="edgpu_sector_build_csr")
       import :: c_int,c_ptr
       type(c_ptr),value :: s
     end function edgpu_sector_build_csr
     integer(c_int) function edgpu_vec_alloc(s,v) bind(C,name="edgpu_vec_alloc")
       import :: c_int,c_ptr
       type(c_ptr),value       :: s
       type(c_ptr),intent(out) :: v
     end function edgpu_vec_alloc
     integer(c_int) function edgpu_vec_free(v) bind(C,name="edgpu_vec_free")
       import :: c_int,c_ptr
       type(c_ptr),value :: v
     end function edgpu_vec_free
     integer(c_int) function edgpu_vec_fill_uniform(v,seed) bind(C,name="edgpu_vec_fill_uniform")
       import :: c_int,c_ptr,c_int64_t
       type(c_ptr),value        :: v
       integer(c_int64_t),value :: seed
     end function edgpu_vec_fill_uniform
     integer(c_int) function edgpu_vec_download(v,host,is_cplx) bind(C,name="edgpu_vec_download")
       import :: c_int,c_ptr,c_int32_t,c_double_complex
       type(c_ptr),value                     :: v
       complex(c_double_complex),intent(out) :: host(*)
       integer(c_int32_t),value              :: is_cplx
     end function edgpu_vec_download
     integer(c_int) function edgpu_hxv(s,nloc,v,hv) bind(C,name="edgpu_hxv")
       import :: c_int,c_ptr,c_int64_t,c_double_complex
       type(c_ptr),value                     :: s
       integer(c_int64_t),value              :: nloc
       complex(c_double_complex),intent(in)  :: v(*)
       complex(c_double_complex),intent(out) :: hv(*)
     end function edgpu_hxv
     integer(c_int) function edgpu_lanczos_gs(s,v0,nitermax,threshold,ncheck,e0,nlanc,alanc,blanc) bind(C,name="edgpu_lanczos_gs")
       import :: c_int,c_ptr,c_int32_t,c_double
       type(c_ptr),value               :: s,v0
       integer(c_int32_t),value        :: nitermax,ncheck
       real(c_double),value            :: threshold
       real(c_double),intent(out)      :: e0
       integer(c_int32_t),intent(out)  :: nlanc
       type(c_ptr),value               :: alanc,blanc          ! c_null_ptr: not wanted
     end function edgpu_lanczos_gs
     integer(c_int) function edgpu_lanczos_eigs(s,neigen,ncv,maxrestart,tol,seed,evals,vecs,nconv,nmatvec) bind(C,name="edgpu_lanczos_eigs")
       import :: c_int,c_int32_t,c_int64_t,c_double,c_ptr
       type(c_ptr),value         :: s
       integer(c_int32_t),value  :: neigen,ncv,maxrestart
       real(c_double),value      :: tol
       integer(c_int64_t),value  :: seed
       real(c_double)            :: evals(*)
       type(c_ptr)               :: vecs(*)
       integer(c_int32_t)        :: nconv,nmatvec
     end function edgpu_lanczos_eigs
     integer(c_int) function edgpu_lanczos_tridiag(s,v,nlanc,threshold,alfa,beta,nused) bind(C,name="edgpu_lanczos_tridiag")
       import :: c_int,c_ptr,c_int32_t,c_double
       type(c_ptr),value              :: s,v
       integer(c_int32_t),value       :: nlanc
       real(c_double),value           :: threshold
       real(c_double),intent(out)     :: alfa(*),beta(*)
       integer(c_int32_t),intent(out) :: nused
     end function edgpu_lanczos_tridiag
     integer(c_int) function edgpu_apply_c(si,so,isite,dagger,vin,vout,normalise,norm2) bind(C,name="edgpu_apply_c")
       import :: c_int,c_ptr,c_int32_t,c_double
       type(c_ptr),value          :: si,so,vin,vout
       integer(c_int32_t),value   :: isite,dagger,normalise
       real(c_double),intent(out) :: norm2
     end function edgpu_apply_c
     integer(c_int) function edgpu_apply_sz(sec,iorb,vin,vout,normalise,norm) bind(C,name="edgpu_apply_sz")
       import :: c_ptr,c_int,c_int32_t,c_double
       type(c_ptr),value        :: sec,vin,vout
       integer(c_int32_t),value :: iorb,normalise
       real(c_double)           :: norm
     end function edgpu_apply_sz
     integer(c_int) function edgpu_apply_n(sec,iorb,vin,vout,normalise,norm) bind(C,name="edgpu_apply_n")
       import :: c_ptr,c_int,c_int32_t,c_double
       type(c_ptr),value        :: sec,vin,vout
       integer(c_int32_t),value :: iorb,normalise
       real(c_double)           :: norm
     end function edgpu_apply_n
     integer(c_int) function edgpu_sector_drop_csr(s) bind(C,name="edgpu_sector_drop_csr")
       import :: c_int,c_ptr
       type(c_ptr),value :: s
     end function edgpu_sector_drop_csr
     integer(c_int) function edgpu_sector_build_shard(ctx,nup,ndw,rank,nranks,s) bind(C,name="edgpu_sector_build_shard")
       import :: c_int,c_ptr,c_int32_t
       type(c_ptr),value        :: ctx
       integer(c_int32_t),value :: nup,ndw,rank,nranks
       type(c_ptr),intent(out)  :: s
     end function edgpu_sector_build_shard
     integer(c_int) function edgpu_comm_unique_id(ctx,id) bind(C,name="edgpu_comm_unique_id")
       import :: c_int,c_ptr,c_char
       type(c_ptr),value      :: ctx
       character(kind=c_char) :: id(128)
     end function edgpu_comm_unique_id
     integer(c_int) function edgpu_comm_init(ctx,id,rank,nranks) bind(C,name="edgpu_comm_init")
       import :: c_int,c_ptr,c_char,c_int32_t
       type(c_ptr),value        :: ctx
       character(kind=c_char)   :: id(128)
       integer(c_int32_t),value :: rank,nranks
     end function edgpu_comm_init
     integer(c_int) function edgpu_vec_upload(v,host,is_cplx) bind(C,name="edgpu_vec_upload")
       import :: c_int,c_ptr,c_int32_t,c_double
       type(c_ptr),value        :: v
       real(c_double)           :: host(*)
       integer(c_int32_t),value :: is_cplx
     end function edgpu_vec_upload
     integer(c_int) function edgpu_observables(s,gs,peso,dens,dens_up,dens_dw,docc,magz,sz2,n2,s2tot) bind(C,name="edgpu_observables")
       import :: c_int,c_ptr,c_double
       type(c_ptr),value            :: s,gs
       real(c_double),value         :: peso
       real(c_double),intent(inout) :: dens(*),dens_up(*),dens_dw(*),docc(*),magz(*),sz2(*),n2(*),s2tot
     end function edgpu_observables
  end interface

  public :: gpu_ctx
  public :: gpu_init, gpu_finalize, gpu_set_hamiltonian, gpu_check
  public :: gpu_lanc_eigs
  public :: gpu_lanc_eigh, gpu_lanc_gf_chain, gpu_lanc_spinchi_chain, gpu_lanc_denschi_chain, gpu_observables
  public :: gpu_comm_init, gpu_shard_lanc_tridiag
  public :: edgpu_sector_build, edgpu_sector_free, edgpu_vec_free, edgpu_hxv

contains

  !> `stop` with the library's message: the reference's error model (e.g. ED_HAMILTONIAN_DIRECT_HxV.f90:45,50)
  subroutine gpu_check(ierr,where)
    integer(c_int)   :: ierr
    character(len=*) :: where
    character(kind=c_char),pointer :: msg(:)
    integer :: i
    if(ierr==0)return
    call c_f_pointer(edgpu_last_error(gpu_ctx),msg,[1024])
    i=1
    do while(i<1024 .AND. msg(i)/=c_null_char)
       i=i+1
    enddo
    write(*,*)"ED GPU ERROR in "//where//": ",msg(1:i-1)
    stop
  end subroutine gpu_check

  !> called once from ed_init_solver after init_ed_structure (ED_MAIN.f90:73)
  subroutine gpu_init(device)
    integer,optional   :: device
    type(edgpu_params) :: p
    integer(c_int)     :: dev
    dev=-1;if(present(device))dev=device
    p%norb=Norb; p%nbath=Nbath; p%nspin=Nspin; p%hfmode=merge(1,0,hfmode)
    p%layout=0; p%hxv_kernel=0; p%reserved=0
    call gpu_check(edgpu_init(p,dev,c_null_ptr,gpu_ctx),"gpu_init")
  end subroutine gpu_init

  subroutine gpu_finalize()
    integer(c_int) :: ierr
    ierr=edgpu_finalize(gpu_ctx)
    gpu_ctx=c_null_ptr
  end subroutine gpu_finalize

  !> called from ed_solve after set_dmft_bath (ED_MAIN.f90:267): `bath` is the user bath vector
  subroutine gpu_set_hamiltonian(bath)
    real(8),dimension(:),intent(in) :: bath
    call gpu_check(edgpu_set_hamiltonian(gpu_ctx,bath,int(size(bath),c_int32_t),impHloc,Uloc,Ust,Jh,Jx,Jp,xmu),"gpu_set_hamiltonian")
  end subroutine gpu_set_hamiltonian

  !> replaces  call build_Hv_sector(isector) ; call sp_lanc_eigh(spHtimesV_cc,e,vec,Nitermax,threshold=lanc_tolerance) ;
  !>           call delete_Hv_sector()                                            (ED_DIAG.f90:134-186)
  !> The eigenvector stays on the device: (sec,vec) are handles to be stored in the state list instead of cvec.
  subroutine gpu_lanc_eigh(nup,ndw,sparse_H,Nitermax,tol,e0,sec,vec)
    integer,intent(in)      :: nup,ndw,Nitermax
    logical,intent(in)      :: sparse_H
    real(8),intent(in)      :: tol
    real(8),intent(out)     :: e0
    type(c_ptr),intent(out) :: sec,vec
    integer(c_int32_t)      :: nlanc
    call gpu_check(edgpu_sector_build(gpu_ctx,int(nup,c_int32_t),int(ndw,c_int32_t),sec),"build_Hv_sector")
    if(sparse_H)call gpu_check(edgpu_sector_build_csr(sec),"ed_buildH_c")
    call gpu_check(edgpu_vec_alloc(sec,vec),"gpu_lanc_eigh")
    call gpu_check(edgpu_vec_fill_uniform(vec,1234567_c_int64_t),"gpu_lanc_eigh")
    call gpu_check(edgpu_lanczos_gs(sec,vec,int(Nitermax,c_int32_t),tol,10_c_int32_t,e0,nlanc,c_null_ptr,c_null_ptr),"sp_lanc_eigh")
    ! delete_Hv_sector (ED_DIAG.f90:186): the stored matrix goes, the sector handle stays in the state list
    if(sparse_H)call gpu_check(edgpu_sector_drop_csr(sec),"delete_Hv_sector")
  end subroutine gpu_lanc_eigh

  !> replaces the sp_eigh (ARPACK) branch of ed_diag_c (ED_DIAG.f90:149-166): Neigen lowest pairs of sector (nup,ndw) from a
  !> basis of Nblock vectors.  eig_values(1:Neigen) ascending; vecs(1:Neigen) are device handles (the state list keeps them;
  !> edgpu_vec_download(vecs(i),cvec,1) gives the complex(8) copy in the reference order).
  subroutine gpu_lanc_eigs(nup,ndw,sparse_H,Neigen,Nblock,Nitermax,tol,eig_values,sec,vecs)
    integer,intent(in)      :: nup,ndw,Neigen,Nblock,Nitermax
    logical,intent(in)      :: sparse_H
    real(8),intent(in)      :: tol
    real(8),intent(out)     :: eig_values(Neigen)
    type(c_ptr),intent(out) :: sec,vecs(Neigen)
    integer(c_int32_t)      :: nconv,nmatvec
    call gpu_check(edgpu_sector_build(gpu_ctx,int(nup,c_int32_t),int(ndw,c_int32_t),sec),"build_Hv_sector")
    if(sparse_H)call gpu_check(edgpu_sector_build_csr(sec),"ed_buildH_c")
    call gpu_check(edgpu_lanczos_eigs(sec,int(Neigen,c_int32_t),int(Nblock,c_int32_t),int(max(1,Nitermax),c_int32_t),tol,&
         1234567_c_int64_t,eig_values,vecs,nconv,nmatvec),"sp_eigh")
    if(sparse_H)call gpu_check(edgpu_sector_drop_csr(sec),"delete_Hv_sector")
  end subroutine gpu_lanc_eigs

  !> replaces one half of lanc_build_gf_normal_c (ED_GF_NORMAL.f90:150-200 for cdg, :203-253 for c):
  !> seed = c^+_isite|gs> or c_isite|gs>, normalise, tridiagonalise in the target sector (jup,jdw).
  subroutine gpu_lanc_gf_chain(sec_i,vec_i,jup,jdw,isite,dagger,sparse_H,nlanc,norm2,alfa_,beta_)
    type(c_ptr),intent(in)  :: sec_i,vec_i
    integer,intent(in)      :: jup,jdw,isite,nlanc
    logical,intent(in)      :: dagger,sparse_H
    real(8),intent(out)     :: norm2,alfa_(nlanc),beta_(nlanc)
    type(c_ptr)             :: sec_j,vv
    integer(c_int32_t)      :: nused
    integer(c_int)          :: ierr
    integer(c_int64_t)      :: jdim,jdimup,jdimdw
    call gpu_check(edgpu_sector_build(gpu_ctx,int(jup,c_int32_t),int(jdw,c_int32_t),sec_j),"build_sector")
    call gpu_check(edgpu_vec_alloc(sec_j,vv),"gpu_lanc_gf_chain")
    call gpu_check(edgpu_apply_c(sec_i,sec_j,int(isite,c_int32_t),merge(1_c_int32_t,0_c_int32_t,dagger),vec_i,vv,1_c_int32_t,norm2),"apply_c")
    alfa_=0d0; beta_=0d0
    if(norm2>0d0)then              ! an annihilated seed contributes nothing: no 0/0 chain (host/ed_main.cpp build_gf)
       call gpu_check(edgpu_sector_dim(sec_j,jdim,jdimup,jdimdw),"vecDim_Hv_sector")
       if(sparse_H.AND.jdim>1)call gpu_check(edgpu_sector_build_csr(sec_j),"ed_buildH_c")
       call gpu_check(edgpu_lanczos_tridiag(sec_j,vv,int(nlanc,c_int32_t),1d-13,alfa_,beta_,nused),"sp_lanc_tridiag")
    endif
    ierr=edgpu_vec_free(vv)
    ierr=edgpu_sector_free(sec_j)    ! frees the CSR with the sector
  end subroutine gpu_lanc_gf_chain

  !> replaces the seed loop + sp_lanc_tridiag of lanc_ed_build_spinChi_c (ED_GF_CHISPIN.f90:89-122; iorb=0: the S_z^tot
  !> variant, :194-227).  norm is |S_z gs| -- the caller keeps the reference's norm2 convention (:101 vs :206).
  subroutine gpu_lanc_spinchi_chain(sec_i,vec_i,iorb,sparse_H,nlanc,norm,alfa_,beta_)
    type(c_ptr),intent(in)  :: sec_i,vec_i
    integer,intent(in)      :: iorb,nlanc
    logical,intent(in)      :: sparse_H
    real(8),intent(out)     :: norm,alfa_(nlanc),beta_(nlanc)
    type(c_ptr)             :: vv
    integer(c_int32_t)      :: nused
    integer(c_int)          :: ierr
    call gpu_check(edgpu_vec_alloc(sec_i,vv),"gpu_lanc_spinchi_chain")
    call gpu_check(edgpu_apply_sz(sec_i,int(iorb,c_int32_t),vec_i,vv,1_c_int32_t,norm),"apply_sz")
    alfa_=0d0; beta_=0d0
    if(norm>0d0)then
       if(sparse_H)call gpu_check(edgpu_sector_build_csr(sec_i),"ed_buildH_c")
       call gpu_check(edgpu_lanczos_tridiag(sec_i,vv,int(nlanc,c_int32_t),1d-13,alfa_,beta_,nused),"sp_lanc_tridiag")
       if(sparse_H)call gpu_check(edgpu_sector_drop_csr(sec_i),"delete_Hv_sector")
    endif
    ierr=edgpu_vec_free(vv)
  end subroutine gpu_lanc_spinchi_chain

  !> replaces the seed loop + sp_lanc_tridiag of lanc_ed_build_densChi_diag_c (ED_GF_CHIDENS.f90:122-152; iorb=0: the total
  !> charge variant lanc_ed_build_densChi_tot_c, :223-253).  Seed = (n_up+n_dw)|gs> of orbital iorb; norm = |seed|.
  subroutine gpu_lanc_denschi_chain(sec_i,vec_i,iorb,sparse_H,nlanc,norm,alfa_,beta_)
    type(c_ptr),intent(in)  :: sec_i,vec_i
    integer,intent(in)      :: iorb,nlanc
    logical,intent(in)      :: sparse_H
    real(8),intent(out)     :: norm,alfa_(nlanc),beta_(nlanc)
    type(c_ptr)             :: vv
    integer(c_int32_t)      :: nused
    integer(c_int)          :: ierr
    call gpu_check(edgpu_vec_alloc(sec_i,vv),"gpu_lanc_denschi_chain")
    call gpu_check(edgpu_apply_n(sec_i,int(iorb,c_int32_t),vec_i,vv,1_c_int32_t,norm),"apply_n")
    alfa_=0d0; beta_=0d0
    if(norm>0d0)then
       if(sparse_H)call gpu_check(edgpu_sector_build_csr(sec_i),"ed_buildH_c")
       call gpu_check(edgpu_lanczos_tridiag(sec_i,vv,int(nlanc,c_int32_t),1d-13,alfa_,beta_,nused),"sp_lanc_tridiag")
       if(sparse_H)call gpu_check(edgpu_sector_drop_csr(sec_i),"delete_Hv_sector")
    endif
    ierr=edgpu_vec_free(vv)
  end subroutine gpu_lanc_denschi_chain

  !> MPI build (ED_MAIN.f90:62-71 ed_set_MpiComm): one process per GPU.  Rank 0 creates the NCCL id, MPI broadcasts it.
  !> Call once after gpu_init.  `comm` is the MPI communicator of the solver (MpiComm).
  subroutine gpu_comm_init(comm,rank,nranks)
    integer,intent(in)     :: comm,rank,nranks
    character(kind=c_char) :: id(128)
    integer                :: ierr
    if(rank==0)call gpu_check(edgpu_comm_unique_id(gpu_ctx,id),"ncclGetUniqueId")
    call MPI_Bcast(id,128,MPI_CHARACTER,0,comm,ierr)
    call gpu_check(edgpu_comm_init(gpu_ctx,id,int(rank,c_int32_t),int(nranks,c_int32_t)),"ncclCommInitRank")
  end subroutine gpu_comm_init

  !> replaces  call sp_lanc_tridiag(MpiComm,spHtimesV_cc,vvloc,alfa_,beta_)  of the MPI build (ED_GF_NORMAL.f90:187-192 with
  !> spHtimesV_cc => directMatVec_MPI_cc, ED_HAMILTONIAN_DIRECT_HxV.f90:97-195): the sector is sharded by conserved
  !> occupation pairs, every rank passes the FULL seed in the reference order (it keeps its own pairs), no Allgatherv.
  subroutine gpu_shard_lanc_tridiag(jup,jdw,rank,nranks,vvinit,nlanc,alfa_,beta_)
    integer,intent(in)      :: jup,jdw,rank,nranks,nlanc
    complex(8),intent(in)   :: vvinit(:)
    real(8),intent(out)     :: alfa_(nlanc),beta_(nlanc)
    type(c_ptr)             :: sec_j,vv
    integer(c_int32_t)      :: nused
    integer(c_int)          :: ierr
    real(c_double),pointer  :: raw(:)
    call gpu_check(edgpu_sector_build_shard(gpu_ctx,int(jup,c_int32_t),int(jdw,c_int32_t),int(rank,c_int32_t),int(nranks,c_int32_t),sec_j),"build_sector")
    call gpu_check(edgpu_vec_alloc(sec_j,vv),"gpu_shard_lanc_tridiag")
    call c_f_pointer(c_loc(vvinit),raw,[2*size(vvinit)])
    call gpu_check(edgpu_vec_upload(vv,raw,1_c_int32_t),"vvinit")
    call gpu_check(edgpu_lanczos_tridiag(sec_j,vv,int(nlanc,c_int32_t),1d-13,alfa_,beta_,nused),"sp_lanc_tridiag")
    ierr=edgpu_vec_free(vv)
    ierr=edgpu_sector_free(sec_j)
  end subroutine gpu_shard_lanc_tridiag

  !> replaces the i-loop of observables_impurity (ED_OBSERVABLES.f90:127-158) for one state
  subroutine gpu_observables(sec,vec,peso,dens,dens_up,dens_dw,docc,magz,sz2,n2,s2tot)
    type(c_ptr),intent(in) :: sec,vec
    real(8),intent(in)     :: peso
    real(8),intent(inout)  :: dens(Norb),dens_up(Norb),dens_dw(Norb),docc(Norb),magz(Norb),sz2(Norb,Norb),n2(Norb,Norb),s2tot
    call gpu_check(edgpu_observables(sec,vec,peso,dens,dens_up,dens_dw,docc,magz,sz2,n2,s2tot),"observables_impurity")
  end subroutine gpu_observables

END MODULE ED_GPU_BINDING
