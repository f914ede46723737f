/*
 * ed_b200.h -- host-side mirror of the reference's solver interface for the Lanczos hot path.
 *
 * The reference is Fortran with module-global state (ED_VARS_GLOBAL.f90); no Fortran compiler exists in this
 * image, so the solver phases that drive the hot path (ED_MAIN.ed_init_solver / ed_solve, ED_DIAG.ed_diag_c,
 * ED_GF_NORMAL.build_gf_normal / build_sigma_normal, ED_OBSERVABLES.observables_impurity, ED_IO getters) are
 * mirrored in C++ (dmft-ed_b200/csrc/host/ed_main.cpp) on top of the C-ABI in edgpu.h.  Same names, argument
 * meaning and error behaviour; the module globals become an explicit `ed_solver` handle.  All functions return
 * 0 on success; ed_last_error() gives the reference's `stop` message otherwise.
 *
 * Array conventions are the reference's: complex arrays are interleaved (re,im) doubles in Fortran
 * column-major order, e.g. Smats(Nspin,Nspin,Norb,Norb,Lmats) (ED_IO/get_sigma_matsubara.f90:2-5).
 */
#ifndef ED_B200_H
#define ED_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ed_solver ed_solver;

/* ED_INPUT_VARS.f90:121-196 -- the flags that reach the hot path (defaults set by ed_input_defaults). */
typedef struct {
    int32_t Norb, Nbath, Nspin;
    double uloc[5];
    double ust, jh, jx, jp;
    double beta, xmu;
    int32_t hfmode;
    int32_t Lmats, Lreal;
    double wini, wfin, eps;
    double gs_threshold;
    double hwband;
    int32_t lanc_method;            /* 0 = "arpack" (default), 1 = "lanczos" */
    int32_t lanc_nstates_sector, lanc_nstates_total;
    int32_t lanc_niter, lanc_ngfiter;
    double lanc_tolerance;
    int32_t lanc_dim_threshold;
    int32_t ed_twin;                /* must be 0 on this path */
    int32_t ed_sparse_H;            /* 1: stored CSR H*v (ED_HAMILTONIAN_STORED_HxV), 0: direct on-the-fly H*v */
    int32_t ed_verbose;
    int32_t gpu_layout;             /* edgpu_params.layout  (0 auto) */
    int32_t gpu_hxv_kernel;         /* edgpu_params.hxv_kernel (0 auto) */
    int32_t chispin_flag;           /* CHISPIN_FLAG (ED_INPUT_VARS.f90:156): build the spin susceptibility in ed_solve */
    int32_t Ltau;                   /* LTAU (ED_INPUT_VARS.f90:148,211): imaginary-time points, raised to int(beta) */
    int32_t chidens_flag;           /* CHIDENS_FLAG (ED_INPUT_VARS.f90:157): charge susceptibility, diagonal + total channels */
    int32_t reserved[5];            /* [0], [1]: lanc_ncv_factor, lanc_ncv_add (0 = reference defaults 10, 0); [2]: host worker threads of
                                       ed_solve, each with a device context and stream of its own, over which the sectors of the scan
                                       and the Green's-function chains are dealt (0 = 4, 1 = serial; env ED_B200_WORKERS overrides) */
} ed_input;

void ed_input_defaults(ed_input *in);

/* get_bath_dimension (ED_BATH/user_aux.f90:11-30), bath_type=normal, ed_mode=normal: 2*Nspin*Norb*Nbath */
int32_t ed_get_bath_dimension(const ed_input *in);

/* ed_init_solver(bath,Hloc) (ED_MAIN.f90:61-101): allocates the solver, sets impHloc, fills `bath` with
 * init_dmft_bath (ED_BATH/dmft_aux.f90:105-127, noise 0).  hloc_cplx may be NULL (zero Hloc).
 * device < 0: current device; stream: cudaStream_t or NULL. */
int ed_init_solver(const ed_input *in, int device, void *stream, double *bath, int32_t bath_len,
                   const double *hloc_cplx, ed_solver **solver);
int ed_finalize_solver(ed_solver *solver);
/* MPI build of the reference (ED_MAIN.f90:62-71 ed_set_MpiComm + ED_MAIN.f90:598-636): one process per GPU.  Rank 0 obtains
 * the 128-byte id with ed_comm_unique_id, the host broadcasts it, every rank calls ed_set_comm.  ed_solve then deals the
 * sectors of the scan (ED_DIAG.f90:71-75) round-robin over the ranks, every rank builds the Green's-function /
 * susceptibility chains and observables of the ground states IT found (ED_GF_NORMAL.f90:150-253), and the results are
 * summed over the ranks, so that every getter returns the global answer on every rank (ed_get_state* list the local
 * states only). */
int ed_comm_unique_id(ed_solver *solver, unsigned char id[128]);
int ed_set_comm(ed_solver *solver, const unsigned char id[128], int32_t rank, int32_t nranks);
const char *ed_last_error(const ed_solver *solver);

/* ed_solve(bath[,Hloc]) (ED_MAIN.f90:253-282): set_dmft_bath, diagonalize_impurity, buildgf_impurity,
 * observables_impurity. */
int ed_solve(ed_solver *solver, const double *bath, int32_t bath_len, const double *hloc_cplx);

/* ED_IO getters (ED_IO/get_sigma_matsubara.f90:2-5, get_sigma_realaxis, get_gimp_matsubara.f90:2-5, ...):
 * out(Nspin,Nspin,Norb,Norb,L) complex, column-major. */
int ed_get_sigma_matsubara(const ed_solver *s, double *Smats);
int ed_get_sigma_real(const ed_solver *s, double *Sreal);
int ed_get_gimp_matsubara(const ed_solver *s, double *Gmats);
int ed_get_gimp_real(const ed_solver *s, double *Greal);
int ed_get_g0imp_matsubara(const ed_solver *s, double *G0mats);
int ed_get_g0imp_real(const ed_solver *s, double *G0real);
/* ED_IO/get_dens.f90:1-4, get_docc.f90:1-4, get_mag: real(8)(Norb) */
int ed_get_dens(const ed_solver *s, double *dens);
int ed_get_dens_up(const ed_solver *s, double *dens);
int ed_get_dens_dw(const ed_solver *s, double *dens);
int ed_get_docc(const ed_solver *s, double *docc);
int ed_get_mag(const ed_solver *s, double *magz);
/* sz2(Norb,Norb), n2(Norb,Norb) column-major, s2tot (ED_OBSERVABLES.f90:150-157) */
int ed_get_sz2_n2(const ed_solver *s, double *sz2, double *n2, double *s2tot);
/* grids: wm(Lmats), wr(Lreal) (ED_AUX_FUNX.f90:449-461) */
int ed_get_grids(const ed_solver *s, double *wm, double *wr);
/* Spin susceptibility <S_z,a(tau) S_z,a(0)> (build_chi_spin, ED_GF_CHISPIN.f90:22-40; printed by ED_IO/print_impChi.f90),
 * available after ed_solve when chispin_flag != 0.  Row Norb+1 is S_z^tot (only filled for Norb > 1, like the reference).
 * chi_iv (Norb+1, 0:Lmats) complex, chi_tau (Norb+1, 0:Ltau) real, chi_w (Norb+1, Lreal) complex, column-major;
 * vm(0:Lmats) bosonic Matsubara frequencies, tau(0:Ltau).  ltau returns the effective Ltau. */
int ed_get_spinchi(const ed_solver *s, double *chi_iv, double *chi_tau, double *chi_w, double *vm, double *tau, int32_t *ltau);
/* Charge susceptibility <n_a(tau) n_a(0)> (build_chi_dens, ED_GF_CHIDENS.f90:21-66), chidens_flag != 0: the channels whose
 * seeds are real -- densChi(a,a) (lanc_ed_build_densChi_diag_c :90-169) and, for Norb > 1, densChi_tot (:191-269).
 * The inter-orbital and spin-mixed channels (:291-673) use complex seeds (n_a + i n_b)|gs> and are NOT built: their
 * entries stay zero.  chi_iv (Norb,Norb,0:Lmats) complex, chi_tau (Norb,Norb,0:Ltau) real, chi_w (Norb,Norb,Lreal)
 * complex; tot_iv (0:Lmats) complex, tot_tau (0:Ltau) real, tot_w (Lreal) complex.  Grids as in ed_get_spinchi. */
int ed_get_denschi(const ed_solver *s, double *chi_iv, double *chi_tau, double *chi_w, double *tot_iv, double *tot_tau, double *tot_w);

/* state_list after diagonalize_impurity (ED_DIAG.f90:220-236, 383-416): number of kept states, zeta_function,
 * and per state: energy, nup, ndw. */
int ed_get_state_count(const ed_solver *s, int32_t *nstates, double *zeta, double *egs);
int ed_get_state(const ed_solver *s, int32_t istate, double *e, int32_t *nup, int32_t *ndw);
/* copy of eigenvector istate in the reference ordering, real(8)(dim) */
int ed_get_state_vector(const ed_solver *s, int32_t istate, double *vec, int64_t len);
/* lowest eigenvalue found in sector (nup,ndw) during the last ed_solve (eigenvalues_list.ed, ED_DIAG.f90:240) */
int ed_get_sector_energy(const ed_solver *s, int32_t nup, int32_t ndw, double *e);
/* number of plain-Lanczos steps sp_lanc_eigh took in that sector (0: LAPACK sector, ED_DIAG.f90:99-101) */
int ed_get_sector_nlanc(const ed_solver *s, int32_t nup, int32_t ndw, int32_t *nlanc);
/* GF Lanczos chains of the last ed_solve (ED_GF_NORMAL.f90:187-194): count, then per chain the meta data and
 * alfa/beta (length nlanc, beta[0] unused). */
int ed_get_chain_count(const ed_solver *s, int32_t *nchains);
int ed_get_chain(const ed_solver *s, int32_t ichain, int32_t *iorb, int32_t *ispin, int32_t *isign, int32_t *istate,
                 int32_t *nlanc, int32_t *nused, double *norm2, double *alfa, double *beta, int32_t cap);
/* restrict the sector scan (ED_SECTORS / sectors_mask, ED_DIAG.f90:71): list of (nup,ndw) pairs, n=0 clears */
int ed_set_sectors_mask(ed_solver *s, const int32_t *nup_ndw_pairs, int32_t n);
/* wall-clock seconds of the phases of the last ed_solve: [0] diag, [1] gf, [2] sigma, [3] observables */
int ed_get_timings(const ed_solver *s, double *t4);

/* host linear algebra used by the mirror (LAPACK eigh in the reference: ED_DIAG.f90:194, ED_GF_NORMAL.f90:618);
 * exported for CPU-side tests.  a: n x n column-major symmetric, overwritten by eigenvectors; w: eigenvalues. */
int ed_host_eigh(int32_t n, double *a, double *w);
/* eigenvalues only (a is not modified): the pre-pass of the LAPACK sectors, whose vectors are computed only when a state can be kept */
int ed_host_eigvals(int32_t n, const double *a, double *w);
int ed_host_eigh_tridiag(int32_t n, const double *diag, const double *sub /* sub[1..n-1] */, double *w, double *z);

#ifdef __cplusplus
}
#endif
#endif
