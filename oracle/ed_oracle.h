/*
 * ed_oracle.h -- CPU ORACLE for the dmft-ed Lanczos hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may call into this library.
 * The product (dmft-ed_b200/csrc) never links or loads it.
 *
 * PARITY UNPINNED: the reference (yaoyongxin/dmft-ed, Fortran 90 + SciFortran)
 * ships no tests, golden vectors or fixtures for this path and cannot be built in
 * this image (no Fortran compiler, SciFortran/DMFT_Tools not vendored).  This file
 * is a literal C restatement of the reference's rules; each function cites the
 * reference file:line it follows (paths relative to the reference root).  It is
 * pinned only by independent invariants (tests/test_oracle_*.py): Hermiticity,
 * dense diagonalisation, U=0 analytic limits, particle-hole symmetry, sum rules.
 */
#ifndef ED_ORACLE_H
#define ED_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAXORB 5

typedef struct {
    int Norb, Nbath, Nspin, Ns;          /* Ns = (Nbath+1)*Norb  (ED_SETUP.f90:99-101, bath_type=normal) */
    int hfmode;                          /* ED_INPUT_VARS.f90:160 */
    int jhflag;                          /* ED_SETUP.f90:289-290: Norb>1 and (Jx!=0 or Jp!=0) */
    double uloc[ORA_MAXORB];
    double ust, jh, jx, jp, xmu;
    double *e;                           /* e[(ispin*Norb+iorb)*Nbath+k]   (dmft_aux.f90:494-501) */
    double *v;                           /* v[(ispin*Norb+iorb)*Nbath+k]   (dmft_aux.f90:503-511) */
    double *hloc_re, *hloc_im;           /* impHloc(ispin,jspin,iorb,jorb), Fortran column-major */
} ora_model;

ora_model *ora_model_new(int Norb, int Nbath, int Nspin, int hfmode,
                         const double *uloc, double ust, double jh, double jx, double jp, double xmu,
                         const double *bath, const double *hloc_re, const double *hloc_im);
void ora_model_free(ora_model *m);

/* ED_BATH/dmft_aux.f90:105-127 with ed_bath_noise_thr = 0; writes the user bath vector [e..., v...] */
void ora_init_bath(int Norb, int Nbath, int Nspin, double hwband, double *bath);

/* ED_SETUP.f90:1283-1300 / 818-830 */
int64_t ora_binomial(int n1, int n2);
int64_t ora_sector_dim(int Ns, int nup, int ndw);

/* ED_SETUP.f90:899-916, evaluated in 64-bit. literal=1: the reference's Theta(4^Ns) scan. */
int64_t ora_build_sector(int Ns, int nup, int ndw, uint64_t *map, int literal);

/* ED_SETUP.f90:1080-1106; pos is 1-based; returns 0 on the reference's "stop" condition */
int ora_c(int pos, uint64_t in, uint64_t *out, double *sgn);
int ora_cdg(int pos, uint64_t in, uint64_t *out, double *sgn);
/* ED_SETUP.f90:1307-1324: 1-based position, 0 if absent */
int64_t ora_binary_search(const uint64_t *a, int64_t n, uint64_t value);

/* ED_HAMILTONIAN_DIRECT_HxV.f90:21-92 + ED_HAMILTONIAN/direct/<part>.f90, scatter form, states j in [j0,j1) (0-based).
 * vin/hv are interleaved complex(8).  hv is NOT cleared here (caller does Hv=0 as at :66). */
void ora_direct_hxv(const ora_model *m, const uint64_t *map, int64_t dim,
                    const double *vin, double *hv, int64_t j0, int64_t j1);
/* Same operator, gather form for rows [i0,i1) (the row-block decomposition of ED_HAMILTONIAN.f90:56-62 that
 * is race-free across workers); this is what ED_HAMILTONIAN/stored/<part>.f90 stores. hv rows are overwritten. */
void ora_gather_hxv(const ora_model *m, const uint64_t *map, int64_t dim,
                    const double *vin, double *hv, int64_t i0, int64_t i1);

/* ED_HAMILTONIAN_STORED_HxV.f90:28-113 + stored/<part>.f90 + ED_SPARSE_MATRIX.f90:249-280: row-wise storage in
 * the reference's insertion order with accumulation of repeated columns. Returns nnz; fills rowptr[dim+1],
 * cols (0-based), vals (interleaved complex).  Call with cols=vals=NULL to count only. */
int64_t ora_stored_build(const ora_model *m, const uint64_t *map, int64_t dim,
                         int64_t *rowptr, int64_t *cols, double *vals);
/* ED_HAMILTONIAN_STORED_HxV.f90:132-143 */
void ora_stored_hxv(int64_t dim, const int64_t *rowptr, const int64_t *cols, const double *vals,
                    const double *vin, double *hv);

/* .repo/PLAIN_LANCZOS.f90:427-565 (tql2): d[n] diagonal, e[n] sub-diagonal in e[1..n-1], z[n*n] column-major */
int ora_tql2(int n, double *d, double *e, double *z);

/* Lanczos on the direct operator (complex vectors, like the reference).
 * .repo/PLAIN_LANCZOS.f90:87-118 (iteration), :154-180 (tridiag). Returns number of iterations done. */
int ora_lanc_tridiag(const ora_model *m, const uint64_t *map, int64_t dim, double *vin,
                     int nitermax, double threshold, double *alanc, double *blanc);
/* .repo/PLAIN_LANCZOS.f90:286-385 with the Ritz-vector bookkeeping fixed.  vect in: start
 * vector (interleaved complex), out: normalised ground state.  Returns nlanc; *egs lowest Ritz value. */
int ora_lanc_gs(const ora_model *m, const uint64_t *map, int64_t dim, double *vect,
                int nitermax, double threshold, int ncheck, double *egs, double *alanc, double *blanc);

/* ED_GF_NORMAL.f90:159-174 / 212-227: vvinit(j) = sgn * gs(m), un-normalised; returns <vv|vv>.
 * isite 1-based level (iorb or iorb+Ns); dagger=1 for cdg. */
double ora_apply_op(int Ns, int isite, int dagger,
                    const uint64_t *mapI, int64_t idim, const uint64_t *mapJ, int64_t jdim,
                    const double *gs, double *vv);

/* ED_OBSERVABLES.f90:127-158: accumulates (+=) dens,dens_up,dens_dw,docc,magz [Norb]; sz2,n2 [Norb*Norb]; s2tot */
void ora_observables(int Ns, int Norb, const uint64_t *map, int64_t dim, const double *gs, double peso,
                     double *dens, double *dens_up, double *dens_dw, double *docc, double *magz,
                     double *sz2, double *n2, double *s2tot);

/* Window oracle for sectors whose map + vectors do not fit the host: rows of H v with vin = Philox uniforms
 * (counter = reference index), map entries and binary searches replaced by colex (un)rank closed forms.
 * ora_map_entry(i) = map(i+1) of build_sector (ED_SETUP.f90:899-916). */
uint64_t ora_map_entry(int Ns, int nup, int ndw, int64_t i);
void ora_window_hxv(const ora_model *m, int nup, int ndw, uint64_t seed, const int64_t *rows, int64_t n, double *out);

/* Counter-based N(0,1) generator shared with the CUDA library (Philox4x32-10 + Box-Muller). */
void ora_philox_normal(uint64_t seed, int64_t i0, int64_t n, double *out);
/* uniform in (-1,1); exact arithmetic, bit-identical on CPU and GPU (Lanczos start vectors) */
void ora_philox_uniform(uint64_t seed, int64_t i0, int64_t n, double *out);

#ifdef __cplusplus
}
#endif
#endif
