/*
 * edgpu.h -- C-ABI of the B200-native Lanczos hot path of dmft-ed (libedgpu.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch/C++ types.  Every entry point names the
 * reference interface it replaces (file:line relative to the reference root).  The Fortran ISO_C_BINDING shim
 * that binds these symbols at the ED_DIAG / ED_GF_NORMAL / ED_OBSERVABLES call sites is in
 * fortran/ed_gpu_binding.f90 and described in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; edgpu_last_error() gives the message (the
 *    reference's `stop "..."`, e.g. ED_HAMILTONIAN_DIRECT_HxV.f90:45,50).
 *  - "cplx" host arrays are interleaved (re,im) doubles = Fortran complex(8) = C double _Complex.
 *  - levels are 1-based like the reference (level l <-> bit l-1; up levels 1..Ns, down levels Ns+1..2Ns).
 *  - sector vectors live on the device behind edgpu_vec handles as REAL fp64 (H is real symmetric for
 *    ed_mode=normal, bath_type=normal, real Hloc); host import/export uses the reference ordering
 *    i = r_up + r_dw*DimUp (0-based; ED_SETUP.f90:905-916).
 *  - there is no CPU fallback: every compute entry point needs a CUDA device.
 *  - one host thread drives a context (the reference's global-state, single-thread model).
 */
#ifndef EDGPU_H
#define EDGPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct edgpu_ctx edgpu_ctx;
typedef struct edgpu_sector edgpu_sector;
typedef struct edgpu_vec edgpu_vec;

#define EDGPU_MAXORB 5

/* Problem shape (ED_INPUT_VARS.f90:121-196; only ed_mode=normal, bath_type=normal are on this path). */
typedef struct {
    int32_t norb, nbath, nspin;
    int32_t hfmode;          /* HFMODE */
    int32_t layout;          /* 0 = auto, 1 = reference (colex) order on device, 2 = star-product order */
    int32_t hxv_kernel;      /* 0 = auto, 1 = generic table kernel, 2 = star-product tile kernels (one Dimdw x ld tile),
                                3 = fiber kernels on the pair-tile layout (hxv_fiber.cu; auto picks them for sectors >= 2^21 states) */
    int32_t reserved[8];     /* reserved[0]: test hooks (bit 0: 2-column down strips, bit 1: single-stage up pass, bit 2: generic tile
                                pass only, bit 3: no copy-engine kernels, bit 4: copy-engine kernels for every block size,
                                bit 5: programmatic dependent launch of the copy-engine kernels, bit 6 (+ bit 7): at most
                                2 (3) stages in the up pipeline, bit 9: 2 instead of 3 x images in the up pipeline);
                                fiber kernels: bit 2 = thread-per-element pair kernels only, bit 4 = fiber kernels for every block size,
                                bit 10 / 11 = pass 1 / pass 2 by the thread-per-element kernels, bit 12 = small pipeline slots (two-slot images) */
} edgpu_params;

/* init_ed_structure + setup_pointers_normal (ED_MAIN.f90:73,91; ED_SETUP.f90:150-360,372-496).
 * device < 0: use the current CUDA device.  stream: a cudaStream_t (may be NULL = default stream). */
int edgpu_init(const edgpu_params *p, int device, void *stream, edgpu_ctx **ctx);
int edgpu_finalize(edgpu_ctx *ctx);
/* Hosts that drive several contexts from several threads (ed_solve's work units, ED_MAIN.f90:598-636 deals them over MPI
 * ranks): edgpu_params.reserved[2] bit 0 makes edgpu_init create a non-blocking stream owned by the context (the `stream`
 * argument is ignored); a thread calls edgpu_bind_thread once before it uses a context (cudaSetDevice). */
int edgpu_bind_thread(edgpu_ctx *ctx);
edgpu_ctx *edgpu_sector_context(const edgpu_sector *s);
const char *edgpu_last_error(const edgpu_ctx *ctx);
int edgpu_version(void);
int edgpu_ns(const edgpu_ctx *ctx);                         /* Ns = (Nbath+1)*Norb, ED_SETUP.f90:99-101 */

/* set_dmft_bath + set_Hloc (ED_MAIN.f90:260-267; dmft_aux.f90:494-511; ED_AUX_FUNX.f90:139-158).
 * bath: [e(ispin,iorb,k) ..., v(ispin,iorb,k) ...], k fastest; hloc_cplx: impHloc(Nspin,Nspin,Norb,Norb)
 * column-major, may be NULL (= 0).  Complex Hloc entries are rejected (error) on this path. */
int edgpu_set_hamiltonian(edgpu_ctx *ctx, const double *bath, int32_t bath_len, const double *hloc_cplx,
                          const double *uloc, double ust, double jh, double jx, double jp, double xmu);

/* build_sector / build_Hv_sector / delete_Hv_sector / vecDim_Hv_sector (ED_SETUP.f90:886-916;
 * ED_HAMILTONIAN.f90:42-149).  Device-resident per-spin tables; the full map is never needed by H*v. */
int edgpu_sector_build(edgpu_ctx *ctx, int32_t nup, int32_t ndw, edgpu_sector **s);
int edgpu_sector_free(edgpu_sector *s);
/* The same sector SHARDED over `nranks` processes (one per GPU): H is block diagonal over (down-block, up-block) pairs of
 * conserved star occupations, so whole pairs are dealt to the ranks (longest-processing-time-first over their sizes) and a
 * rank stores and multiplies only its own pairs -- H*v needs no exchange at all, only the Lanczos scalars are summed
 * (edgpu_shard_lanczos_tridiag).  Replaces the row-block decomposition + MPI_Allgatherv of directMatVec_MPI_cc
 * (ED_HAMILTONIAN_DIRECT_HxV.f90:97-195, ED_HAMILTONIAN.f90:56-62).  Vectors of such a sector hold the local pairs only;
 * upload / fill / download address them by the global reference index (download writes 0 for foreign elements). */
int edgpu_sector_build_shard(edgpu_ctx *ctx, int32_t nup, int32_t ndw, int32_t rank, int32_t nranks, edgpu_sector **s);
/* Communicator of the sharded path: the ONLY collective left is the sum of the Lanczos scalars (2 doubles per step), done
 * with NCCL over NVLink/NVSwitch (libnccl.so.2 is loaded at run time).  Rank 0 calls edgpu_comm_unique_id, the host
 * broadcasts the 128 bytes (MPI_Bcast in the reference's MPI world, ED_MAIN.f90:62-71 ed_set_MpiComm), every rank calls
 * edgpu_comm_init.  Afterwards edgpu_vec_dot, edgpu_lanczos_tridiag and edgpu_lanczos_gs on a sharded sector return the
 * GLOBAL results on every rank (same call sequence on all ranks, like the reference's MPI Lanczos). */
int edgpu_comm_unique_id(edgpu_ctx *ctx, unsigned char id[128]);
int edgpu_comm_init(edgpu_ctx *ctx, const unsigned char id[128], int32_t rank, int32_t nranks);
int edgpu_comm_finalize(edgpu_ctx *ctx);
int edgpu_comm_info(const edgpu_ctx *ctx, int32_t *rank, int32_t *nranks);
/* host[0..n) <- reduction over the ranks, in place (op 0: sum, 1: min, 2: max): what MPI_Allreduce does for the reference's
 * MPI build (e.g. the sums over the states of the distributed ed_solve, host/ed_main.cpp) */
int edgpu_comm_allreduce_host(edgpu_ctx *ctx, double *host, int64_t n, int32_t op);
/* layout_kind: 0 = one Dimdw x ld tile, 3 = pair tiles; nalloc: doubles stored per vector on this rank */
int edgpu_sector_info(const edgpu_sector *s, int32_t *layout_kind, int64_t *nalloc, int32_t *shard_rank, int32_t *shard_nranks);
int edgpu_sector_dim(const edgpu_sector *s, int64_t *dim, int64_t *dim_up, int64_t *dim_dw);
/* type(sector_map)%map (ED_VARS_GLOBAL.f90:28-31), 64-bit, entries [first, first+count): built by a device
 * kernel and copied to host_out. */
int edgpu_sector_map(const edgpu_sector *s, int64_t first, int64_t count, uint64_t *host_out);
/* order-sensitive checksum of the whole map computed on the device (for full-size sectors):
 * sum_i map[i]*(2i+1) mod 2^64, plus sortedness / popcount violations count. */
int edgpu_sector_map_check(const edgpu_sector *s, uint64_t *checksum, int64_t *violations);

/* vectors */
int edgpu_vec_alloc(edgpu_sector *s, edgpu_vec **v);
int edgpu_vec_free(edgpu_vec *v);
int edgpu_vec_upload(edgpu_vec *v, const double *host, int32_t is_cplx);    /* complex: real part is taken */
int edgpu_vec_download(const edgpu_vec *v, double *host, int32_t is_cplx); /* complex: imag = 0 */
/* reference rows [rd0, rd1) (down-spin colex ranks) of the vector: host[(rd - rd0)*DimUp + ru], real */
int edgpu_vec_download_rows(const edgpu_vec *v, int64_t rd0, int64_t rd1, double *host);
int edgpu_vec_fill_normal(edgpu_vec *v, uint64_t seed);   /* Philox4x32-10 N(0,1), element index = counter */
/* Philox uniforms in (-1,1), exact arithmetic (bit-identical to the oracle's generator): Lanczos start vector,
 * standing in for the random_number() start of sp_lanc_eigh (.repo/PLAIN_LANCZOS.f90:310-318) */
int edgpu_vec_fill_uniform(edgpu_vec *v, uint64_t seed);
int edgpu_vec_copy(edgpu_vec *dst, const edgpu_vec *src);
int edgpu_vec_dot(const edgpu_vec *a, const edgpu_vec *b, double *out);
int edgpu_vec_scale(edgpu_vec *a, double alpha);

/* spHtimesV_cc (ED_VARS_GLOBAL.f90:48-54,105) = directMatVec_cc (ED_HAMILTONIAN_DIRECT_HxV.f90:21-92) or
 * spMatVec_cc (ED_HAMILTONIAN_STORED_HxV.f90:132-143).  Host complex in/out: parity hook / fine seam. */
int edgpu_hxv(edgpu_sector *s, int64_t nloc, const double *v_cplx, double *hv_cplx);
/* device-resident real vectors: the product path */
int edgpu_hxv_dev(edgpu_sector *s, const edgpu_vec *x, edgpu_vec *y);
/* ed_sparse_H=T: ed_buildH_c (ED_HAMILTONIAN_STORED_HxV.f90:28-113) on the device; afterwards edgpu_hxv*
 * use the stored CSR.  edgpu_sector_csr_download mirrors the spH0 rows (insertion order, 0-based cols). */
int edgpu_sector_build_csr(edgpu_sector *s);
int edgpu_sector_drop_csr(edgpu_sector *s);
int edgpu_sector_csr_nnz(const edgpu_sector *s, int64_t *nnz);
int edgpu_sector_csr_download(const edgpu_sector *s, int64_t *rowptr, int64_t *cols, double *vals);
/* dense Hmat (build_Hv_sector(isector,Hmat), ED_HAMILTONIAN.f90:75-79) for small sectors: column-major
 * real dim x dim, computed as H applied to the identity on the device. */
int edgpu_sector_dense(edgpu_sector *s, double *hmat);

/* sp_lanc_eigh (ED_DIAG.f90:173-181; ancestor .repo/PLAIN_LANCZOS.f90:286-385): two-pass plain Lanczos.
 * v0: start vector handle (overwritten by the normalised ground state).  alanc/blanc: host arrays of
 * nitermax+1 (blanc[0] unused, Fortran blanc(1)); may be NULL. */
int edgpu_lanczos_gs(edgpu_sector *s, edgpu_vec *v0, int32_t nitermax, double threshold, int32_t ncheck,
                     double *e0, int32_t *nlanc, double *alanc, double *blanc);
/* sp_eigh = ARPACK 'SA' (ED_DIAG.f90:149-166): the `neigen` lowest eigenpairs from a basis of `ncv` vectors
 * (Nblock, ED_DIAG.f90:98) by thick-restart Lanczos with full re-orthogonalisation on the device (eigs.cu).
 * tol = lanc_tolerance (ARPACK's residual criterion); evals[neigen] ascending; vecs[neigen] receives new vector handles
 * (free with edgpu_vec_free); nconv = converged pairs, nmatvec = H*v applications. */
int edgpu_lanczos_eigs(edgpu_sector *s, int32_t neigen, int32_t ncv, int32_t maxrestart, double tol, uint64_t seed,
                       double *evals, edgpu_vec **vecs, int32_t *nconv, int32_t *nmatvec);
/* sp_lanc_tridiag (ED_GF_NORMAL.f90:187-192,240-245; ancestor .repo/PLAIN_LANCZOS.f90:154-180).
 * v is normalised and destroyed.  alfa[nlanc], beta[nlanc] (beta[0] unused) are zero-filled first. */
int edgpu_lanczos_tridiag(edgpu_sector *s, edgpu_vec *v, int32_t nlanc, double threshold,
                          double *alfa, double *beta, int32_t *nused);

/* GF seed loops (ED_GF_NORMAL.f90:159-174 cdg, :212-227 c): out(j) = sgn*in(m); returns <out|out> in
 * *norm2; normalise != 0 divides by sqrt(norm2) like :174.  isite is the 1-based level. */
int edgpu_apply_c(edgpu_sector *s_in, edgpu_sector *s_out, int32_t isite, int32_t dagger,
                  const edgpu_vec *in, edgpu_vec *out, int32_t normalise, double *norm2);

/* Seed of the spin-susceptibility chains (ED_GF_CHISPIN.f90:93-104, 198-208): out = S_z |in> in the SAME sector, with
 * S_z = 1/2 (n_up - n_dw) of impurity orbital iorb (1..Norb) or of all impurity orbitals (iorb = 0).
 * norm = sqrt(<out|out>) (before normalisation); normalise != 0 scales out to unit norm.  The chain itself is
 * edgpu_lanczos_tridiag on `out`. */
int edgpu_apply_sz(edgpu_sector *s, int32_t iorb, const edgpu_vec *in, edgpu_vec *out, int32_t normalise, double *norm);
/* Seed of the charge-susceptibility chains (ED_GF_CHIDENS.f90:126-136, 227-237): out = (n_up + n_dw) |in>, same conventions. */
int edgpu_apply_n(edgpu_sector *s, int32_t iorb, const edgpu_vec *in, edgpu_vec *out, int32_t normalise, double *norm);

/* observables_impurity core (ED_OBSERVABLES.f90:127-158): accumulates (+=) like the reference.
 * dens,dens_up,dens_dw,docc,magz: [Norb]; sz2,n2: [Norb*Norb] column-major; s2tot scalar. */
int edgpu_observables(edgpu_sector *s, const edgpu_vec *gs, double peso,
                      double *dens, double *dens_up, double *dens_dw, double *docc, double *magz,
                      double *sz2, double *n2, double *s2tot);

/* Sharded sector vector (multi-GPU, one process per GPU).  The vector is split by up-spin COLUMN
 * blocks: a rank owns columns [col0, col0+ncols) of every down-spin row as a [DimDw][ldc] tile (ldc multiple of 4).
 * The down-spin term never changes the column -> it is applied on the column shard; the up-spin term needs whole
 * rows -> after an all-to-all transpose a rank holds rows [row0,row0+nrows) of all columns as a [nrows][ld] tile.
 * These two entry points run the star-product kernels on caller-owned DEVICE pointers (e.g. torch tensors that
 * torch.distributed exchanges over NCCL); they replace directMatVec_MPI_cc (ED_HAMILTONIAN_DIRECT_HxV.f90:97-195),
 * whose full-vector MPI_Allgatherv (:163-166) is the thing this layout avoids.  Columns/rows are in the device
 * (internal) order of the sector; edgpu_shard_perm returns the reference<->internal permutations. */
int edgpu_shard_ld(const edgpu_sector *s, int64_t *ld_full);
int edgpu_shard_hxv_dw(edgpu_sector *s, int64_t ncols, int64_t ldc, const void *x_dev, void *y_dev);
int edgpu_shard_hxv_up(edgpu_sector *s, int64_t row0, int64_t nrows, const void *x_dev, void *y_dev, int32_t accumulate);
/* Same as edgpu_shard_hxv_up, but x and y are the all-to-all buffers themselves: nslab slabs packed back to back,
 * slab p = columns [col0[p], col0[p]+ldc[p]) of the nrows rows as [nrows][ldc[p]] (what rank p sent) -- saves the
 * unpack/pack passes around the up-spin term. */
int edgpu_shard_hxv_up_slabs(edgpu_sector *s, int64_t row0, int64_t nrows, int32_t nslab, const int64_t *col0,
                             const int64_t *ldc, const void *x_dev, void *y_dev, int32_t accumulate);
/* Peer mode of the sharded product (one process per GPU, NVLink/NVSwitch): the up-spin term of rows [row0,row0+nrows) reads
 * x from, and writes its result to, the column shards of ALL ranks directly (x_shards[p] / y_shards[p]: base pointers of rank
 * p's [DimDw][ldc[p]] shard, valid on this device -- own memory for p == this rank, CUDA IPC mappings otherwise).  The two
 * MPI exchanges of directMatVec_MPI_cc (ED_HAMILTONIAN_DIRECT_HxV.f90:163-166) are thereby fused into the kernel.
 * accumulate = 0: y = (D + H_up) x ; 1: y += ... .  The caller orders the ranks (a barrier before and after).
 * x_row0[p] (NULL = all 0): global index of the first row held by x_shards[p] -- 0 for a whole shard, row0 for a local
 * copy of just this rank's rows (prefetched by copy-engine DMA while the down pass runs, edgpu_copy_async). */
int edgpu_shard_hxv_up_peers(edgpu_sector *s, int64_t row0, int64_t nrows, int32_t nranks, const int64_t *col0,
                             const int64_t *ldc, const void *const *x_shards, const int64_t *x_row0, void *const *y_shards,
                             int32_t accumulate);
int edgpu_copy_async(edgpu_ctx *ctx, void *dst, const void *src, int64_t bytes, void *stream);
/* Device buffers for peer mode: separate cudaMalloc allocations (zero-filled) whose CUDA IPC handle (64 bytes) another
 * process on the node opens with edgpu_ipc_open (peer access is enabled lazily). */
int edgpu_dev_alloc(edgpu_ctx *ctx, int64_t bytes, void **dev_ptr);
int edgpu_dev_free(edgpu_ctx *ctx, void *dev_ptr);
int edgpu_ipc_export(edgpu_ctx *ctx, void *dev_ptr, unsigned char handle[64]);
int edgpu_ipc_open(edgpu_ctx *ctx, const unsigned char handle[64], void **dev_ptr);
int edgpu_ipc_close(edgpu_ctx *ctx, void *dev_ptr);
/* ref2int_up[DimUp], ref2int_dw[DimDw]: reference (colex) rank -> device index (host arrays, uint32) */
int edgpu_shard_perm(const edgpu_sector *s, uint32_t *ref2int_up, uint32_t *ref2int_dw);

/* measurement helpers (bench.py): average device time of `iters` H*v launches between CUDA events on the
 * context stream; flush_l2 != 0 writes a >L2 scratch buffer between launches (outside the events). */
int edgpu_bench_hxv(edgpu_sector *s, const edgpu_vec *x, edgpu_vec *y, int32_t iters, int32_t flush_l2,
                    double *ms_avg, int64_t *launches);
int edgpu_device_info(edgpu_ctx *ctx, int32_t *sm_count, int64_t *l2_bytes, int64_t *mem_bytes);
int edgpu_sync(edgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
