// edgpu_internal.h -- internal structures of libedgpu (B200-native dmft-ed Lanczos hot path).
// Not part of the C-ABI (include/edgpu.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../include/edgpu.h"

#define EDGPU_VERSION 211

// ---- error plumbing: every C-ABI call returns int; message kept in the context (the reference `stop`s) ----
struct edgpu_ctx;
int edgpu_fail(edgpu_ctx *ctx, const char *fmt, ...);
#define CUDA_TRY(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return edgpu_fail((ctx), "%s:%d CUDA error %s: %s", __FILE__, __LINE__,                 \
                              cudaGetErrorName(e__), cudaGetErrorString(e__));                      \
    } while (0)

// Hamiltonian parameters on the host (set_dmft_bath + set_Hloc, ED_MAIN.f90:260-267).
struct HamParams {
    int norb = 0, nbath = 0, nspin = 0, ns = 0, hfmode = 1;
    std::vector<double> e, v;      // [(ispin*norb+iorb)*nbath+k]
    std::vector<double> hloc;      // real part of impHloc, [ispin][iorb][jorb] for (ispin,ispin) blocks
    double uloc[EDGPU_MAXORB] = {0, 0, 0, 0, 0};
    double ust = 0, jh = 0, jx = 0, jp = 0, xmu = 0;
    bool jhflag = false;           // ED_SETUP.f90:289-290
    uint64_t version = 0;
    double E(int ispin, int iorb, int k) const { return e[(size_t)(ispin * norb + iorb) * nbath + k]; }
    double V(int ispin, int iorb, int k) const { return v[(size_t)(ispin * norb + iorb) * nbath + k]; }
    double H(int ispin, int a, int b) const { return hloc[(size_t)(ispin * norb + a) * norb + b]; }
};

// One same-spin hop pair {p,q} (bit positions, p<q) with real amplitude index.
struct HopPair {
    int p, q;
    int amp;          // index into amp table
    int star;         // orbital whose star this hop belongs to (imp a <-> bath (a,k)), -1 for inter-orbital Hloc
    int k;            // bath index within the star (0-based), -1 otherwise
};

// Star-product description of one spin basis (layout 2). A "star" = impurity orbital a + its Nbath bath levels.
struct StarInfo;
// Per-spin tables of the fiber kernels (hxv_fiber.cu) and the pair-tile layout of a sector ("layout 3").
struct FibSpin;
struct PairLayout;

// Address of element (internal down index id, internal up index iu) of a sector vector.
//   mode 0: one Dimdw x ld tile, id*ld + iu                                   (layouts 1 and 2)
//   mode 3: pair tiles -- the vector is the concatenation of the OWNED (down-block, up-block) pairs; pair (bi,bj) is an
//           R x C matrix in 4x4 micro-tiles: ((rp/4)*C4 + cp/4)*16 + (rp%4)*4 + cp%4, with rp / cp the padded
//           position of the configuration inside its block (hxv_fiber.cu).  Returns -1 for a pair this rank does not own.
struct VAddr {
    int mode = 0;
    int nbu = 0;
    int64_t ld = 0;
    const int2 *rowinfo = nullptr;      // [dim_dw] (down-block, rp)
    const int2 *colinfo = nullptr;      // [dim_up] (up-block, cp)
    const int64_t *pbase = nullptr;     // [nbd*nbu] first element of the pair, -1 = not owned
    const int *c4 = nullptr;            // [nbu] micro-tile columns of the up-block
#ifdef __CUDACC__
    __device__ __forceinline__ int64_t operator()(int64_t id, int64_t iu) const
    {
        if (mode == 0) return id * ld + iu;
        const int2 r = rowinfo[id], c = colinfo[iu];
        const int64_t b = pbase[(int64_t)r.x * nbu + c.x];
        if (b < 0) return -1;
        return b + ((int64_t)(r.y >> 2) * c4[c.x] + (c.y >> 2)) * 16 + (r.y & 3) * 4 + (c.y & 3);
    }
#endif
};

// Per-spin, per-particle-number tables shared by all sectors that use them.
struct SpinBasis {
    int ns = 0, n = 0, pspin = 0;      // pspin: parameter spin index (0 or nspin-1)
    int layout = 1;                    // 1 = reference colex order, 2 = star-product order
    int64_t dim = 0;
    uint64_t ham_version = 0;
    uint32_t *cfg = nullptr;           // [dim]   device: internal index -> Ns-bit word
    uint32_t *cfg_ref = nullptr;       // [dim]   device: colex rank -> word (== cfg when layout 1)
    uint32_t *rank = nullptr;          // [2^ns]  device: word -> internal index (0xFFFFFFFF if popcount != n)
    uint32_t *ref2int = nullptr;       // [dim]   device: colex rank -> internal index (nullptr when layout 1)
    double *ediag = nullptr;           // [dim]   device: per-spin diagonal energy
    uint32_t *hop = nullptr;           // [maxhop][dim] device ELL: (target << 8) | code
    uint8_t *nhop = nullptr;           // [dim]
    int maxhop = 0;
    double *amp = nullptr;             // [256] device: signed amplitudes, code = 2*idx + (negative)
    std::shared_ptr<StarInfo> star;    // non-null when layout 2
    std::shared_ptr<FibSpin> fib;      // fiber-kernel tables (built on demand from `star`)
    ~SpinBasis();
};

struct CsrMatrix {
    int64_t dim = 0;
    int64_t nnz = 0;                   // stored entries including the 4-entry row padding
    int64_t nnz_true = 0;              // entries the reference would hold
    int64_t *rowptr = nullptr;         // [dim+1] device, multiples of 4
    uint32_t *cols = nullptr;          // [nnz] device, internal index (pads: own row)
    double *vals = nullptr;            // [nnz] device (pads: 0)
    uint8_t *rowlen = nullptr;         // [dim] device, true row lengths
    ~CsrMatrix();
};

struct EdComm;                         // comm.cu: NCCL communicator of the sharded Lanczos path
static constexpr int kLancMaxSteps = 4096;                         // longest Lanczos chain (d_scal holds the coefficients)
static constexpr int kScalSlots = 8 + 2 * (kLancMaxSteps + 1);     // device scalars: 8 working slots + alanc + blanc

struct edgpu_ctx {
    edgpu_params par;
    HamParams ham;
    int device = 0;
    cudaStream_t stream = 0;
    int sm_count = 148;
    int64_t l2_bytes = 0, mem_bytes = 0;
    std::string err;
    std::map<std::pair<int, int>, std::shared_ptr<SpinBasis>> bases;   // (pspin, n) -> tables
    // reduction scratch
    double *d_partials = nullptr;      // [kRedBlocks * 4]
    double *d_dotpart = nullptr;       // [16384] per-CTA partials of the dot product fused into the star up pass
    double *d_scal = nullptr;          // device scalars
    double *h_scal = nullptr;          // pinned host mirror
    void *d_flush = nullptr;           // L2 flush scratch
    size_t flush_bytes = 0;
    double *d_xtab = nullptr;          // [32*32] cross-spin interaction table X(u_imp,d_imp) + constant
    // chunked host<->device vector transfers (edgpu_vec_upload): two staging buffers, copy stream, hand-over events
    double *d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    // arena of the stored Lanczos basis (edgpu_lanczos_gs): chunks that live as long as the context, carved anew by every
    // call -- one allocation per ~16 vectors of the largest sector instead of one per Lanczos vector and sector
    std::vector<std::pair<void *, size_t>> arena;
    size_t arena_bytes = 0;
    // side stream of hxv_fiber: the thread-per-element pair kernels (disjoint pairs) fill the tails of the fiber kernels
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_gdw = nullptr, ev_fdw = nullptr;
    double *d_flag = nullptr;               // edgpu_vec_upload: set by the conversion kernel when an imaginary part is not zero
    bool own_stream = false;                // edgpu_params.reserved[2] bit 0: the context created its (non-blocking) stream
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    EdComm *comm = nullptr;            // set by edgpu_comm_init: reductions of sharded sectors are summed over the ranks
    // stream-ordered pool of the large device buffers (sector vectors, Lanczos workspaces, staging): an ed_solve scan builds
    // and frees hundreds of sectors, and cudaMalloc / cudaFree of ~100 MB buffers (10-50 ms each, device-synchronising)
    // dominated its wall time.  Every user of a pooled buffer works on ctx->stream, so reuse is ordered by the stream.
    std::multimap<size_t, void *> pool_free;
    std::map<void *, size_t> pool_size;
    size_t pool_held = 0;
};
int pool_alloc(edgpu_ctx *ctx, size_t bytes, void **p);
void pool_release(edgpu_ctx *ctx, void *p);
void pool_trim(edgpu_ctx *ctx, size_t keep_bytes);

struct edgpu_sector {
    edgpu_ctx *ctx = nullptr;
    int nup = 0, ndw = 0;
    std::shared_ptr<SpinBasis> up, dw;
    int64_t dim_up = 0, dim_dw = 0, dim = 0;
    int64_t ld = 0;                    // leading dimension (elements) of the Dimdw x Dimup tile, >= dim_up
    int64_t nalloc = 0;                // elements of a vector of this sector: dim_dw * ld, or the owned pair tiles
    std::shared_ptr<PairLayout> pl;    // non-null: vectors live in the pair-tile layout (hxv_fiber.cu)
    int shard_rank = 0, shard_nranks = 1;   // pair tiles dealt over `shard_nranks` processes (this one owns `shard_rank`)
    std::unique_ptr<CsrMatrix> csr;
    // scratch vectors owned by the sector (Lanczos workspace), allocated lazily
    double *work[3] = {nullptr, nullptr, nullptr};
};

struct edgpu_vec {
    edgpu_sector *s = nullptr;
    double *d = nullptr;               // [s->nalloc], pads (ld > dim_up) are kept at zero
};

VAddr sector_vaddr(const edgpu_sector *s);
int pair_layout_build(edgpu_sector *s, int rank, int nranks);       // decides fiber support, fills s->pl / s->nalloc
bool pair_layout_supported(const edgpu_sector *s);
int hxv_fiber(edgpu_sector *s, const double *x, double *y, double *dot, int *ndot);
int hxv_fiber_launches(const edgpu_sector *s);

// ---- internal entry points (implemented across the .cu files) ----
int build_spin_basis(edgpu_ctx *ctx, int pspin, int n, std::shared_ptr<SpinBasis> &out);
int hxv_generic(edgpu_sector *s, const double *x, double *y);
int hxv_star(edgpu_sector *s, const double *x, double *y);
int hxv_star_dot(edgpu_sector *s, const double *x, double *y, double *dot, int *ndot);
bool hxv_uses_star(const edgpu_sector *s);
int hxv_csr(edgpu_sector *s, const double *x, double *y);
int hxv_dispatch(edgpu_sector *s, const double *x, double *y);
int upload_xtab(edgpu_ctx *ctx);

int comm_nranks(const edgpu_ctx *ctx);
int comm_allreduce_sum(edgpu_ctx *ctx, double *d_buf, int n);

// vector kernels (lanczos.cu)
constexpr int kRedBlocks = 1184;       // 148 SMs x 8
constexpr int kRedThreads = 256;
int vec_dot(edgpu_ctx *ctx, const double *a, const double *b, int64_t n, double *d_out);
int sector_work(edgpu_sector *s, int i, double **p);

// layout conversion (ops.cu): reference order host <-> internal device order
int vec_import_ref(edgpu_sector *s, const double *d_ref, double *d_int);   // both device pointers
int vec_export_ref(edgpu_sector *s, const double *d_int, double *d_ref);
