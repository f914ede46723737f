// api.cu -- the C-ABI of libedgpu (include/edgpu.h): context, Hamiltonian, sectors, vectors, H*v entry points.
#include "edgpu_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

uint64_t edgpu_binom(int n, int k);
int sector_map_kernel(edgpu_sector *s, int64_t first, int64_t count, uint64_t *d_out);
int sector_map_check_kernel(edgpu_sector *s, unsigned long long *d_sum, unsigned long long *d_viol);
int vec_convert(edgpu_sector *s, int mode, const double *src, double *dst);
int vec_convert_rows(edgpu_sector *s, int mode, int64_t rd0, int64_t rd1, const double *src, double *dst, cudaStream_t st, double *flag = nullptr);
int vec_fill_random(edgpu_sector *s, int uniform, uint64_t seed, double *dst);
int vec_scale(edgpu_ctx *ctx, double *v, double alpha, int64_t n);
int csr_build(edgpu_sector *s);
int hxv_star_launches(const edgpu_sector *s);
int hxv_star_dw(edgpu_sector *s, const double *x, double *y, int64_t ncols, int64_t ld);
int hxv_star_up(edgpu_sector *s, const double *x, double *y, int64_t row0, int64_t nrows, int64_t ld, int accumulate);
int hxv_star_up_peers(edgpu_sector *s, const double *const *xp, double *const *yp, int64_t row0, int64_t nrows, int nslab,
                      const int64_t *col0, const int64_t *ldc, int accumulate, const int64_t *x_row0);
int hxv_star_up_slabs(edgpu_sector *s, const double *x, double *y, int64_t row0, int64_t nrows, int nslab,
                      const int64_t *col0, const int64_t *ldc, int accumulate);
int csr_download(const edgpu_sector *s, int64_t *rowptr, int64_t *cols, double *vals);
int dense_rows(edgpu_sector *s, double *d_H);

static thread_local std::string g_null_err;

int edgpu_fail(edgpu_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_null_err = buf;
    return 1;
}

// ---- pooled device buffers ----------------------------------------------------------------------------------
int pool_alloc(edgpu_ctx *ctx, size_t bytes, void **p)
{
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    auto it = ctx->pool_free.lower_bound(bytes);
    if (it != ctx->pool_free.end() && it->first <= bytes + bytes / 4) {          // at most 25 % larger than asked
        *p = it->second;
        ctx->pool_held -= it->first;
        ctx->pool_free.erase(it);
        return 0;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {                                                      // give the cached buffers back and retry
        cudaGetLastError();
        pool_trim(ctx, 0);
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) return edgpu_fail(ctx, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    ctx->pool_size[*p] = bytes;
    return 0;
}

void pool_release(edgpu_ctx *ctx, void *p)
{
    if (!p) return;
    auto it = ctx->pool_size.find(p);
    if (it == ctx->pool_size.end()) { cudaFree(p); return; }
    ctx->pool_free.emplace(it->second, p);
    ctx->pool_held += it->second;
    // cached (free) buffers of ONE context: an eighth of the device memory -- several contexts share a device (the worker
    // threads of ed_solve), and a context can only trim its own cache when an allocation fails
    const size_t cap = ctx->mem_bytes > 0 ? (size_t)ctx->mem_bytes / 8 : ((size_t)16 << 30);
    if (ctx->pool_held > cap) pool_trim(ctx, cap / 2);
}

void pool_trim(edgpu_ctx *ctx, size_t keep_bytes)
{
    if (ctx->pool_free.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    while (!ctx->pool_free.empty() && ctx->pool_held > keep_bytes) {
        auto it = std::prev(ctx->pool_free.end());                               // largest first
        cudaFree(it->second);
        ctx->pool_size.erase(it->second);
        ctx->pool_held -= it->first;
        ctx->pool_free.erase(it);
    }
}

CsrMatrix::~CsrMatrix() { cudaFree(rowptr); cudaFree(cols); cudaFree(rowlen); cudaFree(vals); }

extern "C" const char *edgpu_last_error(const edgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_null_err.c_str(); }
extern "C" int edgpu_version(void) { return EDGPU_VERSION; }
extern "C" int edgpu_ns(const edgpu_ctx *ctx) { return ctx ? ctx->ham.ns : -1; }

extern "C" int edgpu_init(const edgpu_params *p, int device, void *stream, edgpu_ctx **out)
{
    if (!p || !out) return edgpu_fail(nullptr, "edgpu_init: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return edgpu_fail(nullptr, "edgpu_init: no CUDA device available (%s); this library has no CPU fallback",
                          e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (p->norb < 1 || p->norb > EDGPU_MAXORB || p->nbath < 1 || p->nspin < 1 || p->nspin > 2)
        return edgpu_fail(nullptr, "edgpu_init: bad shape norb=%d nbath=%d nspin=%d", p->norb, p->nbath, p->nspin);
    edgpu_ctx *ctx = new edgpu_ctx();
    ctx->par = *p;
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return edgpu_fail(nullptr, "edgpu_init: cudaSetDevice(%d) failed", device); }
    ctx->stream = (cudaStream_t)stream;
    if (p->reserved[2] & 1) {
        // worker context of a multi-threaded host (ed_solve): a stream of its own that does not synchronise with stream 0
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return edgpu_fail(nullptr, "edgpu_init: cudaStreamCreate failed"); }
        ctx->own_stream = true;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return edgpu_fail(nullptr, "edgpu_init: cudaGetDeviceProperties failed"); }
    if (prop.major < 10) { delete ctx; return edgpu_fail(nullptr, "edgpu_init: device sm_%d%d is not Blackwell (sm_100a build)", prop.major, prop.minor); }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = prop.l2CacheSize;
    ctx->mem_bytes = (int64_t)prop.totalGlobalMem;
    HamParams &h = ctx->ham;
    h.norb = p->norb; h.nbath = p->nbath; h.nspin = p->nspin; h.hfmode = p->hfmode;
    h.ns = (p->nbath + 1) * p->norb;                                    // ED_SETUP.f90:99-101
    if (h.ns > 24) { const int nsbad = h.ns; delete ctx; return edgpu_fail(nullptr, "edgpu_init: Ns=%d > 24 unsupported", nsbad); }
    h.e.assign((size_t)h.nspin * h.norb * h.nbath, 0.0);
    h.v.assign((size_t)h.nspin * h.norb * h.nbath, 0.0);
    h.hloc.assign((size_t)h.nspin * h.norb * h.norb, 0.0);
    const size_t nscal = kScalSlots;
    if (cudaMalloc(&ctx->d_partials, sizeof(double) * kRedBlocks * 4) != cudaSuccess ||
        cudaMalloc(&ctx->d_dotpart, sizeof(double) * 16384) != cudaSuccess ||
        cudaMalloc(&ctx->d_scal, sizeof(double) * nscal) != cudaSuccess ||
        cudaMalloc(&ctx->d_flag, sizeof(double)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_scal, sizeof(double) * 64) != cudaSuccess) {
        delete ctx;
        return edgpu_fail(nullptr, "edgpu_init: scratch allocation failed");
    }
    cudaMemset(ctx->d_scal, 0, sizeof(double) * nscal);
    *out = ctx;
    return 0;
}

extern "C" edgpu_ctx *edgpu_sector_context(const edgpu_sector *s) { return s ? s->ctx : nullptr; }

extern "C" int edgpu_bind_thread(edgpu_ctx *ctx)
{
    if (!ctx) return edgpu_fail(nullptr, "edgpu_bind_thread: null context");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return 0;
}

extern "C" int edgpu_finalize(edgpu_ctx *ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    edgpu_comm_finalize(ctx);
    pool_trim(ctx, 0);
    ctx->bases.clear();
    for (int b = 0; b < 2; b++) { cudaFree(ctx->d_stage[b]); if (ctx->copy_stream) { cudaEventDestroy(ctx->ev_copied[b]); cudaEventDestroy(ctx->ev_free[b]); } }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) { cudaStreamDestroy(ctx->aux_stream); cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join); cudaEventDestroy(ctx->ev_gdw); cudaEventDestroy(ctx->ev_fdw); }
    for (auto &c : ctx->arena) cudaFree(c.first);
    ctx->arena.clear();
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    cudaFree(ctx->d_partials); cudaFree(ctx->d_dotpart); cudaFree(ctx->d_scal); cudaFree(ctx->d_flag); cudaFreeHost(ctx->h_scal); cudaFree(ctx->d_flush); cudaFree(ctx->d_xtab);
    delete ctx;
    return 0;
}

extern "C" int edgpu_device_info(edgpu_ctx *ctx, int32_t *sm_count, int64_t *l2_bytes, int64_t *mem_bytes)
{
    if (!ctx) return 1;
    if (sm_count) *sm_count = ctx->sm_count;
    if (l2_bytes) *l2_bytes = ctx->l2_bytes;
    if (mem_bytes) *mem_bytes = ctx->mem_bytes;
    return 0;
}

extern "C" int edgpu_sync(edgpu_ctx *ctx)
{
    if (!ctx) return 1;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int edgpu_set_hamiltonian(edgpu_ctx *ctx, const double *bath, int32_t bath_len, const double *hloc_cplx,
                                     const double *uloc, double ust, double jh, double jx, double jp, double xmu)
{
    if (!ctx || !bath || !uloc) return ctx ? edgpu_fail(ctx, "edgpu_set_hamiltonian: null argument") : 1;
    HamParams &h = ctx->ham;
    const int nb = h.nspin * h.norb * h.nbath;
    // check_bath_dimension (ED_MAIN.f90:257-258): normal bath, normal mode = 2*Nspin*Norb*Nbath
    if (bath_len != 2 * nb) return edgpu_fail(ctx, "ED_SOLVE_SINGLE Error: wrong bath dimensions (%d, expected %d)", bath_len, 2 * nb);
    for (int i = 0; i < nb; i++) { h.e[i] = bath[i]; h.v[i] = bath[nb + i]; }    // dmft_aux.f90:494-511
    std::fill(h.hloc.begin(), h.hloc.end(), 0.0);
    if (hloc_cplx) {
        for (int is = 0; is < h.nspin; is++)
            for (int js = 0; js < h.nspin; js++)
                for (int a = 0; a < h.norb; a++)
                    for (int b = 0; b < h.norb; b++) {
                        const size_t idx = (size_t)is + h.nspin * ((size_t)js + h.nspin * ((size_t)a + h.norb * (size_t)b));
                        const double re = hloc_cplx[2 * idx], im = hloc_cplx[2 * idx + 1];
                        if (im != 0.0) return edgpu_fail(ctx, "edgpu_set_hamiltonian: complex impHloc is not supported on the GPU path");
                        if (is != js && re != 0.0) return edgpu_fail(ctx, "edgpu_set_hamiltonian: spin-mixing impHloc needs ed_mode=nonsu2 (unsupported)");
                        if (is == js) h.hloc[(size_t)(is * h.norb + a) * h.norb + b] = re;
                    }
    }
    for (int a = 0; a < EDGPU_MAXORB; a++) h.uloc[a] = a < h.norb ? uloc[a] : 0.0;
    h.ust = ust; h.jh = jh; h.jx = jx; h.jp = jp; h.xmu = xmu;
    h.jhflag = (h.norb > 1) && (jx != 0.0 || jp != 0.0);                     // ED_SETUP.f90:289-290
    h.version++;
    ctx->bases.clear();                                                        // tables depend on the parameters
    return upload_xtab(ctx);
}

// ---------------------------------------------------------------------------------------------------------
static int sector_build_impl(edgpu_ctx *ctx, int32_t nup, int32_t ndw, int rank, int nranks, edgpu_sector **out);
void pair_layout_forget(edgpu_sector *s, const double *x);

extern "C" int edgpu_sector_build(edgpu_ctx *ctx, int32_t nup, int32_t ndw, edgpu_sector **out)
{
    return sector_build_impl(ctx, nup, ndw, 0, 1, out);
}

extern "C" int edgpu_sector_build_shard(edgpu_ctx *ctx, int32_t nup, int32_t ndw, int32_t rank, int32_t nranks, edgpu_sector **out)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) return ctx ? edgpu_fail(ctx, "edgpu_sector_build_shard: bad rank %d of %d", rank, nranks) : 1;
    return sector_build_impl(ctx, nup, ndw, rank, nranks, out);
}

static int sector_build_impl(edgpu_ctx *ctx, int32_t nup, int32_t ndw, int rank, int nranks, edgpu_sector **out)
{
    if (!ctx || !out) return ctx ? edgpu_fail(ctx, "edgpu_sector_build: null argument") : 1;
    *out = nullptr;
    const HamParams &h = ctx->ham;
    if (h.version == 0) return edgpu_fail(ctx, "edgpu_sector_build: Hamiltonian not set");
    if (nup < 0 || nup > h.ns || ndw < 0 || ndw > h.ns) return edgpu_fail(ctx, "edgpu_sector_build: (nup,ndw)=(%d,%d) outside 0..%d", nup, ndw, h.ns);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    auto s = new edgpu_sector();
    s->ctx = ctx; s->nup = nup; s->ndw = ndw;
    if (int rc = build_spin_basis(ctx, 0, nup, s->up)) { delete s; return rc; }
    if (int rc = build_spin_basis(ctx, h.nspin - 1, ndw, s->dw)) { delete s; return rc; }    // DW uses index Nspin (HxVbath.f90:8-9)
    s->dim_up = s->up->dim; s->dim_dw = s->dw->dim;
    s->dim = s->dim_up * s->dim_dw;                                           // ED_SETUP.f90:818-830
    s->ld = s->dim_up;
    if (s->up->layout == 2) s->ld = (s->dim_up + 3) / 4 * 4;                  // 32-byte aligned rows for the tiled kernels
    s->nalloc = s->dim_dw * s->ld;
    // pair-tile layout + fiber kernels (hxv_fiber.cu): hxv_kernel = 3 forces them, auto takes them for large sectors;
    // a sharded sector (nranks > 1) exists only in that layout
    // reserved[1] bit 0: the caller will store H (ed_sparse_H=T -> CSR), which needs the single-tile layout
    // (auto on ONE GPU: Norb = 2 only -- the Norb = 3 blocks read their slot tables from global memory and half of their tiles
    // take both pipeline slots; measured on Ns=18: 45.6 ms against 37.9 ms of the round-1 tile kernels)
    const bool want_pairs = ctx->par.hxv_kernel == 3 || nranks > 1 ||
                            (ctx->par.hxv_kernel == 0 && s->dim >= (1ll << 21) && !(ctx->par.reserved[1] & 1) && h.norb == 2);
    if (want_pairs && pair_layout_supported(s)) {
        if (int rc = pair_layout_build(s, rank, nranks)) { delete s; return rc; }
    } else if (ctx->par.hxv_kernel == 3 || nranks > 1) {
        delete s;
        return edgpu_fail(ctx, "edgpu_sector_build: the pair-tile layout (fiber kernels, sharding) needs bath_type=normal with diagonal impHloc, Norb 2-3, Nbath 2-7, no Jx/Jp");
    }
    *out = s;
    return 0;
}

/* layout facts of a sector: kind 0 = one Dimdw x ld tile, 3 = pair tiles; elements stored per vector on this rank */
extern "C" int edgpu_sector_info(const edgpu_sector *s, int32_t *layout_kind, int64_t *nalloc, int32_t *shard_rank, int32_t *shard_nranks)
{
    if (!s) return 1;
    if (layout_kind) *layout_kind = s->pl ? 3 : 0;
    if (nalloc) *nalloc = s->nalloc;
    if (shard_rank) *shard_rank = s->shard_rank;
    if (shard_nranks) *shard_nranks = s->shard_nranks;
    return 0;
}

extern "C" int edgpu_sector_free(edgpu_sector *s)
{
    if (!s) return 0;
    cudaStreamSynchronize(s->ctx->stream);
    for (int i = 0; i < 3; i++) pool_release(s->ctx, s->work[i]);
    delete s;
    return 0;
}

extern "C" int edgpu_sector_dim(const edgpu_sector *s, int64_t *dim, int64_t *dim_up, int64_t *dim_dw)
{
    if (!s) return 1;
    if (dim) *dim = s->dim;
    if (dim_up) *dim_up = s->dim_up;
    if (dim_dw) *dim_dw = s->dim_dw;
    return 0;
}

extern "C" int edgpu_sector_map(const edgpu_sector *s, int64_t first, int64_t count, uint64_t *host_out)
{
    if (!s || !host_out) return 1;
    edgpu_ctx *ctx = s->ctx;
    if (first < 0 || count < 0 || first + count > s->dim) return edgpu_fail(ctx, "edgpu_sector_map: range outside the sector");
    const int64_t chunk = 1 << 24;
    uint64_t *d = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d, sizeof(uint64_t) * (size_t)(count < chunk ? (count > 0 ? count : 1) : chunk)));
    for (int64_t off = 0; off < count; off += chunk) {
        const int64_t c = (count - off) < chunk ? (count - off) : chunk;
        if (int rc = sector_map_kernel(const_cast<edgpu_sector *>(s), first + off, c, d)) { cudaFree(d); return rc; }
        CUDA_TRY(ctx, cudaMemcpyAsync(host_out + off, d, sizeof(uint64_t) * (size_t)c, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(d);
    return 0;
}

extern "C" int edgpu_sector_map_check(const edgpu_sector *s, uint64_t *checksum, int64_t *violations)
{
    if (!s) return 1;
    edgpu_ctx *ctx = s->ctx;
    unsigned long long *d = nullptr, h[2] = {0, 0};
    CUDA_TRY(ctx, cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    CUDA_TRY(ctx, cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (int rc = sector_map_check_kernel(const_cast<edgpu_sector *>(s), d, d + 1)) { cudaFree(d); return rc; }
    CUDA_TRY(ctx, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    if (checksum) *checksum = h[0];
    if (violations) *violations = (int64_t)h[1];
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int edgpu_vec_alloc(edgpu_sector *s, edgpu_vec **out)
{
    if (!s || !out) return 1;
    edgpu_ctx *ctx = s->ctx;
    auto v = new edgpu_vec();
    v->s = s;
    if (int rc = pool_alloc(ctx, sizeof(double) * (size_t)s->nalloc, (void **)&v->d)) { delete v; return rc; }
    CUDA_TRY(ctx, cudaMemsetAsync(v->d, 0, sizeof(double) * (size_t)s->nalloc, ctx->stream));
    *out = v;
    return 0;
}

extern "C" int edgpu_vec_free(edgpu_vec *v)
{
    if (!v) return 0;
    pool_release(v->s->ctx, v->d);              // stream-ordered reuse: no synchronisation needed
    delete v;
    return 0;
}

// Staging buffer in the reference order (real or interleaved complex), then a conversion kernel.
static int stage_alloc(edgpu_ctx *ctx, size_t bytes, double **p)
{
    return pool_alloc(ctx, bytes, (void **)p);
}

extern "C" int edgpu_vec_upload(edgpu_vec *v, const double *host, int32_t is_cplx)
{
    if (!v || !host) return 1;
    edgpu_sector *s = v->s;
    edgpu_ctx *ctx = s->ctx;
    const size_t bytes = sizeof(double) * (size_t)s->dim * (is_cplx ? 2 : 1);
    if (!is_cplx && !s->pl && s->ld == s->dim_up && !s->up->ref2int && !s->dw->ref2int) {     // reference layout: direct copy
        CUDA_TRY(ctx, cudaMemcpyAsync(v->d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    // chunks of reference rows through two persistent staging buffers: the host-to-device copy of chunk k+1 (copy stream)
    // overlaps the layout conversion of chunk k (compute stream)
    const size_t rowb = sizeof(double) * (size_t)s->dim_up * (is_cplx ? 2 : 1);
    const size_t kChunk = (size_t)64 << 20;
    const int64_t crow = (int64_t)std::max<size_t>(1, kChunk / rowb);
    const size_t cbytes = rowb * (size_t)crow;
    if (ctx->stage_bytes < cbytes) {
        for (int b = 0; b < 2; b++) { cudaFree(ctx->d_stage[b]); ctx->d_stage[b] = nullptr; }
        for (int b = 0; b < 2; b++) CUDA_TRY(ctx, cudaMalloc(&ctx->d_stage[b], cbytes));
        ctx->stage_bytes = cbytes;
    }
    if (!ctx->copy_stream) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_free[b], cudaEventDisableTiming));
        }
    }
    CUDA_TRY(ctx, cudaMemsetAsync(v->d, 0, sizeof(double) * (size_t)s->nalloc, ctx->stream));        // pad columns stay zero
    if (is_cplx) CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(double), ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_free[0], ctx->stream));                                    // order after earlier work
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[0], 0));
    int k = 0;
    for (int64_t r0 = 0; r0 < s->dim_dw; r0 += crow, k++) {
        const int b = k & 1;
        const int64_t r1 = std::min<int64_t>(s->dim_dw, r0 + crow);
        if (k >= 2) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[b], 0));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_stage[b], reinterpret_cast<const char *>(host) + rowb * (size_t)r0, rowb * (size_t)(r1 - r0),
                                      cudaMemcpyHostToDevice, ctx->copy_stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
        if (int rc = vec_convert_rows(s, is_cplx ? 1 : 0, r0, r1, ctx->d_stage[b], v->d, ctx->stream, is_cplx ? ctx->d_flag : nullptr)) return rc;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_free[b], ctx->stream));
    }
    double bad = 0.0;
    if (is_cplx) CUDA_TRY(ctx, cudaMemcpyAsync(&bad, ctx->d_flag, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    // the vectors of this path are real (ed_mode=normal, real Hloc: H is real symmetric); a complex(8) start vector whose
    // imaginary part is not zero would silently lose it
    if (bad != 0.0) return edgpu_fail(ctx, "edgpu_vec_upload: the complex vector has a non-zero imaginary part (the device vectors of this path are real; use edgpu_hxv for complex operands)");
    return 0;
}

extern "C" int edgpu_vec_download(const edgpu_vec *v, double *host, int32_t is_cplx)
{
    if (!v || !host) return 1;
    edgpu_sector *s = v->s;
    edgpu_ctx *ctx = s->ctx;
    const size_t bytes = sizeof(double) * (size_t)s->dim * (is_cplx ? 2 : 1);
    if (!is_cplx && !s->pl && s->ld == s->dim_up && !s->up->ref2int && !s->dw->ref2int) {
        CUDA_TRY(ctx, cudaMemcpyAsync(host, v->d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    double *stage = nullptr;
    if (int rc = stage_alloc(ctx, bytes, &stage)) return rc;
    int rc = vec_convert(s, is_cplx ? 3 : 2, v->d, stage);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(host, stage, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = edgpu_fail(ctx, "edgpu_vec_download: %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(ctx->stream);
    pool_release(ctx, stage);
    return rc;
}

/* reference rows [rd0, rd1) of the vector (reference order, real): host[(rd - rd0)*DimUp + ru] */
extern "C" int edgpu_vec_download_rows(const edgpu_vec *v, int64_t rd0, int64_t rd1, double *host)
{
    if (!v || !host) return 1;
    edgpu_sector *s = v->s;
    edgpu_ctx *ctx = s->ctx;
    if (rd0 < 0 || rd1 < rd0 || rd1 > s->dim_dw) return edgpu_fail(ctx, "edgpu_vec_download_rows: rows [%lld,%lld) outside 0..%lld", (long long)rd0, (long long)rd1, (long long)s->dim_dw);
    if (rd1 == rd0) return 0;
    const size_t bytes = sizeof(double) * (size_t)(rd1 - rd0) * (size_t)s->dim_up;
    double *stage = nullptr;
    if (int rc = stage_alloc(ctx, bytes, &stage)) return rc;
    int rc = vec_convert_rows(s, 2, rd0, rd1, v->d, stage, ctx->stream);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(host, stage, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = edgpu_fail(ctx, "edgpu_vec_download_rows: %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(ctx->stream);
    pool_release(ctx, stage);
    return rc;
}

extern "C" int edgpu_vec_fill_normal(edgpu_vec *v, uint64_t seed)
{
    if (!v) return 1;
    return vec_fill_random(v->s, 0, seed, v->d);
}

extern "C" int edgpu_vec_fill_uniform(edgpu_vec *v, uint64_t seed)
{
    if (!v) return 1;
    return vec_fill_random(v->s, 1, seed, v->d);
}

extern "C" int edgpu_vec_copy(edgpu_vec *dst, const edgpu_vec *src)
{
    if (!dst || !src || dst->s->nalloc != src->s->nalloc) return 1;
    CUDA_TRY(dst->s->ctx, cudaMemcpyAsync(dst->d, src->d, sizeof(double) * (size_t)dst->s->nalloc, cudaMemcpyDeviceToDevice, dst->s->ctx->stream));
    return 0;
}

extern "C" int edgpu_vec_dot(const edgpu_vec *a, const edgpu_vec *b, double *out)
{
    if (!a || !b || !out || a->s->nalloc != b->s->nalloc) return 1;
    edgpu_ctx *ctx = a->s->ctx;
    if (int rc = vec_dot(ctx, a->d, b->d, a->s->nalloc, ctx->d_scal)) return rc;
    if (a->s->shard_nranks > 1) {                                  // sharded sector: the global dot product on every rank
        if (comm_nranks(ctx) != a->s->shard_nranks) return edgpu_fail(ctx, "edgpu_vec_dot: sharded sector without a matching communicator (edgpu_comm_init)");
        if (int rc = comm_allreduce_sum(ctx, ctx->d_scal, 1)) return rc;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = ctx->h_scal[0];
    return 0;
}

extern "C" int edgpu_vec_scale(edgpu_vec *a, double alpha)
{
    if (!a) return 1;
    return vec_scale(a->s->ctx, a->d, alpha, a->s->nalloc);
}

// ---------------------------------------------------------------------------------------------------------
// Tiled star kernels: two launches per occupation block, worth it once the sector is large; small sectors (the
// many sectors of an ed_solve scan) are launch-bound and use the single-launch generic kernel.
static bool use_star(const edgpu_sector *s)
{
    const edgpu_ctx *ctx = s->ctx;
    if (s->pl) return false;
    if (s->up->layout != 2 || s->dw->layout != 2 || ctx->par.hxv_kernel == 1 || ctx->ham.jhflag) return false;
    return ctx->par.hxv_kernel == 2 || s->dim >= (1ll << 21);
}

bool hxv_uses_star(const edgpu_sector *s) { return !s->csr && use_star(s); }

int hxv_dispatch(edgpu_sector *s, const double *x, double *y)
{
    if (s->pl) return hxv_fiber(s, x, y, nullptr, nullptr);
    if (s->csr) return hxv_csr(s, x, y);
    if (use_star(s)) return hxv_star(s, x, y);
    return hxv_generic(s, x, y);
}

extern "C" int edgpu_hxv_dev(edgpu_sector *s, const edgpu_vec *x, edgpu_vec *y)
{
    if (!s || !x || !y || x->s != s || y->s != s) return s ? edgpu_fail(s->ctx, "edgpu_hxv_dev: vectors do not belong to the sector") : 1;
    if (x->d == y->d) return edgpu_fail(s->ctx, "edgpu_hxv_dev: in-place product is not allowed");
    return hxv_dispatch(s, x->d, y->d);
}

extern "C" int edgpu_hxv(edgpu_sector *s, int64_t nloc, const double *v_cplx, double *hv_cplx)
{
    if (!s || !v_cplx || !hv_cplx) return 1;
    edgpu_ctx *ctx = s->ctx;
    if (nloc != s->dim) return edgpu_fail(ctx, "directMatVec_cc ERROR: Nloc != dim(isector)");     // DIRECT_HxV.f90:50
    double *stage = nullptr, *xr, *xi, *yr;
    const size_t cb = sizeof(double) * 2 * (size_t)s->dim;
    if (int rc = stage_alloc(ctx, cb, &stage)) return rc;
    if (int rc = sector_work(s, 0, &xr)) return rc;
    if (int rc = sector_work(s, 1, &xi)) return rc;
    if (int rc = sector_work(s, 2, &yr)) return rc;
    int rc = 0;
    do {
        if (cudaMemcpyAsync(stage, v_cplx, cb, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { rc = edgpu_fail(ctx, "edgpu_hxv: H2D failed"); break; }
        if ((rc = vec_convert(s, 1, stage, xr))) break;
        if ((rc = vec_convert(s, 4, stage, xi))) break;
        if ((rc = hxv_dispatch(s, xr, yr))) break;                 // H is real: act on Re and Im separately
        if ((rc = vec_convert(s, 3, yr, stage))) break;
        if ((rc = hxv_dispatch(s, xi, yr))) break;
        if ((rc = vec_convert(s, 5, yr, stage))) break;
        if (cudaMemcpyAsync(hv_cplx, stage, cb, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { rc = edgpu_fail(ctx, "edgpu_hxv: D2H failed"); break; }
    } while (0);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    pool_release(ctx, stage);
    if (!rc && e != cudaSuccess) rc = edgpu_fail(ctx, "edgpu_hxv: %s", cudaGetErrorString(e));
    return rc;
}

extern "C" int edgpu_sector_dense(edgpu_sector *s, double *hmat)
{
    if (!s || !hmat) return 1;
    edgpu_ctx *ctx = s->ctx;
    if (s->dim > 4096) return edgpu_fail(ctx, "edgpu_sector_dense: dim=%lld too large", (long long)s->dim);
    const int64_t n = s->dim;
    {
        // one launch: every thread writes the (sparse) row of its reference state into the dense matrix
        double *d_H = nullptr;
        if (int rc0 = pool_alloc(ctx, sizeof(double) * (size_t)n * (size_t)n, (void **)&d_H)) return rc0;
        const int rc = dense_rows(s, d_H);
        if (rc == 0) {
            cudaError_t e = cudaMemcpy(hmat, d_H, sizeof(double) * (size_t)n * (size_t)n, cudaMemcpyDeviceToHost);
            pool_release(ctx, d_H);
            if (e != cudaSuccess) return edgpu_fail(ctx, "edgpu_sector_dense: %s", cudaGetErrorString(e));
            return 0;
        }
        pool_release(ctx, d_H);
        if (!ctx->ham.jhflag) return rc;          // a real error; with Jx/Jp fall through to H applied to unit vectors
    }
    double *x, *y, *yr;
    if (int rc = sector_work(s, 0, &x)) return rc;
    if (int rc = sector_work(s, 1, &y)) return rc;
    if (int rc = sector_work(s, 2, &yr)) return rc;
    std::vector<double> col(n, 0.0);
    for (int64_t j = 0; j < n; j++) {
        // unit vector e_j in the reference order -> internal layout
        std::fill(col.begin(), col.end(), 0.0);
        col[j] = 1.0;
        CUDA_TRY(ctx, cudaMemcpyAsync(yr, col.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(x, 0, sizeof(double) * (size_t)s->nalloc, ctx->stream));
        if (int rc = vec_convert(s, 0, yr, x)) return rc;
        if (int rc = hxv_dispatch(s, x, y)) return rc;
        if (int rc = vec_convert(s, 2, y, yr)) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(hmat + j * n, yr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void k_flush(double *p, int64_t n, double v)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

extern "C" int edgpu_bench_hxv(edgpu_sector *s, const edgpu_vec *x, edgpu_vec *y, int32_t iters, int32_t flush_l2,
                               double *ms_avg, int64_t *launches)
{
    if (!s || !x || !y || iters < 1) return 1;
    edgpu_ctx *ctx = s->ctx;
    if (flush_l2 && !ctx->d_flush) {
        ctx->flush_bytes = (size_t)(ctx->l2_bytes > 0 ? ctx->l2_bytes : (128ll << 20)) * 2;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_flush, ctx->flush_bytes));
    }
    cudaEvent_t e0, e1;
    CUDA_TRY(ctx, cudaEventCreate(&e0));
    CUDA_TRY(ctx, cudaEventCreate(&e1));
    double total = 0.0;
    int rc = 0;
    if (flush_l2) {
        for (int i = 0; i < iters && !rc; i++) {
            k_flush<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>((double *)ctx->d_flush, (int64_t)(ctx->flush_bytes / 8), (double)i);
            cudaEventRecord(e0, ctx->stream);
            rc = hxv_dispatch(s, x->d, y->d);
            cudaEventRecord(e1, ctx->stream);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            total += ms;
        }
    } else {
        cudaEventRecord(e0, ctx->stream);
        for (int i = 0; i < iters && !rc; i++) rc = hxv_dispatch(s, x->d, y->d);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        total = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaGetLastError());
    if (ms_avg) *ms_avg = total / iters;
    if (launches) {
        int per = 1;
        if (s->pl) per = hxv_fiber_launches(s);
        else if (s->csr) per = 1;
        else if (use_star(s)) per = hxv_star_launches(s);
        else per = 1 + (ctx->ham.jhflag ? 1 : 0);
        *launches = (int64_t)per * iters;
    }
    return 0;
}

// ---- sharded sector vector (multi-GPU): star kernels on caller-owned device pointers ----------------------------
extern "C" int edgpu_shard_ld(const edgpu_sector *s, int64_t *ld_full)
{
    if (!s || !ld_full) return 1;
    *ld_full = s->ld;
    return 0;
}

extern "C" int edgpu_shard_hxv_dw(edgpu_sector *s, int64_t ncols, int64_t ldc, const void *x_dev, void *y_dev)
{
    if (!s || !x_dev || !y_dev) return 1;
    if (s->pl || s->up->layout != 2 || s->ctx->ham.jhflag) return edgpu_fail(s->ctx, "edgpu_shard_hxv_dw: needs the star-product layout (no Jx/Jp, no inter-orbital Hloc)");
    if (ncols < 0 || ldc < ncols || (ldc & 3)) return edgpu_fail(s->ctx, "edgpu_shard_hxv_dw: bad column shard (ncols=%lld, ldc=%lld)", (long long)ncols, (long long)ldc);
    return hxv_star_dw(s, (const double *)x_dev, (double *)y_dev, ncols, ldc);
}

extern "C" int edgpu_shard_hxv_up(edgpu_sector *s, int64_t row0, int64_t nrows, const void *x_dev, void *y_dev, int32_t accumulate)
{
    if (!s || !x_dev || !y_dev) return 1;
    if (s->pl || s->up->layout != 2 || s->ctx->ham.jhflag) return edgpu_fail(s->ctx, "edgpu_shard_hxv_up: needs the star-product layout (no Jx/Jp, no inter-orbital Hloc)");
    if (row0 < 0 || nrows < 0 || row0 + nrows > s->dim_dw) return edgpu_fail(s->ctx, "edgpu_shard_hxv_up: bad row shard");
    if (nrows == 0) return 0;
    return hxv_star_up(s, (const double *)x_dev, (double *)y_dev, row0, nrows, s->ld, accumulate);
}

extern "C" int edgpu_shard_hxv_up_slabs(edgpu_sector *s, int64_t row0, int64_t nrows, int32_t nslab, const int64_t *col0,
                                        const int64_t *ldc, const void *x_dev, void *y_dev, int32_t accumulate)
{
    if (!s || !x_dev || !y_dev || !col0 || !ldc) return 1;
    if (s->pl || s->up->layout != 2 || s->ctx->ham.jhflag) return edgpu_fail(s->ctx, "edgpu_shard_hxv_up_slabs: needs the star-product layout");
    if (row0 < 0 || nrows < 0 || row0 + nrows > s->dim_dw) return edgpu_fail(s->ctx, "edgpu_shard_hxv_up_slabs: bad row shard");
    if (nrows == 0) return 0;
    return hxv_star_up_slabs(s, (const double *)x_dev, (double *)y_dev, row0, nrows, nslab, col0, ldc, accumulate);
}

extern "C" int edgpu_shard_hxv_up_peers(edgpu_sector *s, int64_t row0, int64_t nrows, int32_t nranks, const int64_t *col0,
                                        const int64_t *ldc, const void *const *x_shards, const int64_t *x_row0, void *const *y_shards,
                                        int32_t accumulate)
{
    if (!s || !x_shards || !y_shards || !col0 || !ldc) return 1;
    if (s->pl || s->up->layout != 2 || s->ctx->ham.jhflag) return edgpu_fail(s->ctx, "edgpu_shard_hxv_up_peers: needs the star-product layout");
    if (row0 < 0 || nrows < 0 || row0 + nrows > s->dim_dw) return edgpu_fail(s->ctx, "edgpu_shard_hxv_up_peers: bad row shard");
    if (nrows == 0) return 0;
    return hxv_star_up_peers(s, (const double *const *)x_shards, (double *const *)y_shards, row0, nrows, nranks, col0, ldc, accumulate, x_row0);
}

// ---- device buffers that other processes of the node can map (CUDA IPC) -----------------------------------------
extern "C" int edgpu_dev_alloc(edgpu_ctx *ctx, int64_t bytes, void **dev_ptr)
{
    if (!ctx || !dev_ptr || bytes <= 0) return 1;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMalloc(dev_ptr, (size_t)bytes));             // a separate allocation: IPC handles map whole allocations
    CUDA_TRY(ctx, cudaMemsetAsync(*dev_ptr, 0, (size_t)bytes, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int edgpu_dev_free(edgpu_ctx *ctx, void *dev_ptr)
{
    if (!ctx) return 1;
    CUDA_TRY(ctx, cudaFree(dev_ptr));
    return 0;
}

extern "C" int edgpu_ipc_export(edgpu_ctx *ctx, void *dev_ptr, unsigned char handle[64])
{
    if (!ctx || !dev_ptr || !handle) return 1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, 64);
    return 0;
}

extern "C" int edgpu_ipc_open(edgpu_ctx *ctx, const unsigned char handle[64], void **dev_ptr)
{
    if (!ctx || !dev_ptr || !handle) return 1;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CUDA_TRY(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

// dst <- src (any pair of device pointers valid in this process, e.g. a peer mapping): copy-engine DMA on `stream`
extern "C" int edgpu_copy_async(edgpu_ctx *ctx, void *dst, const void *src, int64_t bytes, void *stream)
{
    if (!ctx || !dst || !src || bytes < 0) return 1;
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return 0;
}

extern "C" int edgpu_ipc_close(edgpu_ctx *ctx, void *dev_ptr)
{
    if (!ctx) return 1;
    CUDA_TRY(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

extern "C" int edgpu_shard_perm(const edgpu_sector *s, uint32_t *r2i_up, uint32_t *r2i_dw)
{
    if (!s) return 1;
    edgpu_ctx *ctx = s->ctx;
    if (r2i_up) {
        if (s->up->ref2int) CUDA_TRY(ctx, cudaMemcpy(r2i_up, s->up->ref2int, sizeof(uint32_t) * (size_t)s->dim_up, cudaMemcpyDeviceToHost));
        else for (int64_t i = 0; i < s->dim_up; i++) r2i_up[i] = (uint32_t)i;
    }
    if (r2i_dw) {
        if (s->dw->ref2int) CUDA_TRY(ctx, cudaMemcpy(r2i_dw, s->dw->ref2int, sizeof(uint32_t) * (size_t)s->dim_dw, cudaMemcpyDeviceToHost));
        else for (int64_t i = 0; i < s->dim_dw; i++) r2i_dw[i] = (uint32_t)i;
    }
    return 0;
}

// ---- CSR entry points (csr.cu) ----------------------------------------------------------------------------
extern "C" int edgpu_sector_build_csr(edgpu_sector *s)
{
    if (!s) return 1;
    if (s->pl) return edgpu_fail(s->ctx, "edgpu_sector_build_csr: the stored (CSR) path needs the single-tile layout (hxv_kernel 0-2 on a sector below 2^21 states, or hxv_kernel = 1/2)");
    return csr_build(s);
}
extern "C" int edgpu_sector_drop_csr(edgpu_sector *s) { if (!s) return 1; cudaStreamSynchronize(s->ctx->stream); s->csr.reset(); return 0; }
extern "C" int edgpu_sector_csr_nnz(const edgpu_sector *s, int64_t *nnz)
{
    if (!s || !s->csr || !nnz) return s ? edgpu_fail(s->ctx, "edgpu_sector_csr_nnz: CSR not built") : 1;
    *nnz = s->csr->nnz_true;
    return 0;
}
extern "C" int edgpu_sector_csr_download(const edgpu_sector *s, int64_t *rowptr, int64_t *cols, double *vals)
{
    if (!s || !s->csr) return s ? edgpu_fail(s->ctx, "edgpu_sector_csr_download: CSR not built") : 1;
    return csr_download(s, rowptr, cols, vals);
}
