// lanczos.cu -- device Lanczos: fused vector kernels, deterministic reductions, ground-state and
// tridiagonalisation drivers.
//
// Replaces SciFortran sp_lanc_eigh / sp_lanc_tridiag as used at ED_DIAG.f90:173-181 and
// ED_GF_NORMAL.f90:187-192,240-245 (in-tree ancestor .repo/PLAIN_LANCZOS.f90:87-118,154-180,286-385).
// Recurrence (no re-orthogonalisation, like the reference):
//   tmp = H vin - b vout ; a = <vin,tmp> ; tmp -= a vin ; b = |tmp| ; vout = vin ; vin = tmp/b
// evaluated on UNNORMALISED vectors w_k = b_k v_k so that no pass is spent on scaling (lanczos_step below):
//   u = H w_k ; a_k = <w_k,u>/b_k^2 ; w_{k+1} = u/b_k - (b_k/b_{k-1}) w_{k-1} - (a_k/b_k) w_k ; b_{k+1} = |w_{k+1}|
// (<v_k, v_{k-1}> = 0 to rounding, so a_k equals the reference's <v_k, H v_k - b_k v_{k-1}> to ~1e-16), and with
// <w_k,u> accumulated inside the H*v kernels of the star path: 4 vector passes per step instead of 9.
// All Lanczos scalars live in device memory so that the GF tridiagonalisation runs without host round trips;
// dot products use a fixed launch shape + fixed-order second stage => bit-reproducible run to run.
#include "edgpu_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void block_store_partial(double v, double *out)
{
    __shared__ double sh[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
        v = warp_sum(v);
        if (lane == 0) *out = v;
    }
}

// Second stage folded into the producing kernel (saves a launch per reduction -- the many small sectors of an ed_solve
// scan are launch-bound): the block that takes the last ticket sums the partials in index order, so the result does not
// depend on which block that is.  mode 0: sum ; 1: sqrt(sum) ; 2: sum / (*nrm)^2.  Call with all threads of the block,
// after block_store_partial.
__device__ __forceinline__ void finalize_by_last_block(const double *__restrict__ partials, int nblocks, int mode,
                                                       const double *__restrict__ nrm, double *__restrict__ out,
                                                       unsigned int *__restrict__ ticket)
{
    __shared__ bool s_last;
    __shared__ double sh2[32];
    __threadfence();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == (unsigned)nblocks - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double v = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += __ldcg(partials + i);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh2[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (blockDim.x >> 5)) ? sh2[lane] : 0.0;
        v = warp_sum(v);
        if (lane == 0) {
            if (mode == 1) v = sqrt(v);
            else if (mode == 2) { const double d = *nrm; v = v / (d * d); }
            *out = v;
            *ticket = 0u;
        }
    }
}

// second stage: one block, fixed order.  mode 0: out = sum ; mode 1: out = sqrt(sum)
__global__ void k_reduce_final(const double *__restrict__ partials, int n, double *__restrict__ out, int mode)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
    __shared__ double sh[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
        v = warp_sum(v);
        if (lane == 0) *out = mode ? sqrt(v) : v;
    }
}

__global__ void __launch_bounds__(kRedThreads) k_dot(const double *__restrict__ a, const double *__restrict__ b,
                                                     int64_t n, double *__restrict__ partials)
{
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += a[i] * b[i];
    block_store_partial(acc, partials + blockIdx.x);
}

// v = v / (*den)      (vin = tmp/b, also the initial normalisation)
__global__ void __launch_bounds__(kRedThreads) k_div(double *__restrict__ v, const double *__restrict__ den, int64_t n)
{
    const double d = *den;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = v[i] / d;
}

__global__ void __launch_bounds__(kRedThreads) k_scale(double *__restrict__ v, double alpha, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v[i] *= alpha;
}

// acc += z * v
__global__ void __launch_bounds__(kRedThreads) k_axpy(double *__restrict__ acc, const double *__restrict__ v, double z, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc[i] += z * v[i];
}

static inline int red_blocks(int64_t n)
{
    int64_t b = (n + kRedThreads - 1) / kRedThreads;
    return (int)(b < kRedBlocks ? (b > 0 ? b : 1) : kRedBlocks);
}

int vec_dot(edgpu_ctx *ctx, const double *a, const double *b, int64_t n, double *d_out)
{
    int nb = red_blocks(n);
    k_dot<<<nb, kRedThreads, 0, ctx->stream>>>(a, b, n, ctx->d_partials);
    k_reduce_final<<<1, 256, 0, ctx->stream>>>(ctx->d_partials, nb, d_out, 0);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

int vec_scale(edgpu_ctx *ctx, double *v, double alpha, int64_t n)
{
    k_scale<<<red_blocks(n), kRedThreads, 0, ctx->stream>>>(v, alpha, n);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

int vec_axpy(edgpu_ctx *ctx, double *acc, const double *v, double z, int64_t n)
{
    k_axpy<<<red_blocks(n), kRedThreads, 0, ctx->stream>>>(acc, v, z, n);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

int sector_work(edgpu_sector *s, int i, double **p)
{
    if (!s->work[i]) {
        if (int rc = pool_alloc(s->ctx, sizeof(double) * (size_t)s->nalloc, (void **)&s->work[i])) return rc;
        CUDA_TRY(s->ctx, cudaMemsetAsync(s->work[i], 0, sizeof(double) * (size_t)s->nalloc, s->ctx->stream));
    }
    *p = s->work[i];
    return 0;
}

// Device scalar slots (ctx->d_scal): [0] norm / b , [1] a , [2] zero constant, [8 + k] alanc[k], [8 + NMAX + k] blanc[k]
static constexpr int kScalB = 0, kScalArr = 8, kLancMax = kLancMaxSteps;

// a = (sum of the partials) / nrm^2
__global__ void k_lanc_alpha(const double *__restrict__ partials, int n, const double *__restrict__ nrm, double *__restrict__ out)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
    __shared__ double sh[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
        v = warp_sum(v);
        if (lane == 0) { const double d = *nrm; *out = v / (d * d); }
    }
}

// a = <a_vec, b_vec> / (*nrm)^2 in one launch
__global__ void __launch_bounds__(kRedThreads) k_dot_alpha(const double *__restrict__ a, const double *__restrict__ b, int64_t n,
                                                           double *__restrict__ partials, const double *__restrict__ nrm,
                                                           double *__restrict__ out, unsigned int *__restrict__ ticket)
{
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += a[i] * b[i];
    block_store_partial(acc, partials + blockIdx.x);
    finalize_by_last_block(partials, gridDim.x, 2, nrm, out, ticket);
}

// w_new = u/nc - (bp/no) w_old - (a/nc) w_cur, written to `out` (== old: in place) ; b = sqrt(sum w_new^2), one launch
__global__ void __launch_bounds__(kRedThreads) k_lanc_c_norm(const double *old, double *out, const double *__restrict__ u,
                                                             const double *__restrict__ cur, const double *__restrict__ p_bprev,
                                                             const double *__restrict__ p_ncur, const double *__restrict__ p_nold,
                                                             const double *__restrict__ p_a, int64_t n, double *__restrict__ partials,
                                                             double *__restrict__ b_out, unsigned int *__restrict__ ticket, int mode)
{
    const double nc = *p_ncur, c_old = *p_bprev / *p_nold, c_cur = *p_a / nc;
    double acc = 0.0;
    // 16-byte accesses, two independent ones in flight per thread (the buffers are 256-byte aligned; the scalar loop takes an odd tail)
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * blockDim.x;
    const double2 *old2 = reinterpret_cast<const double2 *>(old);
    double2 *out2 = reinterpret_cast<double2 *>(out);
    const double2 *__restrict__ u2 = reinterpret_cast<const double2 *>(u), *__restrict__ cur2 = reinterpret_cast<const double2 *>(cur);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n2; i += 2 * stride) {
        const double2 ua = u2[i], ub = u2[i + stride], oa = old2[i], ob = old2[i + stride], ca = cur2[i], cb = cur2[i + stride];
        double2 ta, tb;
        ta.x = ua.x / nc - c_old * oa.x - c_cur * ca.x; ta.y = ua.y / nc - c_old * oa.y - c_cur * ca.y;
        tb.x = ub.x / nc - c_old * ob.x - c_cur * cb.x; tb.y = ub.y / nc - c_old * ob.y - c_cur * cb.y;
        out2[i] = ta; out2[i + stride] = tb;
        acc += ta.x * ta.x; acc += ta.y * ta.y; acc += tb.x * tb.x; acc += tb.y * tb.y;
    }
    for (; i < n2; i += stride) {
        const double2 ua = u2[i], oa = old2[i], ca = cur2[i];
        double2 ta;
        ta.x = ua.x / nc - c_old * oa.x - c_cur * ca.x; ta.y = ua.y / nc - c_old * oa.y - c_cur * ca.y;
        out2[i] = ta;
        acc += ta.x * ta.x; acc += ta.y * ta.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double t = u[n - 1] / nc - c_old * old[n - 1] - c_cur * cur[n - 1];
        out[n - 1] = t;
        acc += t * t;
    }
    block_store_partial(acc, partials + blockIdx.x);
    finalize_by_last_block(partials, gridDim.x, mode, nullptr, b_out, ticket);
}

__global__ void k_set_one(double *p) { *p = 1.0; }

// out = sum_k c[k] * V_k (k ascending, the order of the axpy sequence it replaces): the Ritz vector from a stored Lanczos basis
__global__ void __launch_bounds__(kRedThreads) k_ritz(const double *const *__restrict__ V, const double *__restrict__ c, int nvec,
                                                      double *__restrict__ out, int64_t n)
{
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 acc = make_double2(0.0, 0.0);
        for (int k = 0; k < nvec; k++) {
            const double2 v = reinterpret_cast<const double2 *>(V[k])[i];
            const double ck = c[k];
            acc.x += ck * v.x; acc.y += ck * v.y;
        }
        reinterpret_cast<double2 *>(out)[i] = acc;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double acc = 0.0;
        for (int k = 0; k < nvec; k++) acc += c[k] * V[k][n - 1];
        out[n - 1] = acc;
    }
}

// last step of a reduction that was summed over the ranks (sharded sectors): mode 1: sqrt ; 2: / (*nrm)^2
__global__ void k_fin(double *p, int mode, const double *__restrict__ nrm)
{
    if (mode == 1) *p = sqrt(*p);
    else if (mode == 2) { const double d = *nrm; *p = *p / (d * d); }
}

static inline bool is_sharded(const edgpu_sector *s) { return s->shard_nranks > 1; }
static int check_comm(edgpu_sector *s)
{
    if (is_sharded(s) && comm_nranks(s->ctx) != s->shard_nranks)
        return edgpu_fail(s->ctx, "sharded sector (%d ranks) but the communicator has %d ranks: call edgpu_comm_init first", s->shard_nranks, comm_nranks(s->ctx));
    return 0;
}

// Device scalar slots: [0] norm, [3] the constant 1, [kScalArr + k] alanc[k], [kScalArr + kLancMax + k] blanc[k]
static int reset_scalars(edgpu_ctx *ctx)
{
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scal, 0, sizeof(double) * kScalSlots, ctx->stream));
    k_set_one<<<1, 1, 0, ctx->stream>>>(ctx->d_scal + 3);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

// One Lanczos step (iteration `iter`, 1-based) on unnormalised vectors: cur = w_iter (norm b_iter, 1 for iter 1),
// old = w_{iter-1}, u = scratch.  d_b[k] holds blanc(k+1) (d_b[0] = 0), d_a[k] alanc(k+1).  On return w_{iter+1} lies in
// `old`: the caller rotates (cur, old) -> (old, cur).
// one step; w_{k+1} goes to `out` (nullptr: over w_{k-1} in `old`, the rotating three-buffer form)
static int lanczos_step(edgpu_sector *s, double *cur, double *old, double *u, int iter, double *d_a, double *d_b, double *out = nullptr)
{
    if (!out) out = old;
    edgpu_ctx *ctx = s->ctx;
    const int64_t n = s->nalloc;
    const int nb = red_blocks(n);
    const double *one = ctx->d_scal + 3;
    const double *p_bprev = d_b + (iter - 1);
    const double *p_ncur = iter == 1 ? one : d_b + (iter - 1);
    const double *p_nold = iter <= 2 ? one : d_b + (iter - 2);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(ctx->d_scal + 4);    // zero between launches (reset by the last block)
    int ndot = -1;
    if (s->pl) {
        if (int rc = hxv_fiber(s, cur, u, ctx->d_dotpart, &ndot)) return rc;
    } else if (hxv_uses_star(s)) {
        if (int rc = hxv_star_dot(s, cur, u, ctx->d_dotpart, &ndot)) return rc;
    } else if (int rc = hxv_dispatch(s, cur, u)) return rc;
    const bool sh = is_sharded(s);
    if (ndot >= 0 && sh) {
        // partial <w,u> of the local pair tiles -> sum over the ranks -> / b^2
        k_reduce_final<<<1, 256, 0, ctx->stream>>>(ctx->d_dotpart, ndot, d_a + (iter - 1), 0);
        if (int rc = comm_allreduce_sum(ctx, d_a + (iter - 1), 1)) return rc;
        k_fin<<<1, 1, 0, ctx->stream>>>(d_a + (iter - 1), 2, p_ncur);
    } else if (ndot >= 0) {
        k_lanc_alpha<<<1, 256, 0, ctx->stream>>>(ctx->d_dotpart, ndot, p_ncur, d_a + (iter - 1));
    } else {
        k_dot_alpha<<<nb, kRedThreads, 0, ctx->stream>>>(cur, u, n, ctx->d_partials, p_ncur, d_a + (iter - 1), ticket);
    }
    k_lanc_c_norm<<<nb, kRedThreads, 0, ctx->stream>>>(old, out, u, cur, p_bprev, p_ncur, p_nold, d_a + (iter - 1), n, ctx->d_partials,
                                                       d_b + iter, ticket, sh ? 0 : 1);
    if (sh) {
        if (int rc = comm_allreduce_sum(ctx, d_b + iter, 1)) return rc;
        k_fin<<<1, 1, 0, ctx->stream>>>(d_b + iter, 1, nullptr);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

static int normalise(edgpu_sector *s, double *v)
{
    edgpu_ctx *ctx = s->ctx;
    const int64_t n = s->nalloc;
    const int nb = red_blocks(n);
    k_dot<<<nb, kRedThreads, 0, ctx->stream>>>(v, v, n, ctx->d_partials);
    if (is_sharded(s)) {
        k_reduce_final<<<1, 256, 0, ctx->stream>>>(ctx->d_partials, nb, ctx->d_scal + kScalB, 0);
        if (int rc = comm_allreduce_sum(ctx, ctx->d_scal + kScalB, 1)) return rc;
        k_fin<<<1, 1, 0, ctx->stream>>>(ctx->d_scal + kScalB, 1, nullptr);
    } else {
        k_reduce_final<<<1, 256, 0, ctx->stream>>>(ctx->d_partials, nb, ctx->d_scal + kScalB, 1);
    }
    k_div<<<nb, kRedThreads, 0, ctx->stream>>>(v, ctx->d_scal + kScalB, n);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

// tql2 (EISPACK, as carried in .repo/PLAIN_LANCZOS.f90:427-565): host, tiny.  z == nullptr: eigenvalues only.
int host_tql2(int n, double *d, double *e, double *z);

extern "C" int edgpu_lanczos_tridiag(edgpu_sector *s, edgpu_vec *v, int32_t nlanc, double threshold,
                                     double *alfa, double *beta, int32_t *nused)
{
    if (!s || !v || v->s != s) return s ? edgpu_fail(s->ctx, "edgpu_lanczos_tridiag: bad vector/sector") : 1;
    edgpu_ctx *ctx = s->ctx;
    if (nlanc < 1 || nlanc > kLancMax) return edgpu_fail(ctx, "edgpu_lanczos_tridiag: nlanc=%d out of range", nlanc);
    if (int rc = check_comm(s)) return rc;
    double *cur = v->d, *old, *u;
    if (int rc = sector_work(s, 0, &old)) return rc;
    if (int rc = sector_work(s, 1, &u)) return rc;
    const int64_t n = s->nalloc;
    CUDA_TRY(ctx, cudaMemsetAsync(old, 0, sizeof(double) * (size_t)n, ctx->stream));
    if (int rc = reset_scalars(ctx)) return rc;
    if (int rc = normalise(s, cur)) return rc;                 // iter==1: vin=vin/norm ; b=0   (:94-99)
    double *d_a = ctx->d_scal + kScalArr, *d_b = ctx->d_scal + kScalArr + kLancMax;
    // d_b[k] holds Fortran blanc(k+1): d_b[0] = 0 (b entering iteration 1)
    for (int iter = 1; iter <= nlanc; iter++) {
        if (int rc = lanczos_step(s, cur, old, u, iter, d_a, d_b)) return rc;
        double *t = cur; cur = old; old = t;                    // vout = vin ; vin = tmp/b  (unnormalised)
    }
    std::vector<double> ha(nlanc), hb(nlanc + 1);
    CUDA_TRY(ctx, cudaMemcpyAsync(ha.data(), d_a, sizeof(double) * nlanc, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(hb.data(), d_b, sizeof(double) * (nlanc + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    // the vector handle must keep pointing at memory it owns: the buffers rotated, data is destroyed anyway
    // reproduce the reference's exit: alanc(iter)=a ; if(iter<n) blanc(iter+1)=b ; if(|b|<thr) exit   (:172-174)
    int used = 0;
    for (int i = 0; i < nlanc; i++) { alfa[i] = 0.0; beta[i] = 0.0; }
    for (int iter = 1; iter <= nlanc; iter++) {
        alfa[iter - 1] = ha[iter - 1];
        if (iter < nlanc) beta[iter] = hb[iter];
        used = iter;
        if (std::fabs(hb[iter]) < threshold || !std::isfinite(hb[iter])) break;
    }
    if (nused) *nused = used;
    return 0;
}

extern "C" int edgpu_lanczos_gs(edgpu_sector *s, edgpu_vec *v0, int32_t nitermax, double threshold, int32_t ncheck,
                                double *e0, int32_t *nlanc_out, double *alanc_out, double *blanc_out)
{
    if (!s || !v0 || v0->s != s) return s ? edgpu_fail(s->ctx, "edgpu_lanczos_gs: bad vector/sector") : 1;
    edgpu_ctx *ctx = s->ctx;
    if (nitermax < 1 || nitermax > kLancMax - 1) return edgpu_fail(ctx, "edgpu_lanczos_gs: nitermax=%d out of range", nitermax);
    if (int rc = check_comm(s)) return rc;
    if (ncheck < 1) ncheck = 10;
    const int64_t n = s->nalloc;
    double *w0, *w1, *w2;
    if (int rc = sector_work(s, 0, &w0)) return rc;
    if (int rc = sector_work(s, 1, &w1)) return rc;
    if (int rc = sector_work(s, 2, &w2)) return rc;
    double *d_a = ctx->d_scal + kScalArr, *d_b = ctx->d_scal + kScalArr + kLancMax;
    std::vector<double> alanc(nitermax + 1, 0.0), blanc(nitermax + 2, 0.0), diag(nitermax + 1), sub(nitermax + 1),
        esave(nitermax + 2, 0.0), Z;
    int nlanc = 0;

    // Stored basis: while the Lanczos vectors fit a budget (a twentieth of the device memory per context) every w_k goes to a
    // buffer of its own instead of over w_{k-2}, and the Ritz vector is ONE pass over them -- the reference (and the fallback
    // below) regenerates all of them in a second Lanczos run, which doubled the time of the sector scan.
    const size_t vbytes = (sizeof(double) * (size_t)n + 255) & ~(size_t)255;
    const size_t budget = ctx->mem_bytes > 0 ? (size_t)ctx->mem_bytes / 20 : ((size_t)8 << 30);
    std::vector<double *> basis;                    // basis[k - 1] = w_k (unnormalised), k = 1, 2, ...
    bool store = vbytes * 8 <= budget && !(ctx->par.reserved[0] & 262144);
    auto drop_basis = [&]() { basis.clear(); };      // the arena keeps its chunks for the next sector
    size_t a_chunk = 0, a_off = 0;
    // next vector of the arena; false: the budget is spent (the caller falls back to the rotating buffers)
    auto carve = [&](double **q) -> bool {
        for (;;) {
            if (a_chunk < ctx->arena.size()) {
                auto &c = ctx->arena[a_chunk];
                if (a_off + vbytes <= c.second) { *q = reinterpret_cast<double *>(static_cast<char *>(c.first) + a_off); a_off += vbytes; return true; }
                a_chunk++; a_off = 0;
                continue;
            }
            if (ctx->arena_bytes + vbytes > budget) {
                // chunks of earlier, smaller sectors that cannot hold one vector of this one give their share back
                bool freed = false;
                for (size_t k = ctx->arena.size(); k-- > 0;)
                    if (ctx->arena[k].second < vbytes) {
                        if (!freed) cudaStreamSynchronize(ctx->stream);
                        cudaFree(ctx->arena[k].first);
                        ctx->arena_bytes -= ctx->arena[k].second;
                        ctx->arena.erase(ctx->arena.begin() + k);
                        if (k < a_chunk) a_chunk--;
                        freed = true;
                    }
                if (!freed || ctx->arena_bytes + vbytes > budget) return false;
            }
            size_t want = std::max<size_t>(vbytes * 16, (size_t)256 << 20);
            want = std::min(want, budget - ctx->arena_bytes);
            want -= want % vbytes;
            void *p = nullptr;
            if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return false; }
            ctx->arena.emplace_back(p, want);
            ctx->arena_bytes += want;
        }
    };

    // ---- pass 1: tridiagonalise until the lowest Ritz value stops moving (:328-360) ----
    double *vin = w0, *vout = w1, *tmp = w2;
    if (store) {
        double *q = nullptr;
        if (carve(&q)) { basis.push_back(q); vin = q; }
        else store = false;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(vin, v0->d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(vout, 0, sizeof(double) * (size_t)n, ctx->stream));
    if (int rc = reset_scalars(ctx)) return rc;
    if (int rc = normalise(s, vin)) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scal, ctx->d_scal + kScalB, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_scal[0] == 0.0) { drop_basis(); return edgpu_fail(ctx, "lanczos_plain_iteration: norm =0!!"); }
    // The reference diagonalises the growing tridiagonal and tests |dE0| after EVERY step.  Small sectors are bound by
    // that host round trip, so the device runs kBatch steps ahead and the host then replays the reference's rule step
    // by step over the new (a, b) pairs: the stopping iteration is identical, at most kBatch - 1 steps are computed in
    // vain (pass 2 regenerates exactly nlanc vectors).  Only eigenVALUES are needed here (tql2 without vectors gives
    // bit-identical values); the vectors are computed once, after the loop.
    const int kBatch = s->dim < (1ll << 22) ? 8 : 1;
    bool stop = false;
    for (int iter0 = 1; iter0 <= nitermax && !stop; iter0 += kBatch) {
        const int nb = std::min(kBatch, nitermax - iter0 + 1);
        for (int k = 0; k < nb; k++) {
            double *qn = nullptr;
            if (store && !carve(&qn)) {
                // out of budget: fall back to the rotating three buffers (and to the regenerating second pass)
                CUDA_TRY(ctx, cudaMemcpyAsync(w0, vin, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
                if (vout != w1) CUDA_TRY(ctx, cudaMemcpyAsync(w1, vout, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
                vin = w0; vout = w1;
                drop_basis();
                store = false;
            }
            if (store) {
                basis.push_back(qn);
                if (int rc = lanczos_step(s, vin, vout, tmp, iter0 + k, d_a, d_b, basis.back())) { drop_basis(); return rc; }
                vout = vin; vin = basis.back();
                continue;
            }
            if (int rc = lanczos_step(s, vin, vout, tmp, iter0 + k, d_a, d_b)) return rc;     // (cur, old, scratch)
            double *t = vin; vin = vout; vout = t;
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scal, d_a + (iter0 - 1), sizeof(double) * nb, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scal + 8, d_b + iter0, sizeof(double) * nb, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < nb; k++) {
            const int iter = iter0 + k;
            const double a = ctx->h_scal[k], b = ctx->h_scal[8 + k];
            if (std::fabs(b) < threshold) { stop = true; break; }               // :333
            nlanc++;
            alanc[iter - 1] = a;
            blanc[iter] = b;
            for (int i = 0; i < nlanc; i++) { diag[i] = alanc[i]; sub[i] = i > 0 ? blanc[i] : 0.0; }
            host_tql2(nlanc, diag.data(), sub.data(), nullptr);
            if (nlanc >= ncheck) {                                              // :352-359
                esave[nlanc - (ncheck - 1)] = diag[0];
                if (nlanc >= ncheck + 1) {
                    double diff = esave[nlanc - (ncheck - 1)] - esave[nlanc - (ncheck - 1) - 1];
                    if (std::fabs(diff) <= threshold) { stop = true; break; }
                }
            }
        }
    }
    if (nlanc == 0) { drop_basis(); return edgpu_fail(ctx, "edgpu_lanczos_gs: Lanczos broke down at the first step"); }
    for (int i = 0; i < nlanc; i++) { diag[i] = alanc[i]; sub[i] = i > 0 ? blanc[i] : 0.0; }
    Z.assign((size_t)nlanc * nlanc, 0.0);
    for (int i = 0; i < nlanc; i++) Z[i + (size_t)nlanc * i] = 1.0;
    host_tql2(nlanc, diag.data(), sub.data(), Z.data());
    if (e0) *e0 = diag[0];

    if (store) {
        // ---- Ritz vector sum_k Z(k,1) v_k = sum_k Z(k,1)/b_k w_k from the stored basis (b_1 = 1) ----
        std::vector<double> coef(nlanc);
        for (int k = 1; k <= nlanc; k++) coef[k - 1] = Z[k - 1] / (k == 1 ? 1.0 : blanc[k - 1]);
        void *d_tab = nullptr;
        int rc = pool_alloc(ctx, (sizeof(double *) + sizeof(double)) * (size_t)nlanc, &d_tab);
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(d_tab, basis.data(), sizeof(double *) * (size_t)nlanc, cudaMemcpyHostToDevice, ctx->stream);
            double *d_coef = reinterpret_cast<double *>(reinterpret_cast<double **>(d_tab) + nlanc);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_coef, coef.data(), sizeof(double) * (size_t)nlanc, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) {
                k_ritz<<<red_blocks(n), kRedThreads, 0, ctx->stream>>>(reinterpret_cast<const double *const *>(d_tab), d_coef, nlanc, v0->d, n);
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);       // `coef`, `basis` are host sources of async copies
            if (e != cudaSuccess) rc = edgpu_fail(ctx, "edgpu_lanczos_gs: %s", cudaGetErrorString(e));
            pool_release(ctx, d_tab);
        }
        drop_basis();
        if (rc) return rc;
        if (int rc2 = normalise(s, v0->d)) return rc2;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (nlanc_out) *nlanc_out = nlanc;
        if (alanc_out) for (int i = 0; i < nlanc; i++) alanc_out[i] = alanc[i];
        if (blanc_out) for (int i = 0; i < nlanc; i++) blanc_out[i] = i > 0 ? blanc[i] : 0.0;
        return 0;
    }
    // ---- pass 2: regenerate the Lanczos vectors and accumulate the Ritz vector sum_k Z(k,1) v_k (:375-384,
    //      with Z(k,1) paired with the k-th basis vector) ----
    vin = w0; vout = w1; tmp = w2;
    CUDA_TRY(ctx, cudaMemcpyAsync(vin, v0->d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(vout, 0, sizeof(double) * (size_t)n, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(v0->d, 0, sizeof(double) * (size_t)n, ctx->stream));
    if (int rc = reset_scalars(ctx)) return rc;
    if (int rc = normalise(s, vin)) return rc;
    for (int iter = 1; iter <= nlanc; iter++) {
        // vin holds the unnormalised w_iter = b_iter v_iter (b_1 = 1); the kernels are deterministic, so the b of this
        // pass are bit-identical to the blanc of pass 1
        const double nrm = iter == 1 ? 1.0 : blanc[iter - 1];
        if (int rc = vec_axpy(ctx, v0->d, vin, Z[iter - 1] / nrm, n)) return rc;     // Z(iter,1)
        if (iter == nlanc) break;
        if (int rc = lanczos_step(s, vin, vout, tmp, iter, d_a, d_b)) return rc;
        { double *t = vin; vin = vout; vout = t; }
    }
    if (int rc = normalise(s, v0->d)) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (nlanc_out) *nlanc_out = nlanc;
    if (alanc_out) for (int i = 0; i < nlanc; i++) alanc_out[i] = alanc[i];
    if (blanc_out) for (int i = 0; i < nlanc; i++) blanc_out[i] = i > 0 ? blanc[i] : 0.0;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// EISPACK tql2 (host).  d(1:n) diagonal, e(2:n) sub-diagonal in e[1..n-1], z column-major n x n.
// ---------------------------------------------------------------------------------------------------------
static double h_pythag(double a, double b)
{
    double p = std::fmax(std::fabs(a), std::fabs(b));
    if (p == 0.0) return 0.0;
    double r = std::fmin(std::fabs(a), std::fabs(b)) / p;
    r *= r;
    for (;;) {
        double t = 4.0 + r;
        if (t == 4.0) break;
        double s = r / t, u = 1.0 + 2.0 * s;
        p *= u;
        r *= (s / u) * (s / u);
    }
    return p;
}

int host_tql2(int n, double *d, double *e, double *z)
{
    if (n == 1) return 0;
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    for (int l = 0; l < n; l++) {
        int iters = 0;
        double h = std::fabs(d[l]) + std::fabs(e[l]);
        if (tst1 < h) tst1 = h;
        int m = l;
        while (m < n && !(tst1 + std::fabs(e[m]) == tst1)) m++;
        if (m != l) {
            double tst2;
            do {
                if (iters++ == 30) return l + 1;
                const int l1 = l + 1;
                double g = d[l];
                double p = (d[l1] - g) / (2.0 * e[l]);
                double r = h_pythag(p, 1.0);
                const double sr = std::copysign(r, p);
                d[l] = e[l] / (p + sr);
                d[l1] = e[l] * (p + sr);
                const double dl1 = d[l1];
                h = g - d[l];
                for (int i = l1 + 1; i < n; i++) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
                const double el1 = e[l1];
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = h_pythag(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    if (z) {
                        double *zi = z + (size_t)n * i, *zi1 = z + (size_t)n * (i + 1);
                        for (int k = 0; k < n; k++) {
                            const double hh = zi1[k];
                            zi1[k] = s * zi[k] + c * hh;
                            zi[k] = c * zi[k] - s * hh;
                        }
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
                tst2 = tst1 + std::fabs(e[l]);
            } while (tst2 > tst1);
        }
        d[l] += f;
    }
    for (int ii = 1; ii < n; ii++) {
        int i = ii - 1, k = i;
        double p = d[i];
        for (int j = ii; j < n; j++) if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i]; d[i] = p;
            if (z) for (int j = 0; j < n; j++) std::swap(z[j + (size_t)n * i], z[j + (size_t)n * k]);
        }
    }
    return 0;
}
