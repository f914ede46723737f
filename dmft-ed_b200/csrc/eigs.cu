// eigs.cu -- several lowest eigenpairs of a sector on the device: thick-restart Lanczos with full re-orthogonalisation.
//
// Replaces sp_eigh (SciFortran's ARPACK 'SA' driver) as called at ED_DIAG.f90:149-166 with Neigen = lanc_nstates_sector
// eigenpairs and a basis of Nblock = min(dim, lanc_ncv_factor*Neigen + lanc_ncv_add) vectors.  ARPACK itself is an
// un-vendored dependency; the algorithm restated here is the symmetric thick-restart variant of the implicitly restarted
// Lanczos method it implements (same Krylov subspaces and the same Ritz extraction; only the restart is explicit):
//   expand the orthonormal basis V_m with H v_j, orthogonalised against ALL basis vectors (classical Gram-Schmidt
//   applied twice), T = V^T H V (m x m, dense on the host), Ritz pairs (theta_i, V s_i); a pair is converged when the
//   residual estimate |beta_m s_i(m)| <= tol * max(eps^(2/3), |theta_i|) (ARPACK's criterion, dsconv); otherwise keep the
//   K lowest Ritz vectors + the residual direction and continue (T becomes diag(theta) + an arrow row).
// All vector work is on the device: one fused pass per Gram-Schmidt sweep (k_mdot / k_maxpy read every basis vector
// once), the restart V <- V S in place through shared-memory tiles.
#include "edgpu_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

extern "C" int ed_host_eigh(int32_t n, double *a_colmajor, double *w);
int vec_fill_random(edgpu_sector *s, int uniform, uint64_t seed, double *dst);
int vec_scale(edgpu_ctx *ctx, double *v, double alpha, int64_t n);

static constexpr int kMaxBasis = 96;
static constexpr int kMdBlocks = 592;          // 148 SMs x 4
static constexpr int kMdThreads = 256;

// partial[b][i] = sum over the elements of block b of V_i[e] * w[e], i < nv
__global__ void __launch_bounds__(kMdThreads) k_mdot(const double *__restrict__ base, int64_t stride, int nv, const double *__restrict__ w,
                                                     int64_t n, double *__restrict__ partial)
{
    __shared__ double sh[kMdThreads / 32];
    for (int i0 = 0; i0 < nv; i0 += 8) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
            const double we = w[e];
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (i0 + q < nv) acc[q] = fma(base[(int64_t)(i0 + q) * stride + e], we, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            double v = acc[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
            __syncthreads();
            if (threadIdx.x == 0 && i0 + q < nv) {
                double t = 0.0;
                for (int k = 0; k < kMdThreads / 32; k++) t += sh[k];
                partial[(size_t)blockIdx.x * kMaxBasis + i0 + q] = t;
            }
            __syncthreads();
        }
    }
}

// h[i] = sum_b partial[b][i] (fixed order); hacc[i] += h[i]
__global__ void k_mdot_final(const double *__restrict__ partial, int nb, int nv, double *__restrict__ h, double *__restrict__ hacc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    double t = 0.0;
    for (int b = 0; b < nb; b++) t += partial[(size_t)b * kMaxBasis + i];
    h[i] = t;
    if (hacc) hacc[i] += t;
}

// w -= sum_i h[i] V_i
__global__ void __launch_bounds__(kMdThreads) k_maxpy(double *__restrict__ w, const double *__restrict__ base, int64_t stride, int nv,
                                                      const double *__restrict__ h, int64_t n)
{
    __shared__ double sh[kMaxBasis];
    for (int i = threadIdx.x; i < nv; i += blockDim.x) sh[i] = h[i];
    __syncthreads();
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        double acc = w[e];
        for (int i = 0; i < nv; i++) acc = fma(-sh[i], base[(int64_t)i * stride + e], acc);
        w[e] = acc;
    }
}

// in place V[:, 0..K) <- V[:, 0..m) S[0..m, 0..K)   (S column-major m x m in global memory); tile of kCombT elements
static constexpr int kCombT = 128;
__global__ void __launch_bounds__(kCombT) k_combine(double *__restrict__ base, int64_t stride, int m, int K, const double *__restrict__ S, int64_t n)
{
    extern __shared__ double tile[];                     // [m][kCombT]
    for (int64_t e0 = (int64_t)blockIdx.x * kCombT; e0 < n; e0 += (int64_t)gridDim.x * kCombT) {
        const int64_t e = e0 + threadIdx.x;
        if (e < n)
            for (int j = 0; j < m; j++) tile[j * kCombT + threadIdx.x] = base[(int64_t)j * stride + e];
        __syncthreads();
        if (e < n)
            for (int k = 0; k < K; k++) {
                double acc = 0.0;
                for (int j = 0; j < m; j++) acc = fma(S[j + (size_t)m * k], tile[j * kCombT + threadIdx.x], acc);
                base[(int64_t)k * stride + e] = acc;
            }
        __syncthreads();
    }
}

__global__ void k_copy_scaled(double *__restrict__ dst, const double *__restrict__ src, double a, int64_t n)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) dst[e] = a * src[e];
}

/* sp_eigh (ED_DIAG.f90:149-166): the `neigen` lowest eigenpairs of the sector.  ncv: basis size (Nblock); tol: lanc_tolerance;
 * evals[neigen] ascending; vecs[neigen]: handles allocated here (caller frees with edgpu_vec_free); nconv: converged pairs;
 * nmatvec: H*v applications. */
extern "C" int edgpu_lanczos_eigs(edgpu_sector *s, int32_t neigen, int32_t ncv, int32_t maxrestart, double tol, uint64_t seed,
                                  double *evals, edgpu_vec **vecs, int32_t *nconv, int32_t *nmatvec)
{
    if (!s || !evals || !vecs || neigen < 1) return s ? edgpu_fail(s->ctx, "edgpu_lanczos_eigs: bad arguments") : 1;
    edgpu_ctx *ctx = s->ctx;
    if (s->shard_nranks > 1) return edgpu_fail(ctx, "edgpu_lanczos_eigs: not available on sharded sectors");
    const int64_t n = s->nalloc;
    int m = (int)std::min<int64_t>(std::min<int64_t>(ncv, kMaxBasis), s->dim);
    if (m < neigen + 1 && m < s->dim) m = (int)std::min<int64_t>(s->dim, neigen + 1);
    if (neigen > m) return edgpu_fail(ctx, "edgpu_lanczos_eigs: neigen=%d exceeds the basis size %d", neigen, m);
    if (maxrestart < 1) maxrestart = 300;
    double *V = nullptr, *d_part = nullptr, *d_h = nullptr, *d_S = nullptr;
    cudaError_t ce = cudaMalloc(&V, sizeof(double) * (size_t)n * (size_t)(m + 1));
    if (ce != cudaSuccess) return edgpu_fail(ctx, "edgpu_lanczos_eigs: cudaMalloc of %d basis vectors (%lld doubles each) failed: %s", m + 1, (long long)n, cudaGetErrorString(ce));
    auto cleanup = [&]() { cudaFree(V); cudaFree(d_part); cudaFree(d_h); cudaFree(d_S); };
    if (cudaMalloc(&d_part, sizeof(double) * (size_t)kMdBlocks * kMaxBasis) != cudaSuccess || cudaMalloc(&d_h, sizeof(double) * 2 * kMaxBasis) != cudaSuccess ||
        cudaMalloc(&d_S, sizeof(double) * (size_t)kMaxBasis * kMaxBasis) != cudaSuccess) { cleanup(); return edgpu_fail(ctx, "edgpu_lanczos_eigs: scratch allocation failed"); }
    cudaStream_t st = ctx->stream;
    const int nb = (int)std::min<int64_t>(kMdBlocks, (n + kMdThreads - 1) / kMdThreads);
    auto vec = [&](int i) { return V + (size_t)i * n; };
    int rc = 0, matvecs = 0;
    auto fail = [&](const char *msg) { cleanup(); return edgpu_fail(ctx, "edgpu_lanczos_eigs: %s", msg); };
    std::vector<double> h(kMaxBasis), T((size_t)m * m, 0.0), S((size_t)m * m), theta(m);
    // orthogonalise w against V_0..V_{nv-1} twice; the summed coefficients land in hout[0..nv)
    auto orth = [&](double *w, int nv, double *hout) -> int {
        cudaMemsetAsync(d_h + kMaxBasis, 0, sizeof(double) * kMaxBasis, st);
        for (int pass = 0; pass < 2; pass++) {
            k_mdot<<<nb, kMdThreads, 0, st>>>(V, n, nv, w, n, d_part);
            k_mdot_final<<<(nv + 63) / 64, 64, 0, st>>>(d_part, nb, nv, d_h, d_h + kMaxBasis);
            k_maxpy<<<nb, kMdThreads, 0, st>>>(w, V, n, nv, d_h, n);
        }
        if (cudaMemcpyAsync(hout, d_h + kMaxBasis, sizeof(double) * nv, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
        return cudaStreamSynchronize(st) != cudaSuccess;
    };
    auto norm_of = [&](const double *w, double &out) -> int {
        if (vec_dot(ctx, w, w, n, ctx->d_scal)) return 1;
        if (cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
        out = std::sqrt(ctx->h_scal[0]);
        return 0;
    };
    // start vector
    CUDA_TRY(ctx, cudaMemsetAsync(V, 0, sizeof(double) * (size_t)n * (size_t)(m + 1), st));
    if ((rc = vec_fill_random(s, 1, seed, vec(0)))) { cleanup(); return rc; }
    double nrm = 0.0;
    if (norm_of(vec(0), nrm) || nrm == 0.0) return fail("zero start vector");
    if ((rc = vec_scale(ctx, vec(0), 1.0 / nrm, n))) { cleanup(); return rc; }
    int k = 0, mcur = m, converged = 0;
    double beta_m = 0.0, tnorm = 0.0;
    bool done = false;
    for (int restart = 0; restart <= maxrestart && !done; restart++) {
        bool invariant = false;
        for (int j = k; j < mcur; j++) {
            double *w = vec(j + 1 <= m ? j + 1 : m);             // the next basis slot doubles as the work vector
            if ((rc = hxv_dispatch(s, vec(j), w))) { cleanup(); return rc; }
            matvecs++;
            if (orth(w, j + 1, h.data())) return fail("orthogonalisation failed");
            for (int i = 0; i <= j; i++) { T[i + (size_t)m * j] = h[i]; T[j + (size_t)m * i] = h[i]; }
            double b = 0.0;
            if (norm_of(w, b)) return fail("norm failed");
            tnorm = std::max(tnorm, std::fabs(h[j]) + b);
            if (b <= 1e-13 * std::max(1.0, tnorm)) {             // invariant subspace: the Ritz pairs of T[0..j] are exact
                mcur = j + 1; beta_m = 0.0; invariant = true;
                break;
            }
            if ((rc = vec_scale(ctx, w, 1.0 / b, n))) { cleanup(); return rc; }
            if (j + 1 < mcur) { T[(j + 1) + (size_t)m * j] = b; T[j + (size_t)m * (j + 1)] = b; }
            else beta_m = b;                                      // w = V_m: residual direction
        }
        // Ritz pairs of the mcur x mcur projection
        std::vector<double> A((size_t)mcur * mcur), wv(mcur);
        for (int i = 0; i < mcur; i++)
            for (int j = 0; j < mcur; j++) A[i + (size_t)mcur * j] = T[i + (size_t)m * j];
        if (ed_host_eigh(mcur, A.data(), wv.data())) return fail("dense eigensolver failed");
        const int want = std::min(neigen, mcur);
        converged = 0;
        for (int i = 0; i < want; i++) {
            const double res = std::fabs(beta_m * A[(mcur - 1) + (size_t)mcur * i]);
            const double thr = std::max(tol, 2.3e-16) * std::max(3.7e-11, std::fabs(wv[i]));      // eps^(2/3)
            if (res <= thr) converged++;
        }
        const bool last = invariant || converged == want || restart == maxrestart;
        const int K = last ? want : std::min(mcur - 1, neigen + std::max(1, (mcur - neigen) / 2));
        // V[:, 0..K) <- V S
        CUDA_TRY(ctx, cudaMemcpyAsync(d_S, A.data(), sizeof(double) * (size_t)mcur * mcur, cudaMemcpyHostToDevice, st));
        {
            const size_t smem = sizeof(double) * (size_t)mcur * kCombT;
            static size_t cur_smem = 0;
            if (smem > cur_smem) { CUDA_TRY(ctx, cudaFuncSetAttribute((const void *)k_combine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur_smem = smem; }
            const int gb = (int)std::min<int64_t>(ctx->sm_count * 2, (n + kCombT - 1) / kCombT);
            k_combine<<<gb, kCombT, smem, st>>>(V, n, mcur, K, d_S, n);
            CUDA_TRY(ctx, cudaGetLastError());
        }
        if (last) {
            for (int i = 0; i < want; i++) evals[i] = wv[i];
            done = true;
            break;
        }
        // residual direction becomes basis vector K; T = diag(theta) + arrow
        k_copy_scaled<<<nb, kMdThreads, 0, st>>>(vec(K), vec(mcur), 1.0, n);
        std::fill(T.begin(), T.end(), 0.0);
        for (int i = 0; i < K; i++) {
            T[i + (size_t)m * i] = wv[i];
            const double a = beta_m * A[(mcur - 1) + (size_t)mcur * i];
            T[K + (size_t)m * i] = a; T[i + (size_t)m * K] = a;
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        k = K;
    }
    const int want = std::min(neigen, mcur);
    for (int i = 0; i < neigen; i++) vecs[i] = nullptr;
    for (int i = 0; i < want; i++) {
        if ((rc = edgpu_vec_alloc(s, &vecs[i]))) { cleanup(); return rc; }
        CUDA_TRY(ctx, cudaMemcpyAsync(vecs[i]->d, vec(i), sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    cleanup();
    if (nconv) *nconv = converged;
    if (nmatvec) *nmatvec = matvecs;
    if (want < neigen) return edgpu_fail(ctx, "edgpu_lanczos_eigs: the Krylov space closed after %d vectors (< neigen=%d)", want, neigen);
    return 0;
}
