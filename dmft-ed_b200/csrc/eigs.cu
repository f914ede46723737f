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
 * nmatvec: H*v applications.
 * A Krylov space of ONE start vector holds one vector per eigenspace, so a degenerate level shows up once (ARPACK finds
 * the copies only through rounding noise).  Multiplicities matter here -- every member of a degenerate ground level must
 * enter the state list (ED_DIAG.f90:224-235) -- so the converged vectors are LOCKED and the run is repeated from a fresh
 * random vector in their orthogonal complement until the lowest value found there lies above the neigen-th locked one. */
extern "C" int edgpu_lanczos_eigs(edgpu_sector *s, int32_t neigen, int32_t ncv, int32_t maxrestart, double tol, uint64_t seed,
                                  double *evals, edgpu_vec **vecs, int32_t *nconv, int32_t *nmatvec)
{
    if (!s || !evals || !vecs || neigen < 1) return s ? edgpu_fail(s->ctx, "edgpu_lanczos_eigs: bad arguments") : 1;
    edgpu_ctx *ctx = s->ctx;
    if (s->shard_nranks > 1) return edgpu_fail(ctx, "edgpu_lanczos_eigs: not available on sharded sectors");
    const int64_t n = s->nalloc;
    if (neigen > s->dim) return edgpu_fail(ctx, "edgpu_lanczos_eigs: neigen=%d exceeds the sector dimension", neigen);
    const int maxlock = (int)std::min<int64_t>(s->dim, 2 * neigen + 2);            // locked vectors kept at V[0..nlock)
    int m = (int)std::min<int64_t>(std::min<int64_t>(ncv, kMaxBasis - maxlock), s->dim);
    if (m < neigen + 1) m = (int)std::min<int64_t>(s->dim, neigen + 1);
    if (m + maxlock > kMaxBasis) return edgpu_fail(ctx, "edgpu_lanczos_eigs: basis of %d + %d locked vectors exceeds %d", m, maxlock, kMaxBasis);
    if (maxrestart < 1) maxrestart = 300;
    const int nvtot = maxlock + m + 1;
    double *V = nullptr, *d_part = nullptr, *d_h = nullptr, *d_S = nullptr;
    if (int rc0 = pool_alloc(ctx, sizeof(double) * (size_t)n * (size_t)nvtot, (void **)&V)) return rc0;
    auto cleanup = [&]() { pool_release(ctx, V); cudaFree(d_part); cudaFree(d_h); cudaFree(d_S); };
    if (cudaMalloc(&d_part, sizeof(double) * (size_t)kMdBlocks * kMaxBasis) != cudaSuccess || cudaMalloc(&d_h, sizeof(double) * 2 * kMaxBasis) != cudaSuccess ||
        cudaMalloc(&d_S, sizeof(double) * (size_t)kMaxBasis * kMaxBasis) != cudaSuccess) { cleanup(); return edgpu_fail(ctx, "edgpu_lanczos_eigs: scratch allocation failed"); }
    cudaStream_t st = ctx->stream;
    const int nb = (int)std::min<int64_t>(kMdBlocks, (n + kMdThreads - 1) / kMdThreads);
    auto vec = [&](int i) { return V + (size_t)i * n; };
    int rc = 0, matvecs = 0;
    std::vector<double> h(kMaxBasis);
    // orthogonalise w against V_0..V_{nv-1} (locked + active) twice; the summed coefficients land in hout[0..nv)
    auto orth = [&](double *w, int nv, double *hout) -> int {
        if (nv == 0) return 0;
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
    CUDA_TRY(ctx, cudaMemsetAsync(V, 0, sizeof(double) * (size_t)n * (size_t)nvtot, st));
    std::vector<double> lockval;                      // eigenvalues of the locked vectors V[0..nlock), ascending
    int nlock = 0, converged_total = 0;
    double tnorm = 0.0;
    // one thick-restart run in the complement of the locked vectors: the `want` lowest Ritz pairs end up in
    // V[nlock .. nlock+got), values in out[]
    auto run = [&](int want, uint64_t sd, std::vector<double> &out, int &got, int &nconverged) -> int {
        const int base = nlock;
        auto act = [&](int i) { return vec(base + i); };
        const int mrun = (int)std::min<int64_t>(m, s->dim - nlock);
        want = std::min(want, mrun);
        got = 0; nconverged = 0;
        if (want < 1) return 0;
        std::vector<double> T((size_t)mrun * mrun, 0.0);
        if (int r = vec_fill_random(s, 1, sd, act(0))) return r;
        if (orth(act(0), base, h.data())) return edgpu_fail(ctx, "edgpu_lanczos_eigs: orthogonalisation failed");
        double nrm = 0.0;
        if (norm_of(act(0), nrm) || nrm == 0.0) return edgpu_fail(ctx, "edgpu_lanczos_eigs: zero start vector");
        if (int r = vec_scale(ctx, act(0), 1.0 / nrm, n)) return r;
        int k = 0, mcur = mrun;
        double beta_m = 0.0;
        for (int restart = 0; restart <= maxrestart; restart++) {
            bool invariant = false;
            for (int j = k; j < mcur; j++) {
                double *w = act(j + 1);                               // the next basis slot doubles as the work vector
                if (int r = hxv_dispatch(s, act(j), w)) return r;
                matvecs++;
                if (orth(w, base + j + 1, h.data())) return edgpu_fail(ctx, "edgpu_lanczos_eigs: orthogonalisation failed");
                for (int i = 0; i <= j; i++) { T[i + (size_t)mrun * j] = h[base + i]; T[j + (size_t)mrun * i] = h[base + i]; }
                double b = 0.0;
                if (norm_of(w, b)) return edgpu_fail(ctx, "edgpu_lanczos_eigs: norm failed");
                tnorm = std::max(tnorm, std::fabs(h[base + j]) + b);
                if (b <= 1e-12 * std::max(1.0, tnorm)) {              // invariant subspace: the Ritz pairs of T[0..j] are exact
                    mcur = j + 1; beta_m = 0.0; invariant = true;
                    break;
                }
                if (int r = vec_scale(ctx, w, 1.0 / b, n)) return r;
                if (j + 1 < mcur) { T[(j + 1) + (size_t)mrun * j] = b; T[j + (size_t)mrun * (j + 1)] = b; }
                else beta_m = b;                                       // w = residual direction
            }
            std::vector<double> A((size_t)mcur * mcur), wv(mcur);
            for (int i = 0; i < mcur; i++)
                for (int j = 0; j < mcur; j++) A[i + (size_t)mcur * j] = T[i + (size_t)mrun * j];
            if (ed_host_eigh(mcur, A.data(), wv.data())) return edgpu_fail(ctx, "edgpu_lanczos_eigs: dense eigensolver failed");
            const int w2 = std::min(want, mcur);
            int conv = 0;
            for (int i = 0; i < w2; i++) {
                const double res = std::fabs(beta_m * A[(mcur - 1) + (size_t)mcur * i]);
                const double thr = std::max(tol, 2.3e-16) * std::max(3.7e-11, std::fabs(wv[i]));      // eps^(2/3), ARPACK dsconv
                if (res <= thr) conv++;
            }
            const bool last = invariant || conv == w2 || restart == maxrestart;
            const int K = last ? w2 : std::min(mcur - 1, want + std::max(1, (mcur - want) / 2));
            CUDA_TRY(ctx, cudaMemcpyAsync(d_S, A.data(), sizeof(double) * (size_t)mcur * mcur, cudaMemcpyHostToDevice, st));
            {
                const size_t smem = sizeof(double) * (size_t)mcur * kCombT;
                static size_t cur_smem = 0;
                if (smem > cur_smem) { CUDA_TRY(ctx, cudaFuncSetAttribute((const void *)k_combine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur_smem = smem; }
                const int gb = (int)std::min<int64_t>(ctx->sm_count * 2, (n + kCombT - 1) / kCombT);
                k_combine<<<gb, kCombT, smem, st>>>(act(0), n, mcur, K, d_S, n);
                CUDA_TRY(ctx, cudaGetLastError());
            }
            if (last) {
                out.assign(wv.begin(), wv.begin() + w2);
                got = w2; nconverged = conv;
                CUDA_TRY(ctx, cudaStreamSynchronize(st));
                return 0;
            }
            k_copy_scaled<<<nb, kMdThreads, 0, st>>>(act(K), act(mcur), 1.0, n);
            std::fill(T.begin(), T.end(), 0.0);
            for (int i = 0; i < K; i++) {
                T[i + (size_t)mrun * i] = wv[i];
                const double a = beta_m * A[(mcur - 1) + (size_t)mcur * i];
                T[K + (size_t)mrun * i] = a; T[i + (size_t)mrun * K] = a;
            }
            CUDA_TRY(ctx, cudaStreamSynchronize(st));
            k = K;
        }
        return 0;
    };
    // sweep 1: the neigen lowest (distinct) levels; further sweeps: missed copies of degenerate levels
    for (int sweep = 0; sweep < 2 * neigen + 2; sweep++) {
        std::vector<double> out;
        int got = 0, nconvd = 0;
        const int want = sweep == 0 ? neigen : std::min(neigen, 2);
        if ((rc = run(want, seed + 7919ull * (uint64_t)sweep, out, got, nconvd))) { cleanup(); return rc; }
        if (got == 0) break;                                        // the complement is empty
        if (sweep == 0) {
            lockval = out; nlock = got; converged_total = nconvd;
            if (nlock >= s->dim) break;
            continue;
        }
        // keep the new vectors whose value belongs to the neigen lowest of (locked + new)
        std::vector<double> sorted(lockval);
        std::sort(sorted.begin(), sorted.end());
        const double cut = (int)sorted.size() >= neigen ? sorted[neigen - 1] : 1e300;
        const double slack = 1e-9 * std::max(1.0, std::fabs(cut));
        int take = 0;
        while (take < got && (out[take] <= cut + slack) && nlock + take < maxlock) take++;
        if (take == 0) break;                                       // nothing below the neigen-th level was missed
        for (int i = 0; i < take; i++) lockval.push_back(out[i]);
        nlock += take;
        converged_total += std::min(take, nconvd);
        // the locked set stays orthonormal (new vectors were built in the complement); order is fixed at the end
    }
    // lowest neigen of the locked pairs, ascending
    std::vector<int> idx(lockval.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return lockval[a] < lockval[b]; });
    const int have = (int)std::min<size_t>(idx.size(), (size_t)neigen);
    for (int i = 0; i < neigen; i++) vecs[i] = nullptr;
    for (int i = 0; i < have; i++) {
        evals[i] = lockval[idx[i]];
        if ((rc = edgpu_vec_alloc(s, &vecs[i]))) { cleanup(); return rc; }
        CUDA_TRY(ctx, cudaMemcpyAsync(vecs[i]->d, vec(idx[i]), sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    cleanup();
    if (nconv) *nconv = std::min(converged_total, have);
    if (nmatvec) *nmatvec = matvecs;
    if (have < neigen) return edgpu_fail(ctx, "edgpu_lanczos_eigs: only %d of %d eigenpairs found", have, neigen);
    return 0;
}
