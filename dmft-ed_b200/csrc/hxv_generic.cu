// hxv_generic.cu -- table-driven on-the-fly H*v (any same-spin hop structure, any layout).
//
// Replaces directMatVec_cc (ED_HAMILTONIAN_DIRECT_HxV.f90:21-92 + ED_HAMILTONIAN/direct/{HxVimp,HxVint,
// HxVbath,HxVimp_bath}.f90) in GATHER form (no atomics; the reference's scatter form defines the same
// Hermitian operator):
//   y[rd][ru] = (Eup[ru] + Edw[rd] + X[imp(u)][imp(d)]) x[rd][ru]
//             + sum_j ampU[j] x[rd][tgtU_j(ru)] + sum_j ampD[j] x[tgtD_j(rd)][ru]     (+ Jx/Jp two-spin terms)
// This is the general (and slower) kernel; hxv_star.cu holds the tiled star-product kernels that the
// BASELINE configurations use.
#include "edgpu_internal.h"

__global__ void __launch_bounds__(256)
k_hxv_generic(int64_t dim_up, int64_t dim_dw, int64_t ld, int norb,
              const uint32_t *__restrict__ cfg_up, const uint32_t *__restrict__ cfg_dw,
              const double *__restrict__ e_up, const double *__restrict__ e_dw, const double *__restrict__ xtab,
              const uint32_t *__restrict__ hop_up, const uint8_t *__restrict__ nhop_up, const double *__restrict__ amp_up,
              const uint32_t *__restrict__ hop_dw, const uint8_t *__restrict__ nhop_dw, const double *__restrict__ amp_dw,
              const double *__restrict__ x, double *__restrict__ y)
{
    __shared__ double s_ampu[256], s_ampd[256], s_x[32 * 32];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_ampu[i] = amp_up[i]; s_ampd[i] = amp_dw[i]; }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_x[i] = xtab[i];
    __syncthreads();
    const int64_t ru = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t impmask = (1u << norb) - 1u;
    if (ru >= dim_up) return;
    const uint32_t ui = cfg_up[ru] & impmask;
    const double eu = e_up[ru];
    const int nu = nhop_up[ru];
    for (int64_t rd = blockIdx.y; rd < dim_dw; rd += gridDim.y) {
        const double *xr = x + rd * ld;
        const uint32_t di = cfg_dw[rd] & impmask;
        double acc = (eu + e_dw[rd] + s_x[di * 32 + ui]) * xr[ru];
        for (int j = 0; j < nu; j++) {
            uint32_t h = hop_up[(int64_t)j * dim_up + ru];
            acc += s_ampu[h & 255u] * xr[h >> 8];
        }
        const int nd = nhop_dw[rd];
        for (int j = 0; j < nd; j++) {
            uint32_t h = hop_dw[(int64_t)j * dim_dw + rd];
            acc += s_ampd[h & 255u] * x[(int64_t)(h >> 8) * ld + ru];
        }
        y[rd * ld + ru] = acc;
    }
}

// Spin-exchange and pair-hopping (direct/HxVint.f90:46-98), gather form: for output state i find the source m.
// Signs follow the reference operator order on the full 2Ns-bit word (c, c, cdg, cdg with the O(pos) sign rule
// of ED_SETUP.f90:1080-1106 written as popcounts).
__device__ __forceinline__ double op_sign(uint64_t w, int pos0)     // (-1)^{popcount of bits below pos0}
{
    return (__popcll(w & ((1ull << pos0) - 1ull)) & 1) ? -1.0 : 1.0;
}

__global__ void __launch_bounds__(256)
k_hxv_jxjp(int ns, int norb, int64_t dim_up, int64_t dim_dw, int64_t ld, double jx, double jp,
           const uint32_t *__restrict__ cfg_up, const uint32_t *__restrict__ cfg_dw,
           const uint32_t *__restrict__ rank_up, const uint32_t *__restrict__ rank_dw,
           const double *__restrict__ x, double *__restrict__ y)
{
    const int64_t ru = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ru >= dim_up) return;
    const uint32_t u = cfg_up[ru];
    for (int64_t rd = blockIdx.y; rd < dim_dw; rd += gridDim.y) {
        const uint32_t d = cfg_dw[rd];
        double acc = 0.0;
        for (int a = 0; a < norb; a++)            // a = iorb-1, b = jorb-1 of the reference loops
            for (int b = 0; b < norb; b++) {
                if (a == b) continue;
                const uint32_t ua = (u >> a) & 1u, ub = (u >> b) & 1u, da = (d >> a) & 1u, db = (d >> b) & 1u;
                // spin exchange: source m has (b_up=1, a_dw=1, b_dw=0, a_up=0); output i = c+_{a,up} c+_{b,dw} c_{a,dw} c_{b,up} m
                if (jx != 0.0 && ua == 1 && ub == 0 && da == 0 && db == 1) {
                    uint32_t mu = (u & ~(1u << a)) | (1u << b), md = (d & ~(1u << b)) | (1u << a);
                    uint64_t w = (uint64_t)mu | ((uint64_t)md << ns);
                    double sg = op_sign(w, b); w &= ~(1ull << b);                 // c(jorb)
                    sg *= op_sign(w, a + ns); w &= ~(1ull << (a + ns));           // c(iorb+Ns)
                    sg *= op_sign(w, b + ns); w |= (1ull << (b + ns));            // cdg(jorb+Ns)
                    sg *= op_sign(w, a);                                          // cdg(iorb)
                    acc += jx * sg * x[(int64_t)rank_dw[md] * ld + rank_up[mu]];
                }
                // pair hopping: source m has (b_up=1, b_dw=1, a_dw=0, a_up=0); output i = c+_{a,up} c+_{a,dw} c_{b,dw} c_{b,up} m
                if (jp != 0.0 && ua == 1 && da == 1 && ub == 0 && db == 0) {
                    uint32_t mu = (u & ~(1u << a)) | (1u << b), md = (d & ~(1u << a)) | (1u << b);
                    uint64_t w = (uint64_t)mu | ((uint64_t)md << ns);
                    double sg = op_sign(w, b); w &= ~(1ull << b);                 // c(jorb)
                    sg *= op_sign(w, b + ns); w &= ~(1ull << (b + ns));           // c(jorb+Ns)
                    sg *= op_sign(w, a + ns); w |= (1ull << (a + ns));            // cdg(iorb+Ns)
                    sg *= op_sign(w, a);                                          // cdg(iorb)
                    acc += jp * sg * x[(int64_t)rank_dw[md] * ld + rank_up[mu]];
                }
            }
        if (acc != 0.0) y[rd * ld + ru] += acc;
    }
}

int hxv_jxjp(edgpu_sector *s, const double *x, double *y)
{
    edgpu_ctx *ctx = s->ctx;
    if (!ctx->ham.jhflag) return 0;
    dim3 block(256), grid((unsigned)((s->dim_up + 255) / 256), (unsigned)(s->dim_dw < 32768 ? s->dim_dw : 32768));
    k_hxv_jxjp<<<grid, block, 0, ctx->stream>>>(ctx->ham.ns, ctx->ham.norb, s->dim_up, s->dim_dw, s->ld,
                                               ctx->ham.jx, ctx->ham.jp, s->up->cfg, s->dw->cfg,
                                               s->up->rank, s->dw->rank, x, y);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

int hxv_generic(edgpu_sector *s, const double *x, double *y)
{
    edgpu_ctx *ctx = s->ctx;
    dim3 block(256), grid((unsigned)((s->dim_up + 255) / 256), (unsigned)(s->dim_dw < 32768 ? s->dim_dw : 32768));
    k_hxv_generic<<<grid, block, 0, ctx->stream>>>(s->dim_up, s->dim_dw, s->ld, ctx->ham.norb,
                                                  s->up->cfg, s->dw->cfg, s->up->ediag, s->dw->ediag, ctx->d_xtab,
                                                  s->up->hop, s->up->nhop, s->up->amp,
                                                  s->dw->hop, s->dw->nhop, s->dw->amp, x, y);
    CUDA_TRY(ctx, cudaGetLastError());
    return hxv_jxjp(s, x, y);
}

// ------------------------------------------------------------------------------------------------------------
// Dense Hmat of a small sector in ONE launch (build_Hv_sector(isector,Hmat), ED_HAMILTONIAN.f90:75-79; used for the
// LAPACK sectors of ED_DIAG.f90:188-193).  One thread per reference row i = ru + rd*DimUp: the row of H is sparse, its
// entries are the diagonal and the gather-form hop terms of the two spins.  Column-major output in the reference order.
// (Replaces dim products H e_j with a synchronisation each: the sector scan of ed_solve was bound by those.)
__global__ void k_invert_perm(int64_t n, const uint32_t *__restrict__ r2i, uint32_t *__restrict__ i2r)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) i2r[r2i ? r2i[r] : r] = (uint32_t)r;
}

__global__ void __launch_bounds__(256)
k_dense_rows(int64_t dim_up, int64_t dim_dw, int norb,
             const uint32_t *__restrict__ r2i_up, const uint32_t *__restrict__ r2i_dw,
             const uint32_t *__restrict__ i2r_up, const uint32_t *__restrict__ i2r_dw,
             const uint32_t *__restrict__ cfg_up, const uint32_t *__restrict__ cfg_dw,
             const double *__restrict__ e_up, const double *__restrict__ e_dw, const double *__restrict__ xtab,
             const uint32_t *__restrict__ hop_up, const uint8_t *__restrict__ nhop_up, const double *__restrict__ amp_up,
             const uint32_t *__restrict__ hop_dw, const uint8_t *__restrict__ nhop_dw, const double *__restrict__ amp_dw,
             double *__restrict__ H)
{
    const int64_t dim = dim_up * dim_dw;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    const int64_t ru = i % dim_up, rd = i / dim_up;
    const int64_t iu = r2i_up ? r2i_up[ru] : ru, id = r2i_dw ? r2i_dw[rd] : rd;
    const uint32_t impmask = (1u << norb) - 1u;
    H[i + dim * i] += e_up[iu] + e_dw[id] + xtab[(cfg_dw[id] & impmask) * 32u + (cfg_up[iu] & impmask)];
    const int nu = nhop_up[iu];
    for (int j = 0; j < nu; j++) {
        const uint32_t h = hop_up[(int64_t)j * dim_up + iu];
        const int64_t col = (int64_t)i2r_up[h >> 8] + rd * dim_up;
        H[i + dim * col] += amp_up[h & 255u];
    }
    const int nd = nhop_dw[id];
    for (int j = 0; j < nd; j++) {
        const uint32_t h = hop_dw[(int64_t)j * dim_dw + id];
        const int64_t col = ru + (int64_t)i2r_dw[h >> 8] * dim_up;
        H[i + dim * col] += amp_dw[h & 255u];
    }
}

// d_H: dim x dim device buffer (zeroed here).  Returns 1 (without an error message) when the sector needs the
// two-spin Jx/Jp terms: the caller falls back to products with unit vectors.
int dense_rows(edgpu_sector *s, double *d_H)
{
    edgpu_ctx *ctx = s->ctx;
    if (ctx->ham.jhflag) return 1;
    const int64_t dim = s->dim;
    uint32_t *i2r = nullptr;
    if (int rc = pool_alloc(ctx, sizeof(uint32_t) * (size_t)(s->dim_up + s->dim_dw), (void **)&i2r)) return rc;
    cudaStream_t st = ctx->stream;
    k_invert_perm<<<(unsigned)((s->dim_up + 255) / 256), 256, 0, st>>>(s->dim_up, s->up->ref2int, i2r);
    k_invert_perm<<<(unsigned)((s->dim_dw + 255) / 256), 256, 0, st>>>(s->dim_dw, s->dw->ref2int, i2r + s->dim_up);
    CUDA_TRY(ctx, cudaMemsetAsync(d_H, 0, sizeof(double) * (size_t)dim * (size_t)dim, st));
    k_dense_rows<<<(unsigned)((dim + 255) / 256), 256, 0, st>>>(s->dim_up, s->dim_dw, ctx->ham.norb, s->up->ref2int, s->dw->ref2int, i2r, i2r + s->dim_up,
                                                                s->up->cfg, s->dw->cfg, s->up->ediag, s->dw->ediag, ctx->d_xtab,
                                                                s->up->hop, s->up->nhop, s->up->amp, s->dw->hop, s->dw->nhop, s->dw->amp, d_H);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    pool_release(ctx, i2r);
    return 0;
}
