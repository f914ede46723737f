// ops.cu -- sector-vector utilities: layout import/export, Philox fill, GF seed operators, observables.
//
// apply_c replaces the seed loops of lanc_build_gf_normal_c (ED_GF_NORMAL.f90:159-174 for c^+, :212-227 for c);
// observables replaces the normal-mode core of observables_impurity (ED_OBSERVABLES.f90:127-158).
#include "edgpu_internal.h"
#include <algorithm>
#include <cmath>
#include <vector>

int vec_scale(edgpu_ctx *ctx, double *v, double alpha, int64_t n);

// ---- layout conversion: reference order (i = ru + rd*DimUp, colex ranks) <-> internal order ---------------
// mode 0: internal <- ref (real source) ; 1: internal <- real part of interleaved complex source
// mode 2: ref (real) <- internal        ; 3: interleaved complex (imag=0) <- internal
// mode 4: internal <- imag part of complex source ; 5: complex.imag <- internal (real part untouched)
__global__ void __launch_bounds__(256)
k_convert(int mode, int64_t dim_up, int64_t rd_begin, int64_t rd_end, VAddr va,
          const uint32_t *__restrict__ r2i_up, const uint32_t *__restrict__ r2i_dw,
          const double *__restrict__ src, double *__restrict__ dst, double *__restrict__ flag)
{
    // reference rows [rd_begin, rd_end); the reference-side array starts at row rd_begin (chunked transfers)
    const int64_t ru = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ru >= dim_up) return;
    const int64_t iu = r2i_up ? (int64_t)r2i_up[ru] : ru;
    for (int64_t rd = rd_begin + blockIdx.y; rd < rd_end; rd += gridDim.y) {
        const int64_t id = r2i_dw ? (int64_t)r2i_dw[rd] : rd;
        const int64_t iref = (rd - rd_begin) * dim_up + ru, iint = va(id, iu);
        // iint < 0: the element belongs to a pair tile of another rank (sharded vectors): nothing to import, zero on export
        switch (mode) {
        case 0: if (iint >= 0) dst[iint] = src[iref]; break;
        case 1:
            if (iint >= 0) dst[iint] = src[2 * iref];
            if (flag && src[2 * iref + 1] != 0.0) *flag = 1.0;       // this path is real (ed_mode=normal): reported by the caller
            break;
        case 2: dst[iref] = iint >= 0 ? src[iint] : 0.0; break;
        case 3: dst[2 * iref] = iint >= 0 ? src[iint] : 0.0; dst[2 * iref + 1] = 0.0; break;
        case 4: if (iint >= 0) dst[iint] = src[2 * iref + 1]; break;
        case 5: dst[2 * iref + 1] = iint >= 0 ? src[iint] : 0.0; break;
        }
    }
}

static dim3 grid2d(const edgpu_sector *s) {
    return dim3((unsigned)((s->dim_up + 255) / 256), (unsigned)(s->dim_dw < 32768 ? s->dim_dw : 32768));
}

int vec_convert(edgpu_sector *s, int mode, const double *src, double *dst)
{
    k_convert<<<grid2d(s), 256, 0, s->ctx->stream>>>(mode, s->dim_up, 0, s->dim_dw, sector_vaddr(s), s->up->ref2int, s->dw->ref2int, src, dst, nullptr);
    CUDA_TRY(s->ctx, cudaGetLastError());
    return 0;
}

// the same for reference rows [rd0, rd1) only; the reference-side array is the chunk holding just those rows
int vec_convert_rows(edgpu_sector *s, int mode, int64_t rd0, int64_t rd1, const double *src, double *dst, cudaStream_t st, double *flag)
{
    if (rd1 <= rd0) return 0;
    dim3 grid((unsigned)((s->dim_up + 255) / 256), (unsigned)std::min<int64_t>(rd1 - rd0, 32768));
    k_convert<<<grid, 256, 0, st>>>(mode, s->dim_up, rd0, rd1, sector_vaddr(s), s->up->ref2int, s->dw->ref2int, src, dst, flag);
    CUDA_TRY(s->ctx, cudaGetLastError());
    return 0;
}
int vec_import_ref(edgpu_sector *s, const double *d_ref, double *d_int) { return vec_convert(s, 0, d_ref, d_int); }
int vec_export_ref(edgpu_sector *s, const double *d_int, double *d_ref) { return vec_convert(s, 2, d_int, d_ref); }

// ---- Philox4x32-10 N(0,1): element counter = reference index, so the vector does not depend on the layout --
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__global__ void __launch_bounds__(256)
k_fill_random(int uniform, uint64_t seed, int64_t dim_up, int64_t dim_dw, VAddr va,
              const uint32_t *__restrict__ r2i_up, const uint32_t *__restrict__ r2i_dw, double *__restrict__ dst)
{
    const int64_t ru = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ru >= dim_up) return;
    const int64_t iu = r2i_up ? (int64_t)r2i_up[ru] : ru;
    for (int64_t rd = blockIdx.y; rd < dim_dw; rd += gridDim.y) {
        const int64_t id = r2i_dw ? (int64_t)r2i_dw[rd] : rd;
        const uint64_t idx = (uint64_t)(rd * dim_up + ru);
        const int64_t ia = va(id, iu);
        if (ia < 0) continue;                                 // pair tile of another rank
        uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint64_t a = ((uint64_t)c[0] << 21) ^ (uint64_t)(c[1] >> 11);
        const uint64_t b = ((uint64_t)c[2] << 21) ^ (uint64_t)(c[3] >> 11);
        const double u1 = ((double)a + 0.5) * (1.0 / 9007199254740992.0);
        const double u2 = ((double)b + 0.5) * (1.0 / 9007199254740992.0);
        if (uniform) {
            // (a52 + 0.5)/2^51 - 1 in (-1,1): every step is exact in fp64 => bit-identical to the host generator
            const uint64_t a52 = ((uint64_t)c[0] << 20) ^ (uint64_t)(c[1] >> 12);
            dst[ia] = ((double)a52 + 0.5) * (1.0 / 2251799813685248.0) - 1.0;
        } else {
            dst[ia] = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
        }
    }
}

int vec_fill_random(edgpu_sector *s, int uniform, uint64_t seed, double *dst)
{
    k_fill_random<<<grid2d(s), 256, 0, s->ctx->stream>>>(uniform, seed, s->dim_up, s->dim_dw, sector_vaddr(s), s->up->ref2int, s->dw->ref2int, dst);
    CUDA_TRY(s->ctx, cudaGetLastError());
    return 0;
}

// ---- c / c^+ between neighbouring sectors, gather form on the OUTPUT sector -------------------------------
// sign rule of ED_SETUP.f90:1080-1106 on the full word: an up operator at bit a sees the up bits below a; a down
// operator additionally passes all n_up up bits.  Output is written everywhere (zero where the operator kills).
__global__ void __launch_bounds__(256)
k_apply_c(int bit, int is_dw, int nup_in, int64_t dim_up_o, int64_t dim_dw_o, VAddr va_o, VAddr va_i,
          const uint32_t *__restrict__ cfg_up_o, const uint32_t *__restrict__ cfg_dw_o,
          const uint32_t *__restrict__ rank_up_i, const uint32_t *__restrict__ rank_dw_i,
          const double *__restrict__ in, double *__restrict__ out)
{
    const int64_t ru = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ru >= dim_up_o) return;
    const uint32_t u = cfg_up_o[ru];
    const uint32_t m = 1u << bit;
    for (int64_t rd = blockIdx.y; rd < dim_dw_o; rd += gridDim.y) {
        const uint32_t d = cfg_dw_o[rd];
        uint32_t us = u, ds = d;
        if (is_dw) ds ^= m; else us ^= m;                     // source word = output word with the bit toggled back
        const uint32_t iu = rank_up_i[us], id = rank_dw_i[ds];
        double val = 0.0;
        if (iu != 0xFFFFFFFFu && id != 0xFFFFFFFFu) {         // popcount of the source matches the input sector
            int par = is_dw ? (nup_in + __popc(ds & (m - 1u))) : __popc(us & (m - 1u));
            const double vin = in[va_i((int64_t)id, (int64_t)iu)];
            val = (par & 1) ? -vin : vin;
        }
        out[va_o(rd, ru)] = val;
    }
}

extern "C" int edgpu_apply_c(edgpu_sector *si, edgpu_sector *so, int32_t isite, int32_t dagger,
                             const edgpu_vec *in, edgpu_vec *out, int32_t normalise, double *norm2)
{
    if (!si || !so || !in || !out || in->s != si || out->s != so) return si ? edgpu_fail(si->ctx, "edgpu_apply_c: bad handles") : 1;
    edgpu_ctx *ctx = si->ctx;
    const int ns = ctx->ham.ns;
    if (isite < 1 || isite > 2 * ns) return edgpu_fail(ctx, "edgpu_apply_c: isite=%d out of range", isite);
    const int is_dw = isite > ns, bit = is_dw ? isite - 1 - ns : isite - 1;
    const int dn = dagger ? 1 : -1;
    if (so->nup != si->nup + (is_dw ? 0 : dn) || so->ndw != si->ndw + (is_dw ? dn : 0))
        return edgpu_fail(ctx, "edgpu_apply_c: output sector (%d,%d) is not %s_%d applied to (%d,%d)", so->nup, so->ndw,
                          dagger ? "cdg" : "c", isite, si->nup, si->ndw);
    dim3 grid((unsigned)((so->dim_up + 255) / 256), (unsigned)(so->dim_dw < 32768 ? so->dim_dw : 32768));
    // The toggled bit must be SET in the output for cdg and CLEAR for c: the rank LUT of the input sector rejects
    // sources with the wrong popcount, which is exactly that condition.
    if (si->shard_nranks > 1 || so->shard_nranks > 1) return edgpu_fail(ctx, "edgpu_apply_c: not available on sharded sectors (the seed is built on one rank)");
    k_apply_c<<<grid, 256, 0, ctx->stream>>>(bit, is_dw, si->nup, so->dim_up, so->dim_dw, sector_vaddr(so), sector_vaddr(si),
                                            so->up->cfg, so->dw->cfg, si->up->rank, si->dw->rank, in->d, out->d);
    CUDA_TRY(ctx, cudaGetLastError());
    if (int rc = vec_dot(ctx, out->d, out->d, so->nalloc, ctx->d_scal)) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const double n2 = ctx->h_scal[0];
    if (norm2) *norm2 = n2;
    if (normalise && n2 > 0.0) return vec_scale(ctx, out->d, 1.0 / std::sqrt(n2), so->nalloc);
    return 0;
}

// ---- diagonal seed operators of the susceptibility chains (ED_GF_CHISPIN.f90:93-100, 198-205) ----------------------
// out = 1/2 (n_up - n_dw) in   over the impurity levels selected by `mask` (one orbital, or all of them for S_z^tot)
// charge seeds (ED_GF_CHIDENS.f90:126-133, 227-234): out = (n_up + n_dw) in            (cdw = +1, scale = 1)
__global__ void __launch_bounds__(256)
k_apply_sz(uint32_t mask, double cdw, double scale, int64_t dim_up, int64_t dim_dw, VAddr va, const uint32_t *__restrict__ cfg_up,
           const uint32_t *__restrict__ cfg_dw, const double *__restrict__ in, double *__restrict__ out)
{
    const int64_t ru = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ru >= dim_up) return;
    const int nu = __popc(cfg_up[ru] & mask);
    for (int64_t rd = blockIdx.y; rd < dim_dw; rd += gridDim.y) {
        const double sgn = (double)nu + cdw * (double)__popc(cfg_dw[rd] & mask);
        const int64_t a = va(rd, ru);
        if (a >= 0) out[a] = scale * sgn * in[a];
    }
}

static int apply_diag_seed(edgpu_sector *s, int32_t iorb, double cdw, double scale, const edgpu_vec *in, edgpu_vec *out,
                           int32_t normalise, double *norm);

extern "C" int edgpu_apply_sz(edgpu_sector *s, int32_t iorb, const edgpu_vec *in, edgpu_vec *out, int32_t normalise, double *norm)
{
    return apply_diag_seed(s, iorb, -1.0, 0.5, in, out, normalise, norm);
}

extern "C" int edgpu_apply_n(edgpu_sector *s, int32_t iorb, const edgpu_vec *in, edgpu_vec *out, int32_t normalise, double *norm)
{
    return apply_diag_seed(s, iorb, 1.0, 1.0, in, out, normalise, norm);
}

static int apply_diag_seed(edgpu_sector *s, int32_t iorb, double cdw, double scale, const edgpu_vec *in, edgpu_vec *out,
                           int32_t normalise, double *norm)
{
    if (!s || !in || !out || in->s != s || out->s != s) return s ? edgpu_fail(s->ctx, "edgpu_apply_sz: bad handles") : 1;
    edgpu_ctx *ctx = s->ctx;
    const int norb = ctx->ham.norb;
    if (iorb < 0 || iorb > norb) return edgpu_fail(ctx, "edgpu_apply_sz: iorb=%d out of range (0 = total, 1..Norb)", iorb);
    if (in->d == out->d) return edgpu_fail(ctx, "edgpu_apply_sz: in-place application is not allowed");
    const uint32_t mask = iorb == 0 ? (1u << norb) - 1u : 1u << (iorb - 1);
    dim3 grid((unsigned)((s->dim_up + 255) / 256), (unsigned)(s->dim_dw < 32768 ? s->dim_dw : 32768));
    k_apply_sz<<<grid, 256, 0, ctx->stream>>>(mask, cdw, scale, s->dim_up, s->dim_dw, sector_vaddr(s), s->up->cfg, s->dw->cfg, in->d, out->d);
    CUDA_TRY(ctx, cudaGetLastError());
    if (int rc = vec_dot(ctx, out->d, out->d, s->nalloc, ctx->d_scal)) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const double n2 = ctx->h_scal[0];
    if (norm) *norm = std::sqrt(n2);
    if (normalise && n2 > 0.0) return vec_scale(ctx, out->d, 1.0 / std::sqrt(n2), s->nalloc);
    return 0;
}

// ---- observables: joint distribution of the impurity bits -------------------------------------------------
// rowsum[rd][ui] = sum_{ru : imp(u)=ui} gs[rd][ru]^2 ; the host combines rows by imp(d) (fixed order => deterministic).
__global__ void __launch_bounds__(256)
k_obs_rows(int norb, int64_t dim_up, int64_t dim_dw, VAddr va, const uint32_t *__restrict__ cfg_up,
           const double *__restrict__ gs, double *__restrict__ rowsum)
{
    extern __shared__ double sh[];                      // [nimp][256]
    const int nimp = 1 << norb;
    const uint32_t mask = nimp - 1;
    for (int64_t rd = blockIdx.x; rd < dim_dw; rd += gridDim.x) {
        for (int k = 0; k < nimp; k++) sh[k * 256 + threadIdx.x] = 0.0;
        for (int64_t ru = threadIdx.x; ru < dim_up; ru += 256) {
            const int64_t a = va(rd, ru);
            const double g = a >= 0 ? gs[a] : 0.0;
            sh[(cfg_up[ru] & mask) * 256 + threadIdx.x] += g * g;
        }
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o)
                for (int k = 0; k < nimp; k++) sh[k * 256 + threadIdx.x] += sh[k * 256 + threadIdx.x + o];
            __syncthreads();
        }
        if ((int)threadIdx.x < nimp) rowsum[rd * nimp + threadIdx.x] = sh[threadIdx.x * 256];
        __syncthreads();
    }
}

extern "C" int edgpu_observables(edgpu_sector *s, const edgpu_vec *gs, double peso,
                                 double *dens, double *dens_up, double *dens_dw, double *docc, double *magz,
                                 double *sz2, double *n2, double *s2tot)
{
    if (!s || !gs || gs->s != s) return s ? edgpu_fail(s->ctx, "edgpu_observables: bad handles") : 1;
    edgpu_ctx *ctx = s->ctx;
    const int norb = ctx->ham.norb, nimp = 1 << norb;
    double *d_rows = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_rows, sizeof(double) * (size_t)s->dim_dw * nimp));
    const unsigned nb = (unsigned)(s->dim_dw < 4096 ? s->dim_dw : 4096);
    // [nimp][256] doubles of dynamic shared memory: 64 KB for Norb = 5, above the 48 KB a kernel gets without opting in
    if (sizeof(double) * 256 * nimp > 48 * 1024)
        CUDA_TRY(ctx, cudaFuncSetAttribute((const void *)k_obs_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 256 * nimp)));
    k_obs_rows<<<nb, 256, sizeof(double) * 256 * nimp, ctx->stream>>>(norb, s->dim_up, s->dim_dw, sector_vaddr(s), s->up->cfg, gs->d, d_rows);
    CUDA_TRY(ctx, cudaGetLastError());
    std::vector<double> rows((size_t)s->dim_dw * nimp);
    std::vector<uint32_t> cfgd((size_t)s->dim_dw);
    CUDA_TRY(ctx, cudaMemcpyAsync(rows.data(), d_rows, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(cfgd.data(), s->dw->cfg, sizeof(uint32_t) * cfgd.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_rows);
    std::vector<double> P((size_t)nimp * nimp, 0.0);                     // P[di][ui]
    for (int64_t rd = 0; rd < s->dim_dw; rd++) {
        const int di = (int)(cfgd[rd] & (uint32_t)(nimp - 1));
        for (int ui = 0; ui < nimp; ui++) P[(size_t)di * nimp + ui] += rows[(size_t)rd * nimp + ui];
    }
    // ED_OBSERVABLES.f90:134-157 evaluated on the joint distribution
    for (int di = 0; di < nimp; di++)
        for (int ui = 0; ui < nimp; ui++) {
            const double w = peso * P[(size_t)di * nimp + ui];
            if (w == 0.0) continue;
            double nu[EDGPU_MAXORB], nd[EDGPU_MAXORB], sz[EDGPU_MAXORB], nt[EDGPU_MAXORB], ssz = 0.0;
            for (int a = 0; a < norb; a++) {
                nu[a] = (ui >> a) & 1; nd[a] = (di >> a) & 1;
                sz[a] = (nu[a] - nd[a]) / 2.0; nt[a] = nu[a] + nd[a]; ssz += sz[a];
            }
            for (int a = 0; a < norb; a++) {
                if (dens) dens[a] += nt[a] * w;
                if (dens_up) dens_up[a] += nu[a] * w;
                if (dens_dw) dens_dw[a] += nd[a] * w;
                if (docc) docc[a] += nu[a] * nd[a] * w;
                if (magz) magz[a] += (nu[a] - nd[a]) * w;
                for (int b = 0; b < norb; b++) {
                    if (sz2) sz2[a + norb * b] += sz[a] * sz[b] * w;
                    if (n2) n2[a + norb * b] += nt[a] * nt[b] * w;
                }
            }
            if (s2tot) *s2tot += ssz * ssz * w;
        }
    return 0;
}
