// csr.cu -- stored-Hamiltonian path: device-built CSR + SpMV with 128-bit loads.
//
// Replaces ed_buildH_c (ED_HAMILTONIAN_STORED_HxV.f90:28-113 + ED_HAMILTONIAN/stored/{Himp,Hint,Hbath,
// Himp_bath}.f90), the row-list sparse matrix of ED_SPARSE_MATRIX.f90 (sp_insert_element :249-280) and
// spMatVec_cc (ED_HAMILTONIAN_STORED_HxV.f90:132-143).
// Row i keeps the reference's insertion order: one diagonal entry (the three diagonal inserts accumulate into
// it, ED_SPARSE_MATRIX.f90:268-269), then Himp off-diagonal, spin-exchange, pair-hopping, Himp_bath entries.
// Storage: rows start at multiples of 4 entries (pads: col = own row, val = 0) so that a 4-lane group reads
// one uint4 of columns and two double2 of values per lane -- every global load is 128 bits wide and aligned.
#include "edgpu_internal.h"
#include <cub/device/device_scan.cuh>
#include <vector>

struct CsrParams {
    int ns, norb, nbath, nspin_dw;      // nspin_dw: parameter spin index used by the DW species
    double hloc[2][EDGPU_MAXORB][EDGPU_MAXORB];
    double v[2][EDGPU_MAXORB][32];
    double jx, jp;
    int jhflag;
};

__device__ __forceinline__ double sgn_below(uint64_t w, int bit)
{
    return (__popcll(w & ((1ull << bit) - 1ull)) & 1) ? -1.0 : 1.0;
}

// Enumerate the off-diagonal entries of the row whose state is (u,d), in the reference's insertion order.
// F is called as f(target_u, target_d, value).
template <class F>
__device__ __forceinline__ void csr_row_offdiag(const CsrParams &P, uint32_t u, uint32_t d, F f)
{
    const int ns = P.ns;
    const uint64_t m = (uint64_t)u | ((uint64_t)d << ns);
    // stored/Himp.f90:26-70
    for (int io = 0; io < P.norb; io++)
        for (int jo = 0; jo < P.norb; jo++) {
            for (int sp = 0; sp < 2; sp++) {
                const double h = P.hloc[sp][io][jo];
                const int off = sp ? ns : 0;
                if (h != 0.0 && ((m >> (jo + off)) & 1ull) && !((m >> (io + off)) & 1ull)) {
                    double sg = sgn_below(m, jo + off);
                    uint64_t k1 = m & ~(1ull << (jo + off));
                    sg *= sgn_below(k1, io + off);
                    uint64_t k2 = k1 | (1ull << (io + off));
                    f((uint32_t)(k2 & ((1ull << ns) - 1ull)), (uint32_t)(k2 >> ns), h * sg);
                }
            }
        }
    // stored/Hint.f90:60-118
    if (P.norb > 1 && P.jhflag) {
        for (int io = 0; io < P.norb; io++)
            for (int jo = 0; jo < P.norb; jo++)
                if (io != jo && ((m >> jo) & 1ull) && ((m >> (io + ns)) & 1ull) && !((m >> (jo + ns)) & 1ull) && !((m >> io) & 1ull)) {
                    uint64_t w = m;
                    double sg = sgn_below(w, jo); w &= ~(1ull << jo);
                    sg *= sgn_below(w, io + ns); w &= ~(1ull << (io + ns));
                    sg *= sgn_below(w, jo + ns); w |= (1ull << (jo + ns));
                    sg *= sgn_below(w, io); w |= (1ull << io);
                    f((uint32_t)(w & ((1ull << ns) - 1ull)), (uint32_t)(w >> ns), P.jx * sg);
                }
        for (int io = 0; io < P.norb; io++)
            for (int jo = 0; jo < P.norb; jo++)
                if (io != jo && ((m >> jo) & 1ull) && ((m >> (jo + ns)) & 1ull) && !((m >> (io + ns)) & 1ull) && !((m >> io) & 1ull)) {
                    uint64_t w = m;
                    double sg = sgn_below(w, jo); w &= ~(1ull << jo);
                    sg *= sgn_below(w, jo + ns); w &= ~(1ull << (jo + ns));
                    sg *= sgn_below(w, io + ns); w |= (1ull << (io + ns));
                    sg *= sgn_below(w, io); w |= (1ull << io);
                    f((uint32_t)(w & ((1ull << ns) - 1ull)), (uint32_t)(w >> ns), P.jp * sg);
                }
    }
    // stored/Himp_bath.f90:10-70
    for (int io = 0; io < P.norb; io++)
        for (int kp = 0; kp < P.nbath; kp++) {
            const int ms = P.norb + io * P.nbath + kp;          // getBathStride-1
            for (int sp = 0; sp < 2; sp++) {
                const double v = P.v[sp][io][kp];
                const int off = sp ? ns : 0;
                if (v == 0.0) continue;
                const int bi = (int)((m >> (io + off)) & 1ull), bb = (int)((m >> (ms + off)) & 1ull);
                if (bi == 1 && bb == 0) {                       // c(iorb) ; cdg(ms)
                    double sg = sgn_below(m, io + off);
                    uint64_t k1 = m & ~(1ull << (io + off));
                    sg *= sgn_below(k1, ms + off);
                    uint64_t k2 = k1 | (1ull << (ms + off));
                    f((uint32_t)(k2 & ((1ull << ns) - 1ull)), (uint32_t)(k2 >> ns), v * sg);
                }
                if (bi == 0 && bb == 1) {                       // c(ms) ; cdg(iorb)
                    double sg = sgn_below(m, ms + off);
                    uint64_t k1 = m & ~(1ull << (ms + off));
                    sg *= sgn_below(k1, io + off);
                    uint64_t k2 = k1 | (1ull << (io + off));
                    f((uint32_t)(k2 & ((1ull << ns) - 1ull)), (uint32_t)(k2 >> ns), v * sg);
                }
            }
        }
}

__global__ void __launch_bounds__(256)
k_csr_count(CsrParams P, int64_t dim_up, int64_t dim_dw, int64_t ld,
            const uint32_t *__restrict__ cfg_up, const uint32_t *__restrict__ cfg_dw,
            int64_t *__restrict__ len4, uint8_t *__restrict__ rowlen, unsigned long long *__restrict__ total)
{
    unsigned long long mine = 0;
    const int64_t ru = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ru >= ld) return;
    for (int64_t rd = blockIdx.y; rd < dim_dw; rd += gridDim.y) {
        int n = 0;
        if (ru < dim_up) {
            n = 1;
            csr_row_offdiag(P, cfg_up[ru], cfg_dw[rd], [&](uint32_t, uint32_t, double) { n++; });
        }
        rowlen[rd * ld + ru] = (uint8_t)n;
        len4[rd * ld + ru] = (n + 3) / 4 * 4;
        mine += (unsigned long long)n;
    }
    if (mine) atomicAdd(total, mine);
}

__global__ void __launch_bounds__(256)
k_csr_fill(CsrParams P, int64_t dim_up, int64_t dim_dw, int64_t ld,
           const uint32_t *__restrict__ cfg_up, const uint32_t *__restrict__ cfg_dw,
           const uint32_t *__restrict__ rank_up, const uint32_t *__restrict__ rank_dw,
           const double *__restrict__ e_up, const double *__restrict__ e_dw, const double *__restrict__ xtab,
           const int64_t *__restrict__ rowptr, uint32_t *__restrict__ cols, double *__restrict__ vals)
{
    const int64_t ru = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ru >= dim_up) return;
    const uint32_t u = cfg_up[ru];
    const uint32_t impmask = (1u << P.norb) - 1u;
    for (int64_t rd = blockIdx.y; rd < dim_dw; rd += gridDim.y) {
        const uint32_t d = cfg_dw[rd];
        const int64_t row = rd * ld + ru;
        int64_t p = rowptr[row];
        const int64_t pend = rowptr[row + 1];
        cols[p] = (uint32_t)row;
        vals[p] = e_up[ru] + e_dw[rd] + xtab[(d & impmask) * 32 + (u & impmask)];
        p++;
        csr_row_offdiag(P, u, d, [&](uint32_t tu, uint32_t td, double val) {
            cols[p] = (uint32_t)((int64_t)rank_dw[td] * ld + rank_up[tu]);
            vals[p] = val;
            p++;
        });
        for (; p < pend; p++) { cols[p] = (uint32_t)row; vals[p] = 0.0; }
    }
}

// SpMV: 4 lanes per row; each lane loads 4 column indices (uint4) and 4 values (2 x double2) per step.
__global__ void __launch_bounds__(256)
k_spmv(int64_t nrows, const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ cols,
       const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y)
{
    const int sub = threadIdx.x & 3;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2; row < nrows;
         row += ((int64_t)gridDim.x * blockDim.x) >> 2) {
        const int64_t p0 = rowptr[row], p1 = rowptr[row + 1];
        double acc = 0.0;
        for (int64_t p = p0 + 4 * sub; p < p1; p += 16) {
            const uint4 c = *reinterpret_cast<const uint4 *>(cols + p);
            const double2 v01 = *reinterpret_cast<const double2 *>(vals + p);
            const double2 v23 = *reinterpret_cast<const double2 *>(vals + p + 2);
            acc += v01.x * x[c.x];
            acc += v01.y * x[c.y];
            acc += v23.x * x[c.z];
            acc += v23.y * x[c.w];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (sub == 0 && p1 > p0) y[row] = acc;
    }
}

static void fill_params(const edgpu_ctx *ctx, CsrParams &P)
{
    const HamParams &h = ctx->ham;
    memset(&P, 0, sizeof(P));
    P.ns = h.ns; P.norb = h.norb; P.nbath = h.nbath; P.nspin_dw = h.nspin - 1;
    P.jx = h.jx; P.jp = h.jp; P.jhflag = h.jhflag ? 1 : 0;
    for (int sp = 0; sp < 2; sp++) {
        const int ps = sp ? h.nspin - 1 : 0;
        for (int a = 0; a < h.norb; a++) {
            for (int b = 0; b < h.norb; b++) P.hloc[sp][a][b] = (a == b) ? 0.0 : h.H(ps, a, b);
            for (int k = 0; k < h.nbath && k < 32; k++) P.v[sp][a][k] = h.V(ps, a, k);
        }
    }
}

int csr_build(edgpu_sector *s)
{
    edgpu_ctx *ctx = s->ctx;
    if (ctx->ham.nbath > 32) return edgpu_fail(ctx, "csr_build: Nbath > 32 unsupported");
    if (s->nalloc >= (1ll << 32)) return edgpu_fail(ctx, "csr_build: sector too large for 32-bit column indices");
    auto m = std::unique_ptr<CsrMatrix>(new CsrMatrix());
    m->dim = s->nalloc;
    CsrParams P;
    fill_params(ctx, P);
    int64_t *len4 = nullptr;
    uint8_t *rowlen = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&len4, sizeof(int64_t) * (size_t)(s->nalloc + 1)));
    CUDA_TRY(ctx, cudaMalloc(&rowlen, (size_t)s->nalloc));
    CUDA_TRY(ctx, cudaMalloc(&m->rowptr, sizeof(int64_t) * (size_t)(s->nalloc + 1)));
    CUDA_TRY(ctx, cudaMemsetAsync(len4, 0, sizeof(int64_t) * (size_t)(s->nalloc + 1), ctx->stream));
    unsigned long long *d_total = nullptr, h_total = 0;
    CUDA_TRY(ctx, cudaMalloc(&d_total, sizeof(unsigned long long)));
    CUDA_TRY(ctx, cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), ctx->stream));
    dim3 grid((unsigned)((s->ld + 255) / 256), (unsigned)(s->dim_dw < 32768 ? s->dim_dw : 32768));
    k_csr_count<<<grid, 256, 0, ctx->stream>>>(P, s->dim_up, s->dim_dw, s->ld, s->up->cfg, s->dw->cfg, len4, rowlen, d_total);
    CUDA_TRY(ctx, cudaGetLastError());
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, len4, m->rowptr, s->nalloc + 1, ctx->stream));
    CUDA_TRY(ctx, cudaMalloc(&tmp, tmp_bytes));
    CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, len4, m->rowptr, s->nalloc + 1, ctx->stream));
    int64_t total = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&total, m->rowptr + s->nalloc, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(&h_total, d_total, sizeof(h_total), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    cudaFree(len4);
    cudaFree(d_total);
    m->nnz = total;                                       // padded entry count
    m->nnz_true = (int64_t)h_total;
    m->rowlen = rowlen;
    cudaError_t e1 = cudaMalloc(&m->cols, sizeof(uint32_t) * (size_t)(total + 4));
    cudaError_t e2 = cudaMalloc(&m->vals, sizeof(double) * (size_t)(total + 4));
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        return edgpu_fail(ctx, "csr_build: out of device memory for %lld stored entries", (long long)total);
    }
    dim3 grid2((unsigned)((s->dim_up + 255) / 256), grid.y);
    k_csr_fill<<<grid2, 256, 0, ctx->stream>>>(P, s->dim_up, s->dim_dw, s->ld, s->up->cfg, s->dw->cfg, s->up->rank, s->dw->rank,
                                              s->up->ediag, s->dw->ediag, ctx->d_xtab, m->rowptr, m->cols, m->vals);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    s->csr = std::move(m);
    return 0;
}

int hxv_csr(edgpu_sector *s, const double *x, double *y)
{
    edgpu_ctx *ctx = s->ctx;
    const int64_t threads = s->nalloc * 4;
    int64_t blocks = (threads + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 64;
    if (blocks > cap) blocks = cap;
    k_spmv<<<(unsigned)blocks, 256, 0, ctx->stream>>>(s->nalloc, s->csr->rowptr, s->csr->cols, s->csr->vals, x, y);
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

// Mirror of the reference's spH0 rows: reference row order, 0-based reference column indices, no pads.
int csr_download(const edgpu_sector *s, int64_t *rowptr, int64_t *cols, double *vals)
{
    edgpu_ctx *ctx = s->ctx;
    const CsrMatrix &m = *s->csr;
    std::vector<int64_t> rp((size_t)s->nalloc + 1);
    std::vector<uint32_t> cc((size_t)m.nnz);
    std::vector<double> vv((size_t)m.nnz);
    std::vector<uint8_t> rl((size_t)s->nalloc);
    std::vector<uint32_t> r2iu((size_t)s->dim_up), r2id((size_t)s->dim_dw);
    CUDA_TRY(ctx, cudaMemcpy(rp.data(), m.rowptr, sizeof(int64_t) * rp.size(), cudaMemcpyDeviceToHost));
    CUDA_TRY(ctx, cudaMemcpy(cc.data(), m.cols, sizeof(uint32_t) * cc.size(), cudaMemcpyDeviceToHost));
    CUDA_TRY(ctx, cudaMemcpy(vv.data(), m.vals, sizeof(double) * vv.size(), cudaMemcpyDeviceToHost));
    CUDA_TRY(ctx, cudaMemcpy(rl.data(), m.rowlen, rl.size(), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < s->dim_up; i++) r2iu[i] = (uint32_t)i;
    for (int64_t i = 0; i < s->dim_dw; i++) r2id[i] = (uint32_t)i;
    if (s->up->ref2int) CUDA_TRY(ctx, cudaMemcpy(r2iu.data(), s->up->ref2int, sizeof(uint32_t) * r2iu.size(), cudaMemcpyDeviceToHost));
    if (s->dw->ref2int) CUDA_TRY(ctx, cudaMemcpy(r2id.data(), s->dw->ref2int, sizeof(uint32_t) * r2id.size(), cudaMemcpyDeviceToHost));
    std::vector<int64_t> i2ru((size_t)s->ld, -1), i2rd((size_t)s->dim_dw, -1);
    for (int64_t r = 0; r < s->dim_up; r++) i2ru[r2iu[r]] = r;
    for (int64_t r = 0; r < s->dim_dw; r++) i2rd[r2id[r]] = r;
    int64_t nnz = 0;
    for (int64_t rd = 0; rd < s->dim_dw; rd++)
        for (int64_t ru = 0; ru < s->dim_up; ru++) {
            const int64_t row = (int64_t)r2id[rd] * s->ld + r2iu[ru];
            if (rowptr) rowptr[rd * s->dim_up + ru] = nnz;
            for (int k = 0; k < rl[row]; k++) {
                const int64_t c = cc[rp[row] + k];
                const int64_t cd = c / s->ld, cu = c - cd * s->ld;
                if (cols) cols[nnz] = i2rd[cd] * s->dim_up + i2ru[cu];
                if (vals) vals[nnz] = vv[rp[row] + k];
                nnz++;
            }
        }
    if (rowptr) rowptr[s->dim] = nnz;
    return 0;
}
