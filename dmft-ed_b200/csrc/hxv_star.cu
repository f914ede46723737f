// hxv_star.cu -- star-product layout and tiled on-the-fly H*v kernels (the fast path of the BASELINE configs).
//
// Replaces directMatVec_cc (ED_HAMILTONIAN_DIRECT_HxV.f90:21-92 + ED_HAMILTONIAN/direct/*.f90) when the only
// same-spin hops are impurity a <-> its own bath levels (bath_type=normal, diagonal impHloc: all BASELINE configs).
//
// Structure exploited.  Per spin the Ns levels split into Norb "stars" (orbital a + its Nbath bath levels,
// getBathStride, ED_SETUP.f90:450-454).  Hops never leave a star, so the star occupations (n_0..n_{Norb-1}) are
// conserved: the per-spin hop matrix is block diagonal over occupation tuples, and inside a block it is a
// Kronecker SUM of small star matrices (<= C(Nbath+1,m) <= 70 configurations for Nbath=7) times a sign that
// depends only on the impurity bits of the other stars (c/cdg sign rule, ED_SETUP.f90:1080-1106).
// Device layout ("layout 2"): configurations of one spin are ordered block by block; inside a block the index
// is mixed radix over star indices, star 0 fastest; inside a star: imp=0 configurations (colex over the bath
// word) first, then imp=1.  The sector vector is the Dimdw x ld tile V[r_dw][r_up] in that order (ld = DimUp
// rounded up to 4 doubles so rows are 32-byte aligned).  The reference ordering only exists at the boundary
// (edgpu_vec_upload/download, edgpu_sector_map).
//
// Two kernels per H*v, each closing one spin in shared memory:
//   k_star_dw : one (down-block, column strip) tile.                    y  = H_dw x             R(x) W(y)
//   k_star_up : 4 rows x one up-block = contiguous runs of the rows.    y += (diag + H_up) x    R(x) R(y) W(y)
// Algorithmic bytes per H*v are 2*Dim*8; this two-pass scheme moves 5*Dim*8 through HBM because a
// tile closed under BOTH spins (70^4 doubles at Ns=16) fits no on-chip memory.
#include <mutex>
#include "edgpu_internal.h"
#include <cuda.h>
#include <algorithm>
#include <cstring>
#include <map>

uint64_t edgpu_binom(int n, int k);

#include "star_info.h"

// ------------------------------------------------------------------------------------------------------------
// layout construction
// ------------------------------------------------------------------------------------------------------------
__global__ void k_star_rank(int ns, int n, int norb, int nbath, const uint16_t *__restrict__ srank,
                            const int *__restrict__ Dm, const int *__restrict__ blockoff, uint32_t *__restrict__ rank)
{
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (1u << ns)) return;
    if (__popc(w) != n) { rank[w] = 0xFFFFFFFFu; return; }
    int key = 0, mult = 1, idx = 0, stride = 1;
    for (int a = 0; a < norb; a++) {
        const uint32_t sub = ((w >> a) & 1u) | (((w >> (norb + a * nbath)) & ((1u << nbath) - 1u)) << 1);
        const int m = __popc(sub);
        key += m * mult;
        mult *= (nbath + 2);
        idx += (int)srank[sub] * stride;
        stride *= Dm[m];
    }
    rank[w] = (uint32_t)(blockoff[key] + idx);
}

static int colex_rank_host(uint32_t w)
{
    int r = 0, i = 1;
    for (int p = 0; p < 32; p++)
        if ((w >> p) & 1u) { r += (int)edgpu_binom(p, i); i++; }
    return r;
}

int build_star_layout(edgpu_ctx *ctx, SpinBasis *b, const std::vector<HopPair> &pairs, const std::vector<double> &amps)
{
    const HamParams &h = ctx->ham;
    const int norb = h.norb, nbath = h.nbath, ns = h.ns, n = b->n, ps = b->pspin;
    if (nbath > 9) return edgpu_fail(ctx, "star-product layout supports Nbath <= 9 (got %d)", nbath);
    if (norb > 3) return edgpu_fail(ctx, "star-product kernels support Norb <= 3 (got %d)", norb);
    auto S = std::make_shared<StarInfo>();
    S->norb = norb; S->nbath = nbath; S->H = nbath; S->ncfg = 1 << (nbath + 1);
    int off = 0;
    for (int m = 0; m <= nbath + 1; m++) {
        S->D[m] = (int)edgpu_binom(nbath + 1, m);
        S->A0[m] = (int)edgpu_binom(nbath, m);
        S->coff[m] = off;
        off += S->D[m];
    }
    // star configuration tables: sub-word = imp | bath<<1 ; index inside its occupation class
    std::vector<uint16_t> srank(S->ncfg);
    std::vector<uint32_t> cfg_of(S->ncfg);          // [coff[m] + idx] -> sub-word
    for (uint32_t sub = 0; sub < (uint32_t)S->ncfg; sub++) {
        const int imp = sub & 1u, m = __builtin_popcount(sub);
        const uint32_t bath = sub >> 1;
        const int idx = imp ? S->A0[m] + colex_rank_host(bath) : colex_rank_host(bath);
        srank[sub] = (uint16_t)idx;
        cfg_of[S->coff[m] + idx] = sub;
    }
    std::vector<uint8_t> hopj((size_t)S->ncfg * S->H, 0), hopc(S->ncfg, 0);
    std::vector<int16_t> hopd((size_t)S->ncfg * S->H, 0);
    std::vector<double> hopv((size_t)norb * S->ncfg * S->H, 0.0), estar((size_t)norb * S->ncfg, 0.0);
    // per-level diagonal coefficients of this spin (same regrouping as k_ediag in tables.cu)
    for (int a = 0; a < norb; a++) {
        double cimp = h.H(ps, a, a) - h.xmu;
        if (h.hfmode) {
            cimp -= 0.5 * h.uloc[a];
            if (norb > 1) cimp -= (norb - 1) * (0.5 * h.ust + 0.5 * (h.ust - h.jh));
        }
        for (int m = 0; m <= nbath + 1; m++)
            for (int i = 0; i < S->D[m]; i++) {
                const uint32_t sub = cfg_of[S->coff[m] + i];
                double e = (sub & 1u) ? cimp : 0.0;
                for (int k = 0; k < nbath; k++)
                    if ((sub >> (k + 1)) & 1u) e += h.E(ps, a, k);
                estar[(size_t)a * S->ncfg + S->coff[m] + i] = e;
            }
    }
    S->pair_e = (norb > 1) ? (h.ust - h.jh) : 0.0;
    // hop lists (gather form): config i <- config j that differs by moving one particle between imp and bath k
    for (int m = 0; m <= nbath + 1; m++)
        for (int i = 0; i < S->D[m]; i++) {
            const uint32_t sub = cfg_of[S->coff[m] + i];
            int cnt = 0;
            for (int k = 0; k < nbath; k++) {
                const uint32_t bi = sub & 1u, bk = (sub >> (k + 1)) & 1u;
                if (!(bi ^ bk)) continue;
                const uint32_t sub2 = sub ^ 1u ^ (1u << (k + 1));
                const int neg = __builtin_popcount((sub >> 1) & ((1u << k) - 1u)) & 1;     // bath bits of this star below k
                const size_t e = (size_t)(S->coff[m] + i) * S->H + cnt;
                hopj[e] = (uint8_t)srank[sub2];
                hopd[e] = (int16_t)((int)srank[sub2] - i);
                for (int a = 0; a < norb; a++) {
                    const double v = h.V(ps, a, k);
                    hopv[(size_t)a * S->ncfg * S->H + e] = neg ? -v : v;       // exactly-zero V contributes 0 (skipped in the reference)
                }
                cnt++;
            }
            hopc[S->coff[m] + i] = (uint8_t)cnt;
        }
    (void)pairs; (void)amps;
    // blocks: occupation tuples with sum n, star (norb-1) major
    std::vector<int> blockoff(1, -1);
    int nkeys = 1;
    for (int a = 0; a < norb; a++) nkeys *= (nbath + 2);
    blockoff.assign(nkeys, -1);
    int cur = 0;
    std::vector<int> occ(norb, 0);
    // enumerate tuples in lexicographic order with the last star as the major key
    std::vector<std::vector<int>> tuples;
    {
        std::vector<int> t(norb, 0);
        const int total = nkeys;
        for (int key = 0; key < total; key++) {
            int kk = key, sum = 0;
            for (int a = 0; a < norb; a++) { t[a] = kk % (nbath + 2); kk /= (nbath + 2); sum += t[a]; }
            if (sum == n) tuples.push_back(t);
        }
    }
    for (auto &t : tuples) {
        StarBlock B;
        memset(&B, 0, sizeof(B));
        B.off = cur;
        int size = 1, key = 0, mult = 1, lower = 0;
        for (int a = 0; a < norb; a++) {
            B.n[a] = t[a];
            B.sgn_lower[a] = lower & 1;
            lower += t[a];
            size *= S->D[t[a]];
            key += t[a] * mult;
            mult *= (nbath + 2);
        }
        B.size = size;
        {
            const uint64_t d0 = (uint64_t)S->D[t[0]];
            B.magic0 = (uint32_t)(((1ull << 32) + d0 - 1) / d0);      // d0 == 1 gives 2^32 -> 0: handled in the kernel
            B.ny = 512 / (int)d0;
            B.nouter = size / (int)d0;
            for (int a = 0; a < norb; a++)
                for (int i = 0; i < S->D[t[a]]; i++) B.nh = std::max(B.nh, (int)hopc[S->coff[t[a]] + i]);
        }
        blockoff[key] = cur;
        cur += size;
        S->blocks.push_back(B);
        S->max_block = std::max(S->max_block, size);
    }
    if (cur != (int)b->dim) return edgpu_fail(ctx, "star layout: internal size mismatch (%d vs %lld)", cur, (long long)b->dim);
    if ((int)S->blocks.size() > kMaxBlocks) return edgpu_fail(ctx, "star layout: too many blocks");
    // up-pass schedule: blocks of >= kBigBlock configurations get the persistent double-buffered kernel; the small
    // ones are packed into groups (up to kBigBlock elements) that one CTA walks for 4 rows at a time
    std::vector<int> groups, uplist;
    {
        int acc = 0, start = 0;
        for (int i = 0; i < (int)S->blocks.size(); i++) {
            const int sz = S->blocks[i].size;
            if (sz >= kBigBlock) { S->big_blocks.push_back(i); continue; }
            if (acc > 0 && acc + sz > kBigBlock) { groups.push_back(start); groups.push_back((int)uplist.size() - start); start = (int)uplist.size(); acc = 0; }
            uplist.push_back(i);
            acc += sz;
            S->max_small = std::max(S->max_small, sz);
        }
        if ((int)uplist.size() > start) { groups.push_back(start); groups.push_back((int)uplist.size() - start); }
        S->ngroups = (int)groups.size() / 2;
        if (groups.empty()) { groups.push_back(0); groups.push_back(0); }
        if (uplist.empty()) uplist.push_back(0);
    }
    // upload
    cudaStream_t st = ctx->stream;
    uint16_t *d_srank = nullptr;
    int *d_D = nullptr, *d_blockoff = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_srank, sizeof(uint16_t) * srank.size()));
    CUDA_TRY(ctx, cudaMalloc(&d_D, sizeof(int) * 16));
    CUDA_TRY(ctx, cudaMalloc(&d_blockoff, sizeof(int) * blockoff.size()));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_srank, srank.data(), sizeof(uint16_t) * srank.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_D, S->D, sizeof(int) * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_blockoff, blockoff.data(), sizeof(int) * blockoff.size(), cudaMemcpyHostToDevice, st));
    k_star_rank<<<((1u << ns) + 255) / 256, 256, 0, st>>>(ns, n, norb, nbath, d_srank, d_D, d_blockoff, b->rank);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMalloc(&S->d_blocks, sizeof(StarBlock) * S->blocks.size()));
    CUDA_TRY(ctx, cudaMalloc(&S->d_hopj, hopj.size()));
    CUDA_TRY(ctx, cudaMalloc(&S->d_hopd, sizeof(int16_t) * hopd.size()));
    CUDA_TRY(ctx, cudaMalloc(&S->d_hopc, hopc.size()));
    CUDA_TRY(ctx, cudaMalloc(&S->d_hopv, sizeof(double) * hopv.size()));
    CUDA_TRY(ctx, cudaMalloc(&S->d_estar, sizeof(double) * estar.size()));
    CUDA_TRY(ctx, cudaMalloc(&S->d_upgroups, sizeof(int) * groups.size()));
    CUDA_TRY(ctx, cudaMalloc(&S->d_uplist, sizeof(int) * uplist.size()));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_blocks, S->blocks.data(), sizeof(StarBlock) * S->blocks.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_hopj, hopj.data(), hopj.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_hopd, hopd.data(), sizeof(int16_t) * hopd.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_hopc, hopc.data(), hopc.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_hopv, hopv.data(), sizeof(double) * hopv.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_estar, estar.data(), sizeof(double) * estar.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_upgroups, groups.data(), sizeof(int) * groups.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_uplist, uplist.data(), sizeof(int) * uplist.size(), cudaMemcpyHostToDevice, st));
    {
        std::vector<int> fr;
        for (const StarBlock &B : S->blocks)
            if (B.size < kBulkMin)
                for (int e = 0; e < B.size; e++) fr.push_back(B.off + e);
        S->nfringe = (int)fr.size();
        if (S->nfringe) {
            CUDA_TRY(ctx, cudaMalloc(&S->d_fringe, sizeof(int) * fr.size()));
            CUDA_TRY(ctx, cudaMemcpyAsync(S->d_fringe, fr.data(), sizeof(int) * fr.size(), cudaMemcpyHostToDevice, st));
        }
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    cudaFree(d_srank); cudaFree(d_D); cudaFree(d_blockoff);
    b->star = S;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------
struct StarKParams {
    int norb, H, ncfg;
    int D[16], A0[16], coff[16];
    double pair_e;
};

static constexpr int kNT = 512;         // threads per CTA of the tiled kernels

// Shared-memory star tables of one block (all NORB stars): per configuration i of star a
//   s_off[a][i][h]  byte offset of the h-th source inside a [element][2] double2 plane: (j - i) * stride_a * 16
//   s_val[a][i][h]  signed amplitude V_{a,k} * (-1)^{popc(bath bits of the star below k)}
//   s_cnt[a][i]     number of sources, s_e[a][i] star diagonal energy (up pass only)
struct __align__(16) HopEnt { double val; int off; int pad; };
struct TabPtrs {
    HopEnt *ent;
    double *e;
    uint8_t *cnt;
};

template <int NORB>
__device__ __forceinline__ void load_tabs(const StarKParams &P, const StarBlock &B, const int *D,
                                          const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc,
                                          const double *__restrict__ hopv, const double *__restrict__ estar,
                                          const TabPtrs &T, int maxD, bool with_e)
{
    const int H = P.H;
    int stride = 1;
#pragma unroll
    for (int a = 0; a < NORB; a++) {
        const int Da = D[a], c0 = P.coff[B.n[a]];
        for (int t = threadIdx.x; t < Da * H; t += kNT) {
            HopEnt en;
            en.val = hopv[((size_t)a * P.ncfg + c0) * H + t];
            en.off = (int)hopd[(size_t)c0 * H + t] * stride * 16;
            en.pad = 0;
            T.ent[a * maxD * H + t] = en;
        }
        for (int t = threadIdx.x; t < Da; t += kNT) {
            T.cnt[a * maxD + t] = hopc[c0 + t];
            if (with_e) T.e[a * maxD + t] = estar[(size_t)a * P.ncfg + c0 + t];
        }
        stride *= Da;
    }
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a 32-byte row segment of a column strip moves in ONE request
// per thread instead of two, halving the L1TEX sector work of the (inherently 32-byte-granular) down pass.
__device__ __forceinline__ void ldg256(const double *p, double2 &a, double2 &b)
{
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
__device__ __forceinline__ void stg256(double *p, double a0, double a1, double a2, double a3)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a0), "d"(a1), "d"(a2), "d"(a3) : "memory");
}

__device__ __forceinline__ double2 lds128(uint32_t addr)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));     // not volatile: free to be scheduled early
    return v;
}

// acc[0..VEC) += sg * sum_h val[h] * tile[(e + delta_h)][0..VEC)   for one star; tile = VEC/2 planes of [e][2]
template <int VEC>
__device__ __forceinline__ void star_gather(double (&acc)[VEC], uint32_t a0, uint32_t plane, const HopEnt *ent, int cnt, double sg)
{
#pragma unroll 4
    for (int h = 0; h < cnt; h++) {
        const HopEnt en = ent[h];                          // one 16-byte shared-memory load: amplitude + byte offset
        const double v = sg * en.val;
        const uint32_t a = a0 + (uint32_t)en.off;
        const double2 p = lds128(a);
        acc[0] += v * p.x; acc[1] += v * p.y;
        if (VEC == 4) {
            const double2 q = lds128(a + plane);
            acc[2] += v * q.x; acc[3] += v * q.y;
        }
    }
}

// One pass over a tile held in shared memory as VEC/2 planes of [batch][element][2] doubles (element = o*D0 + i0).
// A tile holds `nbatch` independent copies of the block (row pairs in the up pass, column strips in the down
// pass).  Thread (ty, i0) walks the combined outer index (batch, o) = ty, ty+NY, ...;
// early = pre(batch, e, smem address) is issued before the gathers, f(batch, e, early, ui, esum, acc) consumes the
// VEC gathered sums; ui = impurity bits of the element, esum = sum of the star energies (up pass).
template <int NORB, int VEC, bool WITH_E, int KCH, class EARLY, class PRE, class F>
__device__ __forceinline__ void tile_pass(const StarBlock &B, const int *D, const int *A0, const TabPtrs &T, int maxD, int H,
                                          uint32_t s_in_addr, uint32_t plane, int nbatch, PRE pre, F f)
{
    const int tid = threadIdx.x, D0 = D[0];
    const int ty = (D0 == 1) ? tid : (int)__umulhi((uint32_t)tid, B.magic0);   // tid / D0 by multiplication
    const int i0 = tid - ty * D0;
    const int NY = B.ny;                                 // kNT / D0
    if (ty >= NY) return;
    const int O = B.nouter;                              // size / D0
    const uint32_t imp0 = i0 >= A0[0] ? 1u : 0u;
    const double e0 = WITH_E ? T.e[i0] : 0.0;
    const int cnt0 = T.cnt[i0];
    const HopEnt *ent0 = T.ent + i0 * H;
    // star 0 is the thread's own axis: its <= kRegH hop entries are loop invariant -> keep them in registers
    // (saves the per-lane-distinct 16-byte table reads, which cost 4-8 shared-memory wavefronts per warp and hop)
    constexpr int kRegH = 8;
    const bool reg0 = H <= kRegH;
    double rval[kRegH];
    int roff[kRegH];
#pragma unroll
    for (int h = 0; h < kRegH; h++) {
        const bool on = reg0 && h < cnt0;
        const HopEnt en = on ? ent0[h] : HopEnt{0.0, 0, 0};
        rval[h] = en.val; roff[h] = en.off;
    }
    const double c0s = (B.sgn_lower[0] & 1) ? -1.0 : 1.0;
    const double c1s = (NORB >= 2 && (B.sgn_lower[1] & 1)) ? -1.0 : 1.0;
    const double c2s = (NORB >= 3 && (B.sgn_lower[2] & 1)) ? -1.0 : 1.0;
    const double s0 = imp0 ? -1.0 : 1.0;
    // (bq, i2, i1) = mixed-radix digits of the outer index, advanced by NY per step
    int i1 = ty, i2 = 0, bq = 0;
    const int D1 = (NORB >= 2) ? D[1] : 1, D2 = (NORB >= 3) ? D[2] : 1;
    const bool slow_digits = NY >= D1 * 4 || (D2 == 1 && NORB >= 3);      // tiny blocks: carry loops would spin
    auto normalise = [&](int &j1, int &j2, int &jb) {
        if (slow_digits) {
            j2 += j1 / D1; j1 = j1 % D1;
            jb += j2 / D2; j2 = j2 % D2;
        } else {
            while (j1 >= D1) { j1 -= D1; j2++; }
            while (j2 >= D2) { j2 -= D2; jb++; }
        }
    };
    normalise(i1, i2, bq);
    const int total = nbatch * O;
    for (int ot0 = ty; ot0 < total; ot0 += NY * KCH) {
        // phase A: issue the long-latency loads of the next KCH steps (y of the up pass) so that ONE memory latency
        // is exposed per chunk instead of one per step
        EARLY early[KCH];
        {
            int j1 = i1, j2 = i2, jb = bq;
#pragma unroll
            for (int k = 0; k < KCH; k++) {
                const int ot = ot0 + k * NY;
                if (ot < total) early[k] = pre(jb, ot * D0 + i0 - jb * B.size);
                j1 += NY;
                normalise(j1, j2, jb);
            }
        }
        // phase B: gathers
#pragma unroll
        for (int k = 0; k < KCH; k++) {
            const int ot = ot0 + k * NY;
            if (ot < total) {
                const int et = ot * D0 + i0;                     // element index inside the whole tile
                const int e = et - bq * B.size;                  // element index inside the block
                const uint32_t a0 = s_in_addr + (uint32_t)et * 16u;
                uint32_t ui = imp0;
                double es = e0, s1 = 1.0, s2 = 1.0;
                if (NORB >= 2) { const bool b = i1 >= A0[1]; ui |= b ? 2u : 0u; s1 = b ? -1.0 : 1.0; if (WITH_E) es += T.e[maxD + i1]; }
                if (NORB >= 3) { const bool b = i2 >= A0[2]; ui |= b ? 4u : 0u; s2 = b ? -1.0 : 1.0; if (WITH_E) es += T.e[2 * maxD + i2]; }
                double acc[VEC];
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[v] = 0.0;
                // sign of star a: (-1)^{sum_{a'<a} n_a'} * prod_{a' != a} (-1)^{imp_a'}   (c/cdg rule, ED_SETUP.f90:1080-1106)
                if (reg0) {
                    const double sg = c0s * s1 * s2;
#pragma unroll
                    for (int h = 0; h < kRegH; h++) {
                        if (h < cnt0) {
                            const double v = sg * rval[h];
                            const uint32_t a = a0 + (uint32_t)roff[h];
                            const double2 p = lds128(a);
                            acc[0] += v * p.x; acc[1] += v * p.y;
                            if (VEC == 4) {
                                const double2 q = lds128(a + plane);
                                acc[2] += v * q.x; acc[3] += v * q.y;
                            }
                        }
                    }
                } else {
                    star_gather<VEC>(acc, a0, plane, ent0, cnt0, c0s * s1 * s2);
                }
                if (NORB >= 2) star_gather<VEC>(acc, a0, plane, T.ent + (maxD + i1) * H, T.cnt[maxD + i1], c1s * s0 * s2);
                if (NORB >= 3) star_gather<VEC>(acc, a0, plane, T.ent + (2 * maxD + i2) * H, T.cnt[2 * maxD + i2], c2s * s0 * s1);
                f(bq, e, a0, early[k], ui, es, acc);
            }
            i1 += NY;
            normalise(i1, i2, bq);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Lean kernels (large blocks: one row pair / one column strip per tile).  Same math as tile_pass, written as
// straight-line code: the hop lists are PADDED to NH entries per configuration (unrolled NH times, slots beyond the
// configuration's hop count are predicated off so they cost no shared-memory wavefronts), the block sign (-1)^{sum_{a'<a} n_a'} is folded into the amplitudes, the star-0 hop
// list of a thread lives in registers, all shared-memory accesses use 32-bit addresses that advance by constants,
// and the sign of a star (constant over its hops) multiplies the partial sum once.
//   lent[a][i][h] = {val, off}   off = (j - i) * stride_a * 16 bytes inside an [element][2] plane
//   laux[a][i]    = {star diagonal energy, (-1)^imp}
struct LeanTabs { uint32_t ent, aux; };                       // shared-memory byte addresses

static size_t lean_tabs_bytes(int norb, int maxD, int NH) { return (size_t)16 * norb * maxD * (NH + 1) + 64; }

template <int NORB, int NH>
__device__ __forceinline__ LeanTabs load_lean_tabs(const StarKParams &P, const StarBlock &B, const int *D, const int *A0,
                                                   const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc,
                                                   const double *__restrict__ hopv, const double *__restrict__ estar,
                                                   unsigned char *base, int maxD, int esz = 16)
{
    HopEnt *ent = reinterpret_cast<HopEnt *>(base);
    double2 *aux = reinterpret_cast<double2 *>(ent + NORB * maxD * NH);
    const int H = P.H;
    int stride = 1;
#pragma unroll
    for (int a = 0; a < NORB; a++) {
        const int Da = D[a], c0 = P.coff[B.n[a]];
        const double sgn = (B.sgn_lower[a] & 1) ? -1.0 : 1.0;
        for (int t = threadIdx.x; t < Da * NH; t += kNT) {
            const int i = t / NH, h = t - i * NH;
            const int cnt = (int)hopc[c0 + i];
            HopEnt en = {0.0, 0, cnt};                       // pad carries the hop count of the configuration
            if (h < cnt) {
                en.val = sgn * hopv[((size_t)a * P.ncfg + c0 + i) * H + h];
                en.off = (int)hopd[(size_t)(c0 + i) * H + h] * stride * esz;
            }
            ent[(a * maxD + i) * NH + h] = en;
        }
        for (int t = threadIdx.x; t < Da; t += kNT)
            aux[a * maxD + t] = make_double2(estar ? estar[(size_t)a * P.ncfg + c0 + t] : 0.0, t >= A0[a] ? -1.0 : 1.0);
        stride *= Da;
    }
    LeanTabs T;
    T.ent = (uint32_t)__cvta_generic_to_shared(ent);
    T.aux = (uint32_t)__cvta_generic_to_shared(aux);
    return T;
}

// predicated load (no branch): *p when flag != 0, else 0.0
__device__ __forceinline__ double ldg_if(const double *p, int flag)
{
    double v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@q ld.global.f64 %0, [%1];\n\t}" : "=d"(v) : "l"(p), "r"(flag) : "memory");
    return v;
}

__device__ __forceinline__ double lds64(uint32_t addr)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// Per-thread invariants of the lean passes: thread (ty, i0) owns star-0 configuration i0 and walks the outer index
// o = ty, ty + NY, ... of the block.
template <int NH>
struct LeanThread {
    int ty, i0, NY, O;
    bool active;
    double rval[NH];
    int roff[NH];
    int cnt0;
    double e0, s0;
    uint32_t dgoff0;                 // byte offset of the thread's impurity bit inside a dg row (8 doubles)
    __device__ __forceinline__ void init(const StarBlock &B, const int *D, const int *A0, const LeanTabs &T)
    {
        const int tid = threadIdx.x, D0 = D[0];
        ty = (D0 == 1) ? tid : (int)__umulhi((uint32_t)tid, B.magic0);
        i0 = tid - ty * D0;
        NY = B.ny;
        O = B.nouter;
        active = ty < NY && ty < O;                      // threads beyond the block's outer range have no element
        const int ii = active ? i0 : 0;
#pragma unroll
        for (int h = 0; h < NH; h++) {
            const double2 w = lds128(T.ent + (uint32_t)(ii * NH + h) * 16u);
            rval[h] = w.x;
            roff[h] = __double2loint(w.y);
            if (h == 0) cnt0 = __double2hiint(w.y);
        }
        const double2 ax = lds128(T.aux + (uint32_t)ii * 16u);
        e0 = ax.x; s0 = ax.y;
        dgoff0 = ii >= A0[0] ? 8u : 0u;
    }
};

// acc[v] (+)= sg * sum_h val[h] * tile[e + delta_h][v]  for star a >= 1 of the element whose star index is ia
template <int NH, int VEC, bool FIRST>
__device__ __forceinline__ void lean_star(double (&acc)[VEC], uint32_t a0, uint32_t plane, uint32_t ent_addr, double sg)
{
    double p[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) p[v] = 0.0;
    int cnt = NH;
#pragma unroll
    for (int h = 0; h < NH; h++) {
        const double2 w = lds128(ent_addr + (uint32_t)h * 16u);            // {amplitude, byte offset | count}: uniform over most of the warp
        if (h == 0) cnt = __double2hiint(w.y);
        if (h < cnt) {
            const uint32_t a = a0 + (uint32_t)__double2loint(w.y);
            if (VEC == 1) { p[0] = fma(w.x, lds64(a), p[0]); continue; }
            const double2 u = lds128(a);
            p[0] = fma(w.x, u.x, p[0]); p[VEC > 1 ? 1 : 0] = fma(w.x, u.y, p[VEC > 1 ? 1 : 0]);
            if (VEC == 4) {
                const double2 q = lds128(a + plane);
                p[2] = fma(w.x, q.x, p[2]); p[3] = fma(w.x, q.y, p[3]);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = FIRST ? sg * p[v] : fma(sg, p[v], acc[v]);
}

// The gathers of one element: acc = sum over stars.  Returns the summed star energy in `es` and the dg byte offset.
template <int NORB, int NH, int VEC>
__device__ __forceinline__ void lean_element(const LeanThread<NH> &L, const LeanTabs &T, const int *A0, int maxD, uint32_t a0, uint32_t plane,
                                             int i1, int i2, const double (&init)[VEC], double (&acc)[VEC], double &es, uint32_t &dgo)
{
    double sg0 = 1.0, sg1 = L.s0, sg2 = L.s0;
    es = L.e0;
    dgo = L.dgoff0;
    if (NORB >= 2) {
        const double2 ax = lds128(T.aux + (uint32_t)(maxD + i1) * 16u);
        es += ax.x; sg0 = ax.y; sg2 *= ax.y;
        dgo += i1 >= A0[1] ? 16u : 0u;
    }
    if (NORB >= 3) {
        const double2 ax = lds128(T.aux + (uint32_t)(2 * maxD + i2) * 16u);
        es += ax.x; sg0 *= ax.y; sg1 *= ax.y;
        dgo += i2 >= A0[2] ? 32u : 0u;
    }
    double p[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) p[v] = 0.0;
#pragma unroll
    for (int h = 0; h < NH; h++) {
        if (h < L.cnt0) {
            const uint32_t a = a0 + (uint32_t)L.roff[h];
            if (VEC == 1) { p[0] = fma(L.rval[h], lds64(a), p[0]); continue; }
            const double2 u = lds128(a);
            p[0] = fma(L.rval[h], u.x, p[0]); p[VEC > 1 ? 1 : 0] = fma(L.rval[h], u.y, p[VEC > 1 ? 1 : 0]);
            if (VEC == 4) {
                const double2 q = lds128(a + plane);
                p[2] = fma(L.rval[h], q.x, p[2]); p[3] = fma(L.rval[h], q.y, p[3]);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = fma(sg0, p[v], init[v]);
    if (NORB >= 2) lean_star<NH, VEC, false>(acc, a0, plane, T.ent + (uint32_t)((maxD + i1) * NH) * 16u, sg1);
    if (NORB >= 3) lean_star<NH, VEC, false>(acc, a0, plane, T.ent + (uint32_t)((2 * maxD + i2) * NH) * 16u, sg2);
}

__device__ __forceinline__ TabPtrs carve_tabs(unsigned char *base, int norb, int maxD, int H)
{
    TabPtrs T;
    T.ent = reinterpret_cast<HopEnt *>(base);
    T.e = reinterpret_cast<double *>(T.ent + norb * maxD * H);
    T.cnt = reinterpret_cast<uint8_t *>(T.e + norb * maxD);
    return T;
}
static size_t tabs_bytes(int norb, int maxD, int H)
{
    return 16 * (size_t)norb * maxD * H + sizeof(double) * (size_t)norb * maxD + (size_t)norb * maxD + 64;
}

__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Row shard of a multi-GPU run: after the all-to-all the rows of this rank arrive as one slab per source rank
// ([nrows][ldc_p] each, source rank p owns columns [col0_p, col0_p + ldc_p)); the up pass reads and writes the slabs
// in place instead of paying an unpack and a pack pass.  n == 1 (single GPU / contiguous rows) is the plain layout.
// peer == 1 ("peer mode", one process per GPU): slab p is not a piece of a received buffer but the column shard of rank p
// itself, [dim_dw][ldc_p] in THAT GPU's memory (CUDA IPC mapping over NVLink): xp[p] / yp[p] are its base pointers and
// base[p] = row0 * ldc[p] selects this rank's rows.  The copy engine then reads x from, and writes the result to, the
// owners directly -- the two all-to-all transposes of the exchange are fused into the up-pass kernel.
struct SlabMap {
    int n, ldc0;
    unsigned magic;                 // ceil(2^32 / ldc0): column / ldc0 by multiplication
    int col0[8], ldc[8];
    long long base[8];              // element offset of slab p (y, and x unless peer)
    int peer;
    const double *xp[8];
    double *yp[8];
    long long xbase[8];             // peer mode: element offset of this rank's first row inside xp[p]
};
__device__ __forceinline__ int slab_of(const SlabMap &M, int c)
{
    const int p = (int)__umulhi((unsigned)c, M.magic);
    return p < M.n - 1 ? p : M.n - 1;
}
__device__ __forceinline__ int64_t slab_off(const SlabMap &M, int64_t row, int c)
{
    const int p = slab_of(M, c);
    return M.base[p] + row * M.ldc[p] + (c - M.col0[p]);
}

// Persistent, double-buffered up pass for one up-block: y[rows][blk] += (diag + H_up) x.
// A tile = RP row pairs x the block, stored [rp][element][2] (the two rows of a pair interleaved).  Every CTA
// loads the star tables once, then walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the next tile streams in
// with cp.async (LDGSTS) while the current one is processed, so the HBM pipe stays busy during the gathers.
//   NH > 0: lean path (RP == 1, hop lists padded to NH);  NH == 0: generic tile_pass path (small blocks, Nbath = 9)
template <int NORB, bool SLAB, int NH>
__global__ void __launch_bounds__(kNT)
k_star_up(StarKParams P, SlabMap Mpar, int64_t dim_dw, int64_t ld, int block_index, int RP, int accumulate, int nstage,
          const StarBlock *__restrict__ blocks,
          const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc, const double *__restrict__ hopv,
          const double *__restrict__ estar, const double *__restrict__ e_dw, const uint32_t *__restrict__ cfg_dw,
          const double *__restrict__ xtab, const double *__restrict__ x, double *__restrict__ y, int maxD)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ SlabMap s_M;
    if (SLAB) {
        if (threadIdx.x == 0) s_M = Mpar;
        __syncthreads();
    }
    const SlabMap &M = s_M;
    const StarBlock B = blocks[block_index];
    const int size = B.size, tid = threadIdx.x;
    const int tile_elems = RP * size;
    double *s_buf = reinterpret_cast<double *>(smem_raw);                      // [2 stages][RP][size][2]
    double *s_dg = s_buf + (size_t)2 * nstage * tile_elems;                    // [stages][RP][2 rows][8]
    unsigned char *tab_base = reinterpret_cast<unsigned char *>(s_dg + (size_t)16 * nstage * RP);
    int D[3] = {1, 1, 1}, A0[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < NORB; a++) { D[a] = P.D[B.n[a]]; A0[a] = P.A0[B.n[a]]; }
    TabPtrs T;
    LeanTabs LT;
    if (NH > 0) LT = load_lean_tabs<NORB, (NH > 0 ? NH : 1)>(P, B, D, A0, hopd, hopc, hopv, estar, tab_base, maxD);
    else { T = carve_tabs(tab_base, NORB, maxD, P.H); load_tabs<NORB>(P, B, D, hopd, hopc, hopv, estar, T, maxD, true); }
    const uint32_t buf_addr = (uint32_t)__cvta_generic_to_shared(s_buf);
    const uint32_t dg_addr = (uint32_t)__cvta_generic_to_shared(s_dg);
    const uint32_t stage_bytes = (uint32_t)tile_elems * 16u;
    const uint32_t impmask = (1u << NORB) - 1u;
    const int64_t npairs = (dim_dw + 1) / 2;
    const int64_t ntiles = (npairs + RP - 1) / RP;
    const int boff = B.off;
    const int64_t last = dim_dw - 1;

    auto issue = [&](int64_t t, int stage) {
        const uint32_t dst = buf_addr + (uint32_t)stage * stage_bytes;
        for (int q = 0; q < RP; q++) {
            // rows past the end duplicate the last row (computed, never stored)
            int64_t ra = 2 * (t * RP + q), rb = ra + 1;
            ra = ra < last ? ra : last;
            rb = rb < last ? rb : last;
            const double *xa = x + ra * ld + boff, *xb = x + rb * ld + boff;
            const uint32_t d2 = dst + (uint32_t)q * (uint32_t)size * 16u;
            for (int e = tid; e < size; e += kNT) {
                if (SLAB) {
                    cp_async8(d2 + (uint32_t)e * 16u, x + slab_off(M, ra, boff + e));
                    cp_async8(d2 + (uint32_t)e * 16u + 8u, x + slab_off(M, rb, boff + e));
                } else {
                    cp_async8(d2 + (uint32_t)e * 16u, xa + e);
                    cp_async8(d2 + (uint32_t)e * 16u + 8u, xb + e);
                }
            }
        }
        if (NH > 0 && !SLAB && accumulate && tid < 2) {
            // pull the y rows of that tile (written by the down pass, long evicted) into L2 now, so that the
            // read-modify-write of the gather loop sees an L2 hit latency instead of a loaded-DRAM one
            int64_t r = 2 * t + tid;
            r = r < last ? r : last;
            const uintptr_t p0 = reinterpret_cast<uintptr_t>(y + r * ld + boff) & ~(uintptr_t)15;
            const uint32_t bytes = ((uint32_t)size * 8u + 31u) & ~15u;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"(bytes) : "memory");
        }
        for (int i = tid; i < 16 * RP; i += kNT) {
            // s_dg[stage][q][v][ui] = E_dw[row] + X[imp_dw(row)][ui] + (Ust-Jh) * C(nimp(ui), 2)
            const int q = i >> 4, v = (i >> 3) & 1, ui = i & 7;
            int64_t r = 2 * (t * RP + q) + v;
            r = r < last ? r : last;
            const int nimp = __popc(ui);
            s_dg[(size_t)stage * 16 * RP + i] = e_dw[r] + xtab[(cfg_dw[r] & impmask) * 32u + ui] + P.pair_e * (double)(nimp * (nimp - 1) / 2);
        }
    };

    LeanThread<(NH > 0 ? NH : 1)> L;
    if (NH > 0) {
        __syncthreads();                                     // lean tables visible
        L.init(B, D, A0, LT);
    }

    // nstage == 2: one CTA per SM, the next tile streams in while the current one is processed.
    // nstage == 1: single stage (blocks whose two stages exceed shared memory).
    int64_t t = blockIdx.x;
    if (nstage == 2) {
        if (t < ntiles) issue(t, 0);
        cp_async_commit();
    }
    int stage = 0;
    for (; t < ntiles; t += gridDim.x, stage ^= (nstage - 1)) {
        if (nstage == 2) {
            const int64_t nxt = t + gridDim.x;
            if (nxt < ntiles) issue(nxt, stage ^ 1);
            cp_async_commit();
            cp_async_wait<1>();                              // this thread's copies of the current stage have landed
        } else {
            issue(t, 0);
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();                                     // ... and everybody's; tables + s_dg visible too
        const int64_t row0 = 2 * t * RP;
        if (NH > 0) {
            // large block: one row pair per tile.  A row past the end (odd dim_dw) aliases the last row: both slots
            // then compute and store the same value.
            constexpr int NHc = NH > 0 ? NH : 1;
            const int64_t rra = row0 < last ? row0 : last, rrb = row0 + 1 < last ? row0 + 1 : last;
            double *pa = y + rra * ld + boff, *pb = y + rrb * ld + boff;
            const uint32_t xs = buf_addr + (uint32_t)stage * stage_bytes;
            const uint32_t dgs = dg_addr + (uint32_t)stage * 128u;
            if (L.active) {
                const int D0 = D[0], D1 = D[1], NY = L.NY, O = L.O;
                int i1 = L.ty, i2 = 0;
                if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
                int e = L.ty * D0 + L.i0;
                const int estep = NY * D0;
                uint32_t a0 = xs + (uint32_t)e * 16u;
                // y addresses: uniform row base + 32-bit byte offset (one IMAD.WIDE per access instead of 64-bit index math)
                char *const ba = reinterpret_cast<char *>(pa), *const bb = reinterpret_cast<char *>(pb);
                auto yaddr = [&](int64_t row, char *rowb, int ee) -> double * {
                    return SLAB ? y + slab_off(M, row, boff + ee) : reinterpret_cast<double *>(rowb + (uint32_t)ee * 8u);
                };
                // software pipeline: the y values of the next PF steps are in flight during this one (slots past the
                // end re-read the thread's current element, which is harmless)
                constexpr int PF = 3;
                double2 ring[PF];
#pragma unroll
                for (int k = 0; k < PF; k++) {
                    const int ek = (L.ty + k * NY < O) ? e + k * estep : e;
                    ring[k].x = ldg_if(yaddr(rra, ba, ek), accumulate);
                    ring[k].y = ldg_if(yaddr(rrb, bb, ek), accumulate);
                }
                int o = L.ty;
                while (o < O) {
#pragma unroll
                    for (int k = 0; k < PF; k++) {
                        if (o < O) {
                            const double init[2] = {ring[k].x, ring[k].y};
                            const int en = (o + PF * NY < O) ? e + PF * estep : e;
                            ring[k].x = ldg_if(yaddr(rra, ba, en), accumulate);
                            ring[k].y = ldg_if(yaddr(rrb, bb, en), accumulate);
                            double acc[2], es;
                            uint32_t dgo;
                            lean_element<NORB, NHc, 2>(L, LT, A0, maxD, a0, 0u, i1, i2, init, acc, es, dgo);
                            const double2 own = lds128(a0);
                            const double da = lds64(dgs + dgo), db = lds64(dgs + 64u + dgo);
                            *yaddr(rra, ba, e) = fma(es + da, own.x, acc[0]);
                            *yaddr(rrb, bb, e) = fma(es + db, own.y, acc[1]);
                            e += estep;
                            a0 += (uint32_t)estep * 16u;
                            i1 += NY;
                            if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
                            o += NY;
                        }
                    }
                }
            }
        } else {
        const double *dg = s_dg + (size_t)stage * 16 * RP;
        struct Own { double y0, y1; };
        auto rows = [&](int q, double *&ya, double *&yb, bool &oka, bool &okb) {
            const int64_t ra = row0 + 2 * q;
            oka = ra < dim_dw; okb = ra + 1 < dim_dw;
            // SLAB: the "pointers" carry the row number; the element address is resolved per access
            ya = SLAB ? reinterpret_cast<double *>(oka ? ra : last) : y + (oka ? ra : last) * ld + boff;
            yb = SLAB ? reinterpret_cast<double *>(okb ? ra + 1 : last) : y + (okb ? ra + 1 : last) * ld + boff;
        };
        auto at = [&](double *rowp, int e) -> double * {
            return SLAB ? y + slab_off(M, reinterpret_cast<int64_t>(rowp), boff + e) : rowp + e;
        };
        int curq = -1, curq2 = -1;                           // row pointers are recomputed only when the row pair changes
        double *ya = nullptr, *yb = nullptr, *ya2 = nullptr, *yb2 = nullptr;
        bool oka = false, okb = false, oka2 = false, okb2 = false;
        tile_pass<NORB, 2, true, 1, Own>(B, D, A0, T, maxD, P.H, buf_addr + (uint32_t)stage * stage_bytes, 0u, RP,
            [&](int q, int e) {
                if (q != curq2) { curq2 = q; rows(q, ya2, yb2, oka2, okb2); }
                Own w;
                w.y0 = accumulate ? *at(ya2, e) : 0.0;       // H_dw x written by the down pass
                w.y1 = accumulate ? *at(yb2, e) : 0.0;
                return w;
            },
            [&](int q, int e, uint32_t a0, const Own &w, uint32_t ui, double es, double (&acc)[2]) {
                if (q != curq) { curq = q; rows(q, ya, yb, oka, okb); }
                const double2 p = lds128(a0);
                const double *dq = dg + q * 16 + ui;
                if (oka) *at(ya, e) = w.y0 + acc[0] + (es + dq[0]) * p.x;
                if (okb) *at(yb, e) = w.y1 + acc[1] + (es + dq[8]) * p.y;
            });
        }
        __syncthreads();                                     // all reads of this stage done before it is refilled
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------------------
// Up pass, bulk-copy (TMA) version for the single-GPU path:  y[rows][blk] += (diag + H_up) x.
// A tile = G consecutive rows x the up-block.  All global traffic moves through the bulk async-copy engine:
//   producer warp : cp.async.bulk  x row segment -> xbuf[stage],  y row segment (H_dw x of the down pass) -> ybuf[stage]
//   16 consumer warps : gather from xbuf, read-modify-write ybuf IN PLACE (each element is owned by one thread)
//   producer warp : cp.async.bulk  ybuf[stage] -> y row segment
// so the LSU/L1 path (whose few in-flight lines throttled the cp.async version) carries shared-memory traffic
// only, and the consumers never meet a CTA-wide barrier: stages are handed over through mbarriers
// (full[stage]: copy bytes landed; done[stage]: one arrival per consumer warp).
// Bulk copies need 16-byte aligned addresses and sizes; a block that starts at an odd column is copied from one
// element earlier (`lead`), and an odd length is rounded up.  The extra elements belong to the neighbouring block or
// the pad columns of the SAME row; they are stored back unchanged, which is safe because the launches of
// different up-blocks are serialised on the stream and no other CTA of this launch touches the row.
static constexpr int kNT3 = kNT + 32;                    // 16 consumer warps + 1 producer warp

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

static constexpr int kMaxXS = 6;                         // x stages of the up pipeline (y stages: 1 or 2)

template <int NORB, int NH, bool GEN>                     // GEN = false: nxs = nys = 2 known at compile time (single GPU)
__global__ void __launch_bounds__(kNT3)
k_star_up3(StarKParams P, SlabMap Mpar, int accumulate, int64_t dim_dw, int64_t ld, int block_index, int G, int nxs_arg, int nys_arg,
           const StarBlock *__restrict__ blocks,
           const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc, const double *__restrict__ hopv,
           const double *__restrict__ estar, const double *__restrict__ e_dw, const uint32_t *__restrict__ cfg_dw,
           const double *__restrict__ xtab, const double *__restrict__ x, double *__restrict__ y, int maxD,
           double *__restrict__ dot_out)
{
    // Stages: tile i (i-th tile of this CTA) uses x stage i % nxs and y stage i % nys.  The x images are loaded nxs - nys
    // tiles further ahead than the y images (remote x segments of the peer mode have NVLink latency to hide; a y image is
    // also the output buffer, it frees only after its store has been read by the copy engine).
    //   fullx[sx], fully[sy]: bytes landed (one expect_tx arrival by the producer)
    //   done[sx]            : one arrival per consumer warp after tile i
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_dot[kNT / 32];
    __shared__ SlabMap s_M;
    __shared__ uint64_t s_bar[2 * kMaxXS + 2];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");            // the next launch may start its prologue
    const int nxs = GEN ? nxs_arg : 2, nys = GEN ? nys_arg : 2;
    if (threadIdx.x == 0) s_M = Mpar;                                          // visible after the barrier below
    const SlabMap &M = s_M;
    const StarBlock B = blocks[block_index];
    const int size = B.size, tid = threadIdx.x;
    const int lead = B.off & 1;                                                // copy starts one element early when the block starts odd
    const int ncopy = (lead + size + 1) & ~1;                                  // elements per row segment moved (even)
    const uint32_t rowb = (uint32_t)(ncopy + 2) * 8u;                          // bytes per row slot (16-byte multiple)
    const uint32_t stageb = rowb * (uint32_t)G;
    double *s_dg = reinterpret_cast<double *>(smem_raw + (size_t)(nxs + nys) * stageb);   // [nxs][G][8]
    unsigned char *tab_base = reinterpret_cast<unsigned char *>(s_dg + (size_t)8 * G * nxs);
    int D[3] = {1, 1, 1}, A0[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < NORB; a++) { D[a] = P.D[B.n[a]]; A0[a] = P.A0[B.n[a]]; }
    const LeanTabs LT = load_lean_tabs<NORB, NH>(P, B, D, A0, hopd, hopc, hopv, estar, tab_base, maxD, 8);
    const uint32_t xbuf = (uint32_t)__cvta_generic_to_shared(smem_raw);        // [nxs][G][rowb]
    const uint32_t ybuf = xbuf + (uint32_t)nxs * stageb;                       // [nys][G][rowb]
    const uint32_t dg_addr = (uint32_t)__cvta_generic_to_shared(s_dg);
    const uint32_t bfx = (uint32_t)__cvta_generic_to_shared(s_bar);            // fullx[kMaxXS]
    const uint32_t bdn = bfx + 8u * kMaxXS;                                    // done[kMaxXS]
    const uint32_t bfy = bdn + 8u * kMaxXS;                                    // fully[2]
    if (tid == 0) {
        for (int k = 0; k < kMaxXS; k++) { mbar_init(bfx + 8 * k, 1); mbar_init(bdn + 8 * k, kNT / 32); }
        mbar_init(bfy, 1); mbar_init(bfy + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                                                           // tables + barriers ready
    // programmatic dependent launch: everything above (tables, barriers) overlapped the tail of the previous launch on
    // the stream; its results (and the boundary elements it stored) are visible after this wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t impmask = (1u << NORB) - 1u;
    const int64_t ntiles = (dim_dw + G - 1) / G;
    const int64_t colb = (int64_t)B.off - lead;                                // first column moved (even)
    const int ntot = blockIdx.x < ntiles ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;   // tiles of this CTA
    auto tile_of = [&](int i) -> int64_t { return (int64_t)blockIdx.x + (int64_t)i * gridDim.x; };

    if (tid >= kNT) {
        // ---------------- producer warp ----------------
        const int lane = tid - kNT;
        // the row segment [colb, colb + ncopy) as pieces of the vector: one piece, or one per slab of a row shard that
        // arrived by all-to-all / per column shard of a peer (boundaries are multiples of 4 columns: 16-byte aligned)
        auto for_segments = [&](int64_t row, auto &&fn) {                      // fn(x pointer, y pointer, column - colb, count)
            if (M.n <= 1 && !M.peer) { fn(x + row * ld + colb, y + row * ld + colb, 0, ncopy); return; }
            const int64_t cend = colb + ncopy;
            for (int p = 0; p < M.n; p++) {
                const int64_t lo = colb > M.col0[p] ? colb : (int64_t)M.col0[p];
                const int64_t he = (int64_t)M.col0[p] + M.ldc[p], hi = cend < he ? cend : he;
                if (lo < hi) {
                    const int64_t off = row * M.ldc[p] + (lo - M.col0[p]);
                    fn(M.peer ? M.xp[p] + M.xbase[p] + off : x + M.base[p] + off, (M.peer ? M.yp[p] : y) + M.base[p] + off, (int)(lo - colb), (int)(hi - lo));
                }
            }
        };
        auto y_elem = [&](int64_t row, int64_t c) -> double * {
            if (M.n <= 1 && !M.peer) return y + row * ld + c;
            const int p = slab_of(M, (int)c);
            return (M.peer ? M.yp[p] : y) + M.base[p] + row * M.ldc[p] + (c - M.col0[p]);
        };
        auto rows_of = [&](int i, int64_t &r0) -> int {
            r0 = tile_of(i) * G;
            return (int)((dim_dw - r0) < G ? (dim_dw - r0) : G);
        };
        auto load_x = [&](int i) {                                             // x image + diagonal terms of tile i
            const int sx = i % nxs;
            int64_t r0;
            const int gc = rows_of(i, r0);
            for (int j = lane; j < 8 * gc; j += 32) {
                // s_dg[stage][g][ui] = E_dw[row] + X[imp_dw(row)][ui] + (Ust-Jh) * C(nimp(ui), 2)
                const int g = j >> 3, ui = j & 7;
                const int64_t r = r0 + g;
                const int nimp = __popc(ui);
                s_dg[(size_t)sx * 8 * G + j] = e_dw[r] + xtab[(cfg_dw[r] & impmask) * 32u + ui] + P.pair_e * (double)(nimp * (nimp - 1) / 2);
            }
            __syncwarp();
            if (lane == 0) {
                const uint32_t fb = bfx + 8 * sx;
                mbar_expect_tx(fb, (uint32_t)gc * (uint32_t)ncopy * 8u);
                for (int g = 0; g < gc; g++) {
                    const uint32_t xd = xbuf + (uint32_t)sx * stageb + (uint32_t)g * rowb;
                    for_segments(r0 + g, [&](const double *xs, double *, int rel, int cnt) { bulk_g2s(xd + (uint32_t)rel * 8u, xs, (uint32_t)cnt * 8u, fb); });
                }
            }
        };
        auto load_y = [&](int i) {
            if (lane != 0) return;
            const int sy = i % nys;
            int64_t r0;
            const int gc = rows_of(i, r0);
            const uint32_t fb = bfy + 8 * sy;
            // accumulate == 0 (the result overwrites y): nothing is loaded -- the barrier only hands the buffer over (the copy
            // engine has finished reading the previous result, see the steady-state loop)
            if (!accumulate) { mbar_arrive(fb); return; }
            mbar_expect_tx(fb, (uint32_t)gc * (uint32_t)ncopy * 8u);
            for (int g = 0; g < gc; g++) {
                const uint32_t yd = ybuf + (uint32_t)sy * stageb + (uint32_t)g * rowb;
                for_segments(r0 + g, [&](const double *, double *ys, int rel, int cnt) { bulk_g2s(yd + (uint32_t)rel * 8u, ys, (uint32_t)cnt * 8u, fb); });
            }
        };
        auto store_tile = [&](int i) {
            if (lane != 0) return;
            int64_t r0;
            const int gc = rows_of(i, r0);
            // accumulate != 0: the image holds the loaded neighbours, the whole (16-byte aligned) range goes back.
            // accumulate == 0: a 16-byte pair shared with the neighbouring block (first pair when the block starts at an odd
            // column, last pair when it ends at an even one) must not be stored as a pair: the bulk range shrinks to the
            // interior and the block's own edge element is written with an ordinary 8-byte store.
            const bool tail = ncopy > lead + size;
            for (int g = 0; g < gc; g++) {
                const uint32_t src = ybuf + (uint32_t)(i % nys) * stageb + (uint32_t)g * rowb;
                for_segments(r0 + g, [&](const double *, double *yd, int rel, int cnt) {
                    int lo = rel, hi = rel + cnt;
                    if (!accumulate) {
                        if (lead && lo == 0) { lo = 2 < hi ? 2 : hi; yd[1 - rel] = lds64(src + 8u); }
                        if (tail && hi == ncopy) { hi = ncopy - 2 > lo ? ncopy - 2 : lo; yd[ncopy - 2 - rel] = lds64(src + (uint32_t)(ncopy - 2) * 8u); }
                    }
                    if (hi > lo) bulk_s2g(yd + (lo - rel), src + (uint32_t)lo * 8u, (uint32_t)(hi - lo) * 8u);
                });
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        };
        auto wait_done = [&](int i) {                                          // consumers have finished tile i
            if (lane == 0) mbar_wait(bdn + 8 * (i % nxs), (uint32_t)(i / nxs) & 1u);
            __syncwarp();
        };
        // prologue
        for (int i = 0; i < nxs && i < ntot; i++) load_x(i);
        for (int i = 0; i < nys && i < ntot; i++) load_y(i);
        // steady state: tile i - nys done -> store it, free its y stage for tile i and its x stage for tile i - nys + nxs
        for (int i = nys; i < ntot + nys; i++) {
            wait_done(i - nys);
            store_tile(i - nys);
            if (i < ntot) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                load_y(i);
            }
            if (i - nys + nxs < ntot) load_x(i - nys + nxs);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        return;
    }

    // ---------------- consumer warps ----------------
    double dsum = 0.0;                                                        // partial <x, H x> (fused Lanczos alpha)
    LeanThread<NH> L;
    L.init(B, D, A0, LT);
    const int D0 = D[0], D1 = D[1], NY = L.NY, O = L.O;
    const int estep = NY * D0;
    for (int i = 0; i < ntot; i++) {
        const int sx = i % nxs, sy = i % nys;
        mbar_wait(bfx + 8 * sx, (uint32_t)(i / nxs) & 1u);
        mbar_wait(bfy + 8 * sy, (uint32_t)(i / nys) & 1u);
        const int64_t r0 = tile_of(i) * G;
        const int gc = (int)((dim_dw - r0) < G ? (dim_dw - r0) : G);
        if (L.active) {
            for (int g = 0; g < gc; g++) {
                const uint32_t xs = xbuf + (uint32_t)sx * stageb + (uint32_t)g * rowb + (uint32_t)lead * 8u;
                const uint32_t ys = ybuf + (uint32_t)sy * stageb + (uint32_t)g * rowb + (uint32_t)lead * 8u;
                const uint32_t dgs = dg_addr + (uint32_t)(sx * G + g) * 64u;
                int i1 = L.ty, i2 = 0;
                if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
                const uint32_t eb = (uint32_t)(L.ty * D0 + L.i0) * 8u;
                uint32_t a0 = xs + eb, ya = ys + eb;
#pragma unroll 2
                for (int o = L.ty; o < O; o += NY) {
                    const double init[1] = {accumulate ? lds64(ya) : 0.0};
                    double acc[1], es;
                    uint32_t dgo;
                    lean_element<NORB, NH, 1>(L, LT, A0, maxD, a0, 0u, i1, i2, init, acc, es, dgo);
                    const double own = lds64(a0);
                    const double res = fma(es + lds64(dgs + dgo), own, acc[0]);
                    sts64(ya, res);
                    dsum = fma(own, res, dsum);
                    a0 += (uint32_t)estep * 8u;
                    ya += (uint32_t)estep * 8u;
                    i1 += NY;
                    if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // results visible to the copy engine
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(bdn + 8 * sx);
    }
    if (dot_out) {
        // fixed-order reduction over the 16 consumer warps (named barrier: the producer warp is not part of it)
        for (int o = 16; o > 0; o >>= 1) dsum += __shfl_down_sync(0xffffffffu, dsum, o);
        if ((tid & 31) == 0) s_dot[tid >> 5] = dsum;
        asm volatile("bar.sync 1, %0;" ::"n"(kNT) : "memory");
        if (tid == 0) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kNT / 32; w++) v += s_dot[w];
            dot_out[blockIdx.x] = v;
        }
    }
}

// y[blk rows][strips] = H_dw x  for one down-block (runs FIRST; every row belongs to exactly one down-block, so
// this pass writes every element of y once).  A tile = SP strips of W columns x the block rows.
//   W = 4: two planes [strip][row][2] (columns 0-1 and 2-3 of each strip), 32-byte row segments (256-bit ld/st)
//   W = 2: one plane, 16-byte row segments -- for blocks whose 4-column tile exceeds shared memory (Norb=3, Nbath=5:
//          8000 rows x 32 B = 256 KB)
template <int NORB, int W, int NH>
__global__ void __launch_bounds__(kNT)
k_star_dw(StarKParams P, int64_t dim_up, int64_t ld, int block_index, int SP,
          const StarBlock *__restrict__ blocks,
          const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc, const double *__restrict__ hopv,
          const double *__restrict__ x, double *__restrict__ y, int maxD, int pf_dist)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const StarBlock B = blocks[block_index];
    const int R = B.size;
    const int tile_rows = SP * R;
    double *s_in = reinterpret_cast<double *>(smem_raw);                       // W/2 planes of [SP][R][2]
    unsigned char *tab_base = reinterpret_cast<unsigned char *>(s_in + (size_t)W * tile_rows);
    const int tid = threadIdx.x;
    int D[3] = {1, 1, 1}, A0[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < NORB; a++) { D[a] = P.D[B.n[a]]; A0[a] = P.A0[B.n[a]]; }
    TabPtrs T;
    LeanTabs LT;
    if (NH > 0) LT = load_lean_tabs<NORB, (NH > 0 ? NH : 1)>(P, B, D, A0, hopd, hopc, hopv, nullptr, tab_base, maxD);
    else { T = carve_tabs(tab_base, NORB, maxD, P.H); load_tabs<NORB>(P, B, D, hopd, hopc, hopv, nullptr, T, maxD, false); }
    const uint32_t s_in_addr = (uint32_t)__cvta_generic_to_shared(s_in);
    const uint32_t plane = (uint32_t)tile_rows * 16u;
    const int64_t cbase = (int64_t)blockIdx.x * SP * W;
    // stage the strips: W contiguous doubles per row and strip; ld is a multiple of 4 so every row segment is aligned to
    // its size; pad columns beyond dim_up are zero in x and are never written in y
    const double *xs = x + (int64_t)B.off * ld;
    for (int q = 0; q < SP; q++) {
        const int64_t c0 = cbase + W * q;
        const bool live = c0 < dim_up;
        for (int r = tid; r < R; r += kNT) {
            double2 a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
            if (live) {
                if (W == 4) ldg256(xs + (int64_t)r * ld + c0, a, b);
                else a = *reinterpret_cast<const double2 *>(xs + (int64_t)r * ld + c0);
            }
            *reinterpret_cast<double2 *>(s_in + (size_t)2 * (q * R + r)) = a;
            if (W == 4) *reinterpret_cast<double2 *>(s_in + (size_t)2 * tile_rows + (size_t)2 * (q * R + r)) = b;
        }
    }
    if (NH > 0 && W == 4) {
        // one CTA per SM and no second buffer: pull the strip that the NEXT CTA of this SM will stage into L2 while this
        // one computes, so that its staging loads see an L2 hit instead of a loaded-DRAM latency
        const int64_t cn = cbase + (int64_t)gridDim.y * 0 + (int64_t)pf_dist * SP * W;
        if (cn < dim_up)
            for (int r = tid; r < R; r += kNT) asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + (int64_t)r * ld + cn));
    }
    __syncthreads();
    double *ys = y + (int64_t)B.off * ld;
    const int64_t ldv = ld;
    auto store = [=](double *yp, int64_t left, const double (&acc)[W]) {
        if (left >= W) {
            if (W == 4) stg256(yp, acc[0], acc[1], acc[W - 2], acc[W - 1]);
            else *reinterpret_cast<double2 *>(yp) = make_double2(acc[0], acc[1]);
        } else if (left > 0) {                                             // last strip: keep the pad columns at zero
            yp[0] = acc[0];
            if (left > 1) yp[1] = acc[1];
            if (W == 4 && left > 2) yp[2] = acc[W - 2];
        }
    };
    if (NH > 0) {
        // large block: one strip per tile, straight-line gathers
        constexpr int NHc = NH > 0 ? NH : 1;
        LeanThread<NHc> L;
        L.init(B, D, A0, LT);
        if (!L.active) return;
        const int64_t left = dim_up - cbase;
        const int D0 = D[0], D1 = D[1], NY = L.NY, O = L.O;
        int i1 = L.ty, i2 = 0;
        if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
        const int e0 = L.ty * D0 + L.i0;
        const int estep = NY * D0;
        uint32_t a0 = s_in_addr + (uint32_t)e0 * 16u;
        double *yp = ys + cbase + (int64_t)e0 * ldv;
        const int64_t ystep = (int64_t)estep * ldv;
        const double zero[W] = {};
#pragma unroll 2
        for (int o = L.ty; o < O; o += NY) {
            double acc[W], es;
            uint32_t dgo;
            lean_element<NORB, NHc, W>(L, LT, A0, maxD, a0, plane, i1, i2, zero, acc, es, dgo);
            store(yp, left, acc);
            yp += ystep;
            a0 += (uint32_t)estep * 16u;
            i1 += NY;
            if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
        }
        return;
    }
    tile_pass<NORB, W, false, 1, int>(B, D, A0, T, maxD, P.H, s_in_addr, plane, SP,
        [&](int, int) { return 0; },
        [&](int q, int e, uint32_t, int, uint32_t, double, double (&acc)[W]) {
            const int64_t c0 = cbase + W * q;
            store(ys + (int64_t)e * ldv + c0, dim_up - c0, acc);
        });
}

// ------------------------------------------------------------------------------------------------------------
// Down pass, tensor-map (TMA) version:  y[blk rows][strip] = H_dw x.
// A tile = the rows of one down-block x one strip of 4 columns = R segments of 32 bytes, ld*8 bytes apart.  The
// LSU handles such a pattern one 32-byte sector per wavefront; here the copy engine gathers the tile
// (cp.async.bulk.tensor.2d, boxes of 4 columns x BR rows of a 2-D tensor map over the vector) straight into a
// dense [row][4] shared-memory image, a producer warp keeps `nstage` tiles in flight (and prefetches the tile after
// into L2 when there is only one buffer), and 16 consumer warps gather from the image.
// Thread (row, half): each thread owns 2 of the 4 columns of a row, so a quarter-warp of an LDS.128 touches
// 4 consecutive rows x 32 bytes = one conflict-free 128-byte wavefront without any swizzle.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *tm, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1) : "memory");
}

template <int NORB, int NH>
__global__ void __launch_bounds__(kNT3)
k_star_dw3(const __grid_constant__ CUtensorMap tmx, StarKParams P, int64_t dim_up, int64_t ld, int block_index, int BR, int nstage,
           const StarBlock *__restrict__ blocks,
           const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc, const double *__restrict__ hopv,
           double *__restrict__ y, int maxD)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const StarBlock B = blocks[block_index];
    const int R = B.size, tid = threadIdx.x;
    const uint32_t tileb = ((uint32_t)R * 32u + 127u) & ~127u;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)nstage * tileb);   // full[2], done[2]
    unsigned char *tab_base = reinterpret_cast<unsigned char *>(s_bar + 4);
    int D[3] = {1, 1, 1}, A0[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < NORB; a++) { D[a] = P.D[B.n[a]]; A0[a] = P.A0[B.n[a]]; }
    const LeanTabs LT = load_lean_tabs<NORB, NH>(P, B, D, A0, hopd, hopc, hopv, nullptr, tab_base, maxD, 32);
    const uint32_t tile0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_bar);
    if (tid == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
        mbar_init(bar0 + 16, kNT / 32); mbar_init(bar0 + 24, kNT / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");                         // see k_star_up3
    const int64_t ntiles = (dim_up + 3) / 4;

    if (tid >= kNT) {
        // ---------------- producer warp ----------------
        if (tid == kNT) {
            const int nbox = R / BR;
            int i = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, i++) {
                const int st = i % nstage, k = i / nstage;
                if (i >= nstage) mbar_wait(bar0 + 16 + 8 * st, (uint32_t)(k - 1) & 1u);      // consumers left the buffer
                const uint32_t fb = bar0 + 8 * st;
                mbar_expect_tx(fb, (uint32_t)R * 32u);
                const uint32_t dst = tile0 + (uint32_t)st * tileb;
                for (int b = 0; b < nbox; b++) tma_load_2d(dst + (uint32_t)b * (uint32_t)BR * 32u, &tmx, (int)(4 * t), B.off + b * BR, fb);
                const int64_t tn = t + (int64_t)nstage * gridDim.x;                          // the tile that follows in this buffer
                if (tn < ntiles)
                    for (int b = 0; b < nbox; b++) tma_prefetch_2d(&tmx, (int)(4 * tn), B.off + b * BR);
            }
        }
        return;
    }

    // ---------------- consumer warps: thread = (row slot, column half) ----------------
    const int half = tid & 1, ri = tid >> 1;
    const int D0 = D[0], D1 = D[1];
    const int ty = (D0 == 1) ? ri : (int)__umulhi((uint32_t)ri, B.magic0);
    const int i0 = ri - ty * D0;
    const int NY = (kNT / 2) / D0, O = B.nouter;
    const bool active = ty < NY && ty < O;
    // star-0 hop list of the thread in registers
    double rval[NH];
    int roff[NH], cnt0 = 0;
    double s0 = 1.0;
    {
        const int ii = active ? i0 : 0;
#pragma unroll
        for (int h = 0; h < NH; h++) {
            const double2 w = lds128(LT.ent + (uint32_t)(ii * NH + h) * 16u);
            rval[h] = w.x; roff[h] = __double2loint(w.y);
            if (h == 0) cnt0 = __double2hiint(w.y);
        }
        s0 = lds128(LT.aux + (uint32_t)ii * 16u).y;
    }
    const int estep = NY * D0;
    int i = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, i++) {
        const int st = i % nstage, k = i / nstage;
        mbar_wait(bar0 + 8 * st, (uint32_t)k & 1u);
        if (active) {
            const uint32_t tb = tile0 + (uint32_t)st * tileb + (uint32_t)half * 16u;
            int i1 = ty, i2 = 0;
            if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
            const int e0 = ty * D0 + i0;
            uint32_t a0 = tb + (uint32_t)e0 * 32u;
            double *yp = y + ((int64_t)B.off + e0) * ld + 4 * t + 2 * half;
            const int64_t ystep = (int64_t)estep * ld;
#pragma unroll 2
            for (int o = ty; o < O; o += NY) {
                double sg0 = 1.0, sg1 = s0, sg2 = s0;
                if (NORB >= 2) { const double s1 = lds128(LT.aux + (uint32_t)(maxD + i1) * 16u).y; sg0 = s1; sg2 *= s1; }
                if (NORB >= 3) { const double s2 = lds128(LT.aux + (uint32_t)(2 * maxD + i2) * 16u).y; sg0 *= s2; sg1 *= s2; }
                double p0 = 0.0, p1 = 0.0;
#pragma unroll
                for (int h = 0; h < NH; h++) {
                    if (h < cnt0) {
                        const double2 u = lds128(a0 + (uint32_t)roff[h]);
                        p0 = fma(rval[h], u.x, p0); p1 = fma(rval[h], u.y, p1);
                    }
                }
                double acc[2] = {sg0 * p0, sg0 * p1};
                if (NORB >= 2) lean_star<NH, 2, false>(acc, a0, 0u, LT.ent + (uint32_t)((maxD + i1) * NH) * 16u, sg1);
                if (NORB >= 3) lean_star<NH, 2, false>(acc, a0, 0u, LT.ent + (uint32_t)((2 * maxD + i2) * NH) * 16u, sg2);
                *reinterpret_cast<double2 *>(yp) = make_double2(acc[0], acc[1]);     // pad columns receive exact zeros
                yp += ystep;
                a0 += (uint32_t)estep * 32u;
                i1 += NY;
                if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(bar0 + 16 + 8 * st);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Fringe kernels: the blocks smaller than kBulkMin hold ~1 % of a large sector but would cost one launch each
// (~20 us apiece for no work); all of them go through ONE thread-per-element launch per pass that reads the ELL hop
// tables of the generic kernel (hxv_generic.cu) straight from global memory / L2.
//   k_fringe_dw: y[rd][c]  = sum_j ampD_j x[tgtD_j(rd)][c]                       rd in the fringe rows, all columns
//   k_fringe_up: y[r][ru] (+)= diag x[r][ru] + sum_j ampU_j x[r][tgtU_j(ru)]      ru in the fringe columns, all rows
__global__ void __launch_bounds__(256)
k_fringe_dw(int nfr, const int *__restrict__ fringe, int64_t ncols, int64_t ld, int64_t dim_dw,
            const uint32_t *__restrict__ hop_dw, const uint8_t *__restrict__ nhop_dw, const double *__restrict__ amp_dw,
            const double *__restrict__ x, double *__restrict__ y)
{
    __shared__ double s_amp[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_amp[i] = amp_dw[i];
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    for (int q = blockIdx.y; q < nfr; q += gridDim.y) {
        const int64_t rd = fringe[q];
        const int nd = nhop_dw[rd];
        double acc = 0.0;
        for (int j = 0; j < nd; j++) {
            const uint32_t h = hop_dw[(int64_t)j * dim_dw + rd];
            acc += s_amp[h & 255u] * x[(int64_t)(h >> 8) * ld + c];
        }
        y[rd * ld + c] = acc;
    }
}

template <bool SLABBED>
__global__ void __launch_bounds__(256)
k_fringe_up(int nfr, const int *__restrict__ fringe, int64_t nrows, int64_t ld, int64_t dim_up, int norb, int accumulate,
            const uint32_t *__restrict__ cfg_up, const uint32_t *__restrict__ cfg_dw,
            const double *__restrict__ e_up, const double *__restrict__ e_dw, const double *__restrict__ xtab,
            const uint32_t *__restrict__ hop_up, const uint8_t *__restrict__ nhop_up, const double *__restrict__ amp_up,
            const double *__restrict__ x, double *__restrict__ y, double *__restrict__ dot_out, SlabMap Mpar)
{
    __shared__ double s_amp[256], s_x[32 * 32], s_red[8];
    __shared__ SlabMap M;
    if (SLABBED && threadIdx.x == 0) M = Mpar;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_amp[i] = amp_up[i];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_x[i] = xtab[i];
    __syncthreads();
    const uint32_t impmask = (1u << norb) - 1u;
    const int64_t total = nrows * nfr;
    double dsum = 0.0;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / nfr;
        const int q = (int)(idx - r * nfr);
        const int64_t ru = fringe[q];
        // element (r, c) of the row shard: plain [nrows][ld] tile, or slab / peer addressing (see SlabMap)
        auto xat = [&](int64_t c) -> const double * {
            if (!SLABBED) return x + r * ld + c;
            const int p = slab_of(M, (int)c);
            return (M.peer ? M.xp[p] + M.xbase[p] : x + M.base[p]) + r * M.ldc[p] + (c - M.col0[p]);
        };
        const double xo = *xat(ru);
        double acc = (e_up[ru] + e_dw[r] + s_x[(cfg_dw[r] & impmask) * 32u + (cfg_up[ru] & impmask)]) * xo;
        double *yo;
        if (!SLABBED) yo = y + r * ld + ru;
        else { const int p = slab_of(M, (int)ru); yo = (M.peer ? M.yp[p] : y) + M.base[p] + r * M.ldc[p] + (ru - M.col0[p]); }
        if (accumulate) acc += *yo;
        const int nu = nhop_up[ru];
        for (int j = 0; j < nu; j++) {
            const uint32_t h = hop_up[(int64_t)j * dim_up + ru];
            acc += s_amp[h & 255u] * *xat(h >> 8);
        }
        *yo = acc;
        dsum = fma(xo, acc, dsum);
    }
    if (dot_out) {
        for (int o = 16; o > 0; o >>= 1) dsum += __shfl_down_sync(0xffffffffu, dsum, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = 0.0;
            for (int w = 0; w < 8; w++) v += s_red[w];
            dot_out[blockIdx.x] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
static void fill_kparams(const StarInfo &S, StarKParams &P)
{
    memset(&P, 0, sizeof(P));
    P.norb = S.norb; P.H = S.H; P.ncfg = S.ncfg; P.pair_e = S.pair_e;
    for (int m = 0; m < 16; m++) { P.D[m] = S.D[m]; P.A0[m] = S.A0[m]; P.coff[m] = S.coff[m]; }
}

static constexpr int kStageElems = 4900;     // elements (x 16 B) per pipeline stage of the up pass / per tile of the down pass

// hop lists of the lean kernels are padded to one of these lengths (0: no lean kernel, generic tile_pass)
static int round_nh(int nh) { return nh <= 4 ? 4 : nh <= 5 ? 5 : nh <= 6 ? 6 : nh <= 8 ? 8 : 0; }

static int ensure_smem(edgpu_ctx *ctx, const void *kern, size_t smem)
{
    static std::map<const void *, size_t> set;
    static std::mutex mtx;                                   // contexts of several host threads share the function attributes
    std::lock_guard<std::mutex> lock(mtx);
    size_t &cur = set[kern];
    if (smem > cur) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    return 0;
}

// 2-D tensor map over a [rows][ld] fp64 tile for the strip loads of the down pass: boxes of 4 columns x BR rows
static int encode_strip_map(edgpu_ctx *ctx, CUtensorMap *tm, const double *base, int64_t ld, int64_t rows, int BR)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_TRY(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) return edgpu_fail(ctx, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeFn>(p);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {4, (cuuint32_t)BR};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return edgpu_fail(ctx, "cuTensorMapEncodeTiled failed (%d) for ld=%lld rows=%lld box=4x%d", (int)r, (long long)ld, (long long)rows, BR);
    return 0;
}

using Dw3Kernel = void (*)(const CUtensorMap, StarKParams, int64_t, int64_t, int, int, int, const StarBlock *, const int16_t *, const uint8_t *,
                           const double *, double *, int);
template <int NORB>
static Dw3Kernel pick_dw3(int NH)
{
    switch (NH) {
        case 4: return k_star_dw3<NORB, 4>;
        case 5: return k_star_dw3<NORB, 5>;
        case 6: return k_star_dw3<NORB, 6>;
        default: return k_star_dw3<NORB, 8>;
    }
}

// launch configuration with (optional) programmatic stream serialisation: the kernel may start (prologue only, up to its
// griddepcontrol.wait) while the previous launch on the stream drains
static void pdl_config(edgpu_ctx *ctx, cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, unsigned grid, unsigned block, size_t smem)
{
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    // measured on cfg4 (same box, A/B): 491 matvec/s with, 511 without -- the early-resident CTAs of the next launch cost
    // more than the hidden prologue saves, so it is OFF unless test-hook bit 5 asks for it
    attr[0].val.programmaticStreamSerializationAllowed = (ctx->par.reserved[0] & 32) ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
}

using DwKernel = void (*)(StarKParams, int64_t, int64_t, int, int, const StarBlock *, const int16_t *, const uint8_t *, const double *,
                          const double *, double *, int, int);
using UpKernel = void (*)(StarKParams, SlabMap, int64_t, int64_t, int, int, int, int, const StarBlock *, const int16_t *, const uint8_t *,
                          const double *, const double *, const double *, const uint32_t *, const double *, const double *, double *, int);

template <int NORB>
static DwKernel pick_dw(int W, int NH)
{
    if (W == 4) switch (NH) {
        case 4: return k_star_dw<NORB, 4, 4>;
        case 5: return k_star_dw<NORB, 4, 5>;
        case 6: return k_star_dw<NORB, 4, 6>;
        case 8: return k_star_dw<NORB, 4, 8>;
        default: return k_star_dw<NORB, 4, 0>;
    }
    switch (NH) {
        case 4: return k_star_dw<NORB, 2, 4>;
        case 5: return k_star_dw<NORB, 2, 5>;
        case 6: return k_star_dw<NORB, 2, 6>;
        case 8: return k_star_dw<NORB, 2, 8>;
        default: return k_star_dw<NORB, 2, 0>;
    }
}

using Up3Kernel = void (*)(StarKParams, SlabMap, int, int64_t, int64_t, int, int, int, int, const StarBlock *, const int16_t *, const uint8_t *, const double *,
                           const double *, const double *, const uint32_t *, const double *, const double *, double *, int, double *);
template <int NORB>
static Up3Kernel pick_up3(int NH, bool gen)
{
    if (gen) switch (NH) {
        case 4: return k_star_up3<NORB, 4, true>;
        case 5: return k_star_up3<NORB, 5, true>;
        case 6: return k_star_up3<NORB, 6, true>;
        default: return k_star_up3<NORB, 8, true>;
    }
    switch (NH) {
        case 4: return k_star_up3<NORB, 4, false>;
        case 5: return k_star_up3<NORB, 5, false>;
        case 6: return k_star_up3<NORB, 6, false>;
        default: return k_star_up3<NORB, 8, false>;
    }
}

template <int NORB, bool SLAB>
static UpKernel pick_up(int NH)
{
    switch (NH) {
        case 4: return k_star_up<NORB, SLAB, 4>;
        case 5: return k_star_up<NORB, SLAB, 5>;
        case 6: return k_star_up<NORB, SLAB, 6>;
        case 8: return k_star_up<NORB, SLAB, 8>;
        default: return k_star_up<NORB, SLAB, 0>;
    }
}

// Down pass on a column range: x, y point at a [DimDw][ld] tile holding `ncols` up-spin columns (the whole sector
// vector on one GPU, or the column shard of one rank -- hops of the down spin never change the column).
template <int NORB>
static int launch_star_dw(edgpu_sector *s, const double *x, double *y, int64_t ncols, int64_t ld)
{
    edgpu_ctx *ctx = s->ctx;
    const StarInfo &Dn = *s->dw->star;
    StarKParams PD;
    fill_kparams(Dn, PD);
    int maxD = 0;
    for (int m = 0; m <= Dn.nbath + 1; m++) maxD = std::max(maxD, Dn.D[m]);
    if (maxD > kNT) return edgpu_fail(ctx, "star kernels: star dimension %d exceeds %d threads", maxD, kNT);
    const bool force_narrow = (ctx->par.reserved[0] & 1) != 0;                 // test hook: exercise the 2-column path
    const bool force_generic = (ctx->par.reserved[0] & 4) != 0;                // test hook: generic tile_pass everywhere
    const bool fringe = !(ctx->par.reserved[0] & (4 | 8 | 16)) && Dn.nfringe > 0;
    if (fringe) {
        dim3 grid((unsigned)((ncols + 255) / 256), (unsigned)std::min(Dn.nfringe, 4096));
        k_fringe_dw<<<grid, 256, 0, ctx->stream>>>(Dn.nfringe, Dn.d_fringe, ncols, ld, s->dim_dw, s->dw->hop, s->dw->nhop, s->dw->amp, x, y);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    for (size_t bi = 0; bi < Dn.blocks.size(); bi++) {                         // one launch per down-block
        const StarBlock &B = Dn.blocks[bi];
        if (fringe && B.size < kBulkMin) continue;
        // strips per tile: up to kStageElems rows (x 32 B) of shared memory, but keep >= 4 CTAs per SM worth of tiles
        int64_t SP = std::max<int64_t>(1, kStageElems / B.size);
        SP = std::max<int64_t>(1, std::min<int64_t>(SP, ((ncols + 3) / 4) / (4 * (int64_t)ctx->sm_count)));
        int bD = 1;                                                            // table stride: largest star of THIS block
        for (int a = 0; a < NORB; a++) bD = std::max(bD, Dn.D[B.n[a]]);
        if (!force_generic && !(ctx->par.reserved[0] & 8) && B.size >= ((ctx->par.reserved[0] & 16) ? 4 : kBulkMin) && round_nh(B.nh) &&
            B.size % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && ld % 2 == 0) {
            // tensor-map pipeline
            int BR = 0;
            for (int d = 4; d <= 256 && d <= B.size; d += 4)
                if (B.size % d == 0) BR = d;
            const int NH3 = round_nh(B.nh);
            const size_t tileb = ((size_t)B.size * 32 + 127) & ~(size_t)127;
            const size_t tabs = lean_tabs_bytes(NORB, bD, NH3) + 64;
            const int nstage = (2 * tileb + tabs <= 227 * 1024) ? 2 : 1;
            const size_t smem = nstage * tileb + tabs;
            if (BR >= 16 && smem <= 227 * 1024) {
                CUtensorMap tm;
                if (int rc = encode_strip_map(ctx, &tm, x, ld, s->dim_dw, BR)) return rc;
                auto kern = pick_dw3<NORB>(NH3);
                if (int rc = ensure_smem(ctx, (const void *)kern, smem)) return rc;
                const int64_t ntiles = (ncols + 3) / 4;
                const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (227 * 1024) / (smem + 1024)));
                const unsigned nctas = (unsigned)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * per_sm);
                cudaLaunchConfig_t cfg;
                cudaLaunchAttribute attr[1];
                pdl_config(ctx, cfg, attr, nctas, kNT3, smem);
                CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, tm, PD, ncols, ld, (int)bi, BR, nstage, (const StarBlock *)Dn.d_blocks,
                                                 (const int16_t *)Dn.d_hopd, (const uint8_t *)Dn.d_hopc, (const double *)Dn.d_hopv, y, bD));
                continue;
            }
        }
        const int NH = (SP == 1 && !force_generic) ? round_nh(B.nh) : 0;
        const size_t tab = NH ? lean_tabs_bytes(NORB, bD, NH) : tabs_bytes(NORB, bD, Dn.H);
        const int W = (!force_narrow && sizeof(double) * (size_t)B.size * 4 * SP + tab <= 227 * 1024) ? 4 : 2;
        const int64_t nstrips = (ncols + W - 1) / W;
        const size_t smem = sizeof(double) * (size_t)B.size * W * SP + tab;
        if (smem > 227 * 1024) return edgpu_fail(ctx, "star down pass: block of %d rows does not fit in shared memory", B.size);
        auto kern = pick_dw<NORB>(W, NH);
        if (int rc = ensure_smem(ctx, (const void *)kern, smem)) return rc;
        const unsigned nctas = (unsigned)((nstrips + SP - 1) / SP);
        kern<<<nctas, kNT, smem, ctx->stream>>>(PD, ncols, ld, (int)bi, (int)SP, Dn.d_blocks, Dn.d_hopd, Dn.d_hopc, Dn.d_hopv, x, y, bD,
                                                 ctx->sm_count * (smem * 2 + 2048 <= 227 * 1024 ? 2 : 1));
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return 0;
}

// Up pass on a row range: x, y point at a [nrows][ld] tile holding ALL up-spin columns of down-spin rows
// [row0, row0+nrows) (the whole vector, or the row shard of one rank after the transpose).
template <int NORB>
// dot != nullptr: the kernels also leave per-CTA partial sums of <x, y_final> in dot[0 .. *ndot); *ndot = -1 when a block
// went through a kernel without that epilogue (the caller then computes the dot product separately)
static int launch_star_up(edgpu_sector *s, const double *x, double *y, int64_t row0, int64_t nrows, int64_t ld, int accumulate,
                          const SlabMap *slabs = nullptr, double *dot = nullptr, int *ndot = nullptr)
{
    int slots = 0;
    bool dot_ok = dot != nullptr;
    edgpu_ctx *ctx = s->ctx;
    const StarInfo &U = *s->up->star;
    StarKParams PU;
    fill_kparams(U, PU);
    int maxD = 0;
    for (int m = 0; m <= U.nbath + 1; m++) maxD = std::max(maxD, U.D[m]);
    if (maxD > kNT) return edgpu_fail(ctx, "star kernels: star dimension %d exceeds %d threads", maxD, kNT);
    const bool force_generic = (ctx->par.reserved[0] & 4) != 0;
    SlabMap M;
    memset(&M, 0, sizeof(M));
    if (slabs) M = *slabs; else { M.n = 1; M.ldc0 = (int)ld; M.magic = 0; M.ldc[0] = (int)ld; }
    const int64_t npairs = (nrows + 1) / 2;
    const bool fringe = !(ctx->par.reserved[0] & (4 | 8 | 16)) && U.nfringe > 0;
    if (fringe) {
        const int64_t total = nrows * U.nfringe;
        const unsigned nb = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 32);
        const bool d = dot_ok && slots + (int)nb <= kDotSlots;
        auto fk = slabs ? k_fringe_up<true> : k_fringe_up<false>;
        fk<<<nb, 256, 0, ctx->stream>>>(U.nfringe, U.d_fringe, nrows, ld, s->dim_up, NORB, accumulate, s->up->cfg, s->dw->cfg + row0,
                                        s->up->ediag, s->dw->ediag + row0, ctx->d_xtab, s->up->hop, s->up->nhop, s->up->amp, x, y,
                                        d ? dot + slots : nullptr, M);
        CUDA_TRY(ctx, cudaGetLastError());
        if (d) slots += (int)nb; else dot_ok = false;
    }
    for (size_t bi = 0; bi < U.blocks.size(); bi++) {                          // persistent kernel, one launch per up-block
        const StarBlock &B = U.blocks[bi];
        if (fringe && B.size < kBulkMin) continue;
        int64_t RP = std::max<int64_t>(1, kStageElems / B.size);
        RP = std::max<int64_t>(1, std::min<int64_t>(RP, npairs / (2 * (int64_t)ctx->sm_count)));
        int bD = 1;
        for (int a = 0; a < NORB; a++) bD = std::max(bD, U.D[B.n[a]]);
        if (!force_generic && !(ctx->par.reserved[0] & 8) && round_nh(B.nh) &&
            (B.size >= ((ctx->par.reserved[0] & 16) ? 1 : kBulkMin) || (slabs && slabs->peer)) &&
            ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
            // bulk-copy pipeline: G rows per tile
            const int NH3 = round_nh(B.nh);
            int64_t G = std::max<int64_t>(1, kStageElems / B.size);
            G = std::max<int64_t>(1, std::min<int64_t>(G, nrows / (2 * (int64_t)ctx->sm_count)));
            const int lead = B.off & 1, ncopy = (lead + B.size + 1) & ~1;
            const size_t rowb = (size_t)(ncopy + 2) * 8;
            const size_t tabs = lean_tabs_bytes(NORB, bD, NH3) + 64;
            // stages: 2 y images when they fit; 3 x images (measured 518 vs 512 matvec/s for 2, no gain beyond), as many as
            // shared memory allows when x comes over NVLink; test-hook bit 9: 2 x images (the compile-time variant)
            const size_t avail = 227 * 1024 - 1024 - tabs;
            const size_t per = rowb * G + 64 * G;                                  // one stage (+ its diagonal terms)
            int nst = (int)std::min<size_t>(avail / per, (size_t)kMaxXS + 2);
            if (ctx->par.reserved[0] & 64) nst = std::min(nst, (ctx->par.reserved[0] & 128) ? 3 : 2);   // test hooks: few stages
            const int nys = nst >= 4 ? 2 : 1;
            const int nxs = std::max(1, std::min(nst - nys, (slabs && slabs->peer) ? kMaxXS : ((ctx->par.reserved[0] & 512) ? 2 : 3)));
            const size_t smem = (size_t)(nxs + nys) * rowb * G + sizeof(double) * 8 * G * nxs + tabs;
            if (nst >= 2 && smem <= 227 * 1024) {
                auto kern = pick_up3<NORB>(NH3, !(nxs == 2 && nys == 2));
                if (int rc = ensure_smem(ctx, (const void *)kern, smem)) return rc;
                const int64_t ntiles = (nrows + G - 1) / G;
                const unsigned nctas = (unsigned)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count);
                const bool d = dot_ok && slots + (int)nctas <= kDotSlots;
                cudaLaunchConfig_t cfg;
                cudaLaunchAttribute attr[1];
                pdl_config(ctx, cfg, attr, nctas, kNT3, smem);
                CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, PU, M, accumulate, nrows, ld, (int)bi, (int)G, nxs, nys, (const StarBlock *)U.d_blocks,
                                                 (const int16_t *)U.d_hopd, (const uint8_t *)U.d_hopc, (const double *)U.d_hopv,
                                                 (const double *)U.d_estar, (const double *)(s->dw->ediag + row0),
                                                 (const uint32_t *)(s->dw->cfg + row0), (const double *)ctx->d_xtab, x, y, bD,
                                                 d ? dot + slots : (double *)nullptr));
                if (d) slots += (int)nctas; else dot_ok = false;
                continue;
            }
        }
        if (slabs && slabs->peer)
            return edgpu_fail(ctx, "star up pass (peer mode): no copy-engine kernel for the block of %d configurations", B.size);
        const int NH = (RP == 1 && !force_generic) ? round_nh(B.nh) : 0;
        const size_t tab = NH ? lean_tabs_bytes(NORB, bD, NH) : tabs_bytes(NORB, bD, U.H);
        const size_t smem1 = sizeof(double) * ((size_t)2 * B.size * RP + (size_t)16 * RP) + tab;
        // double-buffered when two stages fit (nstage 1: single stage, e.g. 8000-configuration blocks of Norb=3, Nbath=5)
        const size_t smem2 = sizeof(double) * ((size_t)4 * B.size * RP + (size_t)32 * RP) + tab;
        const int nstage = (smem2 <= 227 * 1024 && !(ctx->par.reserved[0] & 2)) ? 2 : 1;
        const size_t smem = nstage == 1 ? smem1 : smem2;
        if (smem > 227 * 1024) return edgpu_fail(ctx, "star up pass: block of %d configurations does not fit in shared memory", B.size);
        auto kern = slabs ? pick_up<NORB, true>(NH) : pick_up<NORB, false>(NH);
        if (int rc = ensure_smem(ctx, (const void *)kern, smem)) return rc;
        const int64_t ntiles = (npairs + RP - 1) / RP;
        const unsigned nctas = (unsigned)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count);
        kern<<<nctas, kNT, smem, ctx->stream>>>(PU, M, nrows, ld, (int)bi, (int)RP, accumulate, nstage, U.d_blocks, U.d_hopd, U.d_hopc,
                                                          U.d_hopv, U.d_estar, s->dw->ediag + row0, s->dw->cfg + row0, ctx->d_xtab, x, y, bD);
        CUDA_TRY(ctx, cudaGetLastError());
        dot_ok = false;
    }
    if (ndot) *ndot = dot_ok ? slots : -1;
    return 0;
}

int hxv_star_dw(edgpu_sector *s, const double *x, double *y, int64_t ncols, int64_t ld)
{
    if (!s->up->star || !s->dw->star) return edgpu_fail(s->ctx, "hxv_star: sector is not in the star-product layout");
    switch (s->ctx->ham.norb) {
    case 1: return launch_star_dw<1>(s, x, y, ncols, ld);
    case 2: return launch_star_dw<2>(s, x, y, ncols, ld);
    case 3: return launch_star_dw<3>(s, x, y, ncols, ld);
    default: return edgpu_fail(s->ctx, "star kernels: Norb=%d unsupported", s->ctx->ham.norb);
    }
}

// Up pass on a row shard stored as per-source-rank slabs (x and y share the slab layout): nslab slabs, slab p holds
// columns [col0[p], col0[p]+ldc[p]) of the nrows rows, [nrows][ldc[p]] each, packed back to back.
int hxv_star_up_slabs(edgpu_sector *s, const double *x, double *y, int64_t row0, int64_t nrows, int nslab,
                      const int64_t *col0, const int64_t *ldc, int accumulate)
{
    if (!s->up->star || !s->dw->star) return edgpu_fail(s->ctx, "hxv_star: sector is not in the star-product layout");
    if (nslab < 1 || nslab > 8) return edgpu_fail(s->ctx, "hxv_star_up_slabs: 1..8 slabs supported (got %d)", nslab);
    SlabMap M;
    memset(&M, 0, sizeof(M));
    M.n = nslab;
    M.ldc0 = (int)ldc[0];
    if (M.ldc0 <= 0) return edgpu_fail(s->ctx, "hxv_star_up_slabs: empty first slab");
    M.magic = (unsigned)(((1ull << 32) + (uint64_t)M.ldc0 - 1) / (uint64_t)M.ldc0);
    if (M.ldc0 == 1) return edgpu_fail(s->ctx, "hxv_star_up_slabs: slab width must be > 1");
    long long base = 0;
    for (int p = 0; p < nslab; p++) {
        if (p < nslab - 1 && ldc[p] != ldc[0]) return edgpu_fail(s->ctx, "hxv_star_up_slabs: all but the last slab must have equal width");
        if (col0[p] != (int64_t)p * ldc[0]) return edgpu_fail(s->ctx, "hxv_star_up_slabs: slabs must tile the columns in order");
        M.col0[p] = (int)col0[p]; M.ldc[p] = (int)ldc[p]; M.base[p] = base;
        base += (long long)nrows * ldc[p];
    }
    switch (s->ctx->ham.norb) {
    case 1: return launch_star_up<1>(s, x, y, row0, nrows, s->ld, accumulate, &M);
    case 2: return launch_star_up<2>(s, x, y, row0, nrows, s->ld, accumulate, &M);
    case 3: return launch_star_up<3>(s, x, y, row0, nrows, s->ld, accumulate, &M);
    default: return edgpu_fail(s->ctx, "star kernels: Norb=%d unsupported", s->ctx->ham.norb);
    }
}

// Up pass in peer mode: this rank's rows [row0, row0+nrows) of ALL columns, where the columns [col0[p], col0[p]+ldc[p]) live in
// the column shard of rank p ([dim_dw][ldc[p]], base pointers xp[p] / yp[p] valid on this device, e.g. CUDA IPC mappings).
int hxv_star_up_peers(edgpu_sector *s, const double *const *xp, double *const *yp, int64_t row0, int64_t nrows, int nslab,
                      const int64_t *col0, const int64_t *ldc, int accumulate, const int64_t *x_row0)
{
    if (!s->up->star || !s->dw->star) return edgpu_fail(s->ctx, "hxv_star: sector is not in the star-product layout");
    if (nslab < 1 || nslab > 8) return edgpu_fail(s->ctx, "hxv_star_up_peers: 1..8 ranks supported (got %d)", nslab);
    SlabMap M;
    memset(&M, 0, sizeof(M));
    M.n = nslab;
    M.peer = 1;
    M.ldc0 = (int)ldc[0];
    if (M.ldc0 <= 1) return edgpu_fail(s->ctx, "hxv_star_up_peers: column shards must be wider than 1");
    M.magic = (unsigned)(((1ull << 32) + (uint64_t)M.ldc0 - 1) / (uint64_t)M.ldc0);
    for (int p = 0; p < nslab; p++) {
        if (p < nslab - 1 && ldc[p] != ldc[0]) return edgpu_fail(s->ctx, "hxv_star_up_peers: all but the last shard must have equal width");
        if (col0[p] != (int64_t)p * ldc[0] || ldc[p] % 4) return edgpu_fail(s->ctx, "hxv_star_up_peers: shards must tile the columns in strips of 4");
        if ((reinterpret_cast<uintptr_t>(xp[p]) | reinterpret_cast<uintptr_t>(yp[p])) & 15)
            return edgpu_fail(s->ctx, "hxv_star_up_peers: shard pointers must be 16-byte aligned");
        M.col0[p] = (int)col0[p]; M.ldc[p] = (int)ldc[p]; M.base[p] = (long long)row0 * ldc[p];
        M.xbase[p] = (long long)(row0 - (x_row0 ? x_row0[p] : 0)) * ldc[p];
        M.xp[p] = xp[p]; M.yp[p] = yp[p];
    }
    switch (s->ctx->ham.norb) {
    case 1: return launch_star_up<1>(s, nullptr, nullptr, row0, nrows, s->ld, accumulate, &M);
    case 2: return launch_star_up<2>(s, nullptr, nullptr, row0, nrows, s->ld, accumulate, &M);
    case 3: return launch_star_up<3>(s, nullptr, nullptr, row0, nrows, s->ld, accumulate, &M);
    default: return edgpu_fail(s->ctx, "star kernels: Norb=%d unsupported", s->ctx->ham.norb);
    }
}

int hxv_star_up(edgpu_sector *s, const double *x, double *y, int64_t row0, int64_t nrows, int64_t ld, int accumulate)
{
    if (!s->up->star || !s->dw->star) return edgpu_fail(s->ctx, "hxv_star: sector is not in the star-product layout");
    switch (s->ctx->ham.norb) {
    case 1: return launch_star_up<1>(s, x, y, row0, nrows, ld, accumulate);
    case 2: return launch_star_up<2>(s, x, y, row0, nrows, ld, accumulate);
    case 3: return launch_star_up<3>(s, x, y, row0, nrows, ld, accumulate);
    default: return edgpu_fail(s->ctx, "star kernels: Norb=%d unsupported", s->ctx->ham.norb);
    }
}

// y = H x on one GPU: down pass first (y = H_dw x, write only), then the up pass accumulates (diag + H_up) x.
int hxv_star(edgpu_sector *s, const double *x, double *y)
{
    if (int rc = hxv_star_dw(s, x, y, s->dim_up, s->ld)) return rc;
    return hxv_star_up(s, x, y, 0, s->dim_dw, s->ld, 1);
}

// y = H x and, fused into the up pass, per-CTA partial sums of <x, y> in dot[0 .. *ndot) (Lanczos alpha);
// *ndot = -1: not produced (some block used a kernel without the epilogue)
int hxv_star_dot(edgpu_sector *s, const double *x, double *y, double *dot, int *ndot)
{
    if (int rc = hxv_star_dw(s, x, y, s->dim_up, s->ld)) return rc;
    switch (s->ctx->ham.norb) {
    case 1: return launch_star_up<1>(s, x, y, 0, s->dim_dw, s->ld, 1, nullptr, dot, ndot);
    case 2: return launch_star_up<2>(s, x, y, 0, s->dim_dw, s->ld, 1, nullptr, dot, ndot);
    case 3: return launch_star_up<3>(s, x, y, 0, s->dim_dw, s->ld, 1, nullptr, dot, ndot);
    default: return edgpu_fail(s->ctx, "star kernels: Norb=%d unsupported", s->ctx->ham.norb);
    }
}

int hxv_star_launches(const edgpu_sector *s)
{
    int n = 0;
    for (const StarInfo *S : {s->up->star.get(), s->dw->star.get()}) {
        const bool fringe = !(s->ctx->par.reserved[0] & (4 | 8 | 16)) && S->nfringe > 0;
        for (const StarBlock &B : S->blocks) n += (fringe && B.size < kBulkMin) ? 0 : 1;
        n += fringe ? 1 : 0;
    }
    return n;
}
