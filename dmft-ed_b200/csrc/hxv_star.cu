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
#include "edgpu_internal.h"
#include <algorithm>
#include <cstring>

uint64_t edgpu_binom(int n, int k);

static constexpr int kMaxStarCfg = 1024;      // 2^(Nbath+1), Nbath <= 9
static constexpr int kMaxH = 9;               // hops per star configuration <= Nbath
static constexpr int kMaxBlocks = 4096;
static constexpr int kBigBlock = 2048;        // up-blocks at least this large use the pipelined kernel

struct StarBlock {             // one occupation tuple of one spin
    int off;                   // first internal index
    int size;                  // prod D[n_a]
    int n[EDGPU_MAXORB];       // star occupations
    int sgn_lower[EDGPU_MAXORB];   // (-1)^{sum_{a'<a} n_a'} as 0/1
    uint32_t magic0;               // ceil(2^32 / D0): tid / D0 == __umulhi(tid, magic0) for tid < 1024
    int ny;                        // kNT / D0
    int nouter;                    // size / D0
};

struct StarInfo {
    int norb = 0, nbath = 0, H = 0, ncfg = 0;
    int D[16], A0[16], coff[16];                 // per occupation m: star dim, #imp=0 configs, offset into cfg tables
    std::vector<StarBlock> blocks;
    // device copies
    StarBlock *d_blocks = nullptr;
    uint8_t *d_hopj = nullptr;                   // [ncfg][H]   target index inside the same-occupation star list
    int16_t *d_hopd = nullptr;                   // [ncfg][H]   target index minus own index (what the tiled kernels use)
    uint8_t *d_hopc = nullptr;                   // [ncfg]
    double *d_hopv = nullptr;                    // [norb][ncfg][H] signed amplitudes V_{a,k} * (-1)^{popc(bath below k)}
    double *d_estar = nullptr;                   // [norb][ncfg]    star diagonal energies
    double pair_e = 0.0;                         // (Ust-Jh): same-spin inter-orbital density term
    // tile schedule
    int *d_upgroups = nullptr;                   // [ngroups][2] = (start in d_uplist, count): groups of SMALL blocks
    int *d_uplist = nullptr;                     // block ids of the groups
    int ngroups = 0, max_small = 0;
    std::vector<int> big_blocks;                 // blocks handled by the persistent pipelined up kernel
    int max_block = 0;
    ~StarInfo()
    {
        cudaFree(d_blocks); cudaFree(d_hopj); cudaFree(d_hopd); cudaFree(d_hopc); cudaFree(d_hopv); cudaFree(d_estar); cudaFree(d_upgroups); cudaFree(d_uplist);
    }
};

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
static constexpr int kVec = 4;          // outputs per thread and outer step (rows in the up pass, columns in the down pass)

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
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
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

// Lean single-batch pass (large blocks: one row pair / one strip per tile).  Same math as tile_pass, minimal index
// arithmetic: all addresses advance by constant strides, star-0 hop entries live in registers, and the sign of a
// star (constant over its hops) is applied once to the partial sum instead of once per hop.
//   VEC = 2: up pass (tile = one [element][2] plane);  VEC = 4: down pass (two planes `plane` bytes apart)
// pre(e) -> EARLY is issued before the gathers; f(e, early, ui, esum, acc) stores the VEC results.
template <int NORB, int VEC, bool WITH_E, class EARLY, class PRE, class F>
__device__ __forceinline__ void lean_pass(const StarBlock &B, const int *D, const int *A0, const TabPtrs &T, int maxD, int H,
                                          uint32_t s_in_addr, uint32_t plane, PRE pre, F f)
{
    const int tid = threadIdx.x, D0 = D[0];
    const int ty = (D0 == 1) ? tid : (int)__umulhi((uint32_t)tid, B.magic0);
    const int i0 = tid - ty * D0;
    const int NY = B.ny;
    if (ty >= NY) return;
    const int O = B.nouter;
    constexpr int kRegH = 8;
    const int cnt0 = T.cnt[i0];
    double rval[kRegH];
    int roff[kRegH];
    {
        const HopEnt *ent0 = T.ent + i0 * H;
#pragma unroll
        for (int h = 0; h < kRegH; h++) {
            const HopEnt en = (h < cnt0) ? ent0[h] : HopEnt{0.0, 0, 0};
            rval[h] = en.val; roff[h] = en.off;
        }
    }
    const bool imp0 = i0 >= A0[0];
    const double e0 = WITH_E ? T.e[i0] : 0.0;
    const double c0s = (B.sgn_lower[0] & 1) ? -1.0 : 1.0;
    const double c1s = ((NORB >= 2 && (B.sgn_lower[1] & 1)) ? -1.0 : 1.0) * (imp0 ? -1.0 : 1.0);     // includes s0
    const double c2s = ((NORB >= 3 && (B.sgn_lower[2] & 1)) ? -1.0 : 1.0) * (imp0 ? -1.0 : 1.0);
    const int D1 = (NORB >= 2) ? D[1] : 1;
    int i1 = ty, i2 = 0;
    if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
    int e = ty * D0 + i0;
    const int estep = NY * D0;
    uint32_t a0 = s_in_addr + (uint32_t)e * 16u;
    const uint32_t astep = (uint32_t)estep * 16u;
    EARLY nxt = pre(e);                                   // software pipeline: the loads of step o+NY are in flight during step o
#pragma unroll 2
    for (int o = ty; o < O; o += NY) {
        const EARLY early = nxt;
        if (o + NY < O) nxt = pre(e + estep);
        uint32_t ui = imp0 ? 1u : 0u;
        double es = e0, s1 = 1.0, s2 = 1.0;
        if (NORB >= 2) { const bool b = i1 >= A0[1]; ui |= b ? 2u : 0u; s1 = b ? -1.0 : 1.0; if (WITH_E) es += T.e[maxD + i1]; }
        if (NORB >= 3) { const bool b = i2 >= A0[2]; ui |= b ? 4u : 0u; s2 = b ? -1.0 : 1.0; if (WITH_E) es += T.e[2 * maxD + i2]; }
        double acc[VEC], part[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) part[v] = 0.0;
#pragma unroll
        for (int h = 0; h < kRegH; h++) {
            if (h < cnt0) {
                const uint32_t a = a0 + (uint32_t)roff[h];
                const double2 p = lds128(a);
                part[0] += rval[h] * p.x; part[1] += rval[h] * p.y;
                if (VEC == 4) { const double2 q = lds128(a + plane); part[2] += rval[h] * q.x; part[3] += rval[h] * q.y; }
            }
        }
        {
            const double sg = c0s * s1 * s2;
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] = sg * part[v];
        }
        if (NORB >= 2) {
            const HopEnt *ent = T.ent + (maxD + i1) * H;
            const int cnt = T.cnt[maxD + i1];
#pragma unroll
            for (int v = 0; v < VEC; v++) part[v] = 0.0;
#pragma unroll 4
            for (int h = 0; h < cnt; h++) {
                const HopEnt en = ent[h];
                const uint32_t a = a0 + (uint32_t)en.off;
                const double2 p = lds128(a);
                part[0] += en.val * p.x; part[1] += en.val * p.y;
                if (VEC == 4) { const double2 q = lds128(a + plane); part[2] += en.val * q.x; part[3] += en.val * q.y; }
            }
            const double sg = c1s * s2;
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] += sg * part[v];
        }
        if (NORB >= 3) {
            const HopEnt *ent = T.ent + (2 * maxD + i2) * H;
            const int cnt = T.cnt[2 * maxD + i2];
#pragma unroll
            for (int v = 0; v < VEC; v++) part[v] = 0.0;
#pragma unroll 4
            for (int h = 0; h < cnt; h++) {
                const HopEnt en = ent[h];
                const uint32_t a = a0 + (uint32_t)en.off;
                const double2 p = lds128(a);
                part[0] += en.val * p.x; part[1] += en.val * p.y;
                if (VEC == 4) { const double2 q = lds128(a + plane); part[2] += en.val * q.x; part[3] += en.val * q.y; }
            }
            const double sg = c2s * s1;
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] += sg * part[v];
        }
        f(e, a0, early, ui, es, acc);
        e += estep;
        a0 += astep;
        i1 += NY;
        if (NORB >= 3) while (i1 >= D1) { i1 -= D1; i2++; }
    }
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
struct SlabMap {
    int n, ldc0;
    unsigned magic;                 // ceil(2^32 / ldc0): column / ldc0 by multiplication
    int col0[8], ldc[8];
    long long base[8];              // element offset of slab p
};
__device__ __forceinline__ int64_t slab_off(const SlabMap &M, int64_t row, int c)
{
    int p = (int)__umulhi((unsigned)c, M.magic);
    p = p < M.n - 1 ? p : M.n - 1;
    return M.base[p] + row * M.ldc[p] + (c - M.col0[p]);
}

// Persistent, double-buffered up pass for one up-block: y[rows][blk] += (diag + H_up) x.
// A tile = RP row pairs x the block, stored [rp][element][2] (the two rows of a pair interleaved).  Every CTA
// loads the star tables once, then walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the next tile streams in
// with cp.async (LDGSTS) while the current one is processed, so the HBM pipe stays busy during the gathers.
template <int NORB, bool SLAB>
__global__ void __launch_bounds__(kNT)
k_star_up(StarKParams P, SlabMap M, int64_t dim_dw, int64_t ld, int block_index, int RP, int accumulate, int nstage,
          const StarBlock *__restrict__ blocks,
          const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc, const double *__restrict__ hopv,
          const double *__restrict__ estar, const double *__restrict__ e_dw, const uint32_t *__restrict__ cfg_dw,
          const double *__restrict__ xtab, const double *__restrict__ x, double *__restrict__ y, int maxD)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StarBlock B = blocks[block_index];
    const int size = B.size, tid = threadIdx.x;
    const int tile_elems = RP * size;
    double *s_buf = reinterpret_cast<double *>(smem_raw);                      // [2 stages][RP][size][2]
    double *s_dg = s_buf + (size_t)2 * nstage * tile_elems;                    // [stages][RP][2 rows][8]
    const TabPtrs T = carve_tabs(reinterpret_cast<unsigned char *>(s_dg + (size_t)16 * nstage * RP), NORB, maxD, P.H);
    int D[3] = {1, 1, 1}, A0[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < NORB; a++) { D[a] = P.D[B.n[a]]; A0[a] = P.A0[B.n[a]]; }
    load_tabs<NORB>(P, B, D, hopd, hopc, hopv, estar, T, maxD, true);
    const uint32_t buf_addr = (uint32_t)__cvta_generic_to_shared(s_buf);
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
        for (int i = tid; i < 16 * RP; i += kNT) {
            // s_dg[stage][q][v][ui] = E_dw[row] + X[imp_dw(row)][ui] + (Ust-Jh) * C(nimp(ui), 2)
            const int q = i >> 4, v = (i >> 3) & 1, ui = i & 7;
            int64_t r = 2 * (t * RP + q) + v;
            r = r < last ? r : last;
            const int nimp = __popc(ui);
            s_dg[(size_t)stage * 16 * RP + i] = e_dw[r] + xtab[(cfg_dw[r] & impmask) * 32u + ui] + P.pair_e * (double)(nimp * (nimp - 1) / 2);
        }
    };

    // nstage == 2: one CTA per SM, the next tile streams in while the current one is processed.
    // nstage == 1: two CTAs per SM overlap each other's load and compute phases (large blocks: twice the warps).
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
        const double *dg = s_dg + (size_t)stage * 16 * RP;
        const int64_t row0 = 2 * t * RP;
        struct Own { double y0, y1; };
        if (RP == 1 && P.H <= 8) {
            // large block: one row pair per tile -> lean pass with constant-stride addressing
            const int64_t ra = row0;
            const bool oka1 = ra < dim_dw, okb1 = ra + 1 < dim_dw;
            double *pa = y + (oka1 ? ra : last) * ld + boff, *pb = y + (okb1 ? ra + 1 : last) * ld + boff;
            const double *dgp = dg;
            const int acc_y = accumulate;
            const int64_t rra = oka1 ? ra : last, rrb = okb1 ? ra + 1 : last;
            double *yy = y;
            lean_pass<NORB, 2, true, Own>(B, D, A0, T, maxD, P.H, buf_addr + (uint32_t)stage * stage_bytes, 0u,
                [=](int e) {
                    Own w;
                    if (SLAB) { w.y0 = acc_y ? yy[slab_off(M, rra, boff + e)] : 0.0; w.y1 = acc_y ? yy[slab_off(M, rrb, boff + e)] : 0.0; }
                    else { w.y0 = acc_y ? pa[e] : 0.0; w.y1 = acc_y ? pb[e] : 0.0; }
                    return w;
                },
                [=](int e, uint32_t a0, const Own &w, uint32_t ui, double es, double (&acc)[2]) {
                    const double2 p = lds128(a0);
                    const double o0 = w.y0 + acc[0] + (es + dgp[ui]) * p.x, o1 = w.y1 + acc[1] + (es + dgp[8 + ui]) * p.y;
                    if (SLAB) {
                        if (oka1) yy[slab_off(M, rra, boff + e)] = o0;
                        if (okb1) yy[slab_off(M, rrb, boff + e)] = o1;
                    } else {
                        if (oka1) pa[e] = o0;
                        if (okb1) pb[e] = o1;
                    }
                });
        } else {
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

// y[blk rows][strips] = H_dw x  for one down-block (runs FIRST; every row belongs to exactly one down-block, so
// this pass writes every element of y once).  A tile = SP strips of W columns x the block rows.
//   W = 4: two planes [strip][row][2] (columns 0-1 and 2-3 of each strip), 32-byte row segments (256-bit ld/st)
//   W = 2: one plane, 16-byte row segments -- for blocks whose 4-column tile exceeds shared memory (Norb=3, Nbath=5:
//          8000 rows x 32 B = 256 KB)
template <int NORB, int W>
__global__ void __launch_bounds__(kNT)
k_star_dw(StarKParams P, int64_t dim_up, int64_t ld, int block_index, int SP,
          const StarBlock *__restrict__ blocks,
          const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc, const double *__restrict__ hopv,
          const double *__restrict__ x, double *__restrict__ y, int maxD)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StarBlock B = blocks[block_index];
    const int R = B.size;
    const int tile_rows = SP * R;
    double *s_in = reinterpret_cast<double *>(smem_raw);                       // W/2 planes of [SP][R][2]
    const TabPtrs T = carve_tabs(reinterpret_cast<unsigned char *>(s_in + (size_t)W * tile_rows), NORB, maxD, P.H);
    const int tid = threadIdx.x;
    int D[3] = {1, 1, 1}, A0[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < NORB; a++) { D[a] = P.D[B.n[a]]; A0[a] = P.A0[B.n[a]]; }
    load_tabs<NORB>(P, B, D, hopd, hopc, hopv, nullptr, T, maxD, false);
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
    if (SP == 1 && P.H <= 8) {
        const int64_t left = dim_up - cbase;
        double *yc = ys + cbase;
        lean_pass<NORB, W, false, int>(B, D, A0, T, maxD, P.H, s_in_addr, plane,
            [=](int) { return 0; },
            [=](int e, uint32_t, int, uint32_t, double, double (&acc)[W]) { store(yc + (int64_t)e * ldv, left, acc); });
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
static void fill_kparams(const StarInfo &S, StarKParams &P)
{
    memset(&P, 0, sizeof(P));
    P.norb = S.norb; P.H = S.H; P.ncfg = S.ncfg; P.pair_e = S.pair_e;
    for (int m = 0; m < 16; m++) { P.D[m] = S.D[m]; P.A0[m] = S.A0[m]; P.coff[m] = S.coff[m]; }
}

static constexpr int kStageElems = 4900;     // elements (x 16 B) per pipeline stage of the up pass / per tile of the down pass

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
    const size_t tab = tabs_bytes(NORB, maxD, Dn.H);
    static size_t set_dw[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    const bool force_narrow = (ctx->par.reserved[0] & 1) != 0;                 // test hook: exercise the 2-column path
    for (size_t bi = 0; bi < Dn.blocks.size(); bi++) {                         // one launch per down-block
        const StarBlock &B = Dn.blocks[bi];
        const int W = (!force_narrow && sizeof(double) * (size_t)B.size * 4 + tab <= 227 * 1024) ? 4 : 2;
        const int64_t nstrips = (ncols + W - 1) / W;
        // strips per tile: up to kStageElems rows (x 32 B) of shared memory, but keep >= 4 CTAs per SM worth of tiles
        int64_t SP = std::max<int64_t>(1, kStageElems / B.size);
        SP = std::max<int64_t>(1, std::min<int64_t>(SP, nstrips / (4 * (int64_t)ctx->sm_count)));
        const size_t smem = sizeof(double) * (size_t)B.size * W * SP + tab;
        if (smem > 227 * 1024) return edgpu_fail(ctx, "star down pass: block of %d rows does not fit in shared memory", B.size);
        auto kern = (W == 4) ? k_star_dw<NORB, 4> : k_star_dw<NORB, 2>;
        size_t &set = set_dw[NORB][W == 4 ? 1 : 0];
        if (smem > set) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set = smem;
        }
        const unsigned nctas = (unsigned)((nstrips + SP - 1) / SP);
        kern<<<nctas, kNT, smem, ctx->stream>>>(PD, ncols, ld, (int)bi, (int)SP, Dn.d_blocks, Dn.d_hopd, Dn.d_hopc, Dn.d_hopv, x, y, maxD);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return 0;
}

// Up pass on a row range: x, y point at a [nrows][ld] tile holding ALL up-spin columns of down-spin rows
// [row0, row0+nrows) (the whole vector, or the row shard of one rank after the transpose).
template <int NORB>
static int launch_star_up(edgpu_sector *s, const double *x, double *y, int64_t row0, int64_t nrows, int64_t ld, int accumulate,
                          const SlabMap *slabs = nullptr)
{
    edgpu_ctx *ctx = s->ctx;
    const StarInfo &U = *s->up->star;
    StarKParams PU;
    fill_kparams(U, PU);
    int maxD = 0;
    for (int m = 0; m <= U.nbath + 1; m++) maxD = std::max(maxD, U.D[m]);
    if (maxD > kNT) return edgpu_fail(ctx, "star kernels: star dimension %d exceeds %d threads", maxD, kNT);
    const size_t tab = tabs_bytes(NORB, maxD, U.H);
    static size_t set_up[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    SlabMap M;
    memset(&M, 0, sizeof(M));
    if (slabs) M = *slabs; else { M.n = 1; M.ldc0 = (int)ld; M.magic = 0; M.ldc[0] = (int)ld; }
    const int64_t npairs = (nrows + 1) / 2;
    for (size_t bi = 0; bi < U.blocks.size(); bi++) {                          // persistent kernel, one launch per up-block
        const StarBlock &B = U.blocks[bi];
        int64_t RP = std::max<int64_t>(1, kStageElems / B.size);
        RP = std::max<int64_t>(1, std::min<int64_t>(RP, npairs / (2 * (int64_t)ctx->sm_count)));
        // single stage + 2 CTAs per SM when two tiles (and two table sets) fit; else double-buffered, 1 CTA per SM
        const size_t smem1 = sizeof(double) * ((size_t)2 * B.size * RP + (size_t)16 * RP) + tab;
        // double-buffered when two stages fit (nstage 1: single stage, e.g. 8000-configuration blocks of Norb=3, Nbath=5)
        const size_t smem2 = sizeof(double) * ((size_t)4 * B.size * RP + (size_t)32 * RP) + tab;
        const int nstage = (smem2 <= 227 * 1024 && !(ctx->par.reserved[0] & 2)) ? 2 : 1;
        const size_t smem = nstage == 1 ? smem1 : smem2;
        if (smem > 227 * 1024) return edgpu_fail(ctx, "star up pass: block of %d configurations does not fit in shared memory", B.size);
        auto kern = slabs ? k_star_up<NORB, true> : k_star_up<NORB, false>;
        size_t &set = set_up[NORB][slabs ? 1 : 0];
        if (smem > set) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set = smem;
        }
        const int64_t ntiles = (npairs + RP - 1) / RP;
        const int per_sm = nstage == 1 ? 2 : 1;
        const unsigned nctas = (unsigned)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * per_sm);
        kern<<<nctas, kNT, smem, ctx->stream>>>(PU, M, nrows, ld, (int)bi, (int)RP, accumulate, nstage, U.d_blocks, U.d_hopd, U.d_hopc,
                                                          U.d_hopv, U.d_estar, s->dw->ediag + row0, s->dw->cfg + row0, ctx->d_xtab, x, y, maxD);
        CUDA_TRY(ctx, cudaGetLastError());
    }
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

int hxv_star_launches(const edgpu_sector *s)
{
    return (int)s->up->star->blocks.size() + (int)s->dw->star->blocks.size();
}
