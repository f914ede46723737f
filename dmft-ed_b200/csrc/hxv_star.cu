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

struct StarBlock {             // one occupation tuple of one spin
    int off;                   // first internal index
    int size;                  // prod D[n_a]
    int n[EDGPU_MAXORB];       // star occupations
    int sgn_lower[EDGPU_MAXORB];   // (-1)^{sum_{a'<a} n_a'} as 0/1
    uint32_t magic0;               // ceil(2^32 / D0): tid / D0 == __umulhi(tid, magic0) for tid < 1024
    int ny;                        // kNT / D0
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
    int *d_upgroups = nullptr;                   // [ngroups][2] = (first block, count)
    int ngroups = 0, max_group_elems = 0;
    int max_block = 0;
    ~StarInfo()
    {
        cudaFree(d_blocks); cudaFree(d_hopj); cudaFree(d_hopd); cudaFree(d_hopc); cudaFree(d_hopv); cudaFree(d_estar); cudaFree(d_upgroups);
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
            B.ny = 768 / (int)d0;
        }
        blockoff[key] = cur;
        cur += size;
        S->blocks.push_back(B);
        S->max_block = std::max(S->max_block, size);
    }
    if (cur != (int)b->dim) return edgpu_fail(ctx, "star layout: internal size mismatch (%d vs %lld)", cur, (long long)b->dim);
    if ((int)S->blocks.size() > kMaxBlocks) return edgpu_fail(ctx, "star layout: too many blocks");
    // up-pass tile groups: consecutive blocks packed up to the largest block size
    std::vector<int> groups;
    {
        const int cap = std::max(S->max_block, 1);
        int first = 0, acc = 0;
        for (int i = 0; i < (int)S->blocks.size(); i++) {
            if (acc > 0 && acc + S->blocks[i].size > cap) { groups.push_back(first); groups.push_back(i - first); first = i; acc = 0; }
            acc += S->blocks[i].size;
        }
        groups.push_back(first); groups.push_back((int)S->blocks.size() - first);
        S->ngroups = (int)groups.size() / 2;
        S->max_group_elems = cap;
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
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_blocks, S->blocks.data(), sizeof(StarBlock) * S->blocks.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_hopj, hopj.data(), hopj.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_hopd, hopd.data(), sizeof(int16_t) * hopd.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_hopc, hopc.data(), hopc.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_hopv, hopv.data(), sizeof(double) * hopv.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_estar, estar.data(), sizeof(double) * estar.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(S->d_upgroups, groups.data(), sizeof(int) * groups.size(), cudaMemcpyHostToDevice, st));
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

static constexpr int kNT = 768;         // threads per CTA of the tiled kernels
static constexpr int kVec = 4;          // outputs per thread and outer step (rows in the up pass, columns in the down pass)

// Shared-memory star tables of one block (all NORB stars): per configuration i of star a
//   s_off[a][i][h]  byte offset of the h-th source inside a [element][2] double2 plane: (j - i) * stride_a * 16
//   s_val[a][i][h]  signed amplitude V_{a,k} * (-1)^{popc(bath bits of the star below k)}
//   s_cnt[a][i]     number of sources, s_e[a][i] star diagonal energy (up pass only)
struct TabPtrs {
    int *off;
    double *val, *e;
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
            T.off[a * maxD * H + t] = (int)hopd[(size_t)c0 * H + t] * stride * 16;
            T.val[a * maxD * H + t] = hopv[((size_t)a * P.ncfg + c0) * H + t];
        }
        for (int t = threadIdx.x; t < Da; t += kNT) {
            T.cnt[a * maxD + t] = hopc[c0 + t];
            if (with_e) T.e[a * maxD + t] = estar[(size_t)a * P.ncfg + c0 + t];
        }
        stride *= Da;
    }
}

__device__ __forceinline__ double2 lds128(uint32_t addr)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));     // not volatile: free to be scheduled early
    return v;
}

// acc[0..3] += sg * sum_h val[h] * tile[(e + delta_h)][0..3]   for one star; tile = two [e][2] planes
__device__ __forceinline__ void star_gather(double (&acc)[4], uint32_t a0, uint32_t plane, const int *off, const double *val,
                                            int cnt, double sg)
{
#pragma unroll 4
    for (int h = 0; h < cnt; h++) {
        const double v = sg * val[h];
        const uint32_t a = a0 + (uint32_t)off[h];
        const double2 p = lds128(a), q = lds128(a + plane);
        acc[0] += v * p.x; acc[1] += v * p.y; acc[2] += v * q.x; acc[3] += v * q.y;
    }
}

// One pass over a tile held in shared memory as two planes of [element][2] doubles (element = o*D0 + i0).
// Thread (ty, i0) walks o = ty, ty+NY, ...; F(o*D0+i0, ui, esum, acc) consumes the 4 gathered sums, where ui are
// the impurity bits of the element and esum the sum of the star energies (up pass).
template <int NORB, bool WITH_E, class PRE, class F>
__device__ __forceinline__ void tile_pass(const StarBlock &B, const int *D, const int *A0, const TabPtrs &T, int maxD, int H,
                                          uint32_t s_in_addr, uint32_t plane, PRE pre, F f)
{
    const int tid = threadIdx.x, D0 = D[0];
    const int ty = (D0 == 1) ? tid : (int)__umulhi((uint32_t)tid, B.magic0);   // tid / D0 by multiplication
    const int i0 = tid - ty * D0;
    const int NY = B.ny;                                 // kNT / D0
    if (ty >= NY) return;
    const int O = B.size / D0;
    const uint32_t imp0 = i0 >= A0[0] ? 1u : 0u;
    const double e0 = WITH_E ? T.e[i0] : 0.0;
    const int cnt0 = T.cnt[i0];
    const int *off0 = T.off + i0 * H;
    const double *val0 = T.val + i0 * H;
    const double c0s = (B.sgn_lower[0] & 1) ? -1.0 : 1.0;
    const double c1s = (NORB >= 2 && (B.sgn_lower[1] & 1)) ? -1.0 : 1.0;
    const double c2s = (NORB >= 3 && (B.sgn_lower[2] & 1)) ? -1.0 : 1.0;
    const double s0 = imp0 ? -1.0 : 1.0;
    int i1 = 0, i2 = 0;
    if (NORB == 2) i1 = ty;
    if (NORB >= 3) { i1 = ty; while (i1 >= D[1]) { i1 -= D[1]; i2++; } }
    for (int o = ty; o < O; o += NY) {
        const int e = o * D0 + i0;
        const uint32_t a0 = s_in_addr + (uint32_t)e * 16u;
        uint32_t ui = imp0;
        double es = e0, s1 = 1.0, s2 = 1.0;
        if (NORB >= 2) { const bool b = i1 >= A0[1]; ui |= b ? 2u : 0u; s1 = b ? -1.0 : 1.0; if (WITH_E) es += T.e[maxD + i1]; }
        if (NORB >= 3) { const bool b = i2 >= A0[2]; ui |= b ? 4u : 0u; s2 = b ? -1.0 : 1.0; if (WITH_E) es += T.e[2 * maxD + i2]; }
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const auto early = pre(e, a0);                       // loads issued before the gathers (latency overlap)
        // sign of star a: (-1)^{sum_{a'<a} n_a'} * prod_{a' != a} (-1)^{imp_a'}   (c/cdg rule, ED_SETUP.f90:1080-1106)
        star_gather(acc, a0, plane, off0, val0, cnt0, c0s * s1 * s2);
        if (NORB >= 2) star_gather(acc, a0, plane, T.off + (maxD + i1) * H, T.val + (maxD + i1) * H, T.cnt[maxD + i1], c1s * s0 * s2);
        if (NORB >= 3) star_gather(acc, a0, plane, T.off + (2 * maxD + i2) * H, T.val + (2 * maxD + i2) * H, T.cnt[2 * maxD + i2], c2s * s0 * s1);
        f(e, early, ui, es, acc);
        if (NORB == 2) i1 += NY;
        if (NORB >= 3) { i1 += NY; while (i1 >= D[1]) { i1 -= D[1]; i2++; } }
    }
}

__device__ __forceinline__ TabPtrs carve_tabs(unsigned char *base, int norb, int maxD, int H)
{
    TabPtrs T;
    T.val = reinterpret_cast<double *>(base);
    T.e = T.val + norb * maxD * H;
    T.off = reinterpret_cast<int *>(T.e + norb * maxD);
    T.cnt = reinterpret_cast<uint8_t *>(T.off + norb * maxD * H);
    return T;
}
static size_t tabs_bytes(int norb, int maxD, int H)
{
    return sizeof(double) * ((size_t)norb * maxD * H + (size_t)norb * maxD) + sizeof(int) * (size_t)norb * maxD * H + (size_t)norb * maxD + 64;
}

// y[rows][blk] += (E_up + E_dw[row] + X) x + H_up x   for 4 rows and one group of up-blocks (runs after the down pass).
template <int NORB>
__global__ void __launch_bounds__(kNT)
k_star_up(StarKParams P, int64_t dim_dw, int64_t ld,
          const StarBlock *__restrict__ blocks, const int *__restrict__ groups,
          const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc, const double *__restrict__ hopv,
          const double *__restrict__ estar, const double *__restrict__ e_dw, const uint32_t *__restrict__ cfg_dw,
          const double *__restrict__ xtab, const double *__restrict__ x, double *__restrict__ y, int maxD, int tile_elems)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_in = reinterpret_cast<double *>(smem_raw);                       // 2 planes of [tile_elems][2]
    double *s_dg = s_in + (size_t)4 * tile_elems;                              // [4 rows][8 imp patterns] row diagonal terms
    const TabPtrs T = carve_tabs(reinterpret_cast<unsigned char *>(s_dg + 32), NORB, maxD, P.H);
    const int tid = threadIdx.x;
    const int g = blockIdx.x;
    const int b0 = groups[2 * g], nb = groups[2 * g + 1];
    const int64_t r0 = (int64_t)blockIdx.y * kVec;
    const uint32_t impmask = (1u << NORB) - 1u;
    if (tid < 32) {
        // s_dg[v][ui] = E_dw[row v] + X[imp_dw(row v)][ui] + (Ust-Jh) * C(nimp(ui), 2)
        const int v = tid >> 3, ui = tid & 7;
        const int64_t r = (r0 + v < dim_dw) ? r0 + v : dim_dw - 1;
        const int nimp = __popc(ui);
        s_dg[tid] = e_dw[r] + xtab[(cfg_dw[r] & impmask) * 32u + ui] + P.pair_e * (double)(nimp * (nimp - 1) / 2);
    }
    const uint32_t s_in_addr = (uint32_t)__cvta_generic_to_shared(s_in);
    const uint32_t plane = (uint32_t)tile_elems * 16u;
    const double *xr[4];
    double *yr[4];
#pragma unroll
    for (int v = 0; v < 4; v++) {
        const int64_t r = (r0 + v < dim_dw) ? r0 + v : dim_dw - 1;     // tail rows duplicate the last row (never stored)
        xr[v] = x + r * ld;
        yr[v] = y + r * ld;
    }
    const int nvalid = (int)((dim_dw - r0) < 4 ? (dim_dw - r0) : 4);
    for (int bi = b0; bi < b0 + nb; bi++) {
        const StarBlock B = blocks[bi];
        int D[3] = {1, 1, 1}, A0[3] = {0, 0, 0};
#pragma unroll
        for (int a = 0; a < NORB; a++) { D[a] = P.D[B.n[a]]; A0[a] = P.A0[B.n[a]]; }
        __syncthreads();
        load_tabs<NORB>(P, B, D, hopd, hopc, hopv, estar, T, maxD, true);
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const double *src = xr[v] + B.off;
            double *dst = s_in + (size_t)(v >> 1) * 2 * tile_elems + (v & 1);
            for (int e = tid; e < B.size; e += kNT) dst[2 * e] = src[e];
        }
        __syncthreads();
        const int boff = B.off;
        struct Own { double2 p, q; double y0, y1, y2, y3; };
        tile_pass<NORB, true>(B, D, A0, T, maxD, P.H, s_in_addr, plane,
            [&](int e, uint32_t a0) {
                Own w; w.p = lds128(a0); w.q = lds128(a0 + plane);
                w.y0 = yr[0][boff + e]; w.y1 = yr[1][boff + e]; w.y2 = yr[2][boff + e]; w.y3 = yr[3][boff + e];   // H_dw x from the down pass
                return w;
            },
            [&](int e, const Own &w, uint32_t ui, double es, double (&acc)[4]) {
                const double2 p = w.p, q = w.q;
                const double o0 = w.y0 + acc[0] + (es + s_dg[ui]) * p.x, o1 = w.y1 + acc[1] + (es + s_dg[8 + ui]) * p.y;
                const double o2 = w.y2 + acc[2] + (es + s_dg[16 + ui]) * q.x, o3 = w.y3 + acc[3] + (es + s_dg[24 + ui]) * q.y;
                yr[0][boff + e] = o0;
                if (nvalid > 1) yr[1][boff + e] = o1;
                if (nvalid > 2) yr[2][boff + e] = o2;
                if (nvalid > 3) yr[3][boff + e] = o3;
            });
    }
}

// y[blk rows][c..c+4) = H_dw x  for one down-block (runs FIRST; every row belongs to exactly one down-block, so
// this pass writes every element of y once); each CTA walks `spc` strips of 4 columns.
template <int NORB>
__global__ void __launch_bounds__(kNT)
k_star_dw(StarKParams P, int64_t dim_up, int64_t ld, int block_index, int spc,
          const StarBlock *__restrict__ blocks,
          const int16_t *__restrict__ hopd, const uint8_t *__restrict__ hopc, const double *__restrict__ hopv,
          const double *__restrict__ x, double *__restrict__ y, int maxD)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StarBlock B = blocks[block_index];
    const int R = B.size;
    double *s_in = reinterpret_cast<double *>(smem_raw);                       // 2 planes of [R][2]
    const TabPtrs T = carve_tabs(reinterpret_cast<unsigned char *>(s_in + (size_t)4 * R), NORB, maxD, P.H);
    const int tid = threadIdx.x;
    int D[3] = {1, 1, 1}, A0[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < NORB; a++) { D[a] = P.D[B.n[a]]; A0[a] = P.A0[B.n[a]]; }
    load_tabs<NORB>(P, B, D, hopd, hopc, hopv, nullptr, T, maxD, false);
    const uint32_t s_in_addr = (uint32_t)__cvta_generic_to_shared(s_in);
    const uint32_t plane = (uint32_t)R * 16u;
    for (int sidx = 0; sidx < spc; sidx++) {
        const int64_t c0 = ((int64_t)blockIdx.x * spc + sidx) * 4;
        if (c0 >= dim_up) break;                                               // uniform across the CTA
        __syncthreads();
        // stage the strip: 4 contiguous doubles (32 bytes) per row; ld is a multiple of 4 so every row segment is
        // 32-byte aligned; pad columns beyond dim_up are zero in x and are never written in y
        const double *xs = x + (int64_t)B.off * ld + c0;
        for (int r = tid; r < R; r += kNT) {
            const double2 a = *reinterpret_cast<const double2 *>(xs + (int64_t)r * ld);
            const double2 b = *reinterpret_cast<const double2 *>(xs + (int64_t)r * ld + 2);
            *reinterpret_cast<double2 *>(s_in + (size_t)2 * r) = a;
            *reinterpret_cast<double2 *>(s_in + (size_t)2 * R + (size_t)2 * r) = b;
        }
        __syncthreads();
        double *ys = y + (int64_t)B.off * ld + c0;
        const int64_t left = dim_up - c0;
        tile_pass<NORB, false>(B, D, A0, T, maxD, P.H, s_in_addr, plane,
            [&](int, uint32_t) { return 0; },
            [&](int e, int, uint32_t, double, double (&acc)[4]) {
                double *yp = ys + (int64_t)e * ld;
                double2 a, b;
                a.x = acc[0]; a.y = acc[1]; b.x = acc[2]; b.y = acc[3];
                if (left >= 4) {
                    *reinterpret_cast<double2 *>(yp) = a;
                    *reinterpret_cast<double2 *>(yp + 2) = b;
                } else {                                                   // last strip: keep the pad columns at zero
                    yp[0] = a.x;
                    if (left > 1) yp[1] = a.y;
                    if (left > 2) yp[2] = b.x;
                }
            });
    }
}

// ------------------------------------------------------------------------------------------------------------
static void fill_kparams(const StarInfo &S, StarKParams &P)
{
    memset(&P, 0, sizeof(P));
    P.norb = S.norb; P.H = S.H; P.ncfg = S.ncfg; P.pair_e = S.pair_e;
    for (int m = 0; m < 16; m++) { P.D[m] = S.D[m]; P.A0[m] = S.A0[m]; P.coff[m] = S.coff[m]; }
}

template <int NORB>
static int launch_star(edgpu_sector *s, const double *x, double *y)
{
    edgpu_ctx *ctx = s->ctx;
    const StarInfo &U = *s->up->star, &Dn = *s->dw->star;
    StarKParams PU, PD;
    fill_kparams(U, PU);
    fill_kparams(Dn, PD);
    int maxD = 0;
    for (int m = 0; m <= U.nbath + 1; m++) maxD = std::max(maxD, U.D[m]);
    if (maxD > kNT) return edgpu_fail(ctx, "star kernels: star dimension %d exceeds %d threads", maxD, kNT);
    const size_t tab = tabs_bytes(NORB, maxD, U.H);
    // ---- down pass: one launch per down-block
    {
        const int64_t nstrips = (s->dim_up + 3) / 4;
        for (size_t bi = 0; bi < Dn.blocks.size(); bi++) {
            const StarBlock &B = Dn.blocks[bi];
            const size_t smem = sizeof(double) * (size_t)B.size * 4 + tab;
            if (smem > 227 * 1024) return edgpu_fail(ctx, "star down pass: block of %d rows does not fit in shared memory", B.size);
            static size_t set_dw[4] = {0, 0, 0, 0};
            if (smem > set_dw[NORB]) {
                CUDA_TRY(ctx, cudaFuncSetAttribute(k_star_dw<NORB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                set_dw[NORB] = smem;
            }
            // strips per CTA: ~16K tile elements of work per CTA, but keep at least 4 CTAs per SM in flight
            int64_t spc = std::max<int64_t>(1, 16384 / ((int64_t)B.size * 4));
            spc = std::max<int64_t>(1, std::min<int64_t>(spc, nstrips / (4 * (int64_t)ctx->sm_count)));
            const unsigned nctas = (unsigned)((nstrips + spc - 1) / spc);
            k_star_dw<NORB><<<nctas, kNT, smem, ctx->stream>>>(PD, s->dim_up, s->ld, (int)bi, (int)spc, Dn.d_blocks, Dn.d_hopd,
                                                              Dn.d_hopc, Dn.d_hopv, x, y, maxD);
            CUDA_TRY(ctx, cudaGetLastError());
        }
    }
    // ---- up pass
    {
        const int tile_elems = U.max_block;
        const size_t smem = sizeof(double) * ((size_t)4 * tile_elems + 32) + tab;
        if (smem > 227 * 1024) return edgpu_fail(ctx, "star up pass: block of %d configurations does not fit in shared memory", U.max_block);
        static size_t set_up[4] = {0, 0, 0, 0};
        if (smem > set_up[NORB]) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(k_star_up<NORB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_up[NORB] = smem;
        }
        dim3 grid((unsigned)U.ngroups, (unsigned)((s->dim_dw + kVec - 1) / kVec));
        if (grid.y > 65535) return edgpu_fail(ctx, "star up pass: too many row groups");
        k_star_up<NORB><<<grid, kNT, smem, ctx->stream>>>(PU, s->dim_dw, s->ld, U.d_blocks, U.d_upgroups,
                                                         U.d_hopd, U.d_hopc, U.d_hopv, U.d_estar, s->dw->ediag, s->dw->cfg,
                                                         ctx->d_xtab, x, y, maxD, tile_elems);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return 0;
}

int hxv_star(edgpu_sector *s, const double *x, double *y)
{
    if (!s->up->star || !s->dw->star) return edgpu_fail(s->ctx, "hxv_star: sector is not in the star-product layout");
    switch (s->ctx->ham.norb) {
    case 1: return launch_star<1>(s, x, y);
    case 2: return launch_star<2>(s, x, y);
    case 3: return launch_star<3>(s, x, y);
    default: return edgpu_fail(s->ctx, "hxv_star: Norb=%d unsupported", s->ctx->ham.norb);
    }
}

int hxv_star_launches(const edgpu_sector *s)
{
    return 1 + (int)s->dw->star->blocks.size();
}
