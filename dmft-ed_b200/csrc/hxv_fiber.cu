// hxv_fiber.cu -- "fiber" H*v engine on the pair-tile layout (round 2; the fast path of the Ns=14/16/18 configs).
//
// Replaces directMatVec_cc (ED_HAMILTONIAN_DIRECT_HxV.f90:21-92 + ED_HAMILTONIAN/direct/*.f90) for bath_type=normal
// with diagonal impHloc and Norb = 2 or 3, like hxv_star.cu, with two changes of design:
//
// 1. LAYOUT ("layout 3", pair tiles).  Star occupations are conserved per spin (hxv_star.cu:6-10), so H is block
//    diagonal over (down-block, up-block) PAIRS.  A sector vector is stored pair by pair; pair (bi,bj) is an R x C matrix
//    (rows = down configurations of block bi, columns = up configurations of block bj) in 4x4 micro-tiles of 128 bytes:
//        element (rp, cp) at base + ((rp/4)*C4 + cp/4)*16 + (rp%4)*4 + cp%4.
//    A band of 4 rows is contiguous (one bulk copy); a strip of 4 columns is R/4 runs of 128 bytes (3-D tensor map, box rows
//    of 128 B instead of the 32 B of layout 2, which limited the copy engine).  Positions are padded per star-0 index:
//    rp = o*d0r + k with d0r odd, cp = o*d0p + k with d0p == 2 (mod 4) -- these make the shared-memory accesses below
//    bank-conflict free.  Pads hold zeros and are never written with anything else.  Because pairs are independent, a
//    multi-GPU run deals whole pairs to ranks: H*v then needs NO exchange (only the Lanczos scalars are all-reduced).
//
// 2. KERNELS ("fibers").  Inside a pair the operator is a Kronecker sum of small star matrices.  A thread keeps one
//    FIBER (all <= 70 configurations of star 0 for fixed other indices) in REGISTERS and applies the star-0 hops there:
//    the hop structure (which configuration connects to which, with which sign) depends only on (Nbath, occupation) and is
//    generated at COMPILE TIME (fib::out below), amplitudes come from the constant bank.  The other star(s) of the same spin
//    are gathered from the shared-memory image as whole neighbour fibers (<= 7 per fiber, table driven).  Per output this is
//    ~5 shared-memory reads instead of the ~13 of the table kernels of round 1 (which were shared-memory-pipe bound).
//      pass 1  k_fib_up : tile = G bands of 4 rows x the up-block; y  = (diag + H_up) x          R(x) W(y)
//      pass 2  k_fib_dw : tile = G strips of 4 columns x the down-block; y += H_dw x (+ <x,y>)   R(x) R(y) W(y)
//    Blocks the fiber kernels cannot take (star-0 dimension 1, images larger than shared memory, tiny blocks) go through
//    the thread-per-element pair kernels at the end of this file.
#include "star_info.h"
#include <cuda.h>
#include <algorithm>
#include <cstring>
#include <map>
#include <utility>

uint64_t edgpu_binom(int n, int k);

// ------------------------------------------------------------------------------------------------------------
// compile-time combinatorics of one star (NB bath levels, M particles): configuration index <-> (imp, bath set)
//   index < A0 = C(NB,M): imp = 0, bath set = colex_unrank(index, M) ; else imp = 1, bath set = colex_unrank(index-A0, M-1)
// (the order of build_star_layout, hxv_star.cu)
// ------------------------------------------------------------------------------------------------------------
namespace fib {
__host__ __device__ constexpr int cbinom(int n, int k)
{
    if (k < 0 || k > n) return 0;
    long long r = 1;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return (int)r;
}
__host__ __device__ constexpr int cpopc(unsigned w) { int c = 0; while (w) { c += (int)(w & 1u); w >>= 1; } return c; }
__host__ __device__ constexpr int crank(unsigned w)
{
    int r = 0, i = 1;
    for (int p = 0; p < 16; p++)
        if ((w >> p) & 1u) { r += cbinom(p, i); i++; }
    return r;
}
__host__ __device__ constexpr unsigned cunrank(int r, int m)
{
    unsigned w = 0;
    for (int k = m; k >= 1; k--) {
        int p = k - 1;
        while (cbinom(p + 1, k) <= r) p++;
        w |= 1u << p;
        r -= cbinom(p, k);
    }
    return w;
}
__host__ __device__ constexpr int ccoff(int nl, int m) { int s = 0; for (int i = 0; i < m; i++) s += cbinom(nl, i); return s; }

template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F &&f) { (f(std::integral_constant<int, Is>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) { static_for_impl(std::make_integer_sequence<int, N>{}, f); }

// One hop term of output configuration K: bath level KAPPA.  Source index and sign are compile-time constants.
//   K <  A0 (imp=0, bath S): terms kappa in S,     source = A0 + rank(S \ kappa)
//   K >= A0 (imp=1, bath T): terms kappa not in T, source = rank(T + kappa)
// sign = (-1)^{# bath bits of the star below kappa}  (c/cdg rule of ED_SETUP.f90:1080-1106 restricted to the star; the
// factors from the other stars are applied by the caller).  BASE: index of x[0] of the array passed (0: whole fiber,
// A0: only the imp=1 half is held, i.e. PART A of the down pass).
template <int NB, int M, int K, int BASE, int N>
__device__ __forceinline__ double out(const double (&x)[N], const double *__restrict__ v)
{
    constexpr int A0 = cbinom(NB, M);
    constexpr bool isA = K < A0;
    constexpr unsigned S = isA ? cunrank(K, M) : cunrank(K - A0, M - 1);
    double acc = 0.0;
    static_for<NB>([&](auto kk) {
        constexpr int kap = decltype(kk)::value;
        constexpr bool has = (S >> kap) & 1u;
        if constexpr (isA ? has : !has) {
            constexpr unsigned S2 = S ^ (1u << kap);
            constexpr int src = isA ? A0 + crank(S2) : crank(S2);
            constexpr bool neg = cpopc(S & ((1u << kap) - 1u)) & 1;
            static_assert(src - BASE >= 0 && src - BASE < N, "fiber source outside the register array");
            acc = fma(neg ? -v[kap] : v[kap], x[src - BASE], acc);
        }
    });
    return acc;
}
}   // namespace fib

// ------------------------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------------------------
static constexpr int kHS = 12;                      // gather slots per fiber (stars 1.. of the same spin)
static constexpr int kSlot = 110592;                // bytes per pipeline slot (2 slots = 216 KB of dynamic shared memory)
static constexpr int kStab = 10240;                 // shared-memory copy of the outer table of the current block (SOuter entries)
static constexpr int kStabHS = 7;                   // slots per entry of that copy (blocks with more slots read the global table)
static constexpr int kFibMinBlock = 256;            // blocks smaller than this use the thread-per-element pair kernels

struct FibBlockDev {
    int off, size;              // internal index range in the spin basis
    int m0, D0, A0;             // star-0 occupation, dimension, number of imp=0 configurations
    int nouter;                 // size / D0: combined index of stars 1..
    int d0r, d0p;               // padded star-0 extent as ROW index (odd) / COLUMN index (== 2 mod 4)
    int R, R4, C, C4;           // padded extents: R = nouter*d0r rows, C = nouter*d0p columns; micro-tile counts
    int tab;                    // first entry of the block in the outer table
    int fiber;                  // fiber kernels apply (else the generic pair kernels)
    int BR, nbox;               // down pass: bands per tensor box, boxes per strip
    int hsmax;                  // largest slot count of a fiber of the block
};

struct __align__(16) OuterEnt {  // one value of the outer index o (stars 1..) of a block; 128 bytes
    double eo;                  // sum of the star energies of stars 1..
    int impbits;                // impurity bits of stars 1.. (bit a), star 0 bit clear
    int nslot;
    int neg;                    // 1: (-1)^{sum of the impurity bits of stars 1..} = -1 (sign of the star-0 hops)
    int pad0;
    int delta[kHS];             // neighbour fiber: o' - o
    int code[kHS];              // signed amplitude index (FibSpin::d_amps); sign holds everything except (-1)^{imp_0}
    int pad1[2];
};
static_assert(sizeof(OuterEnt) == 128, "OuterEnt must be 128 bytes");

struct __align__(16) SOuter {     // shared-memory form of OuterEnt with the amplitudes resolved; 112 bytes
    double eo;
    int impbits, nslot, neg, pad;
    int delta[kStabHS];
    int pad2;
    double amp[kStabHS];
};
static_assert(sizeof(SOuter) == 112, "SOuter must be 112 bytes");

struct FibConst {               // by-value kernel argument: compile-time indexed => constant-bank operands
    double e0[256];             // star-0 energies, [ccoff(NL, m) + i]
    double v0[8];               // star-0 hybridisations V_{0,kappa}
    double pair_e;              // (Ust - Jh): same-spin inter-orbital term
};

struct FibSpin {
    int nl = 0, norb = 0;
    std::vector<FibBlockDev> blocks;
    FibBlockDev *d_blocks = nullptr;
    OuterEnt *d_outer = nullptr;
    double *d_amps = nullptr;   // [2 * norb * nbath] signed amplitudes
    FibConst cst;
    ~FibSpin() { cudaFree(d_blocks); cudaFree(d_outer); cudaFree(d_amps); }
};

struct PairDev { int bi, bj; int64_t base; };
struct FibTile {                // pass 1: bands [a, a+b) of the pair ; pass 2: strips [a, a+b)
    int64_t off;                // pass 1: first element of band a in the vector ; pass 2: pair base
    int pair, blk, a, b;
    int bytes, pad;
};

struct PairLayout {
    int nbd = 0, nbu = 0;
    std::vector<int64_t> pbase;
    std::vector<PairDev> pairs;
    int2 *d_rowinfo = nullptr, *d_colinfo = nullptr;
    int64_t *d_pbase = nullptr;
    int *d_c4 = nullptr;
    PairDev *d_pairs = nullptr;
    FibTile *d_t1 = nullptr, *d_t2 = nullptr;
    int n1 = 0, n2 = 0;
    int *d_g1 = nullptr, *d_g2 = nullptr;      // pair ids for the generic up / down kernels
    int ng1 = 0, ng2 = 0;
    int64_t g1_elems = 0, g2_elems = 0;
    int nl = 0;
    int slot = kSlot;                                      // pipeline slot size the tile schedules were built for
    int skip1 = 0, skip2 = 0;                              // test hooks: leave a pass to the thread-per-element kernels
    std::map<const double *, CUtensorMap *> tmaps;         // per source pointer: one 3-D map per pair (device array)
    ~PairLayout()
    {
        cudaFree(d_rowinfo); cudaFree(d_colinfo); cudaFree(d_pbase); cudaFree(d_c4); cudaFree(d_pairs); cudaFree(d_t1); cudaFree(d_t2);
        cudaFree(d_g1); cudaFree(d_g2);
        for (auto &kv : tmaps) cudaFree(kv.second);
    }
};

static int colex_rank_host(uint32_t w)
{
    int r = 0, i = 1;
    for (int p = 0; p < 32; p++)
        if ((w >> p) & 1u) { r += (int)edgpu_binom(p, i); i++; }
    return r;
}

// Per-spin fiber tables from the star description (same conventions as build_star_layout, hxv_star.cu).
static int build_fib_spin(edgpu_ctx *ctx, SpinBasis *b)
{
    if (b->fib) return 0;
    if (!b->star) return edgpu_fail(ctx, "fiber kernels need the star-product order of the spin basis");
    const StarInfo &S = *b->star;
    const HamParams &h = ctx->ham;
    const int norb = S.norb, nbath = S.nbath, nl = nbath + 1, ps = b->pspin;
    auto F = std::make_shared<FibSpin>();
    F->nl = nl; F->norb = norb;
    // star configuration lists per occupation
    std::vector<std::vector<uint32_t>> cfg_of(nl + 1);
    for (int m = 0; m <= nl; m++) cfg_of[m].resize(S.D[m]);
    for (uint32_t sub = 0; sub < (1u << nl); sub++) {
        const int imp = sub & 1u, m = __builtin_popcount(sub);
        const int idx = imp ? S.A0[m] + colex_rank_host(sub >> 1) : colex_rank_host(sub >> 1);
        cfg_of[m][idx] = sub;
    }
    auto srank = [&](uint32_t sub) {
        const int imp = sub & 1u, m = __builtin_popcount(sub);
        return imp ? S.A0[m] + colex_rank_host(sub >> 1) : colex_rank_host(sub >> 1);
    };
    auto estar = [&](int a, uint32_t sub) {
        double cimp = h.H(ps, a, a) - h.xmu;
        if (h.hfmode) {
            cimp -= 0.5 * h.uloc[a];
            if (norb > 1) cimp -= (norb - 1) * (0.5 * h.ust + 0.5 * (h.ust - h.jh));
        }
        double e = (sub & 1u) ? cimp : 0.0;
        for (int k = 0; k < nbath; k++)
            if ((sub >> (k + 1)) & 1u) e += h.E(ps, a, k);
        return e;
    };
    memset(&F->cst, 0, sizeof(F->cst));
    if (nl <= 8) {
        int off = 0;
        for (int m = 0; m <= nl; m++) {
            for (int i = 0; i < S.D[m]; i++) F->cst.e0[off + i] = estar(0, cfg_of[m][i]);
            off += S.D[m];
        }
        for (int k = 0; k < nbath && k < 8; k++) F->cst.v0[k] = h.V(ps, 0, k);
    }
    F->cst.pair_e = S.pair_e;
    std::vector<double> amps((size_t)2 * norb * nbath);
    for (int a = 0; a < norb; a++)
        for (int k = 0; k < nbath; k++) {
            amps[(size_t)2 * (a * nbath + k)] = h.V(ps, a, k);
            amps[(size_t)2 * (a * nbath + k) + 1] = -h.V(ps, a, k);
        }
    std::vector<OuterEnt> outer;
    for (const StarBlock &B : S.blocks) {
        FibBlockDev fb;
        memset(&fb, 0, sizeof(fb));
        fb.off = B.off; fb.size = B.size; fb.m0 = B.n[0];
        fb.D0 = S.D[B.n[0]]; fb.A0 = S.A0[B.n[0]];
        fb.nouter = B.size / fb.D0;
        fb.d0r = fb.D0 | 1;
        fb.d0p = fb.D0 + ((6 - (fb.D0 & 3)) & 3);                 // smallest value >= D0 that is 2 (mod 4)
        if ((fb.d0p & 3) != 2 || fb.d0p < fb.D0) return edgpu_fail(ctx, "fiber layout: bad column padding");
        fb.R = fb.nouter * fb.d0r; fb.R4 = (fb.R + 3) / 4;
        fb.C = fb.nouter * fb.d0p; fb.C4 = (fb.C + 3) / 4;
        fb.nbox = (fb.R4 + 255) / 256;
        fb.BR = (fb.R4 + fb.nbox - 1) / fb.nbox;
        fb.tab = (int)outer.size();
        int hsmax = 0;
        // outer index o: mixed radix over stars 1.., star 1 fastest
        for (int o = 0; o < fb.nouter; o++) {
            OuterEnt e;
            memset(&e, 0, sizeof(e));
            int idx[EDGPU_MAXORB] = {0}, rem = o;
            uint32_t sub[EDGPU_MAXORB] = {0};
            int nimp = 0;
            for (int a = 1; a < norb; a++) {
                const int Da = S.D[B.n[a]];
                idx[a] = rem % Da; rem /= Da;
                sub[a] = cfg_of[B.n[a]][idx[a]];
                e.eo += estar(a, sub[a]);
                if (sub[a] & 1u) { e.impbits |= 1 << a; nimp++; }
            }
            e.neg = nimp & 1;
            int stride = 1, ns = 0;
            for (int a = 1; a < norb; a++) {
                const int Da = S.D[B.n[a]];
                for (int k = 0; k < nbath; k++) {
                    const uint32_t bi = sub[a] & 1u, bk = (sub[a] >> (k + 1)) & 1u;
                    if (!(bi ^ bk)) continue;
                    if (h.V(ps, a, k) == 0.0) continue;                     // the reference skips exactly-zero amplitudes
                    const uint32_t sub2 = sub[a] ^ 1u ^ (1u << (k + 1));
                    int neg = __builtin_popcount((sub[a] >> 1) & ((1u << k) - 1u)) & 1;      // star-local parity
                    neg ^= B.sgn_lower[a] & 1;                                              // (-1)^{sum_{a'<a} n_a'}
                    neg ^= (nimp - (int)(sub[a] & 1u)) & 1;                                 // impurity bits of the OTHER stars >= 1
                    if (ns >= kHS) return edgpu_fail(ctx, "fiber tables: more than %d gather slots", kHS);
                    e.delta[ns] = (srank(sub2) - idx[a]) * stride;
                    e.code[ns] = 2 * (a * nbath + k) + neg;
                    ns++;
                }
                stride *= Da;
            }
            e.nslot = ns;
            hsmax = std::max(hsmax, ns);
            outer.push_back(e);
        }
        // fiber kernels: star 0 must have hops (1 <= m0 <= nl-1), the images must fit, the block must be worth a tile
        const int64_t band_bytes = (int64_t)fb.C4 * 128, strip_bytes = (int64_t)fb.nbox * fb.BR * 128;
        fb.fiber = (nl >= 3 && nl <= 8 && norb >= 2 && fb.m0 >= 1 && fb.m0 <= nl - 1 && band_bytes <= 2 * kSlot && strip_bytes <= 2 * kSlot &&
                    hsmax <= (nl >= 7 ? 7 : kHS)) ? 1 : 0;
        fb.hsmax = hsmax;
        F->blocks.push_back(fb);
    }
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMalloc(&F->d_blocks, sizeof(FibBlockDev) * F->blocks.size()));
    CUDA_TRY(ctx, cudaMalloc(&F->d_outer, sizeof(OuterEnt) * std::max<size_t>(1, outer.size())));
    CUDA_TRY(ctx, cudaMalloc(&F->d_amps, sizeof(double) * amps.size()));
    CUDA_TRY(ctx, cudaMemcpyAsync(F->d_blocks, F->blocks.data(), sizeof(FibBlockDev) * F->blocks.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(F->d_outer, outer.data(), sizeof(OuterEnt) * outer.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(F->d_amps, amps.data(), sizeof(double) * amps.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    b->fib = F;
    return 0;
}

bool pair_layout_supported(const edgpu_sector *s)
{
    const edgpu_ctx *ctx = s->ctx;
    const HamParams &h = ctx->ham;
    return s->up->layout == 2 && s->dw->layout == 2 && s->up->star && s->dw->star && !h.jhflag && h.norb >= 2 && h.norb <= 3 &&
           h.nbath + 1 >= 3 && h.nbath + 1 <= 8;
}

VAddr sector_vaddr(const edgpu_sector *s)
{
    VAddr a;
    if (!s->pl) { a.mode = 0; a.ld = s->ld; return a; }
    a.mode = 3; a.nbu = s->pl->nbu; a.ld = 0;
    a.rowinfo = s->pl->d_rowinfo; a.colinfo = s->pl->d_colinfo; a.pbase = s->pl->d_pbase; a.c4 = s->pl->d_c4;
    return a;
}

// Pair-tile layout of a sector: pair bases, position tables, tile schedules.  rank/nranks: the pairs are dealt to
// `nranks` processes by longest-processing-time-first over their element counts; this process keeps those of `rank`.
int pair_layout_build(edgpu_sector *s, int rank, int nranks)
{
    edgpu_ctx *ctx = s->ctx;
    if (!pair_layout_supported(s)) return edgpu_fail(ctx, "pair-tile layout: unsupported shape (needs bath_type=normal star order, Norb 2-3, Nbath 2-7, no Jx/Jp)");
    if (nranks < 1 || rank < 0 || rank >= nranks) return edgpu_fail(ctx, "pair-tile layout: bad rank %d of %d", rank, nranks);
    if (int rc = build_fib_spin(ctx, s->up.get())) return rc;
    if (int rc = build_fib_spin(ctx, s->dw.get())) return rc;
    const FibSpin &FU = *s->up->fib, &FD = *s->dw->fib;
    auto P = std::make_shared<PairLayout>();
    P->nbd = (int)FD.blocks.size(); P->nbu = (int)FU.blocks.size(); P->nl = FU.nl;
    const int nbd = P->nbd, nbu = P->nbu;
    const bool force_fiber = (ctx->par.reserved[0] & 16) != 0, force_generic = (ctx->par.reserved[0] & 4) != 0;
    const bool gen1 = force_generic || (ctx->par.reserved[0] & 1024) != 0, gen2 = force_generic || (ctx->par.reserved[0] & 2048) != 0;
    if (ctx->par.reserved[0] & 4096) {
        // test hook: shrink the pipeline slot so that the largest image of this sector needs BOTH slots (the two-slot
        // path that only the 4900-configuration blocks of Ns=16 take in production)
        int64_t big = 0;
        for (const FibBlockDev &B : FU.blocks) if (B.fiber) big = std::max<int64_t>(big, (int64_t)B.C4 * 128);
        for (const FibBlockDev &B : FD.blocks) if (B.fiber) big = std::max<int64_t>(big, (int64_t)B.nbox * B.BR * 128);
        if (big > 256) P->slot = (int)std::min<int64_t>(kSlot, ((big / 2 + 127) / 128) * 128);
    }
    const int slot = P->slot;
    // ownership: LPT over pair sizes (deterministic: ties by pair index)
    std::vector<int64_t> psize((size_t)nbd * nbu);
    std::vector<int> order((size_t)nbd * nbu), owner((size_t)nbd * nbu, 0);
    for (int i = 0; i < nbd; i++)
        for (int j = 0; j < nbu; j++) { psize[(size_t)i * nbu + j] = (int64_t)FD.blocks[i].R4 * FU.blocks[j].C4 * 16; order[(size_t)i * nbu + j] = i * nbu + j; }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return psize[a] > psize[b]; });
    {
        std::vector<int64_t> load(nranks, 0);
        for (int p : order) {
            int best = 0;
            for (int r = 1; r < nranks; r++) if (load[r] < load[best]) best = r;
            owner[p] = best; load[best] += psize[p];
        }
    }
    P->pbase.assign((size_t)nbd * nbu, -1);
    int64_t cur = 0;
    for (int i = 0; i < nbd; i++)
        for (int j = 0; j < nbu; j++) {
            const size_t p = (size_t)i * nbu + j;
            if (owner[p] != rank) continue;
            P->pbase[p] = cur;
            PairDev pd; pd.bi = i; pd.bj = j; pd.base = cur;
            P->pairs.push_back(pd);
            cur += psize[p];
        }
    s->nalloc = std::max<int64_t>(cur, 16);
    s->ld = 0;
    s->shard_rank = rank; s->shard_nranks = nranks;
    // position tables
    std::vector<int2> rowinfo((size_t)s->dim_dw), colinfo((size_t)s->dim_up);
    for (int i = 0; i < nbd; i++) {
        const FibBlockDev &B = FD.blocks[i];
        for (int e = 0; e < B.size; e++) rowinfo[(size_t)B.off + e] = make_int2(i, (e / B.D0) * B.d0r + e % B.D0);
    }
    std::vector<int> c4(nbu);
    for (int j = 0; j < nbu; j++) {
        const FibBlockDev &B = FU.blocks[j];
        c4[j] = B.C4;
        for (int e = 0; e < B.size; e++) colinfo[(size_t)B.off + e] = make_int2(j, (e / B.D0) * B.d0p + e % B.D0);
    }
    // tile schedules
    std::vector<FibTile> t1, t2;
    std::vector<int> g1, g2;
    for (int p = 0; p < (int)P->pairs.size(); p++) {
        const PairDev &pd = P->pairs[p];
        const FibBlockDev &BD = FD.blocks[pd.bi], &BU = FU.blocks[pd.bj];
        const bool fu = !gen1 && BU.fiber && (force_fiber || BU.size >= kFibMinBlock);
        const bool fd = !gen2 && BD.fiber && (force_fiber || BD.size >= kFibMinBlock);
        if (fu) {
            const int64_t bb = (int64_t)BU.C4 * 128;
            int G = (int)std::max<int64_t>(1, std::min<int64_t>(8, slot / bb));
            for (int a = 0; a < BD.R4; a += G) {
                FibTile t; memset(&t, 0, sizeof(t));
                t.pair = p; t.blk = pd.bj; t.a = a; t.b = std::min(G, BD.R4 - a);
                t.off = pd.base + (int64_t)a * BU.C4 * 16;
                t.bytes = (int)(bb * t.b);
                t1.push_back(t);
            }
        } else { g1.push_back(p); P->g1_elems += (int64_t)BD.size * BU.size; }
        if (fd) {
            const int64_t sb = (int64_t)BD.nbox * BD.BR * 128;
            int G = (int)std::max<int64_t>(1, std::min<int64_t>(8, slot / sb));
            for (int a = 0; a < BU.C4; a += G) {
                FibTile t; memset(&t, 0, sizeof(t));
                t.pair = p; t.blk = pd.bi; t.a = a; t.b = std::min(G, BU.C4 - a);
                t.off = pd.base;
                t.bytes = (int)(sb * t.b);
                t2.push_back(t);
            }
        } else { g2.push_back(p); P->g2_elems += (int64_t)BD.size * BU.size; }
    }
    // big tiles first (a CTA's sequence is then descending: two-slot tiles precede one-slot tiles, see the pipeline)
    // (equal sizes: by block, so that a CTA meets long runs of the same block)
    auto bysize = [](const FibTile &a, const FibTile &b) { return a.bytes != b.bytes ? a.bytes > b.bytes : a.blk < b.blk; };
    std::stable_sort(t1.begin(), t1.end(), bysize);
    std::stable_sort(t2.begin(), t2.end(), bysize);
    P->n1 = (int)t1.size(); P->n2 = (int)t2.size(); P->ng1 = (int)g1.size(); P->ng2 = (int)g2.size();
    cudaStream_t st = ctx->stream;
    auto up = [&](auto **dptr, const auto &vec) -> cudaError_t {
        using T = typename std::remove_reference<decltype(vec)>::type::value_type;
        cudaError_t e = cudaMalloc(dptr, sizeof(T) * std::max<size_t>(1, vec.size()));
        if (e != cudaSuccess) return e;
        if (vec.empty()) return cudaSuccess;
        return cudaMemcpyAsync(*dptr, vec.data(), sizeof(T) * vec.size(), cudaMemcpyHostToDevice, st);
    };
    CUDA_TRY(ctx, up(&P->d_rowinfo, rowinfo));
    CUDA_TRY(ctx, up(&P->d_colinfo, colinfo));
    CUDA_TRY(ctx, up(&P->d_pbase, P->pbase));
    CUDA_TRY(ctx, up(&P->d_c4, c4));
    CUDA_TRY(ctx, up(&P->d_pairs, P->pairs));
    CUDA_TRY(ctx, up(&P->d_t1, t1));
    CUDA_TRY(ctx, up(&P->d_t2, t2));
    CUDA_TRY(ctx, up(&P->d_g1, g1));
    CUDA_TRY(ctx, up(&P->d_g2, g2));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    s->pl = P;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fmbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void fmbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void fmbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void fmbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "FLAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra FDONE;\n"
        "bra FLAB_WAIT;\n"
        "FDONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// one lane polls, the warp follows (32 spinning lanes per warp only burn issue slots of the warps that still compute)
__device__ __forceinline__ void fmbar_wait_warp(uint32_t bar, uint32_t parity)
{
    if ((threadIdx.x & 31) == 0) fmbar_wait(bar, parity);
    __syncwarp();
}
// the producer thread shares a scheduler with consumer warps: poll with a pause instead of burning its issue slots
__device__ __forceinline__ void fmbar_wait_backoff(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (true) {
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void fbulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ftma_prefetch_3d(const CUtensorMap *tm, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void fbulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ftma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ double2 flds128(uint32_t addr)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ int flds32(uint32_t addr)
{
    int v;
    asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double flds64(uint32_t addr)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void fstg128(double *p, double a, double b)
{
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ double fldg64(const double *p)
{
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

struct FibArgs {
    FibConst cst;                       // of the fiber spin (up in pass 1, down in pass 2)
    const FibBlockDev *blk_f;           // blocks of the fiber spin
    const FibBlockDev *blk_o;           // blocks of the other spin
    const OuterEnt *outer;
    const double *amps;
    const PairDev *pairs;
    const FibTile *tiles;
    int ntiles;
    uint32_t impmask;
    int slot;                           // bytes per pipeline slot (kSlot; smaller in the two-slot test mode)
    int dbg;                            // measurement hooks: bit 0 = consumers skip the fibers (load pipeline only)
    const double *x;
    double *y;
    const double *e_dw;                 // pass 1: per-row diagonal energy and configuration word of the down spin
    const uint32_t *cfg_dw;
    const double *xtab;
    const CUtensorMap *tmaps;           // pass 2: one 3-D tensor map per pair over x
    const CUtensorMap *tmaps_y;         // pass 2: the same over y (L2 prefetch of the read-modify-write operand)
    double *dot_out;                    // pass 2: per-CTA partial <x, y> (nullptr: not wanted)
};

// Long fibers (7-8 levels per star): 8 warps (7 consumers + producer) = 2 per scheduler -> 255 registers per thread and NO
// spills (local memory has no L1 to live in next to 226 KB of shared memory: every spill reload is an L2 round trip, which
// made the 12-warp / 168-register build latency-bound); short fibers: 16 warps, 128 registers.
#ifndef EDGPU_FIB_NC_BIG
#define EDGPU_FIB_NC_BIG 224
#endif
template <int NL> struct FibCfg { static constexpr int NC = (NL >= 7) ? EDGPU_FIB_NC_BIG : 480; static constexpr int NT = NC + 32; };

// ---- pass 1: one fiber (row r4 of the band, outer index o) of the up spin ----
// Two phases so that only HALF of the fiber lives in registers at a time (a whole 70-element fiber plus the slot
// registers exceeds the 168 registers available at 3 warps per scheduler): phase 0 holds the imp=1 half and produces the
// imp=0 outputs, phase 1 the other way round; the own element (diagonal term) is re-read from the image with the gathers.
template <int NL, int M0, int HS, int PART>
__device__ __forceinline__ void fiber_up_phase(const FibArgs &A, uint32_t own /* img + oE */, uint32_t ownO, const uint32_t (&sE)[HS], const uint32_t (&sO)[HS],
                                               const double (&amp)[HS], double dg, double sig, double *yE, double *yO)
{
    constexpr int NB = NL - 1, D0 = fib::cbinom(NL, M0), A0 = fib::cbinom(NB, M0), NP = (D0 + 1) / 2, COFF = fib::ccoff(NL, M0);
    constexpr int IN_LO = PART == 0 ? (A0 & ~1) : 0, IN_HI = PART == 0 ? 2 * NP : ((A0 + 1) & ~1), NIN = IN_HI - IN_LO;
    constexpr int OUT_LO = PART == 0 ? 0 : A0, OUT_HI = PART == 0 ? A0 : D0;
    if constexpr (NIN > 0 && OUT_HI > OUT_LO) {
        double in[NIN];
        fib::static_for<NIN / 2>([&](auto jj) {
            constexpr int K = IN_LO + 2 * decltype(jj)::value;
            const double2 v = flds128(((K & 3) == 0 ? own : ownO) + (uint32_t)(K >> 2) * 128u);
            in[K - IN_LO] = v.x; in[K - IN_LO + 1] = v.y;
        });
        constexpr int P_LO = OUT_LO / 2, P_HI = (OUT_HI + 1) / 2;
        fib::static_for<P_HI - P_LO>([&](auto jj) {
            constexpr int K = 2 * (P_LO + decltype(jj)::value);
            constexpr bool do0 = K >= OUT_LO && K < OUT_HI, do1 = K + 1 >= OUT_LO && K + 1 < OUT_HI;
            const uint32_t rel = (uint32_t)(K >> 2) * 128u;
            const double2 xo = flds128(((K & 3) == 0 ? own : ownO) + rel);
            double g0 = 0.0, g1 = 0.0;
#pragma unroll
            for (int s = 0; s < HS; s++) {
                const double2 v = flds128(((K & 3) == 0 ? sE[s] : sO[s]) + rel);
                if (do0) g0 = fma(amp[s], v.x, g0);
                if (do1) g1 = fma(amp[s], v.y, g1);
            }
            double r0 = 0.0, r1 = 0.0;
            if constexpr (do0) {
                const double in0 = fib::out<NB, M0, K, IN_LO>(in, A.cst.v0);
                r0 = fma(dg + A.cst.e0[COFF + K], xo.x, fma(sig, in0, PART == 0 ? g0 : -g0));
            }
            if constexpr (do1) {
                const double in1 = fib::out<NB, M0, K + 1, IN_LO>(in, A.cst.v0);
                r1 = fma(dg + A.cst.e0[COFF + K + 1], xo.y, fma(sig, in1, PART == 0 ? g1 : -g1));
            }
            // a pair that straddles the imp=0 / imp=1 boundary (A0 odd): each phase stores its own element
            double *yp = ((K & 3) == 0 ? yE : yO) + (K >> 2) * 16;
            if constexpr (do0 && do1) fstg128(yp, r0, r1);
            else if constexpr (do0 && K + 1 >= D0) fstg128(yp, r0, 0.0);        // last pair of an odd fiber: the pad stays zero
            else if constexpr (do0) yp[0] = r0;
            else if constexpr (do1) yp[1] = r1;
        });
    }
}

// Slot data of one fiber: from the shared-memory copy of the block's outer table (stab != 0: address of the SOuter entry)
// or from the global table (blocks whose table does not fit the copy).
struct FiberMeta {
    uint32_t stab;                    // shared address of the SOuter entry, 0 = use `ent`
    const OuterEnt *ent;
    __device__ __forceinline__ int nslot() const { return stab ? flds32(stab + 12u) : ent->nslot; }
    __device__ __forceinline__ int impbits() const { return stab ? flds32(stab + 8u) : ent->impbits; }
    __device__ __forceinline__ int neg() const { return stab ? flds32(stab + 16u) : ent->neg; }
    __device__ __forceinline__ double eo() const { return stab ? flds64(stab) : ent->eo; }
    __device__ __forceinline__ int delta(int s2) const { return stab ? flds32(stab + 24u + 4u * (uint32_t)s2) : ent->delta[s2]; }
    __device__ __forceinline__ double amp(int s2, const double *__restrict__ amps) const
    {
        return stab ? flds64(stab + 56u + 8u * (uint32_t)s2) : __ldg(amps + ent->code[s2]);
    }
};

template <int NL, int M0, int HS, int PART>
__device__ __forceinline__ void fiber_up(const FibArgs &A, const FiberMeta &F, int nslot, bool active, uint32_t img, int r4, int o,
                                         int d0p, double dgbase, uint32_t impd, double *yband)
{
    if (!active) return;
    // addresses: element k of the fiber sits at column c = o*d0p + k; c0 = o*d0p is 0 or 2 (mod 4)
    const int c0 = o * d0p;
    const uint32_t T = (uint32_t)(c0 >> 2) * 128u + (uint32_t)r4 * 32u;
    const bool q2 = (c0 & 3) != 0;
    const uint32_t oE = q2 ? T + 16u : T, oO = q2 ? T + 128u : T + 16u;          // k == 0 / 2 (mod 4); plus (k/4)*128
    uint32_t sE[HS], sO[HS];
    double amp[HS];
#pragma unroll
    for (int s = 0; s < HS; s++) {
        const bool on = s < nslot;
        const int cs = (o + (on ? F.delta(s) : 0)) * d0p;
        const uint32_t Ts = (uint32_t)(cs >> 2) * 128u + (uint32_t)r4 * 32u;
        const bool qs = (cs & 3) != 0;
        sE[s] = img + (qs ? Ts + 16u : Ts);
        sO[s] = img + (qs ? Ts + 128u : Ts + 16u);
        amp[s] = on ? F.amp(s, A.amps) : 0.0;
    }
    const int ib = F.impbits() | PART;                    // PART 1: the outputs have the impurity of star 0 occupied
    const int nimp = __popc(ib);
    const double dg = dgbase + F.eo() + __ldg(A.xtab + impd * 32u + (uint32_t)ib) + A.cst.pair_e * (double)(nimp * (nimp - 1) / 2);
    const double sig = F.neg() ? -1.0 : 1.0;
    double *yE = yband + (oE >> 3), *yO = yband + (oO >> 3);
    fiber_up_phase<NL, M0, HS, PART>(A, img + oE, img + oO, sE, sO, amp, dg, sig, yE, yO);
}

// consumers: copy the outer table of block `B` into shared memory (amplitudes resolved); named barrier 1 = consumers only
template <int NC>
__device__ __forceinline__ bool load_stab(const FibArgs &A, const FibBlockDev &B, uint32_t stab0, int tid)
{
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");                  // nobody still reads the previous table
    const bool fits = B.hsmax <= kStabHS && (size_t)B.nouter * sizeof(SOuter) <= (size_t)kStab;
    if (fits) {
        for (int o = tid; o < B.nouter; o += NC) {
            const OuterEnt *e = A.outer + B.tab + o;
            const uint32_t d = stab0 + (uint32_t)o * (uint32_t)sizeof(SOuter);
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(d), "d"(e->eo) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d + 8u), "r"(e->impbits), "r"(e->nslot) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d + 16u), "r"(e->neg), "r"(0) : "memory");
#pragma unroll
            for (int q = 0; q < kStabHS; q++) {
                const bool on = q < e->nslot;
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(d + 24u + 4u * q), "r"(on ? e->delta[q] : 0) : "memory");
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(d + 56u + 8u * q), "d"(on ? __ldg(A.amps + e->code[q]) : 0.0) : "memory");
            }
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
    return fits;
}

template <int NL>
__global__ void __launch_bounds__(FibCfg<NL>::NT) k_fib_up(const __grid_constant__ FibArgs A)
{
    constexpr int NC = FibCfg<NL>::NC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_bar[4];
    const int tid = threadIdx.x;
    const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t stab0 = slot0 + 2u * kSlot;
    const uint32_t bfull = (uint32_t)__cvta_generic_to_shared(s_bar), bempty = bfull + 16u;
    if (tid == 0) {
        fmbar_init(bfull, 1); fmbar_init(bfull + 8, 1);
        fmbar_init(bempty, NC / 32); fmbar_init(bempty + 8, NC / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int myn = (int)blockIdx.x < A.ntiles ? (A.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid >= NC) {
        if (tid == NC) {
            // ne[s]: tiles that have occupied the memory of slot s so far (= phases its empty barrier must have completed);
            // a two-slot tile occupies both
            int ne[2] = {0, 0}, pos = 0;
            for (int i = 0; i < myn; i++) {
                const FibTile t = A.tiles[blockIdx.x + (size_t)i * gridDim.x];
                const bool two = t.bytes > A.slot;
                const int s = two ? 0 : pos;
                if (ne[s] > 0) fmbar_wait_backoff(bempty + 8 * s, (uint32_t)(ne[s] - 1) & 1u);
                if (two && ne[1] > 0) fmbar_wait_backoff(bempty + 8, (uint32_t)(ne[1] - 1) & 1u);
                fmbar_expect_tx(bfull + 8 * s, (uint32_t)t.bytes);
                const uint32_t dst = slot0 + (uint32_t)s * (uint32_t)A.slot;
                const char *src = reinterpret_cast<const char *>(A.x + t.off);
                for (int ofs = 0; ofs < t.bytes; ofs += 32768) {
                    const int n = t.bytes - ofs < 32768 ? t.bytes - ofs : 32768;
                    fbulk_g2s(dst + (uint32_t)ofs, src + ofs, (uint32_t)n, bfull + 8 * s);
                }
                ne[s]++;
                if (two) ne[1]++;
                else pos ^= 1;
                // the tile after next cannot be loaded before a buffer frees: pull it into L2 meanwhile
                if (i + 1 < myn) {
                    const FibTile tn = A.tiles[blockIdx.x + (size_t)(i + 1) * gridDim.x];
                    const bool need = two || tn.bytes > A.slot || i + 2 < myn;
                    const FibTile tp = (two || tn.bytes > A.slot || i + 2 >= myn) ? tn : A.tiles[blockIdx.x + (size_t)(i + 2) * gridDim.x];
                    if (need) {
                        const char *ps = reinterpret_cast<const char *>(A.x + tp.off);
                        for (int ofs = 0; ofs < tp.bytes; ofs += 32768)
                            fbulk_prefetch_l2(ps + ofs, (uint32_t)(tp.bytes - ofs < 32768 ? tp.bytes - ofs : 32768));
                    }
                }
            }
        }
        return;
    }
    int nfill[2] = {0, 0}, pos = 0, cur_blk = -1;
    bool stab = false;
    for (int i = 0; i < myn; i++) {
        const FibTile t = A.tiles[blockIdx.x + (size_t)i * gridDim.x];
        const bool two = t.bytes > A.slot;
        const int s = two ? 0 : pos;
        const FibBlockDev BU = A.blk_f[t.blk];
        const PairDev pd = A.pairs[t.pair];
        const FibBlockDev BD = A.blk_o[pd.bi];
        if (t.blk != cur_blk) { stab = load_stab<NC>(A, BU, stab0, tid); cur_blk = t.blk; }
        // work units of a tile: [band g][phase][fiber], a warp takes 32 fibers of ONE phase (lanes = 4 rows x 8 outer indices,
        // a quarter-warp = 4 rows x 2 neighbouring outer indices: conflict-free LDS.128)
        const int noE = (BU.nouter + 1) & ~1, nwf = (4 * noE + 31) >> 5, nwu = 2 * nwf * t.b;
        const uint32_t band_bytes = (uint32_t)BU.C4 * 128u;
        const int warp = tid >> 5, lane = tid & 31;
        auto decode = [&](int wu, int &g, int &part, int &r4, int &o) {
            g = wu / (2 * nwf);
            const int rem = wu - g * 2 * nwf;
            part = rem / nwf;
            const int fb = (rem - part * nwf) * 32 + lane;
            r4 = fb & 3; o = 2 * (fb >> 3) + ((fb >> 2) & 1);
        };
        // per-row terms of the first unit, issued before the wait for the image
        double dg_first = 0.0;
        uint32_t impd_first = 0;
        if (warp < nwu) {
            int g, part, r4, o;
            decode(warp, g, part, r4, o);
            const int rp = (t.a + g) * 4 + r4;
            const int od = rp / BD.d0r, kd = rp - od * BD.d0r;
            if (od < BD.nouter && kd < BD.D0) {
                const int id = BD.off + od * BD.D0 + kd;
                dg_first = __ldg(A.e_dw + id);
                impd_first = __ldg(A.cfg_dw + id) & A.impmask;
            }
        }
        fmbar_wait_warp(bfull + 8 * s, (uint32_t)nfill[s] & 1u);
        for (int wu = warp; wu < ((A.dbg & 1) ? 0 : nwu); wu += NC / 32) {
            int g, part, r4, o;
            decode(wu, g, part, r4, o);
            const int rp = (t.a + g) * 4 + r4;
            const int od = rp / BD.d0r, kd = rp - od * BD.d0r;
            const bool active = o < BU.nouter && od < BD.nouter && kd < BD.D0;
            FiberMeta F;
            F.stab = stab ? stab0 + (uint32_t)(active ? o : 0) * (uint32_t)sizeof(SOuter) : 0u;
            F.ent = A.outer + BU.tab + (active ? o : 0);
            const int nslot = active ? F.nslot() : 0;
            const int wmax = __reduce_max_sync(0xffffffffu, nslot);
            double dgbase = dg_first;
            uint32_t impd = impd_first;
            if (wu != warp && active) {
                const int id = BD.off + od * BD.D0 + kd;
                dgbase = __ldg(A.e_dw + id);
                impd = __ldg(A.cfg_dw + id) & A.impmask;
            }
            const uint32_t img = slot0 + (uint32_t)s * (uint32_t)A.slot + (uint32_t)g * band_bytes;
            double *yband = A.y + t.off + (int64_t)g * BU.C4 * 16;
            fib::static_for<NL - 1>([&](auto mm) {
                constexpr int M0 = decltype(mm)::value + 1;
                if (BU.m0 == M0) {
                    if (part == 0) {
                        if (wmax <= 4) fiber_up<NL, M0, 4, 0>(A, F, nslot, active, img, r4, o, BU.d0p, dgbase, impd, yband);
                        else if (NL >= 7 || wmax <= 7) fiber_up<NL, M0, 7, 0>(A, F, nslot, active, img, r4, o, BU.d0p, dgbase, impd, yband);
                        else if constexpr (NL < 7) fiber_up<NL, M0, kHS, 0>(A, F, nslot, active, img, r4, o, BU.d0p, dgbase, impd, yband);
                    } else {
                        if (wmax <= 4) fiber_up<NL, M0, 4, 1>(A, F, nslot, active, img, r4, o, BU.d0p, dgbase, impd, yband);
                        else if (NL >= 7 || wmax <= 7) fiber_up<NL, M0, 7, 1>(A, F, nslot, active, img, r4, o, BU.d0p, dgbase, impd, yband);
                        else if constexpr (NL < 7) fiber_up<NL, M0, kHS, 1>(A, F, nslot, active, img, r4, o, BU.d0p, dgbase, impd, yband);
                    }
                }
            });
        }
        __syncwarp();
        if ((tid & 31) == 0) { fmbar_arrive(bempty + 8 * s); if (two) fmbar_arrive(bempty + 8); }
        nfill[s]++;
        if (!two) pos ^= 1;
    }
}

// ---- pass 2: one column (c4 of the strip) of the fiber (outer index o) of the down spin; PART 0: outputs with imp=0 ----
template <int NL, int M0, int HS, int PART>
__device__ __forceinline__ void fiber_dw(const FibArgs &A, const FiberMeta &F, int nslot, bool active, uint32_t img, int c4, int o,
                                         int d0r, double *ystrip, int64_t bstride, double &dsum)
{
    constexpr int NB = NL - 1, D0 = fib::cbinom(NL, M0), A0 = fib::cbinom(NB, M0), B0 = D0 - A0;
    constexpr int NIN = PART == 0 ? B0 : A0, NOUT = PART == 0 ? A0 : B0, IN0 = PART == 0 ? A0 : 0, OUT0 = PART == 0 ? 0 : A0;
    if (!active) return;
    if constexpr (NIN > 0 && NOUT > 0) {
        const int r0 = o * d0r;
        // global rows of the outputs: row r0 + k lies in band (r0+k)/4 at sub-row (r0+k)%4; Q[j] serves k == j (mod 4)
        double *Q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) Q[j] = ystrip + (int64_t)((r0 + j) >> 2) * bstride + ((r0 + j) & 3) * 4 + c4;
        // y of the outputs, prefetched chunk by chunk (double buffered in registers: CH loads in flight per thread);
        // the producer has pulled the y tile into L2 together with the x image
        constexpr int CH = 9, NCH = (NOUT + CH - 1) / CH;
        double yv[2][CH];
        auto prefetch = [&](auto cc) {
            constexpr int c = decltype(cc)::value;
            fib::static_for<CH>([&](auto ii) {
                constexpr int I = decltype(ii)::value, K = OUT0 + c * CH + I;
                if constexpr (c < NCH && c * CH + I < NOUT) yv[c & 1][I] = fldg64(Q[K & 3] + (int64_t)(K >> 2) * bstride);
            });
        };
        prefetch(std::integral_constant<int, 0>{});
        const uint32_t base = img + (uint32_t)r0 * 32u + (uint32_t)c4 * 8u;
        double in[NIN];
        fib::static_for<NIN>([&](auto jj) { constexpr int J = decltype(jj)::value; in[J] = flds64(base + (uint32_t)(IN0 + J) * 32u); });
        uint32_t sb[HS];
        double amp[HS];
#pragma unroll
        for (int s = 0; s < HS; s++) {
            const bool on = s < nslot;
            sb[s] = img + (uint32_t)((o + (on ? F.delta(s) : 0)) * d0r) * 32u + (uint32_t)c4 * 8u;
            amp[s] = on ? F.amp(s, A.amps) : 0.0;
        }
        const double sig = F.neg() ? -1.0 : 1.0;
        fib::static_for<NCH>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            prefetch(std::integral_constant<int, c + 1>{});
            fib::static_for<CH>([&](auto ii) {
                constexpr int I = decltype(ii)::value, K = OUT0 + c * CH + I;
                if constexpr (c * CH + I < NOUT) {
                    double g = 0.0;
#pragma unroll
                    for (int s = 0; s < HS; s++) g = fma(amp[s], flds64(sb[s] + (uint32_t)K * 32u), g);
                    const double inr = fib::out<NB, M0, K, IN0>(in, A.cst.v0);
                    const double r = yv[c & 1][I] + fma(sig, inr, PART == 0 ? g : -g);
                    Q[K & 3][(int64_t)(K >> 2) * bstride] = r;
                    if (A.dot_out) dsum = fma(flds64(base + (uint32_t)K * 32u), r, dsum);
                }
            });
        });
    }
}

template <int NL>
__global__ void __launch_bounds__(FibCfg<NL>::NT) k_fib_dw(const __grid_constant__ FibArgs A)
{
    constexpr int NC = FibCfg<NL>::NC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_bar[4];
    __shared__ double s_dot[NC / 32];
    const int tid = threadIdx.x;
    const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t stab0 = slot0 + 2u * kSlot;
    const uint32_t bfull = (uint32_t)__cvta_generic_to_shared(s_bar), bempty = bfull + 16u;
    if (tid == 0) {
        fmbar_init(bfull, 1); fmbar_init(bfull + 8, 1);
        fmbar_init(bempty, NC / 32); fmbar_init(bempty + 8, NC / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int myn = (int)blockIdx.x < A.ntiles ? (A.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid >= NC) {
        if (tid == NC) {
            int ne[2] = {0, 0}, pos = 0;
            auto boxes = [&](const FibTile &t, const FibBlockDev &BD, auto &&fn) {
                for (int g = 0; g < t.b; g++)
                    for (int b = 0; b < BD.nbox; b++) fn(g, b);
            };
            for (int i = 0; i < myn; i++) {
                const FibTile t = A.tiles[blockIdx.x + (size_t)i * gridDim.x];
                const FibBlockDev BD = A.blk_f[t.blk];
                const bool two = t.bytes > A.slot;
                const int s = two ? 0 : pos;
                if (ne[s] > 0) fmbar_wait_backoff(bempty + 8 * s, (uint32_t)(ne[s] - 1) & 1u);
                if (two && ne[1] > 0) fmbar_wait_backoff(bempty + 8, (uint32_t)(ne[1] - 1) & 1u);
                fmbar_expect_tx(bfull + 8 * s, (uint32_t)t.bytes);
                const uint32_t dst = slot0 + (uint32_t)s * (uint32_t)A.slot;
                const uint32_t sbytes = (uint32_t)BD.nbox * (uint32_t)BD.BR * 128u;
                boxes(t, BD, [&](int g, int b) {
                    ftma_load_3d(dst + (uint32_t)g * sbytes + (uint32_t)b * (uint32_t)BD.BR * 128u, A.tmaps + t.pair, 0, t.a + g, b * BD.BR, bfull + 8 * s);
                });
                // the read-modify-write operand of THIS tile goes to L2 while its x image lands ...
                boxes(t, BD, [&](int g, int b) { ftma_prefetch_3d(A.tmaps_y + t.pair, 0, t.a + g, b * BD.BR); });
                ne[s]++;
                if (two) ne[1]++;
                else pos ^= 1;
                // ... and so does the x image of the tile that cannot be loaded yet
                if (i + 1 < myn) {
                    const FibTile tn = A.tiles[blockIdx.x + (size_t)(i + 1) * gridDim.x];
                    const bool nextwaits = two || tn.bytes > A.slot;
                    if (nextwaits || i + 2 < myn) {
                        const FibTile tp = nextwaits ? tn : A.tiles[blockIdx.x + (size_t)(i + 2) * gridDim.x];
                        const FibBlockDev BP = A.blk_f[tp.blk];
                        boxes(tp, BP, [&](int g, int b) { ftma_prefetch_3d(A.tmaps + tp.pair, 0, tp.a + g, b * BP.BR); });
                    }
                }
            }
        }
        return;
    }
    double dsum = 0.0;
    int nfill[2] = {0, 0}, pos = 0, cur_blk = -1;
    bool stab = false;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = 0; i < myn; i++) {
        const FibTile t = A.tiles[blockIdx.x + (size_t)i * gridDim.x];
        const bool two = t.bytes > A.slot;
        const int s = two ? 0 : pos;
        const FibBlockDev BD = A.blk_f[t.blk];
        const PairDev pd = A.pairs[t.pair];
        const int C4 = A.blk_o[pd.bj].C4;
        if (t.blk != cur_blk) { stab = load_stab<NC>(A, BD, stab0, tid); cur_blk = t.blk; }
        fmbar_wait_warp(bfull + 8 * s, (uint32_t)nfill[s] & 1u);
        const int64_t bstride = (int64_t)C4 * 16;
        const uint32_t sbytes = (uint32_t)BD.nbox * (uint32_t)BD.BR * 128u;
        // warp-fibers of a strip: [part][8 outer indices per warp]; lanes = (c4 = lane & 3, o = 8*ow + lane/4)
        const int now = (BD.nouter + 7) >> 3, nwf = 2 * now * t.b;
        for (int wf = warp; wf < ((A.dbg & 1) ? 0 : nwf); wf += NC / 32) {
            const int g = wf / (2 * now), rem = wf - g * 2 * now;
            const int part = rem / now, ow = rem - part * now;
            const int c4 = lane & 3, o = 8 * ow + (lane >> 2);
            const bool active = o < BD.nouter;
            FiberMeta F;
            F.stab = stab ? stab0 + (uint32_t)(active ? o : 0) * (uint32_t)sizeof(SOuter) : 0u;
            F.ent = A.outer + BD.tab + (active ? o : 0);
            const int nslot = active ? F.nslot() : 0;
            const int wmax = __reduce_max_sync(0xffffffffu, nslot);
            const uint32_t img = slot0 + (uint32_t)s * (uint32_t)A.slot + (uint32_t)g * sbytes;
            double *ystrip = A.y + pd.base + (int64_t)(t.a + g) * 16;
            fib::static_for<NL - 1>([&](auto mm) {
                constexpr int M0 = decltype(mm)::value + 1;
                if (BD.m0 == M0) {
                    if (part == 0) {
                        if (wmax <= 4) fiber_dw<NL, M0, 4, 0>(A, F, nslot, active, img, c4, o, BD.d0r, ystrip, bstride, dsum);
                        else if (NL >= 7 || wmax <= 7) fiber_dw<NL, M0, 7, 0>(A, F, nslot, active, img, c4, o, BD.d0r, ystrip, bstride, dsum);
                        else if constexpr (NL < 7) fiber_dw<NL, M0, kHS, 0>(A, F, nslot, active, img, c4, o, BD.d0r, ystrip, bstride, dsum);
                    } else {
                        if (wmax <= 4) fiber_dw<NL, M0, 4, 1>(A, F, nslot, active, img, c4, o, BD.d0r, ystrip, bstride, dsum);
                        else if (NL >= 7 || wmax <= 7) fiber_dw<NL, M0, 7, 1>(A, F, nslot, active, img, c4, o, BD.d0r, ystrip, bstride, dsum);
                        else if constexpr (NL < 7) fiber_dw<NL, M0, kHS, 1>(A, F, nslot, active, img, c4, o, BD.d0r, ystrip, bstride, dsum);
                    }
                }
            });
        }
        __syncwarp();
        if (lane == 0) { fmbar_arrive(bempty + 8 * s); if (two) fmbar_arrive(bempty + 8); }
        nfill[s]++;
        if (!two) pos ^= 1;
    }
    if (A.dot_out) {
        for (int o = 16; o > 0; o >>= 1) dsum += __shfl_down_sync(0xffffffffu, dsum, o);
        if (lane == 0) s_dot[warp] = dsum;
        asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
        if (tid == 0) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NC / 32; w++) v += s_dot[w];
            A.dot_out[blockIdx.x] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// thread-per-element pair kernels (ELL hop tables of the generic kernel): blocks without a fiber kernel
// ------------------------------------------------------------------------------------------------------------
struct PairGenArgs {
    const PairDev *pairs; const int *list; int nlist;
    const FibBlockDev *blk_u, *blk_d;
    VAddr va;
    const uint32_t *cfg_up, *cfg_dw; const double *e_up, *e_dw, *xtab;
    const uint32_t *hop; const uint8_t *nhop; const double *amp; int64_t hop_ld;     // of the spin that is applied
    uint32_t impmask;
    const double *x; double *y; double *dot_out;
};

// y = (diag + H_up) x on the listed pairs
__global__ void __launch_bounds__(256) k_pair_up(const PairGenArgs A)
{
    __shared__ double s_amp[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_amp[i] = A.amp[i];
    __syncthreads();
    for (int q = blockIdx.y; q < A.nlist; q += gridDim.y) {
        const PairDev pd = A.pairs[A.list[q]];
        const FibBlockDev BD = A.blk_d[pd.bi], BU = A.blk_u[pd.bj];
        const int64_t total = (int64_t)BD.size * BU.size;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            const int64_t id = BD.off + e / BU.size, iu = BU.off + e % BU.size;
            const int64_t a = A.va(id, iu);
            double acc = (A.e_up[iu] + A.e_dw[id] + A.xtab[(A.cfg_dw[id] & A.impmask) * 32u + (A.cfg_up[iu] & A.impmask)]) * A.x[a];
            const int nu = A.nhop[iu];
            for (int j = 0; j < nu; j++) {
                const uint32_t h = A.hop[(int64_t)j * A.hop_ld + iu];
                acc += s_amp[h & 255u] * A.x[A.va(id, h >> 8)];
            }
            A.y[a] = acc;
        }
    }
}

// y += H_dw x on the listed pairs (+ partial <x, y>)
__global__ void __launch_bounds__(256) k_pair_dw(const PairGenArgs A)
{
    __shared__ double s_amp[256], s_red[8];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_amp[i] = A.amp[i];
    __syncthreads();
    double dsum = 0.0;
    for (int q = blockIdx.y; q < A.nlist; q += gridDim.y) {
        const PairDev pd = A.pairs[A.list[q]];
        const FibBlockDev BD = A.blk_d[pd.bi], BU = A.blk_u[pd.bj];
        const int64_t total = (int64_t)BD.size * BU.size;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            const int64_t id = BD.off + e / BU.size, iu = BU.off + e % BU.size;
            const int64_t a = A.va(id, iu);
            double acc = A.y[a];
            const int nd = A.nhop[id];
            for (int j = 0; j < nd; j++) {
                const uint32_t h = A.hop[(int64_t)j * A.hop_ld + id];
                acc += s_amp[h & 255u] * A.x[A.va(h >> 8, iu)];
            }
            A.y[a] = acc;
            dsum = fma(A.x[a], acc, dsum);
        }
    }
    if (A.dot_out) {
        for (int o = 16; o > 0; o >>= 1) dsum += __shfl_down_sync(0xffffffffu, dsum, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = 0.0;
            for (int w = 0; w < 8; w++) v += s_red[w];
            A.dot_out[blockIdx.y * gridDim.x + blockIdx.x] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static int fib_ensure_smem(edgpu_ctx *ctx, const void *kern, size_t smem)
{
    static std::map<const void *, size_t> set;
    size_t &cur = set[kern];
    if (smem > cur) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    return 0;
}

// one 3-D tensor map per pair over the vector at `x`: dims (16 doubles of a micro-tile, C4 strips, R4 bands)
static int fib_tensor_maps(edgpu_sector *s, const double *x, const CUtensorMap **out)
{
    edgpu_ctx *ctx = s->ctx;
    PairLayout &P = *s->pl;
    auto it = P.tmaps.find(x);
    if (it != P.tmaps.end()) { *out = it->second; return 0; }
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
    const FibSpin &FU = *s->up->fib, &FD = *s->dw->fib;
    std::vector<CUtensorMap> maps(std::max<size_t>(1, P.pairs.size()));
    memset(maps.data(), 0, sizeof(CUtensorMap) * maps.size());
    for (size_t p = 0; p < P.pairs.size(); p++) {
        const PairDev &pd = P.pairs[p];
        const FibBlockDev &BD = FD.blocks[pd.bi], &BU = FU.blocks[pd.bj];
        if (!BD.fiber) continue;
        const cuuint64_t gdim[3] = {16, (cuuint64_t)BU.C4, (cuuint64_t)BD.R4};
        const cuuint64_t gstr[2] = {128, (cuuint64_t)BU.C4 * 128};
        const cuuint32_t box[3] = {16, 1, (cuuint32_t)BD.BR};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = fn(&maps[p], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(x + pd.base), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return edgpu_fail(ctx, "cuTensorMapEncodeTiled failed (%d) for pair (%d,%d): C4=%d R4=%d BR=%d", (int)r, pd.bi, pd.bj, BU.C4, BD.R4, BD.BR);
    }
    if (P.tmaps.size() > 64) {                       // bounded cache: a sector sees only a handful of vector buffers
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (auto &kv : P.tmaps) cudaFree(kv.second);
        P.tmaps.clear();
    }
    CUtensorMap *d = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d, sizeof(CUtensorMap) * maps.size()));
    CUDA_TRY(ctx, cudaMemcpyAsync(d, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));                      // `maps` is a stack-lifetime source
    P.tmaps[x] = d;
    *out = d;
    return 0;
}

// a vector buffer is released: its tensor maps must not be reused for a new allocation at the same address of another size
void pair_layout_forget(edgpu_sector *s, const double *x)
{
    if (!s || !s->pl) return;
    auto it = s->pl->tmaps.find(x);
    if (it == s->pl->tmaps.end()) return;
    cudaStreamSynchronize(s->ctx->stream);
    cudaFree(it->second);
    s->pl->tmaps.erase(it);
}

template <int NL>
static int launch_fiber(edgpu_sector *s, const double *x, double *y, double *dot, int *ndot)
{
    edgpu_ctx *ctx = s->ctx;
    PairLayout &P = *s->pl;
    const FibSpin &FU = *s->up->fib, &FD = *s->dw->fib;
    const size_t smem = 2 * (size_t)kSlot + kStab;
    int nd = 0;
    FibArgs A;
    memset(&A, 0, sizeof(A));
    A.pairs = P.d_pairs; A.x = x; A.y = y; A.impmask = (1u << ctx->ham.norb) - 1u; A.slot = P.slot; A.dbg = (ctx->par.reserved[0] >> 13) & 3;
    A.e_dw = s->dw->ediag; A.cfg_dw = s->dw->cfg; A.xtab = ctx->d_xtab;
    PairGenArgs G{};
    G.pairs = P.d_pairs; G.blk_u = FU.d_blocks; G.blk_d = FD.d_blocks; G.va = sector_vaddr(s);
    G.cfg_up = s->up->cfg; G.cfg_dw = s->dw->cfg; G.e_up = s->up->ediag; G.e_dw = s->dw->ediag; G.xtab = ctx->d_xtab;
    G.impmask = A.impmask; G.x = x; G.y = y;
    // ---- pass 1 ----
    if (P.n1 > 0) {
        if (int rc = fib_ensure_smem(ctx, (const void *)k_fib_up<NL>, smem)) return rc;
        A.cst = FU.cst; A.blk_f = FU.d_blocks; A.blk_o = FD.d_blocks; A.outer = FU.d_outer; A.amps = FU.d_amps;
        A.tiles = P.d_t1; A.ntiles = P.n1; A.tmaps = nullptr; A.dot_out = nullptr;
        const int grid = std::min(ctx->sm_count, P.n1);
        k_fib_up<NL><<<grid, FibCfg<NL>::NT, smem, ctx->stream>>>(A);
    }
    if (P.ng1 > 0) {
        G.list = P.d_g1; G.nlist = P.ng1; G.hop = s->up->hop; G.nhop = s->up->nhop; G.amp = s->up->amp; G.hop_ld = s->up->dim; G.dot_out = nullptr;
        const int64_t per = P.g1_elems / P.ng1 + 1;
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(64, (per + 255) / 256)), (unsigned)std::min(P.ng1, 4096));
        k_pair_up<<<grid, 256, 0, ctx->stream>>>(G);
    }
    // ---- pass 2 ----
    if (P.n2 > 0) {
        if (int rc = fib_ensure_smem(ctx, (const void *)k_fib_dw<NL>, smem)) return rc;
        const CUtensorMap *tm = nullptr, *tmy = nullptr;
        if (int rc = fib_tensor_maps(s, x, &tm)) return rc;
        if (int rc = fib_tensor_maps(s, y, &tmy)) return rc;
        A.cst = FD.cst; A.blk_f = FD.d_blocks; A.blk_o = FU.d_blocks; A.outer = FD.d_outer; A.amps = FD.d_amps;
        A.tiles = P.d_t2; A.ntiles = P.n2; A.tmaps = tm; A.tmaps_y = tmy;
        const int grid = std::min(ctx->sm_count, P.n2);
        A.dot_out = dot ? dot + nd : nullptr;
        k_fib_dw<NL><<<grid, FibCfg<NL>::NT, smem, ctx->stream>>>(A);
        if (dot) nd += grid;
    }
    if (P.ng2 > 0) {
        G.list = P.d_g2; G.nlist = P.ng2; G.hop = s->dw->hop; G.nhop = s->dw->nhop; G.amp = s->dw->amp; G.hop_ld = s->dw->dim;
        const int64_t per = P.g2_elems / P.ng2 + 1;
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(64, (per + 255) / 256)), (unsigned)std::min(P.ng2, 128));
        G.dot_out = dot ? dot + nd : nullptr;
        k_pair_dw<<<grid, 256, 0, ctx->stream>>>(G);
        if (dot) nd += (int)(grid.x * grid.y);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    if (nd > kDotSlots) return edgpu_fail(ctx, "fiber H*v: too many dot partials (%d)", nd);
    if (ndot) *ndot = nd;
    return 0;
}

int hxv_fiber(edgpu_sector *s, const double *x, double *y, double *dot, int *ndot)
{
    if (!s->pl) return edgpu_fail(s->ctx, "hxv_fiber: the sector is not in the pair-tile layout");
    switch (s->pl->nl) {
#ifndef EDGPU_FIB_ONLY_NL8                     // (development switch: compile one instantiation only)
        case 3: return launch_fiber<3>(s, x, y, dot, ndot);
        case 4: return launch_fiber<4>(s, x, y, dot, ndot);
        case 5: return launch_fiber<5>(s, x, y, dot, ndot);
        case 6: return launch_fiber<6>(s, x, y, dot, ndot);
        case 7: return launch_fiber<7>(s, x, y, dot, ndot);
#endif
        case 8: return launch_fiber<8>(s, x, y, dot, ndot);
    }
    return edgpu_fail(s->ctx, "hxv_fiber: %d levels per star are not instantiated", s->pl->nl);
}

int hxv_fiber_launches(const edgpu_sector *s)
{
    if (!s->pl) return 0;
    return (s->pl->n1 > 0) + (s->pl->ng1 > 0) + (s->pl->n2 > 0) + (s->pl->ng2 > 0);
}
