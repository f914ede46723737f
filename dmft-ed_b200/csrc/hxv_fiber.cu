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
#include "fiber_common.h"
#include <algorithm>
#include <cstring>

uint64_t edgpu_binom(int n, int k);
__global__ void k_poshop(int64_t dim, int maxhop, const uint32_t *__restrict__ hop, const int2 *__restrict__ info, uint32_t *__restrict__ out);

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
        // test hook (bit 17): the smallest multiple of 4 instead -- fewer pad columns (Norb = 3: 16 instead of 18 for 15
        // configurations, 20 instead of 22), paid with 2-way bank conflicts of the neighbour-fiber loads
        if (ctx->par.reserved[0] & 131072) fb.d0p = (fb.D0 + 3) & ~3;
        fb.R = fb.nouter * fb.d0r; fb.R4 = (fb.R + 3) / 4;
        fb.C = fb.nouter * fb.d0p; fb.C4 = (fb.C + 3) / 4;
        fb.nbox = (fb.R4 + 255) / 256;
        fb.BR = ((fb.R4 + fb.nbox - 1) / fb.nbox + 1) & ~1;       // even: a half-strip box (BR x 64 bytes) keeps the 128-byte alignment
        fb.hnbox = (fb.C4 + 255) / 256;
        fb.hbox = ((fb.C4 + fb.hnbox - 1) / fb.hnbox + 1) & ~1;   // even, for the same reason (hbox x 64 bytes per box)
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
        // fiber kernels: star 0 must have hops (1 <= m0 <= nl-1), the images must fit (as half tiles at least), the block
        // must be worth a tile
        const int64_t band_bytes = (int64_t)fb.hnbox * fb.hbox * 64, strip_bytes = (int64_t)fb.nbox * fb.BR * 64;
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
        // test hook: shrink the pipeline slot so that the largest HALF image of this sector needs BOTH slots (the path that
        // only the 8000-configuration blocks of Ns=18 take in production); the blocks below it run as two-slot full tiles
        // (the 4900-configuration blocks of Ns=16) or as one-slot tiles
        int64_t big = 0;
        for (const FibBlockDev &B : FU.blocks) if (B.fiber) big = std::max<int64_t>(big, (int64_t)B.hnbox * B.hbox * 64);
        for (const FibBlockDev &B : FD.blocks) if (B.fiber) big = std::max<int64_t>(big, (int64_t)B.nbox * B.BR * 64);
        if (big > 256) P->slot = (int)std::min<int64_t>(kSlot, ((big / 2 + 127) / 128) * 128);
    }
    const int slot = P->slot;
    // ownership: LPT over pair COSTS (deterministic: ties by pair index).  cost = stored elements, weighted by the pass
    // kernels the pair will get: a pass through the memory-order pair kernels is ~6x slower per element than a fiber pass (measured on Ns=18)
    std::vector<int64_t> psize((size_t)nbd * nbu), pcost((size_t)nbd * nbu);
    std::vector<int> order((size_t)nbd * nbu), owner((size_t)nbd * nbu, 0);
    for (int i = 0; i < nbd; i++)
        for (int j = 0; j < nbu; j++) {
            const FibBlockDev &BD = FD.blocks[i], &BU = FU.blocks[j];
            const bool fu = !gen1 && BU.fiber && (force_fiber || BU.size >= kFibMinBlock) && (int64_t)BU.hnbox * BU.hbox * 64 <= 2 * (int64_t)slot;
            const bool fd = !gen2 && BD.fiber && (force_fiber || BD.size >= kFibMinBlock) && (int64_t)BD.nbox * BD.BR * 64 <= 2 * (int64_t)slot;
            psize[(size_t)i * nbu + j] = (int64_t)BD.R4 * BU.C4 * 16;
            pcost[(size_t)i * nbu + j] = psize[(size_t)i * nbu + j] * ((fu ? 1 : 6) + (fd ? 1 : 6));
            order[(size_t)i * nbu + j] = i * nbu + j;
        }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pcost[a] > pcost[b]; });
    {
        std::vector<int64_t> load(nranks, 0);
        for (int p : order) {
            int best = 0;
            for (int r = 1; r < nranks; r++) if (load[r] < load[best]) best = r;
            owner[p] = best; load[best] += pcost[p];
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
    std::vector<FibTile> t1, t2, t1h, t2h;
    std::vector<int> g1, g2;
    for (int p = 0; p < (int)P->pairs.size(); p++) {
        const PairDev &pd = P->pairs[p];
        const FibBlockDev &BD = FD.blocks[pd.bi], &BU = FU.blocks[pd.bj];
        const bool fu = !gen1 && BU.fiber && (force_fiber || BU.size >= kFibMinBlock) && (int64_t)BU.hnbox * BU.hbox * 64 <= 2 * (int64_t)slot;
        const bool fd = !gen2 && BD.fiber && (force_fiber || BD.size >= kFibMinBlock) && (int64_t)BD.nbox * BD.BR * 64 <= 2 * (int64_t)slot;
        if (fu) {
            const int64_t bb = (int64_t)BU.C4 * 128;
            // a band does not fit the two slots: two half bands (rows 0-1, 2-3) per band.  (Bands between one and two slots stay
            // whole and take both slots: halving them so that they double-buffer was measured SLOWER on the 4900-configuration
            // blocks of Ns=16 -- 64-byte runs, 87 % lane use, a second launch: 0.55 ms against 0.40 ms for their share.)
            const bool halves = bb > 2 * (int64_t)slot || (ctx->par.reserved[0] & 32768) != 0 && bb > slot;
            int G = (int)std::max<int64_t>(1, std::min<int64_t>(8, slot / bb));
            if (halves) G = 1;
            for (int a = 0; a < BD.R4; a += G) {
                FibTile t; memset(&t, 0, sizeof(t));
                t.pair = p; t.blk = pd.bj; t.a = a; t.b = std::min(G, BD.R4 - a);
                t.q0 = BD.d0r; t.q1 = BD.nouter; t.q2 = BD.D0; t.q3 = BD.off;
                t.off = pd.base + (int64_t)a * BU.C4 * 16;
                if (!halves) { t.bytes = (int)(bb * t.b); t1.push_back(t); continue; }
                t.bytes = BU.hnbox * BU.hbox * 64;
                for (int h = 1; h <= 2; h++) { t.half = h; t1h.push_back(t); }
            }
        } else { g1.push_back(p); P->g1_elems += (int64_t)BD.size * BU.size; }
        if (fd) {
            const int64_t sb = (int64_t)BD.nbox * BD.BR * 128;
            // a strip does not fit the two slots: two half strips (columns 0-1, 2-3); see above
            const bool halves = sb > 2 * (int64_t)slot || (ctx->par.reserved[0] & 32768) != 0 && sb > slot;
            int G = (int)std::max<int64_t>(1, std::min<int64_t>(8, slot / sb));
            if (halves) G = 1;
            for (int a = 0; a < BU.C4; a += G) {
                FibTile t; memset(&t, 0, sizeof(t));
                t.pair = p; t.blk = pd.bi; t.a = a; t.b = std::min(G, BU.C4 - a);
                t.q0 = BU.C4;
                t.off = pd.base;
                if (!halves) { t.bytes = (int)(sb * t.b); t2.push_back(t); continue; }
                t.bytes = (int)(sb / 2);
                for (int h = 1; h <= 2; h++) { t.half = h; t2h.push_back(t); }
            }
        } else { g2.push_back(p); P->g2_elems += (int64_t)BD.size * BU.size; }
    }
    // big tiles first (a CTA's sequence is then descending: two-slot tiles precede one-slot tiles, see the pipeline)
    // (equal sizes: by block, so that a CTA meets long runs of the same block)
    auto bysize = [](const FibTile &a, const FibTile &b) { return a.bytes != b.bytes ? a.bytes > b.bytes : a.blk < b.blk; };
    std::stable_sort(t1.begin(), t1.end(), bysize);
    std::stable_sort(t2.begin(), t2.end(), bysize);
    std::stable_sort(t1h.begin(), t1h.end(), bysize);
    std::stable_sort(t2h.begin(), t2h.end(), bysize);
    P->n1 = (int)t1.size(); P->n2 = (int)t2.size(); P->ng1 = (int)g1.size(); P->ng2 = (int)g2.size();
    P->n1h = (int)t1h.size(); P->n2h = (int)t2h.size();
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
    CUDA_TRY(ctx, up(&P->d_t1h, t1h));
    CUDA_TRY(ctx, up(&P->d_t2h, t2h));
    CUDA_TRY(ctx, up(&P->d_g1, g1));
    CUDA_TRY(ctx, up(&P->d_g2, g2));
    if (!g1.empty() || !g2.empty()) {
        // memory-order pair kernels: inverse position maps per block and position-resolved hop tables
        std::vector<int> cps(nbu + 1, 0), rps(nbd + 1, 0);
        for (int j = 0; j < nbu; j++) cps[j + 1] = cps[j] + FU.blocks[j].C;
        for (int i = 0; i < nbd; i++) rps[i + 1] = rps[i] + FD.blocks[i].R;
        std::vector<int> icp((size_t)cps[nbu], -1), irp((size_t)rps[nbd], -1);
        for (int j = 0; j < nbu; j++) {
            const FibBlockDev &B = FU.blocks[j];
            for (int e = 0; e < B.size; e++) icp[(size_t)cps[j] + (e / B.D0) * B.d0p + e % B.D0] = B.off + e;
        }
        for (int i = 0; i < nbd; i++) {
            const FibBlockDev &B = FD.blocks[i];
            for (int e = 0; e < B.size; e++) irp[(size_t)rps[i] + (e / B.D0) * B.d0r + e % B.D0] = B.off + e;
        }
        CUDA_TRY(ctx, up(&P->d_idx_of_cp, icp));
        CUDA_TRY(ctx, up(&P->d_idx_of_rp, irp));
        CUDA_TRY(ctx, up(&P->d_cp_start, cps));
        CUDA_TRY(ctx, up(&P->d_rp_start, rps));
        const int mhu = std::max(1, s->up->maxhop), mhd = std::max(1, s->dw->maxhop);
        CUDA_TRY(ctx, cudaMalloc(&P->d_poshop_c, sizeof(uint32_t) * (size_t)s->dim_up * mhu));
        CUDA_TRY(ctx, cudaMalloc(&P->d_poshop_r, sizeof(uint32_t) * (size_t)s->dim_dw * mhd));
        k_poshop<<<(unsigned)((s->dim_up + 255) / 256), 256, 0, st>>>(s->dim_up, s->up->maxhop, s->up->hop, P->d_colinfo, P->d_poshop_c);
        k_poshop<<<(unsigned)((s->dim_dw + 255) / 256), 256, 0, st>>>(s->dim_dw, s->dw->maxhop, s->dw->hop, P->d_rowinfo, P->d_poshop_r);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    s->pl = P;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// thread-per-element pair kernels (ELL hop tables of the generic kernel): blocks without a fiber kernel
// ------------------------------------------------------------------------------------------------------------
struct PairGenArgs {
    const PairDev *pairs; const int *list; int nlist;
    const FibBlockDev *blk_u, *blk_d;
    const int *idx_of_cp, *idx_of_rp, *cp_start, *rp_start;
    const uint32_t *cfg_up, *cfg_dw; const double *e_up, *e_dw, *xtab;
    const uint32_t *poshop; const uint8_t *nhop; const double *amp; int64_t hop_ld;     // of the spin that is applied
    uint32_t impmask;
    const double *x; double *y; double *dot_out;
};

// Hop table with positions as targets: out[j*dim + i] = (position of target(in[j*dim + i]) << 8) | amplitude code
__global__ void k_poshop(int64_t dim, int maxhop, const uint32_t *__restrict__ hop, const int2 *__restrict__ info, uint32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    for (int j = 0; j < maxhop; j++) {
        const uint32_t h = hop[(int64_t)j * dim + i];
        out[(int64_t)j * dim + i] = ((uint32_t)info[h >> 8].y << 8) | (h & 255u);
    }
}

// The two kernels walk a pair in MEMORY order (thread = one element of a micro-tile: coalesced own accesses); the hop
// sources are gathered through L2 inside the same band (up) / the same strip (down).
// y = H_dw x on the listed pairs (first pass: every element of the pair is written; pads receive zero)
__global__ void __launch_bounds__(256) k_pair_dw(const PairGenArgs A)
{
    __shared__ double s_amp[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_amp[i] = A.amp[i];
    __syncthreads();
    for (int q = blockIdx.y; q < A.nlist; q += gridDim.y) {
        const PairDev pd = A.pairs[A.list[q]];
        const FibBlockDev BD = A.blk_d[pd.bi], BU = A.blk_u[pd.bj];
        const int *irp = A.idx_of_rp + A.rp_start[pd.bi];
        const int64_t C4 = BU.C4, ntile = (int64_t)BD.R4 * C4;
        for (int64_t tile = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4); tile < ntile; tile += (int64_t)gridDim.x * 16) {
            const int within = threadIdx.x & 15, r4 = within >> 2, c4 = within & 3;
            const int64_t band = tile / C4, ct = tile - band * C4;
            const int rp = (int)band * 4 + r4;
            const int id = rp < BD.R ? irp[rp] : -1;
            double acc = 0.0;
            if (id >= 0) {
                const int nd = A.nhop[id];
                for (int j = 0; j < nd; j++) {
                    const uint32_t h = A.poshop[(int64_t)j * A.hop_ld + id];
                    const int rp2 = (int)(h >> 8);
                    acc += s_amp[h & 255u] * A.x[pd.base + ((int64_t)(rp2 >> 2) * C4 + ct) * 16 + (rp2 & 3) * 4 + c4];
                }
            }
            A.y[pd.base + tile * 16 + within] = acc;
        }
    }
}

// y += (diag + H_up) x on the listed pairs (second pass; + partial <x, y>)
__global__ void __launch_bounds__(256) k_pair_up(const PairGenArgs A)
{
    __shared__ double s_amp[256], s_red[8];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_amp[i] = A.amp[i];
    __syncthreads();
    double dsum = 0.0;
    for (int q = blockIdx.y; q < A.nlist; q += gridDim.y) {
        const PairDev pd = A.pairs[A.list[q]];
        const FibBlockDev BD = A.blk_d[pd.bi], BU = A.blk_u[pd.bj];
        const int *irp = A.idx_of_rp + A.rp_start[pd.bi], *icp = A.idx_of_cp + A.cp_start[pd.bj];
        const int64_t C4 = BU.C4, ntile = (int64_t)BD.R4 * C4;
        for (int64_t tile = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4); tile < ntile; tile += (int64_t)gridDim.x * 16) {
            const int within = threadIdx.x & 15, r4 = within >> 2, c4 = within & 3;
            const int64_t band = tile / C4, ct = tile - band * C4;
            const int rp = (int)band * 4 + r4, cp = (int)ct * 4 + c4;
            const int id = rp < BD.R ? irp[rp] : -1, iu = cp < BU.C ? icp[cp] : -1;
            if (id < 0 || iu < 0) continue;
            const int64_t a = pd.base + tile * 16 + within, rowb = pd.base + band * C4 * 16 + r4 * 4;
            const double xo = A.x[a];
            double acc = A.y[a] + (A.e_up[iu] + A.e_dw[id] + A.xtab[(A.cfg_dw[id] & A.impmask) * 32u + (A.cfg_up[iu] & A.impmask)]) * xo;
            const int nu = A.nhop[iu];
            for (int j = 0; j < nu; j++) {
                const uint32_t h = A.poshop[(int64_t)j * A.hop_ld + iu];
                const int cp2 = (int)(h >> 8);
                acc += s_amp[h & 255u] * A.x[rowb + (int64_t)(cp2 >> 2) * 16 + (cp2 & 3)];
            }
            A.y[a] = acc;
            dsum = fma(xo, acc, dsum);
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
// tensor maps per pair over the vector at `x` (see PairLayout::tmaps): strips = dims (16 doubles of a micro-tile, C4, R4) with
// box 16 x 1 x BR; half bands = the same dims with box 8 x hbox x 1; half strips = dims (4 columns, 4 rows, C4, R4) with box
// 2 x 4 x 1 x BR
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
    const size_t np = std::max<size_t>(1, P.pairs.size());
    std::vector<CUtensorMap> maps(3 * np);
    memset(maps.data(), 0, sizeof(CUtensorMap) * maps.size());
    for (size_t p = 0; p < P.pairs.size(); p++) {
        const PairDev &pd = P.pairs[p];
        const FibBlockDev &BD = FD.blocks[pd.bi], &BU = FU.blocks[pd.bj];
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        void *base = const_cast<double *>(x + pd.base);
        CUresult r = CUDA_SUCCESS;
        if (BD.fiber) {
            const cuuint64_t gdim[3] = {16, (cuuint64_t)BU.C4, (cuuint64_t)BD.R4};
            const cuuint64_t gstr[2] = {128, (cuuint64_t)BU.C4 * 128};
            const cuuint32_t box[3] = {16, 1, (cuuint32_t)BD.BR};
            r = fn(&maps[p], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS && P.n2h > 0) {
                const cuuint64_t gdim4[4] = {4, 4, (cuuint64_t)BU.C4, (cuuint64_t)BD.R4};
                const cuuint64_t gstr4[3] = {32, 128, (cuuint64_t)BU.C4 * 128};
                const cuuint32_t box4[4] = {2, 4, 1, (cuuint32_t)BD.BR};
                r = fn(&maps[np + p], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, gdim4, gstr4, box4, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            }
        }
        if (r == CUDA_SUCCESS && BU.fiber && P.n1h > 0) {
            const cuuint64_t gdim[3] = {16, (cuuint64_t)BU.C4, (cuuint64_t)BD.R4};
            const cuuint64_t gstr[2] = {128, (cuuint64_t)BU.C4 * 128};
            const cuuint32_t box[3] = {8, (cuuint32_t)BU.hbox, 1};
            r = fn(&maps[2 * np + p], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) return edgpu_fail(ctx, "cuTensorMapEncodeTiled failed (%d) for pair (%d,%d): C4=%d R4=%d BR=%d hbox=%d", (int)r, pd.bi, pd.bj, BU.C4, BD.R4, BD.BR, BU.hbox);
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

#define FIB_DECL(n) int fib_launch_nl##n(int pass, cudaStream_t st, const FibArgs &A, int grid); int fib_launch_nl##n##h(int pass, cudaStream_t st, const FibArgs &A, int grid);
FIB_DECL(3) FIB_DECL(4) FIB_DECL(5) FIB_DECL(6) FIB_DECL(7) FIB_DECL(8)
#undef FIB_DECL

static int fib_launch_any(edgpu_ctx *ctx, int nl, int pass, bool half, const FibArgs &A, int grid)
{
    int e = (int)cudaErrorInvalidValue;
    switch (nl) {
#define FIB_CASE(n) case n: e = half ? fib_launch_nl##n##h(pass, ctx->stream, A, grid) : fib_launch_nl##n(pass, ctx->stream, A, grid); break;
        FIB_CASE(3) FIB_CASE(4) FIB_CASE(5) FIB_CASE(6) FIB_CASE(7) FIB_CASE(8)
#undef FIB_CASE
        default: return edgpu_fail(ctx, "hxv_fiber: %d levels per star are not instantiated", nl);
    }
    if (e != 0) return edgpu_fail(ctx, "fiber kernel launch (NL=%d, pass %d%s): %s", nl, pass, half ? ", half tiles" : "", cudaGetErrorString((cudaError_t)e));
    return 0;
}

int hxv_fiber(edgpu_sector *s, const double *x, double *y, double *dot, int *ndot)
{
    if (!s->pl) return edgpu_fail(s->ctx, "hxv_fiber: the sector is not in the pair-tile layout");
    edgpu_ctx *ctx = s->ctx;
    PairLayout &P = *s->pl;
    const FibSpin &FU = *s->up->fib, &FD = *s->dw->fib;
    int nd = 0;
    FibArgs A;
    memset(&A, 0, sizeof(A));
    A.pairs = P.d_pairs; A.x = x; A.y = y; A.impmask = (1u << ctx->ham.norb) - 1u; A.norb = ctx->ham.norb; A.slot = P.slot;
    A.dbg = ((ctx->par.reserved[0] >> 13) & 3) | (((ctx->par.reserved[0] >> 16) & 1) << 3);
    A.e_dw = s->dw->ediag; A.cfg_dw = s->dw->cfg; A.xtab = ctx->d_xtab;
    PairGenArgs G{};
    G.pairs = P.d_pairs; G.blk_u = FU.d_blocks; G.blk_d = FD.d_blocks;
    G.idx_of_cp = P.d_idx_of_cp; G.idx_of_rp = P.d_idx_of_rp; G.cp_start = P.d_cp_start; G.rp_start = P.d_rp_start;
    G.cfg_up = s->up->cfg; G.cfg_dw = s->dw->cfg; G.e_up = s->up->ediag; G.e_dw = s->dw->ediag; G.xtab = ctx->d_xtab;
    G.impmask = A.impmask; G.x = x; G.y = y;
    const size_t np = std::max<size_t>(1, P.pairs.size());
    const CUtensorMap *tm = nullptr, *tmy = nullptr;
    if (P.n2 + P.n2h + P.n1h > 0)
        if (int rc = fib_tensor_maps(s, x, &tm)) return rc;
    if (P.n1h > 0)
        if (int rc = fib_tensor_maps(s, y, &tmy)) return rc;
    // dot partials: [fiber up][fiber up, half tiles][pair kernels]
    const int g1 = P.n1 > 0 ? std::min(ctx->sm_count, P.n1) : 0, g1h = P.n1h > 0 ? std::min(ctx->sm_count, P.n1h) : 0;
    // ---- thread-per-element pair kernels (blocks without a fiber kernel) on a side stream, so that they fill the tails of the
    //      persistent fiber kernels.  A pair may take one pass from each family (fiber up block, small down block): the down
    //      kernels of BOTH families precede the up kernels of both ----
    cudaStream_t gs = ctx->stream;
    const bool side = (P.ng1 + P.ng2 > 0) && (P.n1 + P.n2 + P.n1h + P.n2h > 0) && !(ctx->par.reserved[0] & 524288);
    if (side) {
        if (!ctx->aux_stream) {
            CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_gdw, cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fdw, cudaEventDisableTiming));
        }
        gs = ctx->aux_stream;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(gs, ctx->ev_fork, 0));
    }
    auto generic_dw = [&]() {
        if (P.ng2 <= 0) return;
        G.list = P.d_g2; G.nlist = P.ng2; G.poshop = P.d_poshop_r; G.nhop = s->dw->nhop; G.amp = s->dw->amp; G.hop_ld = s->dw->dim; G.dot_out = nullptr;
        const int64_t per = P.g2_elems / P.ng2 + 1;
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count * 2, (per + 255) / 256)), (unsigned)std::min(P.ng2, 2048));
        k_pair_dw<<<grid, 256, 0, gs>>>(G);
    };
    int ndg = 0;
    auto generic_up = [&]() {
        if (P.ng1 <= 0) return;
        G.list = P.d_g1; G.nlist = P.ng1; G.poshop = P.d_poshop_c; G.nhop = s->up->nhop; G.amp = s->up->amp; G.hop_ld = s->up->dim;
        const int64_t per = P.g1_elems / P.ng1 + 1;
        const unsigned gy = (unsigned)std::min(P.ng1, 32);
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count * 2, (per + 255) / 256)), gy);
        G.dot_out = dot ? dot + g1 + g1h : nullptr;
        k_pair_up<<<grid, 256, 0, gs>>>(G);
        ndg = (int)(grid.x * grid.y);
    };
    if (side) {
        generic_dw();
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_gdw, gs));
    }
    // ---- first pass: y = H_dw x (strips of 4 columns; write-only) ----
    A.cst = FD.cst; A.blk_f = FD.d_blocks; A.blk_o = FU.d_blocks; A.outer = FD.d_outer; A.amps = FD.d_amps;
    A.tmaps_y = nullptr; A.dot_out = nullptr;
    if (P.n2 > 0) {
        A.tiles = P.d_t2; A.ntiles = P.n2; A.tmaps = tm;
        if (int rc = fib_launch_any(ctx, P.nl, 2, false, A, std::min(ctx->sm_count, P.n2))) return rc;
    }
    if (P.n2h > 0) {
        A.tiles = P.d_t2h; A.ntiles = P.n2h; A.tmaps = tm + np;
        if (int rc = fib_launch_any(ctx, P.nl, 2, true, A, std::min(ctx->sm_count, P.n2h))) return rc;
    }
    if (!side) generic_dw();
    else {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fdw, ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(gs, ctx->ev_fdw, 0));
        generic_up();
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, gs));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_gdw, 0));
    }
    // ---- second pass: y += (diag + H_up) x (bands of 4 rows; read-modify-write, partial <x, y>) ----
    A.cst = FU.cst; A.blk_f = FU.d_blocks; A.blk_o = FD.d_blocks; A.outer = FU.d_outer; A.amps = FU.d_amps;
    if (P.n1 > 0) {
        A.tiles = P.d_t1; A.ntiles = P.n1; A.tmaps = nullptr; A.tmaps_y = nullptr;
        A.dot_out = dot ? dot : nullptr;
        if (int rc = fib_launch_any(ctx, P.nl, 1, false, A, g1)) return rc;
    }
    if (P.n1h > 0) {
        A.tiles = P.d_t1h; A.ntiles = P.n1h; A.tmaps = tm + 2 * np; A.tmaps_y = tmy + 2 * np;
        A.dot_out = dot ? dot + g1 : nullptr;
        if (int rc = fib_launch_any(ctx, P.nl, 1, true, A, g1h)) return rc;
    }
    if (!side) generic_up();
    else CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    if (dot) nd = g1 + g1h + ndg;
    CUDA_TRY(ctx, cudaGetLastError());
    if (nd > kDotSlots) return edgpu_fail(ctx, "fiber H*v: too many dot partials (%d)", nd);
    if (ndot) *ndot = nd;
    return 0;
}

int hxv_fiber_launches(const edgpu_sector *s)
{
    if (!s->pl) return 0;
    return (s->pl->n1 > 0) + (s->pl->n1h > 0) + (s->pl->ng1 > 0) + (s->pl->n2 > 0) + (s->pl->n2h > 0) + (s->pl->ng2 > 0);
}
