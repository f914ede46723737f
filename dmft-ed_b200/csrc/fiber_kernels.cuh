// fiber_kernels.cuh -- device side of the fiber H*v engine (see hxv_fiber.cu for the design); included by fib_nl*.cu, one
// translation unit per number of levels of a star so that the instantiations compile in parallel.
#pragma once
#include "fiber_common.h"

// ------------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------------
#ifndef EDGPU_FIB_EVICT
#define EDGPU_FIB_EVICT 1          // streamed operands (x images, y stores) leave L2 first: prefetched bands survive until they are used
#endif
__device__ __forceinline__ uint64_t fpolicy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void fmbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void fmbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void fmbar_expect_tx_only(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
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
#ifndef EDGPU_FIB_EVICT_DW
#define EDGPU_FIB_EVICT_DW 0       // the same for the y stores of the down pass
#endif
#ifndef EDGPU_FIB_YPF
#define EDGPU_FIB_YPF 0            // up pass: every lane prefetches the y sectors of its next unit into L2 (measured slower: 1.64 vs 1.58 ms)
#endif
#ifndef EDGPU_FIB_PFLAST
#define EDGPU_FIB_PFLAST 0         // L2 prefetches of the up pass (y band, next two-slot x image) with the evict_last priority
#endif
__device__ __forceinline__ uint64_t fpolicy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void fbulk_prefetch_l2_hint(const void *src, uint32_t bytes, uint64_t pol)
{
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void fbulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void ftma_load_3d_hint(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void ftma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void ftma_load_4d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
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
__device__ __forceinline__ void fstg128_hint(double *p, double a, double b, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
}
#ifndef EDGPU_FIB_NOYLD
#define EDGPU_FIB_NOYLD 0          // measurement only (wrong results): the up pass does not load y / does not store y
#endif
#ifndef EDGPU_FIB_NOYST
#define EDGPU_FIB_NOYST 0
#endif
__device__ __forceinline__ double2 fldg128(const double *p)
{
    double2 v;
    if (EDGPU_FIB_NOYLD) return make_double2(0.0, 0.0);
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double fldg64(const double *p)
{
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Long fibers (7-8 levels per star): 8 warps (7 consumers + producer) = 2 per scheduler -> 255 registers per thread and NO
// spills (local memory has no L1 to live in next to 226 KB of shared memory: every spill reload is an L2 round trip, which
// made the 12-warp / 168-register build latency-bound); short fibers: 16 warps, 128 registers.
#ifndef EDGPU_FIB_YW
#define EDGPU_FIB_YW 8            // up pass: y loads in flight per unit (0 = all of them up front)
#endif
#ifndef EDGPU_FIB_YD
#define EDGPU_FIB_YD 4            // up pass: pairs between the arithmetic of a pair and its "+ y, store" stage
#endif
#ifndef EDGPU_FIB_NC_BIG
#define EDGPU_FIB_NC_BIG 224
#endif
// the down pass (write-only, one column per thread) fits 168 registers: 12 warps = 3 per scheduler
#ifndef EDGPU_FIB_NC_DW
#define EDGPU_FIB_NC_DW 352
#endif
// 6 levels per star (Norb = 3, Nbath = 5: up to 12 gather slots per fiber): 12 warps / 168 registers in both passes -- the
// 16-warp build spilled 136 bytes (Ns=18: 50.1 -> 45.6 ms)
template <int NL> struct FibCfg { static constexpr int NC = NL >= 7 ? EDGPU_FIB_NC_BIG : NL == 6 ? EDGPU_FIB_NC_DW : 480; static constexpr int NT = NC + 32; };
template <int NL> struct FibCfgDw { static constexpr int NC = NL >= 6 ? EDGPU_FIB_NC_DW : 480; static constexpr int NT = NC + 32; };

// HS variants (gather slots unrolled per fiber): exact for the long fibers, where every dummy slot costs 18 LDS + 36 DFMA
template <int NL> struct FibHS { static constexpr int n = NL >= 7 ? 4 : 3; };
template <int NL> __device__ __forceinline__ constexpr int fib_hs(int i)
{
    return NL >= 7 ? (i == 0 ? 4 : i == 1 ? 5 : i == 2 ? 6 : 7) : (i == 0 ? 4 : i == 1 ? 7 : kHS);
}

// ------------------------------------------------------------------------------------------------------------
// Shared-memory copy of the outer table of the block a CTA is working on (consumers only; named barrier 1).
// PASS 1: offsets in the micro-tiled band image (mt = bytes per micro-tile of the image: 128, half bands 64); PASS 2: row
// offsets in the strip image (mt = bytes per row: 32, half strips 16).  Blocks whose table does not fit (more than 70 outer
// indices or more than 7 slots: only Norb = 3) use the global table (slower set-up).
// ------------------------------------------------------------------------------------------------------------
template <int NC, int PASS>
__device__ __forceinline__ bool load_stab(const FibArgs &A, const FibBlockDev &B, uint32_t stab0, int tid, int mt)
{
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");                  // nobody still reads the previous table
    const bool fits = B.hsmax <= kStabHS && B.nouter * kSOuterBytes <= kStabXtab;
    if (fits) {
        for (int o = tid; o < B.nouter; o += NC) {
            const OuterEnt *e = A.outer + B.tab + o;
            const uint32_t d = stab0 + (uint32_t)o * (uint32_t)kSOuterBytes;
            const int ns = e->nslot;
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(d), "d"(e->eo) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d + 8u), "r"(e->impbits), "r"(ns | (e->neg << 8)) : "memory");
#pragma unroll
            for (int q = 0; q < kStabHS; q++) {
                const bool on = q < ns;
                const int o2 = o + (on ? e->delta[q] : 0);
                int relE, relO;
                if (PASS == 1) {
                    const int c = o2 * B.d0p;
                    const int T = (c >> 2) * mt;
                    relE = (c & 3) ? T + 16 : T;
                    relO = (c & 3) ? T + mt : T + 16;
                } else { relE = o2 * B.d0r * mt; relO = 0; }
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d + 16u + 16u * q), "r"(relE), "r"(relO) : "memory");
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(d + 24u + 16u * q), "d"(on ? __ldg(A.amps + e->code[q]) : 0.0) : "memory");
            }
        }
        if (PASS == 1) {
            const int nim = 1 << A.norb;
            for (int q = tid; q < nim * nim; q += NC)
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(stab0 + (uint32_t)kStabXtab + 8u * q), "d"(__ldg(A.xtab + (q / nim) * 32 + (q % nim))) : "memory");
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
    return fits;
}

// Slot data of one fiber, from the shared copy (stab != 0) or from the global table.
struct FiberMeta {
    uint32_t stab;                    // shared address of the entry, 0 = use `ent`
    const OuterEnt *ent;
    int o, stride, mt;                // global path: outer index, d0p (pass 1) / d0r (pass 2), image stride (see load_stab)
    __device__ __forceinline__ void head(double &eo, int &impbits, int &nslot, int &neg) const
    {
        if (stab) {
            double a; int b, c;
            asm("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(stab));
            asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(b), "=r"(c) : "r"(stab + 8u));
            eo = a; impbits = b; nslot = c & 255; neg = c >> 8;
        } else { eo = ent->eo; impbits = ent->impbits; nslot = ent->nslot; neg = ent->neg; }
    }
    template <int PASS>
    __device__ __forceinline__ void slot(int s2, bool on, const double *__restrict__ amps, int &relE, int &relO, double &amp) const
    {
        if (stab) {
            asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(relE), "=r"(relO) : "r"(stab + 16u + 16u * (uint32_t)s2));
            asm("ld.shared.f64 %0, [%1];" : "=d"(amp) : "r"(stab + 24u + 16u * (uint32_t)s2));
        } else {
            const int o2 = o + (on ? ent->delta[s2] : 0);
            if (PASS == 1) {
                const int c = o2 * stride, T = (c >> 2) * mt;
                relE = (c & 3) ? T + 16 : T;
                relO = (c & 3) ? T + mt : T + 16;
            } else { relE = o2 * stride * mt; relO = 0; }
            amp = on ? __ldg(amps + ent->code[s2]) : 0.0;
        }
    }
};

// ---- up pass (runs SECOND: y += (diag + H_up) x, fused <x,y>): one PHASE of one fiber (row r4 of the band, outer index o) ----
// PART 0 holds the imp=1 half of the fiber in registers and produces the imp=0 outputs, PART 1 the other way round (only half
// of a <= 70-element fiber lives in registers at a time); the own element (diagonal term) is re-read with the gathers.
// MT: bytes per micro-tile of the IMAGE (128: bands of 4 rows; 64: half bands of 2 rows); y is always in 128-byte micro-tiles.
template <int NL, int M0, int HS, int PART, int MT>
__device__ __forceinline__ void fiber_up(const FibArgs &A, const FiberMeta &F, int nslot, double eo, int impbits, int neg, uint32_t img4 /* image + r4*32 */,
                                         int o, int d0p, double dgbase, double xt, double *yband4 /* band + r4*4 */, double &dsum, uint64_t spol)
{
    constexpr int NB = NL - 1, D0 = fib::cbinom(NL, M0), A0 = fib::cbinom(NB, M0), NP = (D0 + 1) / 2, COFF = fib::ccoff(NL, M0);
    constexpr int IN_LO = PART == 0 ? (A0 & ~1) : 0, IN_HI = PART == 0 ? 2 * NP : ((A0 + 1) & ~1), NIN = IN_HI - IN_LO;
    constexpr int OUT_LO = PART == 0 ? 0 : A0, OUT_HI = PART == 0 ? A0 : D0;
    if constexpr (NIN > 0 && OUT_HI > OUT_LO) {
        // element k of the fiber sits at column c = o*d0p + k; c0 = o*d0p is 0 or 2 (mod 4)
        const int c0 = o * d0p;
        const uint32_t T = (uint32_t)(c0 >> 2) * (uint32_t)MT, Ty = (uint32_t)(c0 >> 2) * 16u;
        const bool q2 = (c0 & 3) != 0;
        const uint32_t oE = q2 ? T + 16u : T, oO = q2 ? T + (uint32_t)MT : T + 16u;      // k == 0 / 2 (mod 4); plus (k/4)*MT
        const uint32_t own = img4 + oE, ownO = img4 + oO;
        double *yE = yband4 + (q2 ? Ty + 2u : Ty), *yO = yband4 + (q2 ? Ty + 16u : Ty + 2u);
        constexpr int P_LO = OUT_LO / 2, P_HI = (OUT_HI + 1) / 2, NPR = P_HI - P_LO;
        // y += ... : the pass-1 result (H_dw x) comes through a window of W loads in flight, and the "+ y, store" stage of a
        // pair runs D pairs BEHIND its arithmetic (results wait in `pend`): the first loads of a unit then have the unit's
        // set-up plus D pairs of arithmetic to arrive -- with the add right behind the arithmetic the L2 latency of each
        // batch of loads was exposed (50 % of the stall samples of the fiber bodies sat on three DADDs per unit).
        // (the producer has pulled the y band into L2 together with the x image)
        constexpr int D = (EDGPU_FIB_YD) < NPR ? (EDGPU_FIB_YD) : NPR;          // 0: the add follows the arithmetic directly
        constexpr int W = (EDGPU_FIB_YW) > 0 && (EDGPU_FIB_YW) < NPR ? ((EDGPU_FIB_YW) > D ? (EDGPU_FIB_YW) : (D > 0 ? D : 1)) : NPR;
        double2 yw[W], pend[D > 0 ? D : 1];
        auto yaddr = [&](auto kk) -> double * {
            constexpr int K = decltype(kk)::value;
            return ((K & 3) == 0 ? yE : yO) + (K >> 2) * 16;
        };
        fib::static_for<W>([&](auto jj) {
            constexpr int J = decltype(jj)::value;
            yw[J] = fldg128(yaddr(std::integral_constant<int, 2 * (P_LO + J)>{}));
        });
        double in[NIN];
        fib::static_for<NIN / 2>([&](auto jj) {
            constexpr int K = IN_LO + 2 * decltype(jj)::value;
            const double2 v = flds128(((K & 3) == 0 ? own : ownO) + (uint32_t)(K >> 2) * (uint32_t)MT);
            in[K - IN_LO] = v.x; in[K - IN_LO + 1] = v.y;
        });
        uint32_t sE[HS], sO[HS];
        double amp[HS];
#pragma unroll
        for (int s = 0; s < HS; s++) {
            int rE, rO;
            F.slot<1>(s, s < nslot, A.amps, rE, rO, amp[s]);
            sE[s] = img4 + (uint32_t)rE; sO[s] = img4 + (uint32_t)rO;
        }
        const int nimp = __popc(impbits) + PART;
        const double dg = dgbase + eo + xt + A.cst.pair_e * (double)(nimp * (nimp - 1) / 2);
        const double sig = neg ? -1.0 : 1.0;
        // second stage of pair F: add the pass-1 result, <x, y>, store; its window slot goes to pair F + W
        auto finish = [&](auto ff, const double2 *xop) {
            constexpr int F2 = decltype(ff)::value, K = 2 * (P_LO + F2);
            constexpr bool do0 = K >= OUT_LO && K < OUT_HI, do1 = K + 1 >= OUT_LO && K + 1 < OUT_HI;
            const double2 xo = xop ? *xop : flds128(((K & 3) == 0 ? own : ownO) + (uint32_t)(K >> 2) * (uint32_t)MT);
            const double2 yo = yw[F2 % W];
            if constexpr (F2 + W < NPR) yw[F2 % W] = fldg128(yaddr(std::integral_constant<int, 2 * (P_LO + F2 + W)>{}));
            double r0 = 0.0, r1 = 0.0;
            if constexpr (do0) { r0 = yo.x + pend[D > 0 ? F2 % (D > 0 ? D : 1) : 0].x; dsum = fma(xo.x, r0, dsum); }
            if constexpr (do1) { r1 = yo.y + pend[D > 0 ? F2 % (D > 0 ? D : 1) : 0].y; dsum = fma(xo.y, r1, dsum); }
            // a pair that straddles the imp=0 / imp=1 boundary (A0 odd): each phase stores its own element
            double *yp = yaddr(std::integral_constant<int, K>{});
            if constexpr (EDGPU_FIB_NOYST) { if (r0 == 1.2345e300) yp[0] = r1; }
            else if constexpr (do0 && do1) { if (EDGPU_FIB_EVICT) fstg128_hint(yp, r0, r1, spol); else fstg128(yp, r0, r1); }
            else if constexpr (do0 && K + 1 >= D0) fstg128(yp, r0, 0.0);        // last pair of an odd fiber: the pad stays zero
            else if constexpr (do0) yp[0] = r0;
            else if constexpr (do1) yp[1] = r1;
        };
        fib::static_for<NPR>([&](auto jj) {
            constexpr int J = decltype(jj)::value, K = 2 * (P_LO + J);
            constexpr bool do0 = K >= OUT_LO && K < OUT_HI, do1 = K + 1 >= OUT_LO && K + 1 < OUT_HI;
            if constexpr (D > 0 && J >= D) finish(std::integral_constant<int, J - D>{}, nullptr);
            const uint32_t rel = (uint32_t)(K >> 2) * (uint32_t)MT;
            const double2 xo = flds128(((K & 3) == 0 ? own : ownO) + rel);
            double g0 = 0.0, g1 = 0.0;
#pragma unroll
            for (int s = 0; s < HS; s++) {
                const double2 v = flds128(((K & 3) == 0 ? sE[s] : sO[s]) + rel);
                if (do0) g0 = fma(amp[s], v.x, g0);
                if (do1) g1 = fma(amp[s], v.y, g1);
            }
            double c0 = 0.0, c1 = 0.0;
            if constexpr (do0) {
                const double in0 = fib::out<NB, M0, K, IN_LO>(in, A.cst.v0);
                c0 = fma(dg + A.cst.e0[COFF + K], xo.x, fma(sig, in0, PART == 0 ? g0 : -g0));
            }
            if constexpr (do1) {
                const double in1 = fib::out<NB, M0, K + 1, IN_LO>(in, A.cst.v0);
                c1 = fma(dg + A.cst.e0[COFF + K + 1], xo.y, fma(sig, in1, PART == 0 ? g1 : -g1));
            }
            pend[D > 0 ? J % (D > 0 ? D : 1) : 0] = make_double2(c0, c1);
            if constexpr (D == 0) finish(jj, &xo);
        });
        fib::static_for<D>([&](auto tt) { finish(std::integral_constant<int, NPR - D + decltype(tt)::value>{}, nullptr); });
    }
}

// Tile descriptors travel from the producer thread to the consumers through a 4-entry ring in shared memory (entry i+1 is
// written before the arrive on full(i), read after full(i) was observed).  Together with volatile re-reads of the few fields a
// work unit needs, this keeps the state that is live across the (huge, fully unrolled) fiber bodies small: the first build kept
// the descriptor in registers, which ptxas spilled -- ~30 local-memory reloads per unit, each an L2 round trip next to 226 KB
// of shared memory (26 % of the stall samples sat at the head of a unit, another 13 % at its tail and at the tile head).
struct UpRing {                 // 64 bytes
    int off_lo, off_hi, bytes, blk;
    int a, b, q0, q1;
    int q2, q3, m0, d0p;        // m0 | half << 8
    int nouter, C4, tab, nwf;   // nwf | A0 << 8 | D0 << 16
};
#define UPR(field) ((uint32_t)offsetof(UpRing, field))
__device__ __forceinline__ int ring_ld(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// HALF: the tiles are half bands (rows 0-1 or 2-3 of a band): lanes = 2 rows x 16 outer indices, image in 64-byte half micro-tiles
// fetched through a tensor map (64-byte runs); a quarter-warp = 2 rows x 4 consecutive outer indices is conflict free for
// d0p == 2 (mod 4) as well.
template <int NL, bool HALF>
__global__ void __launch_bounds__(FibCfg<NL>::NT) k_fib_up(const __grid_constant__ FibArgs A)
{
    constexpr int NC = FibCfg<NL>::NC, NW = NC / 32, MT = HALF ? 64 : 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_bar[4];
    __shared__ double s_dot[NW];
    __shared__ __align__(16) UpRing s_ring[4];
    const int tid = threadIdx.x;
    const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t stab0 = slot0 + 2u * kSlot;
    const uint32_t bfull = (uint32_t)__cvta_generic_to_shared(s_bar), bempty = bfull + 16u;
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(s_ring);
    const int myn = (int)blockIdx.x < A.ntiles ? (A.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto ring_put = [&](int i, const FibTile &t, const FibBlockDev &BU) {
        const uint32_t d = ring0 + (uint32_t)(i & 3) * (uint32_t)sizeof(UpRing);
        const int nwf = HALF ? (BU.nouter + 15) >> 4 : (4 * ((BU.nouter + 1) & ~1) + 31) >> 5;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"((int)(uint32_t)(t.off & 0xffffffffll)), "r"((int)(t.off >> 32)), "r"(t.bytes), "r"(t.blk) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + 16u), "r"(t.a), "r"(t.b), "r"(t.q0), "r"(t.q1) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + 32u), "r"(t.q2), "r"(t.q3), "r"(BU.m0 | (t.half << 8)), "r"(BU.d0p) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + 48u), "r"(BU.nouter), "r"(BU.C4), "r"(BU.tab), "r"(nwf | (BU.A0 << 8) | (BU.D0 << 16)) : "memory");
    };
    if (tid == 0) {
        fmbar_init(bfull, 1); fmbar_init(bfull + 8, 1);
        fmbar_init(bempty, NW); fmbar_init(bempty + 8, NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == NC && myn > 0) {
        const FibTile t0 = A.tiles[blockIdx.x];
        ring_put(0, t0, A.blk_f[t0.blk]);
    }
    __syncthreads();
    if (tid >= NC) {
        if (tid == NC) {
            // ne[s]: tiles that have occupied the memory of slot s so far (= phases its empty barrier must have completed);
            // a two-slot tile occupies both
            int ne[2] = {0, 0}, nfb[2] = {0, 0}, pos = 0;        // nfb[s]: completed uses of full barrier s
            FibTile t = myn > 0 ? A.tiles[blockIdx.x] : FibTile{};
            FibBlockDev bt = A.blk_f[t.blk];
            const uint64_t pol = fpolicy_evict_first(), plast = fpolicy_evict_last();
            for (int i = 0; i < myn; i++) {
                FibTile tn = t;
                if (i + 1 < myn) tn = A.tiles[blockIdx.x + (size_t)(i + 1) * gridDim.x];
                const bool two = t.bytes > A.slot;
                const int s = two ? 0 : pos;
                if (ne[s] > 0) fmbar_wait_backoff(bempty + 8 * s, (uint32_t)(ne[s] - 1) & 1u);
                if (two && ne[1] > 0) fmbar_wait_backoff(bempty + 8, (uint32_t)(ne[1] - 1) & 1u);
                fmbar_expect_tx_only(bfull + 8 * s, (uint32_t)t.bytes);
                const uint32_t dst = slot0 + (uint32_t)s * (uint32_t)A.slot;
                // the read-modify-write operand of THIS tile (the same band of y) goes to L2 while the x image lands
                if (HALF) {
                    const int c0 = 8 * (t.half - 1);
                    for (int k = 0; k < bt.hnbox; k++)
                        ftma_load_3d(dst + (uint32_t)k * (uint32_t)bt.hbox * 64u, A.tmaps + t.pair, c0, k * bt.hbox, t.a, bfull + 8 * s);
                    for (int k = 0; k < bt.hnbox; k++) ftma_prefetch_3d(A.tmaps_y + t.pair, c0, k * bt.hbox, t.a);
                } else {
                    const char *src = reinterpret_cast<const char *>(A.x + t.off);
                    for (int ofs = 0; ofs < t.bytes; ofs += 32768) {
                        const int n = t.bytes - ofs < 32768 ? t.bytes - ofs : 32768;
                        if (EDGPU_FIB_EVICT) fbulk_g2s_hint(dst + (uint32_t)ofs, src + ofs, (uint32_t)n, bfull + 8 * s, pol);
                        else fbulk_g2s(dst + (uint32_t)ofs, src + ofs, (uint32_t)n, bfull + 8 * s);
                    }
                    const char *py = reinterpret_cast<const char *>(A.y + t.off);
                    for (int ofs = 0; ofs < t.bytes; ofs += 32768)
                        { if (EDGPU_FIB_PFLAST) fbulk_prefetch_l2_hint(py + ofs, (uint32_t)(t.bytes - ofs < 32768 ? t.bytes - ofs : 32768), plast); else fbulk_prefetch_l2(py + ofs, (uint32_t)(t.bytes - ofs < 32768 ? t.bytes - ofs : 32768)); }
                }
                // every consumer has finished tile i-2 (or later) here: ring entry (i+1) & 3 = (i-3) & 3 is free.  The entry is
                // written while the copies fly and published by the (release) arrive that lets full(i) complete.
                FibBlockDev bn = bt;
                if (i + 1 < myn) {
                    if (tn.blk != t.blk) bn = A.blk_f[tn.blk];
                    ring_put(i + 1, tn, bn);
                }
                fmbar_arrive(bfull + 8 * s);
                // a two-slot tile cannot be loaded while its predecessor is computed on: at least pull it into L2 meanwhile
                // (once the own image has landed, so that the two do not compete), the exposed fill then runs at L2 speed
                if (!HALF && i + 1 < myn && tn.bytes > A.slot && !(A.dbg & 2)) {
                    fmbar_wait_backoff(bfull + 8 * s, (uint32_t)nfb[s] & 1u);
                    const char *px = reinterpret_cast<const char *>(A.x + tn.off), *py = reinterpret_cast<const char *>(A.y + tn.off);
                    for (int ofs = 0; ofs < tn.bytes; ofs += 32768)
                        { if (EDGPU_FIB_PFLAST) fbulk_prefetch_l2_hint(px + ofs, (uint32_t)(tn.bytes - ofs < 32768 ? tn.bytes - ofs : 32768), plast); else fbulk_prefetch_l2(px + ofs, (uint32_t)(tn.bytes - ofs < 32768 ? tn.bytes - ofs : 32768)); }
                    if (A.dbg & 8)
                        for (int ofs = 0; ofs < tn.bytes; ofs += 32768)
                            { if (EDGPU_FIB_PFLAST) fbulk_prefetch_l2_hint(py + ofs, (uint32_t)(tn.bytes - ofs < 32768 ? tn.bytes - ofs : 32768), plast); else fbulk_prefetch_l2(py + ofs, (uint32_t)(tn.bytes - ofs < 32768 ? tn.bytes - ofs : 32768)); }
                    // (pulling the y band of the next tile in as well was measured SLOWER: 1.69 against 1.61 ms -- four
                    // 157 KB images per SM, 93 MB in all, no longer fit the L2 next to the write-back traffic)
                }
                ne[s]++; nfb[s]++;
                if (two) ne[1]++;
                else pos ^= 1;
                t = tn; bt = bn;
            }
        }
        return;
    }
    const int warp = tid >> 5, lane = tid & 31, rl = HALF ? (lane & 1) : (lane & 3);       // row of the image
    const uint64_t spol = fpolicy_evict_first();
    double dsum = 0.0;
    int nfill0 = 0, nfill1 = 0, pos = 0, cur_blk = -1;
    bool stab = false;
    for (int i = 0; i < myn; i++) {
        const uint32_t re = ring0 + (uint32_t)(i & 3) * (uint32_t)sizeof(UpRing);
        const bool two = ring_ld(re + UPR(bytes)) > A.slot;
        const int s = two ? 0 : pos;
        {
            const int blk = ring_ld(re + UPR(blk));
            if (blk != cur_blk) {
                const FibBlockDev BU = A.blk_f[blk];
                stab = load_stab<NC, 1>(A, BU, stab0, tid, MT);
                cur_blk = blk;
            }
        }
        const int m0h = ring_ld(re + UPR(m0));
        const int r4 = HALF ? 2 * ((m0h >> 8) - 1) + rl : rl;                                // row of the band
        // work units of a tile: [band g][phase][32 fibers]; lanes = 4 rows x 8 outer indices, a quarter-warp = 4 rows x 2
        // neighbouring outer indices (conflict-free LDS.128).  A thread keeps its row r4 = lane & 3 in every unit, so the
        // per-row terms change with the band only: they are fetched one unit ahead (for band 0 before the wait for the image).
        auto rowterms = [&](int g, double &dg, uint32_t &impd) -> bool {
            const int q0 = ring_ld(re + UPR(q0)), q2 = ring_ld(re + UPR(q2));
            const int rp = (ring_ld(re + UPR(a)) + g) * 4 + r4;
            const int od = rp / q0, kd = rp - od * q0;
            if (od >= ring_ld(re + UPR(q1)) || kd >= q2) return false;
            const int id = ring_ld(re + UPR(q3)) + od * q2 + kd;
            dg = __ldg(A.e_dw + id);
            impd = __ldg(A.cfg_dw + id) & A.impmask;
            return true;
        };
        const int nb = (A.dbg & 1) ? 0 : ring_ld(re + UPR(b));
        const int nwfp = ring_ld(re + UPR(nwf));
        const int nwf2 = 2 * (nwfp & 255);
        int g = 0, rem = warp;
        while (rem >= nwf2) { rem -= nwf2; g++; }
        double dgb = 0.0;
        uint32_t impd = 0;
        bool rowok = g < nb ? rowterms(g, dgb, impd) : false;
        // L2 prefetch of the y sectors of ONE unit of this lane (row r4, fiber o, phase): issued a unit ahead -- for the first
        // unit of a tile before the wait for the image -- so that the read-modify-write loads of the unit find their operand
        // in L2 (the band prefetch of the producer is only a hint and does not keep up on the two-slot tiles)
        auto ypf = [&](int gq, int remq, bool okq) {
            if (!EDGPU_FIB_YPF || !okq || gq >= nb) return;
            const int nwfq = nwf2 >> 1;
            const int partq = remq >= nwfq ? 1 : 0;
            const int fbq = (remq - partq * nwfq) * 32 + lane;
            const int oq = HALF ? (fbq >> 1) : 2 * (fbq >> 3) + ((fbq >> 2) & 1);
            if (oq >= ring_ld(re + UPR(nouter))) return;
            const int A0q = (nwfp >> 8) & 255, D0q = (nwfp >> 16) & 255;
            const int c0q = oq * ring_ld(re + UPR(d0p));
            const int lo = c0q + (partq ? A0q : 0), hi = c0q + (partq ? D0q : A0q) - 1;
            const int C4q = ring_ld(re + UPR(C4));
            const int64_t toffq = (int64_t)(((uint64_t)(uint32_t)ring_ld(re + UPR(off_hi)) << 32) | (uint64_t)(uint32_t)ring_ld(re + UPR(off_lo)));
            const double *yq = A.y + toffq + (int64_t)gq * C4q * 16 + r4 * 4;
            for (int sq = lo >> 2; sq <= (hi >> 2); sq++) asm volatile("prefetch.global.L2 [%0];" ::"l"(yq + sq * 16));
        };
        ypf(g, rem, rowok);
        fmbar_wait_warp(bfull + 8 * s, (uint32_t)(s ? nfill1 : nfill0) & 1u);
        const uint32_t img0 = slot0 + (uint32_t)s * (uint32_t)A.slot + (uint32_t)rl * 32u;
        while (g < nb) {
            // the next unit of this warp; its row terms are requested now if it lies in another band
            int gn = g, remn = rem + NW;
            while (remn >= nwf2) { remn -= nwf2; gn++; }
            double dgn = dgb;
            uint32_t impn = impd;
            bool okn = rowok;
            if (gn != g && gn < nb) okn = rowterms(gn, dgn, impn);
            ypf(gn, remn, okn);
            const int nwf = nwf2 >> 1;
            const int part = rem >= nwf ? 1 : 0;
            const int fb = (rem - part * nwf) * 32 + lane;
            const int o = HALF ? (fb >> 1) : 2 * (fb >> 3) + ((fb >> 2) & 1);
            const bool active = rowok && o < ring_ld(re + UPR(nouter));
            FiberMeta F;
            F.stab = stab ? stab0 + (uint32_t)(active ? o : 0) * (uint32_t)kSOuterBytes : 0u;
            F.ent = A.outer + ring_ld(re + UPR(tab)) + (active ? o : 0);
            F.o = o; F.stride = ring_ld(re + UPR(d0p)); F.mt = MT;
            double eo; int impbits, nslot, neg;
            F.head(eo, impbits, nslot, neg);
            if (!active) nslot = 0;
            const int wmax = __reduce_max_sync(0xffffffffu, nslot);
            if (active) {
                const int ib = impbits | part;
                const double xt = stab ? flds64(stab0 + (uint32_t)kStabXtab + 8u * (impd * (1u << A.norb) + (uint32_t)ib)) : __ldg(A.xtab + impd * 32u + (uint32_t)ib);
                const int C4 = ring_ld(re + UPR(C4));
                const uint32_t img4 = img0 + (uint32_t)g * (uint32_t)C4 * (uint32_t)MT;
                const int64_t toff = (int64_t)(((uint64_t)(uint32_t)ring_ld(re + UPR(off_hi)) << 32) | (uint64_t)(uint32_t)ring_ld(re + UPR(off_lo)));
                double *yband4 = A.y + toff + (int64_t)g * C4 * 16 + r4 * 4;
                const int m0 = m0h & 255;
                fib::static_for<NL - 1>([&](auto mm) {
                    constexpr int M0 = decltype(mm)::value + 1;
                    if (m0 == M0) {
                        fib::static_for<FibHS<NL>::n>([&](auto hh) {
                            constexpr int H = fib_hs<NL>(decltype(hh)::value), HP = decltype(hh)::value == 0 ? 0 : fib_hs<NL>(decltype(hh)::value - 1);
                            if (wmax <= H && (decltype(hh)::value == 0 || wmax > HP)) {
                                if (part == 0) fiber_up<NL, M0, H, 0, MT>(A, F, nslot, eo, impbits, neg, img4, o, F.stride, dgb, xt, yband4, dsum, spol);
                                else fiber_up<NL, M0, H, 1, MT>(A, F, nslot, eo, impbits, neg, img4, o, F.stride, dgb, xt, yband4, dsum, spol);
                            }
                        });
                    }
                });
            }
            g = gn; rem = remn; dgb = dgn; impd = impn; rowok = okn;
        }
        __syncwarp();
        if (lane == 0) { fmbar_arrive(bempty + 8 * s); if (two) fmbar_arrive(bempty + 8); }
        if (s) nfill1++; else nfill0++;
        if (!two) pos ^= 1;
    }
    if (A.dot_out) {
        for (int o = 16; o > 0; o >>= 1) dsum += __shfl_down_sync(0xffffffffu, dsum, o);
        if (lane == 0) s_dot[warp] = dsum;
        asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
        if (tid == 0) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NW; w++) v += s_dot[w];
            A.dot_out[blockIdx.x] = v;
        }
    }
}

// ---- down pass (runs FIRST: y = H_dw x, write-only): one column (c4 of the strip) of the fiber (outer index o) of the down
// spin; PART 0: outputs with imp=0 ----
// RS: bytes per row of the image (32: strips of 4 columns; 16: half strips of 2 columns)
template <int NL, int M0, int HS, int PART, int RS>
__device__ __forceinline__ void fiber_dw(const FibArgs &A, const FiberMeta &F, int nslot, int neg, uint32_t img8 /* image + c4*8 */, int o,
                                         int d0r, double *ystrip4 /* strip + c4 */, int64_t bstride, uint64_t spol)
{
    constexpr int NB = NL - 1, D0 = fib::cbinom(NL, M0), A0 = fib::cbinom(NB, M0), B0 = D0 - A0;
    constexpr int NIN = PART == 0 ? B0 : A0, NOUT = PART == 0 ? A0 : B0, IN0 = PART == 0 ? A0 : 0, OUT0 = PART == 0 ? 0 : A0;
    if constexpr (NIN > 0 && NOUT > 0) {
        const int r0 = o * d0r;
        // global rows of the outputs: row r0 + k lies in band (r0+k)/4 at sub-row (r0+k)%4; Q[j] serves k == j (mod 4)
        double *Q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) Q[j] = ystrip4 + (int64_t)((r0 + j) >> 2) * bstride + ((r0 + j) & 3) * 4;
        const uint32_t base = img8 + (uint32_t)r0 * (uint32_t)RS;
        double in[NIN];
        fib::static_for<NIN>([&](auto jj) { constexpr int J = decltype(jj)::value; in[J] = flds64(base + (uint32_t)(IN0 + J) * (uint32_t)RS); });
        uint32_t sb[HS];
        double amp[HS];
#pragma unroll
        for (int s = 0; s < HS; s++) {
            int rE, rO;
            F.slot<2>(s, s < nslot, A.amps, rE, rO, amp[s]);
            sb[s] = img8 + (uint32_t)rE;
        }
        const double sig = neg ? -1.0 : 1.0;
        fib::static_for<NOUT>([&](auto kk) {
            constexpr int K = OUT0 + decltype(kk)::value;
            double g = 0.0;
#pragma unroll
            for (int s = 0; s < HS; s++) g = fma(amp[s], flds64(sb[s] + (uint32_t)K * (uint32_t)RS), g);
            const double inr = fib::out<NB, M0, K, IN0>(in, A.cst.v0);
            const double val = fma(sig, inr, PART == 0 ? g : -g);
            if (EDGPU_FIB_EVICT_DW) asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(Q[K & 3] + (int64_t)(K >> 2) * bstride), "d"(val), "l"(spol) : "memory");
            else Q[K & 3][(int64_t)(K >> 2) * bstride] = val;
        });
    }
}

// HALF: the tiles are half strips (columns 0-1 or 2-3 of a strip): lanes = 2 columns x 16 outer indices, image rows of 16 bytes
// fetched through a 4-D tensor map (box 2 x 4 x 1 x BR: 16-byte runs)
template <int NL, bool HALF>
__global__ void __launch_bounds__(FibCfgDw<NL>::NT) k_fib_dw(const __grid_constant__ FibArgs A)
{
    constexpr int NC = FibCfgDw<NL>::NC, NW = NC / 32, RS = HALF ? 16 : 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_bar[4];
    const int tid = threadIdx.x;
    const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t stab0 = slot0 + 2u * kSlot;
    const uint32_t bfull = (uint32_t)__cvta_generic_to_shared(s_bar), bempty = bfull + 16u;
    if (tid == 0) {
        fmbar_init(bfull, 1); fmbar_init(bfull + 8, 1);
        fmbar_init(bempty, NW); fmbar_init(bempty + 8, NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int myn = (int)blockIdx.x < A.ntiles ? (A.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid >= NC) {
        if (tid == NC) {
            int ne[2] = {0, 0}, nfb[2] = {0, 0}, pos = 0;        // nfb[s]: completed uses of full barrier s
            const uint64_t pol = fpolicy_evict_first();
            auto boxes = [&](const FibTile &t, const FibBlockDev &BD, auto &&fn) {
                for (int g = 0; g < t.b; g++)
                    for (int b = 0; b < BD.nbox; b++) fn(g, b);
            };
            FibTile t = myn > 0 ? A.tiles[blockIdx.x] : FibTile{};
            for (int i = 0; i < myn; i++) {
                FibTile tn = t;
                if (i + 1 < myn) tn = A.tiles[blockIdx.x + (size_t)(i + 1) * gridDim.x];
                const FibBlockDev BD = A.blk_f[t.blk];
                const bool two = t.bytes > A.slot;
                const int s = two ? 0 : pos;
                if (ne[s] > 0) fmbar_wait_backoff(bempty + 8 * s, (uint32_t)(ne[s] - 1) & 1u);
                if (two && ne[1] > 0) fmbar_wait_backoff(bempty + 8, (uint32_t)(ne[1] - 1) & 1u);
                fmbar_expect_tx(bfull + 8 * s, (uint32_t)t.bytes);
                const uint32_t dst = slot0 + (uint32_t)s * (uint32_t)A.slot;
                const uint32_t sbytes = (uint32_t)BD.nbox * (uint32_t)BD.BR * (uint32_t)(RS * 4);
                boxes(t, BD, [&](int g, int b) {
                    const uint32_t d = dst + (uint32_t)g * sbytes + (uint32_t)b * (uint32_t)BD.BR * (uint32_t)(RS * 4);
                    if (HALF) ftma_load_4d(d, A.tmaps + t.pair, 2 * (t.half - 1), 0, t.a + g, b * BD.BR, bfull + 8 * s);
                    else if (EDGPU_FIB_EVICT) ftma_load_3d_hint(d, A.tmaps + t.pair, 0, t.a + g, b * BD.BR, bfull + 8 * s, pol);
                    else ftma_load_3d(d, A.tmaps + t.pair, 0, t.a + g, b * BD.BR, bfull + 8 * s);
                });
                // see the up pass: the next two-slot strip goes to L2 while this one is computed on
                if (!HALF && i + 1 < myn && tn.bytes > A.slot && !(A.dbg & 2)) {
                    fmbar_wait_backoff(bfull + 8 * s, (uint32_t)nfb[s] & 1u);
                    const FibBlockDev BN = tn.blk == t.blk ? BD : A.blk_f[tn.blk];
                    for (int g = 0; g < tn.b; g++)
                        for (int b = 0; b < BN.nbox; b++) ftma_prefetch_3d(A.tmaps + tn.pair, 0, tn.a + g, b * BN.BR);
                }
                ne[s]++; nfb[s]++;
                if (two) ne[1]++;
                else pos ^= 1;
                t = tn;
            }
        }
        return;
    }
    int nfill0 = 0, nfill1 = 0, pos = 0, cur_blk = -1;
    bool stab = false;
    const int warp = tid >> 5, lane = tid & 31, cl = HALF ? (lane & 1) : (lane & 3);         // column of the image
    const uint64_t spol = fpolicy_evict_first();
    int b_m0 = 0, b_d0r = 0, b_nouter = 0, b_tab = 0, now = 1;
    uint32_t sbytes = 0;
    FibTile tnext = myn > 0 ? A.tiles[blockIdx.x] : FibTile{};
    for (int i = 0; i < myn; i++) {
        const FibTile t = tnext;
        if (i + 1 < myn) tnext = A.tiles[blockIdx.x + (size_t)(i + 1) * gridDim.x];      // in flight during this tile
        const bool two = t.bytes > A.slot;
        const int s = two ? 0 : pos;
        if (t.blk != cur_blk) {
            const FibBlockDev BD = A.blk_f[t.blk];
            stab = load_stab<NC, 2>(A, BD, stab0, tid, RS);
            cur_blk = t.blk;
            b_m0 = BD.m0; b_d0r = BD.d0r; b_nouter = BD.nouter; b_tab = BD.tab;
            now = HALF ? (BD.nouter + 15) >> 4 : (BD.nouter + 7) >> 3;
            sbytes = (uint32_t)BD.nbox * (uint32_t)BD.BR * (uint32_t)(RS * 4);
        }
        fmbar_wait_warp(bfull + 8 * s, (uint32_t)(s ? nfill1 : nfill0) & 1u);
        const int64_t bstride = (int64_t)t.q0 * 16;
        // warp-fibers of a strip: [part][8 outer indices per warp]; lanes = (c4 = lane & 3, o = 8*ow + lane/4)
        // (half strips: 16 outer indices per warp, lanes = (lane & 1, o = 16*ow + lane/2))
        const int nwf = 2 * now * t.b;
        for (int wf = warp; wf < ((A.dbg & 1) ? 0 : nwf); wf += NW) {
            int g = 0, rem = wf;
            if (t.b > 1) { g = wf / (2 * now); rem = wf - g * 2 * now; }
            const int part = rem >= now ? 1 : 0, ow = rem - part * now;
            const int o = HALF ? 16 * ow + (lane >> 1) : 8 * ow + (lane >> 2);
            const bool active = o < b_nouter;
            FiberMeta F;
            F.stab = stab ? stab0 + (uint32_t)(active ? o : 0) * (uint32_t)kSOuterBytes : 0u;
            F.ent = A.outer + b_tab + (active ? o : 0);
            F.o = o; F.stride = b_d0r; F.mt = RS;
            double eo; int impbits, nslot, neg;
            F.head(eo, impbits, nslot, neg);
            if (!active) nslot = 0;
            const int wmax = __reduce_max_sync(0xffffffffu, nslot);
            if (active) {
                const uint32_t img8 = slot0 + (uint32_t)s * (uint32_t)A.slot + (uint32_t)g * sbytes + (uint32_t)cl * 8u;
                double *ystrip4 = A.y + t.off + (int64_t)(t.a + g) * 16 + (HALF ? 2 * (t.half - 1) : 0) + cl;
                fib::static_for<NL - 1>([&](auto mm) {
                    constexpr int M0 = decltype(mm)::value + 1;
                    if (b_m0 == M0) {
                        fib::static_for<FibHS<NL>::n>([&](auto hh) {
                            constexpr int H = fib_hs<NL>(decltype(hh)::value), HP = decltype(hh)::value == 0 ? 0 : fib_hs<NL>(decltype(hh)::value - 1);
                            if (wmax <= H && (decltype(hh)::value == 0 || wmax > HP)) {
                                if (part == 0) fiber_dw<NL, M0, H, 0, RS>(A, F, nslot, neg, img8, o, b_d0r, ystrip4, bstride, spol);
                                else fiber_dw<NL, M0, H, 1, RS>(A, F, nslot, neg, img8, o, b_d0r, ystrip4, bstride, spol);
                            }
                        });
                    }
                });
            }
        }
        __syncwarp();
        if (lane == 0) { fmbar_arrive(bempty + 8 * s); if (two) fmbar_arrive(bempty + 8); }
        if (s) nfill1++; else nfill0++;
        if (!two) pos ^= 1;
    }
}

// host entry of one translation unit: pass 1 or 2 of the fiber kernels for NL levels per star, full or half tiles
template <int NL, bool HALF>
static int fib_launch(int pass, cudaStream_t st, const FibArgs &A, int grid)
{
    static bool attr[2] = {false, false};
    const size_t smem = 2 * (size_t)kSlot + kStab;
    if (!attr[pass - 1]) {
        cudaError_t e = pass == 1 ? cudaFuncSetAttribute((const void *)k_fib_up<NL, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                  : cudaFuncSetAttribute((const void *)k_fib_dw<NL, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr[pass - 1] = true;
    }
    if (pass == 1) k_fib_up<NL, HALF><<<grid, FibCfg<NL>::NT, smem, st>>>(A);
    else k_fib_dw<NL, HALF><<<grid, FibCfgDw<NL>::NT, smem, st>>>(A);
    return (int)cudaGetLastError();
}
