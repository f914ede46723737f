// fiber_common.h -- structures shared by the host side (hxv_fiber.cu) and the per-NL kernel translation units (fib_nl*.cu)
// of the fiber H*v engine.  See hxv_fiber.cu for the design.
#pragma once
#include "star_info.h"
#include <cuda.h>
#include <map>
#include <utility>

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
static constexpr int kFibMinBlock = 64;             // blocks smaller than this use the thread-per-element pair kernels

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
    int hbox, hnbox;            // up pass, half bands (2 rows): micro-tile columns per tensor box, boxes per half band
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

// Shared-memory form of one outer-table entry, everything resolved for the pass that uses it: 8 x 16 bytes.
//   [0]      eo (double), impbits (int), nslot | neg << 8 (int)
//   [1 + s]  relE, relO (int), amp (double): byte offsets of the neighbour fiber in the image (pass 1: for k == 0 / 2 mod 4
//            at row 0 of the band; pass 2: relE = byte offset of row o'*d0r, relO unused) and the signed amplitude
static constexpr int kSOuterBytes = 128;
static constexpr int kStabXtab = 8960;              // byte offset of the xtab subset ((1 << norb)^2 doubles) inside the stab area

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
struct FibTile {                // pass 1: bands [a, a+b) of the pair ; pass 2: strips [a, a+b); 48 bytes, self-contained
    int64_t off;                // pass 1: first element of band a in the vector ; pass 2: pair base
    int pair, blk, a, b;
    int bytes;
    int q0, q1, q2, q3;         // pass 1: d0r, nouter, D0, off of the DOWN block (rows of the pair) ; pass 2: q0 = C4 of the up block
    int half;                   // 0: full band / strip (4 rows / columns); 1, 2: its first / second half (half-tile kernels)
};
static_assert(sizeof(FibTile) == 48, "FibTile must be 48 bytes");

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
    // half tiles (2 rows of a band / 2 columns of a strip): blocks whose full image does not fit ONE pipeline slot -- halved, the
    // image double-buffers (4900 configurations of Ns=16) or fits the two slots at all (8000 configurations of Ns=18)
    FibTile *d_t1h = nullptr, *d_t2h = nullptr;
    int n1h = 0, n2h = 0;
    int *d_g1 = nullptr, *d_g2 = nullptr;      // pair ids for the generic up / down kernels
    // tables of the memory-order pair kernels: inverse position maps (padded position -> internal index, -1 = pad) per
    // block, concatenated (start of block b at *_start[b]), and hop tables whose targets are POSITIONS inside the block
    int *d_idx_of_cp = nullptr, *d_idx_of_rp = nullptr, *d_cp_start = nullptr, *d_rp_start = nullptr;
    uint32_t *d_poshop_c = nullptr, *d_poshop_r = nullptr;
    int ng1 = 0, ng2 = 0;
    int64_t g1_elems = 0, g2_elems = 0;
    int nl = 0;
    int slot = kSlot;                                      // pipeline slot size the tile schedules were built for
    int skip1 = 0, skip2 = 0;                              // test hooks: leave a pass to the thread-per-element kernels
    // per source pointer, a device array of 3 maps per pair: [p] strips (3-D, box 16 x 1 x BR), [np + p] half strips (4-D,
    // box 2 x 4 x 1 x BR), [2 np + p] half bands (3-D, box 8 x hbox x 1)
    std::map<const double *, CUtensorMap *> tmaps;
    ~PairLayout()
    {
        cudaFree(d_rowinfo); cudaFree(d_colinfo); cudaFree(d_pbase); cudaFree(d_c4); cudaFree(d_pairs); cudaFree(d_t1); cudaFree(d_t2); cudaFree(d_t1h); cudaFree(d_t2h);
        cudaFree(d_g1); cudaFree(d_g2);
        cudaFree(d_idx_of_cp); cudaFree(d_idx_of_rp); cudaFree(d_cp_start); cudaFree(d_rp_start); cudaFree(d_poshop_c); cudaFree(d_poshop_r);
        for (auto &kv : tmaps) cudaFree(kv.second);
    }
};

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
    int norb;
    int slot;                           // bytes per pipeline slot (kSlot; smaller in the two-slot test mode)
    int dbg;                            // measurement hooks: bit 0 = consumers skip the fibers (load pipeline only)
    const double *x;
    double *y;
    const double *e_dw;                 // pass 1: per-row diagonal energy and configuration word of the down spin
    const uint32_t *cfg_dw;
    const double *xtab;
    const CUtensorMap *tmaps;           // one tensor map per pair over x, of the kind the kernel loads with (see PairLayout::tmaps)
    const CUtensorMap *tmaps_y;         // half-band up pass: the same over y (L2 prefetch of the read-modify-write operand)
    double *dot_out;                    // pass 2: per-CTA partial <x, y> (nullptr: not wanted)
};
