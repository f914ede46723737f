// tables.cu -- device construction of the per-spin sector tables.
//
// Replaces build_sector (ED_SETUP.f90:886-916), bdecomp (:1234-1244), binomial (:1283-1300) and the
// binary_search lookups (:1307-1324) of the reference.  The reference scans all 2^Ns x 2^Ns words per call;
// here each spin species is enumerated once: thread r unranks the r-th popcount-n word (colex order =
// ascending numeric order = the reference's loop order), and a 2^Ns rank LUT replaces binary_search.
// map(i) = iup + idw*2^Ns for i = r_up + r_dw*DimUp is then a pure function of the two per-spin lists.
#include "edgpu_internal.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>

__constant__ uint64_t c_binom[33][33];
static bool g_binom_uploaded[64] = {false};

static uint64_t h_binom[33][33];
static void init_binom_host()
{
    static bool done = false;
    if (done) return;
    for (int n = 0; n <= 32; n++)
        for (int k = 0; k <= 32; k++)
            h_binom[n][k] = (k == 0) ? 1 : (n == 0 ? 0 : h_binom[n - 1][k - 1] + h_binom[n - 1][k]);
    done = true;
}

uint64_t edgpu_binom(int n, int k)
{
    init_binom_host();
    if (k < 0 || n < 0 || k > 32 || n > 32) return 0;
    return h_binom[n][k];
}

static int upload_binom(edgpu_ctx *ctx)
{
    init_binom_host();
    if (ctx->device >= 0 && ctx->device < 64 && g_binom_uploaded[ctx->device]) return 0;
    CUDA_TRY(ctx, cudaMemcpyToSymbol(c_binom, h_binom, sizeof(h_binom)));
    if (ctx->device >= 0 && ctx->device < 64) g_binom_uploaded[ctx->device] = true;
    return 0;
}

// r-th (0-based) Ns-bit word with n set bits in ascending numeric (colex) order.
__device__ __forceinline__ uint32_t unrank_colex(int ns, int n, uint64_t r)
{
    uint32_t w = 0;
    int p = ns - 1;
    for (int i = n; i >= 1; i--) {
        while (c_binom[p][i] > r) p--;          // largest p with C(p,i) <= r
        w |= 1u << p;
        r -= c_binom[p][i];
        p--;
    }
    return w;
}

__device__ __forceinline__ uint32_t rank_colex(uint32_t w)
{
    uint64_t r = 0;
    int i = 1;
    while (w) {
        int p = __ffs(w) - 1;
        r += c_binom[p][i];
        i++;
        w &= w - 1;
    }
    return (uint32_t)r;
}

__global__ void k_unrank(int ns, int n, int64_t dim, uint32_t *__restrict__ cfg)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < dim) cfg[r] = unrank_colex(ns, n, (uint64_t)r);
}

__global__ void k_rank_lut_ref(int ns, int n, uint32_t *__restrict__ rank)
{
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (1u << ns)) return;
    rank[w] = (__popc(w) == n) ? rank_colex(w) : 0xFFFFFFFFu;
}

// Per-spin diagonal energy E_sigma(w) = sum_bits elev[bit]*n_bit + (Ust-Jh) * sum_{a<b} n_a n_b
// (direct/HxVimp.f90:2-8, HxVint.f90:19-23,27-38, HxVbath.f90:4-11 regrouped).
__global__ void k_ediag(int ns, int norb, int64_t dim, const uint32_t *__restrict__ cfg,
                        const double *__restrict__ elev, double upp, double *__restrict__ ediag)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= dim) return;
    uint32_t w = cfg[r];
    double e = 0.0;
    for (int b = 0; b < ns; b++)
        if ((w >> b) & 1u) e += elev[b];
    int nimp = __popc(w & ((1u << norb) - 1u));
    e += upp * (double)(nimp * (nimp - 1) / 2);
    ediag[r] = e;
}

struct HopPairDev { int p, q, amp; };

// Gather-form hop table of one spin: for output configuration w, every pair {p,q} with exactly one of the two
// bits set contributes source w' = w ^ bit(p) ^ bit(q) with sign (-1)^{popcount(w & between(p,q))}
// (c/cdg, ED_SETUP.f90:1080-1106; the sign of c^+_p c_q depends only on the bits strictly between).
__global__ void k_hops(int64_t dim, const uint32_t *__restrict__ cfg, const uint32_t *__restrict__ rank,
                       int npairs, const HopPairDev *__restrict__ pairs, int maxhop,
                       uint32_t *__restrict__ hop, uint8_t *__restrict__ nhop)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= dim) return;
    uint32_t w = cfg[r];
    int cnt = 0;
    for (int i = 0; i < npairs; i++) {
        int p = pairs[i].p, q = pairs[i].q;
        uint32_t bp = (w >> p) & 1u, bq = (w >> q) & 1u;
        if (bp ^ bq) {
            uint32_t w2 = w ^ (1u << p) ^ (1u << q);
            uint32_t between = ((1u << q) - 1u) & ~((1u << (p + 1)) - 1u);
            uint32_t neg = __popc(w & between) & 1u;
            uint32_t tgt = rank[w2];
            hop[(int64_t)cnt * dim + r] = (tgt << 8) | (uint32_t)(2 * pairs[i].amp) | neg;
            cnt++;
        }
    }
    nhop[r] = (uint8_t)cnt;
    for (int j = cnt; j < maxhop; j++) hop[(int64_t)j * dim + r] = 0;
}

__global__ void k_scatter_cfg(int ns, const uint32_t *__restrict__ rank, uint32_t *__restrict__ cfg)
{
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (1u << ns)) return;
    uint32_t r = rank[w];
    if (r != 0xFFFFFFFFu) cfg[r] = w;
}

__global__ void k_ref2int(int64_t dim, const uint32_t *__restrict__ cfg_ref, const uint32_t *__restrict__ rank,
                          uint32_t *__restrict__ ref2int)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < dim) ref2int[r] = rank[cfg_ref[r]];
}

// map(i) = iup + idw*2**Ns  (ED_SETUP.f90:914), i = r_up + r_dw*DimUp, evaluated in 64-bit.
__global__ void k_map(int ns, int64_t dim_up, int64_t first, int64_t count,
                      const uint32_t *__restrict__ cfg_up_ref, const uint32_t *__restrict__ cfg_dw_ref,
                      uint64_t *__restrict__ out)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    int64_t i = first + t;
    int64_t rd = i / dim_up, ru = i - rd * dim_up;
    out[t] = (uint64_t)cfg_up_ref[ru] + ((uint64_t)cfg_dw_ref[rd] << ns);
}

// Order-sensitive checksum + violation count (ascending order, popcounts) of the whole map, no materialisation.
__global__ void k_map_check(int ns, int nup, int ndw, int64_t dim_up, int64_t dim,
                            const uint32_t *__restrict__ cfg_up_ref, const uint32_t *__restrict__ cfg_dw_ref,
                            unsigned long long *__restrict__ sum, unsigned long long *__restrict__ viol)
{
    unsigned long long acc = 0, bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t rd = i / dim_up, ru = i - rd * dim_up;
        uint32_t u = cfg_up_ref[ru], d = cfg_dw_ref[rd];
        uint64_t m = (uint64_t)u + ((uint64_t)d << ns);
        acc += m * (2ull * (uint64_t)i + 1ull);
        if (__popc(u) != nup || __popc(d) != ndw) bad++;
        if (i > 0) {
            int64_t j = i - 1, rd2 = j / dim_up, ru2 = j - rd2 * dim_up;
            uint64_t mprev = (uint64_t)cfg_up_ref[ru2] + ((uint64_t)cfg_dw_ref[rd2] << ns);
            if (!(mprev < m)) bad++;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_down_sync(0xffffffffu, acc, o);
        bad += __shfl_down_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sum, acc);      // integer atomics: order-independent, exact
        atomicAdd(viol, bad);
    }
}

SpinBasis::~SpinBasis()
{
    if (cfg_ref && cfg_ref != cfg) cudaFree(cfg_ref);
    cudaFree(cfg); cudaFree(rank); cudaFree(ref2int); cudaFree(ediag); cudaFree(hop); cudaFree(nhop); cudaFree(amp);
}

int build_star_layout(edgpu_ctx *ctx, SpinBasis *b, const std::vector<HopPair> &pairs, const std::vector<double> &amps);

// Host-side description of the same-spin hops for parameter spin `ps` (direct/HxVimp_bath.f90:1-38 and
// direct/HxVimp.f90:16-50).  Amplitudes exactly zero are skipped like the reference's /=0 tests.
static int collect_pairs(edgpu_ctx *ctx, int ps, std::vector<HopPair> &pairs, std::vector<double> &amps)
{
    const HamParams &h = ctx->ham;
    for (int a = 0; a < h.norb; a++)
        for (int k = 0; k < h.nbath; k++) {
            double v = h.V(ps, a, k);
            if (v == 0.0) continue;
            HopPair hp;
            hp.p = a;                                   // impurity level a+1 -> bit a
            hp.q = h.norb + a * h.nbath + k;            // getBathStride(a+1,k+1)-1, ED_SETUP.f90:450-454
            hp.amp = (int)amps.size();
            hp.star = a; hp.k = k;
            amps.push_back(v);
            pairs.push_back(hp);
        }
    for (int a = 0; a < h.norb; a++)
        for (int b = a + 1; b < h.norb; b++) {
            double t = h.H(ps, a, b), t2 = h.H(ps, b, a);
            if (t == 0.0 && t2 == 0.0) continue;
            if (t != t2) return edgpu_fail(ctx, "impHloc is not symmetric (orbitals %d,%d): unsupported", a + 1, b + 1);
            HopPair hp;
            hp.p = a; hp.q = b; hp.amp = (int)amps.size(); hp.star = -1; hp.k = -1;
            amps.push_back(t);
            pairs.push_back(hp);
        }
    if (amps.size() > 127) return edgpu_fail(ctx, "too many distinct hop amplitudes (%zu > 127)", amps.size());
    return 0;
}

int build_spin_basis(edgpu_ctx *ctx, int pspin, int n, std::shared_ptr<SpinBasis> &out)
{
    const HamParams &h = ctx->ham;
    auto key = std::make_pair(pspin, n);
    auto it = ctx->bases.find(key);
    if (it != ctx->bases.end() && it->second->ham_version == h.version) { out = it->second; return 0; }
    if (int rc = upload_binom(ctx)) return rc;

    const int ns = h.ns;
    if (ns > 24) return edgpu_fail(ctx, "Ns=%d > 24 not supported by the 24-bit hop-table targets", ns);
    auto b = std::make_shared<SpinBasis>();
    b->ns = ns; b->n = n; b->pspin = pspin; b->ham_version = h.version;
    b->dim = (int64_t)edgpu_binom(ns, n);
    const int64_t dim = b->dim;
    const uint32_t nwords = 1u << ns;
    cudaStream_t st = ctx->stream;
    const int T = 256;
    const unsigned gdim = (unsigned)((dim + T - 1) / T), gw = (nwords + T - 1) / T;

    std::vector<HopPair> pairs;
    std::vector<double> amps;
    if (int rc = collect_pairs(ctx, pspin, pairs, amps)) return rc;
    bool all_star = true;
    for (auto &p : pairs) if (p.star < 0) all_star = false;

    int layout = ctx->par.layout;
    // auto: the star-product order whenever the hop structure allows it (all BASELINE configs); it serves both the
    // tiled star kernels and the generic table kernel
    if (layout == 0) layout = (all_star && h.norb <= 3 && h.nbath <= 9 && ctx->par.hxv_kernel != 1) ? 2 : 1;
    if (layout == 2 && !all_star)
        return edgpu_fail(ctx, "star-product layout requested but impHloc has inter-orbital hopping");
    b->layout = layout;

    CUDA_TRY(ctx, cudaMalloc(&b->cfg_ref, sizeof(uint32_t) * (size_t)dim));
    CUDA_TRY(ctx, cudaMalloc(&b->rank, sizeof(uint32_t) * (size_t)nwords));
    k_unrank<<<gdim, T, 0, st>>>(ns, n, dim, b->cfg_ref);
    if (layout == 1) {
        b->cfg = b->cfg_ref;
        k_rank_lut_ref<<<gw, T, 0, st>>>(ns, n, b->rank);
    } else {
        CUDA_TRY(ctx, cudaMalloc(&b->cfg, sizeof(uint32_t) * (size_t)dim));
        CUDA_TRY(ctx, cudaMalloc(&b->ref2int, sizeof(uint32_t) * (size_t)dim));
        if (int rc = build_star_layout(ctx, b.get(), pairs, amps)) return rc;   // fills b->rank (+ star info)
        k_scatter_cfg<<<gw, T, 0, st>>>(ns, b->rank, b->cfg);
        k_ref2int<<<gdim, T, 0, st>>>(dim, b->cfg_ref, b->rank, b->ref2int);
    }

    // per-level energies for this spin
    std::vector<double> elev(ns, 0.0);
    const double upp = (h.norb > 1) ? (h.ust - h.jh) : 0.0;
    for (int a = 0; a < h.norb; a++) {
        double c = h.H(pspin, a, a) - h.xmu;
        if (h.hfmode) {
            c -= 0.5 * h.uloc[a];
            if (h.norb > 1) c -= (h.norb - 1) * (0.5 * h.ust + 0.5 * (h.ust - h.jh));
        }
        elev[a] = c;
        for (int k = 0; k < h.nbath; k++) elev[h.norb + a * h.nbath + k] = h.E(pspin, a, k);
    }
    double *d_elev = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_elev, sizeof(double) * ns));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_elev, elev.data(), sizeof(double) * ns, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMalloc(&b->ediag, sizeof(double) * (size_t)dim));
    k_ediag<<<gdim, T, 0, st>>>(ns, h.norb, dim, b->cfg, d_elev, upp, b->ediag);

    // hop table
    b->maxhop = (int)pairs.size();
    std::vector<HopPairDev> hp(pairs.size());
    for (size_t i = 0; i < pairs.size(); i++) { hp[i].p = pairs[i].p; hp[i].q = pairs[i].q; hp[i].amp = pairs[i].amp; }
    HopPairDev *d_pairs = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_pairs, sizeof(HopPairDev) * (pairs.size() + 1)));
    if (!pairs.empty())
        CUDA_TRY(ctx, cudaMemcpyAsync(d_pairs, hp.data(), sizeof(HopPairDev) * pairs.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMalloc(&b->hop, sizeof(uint32_t) * (size_t)dim * (size_t)(b->maxhop > 0 ? b->maxhop : 1)));
    CUDA_TRY(ctx, cudaMalloc(&b->nhop, (size_t)dim));
    k_hops<<<gdim, T, 0, st>>>(dim, b->cfg, b->rank, (int)pairs.size(), d_pairs, b->maxhop, b->hop, b->nhop);
    std::vector<double> amp_signed(256, 0.0);
    for (size_t i = 0; i < amps.size(); i++) { amp_signed[2 * i] = amps[i]; amp_signed[2 * i + 1] = -amps[i]; }
    CUDA_TRY(ctx, cudaMalloc(&b->amp, sizeof(double) * 256));
    CUDA_TRY(ctx, cudaMemcpyAsync(b->amp, amp_signed.data(), sizeof(double) * 256, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    CUDA_TRY(ctx, cudaGetLastError());
    cudaFree(d_elev);
    cudaFree(d_pairs);
    ctx->bases[key] = b;
    out = b;
    return 0;
}

// X(u_imp, d_imp) = sum_a U_a nu_a nd_a + Ust sum_{a<b} (nu_a nd_b + nu_b nd_a)  (direct/HxVint.f90:4-15)
// + the HF constant sum_a U_a/4 + sum_{a<b} (Ust/4 + (Ust-Jh)/4)                  (HxVint.f90:29,34-35)
int upload_xtab(edgpu_ctx *ctx)
{
    const HamParams &h = ctx->ham;
    std::vector<double> x(32 * 32, 0.0);
    double cst = 0.0;
    if (h.hfmode) {
        for (int a = 0; a < h.norb; a++) cst += 0.25 * h.uloc[a];
        if (h.norb > 1)
            for (int a = 0; a < h.norb; a++)
                for (int b = a + 1; b < h.norb; b++) cst += 0.25 * h.ust + 0.25 * (h.ust - h.jh);
    }
    int nim = 1 << h.norb;
    for (int u = 0; u < nim; u++)
        for (int d = 0; d < nim; d++) {
            double e = cst;
            for (int a = 0; a < h.norb; a++) e += h.uloc[a] * ((u >> a) & 1) * ((d >> a) & 1);
            if (h.norb > 1)
                for (int a = 0; a < h.norb; a++)
                    for (int b = a + 1; b < h.norb; b++)
                        e += h.ust * (((u >> a) & 1) * ((d >> b) & 1) + ((u >> b) & 1) * ((d >> a) & 1));
            x[d * 32 + u] = e;
        }
    if (!ctx->d_xtab) CUDA_TRY(ctx, cudaMalloc(&ctx->d_xtab, sizeof(double) * 32 * 32));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_xtab, x.data(), sizeof(double) * 32 * 32, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int sector_map_kernel(edgpu_sector *s, int64_t first, int64_t count, uint64_t *d_out)
{
    const int T = 256;
    k_map<<<(unsigned)((count + T - 1) / T), T, 0, s->ctx->stream>>>(s->ctx->ham.ns, s->dim_up, first, count,
                                                                    s->up->cfg_ref, s->dw->cfg_ref, d_out);
    CUDA_TRY(s->ctx, cudaGetLastError());
    return 0;
}

int sector_map_check_kernel(edgpu_sector *s, unsigned long long *d_sum, unsigned long long *d_viol)
{
    k_map_check<<<s->ctx->sm_count * 8, 256, 0, s->ctx->stream>>>(s->ctx->ham.ns, s->nup, s->ndw, s->dim_up, s->dim,
                                                                 s->up->cfg_ref, s->dw->cfg_ref, d_sum, d_viol);
    CUDA_TRY(s->ctx, cudaGetLastError());
    return 0;
}
