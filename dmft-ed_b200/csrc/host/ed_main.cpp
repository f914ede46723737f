// ed_main.cpp -- host-side mirror of the reference solver phases that drive the Lanczos hot path.
//
// Mirrors (file:line relative to the reference root):
//   ed_init_solver / ed_solve            ED_MAIN.f90:61-101, 253-282
//   ed_diag_c / ed_post_diag             ED_DIAG.f90:49-251, 383-416
//   build_gf_normal / lanc_build_gf_normal_c / add_to_lanczos_gf_normal / build_sigma_normal
//                                        ED_GF_NORMAL.f90:18-31, 116-260, 580-632, 656-694
//   observables_impurity (normal core)   ED_OBSERVABLES.f90:105-162
//   delta_bath / invg0_bath              ED_BATH_FUNCTIONS.f90:221-258, 1784-1807
//   init_dmft_bath                       ED_BATH/dmft_aux.f90:105-127
//   allocate_grids                       ED_AUX_FUNX.f90:449-461
// Everything heavy goes through the C-ABI of include/edgpu.h (device kernels); only O(nlanc*L) pole sums, the
// tiny dense/tridiagonal eigenproblems and bookkeeping run on the host, as in the reference.
#include "../../../include/ed_b200.h"
#include "../../../include/edgpu.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

int host_tql2(int n, double *d, double *e, double *z);

typedef std::complex<double> cplx;

struct EdState {
    double e = 0;
    int nup = 0, ndw = 0;
    edgpu_sector *sec = nullptr;
    edgpu_vec *vec = nullptr;
    int worker = 0;                 // index of the worker context that owns sec / vec
};

struct EdChain {
    int iorb, ispin, isign, istate, nlanc, nused;
    double norm2;
    std::vector<double> alfa, beta;
};

struct ed_solver {
    ed_input in;
    edgpu_ctx *ctx = nullptr;
    std::string err;
    int Ns = 0;
    std::vector<double> hloc;                 // complex interleaved (Nspin,Nspin,Norb,Norb)
    std::vector<double> bath;
    std::vector<EdState> states;
    double zeta = 0, egs = 0;
    std::map<std::pair<int, int>, double> sector_e;
    std::map<std::pair<int, int>, int> sector_nlanc;      // Lanczos steps used per sector (0 = LAPACK sector)
    std::vector<std::pair<int, int>> mask;
    std::vector<EdChain> chains;
    std::vector<double> wm, wr;
    std::vector<cplx> Gmats, Greal, Smats, Sreal, G0mats, G0real;
    std::vector<double> dens, dens_up, dens_dw, docc, magz, sz2, n2;
    double s2tot = 0;
    // spin susceptibility (ED_VARS_GLOBAL.f90:144-146, ED_SETUP.f90:342-344)
    int Ltau = 0;
    std::vector<double> vm, tau, spinChi_tau;
    std::vector<cplx> spinChi_iv, spinChi_w;
    std::vector<double> densChi_tau, densChi_tot_tau;            // (ED_SETUP.f90:346-354)
    std::vector<cplx> densChi_iv, densChi_w, densChi_tot_iv, densChi_tot_w;
    double timings[4] = {0, 0, 0, 0};
    int rank = 0, nranks = 1;                 // ed_set_comm: sectors / states dealt over the ranks
    // work-unit parallelism on ONE GPU: the sectors of the scan and the GF chains are independent (the reference deals them
    // over MPI ranks, ED_MAIN.f90:598-636); here host threads drive worker contexts with streams of their own, so that the
    // launch latency and the host-side algebra of the many small sectors overlap.  wctx[0] == ctx.
    std::vector<edgpu_ctx *> wctx;
};

static int fail(ed_solver *s, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (s) s->err = buf;
    return 1;
}
#define GPU_TRY(s, call)                                                               \
    do {                                                                               \
        if ((call) != 0) return fail((s), "%s", edgpu_last_error((s)->ctx));           \
    } while (0)

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" void ed_input_defaults(ed_input *in)
{
    memset(in, 0, sizeof(*in));                 // ED_INPUT_VARS.f90:121-196
    in->Norb = 1; in->Nbath = 6; in->Nspin = 1;
    in->uloc[0] = 2.0;
    in->beta = 1000.0; in->xmu = 0.0; in->hfmode = 1;
    in->Lmats = 5000; in->Lreal = 5000;
    in->wini = -5.0; in->wfin = 5.0; in->eps = 0.01;
    in->gs_threshold = 1e-9; in->hwband = 2.0;
    in->lanc_method = 0; in->lanc_nstates_sector = 6; in->lanc_nstates_total = 1;
    in->lanc_niter = 512; in->lanc_ngfiter = 200; in->lanc_tolerance = 1e-12; in->lanc_dim_threshold = 256;
    in->chispin_flag = 0; in->Ltau = 1000; in->chidens_flag = 0;
    in->ed_twin = 0; in->ed_sparse_H = 1; in->ed_verbose = 3;
}

extern "C" int32_t ed_get_bath_dimension(const ed_input *in)
{
    return 2 * in->Nspin * in->Norb * in->Nbath;            // user_aux.f90:11-30 (normal/normal: e and v)
}

static void init_dmft_bath(const ed_input &in, double *bath)
{
    // ED_BATH/dmft_aux.f90:105-127 with ed_bath_noise_thr = 0
    const int N = in.Nbath;
    std::vector<double> e(N + 2, 0.0), v(N + 2, 0.0);
    e[1] = -in.hwband;
    e[N] = in.hwband;
    const int Nh = N / 2;
    if (N % 2 == 0 && N >= 4) {
        const double de = in.hwband / std::max(Nh - 1, 1);
        e[Nh] = -1e-3;
        e[Nh + 1] = 1e-3;
        for (int i = 2; i <= Nh - 1; i++) { e[i] = -in.hwband + (i - 1) * de; e[N - i + 1] = in.hwband - (i - 1) * de; }
    } else if (N % 2 != 0 && N >= 3) {
        const double de = in.hwband / Nh;
        e[Nh + 1] = 0.0;
        for (int i = 2; i <= Nh; i++) { e[i] = -in.hwband + (i - 1) * de; e[N - i + 1] = in.hwband - (i - 1) * de; }
    }
    for (int i = 1; i <= N; i++) v[i] = std::max(0.1, 1.0 / std::sqrt((double)N));
    const int nb = in.Nspin * in.Norb * N;
    for (int is = 0; is < in.Nspin; is++)
        for (int io = 0; io < in.Norb; io++)
            for (int k = 1; k <= N; k++) {
                bath[(is * in.Norb + io) * N + (k - 1)] = e[k];
                bath[nb + (is * in.Norb + io) * N + (k - 1)] = v[k];
            }
}

extern "C" int ed_init_solver(const ed_input *in, int device, void *stream, double *bath, int32_t bath_len,
                              const double *hloc_cplx, ed_solver **out)
{
    if (!in || !out) return 1;
    *out = nullptr;
    ed_solver *s = new ed_solver();
    s->in = *in;
    // ed_checks_global (ED_SETUP.f90:36-83), the subset relevant to ed_mode=normal
    if (in->Nspin > 2) { delete s; return 1; }
    if (in->lanc_method == 1 && (in->lanc_nstates_total > 1 || in->lanc_nstates_sector > 1)) {
        fprintf(stderr, "ED ERROR: lanc_method==lanczos available only for lanc_nstates_total==1 and lanc_nstates_sector==1, T=0\n");
        delete s;
        return 1;
    }
    if (in->lanc_nstates_total > 1) { fprintf(stderr, "ED ERROR (GPU path): finite temperature (lanc_nstates_total>1) is not supported yet\n"); delete s; return 1; }
    if (in->ed_twin) { fprintf(stderr, "ED ERROR (GPU path): ed_twin=T is not supported yet\n"); delete s; return 1; }
    if (bath_len != ed_get_bath_dimension(in)) { fprintf(stderr, "ED ERROR: wrong bath dimensions\n"); delete s; return 1; }
    s->Ns = (in->Nbath + 1) * in->Norb;
    edgpu_params p;
    memset(&p, 0, sizeof(p));
    p.norb = in->Norb; p.nbath = in->Nbath; p.nspin = in->Nspin; p.hfmode = in->hfmode;
    p.layout = in->gpu_layout; p.hxv_kernel = in->gpu_hxv_kernel;
    p.reserved[1] = in->ed_sparse_H ? 1 : 0;          // stored H (CSR) lives in the single-tile layout
    if (edgpu_init(&p, device, stream, &s->ctx) != 0) {
        fprintf(stderr, "ed_init_solver: %s\n", edgpu_last_error(nullptr));
        delete s;
        return 1;
    }
    s->wctx.push_back(s->ctx);
    int nworkers = in->reserved[2] > 0 ? in->reserved[2] : 4;
    if (const char *e = getenv("ED_B200_WORKERS")) nworkers = std::max(1, atoi(e));
    nworkers = std::min(nworkers, 16);
    p.reserved[2] = 1;                                // worker contexts own a non-blocking stream
    for (int w = 1; w < nworkers; w++) {
        edgpu_ctx *c = nullptr;
        if (edgpu_init(&p, device, nullptr, &c) != 0) {
            fprintf(stderr, "ed_init_solver: %s\n", edgpu_last_error(nullptr));
            for (auto *q : s->wctx) edgpu_finalize(q);
            delete s;
            return 1;
        }
        s->wctx.push_back(c);
    }
    const size_t nh = (size_t)in->Nspin * in->Nspin * in->Norb * in->Norb;
    s->hloc.assign(2 * nh, 0.0);
    if (hloc_cplx) memcpy(s->hloc.data(), hloc_cplx, sizeof(double) * 2 * nh);
    if (bath) init_dmft_bath(*in, bath);
    *out = s;
    return 0;
}

static void free_states(ed_solver *s)
{
    std::vector<edgpu_sector *> secs;
    for (auto &st : s->states) {
        if (st.vec) edgpu_vec_free(st.vec);
        if (st.sec && std::find(secs.begin(), secs.end(), st.sec) == secs.end()) secs.push_back(st.sec);
    }
    for (auto *q : secs) edgpu_sector_free(q);
    s->states.clear();
}

extern "C" int ed_comm_unique_id(ed_solver *s, unsigned char id[128])
{
    if (!s) return 1;
    GPU_TRY(s, edgpu_comm_unique_id(s->ctx, id));
    return 0;
}

extern "C" int ed_set_comm(ed_solver *s, const unsigned char id[128], int32_t rank, int32_t nranks)
{
    if (!s) return 1;
    GPU_TRY(s, edgpu_comm_init(s->ctx, id, rank, nranks));
    s->rank = rank; s->nranks = nranks;
    return 0;
}

// sum / min of host arrays over the ranks of the distributed solve (no-op on one rank)
static int allsum(ed_solver *s, double *v, size_t n) { return s->nranks > 1 ? edgpu_comm_allreduce_host(s->ctx, v, (int64_t)n, 0) : 0; }
static int allsum(ed_solver *s, std::vector<double> &v) { return allsum(s, v.data(), v.size()); }
static int allsum(ed_solver *s, std::vector<cplx> &v) { return allsum(s, reinterpret_cast<double *>(v.data()), 2 * v.size()); }

extern "C" int ed_finalize_solver(ed_solver *s)
{
    if (!s) return 0;
    free_states(s);
    for (size_t w = s->wctx.size(); w-- > 1;) edgpu_finalize(s->wctx[w]);
    edgpu_finalize(s->ctx);
    delete s;
    return 0;
}

extern "C" const char *ed_last_error(const ed_solver *s) { return s ? s->err.c_str() : "null solver"; }

extern "C" int ed_set_sectors_mask(ed_solver *s, const int32_t *pairs, int32_t n)
{
    if (!s) return 1;
    s->mask.clear();
    for (int i = 0; i < n; i++) s->mask.push_back({pairs[2 * i], pairs[2 * i + 1]});
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// dense symmetric eigensolver: Householder tridiagonalisation + tql2 (LAPACK eigh in the reference)
// ---------------------------------------------------------------------------------------------------------
static void tred2(int n, std::vector<double> &a, std::vector<double> &d, std::vector<double> &e, bool vectors = true)
{
    // a: row-major n x n symmetric; on return a holds the orthogonal transformation Q (a[k*n+i] = Q_ki)
    // vectors = false: the reduction only (d, e of the tridiagonal matrix; a is left in its reduced form)
    auto A = [&](int i, int j) -> double & { return a[(size_t)i * n + j]; };
    for (int i = n - 1; i >= 1; i--) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; k++) scale += std::fabs(A(i, k));
            if (scale == 0.0) e[i] = A(i, l);
            else {
                for (int k = 0; k <= l; k++) { A(i, k) /= scale; h += A(i, k) * A(i, k); }
                double f = A(i, l);
                double g = (f >= 0.0) ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; j++) {
                    A(j, i) = A(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; k++) g += A(j, k) * A(i, k);
                    for (int k = j + 1; k <= l; k++) g += A(k, j) * A(i, k);
                    e[j] = g / h;
                    f += e[j] * A(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; j++) {
                    f = A(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; k++) A(j, k) -= (f * e[k] + g * A(i, k));
                }
            }
        } else e[i] = A(i, l);
        d[i] = h;
    }
    d[0] = 0.0;
    e[0] = 0.0;
    if (!vectors) {
        // the diagonal is final after the reduction (the accumulation below touches A(k, j) with k, j < i only before it reads A(i, i))
        for (int i = 0; i < n; i++) d[i] = A(i, i);
        return;
    }
    for (int i = 0; i < n; i++) {
        const int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; j++) {
                double g = 0.0;
                for (int k = 0; k <= l; k++) g += A(i, k) * A(k, j);
                for (int k = 0; k <= l; k++) A(k, j) -= g * A(k, i);
            }
        }
        d[i] = A(i, i);
        A(i, i) = 1.0;
        for (int j = 0; j <= l; j++) A(j, i) = A(i, j) = 0.0;
    }
}

extern "C" int ed_host_eigh(int32_t n, double *a_colmajor, double *w)
{
    if (n < 1) return 1;
    if (n == 1) { w[0] = a_colmajor[0]; a_colmajor[0] = 1.0; return 0; }
    std::vector<double> a((size_t)n * n), d(n), e(n), z((size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) a[(size_t)i * n + j] = 0.5 * (a_colmajor[i + (size_t)n * j] + a_colmajor[j + (size_t)n * i]);
    tred2(n, a, d, e);
    for (int k = 0; k < n; k++)
        for (int i = 0; i < n; i++) z[k + (size_t)n * i] = a[(size_t)k * n + i];
    int ierr = host_tql2(n, d.data(), e.data(), z.data());
    if (ierr) return ierr;
    memcpy(w, d.data(), sizeof(double) * n);
    memcpy(a_colmajor, z.data(), sizeof(double) * (size_t)n * n);
    return 0;
}

// eigenvalues only (same reduction and QL iteration as ed_host_eigh, without accumulating the transformations: ~4x cheaper);
// the values are bit-identical to those of ed_host_eigh
extern "C" int ed_host_eigvals(int32_t n, const double *a_colmajor, double *w)
{
    if (n < 1) return 1;
    if (n == 1) { w[0] = a_colmajor[0]; return 0; }
    std::vector<double> a((size_t)n * n), d(n), e(n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) a[(size_t)i * n + j] = 0.5 * (a_colmajor[i + (size_t)n * j] + a_colmajor[j + (size_t)n * i]);
    tred2(n, a, d, e, false);
    int ierr = host_tql2(n, d.data(), e.data(), nullptr);
    if (ierr) return ierr;
    memcpy(w, d.data(), sizeof(double) * n);
    return 0;
}

extern "C" int ed_host_eigh_tridiag(int32_t n, const double *diag, const double *sub, double *w, double *z)
{
    std::vector<double> e(n, 0.0);
    for (int i = 0; i < n; i++) w[i] = diag[i];
    for (int i = 1; i < n; i++) e[i] = sub[i];
    for (size_t q = 0; q < (size_t)n * n; q++) z[q] = 0.0;
    for (int i = 0; i < n; i++) z[i + (size_t)n * i] = 1.0;
    return host_tql2(n, w, e.data(), z);
}

// ---------------------------------------------------------------------------------------------------------
// ed_diag_c (ED_DIAG.f90:49-251), T=0
// ---------------------------------------------------------------------------------------------------------
static void insert_state(std::vector<EdState> &list, const EdState &st)
{
    // es_insert_state_c (ED_EIGENSPACE.f90:169-218): ordered list, new entry goes before the first c with e <= c%e
    size_t pos = 0;
    while (pos < list.size() && !(st.e <= list[pos].e)) pos++;
    list.insert(list.begin() + pos, st);
}

// ED_B200_TRACE=1: per-sector / per-chain wall times on stderr (where does a solve spend its time)
static bool trace_on()
{
    static const bool on = getenv("ED_B200_TRACE") != nullptr;
    return on;
}

// One worker's share of the scan: the reference's loop body (ED_DIAG.f90:86-240) over `secs` with a LOCAL running minimum.
struct DiagLocal {
    std::vector<EdState> states;
    std::map<std::pair<int, int>, double> sector_e;
    std::map<std::pair<int, int>, int> sector_nlanc;
    std::string err;
};

static int diag_sectors(ed_solver *s, int worker, const std::vector<std::pair<int, int>> &secs, DiagLocal &L)
{
    const ed_input &in = s->in;
    edgpu_ctx *ctx = s->wctx[worker];
    auto bad = [&](const char *what) { L.err = std::string(what) + ": " + edgpu_last_error(ctx); return 1; };
    if (edgpu_bind_thread(ctx)) return bad("ed_diag");
    double oldzero = 1000.0;
    for (const auto &nn : secs) {
            const int nup = nn.first, ndw = nn.second;
            edgpu_sector *sec = nullptr;
            const double tr0 = now_s();
            if (edgpu_sector_build(ctx, nup, ndw, &sec)) return bad("ed_diag");
            const double tr1 = now_s();
            int64_t dim = 0;
            edgpu_sector_dim(sec, &dim, nullptr, nullptr);
            int64_t neigen, nitermax;
            if (in.lanc_method == 1) { neigen = 1; nitermax = std::min<int64_t>(dim, in.lanc_niter); }      // :93-97
            else { neigen = std::min<int64_t>(dim, std::min<int64_t>(dim, in.lanc_nstates_sector)); nitermax = std::min<int64_t>(dim, in.lanc_niter); }
            bool lanc_solve = true;
            if (neigen == dim) lanc_solve = false;                                                        // :100
            if (dim <= std::max(in.lanc_dim_threshold, 1)) lanc_solve = false;                             // :101
            std::vector<double> evals, H;
            std::vector<edgpu_vec *> evecs;
            int rc = 0;
            auto need_vec = [&](size_t i) -> int {       // eigenvector i of a dense sector -> device handle
                if (evecs[i]) return 0;
                if (H.empty()) return 1;
                if (edgpu_vec_alloc(sec, &evecs[i])) return 1;
                return edgpu_vec_upload(evecs[i], H.data() + i * (size_t)dim, 0);
            };
            if (lanc_solve) {
                if (in.ed_sparse_H) rc = edgpu_sector_build_csr(sec);                                      // ED_HAMILTONIAN.f90:85-92
                edgpu_vec *v = nullptr;
                if (!rc) rc = edgpu_vec_alloc(sec, &v);
                if (!rc) rc = edgpu_vec_fill_uniform(v, 1234567ull);      // start vector (reference: random_number)
                double e0 = 0;
                int nl = 0;
                if (!rc && in.lanc_method == 0 && neigen > 1) {
                    // sp_eigh (ARPACK, ED_DIAG.f90:149-166): Neigen pairs from a basis of Nblock vectors (:98
                    // Nblock = min(dim, lanc_ncv_factor*Neigen + lanc_ncv_add), defaults 10 and 0) -- needed to keep
                    // EVERY member of a ground state that is degenerate inside the sector (:224-235)
                    const int factor = in.reserved[0] > 0 ? in.reserved[0] : 10, add = in.reserved[1];
                    const int ncv = (int)std::min<int64_t>(dim, (int64_t)factor * neigen + add);
                    edgpu_vec_free(v); v = nullptr;
                    std::vector<double> ev(neigen);
                    std::vector<edgpu_vec *> vv(neigen, nullptr);
                    int nconv = 0, nmv = 0;
                    rc = edgpu_lanczos_eigs(sec, (int)neigen, ncv, std::max(1, in.lanc_niter), in.lanc_tolerance, 1234567ull, ev.data(), vv.data(), &nconv, &nmv);
                    if (rc) { bad("ed_diag"); for (auto *q : vv) if (q) edgpu_vec_free(q); edgpu_sector_free(sec); return 1; }
                    for (int64_t i = 0; i < neigen; i++) { evals.push_back(ev[i]); evecs.push_back(vv[i]); }
                    L.sector_nlanc[{nup, ndw}] = nmv;
                } else {
                    if (!rc) rc = edgpu_lanczos_gs(sec, v, (int)nitermax, in.lanc_tolerance, 10, &e0, &nl, nullptr, nullptr);
                    if (rc) { bad("ed_diag"); if (v) edgpu_vec_free(v); edgpu_sector_free(sec); return 1; }
                    // lanc_nstates_sector = 1 (or lanc_method=lanczos): the lowest pair by plain Lanczos (sp_lanc_eigh, :173-181)
                    evals.push_back(e0);
                    evecs.push_back(v);
                    L.sector_nlanc[{nup, ndw}] = nl;
                }
                if (in.ed_sparse_H) edgpu_sector_drop_csr(sec);
            } else {
                H.assign((size_t)dim * dim, 0.0);
                std::vector<double> w(dim);
                rc = edgpu_sector_dense(sec, H.data());                                                   // :188-193
                if (rc) { bad("ed_diag"); edgpu_sector_free(sec); return 1; }
                // Most LAPACK sectors cannot hold a ground state: their eigenVALUES decide that (a state is kept only if it lies
                // within gs_threshold of the running minimum or below it, :224-235), the vectors are computed only then.
                bool vectors = true;
                if (dim >= 16) {
                    if (ed_host_eigvals((int)dim, H.data(), w.data())) { L.err = "ed_diag: dense eigh failed"; edgpu_sector_free(sec); return 1; }
                    vectors = w[0] <= oldzero + in.gs_threshold;
                }
                if (vectors) {
                    if (ed_host_eigh((int)dim, H.data(), w.data())) { L.err = "ed_diag: dense eigh failed"; edgpu_sector_free(sec); return 1; }
                } else H.clear();
                for (int64_t i = 0; i < neigen; i++) { evals.push_back(w[i]); evecs.push_back(nullptr); }   // uploaded on demand
            }
            L.sector_e[{nup, ndw}] = evals.empty() ? 0.0 : evals[0];
            const double tr2 = now_s();
            bool used = false;
            for (size_t i = 0; i < evals.size(); i++) {                                                   // :224-235
                const double enemin = evals[i];
                EdState st;
                st.e = enemin; st.nup = nup; st.ndw = ndw; st.sec = sec; st.worker = worker;
                const bool lower = enemin < oldzero - 10.0 * in.gs_threshold;
                const bool degen = !lower && std::fabs(enemin - oldzero) <= in.gs_threshold;
                if ((lower || degen) && need_vec(i)) return bad("ed_diag");
                st.vec = evecs[i];
                if (lower) {
                    oldzero = enemin;
                    // es_free_espace: drop every stored state
                    std::vector<edgpu_sector *> old;
                    for (auto &o : L.states) {
                        if (o.vec) edgpu_vec_free(o.vec);
                        if (o.sec != sec && std::find(old.begin(), old.end(), o.sec) == old.end()) old.push_back(o.sec);
                    }
                    for (auto *q : old) edgpu_sector_free(q);
                    L.states.clear();
                    if (!st.vec) { L.err = "ed_diag: internal error (missing eigenvector)"; return 1; }
                    insert_state(L.states, st);
                    used = true;
                } else if (degen) {
                    oldzero = std::min(oldzero, enemin);
                    if (!st.vec) { L.err = "ed_diag: internal error (missing eigenvector)"; return 1; }
                    insert_state(L.states, st);
                    used = true;
                } else if (st.vec) {
                    edgpu_vec_free(st.vec);
                }
            }
            if (!used) edgpu_sector_free(sec);
            if (trace_on())
                fprintf(stderr, "[ed_diag w%d] (%d,%d) dim %lld %s build %.2f ms solve %.2f ms (nlanc %d) keep/free %.2f ms\n", worker, nup, ndw, (long long)dim,
                        lanc_solve ? "lanczos" : "dense", (tr1 - tr0) * 1e3, (tr2 - tr1) * 1e3, lanc_solve ? L.sector_nlanc[{nup, ndw}] : 0, (now_s() - tr2) * 1e3);
    }
    return 0;
}

// Keeps, of the states the workers (and ranks) found with their local running minima, those within gs_threshold of `emin`.
static void keep_ground_states(ed_solver *s, std::vector<EdState> &all, double emin)
{
    std::vector<EdState> keep;
    std::vector<edgpu_sector *> used, drop;
    for (auto &st : all) {
        if (std::fabs(st.e - emin) <= s->in.gs_threshold) { keep.push_back(st); used.push_back(st.sec); }
        else { if (st.vec) edgpu_vec_free(st.vec); drop.push_back(st.sec); }
    }
    for (auto *q : drop)
        if (std::find(used.begin(), used.end(), q) == used.end()) { edgpu_sector_free(q); used.push_back(q); }
    // the list order of the serial scan: insertion in sector order (ties keep the reference's "new entry first" rule)
    std::stable_sort(keep.begin(), keep.end(), [](const EdState &a, const EdState &b) { return a.nup != b.nup ? a.nup < b.nup : a.ndw < b.ndw; });
    s->states.clear();
    for (auto &st : keep) insert_state(s->states, st);
}

static int ed_diag(ed_solver *s)
{
    const int Ns = s->Ns;
    free_states(s);
    s->sector_e.clear();
    s->sector_nlanc.clear();
    // sectors of this rank in the reference's isector order (ED_SETUP.f90:382-393), dealt round-robin to the workers
    const int nw = (int)s->wctx.size();
    std::vector<std::vector<std::pair<int, int>>> mine(nw);
    int isec = -1, cnt = 0;
    for (int nup = 0; nup <= Ns; nup++)
        for (int ndw = 0; ndw <= Ns; ndw++) {
            if (!s->mask.empty() && std::find(s->mask.begin(), s->mask.end(), std::make_pair(nup, ndw)) == s->mask.end()) continue;
            isec++;
            if (s->nranks > 1 && isec % s->nranks != s->rank) continue;        // distributed scan: this sector belongs to another rank
            mine[cnt++ % nw].push_back({nup, ndw});
        }
    std::vector<DiagLocal> loc(nw);
    std::vector<int> rcs(nw, 0);
    {
        std::vector<std::thread> th;
        for (int w = 1; w < nw; w++) th.emplace_back([&, w] { rcs[w] = diag_sectors(s, w, mine[w], loc[w]); });
        rcs[0] = diag_sectors(s, 0, mine[0], loc[0]);
        for (auto &t : th) t.join();
    }
    std::vector<EdState> all;
    int bad = -1;
    for (int w = 0; w < nw; w++) {
        if (rcs[w] && bad < 0) bad = w;
        for (auto &st : loc[w].states) all.push_back(st);
        for (auto &kv : loc[w].sector_e) s->sector_e[kv.first] = kv.second;
        for (auto &kv : loc[w].sector_nlanc) s->sector_nlanc[kv.first] = kv.second;
    }
    if (bad >= 0) { s->states = all; free_states(s); return fail(s, "%s", loc[bad].err.c_str()); }
    double emin = 1e300;
    for (auto &st : all) emin = std::min(emin, st.e);
    if (s->nranks > 1) {
        // the running minimum of :224-235 was local: agree on the global ground energy, keep the local states within
        // gs_threshold of it, and share the sector energies
        GPU_TRY(s, edgpu_comm_allreduce_host(s->ctx, &emin, 1, 1));
        keep_ground_states(s, all, emin);
        double cnt = (double)s->states.size();
        GPU_TRY(s, edgpu_comm_allreduce_host(s->ctx, &cnt, 1, 0));
        const int nsec = (Ns + 1) * (Ns + 1);
        std::vector<double> tab(2 * (size_t)nsec, 0.0);
        for (auto &kv : s->sector_e) { tab[kv.first.first * (Ns + 1) + kv.first.second] = kv.second; tab[nsec + kv.first.first * (Ns + 1) + kv.first.second] = 1.0; }
        GPU_TRY(s, allsum(s, tab));
        for (int a = 0; a <= Ns; a++)
            for (int b = 0; b <= Ns; b++)
                if (tab[nsec + a * (Ns + 1) + b] > 0.5) s->sector_e[{a, b}] = tab[a * (Ns + 1) + b];
        if (cnt < 0.5) return fail(s, "ed_diag: no state found");
        s->egs = emin;
        s->zeta = cnt;
        return 0;
    }
    keep_ground_states(s, all, emin);
    if (s->states.empty()) return fail(s, "ed_diag: no state found");
    // ed_post_diag (ED_DIAG.f90:403-416), T=0
    s->egs = s->states[0].e;
    for (auto &st : s->states) s->egs = std::min(s->egs, st.e);
    s->zeta = (double)s->states.size();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Green's functions
// ---------------------------------------------------------------------------------------------------------
static inline size_t gidx(const ed_input &in, int ispin, int jspin, int iorb, int jorb, int i)
{
    // (Nspin,Nspin,Norb,Norb,L) column-major
    return (size_t)ispin + in.Nspin * ((size_t)jspin + in.Nspin * ((size_t)iorb + in.Norb * ((size_t)jorb + (size_t)in.Norb * i)));
}

static void add_to_lanczos_gf(const ed_solver *s, std::vector<cplx> &Gm, std::vector<cplx> &Gr, double vnorm2, double Ei, const std::vector<double> &alanc,
                              const std::vector<double> &blanc, int isign, int iorb, int jorb, int ispin)
{
    // add_to_lanczos_gf_normal (ED_GF_NORMAL.f90:580-632), T=0: pesoBZ = vnorm2/zeta_function
    const ed_input &in = s->in;
    const int nlanc = (int)alanc.size();
    const double pesoBZ = vnorm2 / s->zeta;
    std::vector<double> w(nlanc), Z((size_t)nlanc * nlanc);
    ed_host_eigh_tridiag(nlanc, alanc.data(), blanc.data(), w.data(), Z.data());
    for (int j = 0; j < nlanc; j++) {
        const double de = w[j] - Ei;
        const double peso = pesoBZ * Z[(size_t)nlanc * j] * Z[(size_t)nlanc * j];      // Z(1,j)^2
        if (peso == 0.0) continue;
        for (int i = 0; i < in.Lmats; i++)
            Gm[gidx(in, ispin, ispin, iorb, jorb, i)] += peso / (cplx(0.0, s->wm[i]) - (double)isign * de);
        for (int i = 0; i < in.Lreal; i++)
            Gr[gidx(in, ispin, ispin, iorb, jorb, i)] += peso / (cplx(s->wr[i], in.eps) - (double)isign * de);
    }
}

// one GF chain: seed c / c^dagger |state>, tridiagonalise in the target sector, add the poles (ED_GF_NORMAL.f90:116-260)
struct GfUnit { int ispin, iorb, istate, pass; };

static int gf_unit(ed_solver *s, const GfUnit &u, EdChain &ch, std::vector<cplx> &Gm, std::vector<cplx> &Gr, std::string &err)
{
    const ed_input &in = s->in;
    const int Ns = s->Ns;
    const int ispin = u.ispin, iorb = u.iorb;
    EdState &st = s->states[u.istate];
    edgpu_ctx *ctx = s->wctx[st.worker];
    const int isite = (ispin == 0) ? iorb + 1 : iorb + 1 + Ns;                // impIndex, ED_SETUP.f90:443-446
    const int dagger = u.pass == 0 ? 1 : 0, isign = u.pass == 0 ? 1 : -1;     // :150 cdg first, :203 c
    const int jup = st.nup + (ispin == 0 ? (dagger ? 1 : -1) : 0);
    const int jdw = st.ndw + (ispin == 1 ? (dagger ? 1 : -1) : 0);
    ch.iorb = iorb; ch.ispin = ispin; ch.isign = isign; ch.istate = u.istate; ch.norm2 = 0; ch.nlanc = 0; ch.nused = -1;
    if (jup < 0 || jup > Ns || jdw < 0 || jdw > Ns) return 0;                 // getCDGsector/getCsector == 0: no chain
    edgpu_sector *sj = nullptr;
    edgpu_vec *vv = nullptr;
    const double tr0 = now_s();
    if (edgpu_sector_build(ctx, jup, jdw, &sj)) { err = edgpu_last_error(ctx); return 1; }
    const double tr1 = now_s();
    int64_t jdim = 0;
    edgpu_sector_dim(sj, &jdim, nullptr, nullptr);
    int rc = edgpu_vec_alloc(sj, &vv);
    double norm2 = 0;
    if (!rc) rc = edgpu_apply_c(st.sec, sj, isite, dagger, st.vec, vv, 1, &norm2);     // :159-174
    ch.norm2 = norm2;
    ch.nlanc = (int)std::min<int64_t>(jdim, in.lanc_ngfiter);                           // :177
    ch.alfa.assign(ch.nlanc, 0.0);
    ch.beta.assign(ch.nlanc, 0.0);
    ch.nused = 0;
    if (!rc && in.ed_sparse_H && jdim > 1) rc = edgpu_sector_build_csr(sj);              // build_Hv_sector :180
    if (!rc && norm2 > 0.0)
        rc = edgpu_lanczos_tridiag(sj, vv, ch.nlanc, 1e-13, ch.alfa.data(), ch.beta.data(), &ch.nused);
    const double tr2 = now_s();
    if (rc) err = edgpu_last_error(ctx);
    if (vv) edgpu_vec_free(vv);
    edgpu_sector_free(sj);
    if (rc) return 1;
    const double tr3 = now_s();
    if (norm2 > 0.0) add_to_lanczos_gf(s, Gm, Gr, norm2, st.e, ch.alfa, ch.beta, isign, iorb, iorb, ispin);  // :194
    if (trace_on())
        fprintf(stderr, "[build_gf w%d] state %d orb %d spin %d sign %d -> (%d,%d) dim %lld: build %.2f ms apply+chain %.2f ms (%d steps) free %.2f ms poles %.2f ms\n",
                st.worker, u.istate, iorb, ispin, isign, jup, jdw, (long long)jdim, (tr1 - tr0) * 1e3, (tr2 - tr1) * 1e3, ch.nused, (tr3 - tr2) * 1e3, (now_s() - tr3) * 1e3);
    return 0;
}

static int build_gf(ed_solver *s)
{
    const ed_input &in = s->in;
    const size_t nblk = (size_t)in.Nspin * in.Nspin * in.Norb * in.Norb;
    // allocate_grids (ED_AUX_FUNX.f90:449-461)
    s->wm.resize(in.Lmats);
    s->wr.resize(in.Lreal);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < in.Lmats; i++) s->wm[i] = pi / in.beta * (2.0 * (i + 1) - 1.0);
    for (int i = 0; i < in.Lreal; i++) s->wr[i] = in.Lreal > 1 ? in.wini + (in.wfin - in.wini) * (double)i / (double)(in.Lreal - 1) : in.wini;
    s->Gmats.assign(nblk * in.Lmats, cplx(0, 0));
    s->Greal.assign(nblk * in.Lreal, cplx(0, 0));
    s->chains.clear();
    // the chains in the reference's loop order (build_gf_normal :24-31, :132, :150/:203); each runs on the worker context
    // that owns its state, the workers side by side, every worker summing into its own G
    std::vector<GfUnit> units;
    for (int ispin = 0; ispin < in.Nspin; ispin++)
        for (int iorb = 0; iorb < in.Norb; iorb++)
            for (size_t istate = 0; istate < s->states.size(); istate++)
                for (int pass = 0; pass < 2; pass++) units.push_back({ispin, iorb, (int)istate, pass});
    const int nw = (int)s->wctx.size();
    std::vector<EdChain> chains(units.size());
    std::vector<std::vector<cplx>> Gm(nw), Gr(nw);
    std::vector<std::string> errs(nw);
    std::vector<int> rcs(nw, 0);
    auto run = [&](int w) {
        bool any = false;
        for (size_t k = 0; k < units.size() && !rcs[w]; k++) {
            if (s->states[units[k].istate].worker != w) continue;
            if (!any) {
                any = true;
                Gm[w].assign(s->Gmats.size(), cplx(0, 0)); Gr[w].assign(s->Greal.size(), cplx(0, 0));
                if (edgpu_bind_thread(s->wctx[w])) { errs[w] = edgpu_last_error(s->wctx[w]); rcs[w] = 1; return; }
            }
            rcs[w] = gf_unit(s, units[k], chains[k], Gm[w], Gr[w], errs[w]);
        }
    };
    {
        std::vector<std::thread> th;
        for (int w = 1; w < nw; w++) th.emplace_back(run, w);
        run(0);
        for (auto &t : th) t.join();
    }
    for (int w = 0; w < nw; w++)
        if (rcs[w]) return fail(s, "%s", errs[w].c_str());
    for (int w = 0; w < nw; w++) {
        if (Gm[w].empty()) continue;
        for (size_t i = 0; i < s->Gmats.size(); i++) s->Gmats[i] += Gm[w][i];
        for (size_t i = 0; i < s->Greal.size(); i++) s->Greal[i] += Gr[w][i];
    }
    for (auto &ch : chains)
        if (ch.nused >= 0) s->chains.push_back(std::move(ch));
    // distributed solve: every rank summed over ITS states, G is the sum over all of them
    GPU_TRY(s, allsum(s, s->Gmats));
    GPU_TRY(s, allsum(s, s->Greal));
    return 0;
}

// add_to_lanczos_spinChi (ED_GF_CHISPIN.f90:247-319), T = 0 (pesoBZ = 1).  isign = +1 and -1 in one go.
static void add_to_lanczos_spinchi(ed_solver *s, double vnorm, double Ei, const std::vector<double> &alanc,
                                   const std::vector<double> &blanc, int iorb)
{
    const ed_input &in = s->in;
    const int nlanc = (int)alanc.size(), L1 = in.Norb + 1;
    const double beta = in.beta;
    const double pesoF = vnorm * vnorm / s->zeta;
    std::vector<double> w(nlanc), Z((size_t)nlanc * nlanc);
    ed_host_eigh_tridiag(nlanc, alanc.data(), blanc.data(), w.data(), Z.data());
    for (int isign = 1; isign >= -1; isign -= 2)
        for (int j = 0; j < nlanc; j++) {
            const double dE = w[j] - Ei;
            const double peso = pesoF * Z[(size_t)nlanc * j] * Z[(size_t)nlanc * j];
            const double ex = std::exp(-beta * dE);
            s->spinChi_iv[iorb] += (beta * dE < 1e-1) ? peso * beta : peso * (1.0 - ex) / dE;          // :273-277, :294-298
            for (int i = 1; i <= in.Lmats; i++)
                s->spinChi_iv[iorb + (size_t)L1 * i] += isign == 1 ? peso * (ex - 1.0) / (cplx(0.0, s->vm[i]) - dE)
                                                                   : peso * (1.0 - ex) / (cplx(0.0, s->vm[i]) + dE);
            for (int i = 0; i <= s->Ltau; i++)
                s->spinChi_tau[iorb + (size_t)L1 * i] += isign == 1 ? peso * std::exp(-s->tau[i] * dE) : peso * std::exp(-(beta - s->tau[i]) * dE);
            for (int i = 0; i < in.Lreal; i++)
                s->spinChi_w[iorb + (size_t)L1 * i] += isign == 1 ? peso * (ex - 1.0) / (cplx(s->wr[i], in.eps) - dE)
                                                                  : peso * (1.0 - ex) / (cplx(s->wr[i], in.eps) + dE);
        }
}

// build_chi_spin (ED_GF_CHISPIN.f90:22-40): one chain per kept state and orbital, seed S_z,a |gs> in the state's own sector
// (lanc_ed_build_spinChi_c :57-141), plus S_z^tot for Norb > 1 (lanc_ed_build_spinChi_tot_c :160-237).
static int build_chi_spin(ed_solver *s)
{
    const ed_input &in = s->in;
    const int L1 = in.Norb + 1;
    const double pi = 3.14159265358979323846;
    s->Ltau = std::max((int)in.beta, (int)in.Ltau);                                  // ED_INPUT_VARS.f90:211
    s->vm.resize(in.Lmats + 1);
    s->tau.resize(s->Ltau + 1);
    for (int i = 0; i <= in.Lmats; i++) s->vm[i] = pi / in.beta * 2.0 * (double)i;   // ED_AUX_FUNX.f90:456-458
    for (int i = 0; i <= s->Ltau; i++) s->tau[i] = s->Ltau > 0 ? in.beta * (double)i / (double)s->Ltau : 0.0;
    s->spinChi_iv.assign((size_t)L1 * (in.Lmats + 1), cplx(0, 0));
    s->spinChi_tau.assign((size_t)L1 * (s->Ltau + 1), 0.0);
    s->spinChi_w.assign((size_t)L1 * in.Lreal, cplx(0, 0));
    if (!in.chispin_flag) return 0;
    const int nchan = in.Norb + (in.Norb > 1 ? 1 : 0);
    for (int ic = 0; ic < nchan; ic++) {
        const bool tot = ic == in.Norb;
        for (size_t istate = 0; istate < s->states.size(); istate++) {
            EdState &st = s->states[istate];
            int64_t idim = 0;
            edgpu_sector_dim(st.sec, &idim, nullptr, nullptr);
            edgpu_vec *vv = nullptr;
            GPU_TRY(s, edgpu_vec_alloc(st.sec, &vv));
            double nrm = 0;
            int rc = edgpu_apply_sz(st.sec, tot ? 0 : ic + 1, st.vec, vv, 1, &nrm);
            const int nlanc = (int)std::min<int64_t>(idim, in.lanc_ngfiter);
            std::vector<double> alfa(nlanc, 0.0), beta(nlanc, 0.0);
            int nused = 0;
            if (!rc && in.ed_sparse_H && idim > 1) rc = edgpu_sector_build_csr(st.sec);
            if (!rc && nrm > 0.0) rc = edgpu_lanczos_tridiag(st.sec, vv, nlanc, 1e-13, alfa.data(), beta.data(), &nused);
            if (!rc && in.ed_sparse_H && idim > 1) rc = edgpu_sector_drop_csr(st.sec);
            edgpu_vec_free(vv);
            if (rc) return fail(s, "%s", edgpu_last_error(s->wctx[st.worker]));
            // the single-orbital routine hands the NORM to add_to_lanczos_spinChi (:101), the total one its SQUARE (:206);
            // both are squared again there (pesoF = vnorm**2, :263) -- reproduced as is
            if (nrm > 0.0) add_to_lanczos_spinchi(s, tot ? nrm * nrm : nrm, st.e, alfa, beta, ic);
        }
    }
    GPU_TRY(s, allsum(s, s->spinChi_tau));
    GPU_TRY(s, allsum(s, s->spinChi_w));
    GPU_TRY(s, allsum(s, s->spinChi_iv));
    for (auto &v : s->spinChi_tau) v /= s->zeta;                                     // :36-38
    for (auto &v : s->spinChi_w) v /= s->zeta;
    for (auto &v : s->spinChi_iv) v /= s->zeta;
    return 0;
}

// add_to_lanczos_densChi / _tot (ED_GF_CHIDENS.f90:692-765, 876-948), T = 0: both signs.  iv/tau/w point at the channel's
// first element, `stride` is the distance between consecutive frequencies.
static void add_to_lanczos_denschi(ed_solver *s, double vnorm2, double Ei, const std::vector<double> &alanc,
                                   const std::vector<double> &blanc, cplx *iv, double *tau_out, cplx *w_out, size_t stride)
{
    const ed_input &in = s->in;
    const int nlanc = (int)alanc.size();
    const double beta = in.beta;
    const double pesoF = vnorm2 / s->zeta;
    std::vector<double> d(alanc), e(nlanc, 0.0), Z((size_t)nlanc * nlanc, 0.0);
    for (int i = 1; i < nlanc; i++) e[i] = blanc[i];
    for (int i = 0; i < nlanc; i++) Z[i + (size_t)nlanc * i] = 1.0;
    host_tql2(nlanc, d.data(), e.data(), Z.data());                                  // tql2, not eigh (:713-718)
    for (int isign = 1; isign >= -1; isign -= 2)
        for (int j = 0; j < nlanc; j++) {
            const double dE = d[j] - Ei;
            const double peso = pesoF * Z[(size_t)nlanc * j] * Z[(size_t)nlanc * j];
            const double ex = std::exp(-beta * dE);
            if (isign == 1) iv[0] += (beta * dE < 1e-1) ? -peso * beta : peso * (ex - 1.0) / dE;      // :727-731 (sic)
            else            iv[0] += (beta * dE < 1e-1) ? peso * beta : peso * (1.0 - ex) / dE;       // :748-752
            for (int i = 1; i <= in.Lmats; i++)
                iv[stride * i] += isign == 1 ? peso * (ex - 1.0) / (cplx(0.0, s->vm[i]) - dE) : peso * (1.0 - ex) / (cplx(0.0, s->vm[i]) + dE);
            for (int i = 0; i <= s->Ltau; i++)
                tau_out[stride * i] += isign == 1 ? peso * std::exp(-s->tau[i] * dE) : peso * std::exp(-(beta - s->tau[i]) * dE);
            for (int i = 0; i < in.Lreal; i++)
                w_out[stride * i] += isign == 1 ? peso * (ex - 1.0) / (cplx(s->wr[i], in.eps) - dE) : peso * (1.0 - ex) / (cplx(s->wr[i], in.eps) + dE);
        }
}

// build_chi_dens (ED_GF_CHIDENS.f90:21-66), the channels with real seeds: diagonal (:90-169) and total (:191-269).
static int build_chi_dens(ed_solver *s)
{
    const ed_input &in = s->in;
    const size_t nn = (size_t)in.Norb * in.Norb;
    s->densChi_iv.assign(nn * (in.Lmats + 1), cplx(0, 0));
    s->densChi_tau.assign(nn * (s->Ltau + 1), 0.0);
    s->densChi_w.assign(nn * in.Lreal, cplx(0, 0));
    s->densChi_tot_iv.assign(in.Lmats + 1, cplx(0, 0));
    s->densChi_tot_tau.assign(s->Ltau + 1, 0.0);
    s->densChi_tot_w.assign(in.Lreal, cplx(0, 0));
    if (!in.chidens_flag) return 0;
    const int nchan = in.Norb + (in.Norb > 1 ? 1 : 0);
    for (int ic = 0; ic < nchan; ic++) {
        const bool tot = ic == in.Norb;
        for (size_t istate = 0; istate < s->states.size(); istate++) {
            EdState &st = s->states[istate];
            int64_t idim = 0;
            edgpu_sector_dim(st.sec, &idim, nullptr, nullptr);
            edgpu_vec *vv = nullptr;
            GPU_TRY(s, edgpu_vec_alloc(st.sec, &vv));
            double nrm = 0;
            int rc = edgpu_apply_n(st.sec, tot ? 0 : ic + 1, st.vec, vv, 1, &nrm);
            const int nlanc = (int)std::min<int64_t>(idim, in.lanc_ngfiter);
            std::vector<double> alfa(nlanc, 0.0), beta(nlanc, 0.0);
            int nused = 0;
            if (!rc && in.ed_sparse_H && idim > 1) rc = edgpu_sector_build_csr(st.sec);
            if (!rc && nrm > 0.0) rc = edgpu_lanczos_tridiag(st.sec, vv, nlanc, 1e-13, alfa.data(), beta.data(), &nused);
            if (!rc && in.ed_sparse_H && idim > 1) rc = edgpu_sector_drop_csr(st.sec);
            edgpu_vec_free(vv);
            if (rc) return fail(s, "%s", edgpu_last_error(s->wctx[st.worker]));
            if (nrm <= 0.0) continue;
            if (tot) add_to_lanczos_denschi(s, nrm * nrm, st.e, alfa, beta, s->densChi_tot_iv.data(), s->densChi_tot_tau.data(), s->densChi_tot_w.data(), 1);
            else {
                const size_t o = (size_t)ic + (size_t)in.Norb * ic;                  // (iorb,iorb), column-major
                add_to_lanczos_denschi(s, nrm * nrm, st.e, alfa, beta, s->densChi_iv.data() + o, s->densChi_tau.data() + o, s->densChi_w.data() + o, nn);
            }
        }
    }
    GPU_TRY(s, allsum(s, s->densChi_tau)); GPU_TRY(s, allsum(s, s->densChi_w)); GPU_TRY(s, allsum(s, s->densChi_iv));
    GPU_TRY(s, allsum(s, s->densChi_tot_tau)); GPU_TRY(s, allsum(s, s->densChi_tot_w)); GPU_TRY(s, allsum(s, s->densChi_tot_iv));
    for (auto &v : s->densChi_tau) v /= s->zeta;                                     // :62-64 (the total channel is not divided again)
    for (auto &v : s->densChi_w) v /= s->zeta;
    for (auto &v : s->densChi_iv) v /= s->zeta;
    return 0;
}

static cplx delta_bath(const ed_solver *s, cplx x, int ispin, int iorb)
{
    // delta_bath_mats_main, normal/normal (ED_BATH_FUNCTIONS.f90:245-256)
    const ed_input &in = s->in;
    const int nb = in.Nspin * in.Norb * in.Nbath;
    cplx d(0, 0);
    for (int k = 0; k < in.Nbath; k++) {
        const double e = s->bath[(size_t)(ispin * in.Norb + iorb) * in.Nbath + k];
        const double v = s->bath[(size_t)nb + (size_t)(ispin * in.Norb + iorb) * in.Nbath + k];
        d += v * v / (x - e);
    }
    return d;
}

static void build_sigma(ed_solver *s)
{
    // build_sigma_normal (ED_GF_NORMAL.f90:656-694, 725-726) + invg0_bath (ED_BATH_FUNCTIONS.f90:1784-1807)
    const ed_input &in = s->in;
    const size_t nblk = (size_t)in.Nspin * in.Nspin * in.Norb * in.Norb;
    s->Smats.assign(nblk * in.Lmats, cplx(0, 0));
    s->Sreal.assign(nblk * in.Lreal, cplx(0, 0));
    s->G0mats.assign(nblk * in.Lmats, cplx(0, 0));
    s->G0real.assign(nblk * in.Lreal, cplx(0, 0));
    for (int ispin = 0; ispin < in.Nspin; ispin++)
        for (int iorb = 0; iorb < in.Norb; iorb++) {
            const size_t hidx = (size_t)ispin + in.Nspin * ((size_t)ispin + in.Nspin * ((size_t)iorb + in.Norb * (size_t)iorb));
            const cplx hl(s->hloc[2 * hidx], s->hloc[2 * hidx + 1]);
            for (int i = 0; i < in.Lmats; i++) {
                const cplx z(0.0, s->wm[i]);
                const cplx invg0 = z + in.xmu - hl - delta_bath(s, z, ispin, iorb);
                const size_t q = gidx(in, ispin, ispin, iorb, iorb, i);
                s->Smats[q] = invg0 - 1.0 / s->Gmats[q];
                s->G0mats[q] = 1.0 / invg0;
            }
            for (int i = 0; i < in.Lreal; i++) {
                const cplx z(s->wr[i], in.eps);
                const cplx invg0 = z + in.xmu - hl - delta_bath(s, z, ispin, iorb);
                const size_t q = gidx(in, ispin, ispin, iorb, iorb, i);
                s->Sreal[q] = invg0 - 1.0 / s->Greal[q];
                s->G0real[q] = 1.0 / invg0;
            }
        }
}

static int observables(ed_solver *s)
{
    // observables_impurity (ED_OBSERVABLES.f90:105-162), T=0: peso = 1/zeta_function
    const int n = s->in.Norb;
    s->dens.assign(n, 0.0); s->dens_up.assign(n, 0.0); s->dens_dw.assign(n, 0.0); s->docc.assign(n, 0.0); s->magz.assign(n, 0.0);
    s->sz2.assign((size_t)n * n, 0.0); s->n2.assign((size_t)n * n, 0.0);
    s->s2tot = 0.0;
    for (auto &st : s->states)
        GPU_TRY(s, edgpu_observables(st.sec, st.vec, 1.0 / s->zeta, s->dens.data(), s->dens_up.data(), s->dens_dw.data(),
                                     s->docc.data(), s->magz.data(), s->sz2.data(), s->n2.data(), &s->s2tot));
    if (s->nranks > 1) {
        GPU_TRY(s, allsum(s, s->dens)); GPU_TRY(s, allsum(s, s->dens_up)); GPU_TRY(s, allsum(s, s->dens_dw));
        GPU_TRY(s, allsum(s, s->docc)); GPU_TRY(s, allsum(s, s->magz)); GPU_TRY(s, allsum(s, s->sz2)); GPU_TRY(s, allsum(s, s->n2));
        GPU_TRY(s, allsum(s, &s->s2tot, 1));
    }
    return 0;
}

extern "C" int ed_solve(ed_solver *s, const double *bath, int32_t bath_len, const double *hloc_cplx)
{
    if (!s || !bath) return 1;
    const ed_input &in = s->in;
    if (bath_len != ed_get_bath_dimension(&in)) return fail(s, "ED_SOLVE_SINGLE Error: wrong bath dimensions");   // ED_MAIN.f90:258
    if (hloc_cplx) memcpy(s->hloc.data(), hloc_cplx, sizeof(double) * s->hloc.size());                           // set_Hloc :256
    s->bath.assign(bath, bath + bath_len);
    for (edgpu_ctx *c : s->wctx)
        if (edgpu_set_hamiltonian(c, bath, bath_len, s->hloc.data(), in.uloc, in.ust, in.jh, in.jx, in.jp, in.xmu)) return fail(s, "%s", edgpu_last_error(c));
    double t0 = now_s();
    if (int rc = ed_diag(s)) return rc;                     // diagonalize_impurity
    double t1 = now_s();
    if (int rc = build_gf(s)) return rc;                    // buildgf_impurity
    double t2 = now_s();
    build_sigma(s);
    if (int rc = build_chi_spin(s)) return rc;              // buildchi_impurity (ED_MAIN.f90:274)
    if (int rc = build_chi_dens(s)) return rc;
    double t3 = now_s();
    if (int rc = observables(s)) return rc;                 // observables_impurity
    double t4 = now_s();
    s->timings[0] = t1 - t0; s->timings[1] = t2 - t1; s->timings[2] = t3 - t2; s->timings[3] = t4 - t3;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// getters (ED_IO)
// ---------------------------------------------------------------------------------------------------------
static int copy_c(const std::vector<cplx> &src, double *dst)
{
    if (!dst || src.empty()) return 1;
    memcpy(dst, src.data(), sizeof(cplx) * src.size());
    return 0;
}
static int copy_d(const std::vector<double> &src, double *dst)
{
    if (!dst || src.empty()) return 1;
    memcpy(dst, src.data(), sizeof(double) * src.size());
    return 0;
}
extern "C" int ed_get_sigma_matsubara(const ed_solver *s, double *o) { return s ? copy_c(s->Smats, o) : 1; }
extern "C" int ed_get_sigma_real(const ed_solver *s, double *o) { return s ? copy_c(s->Sreal, o) : 1; }
extern "C" int ed_get_gimp_matsubara(const ed_solver *s, double *o) { return s ? copy_c(s->Gmats, o) : 1; }
extern "C" int ed_get_gimp_real(const ed_solver *s, double *o) { return s ? copy_c(s->Greal, o) : 1; }
extern "C" int ed_get_g0imp_matsubara(const ed_solver *s, double *o) { return s ? copy_c(s->G0mats, o) : 1; }
extern "C" int ed_get_g0imp_real(const ed_solver *s, double *o) { return s ? copy_c(s->G0real, o) : 1; }
extern "C" int ed_get_dens(const ed_solver *s, double *o) { return s ? copy_d(s->dens, o) : 1; }
extern "C" int ed_get_dens_up(const ed_solver *s, double *o) { return s ? copy_d(s->dens_up, o) : 1; }
extern "C" int ed_get_dens_dw(const ed_solver *s, double *o) { return s ? copy_d(s->dens_dw, o) : 1; }
extern "C" int ed_get_docc(const ed_solver *s, double *o) { return s ? copy_d(s->docc, o) : 1; }
extern "C" int ed_get_mag(const ed_solver *s, double *o) { return s ? copy_d(s->magz, o) : 1; }
extern "C" int ed_get_sz2_n2(const ed_solver *s, double *sz2, double *n2, double *s2tot)
{
    if (!s) return 1;
    if (sz2) copy_d(s->sz2, sz2);
    if (n2) copy_d(s->n2, n2);
    if (s2tot) *s2tot = s->s2tot;
    return 0;
}
extern "C" int ed_get_grids(const ed_solver *s, double *wm, double *wr)
{
    if (!s) return 1;
    if (wm) copy_d(s->wm, wm);
    if (wr) copy_d(s->wr, wr);
    return 0;
}
extern "C" int ed_get_spinchi(const ed_solver *s, double *chi_iv, double *chi_tau, double *chi_w, double *vm, double *tau, int32_t *ltau)
{
    if (!s || s->spinChi_iv.empty()) return 1;
    if (chi_iv) copy_c(s->spinChi_iv, chi_iv);
    if (chi_tau) copy_d(s->spinChi_tau, chi_tau);
    if (chi_w) copy_c(s->spinChi_w, chi_w);
    if (vm) copy_d(s->vm, vm);
    if (tau) copy_d(s->tau, tau);
    if (ltau) *ltau = s->Ltau;
    return 0;
}
extern "C" int ed_get_denschi(const ed_solver *s, double *chi_iv, double *chi_tau, double *chi_w, double *tot_iv, double *tot_tau, double *tot_w)
{
    if (!s || s->densChi_iv.empty()) return 1;
    if (chi_iv) copy_c(s->densChi_iv, chi_iv);
    if (chi_tau) copy_d(s->densChi_tau, chi_tau);
    if (chi_w) copy_c(s->densChi_w, chi_w);
    if (tot_iv) copy_c(s->densChi_tot_iv, tot_iv);
    if (tot_tau) copy_d(s->densChi_tot_tau, tot_tau);
    if (tot_w) copy_c(s->densChi_tot_w, tot_w);
    return 0;
}
extern "C" int ed_get_state_count(const ed_solver *s, int32_t *n, double *zeta, double *egs)
{
    if (!s) return 1;
    if (n) *n = (int32_t)s->states.size();
    if (zeta) *zeta = s->zeta;
    if (egs) *egs = s->egs;
    return 0;
}
extern "C" int ed_get_state(const ed_solver *s, int32_t i, double *e, int32_t *nup, int32_t *ndw)
{
    if (!s || i < 0 || i >= (int)s->states.size()) return 1;
    if (e) *e = s->states[i].e;
    if (nup) *nup = s->states[i].nup;
    if (ndw) *ndw = s->states[i].ndw;
    return 0;
}
extern "C" int ed_get_state_vector(const ed_solver *s, int32_t i, double *vec, int64_t len)
{
    if (!s || i < 0 || i >= (int)s->states.size() || !vec) return 1;
    int64_t dim = 0;
    edgpu_sector_dim(s->states[i].sec, &dim, nullptr, nullptr);
    if (len != dim) return 1;
    return edgpu_vec_download(s->states[i].vec, vec, 0);
}
extern "C" int ed_get_sector_energy(const ed_solver *s, int32_t nup, int32_t ndw, double *e)
{
    if (!s || !e) return 1;
    auto it = s->sector_e.find({nup, ndw});
    if (it == s->sector_e.end()) return 1;
    *e = it->second;
    return 0;
}
extern "C" int ed_get_sector_nlanc(const ed_solver *s, int32_t nup, int32_t ndw, int32_t *nlanc)
{
    if (!s || !nlanc) return 1;
    auto it = s->sector_nlanc.find({nup, ndw});
    *nlanc = it == s->sector_nlanc.end() ? 0 : it->second;
    return 0;
}
extern "C" int ed_get_chain_count(const ed_solver *s, int32_t *n)
{
    if (!s || !n) return 1;
    *n = (int32_t)s->chains.size();
    return 0;
}
extern "C" int ed_get_chain(const ed_solver *s, int32_t i, int32_t *iorb, int32_t *ispin, int32_t *isign, int32_t *istate,
                            int32_t *nlanc, int32_t *nused, double *norm2, double *alfa, double *beta, int32_t cap)
{
    if (!s || i < 0 || i >= (int)s->chains.size()) return 1;
    const EdChain &c = s->chains[i];
    if (iorb) *iorb = c.iorb;
    if (ispin) *ispin = c.ispin;
    if (isign) *isign = c.isign;
    if (istate) *istate = c.istate;
    if (nlanc) *nlanc = c.nlanc;
    if (nused) *nused = c.nused;
    if (norm2) *norm2 = c.norm2;
    for (int k = 0; k < c.nlanc && k < cap; k++) {
        if (alfa) alfa[k] = c.alfa[k];
        if (beta) beta[k] = c.beta[k];
    }
    return 0;
}
extern "C" int ed_get_timings(const ed_solver *s, double *t4)
{
    if (!s || !t4) return 1;
    for (int i = 0; i < 4; i++) t4[i] = s->timings[i];
    return 0;
}
