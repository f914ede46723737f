/*
 * ed_oracle.c -- CPU ORACLE (test infrastructure, NOT product code) for the dmft-ed Lanczos hot path.
 *
 * PARITY UNPINNED (see ed_oracle.h): literal restatement of the reference rules, cross-checked only by
 * independent invariants.  Citations are file:line relative to the reference root.
 *
 * Build: make -C oracle      (gcc -O3 -funroll-loops = the reference's RELEASE flags, CMakeLists.txt:46)
 */
#include "ed_oracle.h"
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

/* ------------------------------------------------------------------ model ---------------------------------- */

ora_model *ora_model_new(int Norb, int Nbath, int Nspin, int hfmode,
                         const double *uloc, double ust, double jh, double jx, double jp, double xmu,
                         const double *bath, const double *hloc_re, const double *hloc_im)
{
    ora_model *m = (ora_model *)calloc(1, sizeof(ora_model));
    m->Norb = Norb; m->Nbath = Nbath; m->Nspin = Nspin;
    m->Ns = (Nbath + 1) * Norb;                      /* ED_SETUP.f90:99-101 */
    m->hfmode = hfmode;
    for (int i = 0; i < Norb && i < ORA_MAXORB; i++) m->uloc[i] = uloc[i];
    m->ust = ust; m->jh = jh; m->jx = jx; m->jp = jp; m->xmu = xmu;
    m->jhflag = (Norb > 1) && (jx != 0.0 || jp != 0.0);   /* ED_SETUP.f90:289-290 */
    int nb = Nspin * Norb * Nbath;
    m->e = (double *)malloc(sizeof(double) * nb);
    m->v = (double *)malloc(sizeof(double) * nb);
    /* set_dmft_bath, dmft_aux.f90:494-511: io = i + (iorb-1)*Nbath + (ispin-1)*Nbath*Norb, e first then v */
    memcpy(m->e, bath, sizeof(double) * nb);
    memcpy(m->v, bath + nb, sizeof(double) * nb);
    int nh = Nspin * Nspin * Norb * Norb;
    m->hloc_re = (double *)calloc(nh, sizeof(double));
    m->hloc_im = (double *)calloc(nh, sizeof(double));
    if (hloc_re) memcpy(m->hloc_re, hloc_re, sizeof(double) * nh);
    if (hloc_im) memcpy(m->hloc_im, hloc_im, sizeof(double) * nh);
    return m;
}

void ora_model_free(ora_model *m)
{
    if (!m) return;
    free(m->e); free(m->v); free(m->hloc_re); free(m->hloc_im); free(m);
}

/* impHloc(ispin,jspin,iorb,jorb), 1-based Fortran indices, column-major (ED_VARS_GLOBAL.f90, set_Hloc) */
static inline cplx hloc(const ora_model *m, int ispin, int jspin, int iorb, int jorb)
{
    int idx = (ispin - 1) + m->Nspin * ((jspin - 1) + m->Nspin * ((iorb - 1) + m->Norb * (jorb - 1)));
    return m->hloc_re[idx] + I * m->hloc_im[idx];
}
static inline double bath_e(const ora_model *m, int ispin, int iorb, int k)
{ return m->e[((ispin - 1) * m->Norb + (iorb - 1)) * m->Nbath + (k - 1)]; }
static inline double bath_v(const ora_model *m, int ispin, int iorb, int k)
{ return m->v[((ispin - 1) * m->Norb + (iorb - 1)) * m->Nbath + (k - 1)]; }
/* getBathStride(iorb,i) = Norb + (iorb-1)*Nbath + i   (ED_SETUP.f90:450-454, bath_type=normal) */
static inline int bath_stride(const ora_model *m, int iorb, int k)
{ return m->Norb + (iorb - 1) * m->Nbath + k; }

void ora_init_bath(int Norb, int Nbath, int Nspin, double hwband, double *bath)
{
    /* init_dmft_bath, ED_BATH/dmft_aux.f90:105-127, noise = 0; same values for every (ispin,iorb) */
    double *ek = (double *)calloc(Nbath + 2, sizeof(double));   /* 1-based */
    double *vk = (double *)calloc(Nbath + 2, sizeof(double));
    ek[1] = -hwband;
    ek[Nbath] = hwband;
    int Nh = Nbath / 2;
    if (Nbath % 2 == 0 && Nbath >= 4) {
        double de = hwband / (double)((Nh - 1) > 1 ? (Nh - 1) : 1);
        ek[Nh] = -1.e-3;
        ek[Nh + 1] = 1.e-3;
        for (int i = 2; i <= Nh - 1; i++) {
            ek[i] = -hwband + (i - 1) * de;
            ek[Nbath - i + 1] = hwband - (i - 1) * de;
        }
    } else if (Nbath % 2 != 0 && Nbath >= 3) {
        double de = hwband / (double)Nh;
        ek[Nh + 1] = 0.0;
        for (int i = 2; i <= Nh; i++) {
            ek[i] = -hwband + (i - 1) * de;
            ek[Nbath - i + 1] = hwband - (i - 1) * de;
        }
    }
    for (int i = 1; i <= Nbath; i++) {
        double a = 1.0 / sqrt((double)Nbath);
        vk[i] = a > 0.1 ? a : 0.1;
    }
    int nb = Nspin * Norb * Nbath;
    for (int is = 0; is < Nspin; is++)
        for (int io = 0; io < Norb; io++)
            for (int k = 1; k <= Nbath; k++) {
                bath[(is * Norb + io) * Nbath + (k - 1)] = ek[k];
                bath[nb + (is * Norb + io) * Nbath + (k - 1)] = vk[k];
            }
    free(ek); free(vk);
}

/* ------------------------------------------------------------------ sectors -------------------------------- */

int64_t ora_binomial(int n1, int n2)
{
    /* ED_SETUP.f90:1283-1300: floating product, rounded */
    double xh = 1.0;
    if (n2 < 0) return 0;
    if (n2 == 0) return 1;
    for (int i = 1; i <= n2; i++) xh = xh * (double)(n1 + 1 - i) / (double)i;
    return (int64_t)(xh + 0.5);
}

int64_t ora_sector_dim(int Ns, int nup, int ndw)
{
    return ora_binomial(Ns, nup) * ora_binomial(Ns, ndw);     /* ED_SETUP.f90:818-830 */
}

static inline int bdecomp_sum(uint64_t i, int ntot)
{
    /* bdecomp (ED_SETUP.f90:1234-1244) followed by sum(ivec) as in build_sector */
    int s = 0;
    for (int l = 0; l < ntot; l++) if ((i >> l) & 1ull) s++;
    return s;
}

int64_t ora_build_sector(int Ns, int nup, int ndw, uint64_t *map, int literal)
{
    /* ED_SETUP.f90:899-916 (normal branch), idw outer, iup inner; map = iup + idw*2**Ns in 64-bit */
    int64_t dim = 0;
    uint64_t n = 1ull << Ns;
    if (literal) {
        for (uint64_t idw = 0; idw < n; idw++) {
            if (bdecomp_sum(idw, Ns) != ndw) continue;
            for (uint64_t iup = 0; iup < n; iup++) {
                if (bdecomp_sum(iup, Ns) != nup) continue;
                if (map) map[dim] = iup + idw * n;
                dim++;
            }
        }
        return dim;
    }
    /* same order, skipping words of the wrong popcount up front */
    int64_t dup = ora_binomial(Ns, nup);
    uint64_t *ups = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(dup > 0 ? dup : 1));
    int64_t k = 0;
    for (uint64_t iup = 0; iup < n; iup++)
        if (__builtin_popcountll(iup) == nup) ups[k++] = iup;
    for (uint64_t idw = 0; idw < n; idw++) {
        if (__builtin_popcountll(idw) != ndw) continue;
        if (map) for (int64_t j = 0; j < dup; j++) map[dim + j] = ups[j] + idw * n;
        dim += dup;
    }
    free(ups);
    return dim;
}

int ora_c(int pos, uint64_t in, uint64_t *out, double *sgn)
{
    /* ED_SETUP.f90:1080-1092 */
    if (!((in >> (pos - 1)) & 1ull)) return 0;             /* stop "C error" */
    double f = 1.0;
    for (int l = 1; l <= pos - 1; l++) if ((in >> (l - 1)) & 1ull) f = -f;
    *sgn = f;
    *out = in & ~(1ull << (pos - 1));
    return 1;
}

int ora_cdg(int pos, uint64_t in, uint64_t *out, double *sgn)
{
    /* ED_SETUP.f90:1094-1106 */
    if ((in >> (pos - 1)) & 1ull) return 0;                /* stop "C^+ error" */
    double f = 1.0;
    for (int l = 1; l <= pos - 1; l++) if ((in >> (l - 1)) & 1ull) f = -f;
    *sgn = f;
    *out = in | (1ull << (pos - 1));
    return 1;
}

int64_t ora_binary_search(const uint64_t *a, int64_t n, uint64_t value)
{
    /* ED_SETUP.f90:1307-1324, recursive, 1-based result, 0 = not found */
    if (n == 0) return 0;
    int64_t mid = n / 2 + 1;
    if (a[mid - 1] > value) return ora_binary_search(a, mid - 1, value);
    if (a[mid - 1] < value) {
        int64_t r = ora_binary_search(a + mid, n - mid, value);
        return r != 0 ? mid + r : 0;
    }
    return mid;
}

/* ------------------------------------------------------------------ Hamiltonian terms ----------------------- */

/* Diagonal pieces, in the order the reference adds them (three separate "Hv(j) += htmp*vin(j)"). */
static void diag_terms(const ora_model *M, uint64_t m, cplx h[3])
{
    const int Ns = M->Ns, Norb = M->Norb, Nspin = M->Nspin;
    double nup[ORA_MAXORB], ndw[ORA_MAXORB];
    for (int io = 1; io <= Norb; io++) {
        nup[io - 1] = (double)((m >> (io - 1)) & 1ull);
        ndw[io - 1] = (double)((m >> (io - 1 + Ns)) & 1ull);
    }
    /* direct/HxVimp.f90:2-8 */
    cplx htmp = 0.0;
    double snup = 0, sndw = 0;
    for (int io = 0; io < Norb; io++) { snup += nup[io]; sndw += ndw[io]; }
    htmp = htmp - M->xmu * (snup + sndw);
    for (int io = 1; io <= Norb; io++) {
        htmp = htmp + hloc(M, 1, 1, io, io) * nup[io - 1];
        htmp = htmp + hloc(M, Nspin, Nspin, io, io) * ndw[io - 1];
    }
    h[0] = htmp;
    /* direct/HxVint.f90:3-39 */
    htmp = 0.0;
    for (int io = 0; io < Norb; io++) htmp = htmp + M->uloc[io] * nup[io] * ndw[io];
    if (Norb > 1) {
        for (int io = 0; io < Norb; io++)
            for (int jo = io + 1; jo < Norb; jo++)
                htmp = htmp + M->ust * (nup[io] * ndw[jo] + nup[jo] * ndw[io]);
        for (int io = 0; io < Norb; io++)
            for (int jo = io + 1; jo < Norb; jo++)
                htmp = htmp + (M->ust - M->jh) * (nup[io] * nup[jo] + ndw[io] * ndw[jo]);
    }
    if (M->hfmode) {
        for (int io = 0; io < Norb; io++)
            htmp = htmp - 0.5 * M->uloc[io] * (nup[io] + ndw[io]) + 0.25 * M->uloc[io];
        if (Norb > 1) {
            for (int io = 0; io < Norb; io++)
                for (int jo = io + 1; jo < Norb; jo++) {
                    htmp = htmp - 0.5 * M->ust * (nup[io] + ndw[io] + nup[jo] + ndw[jo]) + 0.25 * M->ust;
                    htmp = htmp - 0.5 * (M->ust - M->jh) * (nup[io] + ndw[io] + nup[jo] + ndw[jo]) + 0.25 * (M->ust - M->jh);
                }
        }
    }
    h[1] = htmp;
    /* direct/HxVbath.f90:4-11 */
    htmp = 0.0;
    for (int io = 1; io <= Norb; io++)
        for (int kp = 1; kp <= M->Nbath; kp++) {
            int alfa = bath_stride(M, io, kp);
            htmp = htmp + bath_e(M, 1, io, kp) * (double)((m >> (alfa - 1)) & 1ull);
            htmp = htmp + bath_e(M, Nspin, io, kp) * (double)((m >> (alfa - 1 + Ns)) & 1ull);
        }
    h[2] = htmp;
}

/* One off-diagonal term produced from state m: target word k and amplitude (direct form: Hv(k) += amp*v(m);
 * stored form inserts conjg(amp) at (row m, column k)). */
typedef struct { uint64_t k; cplx amp; } ora_term;
#define ORA_MAXTERMS 512

#define IB(pos) ((int)((m >> ((pos) - 1)) & 1ull))

/* Off-diagonal terms generated from word m, split by include file so callers can interleave the diagonal
 * accumulations exactly like the reference: part 0 = HxVimp.f90:16-50, part 1 = HxVint.f90:46-98,
 * part 2 = HxVimp_bath.f90:1-38.  (stored/Himp.f90:26-70, stored/Hint.f90:60-118, stored/Himp_bath.f90.) */
static int offdiag_terms(const ora_model *M, uint64_t m, int part, ora_term *t)
{
    const int Ns = M->Ns, Norb = M->Norb, Nspin = M->Nspin, Nbath = M->Nbath;
    int n = 0;
    uint64_t k1, k2, k3, k4;
    double sg1, sg2, sg3, sg4;
    if (part == 0) {
        for (int io = 1; io <= Norb; io++)
            for (int jo = 1; jo <= Norb; jo++) {
                cplx hu = hloc(M, 1, 1, io, jo);
                if (hu != 0.0 && IB(jo) == 1 && IB(io) == 0) {
                    ora_c(jo, m, &k1, &sg1);
                    ora_cdg(io, k1, &k2, &sg2);
                    t[n].k = k2; t[n].amp = hu * sg1 * sg2; n++;
                }
                cplx hd = hloc(M, Nspin, Nspin, io, jo);
                if (hd != 0.0 && IB(jo + Ns) == 1 && IB(io + Ns) == 0) {
                    ora_c(jo + Ns, m, &k1, &sg1);
                    ora_cdg(io + Ns, k1, &k2, &sg2);
                    t[n].k = k2; t[n].amp = hd * sg1 * sg2; n++;
                }
            }
    } else if (part == 1) {
        if (Norb > 1 && M->jhflag) {
            for (int io = 1; io <= Norb; io++)            /* spin-exchange, HxVint.f90:49-71 */
                for (int jo = 1; jo <= Norb; jo++)
                    if (io != jo && IB(jo) == 1 && IB(io + Ns) == 1 && IB(jo + Ns) == 0 && IB(io) == 0) {
                        ora_c(jo, m, &k1, &sg1);
                        ora_c(io + Ns, k1, &k2, &sg2);
                        ora_cdg(jo + Ns, k2, &k3, &sg3);
                        ora_cdg(io, k3, &k4, &sg4);
                        t[n].k = k4; t[n].amp = M->jx * sg1 * sg2 * sg3 * sg4; n++;
                    }
            for (int io = 1; io <= Norb; io++)            /* pair-hopping, HxVint.f90:76-98 */
                for (int jo = 1; jo <= Norb; jo++)
                    if (io != jo && IB(jo) == 1 && IB(jo + Ns) == 1 && IB(io + Ns) == 0 && IB(io) == 0) {
                        ora_c(jo, m, &k1, &sg1);
                        ora_c(jo + Ns, k1, &k2, &sg2);
                        ora_cdg(io + Ns, k2, &k3, &sg3);
                        ora_cdg(io, k3, &k4, &sg4);
                        t[n].k = k4; t[n].amp = M->jp * sg1 * sg2 * sg3 * sg4; n++;
                    }
        }
    } else {
        for (int io = 1; io <= Norb; io++)
            for (int kp = 1; kp <= Nbath; kp++) {
                int ms = bath_stride(M, io, kp);
                double vu = bath_v(M, 1, io, kp), vd = bath_v(M, Nspin, io, kp);
                if (vu != 0.0 && IB(io) == 1 && IB(ms) == 0) {
                    ora_c(io, m, &k1, &sg1); ora_cdg(ms, k1, &k2, &sg2);
                    t[n].k = k2; t[n].amp = vu * sg1 * sg2; n++;
                }
                if (vu != 0.0 && IB(io) == 0 && IB(ms) == 1) {
                    ora_c(ms, m, &k1, &sg1); ora_cdg(io, k1, &k2, &sg2);
                    t[n].k = k2; t[n].amp = vu * sg1 * sg2; n++;
                }
                if (vd != 0.0 && IB(io + Ns) == 1 && IB(ms + Ns) == 0) {
                    ora_c(io + Ns, m, &k1, &sg1); ora_cdg(ms + Ns, k1, &k2, &sg2);
                    t[n].k = k2; t[n].amp = vd * sg1 * sg2; n++;
                }
                if (vd != 0.0 && IB(io + Ns) == 0 && IB(ms + Ns) == 1) {
                    ora_c(ms + Ns, m, &k1, &sg1); ora_cdg(io + Ns, k1, &k2, &sg2);
                    t[n].k = k2; t[n].amp = vd * sg1 * sg2; n++;
                }
            }
    }
    return n;
}

void ora_direct_hxv(const ora_model *M, const uint64_t *map, int64_t dim,
                    const double *vin_, double *hv_, int64_t j0, int64_t j1)
{
    /* ED_HAMILTONIAN_DIRECT_HxV.f90:68-89: scatter form, one state j at a time */
    const cplx *vin = (const cplx *)vin_;
    cplx *hv = (cplx *)hv_;
    ora_term t[ORA_MAXTERMS];
    for (int64_t j = j0; j < j1; j++) {
        uint64_t m = map[j];
        cplx h[3];
        diag_terms(M, m, h);
        for (int part = 0; part < 3; part++) {
            /* include order: HxVimp (diag, offdiag) ; HxVint (diag, Jx/Jp) ; HxVbath (diag) ; HxVimp_bath */
            hv[j] = hv[j] + h[part] * vin[j];
            int n = offdiag_terms(M, m, part, t);
            for (int q = 0; q < n; q++) {
                int64_t i = ora_binary_search(map, dim, t[q].k);
                if (i != 0) hv[i - 1] = hv[i - 1] + t[q].amp * vin[j];
            }
        }
    }
}

void ora_gather_hxv(const ora_model *M, const uint64_t *map, int64_t dim,
                    const double *vin_, double *hv_, int64_t i0, int64_t i1)
{
    /* Row form of the same operator = what stored/<part>.f90 inserts: H(i,j) = conjg(amp) for j = target(i). */
    const cplx *vin = (const cplx *)vin_;
    cplx *hv = (cplx *)hv_;
    ora_term t[ORA_MAXTERMS];
    for (int64_t i = i0; i < i1; i++) {
        uint64_t m = map[i];
        cplx h[3];
        diag_terms(M, m, h);
        cplx acc = 0.0;
        for (int part = 0; part < 3; part++) {
            acc = acc + h[part] * vin[i];
            int n = offdiag_terms(M, m, part, t);
            for (int q = 0; q < n; q++) {
                int64_t j = ora_binary_search(map, dim, t[q].k);
                if (j != 0) acc = acc + conj(t[q].amp) * vin[j - 1];
            }
        }
        hv[i] = acc;
    }
}

int64_t ora_stored_build(const ora_model *M, const uint64_t *map, int64_t dim,
                         int64_t *rowptr, int64_t *cols, double *vals_)
{
    /* ED_HAMILTONIAN_STORED_HxV.f90:28-113: sp_insert_element accumulates when the column already exists in
     * the row (ED_SPARSE_MATRIX.f90:263-269), else appends (:270-276).  Per-row insertion order:
     * Himp diag, Himp offdiag, Hint diag, Jx/Jp, Hbath diag, Himp_bath. */
    cplx *vals = (cplx *)vals_;
    ora_term t[ORA_MAXTERMS];
    int64_t rc[ORA_MAXTERMS + 1];
    cplx rv[ORA_MAXTERMS + 1];
    int64_t nnz = 0;
    for (int64_t i = 0; i < dim; i++) {
        uint64_t m = map[i];
        cplx h[3];
        diag_terms(M, m, h);
        /* stored/Himp.f90:11-16 adds the three pieces per orbital in a different order than direct/HxVimp.f90;
         * the sum is the same up to rounding.  Use the stored order here. */
        {
            cplx htmp = 0.0;
            for (int io = 1; io <= M->Norb; io++) {
                double nu = (double)((m >> (io - 1)) & 1ull), nd = (double)((m >> (io - 1 + M->Ns)) & 1ull);
                htmp = htmp + hloc(M, 1, 1, io, io) * nu;
                htmp = htmp + hloc(M, M->Nspin, M->Nspin, io, io) * nd;
                htmp = htmp - M->xmu * (nu + nd);
            }
            h[0] = htmp;
        }
        int nr = 0;
        for (int part = 0; part < 3; part++) {
            /* diagonal insert at (i,i) */
            int found = -1;
            for (int q = 0; q < nr; q++) if (rc[q] == i) { found = q; break; }
            if (found >= 0) rv[found] += h[part]; else { rc[nr] = i; rv[nr] = h[part]; nr++; }
            int n = offdiag_terms(M, m, part, t);
            for (int q = 0; q < n; q++) {
                int64_t j = ora_binary_search(map, dim, t[q].k);
                if (j == 0) continue;
                j -= 1;
                cplx val = conj(t[q].amp);
                found = -1;
                for (int p = 0; p < nr; p++) if (rc[p] == j) { found = p; break; }
                if (found >= 0) rv[found] += val; else { rc[nr] = j; rv[nr] = val; nr++; }
            }
        }
        if (rowptr) rowptr[i] = nnz;
        if (cols && vals) for (int q = 0; q < nr; q++) { cols[nnz + q] = rc[q]; vals[nnz + q] = rv[q]; }
        nnz += nr;
    }
    if (rowptr) rowptr[dim] = nnz;
    return nnz;
}

void ora_stored_hxv(int64_t dim, const int64_t *rowptr, const int64_t *cols, const double *vals_,
                    const double *vin_, double *hv_)
{
    /* ED_HAMILTONIAN_STORED_HxV.f90:132-143 */
    const cplx *vals = (const cplx *)vals_, *vin = (const cplx *)vin_;
    cplx *hv = (cplx *)hv_;
    for (int64_t i = 0; i < dim; i++) {
        cplx acc = 0.0;
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; p++) acc = acc + vals[p] * vin[cols[p]];
        hv[i] = acc;
    }
}

/* ------------------------------------------------------------------ tql2 ------------------------------------ */

static double pythag(double a, double b)
{
    /* .repo/PLAIN_LANCZOS.f90:567-605 (EISPACK pythag) */
    double p = fmax(fabs(a), fabs(b));
    if (p == 0.0) return 0.0;
    double r = fmin(fabs(a), fabs(b)) / p; r = r * r;
    for (;;) {
        double t = 4.0 + r;
        if (t == 4.0) break;
        double s = r / t, u = 1.0 + 2.0 * s;
        p = u * p; r = (s / u) * (s / u) * r;
    }
    return p;
}

int ora_tql2(int n, double *d, double *e, double *z)
{
    /* EISPACK tql2 as carried in .repo/PLAIN_LANCZOS.f90:427-565.  d(1:n) diagonal, e(2:n) sub-diagonal
     * (e[0] unused on input), z(n,n) column-major, identity on input for a tridiagonal matrix.
     * On return d ascending, z columns = eigenvectors. */
    int ierr = 0;
    if (n == 1) return 0;
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    double f = 0.0, tst1 = 0.0;
    e[n - 1] = 0.0;
    for (int l = 0; l < n; l++) {
        int j = 0;
        double h = fabs(d[l]) + fabs(e[l]);
        if (tst1 < h) tst1 = h;
        int m;
        for (m = l; m < n; m++) {
            double tst2 = tst1 + fabs(e[m]);
            if (tst2 == tst1) break;
        }
        if (m != l) {
            for (;;) {
                if (j == 30) { ierr = l + 1; return ierr; }
                j++;
                int l1 = l + 1, l2 = l1 + 1;
                double g = d[l];
                double p = (d[l1] - g) / (2.0 * e[l]);
                double r = pythag(p, 1.0);
                d[l] = e[l] / (p + copysign(r, p));
                d[l1] = e[l] * (p + copysign(r, p));
                double dl1 = d[l1];
                h = g - d[l];
                for (int i = l2; i < n; i++) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, el1 = e[l1], s = 0.0, c3 = c, s2 = 0.0;
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = pythag(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; k++) {
                        h = z[k + (size_t)n * (i + 1)];
                        z[k + (size_t)n * (i + 1)] = s * z[k + (size_t)n * i] + c * h;
                        z[k + (size_t)n * i] = c * z[k + (size_t)n * i] - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
                double tst2 = tst1 + fabs(e[l]);
                if (!(tst2 > tst1)) break;
            }
        }
        d[l] += f;
    }
    /* order eigenvalues and eigenvectors */
    for (int ii = 1; ii < n; ii++) {
        int i = ii - 1, k = i;
        double p = d[i];
        for (int j = ii; j < n; j++) if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i]; d[i] = p;
            for (int j = 0; j < n; j++) {
                double tmp = z[j + (size_t)n * i];
                z[j + (size_t)n * i] = z[j + (size_t)n * k];
                z[j + (size_t)n * k] = tmp;
            }
        }
    }
    return ierr;
}

/* ------------------------------------------------------------------ Lanczos --------------------------------- */

static double cdotr(int64_t n, const cplx *a, const cplx *b)
{
    /* Fortran dot_product(a,b) = sum(conjg(a)*b); callers only use it where the result is real */
    cplx s = 0.0;
    for (int64_t i = 0; i < n; i++) s += conj(a[i]) * b[i];
    return creal(s);
}

static int lanc_iteration(const ora_model *M, const uint64_t *map, int64_t dim, int iter,
                          cplx *vin, cplx *vout, cplx *tmp, double *a, double *b)
{
    /* .repo/PLAIN_LANCZOS.f90:87-118 (complex version) */
    if (iter == 1) {
        double norm = sqrt(cdotr(dim, vin, vin));
        if (norm == 0.0) return 0;                                  /* stop "norm =0!!" */
        for (int64_t i = 0; i < dim; i++) vin[i] = vin[i] / norm;
        *b = 0.0;
    }
    memset(tmp, 0, sizeof(cplx) * (size_t)dim);
    ora_direct_hxv(M, map, dim, (const double *)vin, (double *)tmp, 0, dim);
    for (int64_t i = 0; i < dim; i++) tmp[i] = tmp[i] - (*b) * vout[i];
    *a = cdotr(dim, vin, tmp);
    for (int64_t i = 0; i < dim; i++) tmp[i] = tmp[i] - (*a) * vin[i];
    *b = sqrt(cdotr(dim, tmp, tmp));
    for (int64_t i = 0; i < dim; i++) { vout[i] = vin[i]; vin[i] = tmp[i] / (*b); }
    return 1;
}

int ora_lanc_tridiag(const ora_model *M, const uint64_t *map, int64_t dim, double *vin_,
                     int nitermax, double threshold, double *alanc, double *blanc)
{
    /* .repo/PLAIN_LANCZOS.f90:154-180; alanc(1:n), blanc(2:n) 1-based -> alanc[0..n-1], blanc[1..n-1] here.
     * Arrays are zero-filled first (the call sites leave them uninitialised). */
    cplx *vin = (cplx *)vin_;
    cplx *vout = (cplx *)calloc((size_t)dim, sizeof(cplx));
    cplx *tmp = (cplx *)malloc(sizeof(cplx) * (size_t)dim);
    double a = 0.0, b = 0.0;
    int done = 0;
    for (int i = 0; i < nitermax; i++) { alanc[i] = 0.0; blanc[i] = 0.0; }
    for (int iter = 1; iter <= nitermax; iter++) {
        if (!lanc_iteration(M, map, dim, iter, vin, vout, tmp, &a, &b)) break;
        alanc[iter - 1] = a;
        if (iter < nitermax) blanc[iter] = b;
        done = iter;
        if (fabs(b) < threshold) break;
    }
    free(vout); free(tmp);
    return done;
}

int ora_lanc_gs(const ora_model *M, const uint64_t *map, int64_t dim, double *vect_,
                int nitermax, double threshold, int ncheck, double *egs, double *alanc, double *blanc)
{
    /* .repo/PLAIN_LANCZOS.f90:286-385.  alanc/blanc sized nitermax+1. The start vector must be supplied
     * (the reference draws random numbers when it is zero, :310-318 -- not reproducible). */
    cplx *vect = (cplx *)vect_;
    cplx *vin = (cplx *)malloc(sizeof(cplx) * (size_t)dim);
    cplx *vout = (cplx *)calloc((size_t)dim, sizeof(cplx));
    cplx *tmp = (cplx *)malloc(sizeof(cplx) * (size_t)dim);
    double *diag = (double *)malloc(sizeof(double) * (size_t)(nitermax + 1));
    double *sub = (double *)malloc(sizeof(double) * (size_t)(nitermax + 1));
    double *esave = (double *)calloc((size_t)(nitermax + 2), sizeof(double));
    double *Z = (double *)malloc(sizeof(double) * (size_t)nitermax * (size_t)nitermax);
    double a = 0.0, b = 0.0;
    int nlanc = 0;
    memcpy(vin, vect, sizeof(cplx) * (size_t)dim);
    for (int i = 0; i <= nitermax; i++) { alanc[i] = 0.0; blanc[i] = 0.0; }
    for (int iter = 1; iter <= nitermax; iter++) {
        if (!lanc_iteration(M, map, dim, iter, vin, vout, tmp, &a, &b)) break;
        if (fabs(b) < threshold) break;                                         /* :333 */
        nlanc++;
        alanc[iter - 1] = a;
        blanc[iter] = b;
        for (int i = 0; i < nlanc; i++) { diag[i] = alanc[i]; sub[i] = (i > 0) ? blanc[i] : 0.0; }
        for (size_t q = 0; q < (size_t)nlanc * nlanc; q++) Z[q] = 0.0;
        for (int i = 0; i < nlanc; i++) Z[i + (size_t)nlanc * i] = 1.0;
        ora_tql2(nlanc, diag, sub, Z);
        if (nlanc >= ncheck) {                                                  /* :352-359 */
            esave[nlanc - (ncheck - 1)] = diag[0];
            if (nlanc >= ncheck + 1) {
                double diff = esave[nlanc - (ncheck - 1)] - esave[nlanc - (ncheck - 1) - 1];
                if (fabs(diff) <= threshold) break;
            }
        }
    }
    for (int i = 0; i < nlanc; i++) { diag[i] = alanc[i]; sub[i] = (i > 0) ? blanc[i] : 0.0; }
    for (size_t q = 0; q < (size_t)nlanc * nlanc; q++) Z[q] = 0.0;
    for (int i = 0; i < nlanc; i++) Z[i + (size_t)nlanc * i] = 1.0;
    ora_tql2(nlanc, diag, sub, Z);
    *egs = diag[0];
    /* eigenvector pass (:375-384), pairing Z(k,1) with the k-th Lanczos vector v_k (SURVEY App. C note) */
    memcpy(vin, vect, sizeof(cplx) * (size_t)dim);
    memset(vout, 0, sizeof(cplx) * (size_t)dim);
    memset(vect, 0, sizeof(cplx) * (size_t)dim);
    for (int iter = 1; iter <= nlanc; iter++) {
        if (iter == 1) {
            double norm = sqrt(cdotr(dim, vin, vin));
            for (int64_t i = 0; i < dim; i++) vin[i] = vin[i] / norm;
        }
        double zk = Z[(iter - 1)];                         /* Z(iter,1) */
        for (int64_t i = 0; i < dim; i++) vect[i] += vin[i] * zk;
        if (iter == nlanc) break;
        double aa, bb = (iter == 1) ? 0.0 : blanc[iter - 1];
        memset(tmp, 0, sizeof(cplx) * (size_t)dim);
        ora_direct_hxv(M, map, dim, (const double *)vin, (double *)tmp, 0, dim);
        for (int64_t i = 0; i < dim; i++) tmp[i] = tmp[i] - bb * vout[i];
        aa = cdotr(dim, vin, tmp);
        for (int64_t i = 0; i < dim; i++) tmp[i] = tmp[i] - aa * vin[i];
        bb = sqrt(cdotr(dim, tmp, tmp));
        for (int64_t i = 0; i < dim; i++) { vout[i] = vin[i]; vin[i] = tmp[i] / bb; }
    }
    {
        double norm = sqrt(cdotr(dim, vect, vect));
        for (int64_t i = 0; i < dim; i++) vect[i] = vect[i] / norm;
    }
    free(vin); free(vout); free(tmp); free(diag); free(sub); free(esave); free(Z);
    return nlanc;
}

/* ------------------------------------------------------------------ GF seeds, observables ------------------- */

double ora_apply_op(int Ns, int isite, int dagger,
                    const uint64_t *mapI, int64_t idim, const uint64_t *mapJ, int64_t jdim,
                    const double *gs_, double *vv_)
{
    /* ED_GF_NORMAL.f90:159-174 (cdg) and :212-227 (c): vvinit = 0; vvinit(j) = sgn*state_cvec(m) */
    const cplx *gs = (const cplx *)gs_;
    cplx *vv = (cplx *)vv_;
    (void)Ns;
    for (int64_t j = 0; j < jdim; j++) vv[j] = 0.0;
    for (int64_t mi = 0; mi < idim; mi++) {
        uint64_t i = mapI[mi], r;
        double sgn;
        int bit = (int)((i >> (isite - 1)) & 1ull);
        if (dagger ? (bit == 0) : (bit == 1)) {
            if (dagger) ora_cdg(isite, i, &r, &sgn); else ora_c(isite, i, &r, &sgn);
            int64_t j = ora_binary_search(mapJ, jdim, r);
            vv[j - 1] = sgn * gs[mi];
        }
    }
    return cdotr(jdim, vv, vv);
}

void ora_observables(int Ns, int Norb, const uint64_t *map, int64_t dim, const double *gs_, double peso,
                     double *dens, double *dens_up, double *dens_dw, double *docc, double *magz,
                     double *sz2, double *n2, double *s2tot)
{
    /* ED_OBSERVABLES.f90:127-158 */
    const cplx *gs = (const cplx *)gs_;
    double nup[ORA_MAXORB], ndw[ORA_MAXORB], sz[ORA_MAXORB], nt[ORA_MAXORB];
    for (int64_t i = 0; i < dim; i++) {
        uint64_t m = map[i];
        double w = peso * cabs(gs[i]) * cabs(gs[i]);
        double ssz = 0.0;
        for (int io = 0; io < Norb; io++) {
            nup[io] = (double)((m >> io) & 1ull);
            ndw[io] = (double)((m >> (io + Ns)) & 1ull);
            sz[io] = (nup[io] - ndw[io]) / 2.0;
            nt[io] = nup[io] + ndw[io];
            ssz += sz[io];
        }
        for (int io = 0; io < Norb; io++) {
            dens[io] += nt[io] * w;
            dens_up[io] += nup[io] * w;
            dens_dw[io] += ndw[io] * w;
            docc[io] += nup[io] * ndw[io] * w;
            magz[io] += (nup[io] - ndw[io]) * w;
            sz2[io + Norb * io] += sz[io] * sz[io] * w;
            n2[io + Norb * io] += nt[io] * nt[io] * w;
            for (int jo = io + 1; jo < Norb; jo++) {
                sz2[io + Norb * jo] += sz[io] * sz[jo] * w;
                sz2[jo + Norb * io] += sz[jo] * sz[io] * w;
                n2[io + Norb * jo] += nt[io] * nt[jo] * w;
                n2[jo + Norb * io] += nt[jo] * nt[io] * w;
            }
        }
        *s2tot += ssz * ssz * w;
    }
}

/* ------------------------------------------------------------------ Philox ---------------------------------- */

static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void ora_philox_uniform(uint64_t seed, int64_t i0, int64_t n, double *out)
{
    /* uniform in (-1,1), exactly representable arithmetic => bit-identical to the CUDA generator.
     * (The reference seeds Lanczos with random_number() uniforms, .repo/PLAIN_LANCZOS.f90:310-318.) */
    for (int64_t q = 0; q < n; q++) {
        uint64_t idx = (uint64_t)(i0 + q);
        uint32_t c[4] = { (uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u };
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        uint64_t a = ((uint64_t)c[0] << 20) ^ (uint64_t)(c[1] >> 12);     /* 52 bits */
        out[q] = ((double)a + 0.5) * (1.0 / 2251799813685248.0) - 1.0;    /* (a+0.5)/2^51 - 1 */
    }
}

void ora_philox_normal(uint64_t seed, int64_t i0, int64_t n, double *out)
{
    /* element index is the counter: reproducible for any partition of the vector */
    for (int64_t q = 0; q < n; q++) {
        uint64_t idx = (uint64_t)(i0 + q);
        uint32_t c[4] = { (uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u };
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        uint64_t a = ((uint64_t)c[0] << 21) ^ (uint64_t)(c[1] >> 11);     /* 53 bits */
        uint64_t b = ((uint64_t)c[2] << 21) ^ (uint64_t)(c[3] >> 11);
        double u1 = ((double)a + 0.5) * (1.0 / 9007199254740992.0);
        double u2 = ((double)b + 0.5) * (1.0 / 9007199254740992.0);
        out[q] = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
    }
}

/* ------------------------------------------------------------------ window oracle ---------------------------- */
/* Rows of H v for sectors too large to hold map + complex vectors on the host (Ns=16: 4 GB, Ns=18: 57 GB).
 * Same per-row rule as ora_gather_hxv (diag_terms / offdiag_terms, i.e. direct/HxV*.f90 in stored/row form);
 * only the two LOOKUPS are replaced by closed forms that tests/test_oracle_invariants.py pins to the literal ones:
 *   map(i)            -> colex unrank of (i mod DimUp, i / DimUp)     (ED_SETUP.f90:899-916 ordering)
 *   binary_search(k)  -> colex rank of the two Ns-bit halves of k     (ED_SETUP.f90:1307-1324)
 * and vin(j) is the Philox uniform with counter j (ora_philox_uniform), imaginary part 0, so nothing of size Dim is
 * ever stored.  rows[n]: 0-based reference indices i; out[n]: Re (H v)(i)  (Im is identically 0 for real vin). */
static uint64_t colex_unrank(int64_t r, int n)
{
    uint64_t w = 0;
    for (int k = n; k >= 1; k--) {
        int p = k - 1;
        while (ora_binomial(p + 1, k) <= r) p++;
        w |= 1ull << p;
        r -= ora_binomial(p, k);
    }
    return w;
}

static int64_t colex_rank(uint64_t w)
{
    int64_t r = 0;
    int i = 1;
    for (int p = 0; p < 64; p++)
        if ((w >> p) & 1ull) { r += ora_binomial(p, i); i++; }
    return r;
}

uint64_t ora_map_entry(int Ns, int nup, int ndw, int64_t i)
{
    const int64_t dim_up = ora_binomial(Ns, nup);
    (void)ndw;
    return colex_unrank(i % dim_up, nup) | (colex_unrank(i / dim_up, ndw) << Ns);
}

void ora_window_hxv(const ora_model *M, int nup, int ndw, uint64_t seed, const int64_t *rows, int64_t n, double *out)
{
    const int Ns = M->Ns;
    const int64_t dim_up = ora_binomial(Ns, nup);
    const uint64_t lo = (1ull << Ns) - 1ull;
    ora_term t[ORA_MAXTERMS];
    for (int64_t q = 0; q < n; q++) {
        const int64_t i = rows[q];
        const uint64_t m = ora_map_entry(Ns, nup, ndw, i);
        cplx h[3];
        diag_terms(M, m, h);
        double vi;
        ora_philox_uniform(seed, i, 1, &vi);
        cplx acc = 0.0;
        for (int part = 0; part < 3; part++) {
            acc = acc + h[part] * vi;
            int nt = offdiag_terms(M, m, part, t);
            for (int k = 0; k < nt; k++) {
                const uint64_t w = t[k].k;
                if (__builtin_popcountll(w & lo) != nup || __builtin_popcountll(w >> Ns) != ndw) continue;   /* binary_search miss */
                const int64_t j = colex_rank(w & lo) + colex_rank(w >> Ns) * dim_up;
                double vj;
                ora_philox_uniform(seed, j, 1, &vj);
                acc = acc + conj(t[k].amp) * vj;
            }
        }
        out[q] = creal(acc);
    }
}
