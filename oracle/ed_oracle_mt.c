/*
 * ed_oracle_mt.c -- CPU baseline driver for the oracle (test/bench infrastructure, NOT product code).
 *
 * Times the reference algorithm (ed_oracle.c: integer map + bdecomp + O(Ns) sign loops + recursive
 * binary_search + complex(8) vectors) on P host threads, each owning the row block MpiIstart..MpiIend of
 * ED_HAMILTONIAN.f90:56-62 and reading a shared full input vector, written in gather form because the
 * reference's scatter form races across workers (and is out of bounds in its own MPI variant).
 */
#include "ed_oracle.h"
#include <pthread.h>
#include <time.h>
#include <unistd.h>

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int ora_num_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }

/* Applies rows [i0,i1) of H to vin with nthreads workers (row blocks Q=(i1-i0)/P, remainder to the last
 * worker, ED_HAMILTONIAN.f90:56-62).  Returns wall seconds. */
typedef struct {
    const ora_model *m; const uint64_t *map; int64_t dim; const double *vin; double *hv; int64_t a, b;
} mt_job;

static void *mt_worker(void *p)
{
    mt_job *j = (mt_job *)p;
    ora_gather_hxv(j->m, j->map, j->dim, j->vin, j->hv, j->a, j->b);
    return 0;
}

double ora_gather_hxv_mt(const ora_model *m, const uint64_t *map, int64_t dim,
                         const double *vin, double *hv, int64_t i0, int64_t i1, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t th[1024];
    mt_job jobs[1024];
    int64_t n = i1 - i0, q = n / nthreads;
    double t0 = now_s();
    for (int r = 0; r < nthreads; r++) {
        int64_t a = i0 + r * q, b = (r == nthreads - 1) ? i1 : a + q;
        jobs[r] = (mt_job){ m, map, dim, vin, hv, a, b };
        pthread_create(&th[r], 0, mt_worker, &jobs[r]);
    }
    for (int r = 0; r < nthreads; r++) pthread_join(th[r], 0);
    return now_s() - t0;
}

/* Serial literal scatter form on states [j0,j1): the reference's serial directMatVec_cc. Returns seconds. */
double ora_direct_hxv_timed(const ora_model *m, const uint64_t *map, int64_t dim,
                            const double *vin, double *hv, int64_t j0, int64_t j1)
{
    double t0 = now_s();
    ora_direct_hxv(m, map, dim, vin, hv, j0, j1);
    return now_s() - t0;
}
