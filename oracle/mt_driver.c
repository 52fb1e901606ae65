/*
 * oracle/mt_driver.c -- TEST INFRASTRUCTURE ONLY (CPU-baseline timing helper).
 *
 * The reference does not thread the EZ path itself (the only OpenMP directives in
 * src/interp are disabled).  Its session state is __thread (src/interp/ez_statics.c:12-17), so the way a
 * caller parallelises the stock library is one c_ezdefset + c_ezsint per thread on its own fields.  This
 * helper does exactly that over `nfields` independent fields with `nthreads` pthreads, calling the
 * reference's own public entry points in libezref.so.  The gridset must already have been built by one
 * warm single-threaded call (the registry/gset tables are unguarded globals).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int32_t c_ezdefset(int32_t gdout, int32_t gdin);
int32_t c_ezsint(float *zout, float *zin);
int32_t c_ezuvint(float *uuout, float *vvout, float *uuin, float *vvin);
int32_t c_ezsetopt(char *option, char *value);

typedef struct {
    int32_t gdout, gdin, f0, f1, vector, rc;
    size_t nin, nout;
    float *zout, *zin, *vout, *vin;
    const char *interp, *extrap;
} job_t;

static void *worker(void *p) {
    job_t *j = (job_t *)p;
    char o1[] = "interp_degree", o2[] = "extrap_degree", v[32];
    /* options are thread-local in the reference: every worker sets its own */
    strncpy(v, j->interp, 31); v[31] = 0; c_ezsetopt(o1, v);
    strncpy(v, j->extrap, 31); v[31] = 0; c_ezsetopt(o2, v);
    c_ezdefset(j->gdout, j->gdin);
    for (int f = j->f0; f < j->f1; f++) {
        if (j->vector)
            j->rc = c_ezuvint(j->zout + f * j->nout, j->vout + f * j->nout, j->zin + f * j->nin, j->vin + f * j->nin);
        else
            j->rc = c_ezsint(j->zout + f * j->nout, j->zin + f * j->nin);
    }
    return NULL;
}

/* Interpolate nfields fields (contiguous, nin / nout floats each) with nthreads threads. */
int32_t ezref_mt(int32_t gdout, int32_t gdin, float *zout, float *zin, float *vout, float *vin,
                 int64_t nin, int64_t nout, int32_t nfields, int32_t nthreads,
                 const char *interp, const char *extrap) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nfields) nthreads = nfields;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    job_t *jobs = (job_t *)calloc(nthreads, sizeof(job_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t].gdout = gdout; jobs[t].gdin = gdin;
        jobs[t].f0 = (int)((int64_t)nfields * t / nthreads);
        jobs[t].f1 = (int)((int64_t)nfields * (t + 1) / nthreads);
        jobs[t].vector = (vout != NULL);
        jobs[t].nin = (size_t)nin; jobs[t].nout = (size_t)nout;
        jobs[t].zout = zout; jobs[t].zin = zin; jobs[t].vout = vout; jobs[t].vin = vin;
        jobs[t].interp = interp; jobs[t].extrap = extrap;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    int32_t rc = 0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc < 0) rc = jobs[t].rc;
        else if (rc >= 0 && jobs[t].rc > rc) rc = jobs[t].rc;
    }
    free(th); free(jobs);
    return rc;
}
