/*
 * oracle/fortran_shim.c -- TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).
 *
 * A plain-C restatement of the Fortran leaf routines that the reference's EZ
 * interpolator C layer (src/interp/ *.c) calls.  This container has no Fortran
 * compiler, so these leaves are re-expressed in C with gfortran's calling
 * convention (trailing underscore, every argument by reference, hidden string
 * lengths last as int32, column-major arrays, 1-based indices).  Linked with
 * the reference's own, unmodified C sources (compiled in place from
 * /root/reference by oracle/Makefile) it forms oracle/_ref/libezref.so.
 * Built alone it forms oracle/libezshim.so (leaf functions only).
 *
 * Nothing in the product (librmn_b200/) links, loads or calls this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * PARITY STATUS: "parity unpinned" upstream -- the reference ships no tests,
 * golden vectors or fixtures for ezscint (tests/CMakeLists.txt:1-75 registers
 * only fst/wb/gmm/ccard/burp programs).  The pins that exist are the ones in
 * tests/test_oracle_*.py (analytic identities, polynomial exactness, inverse
 * round trips) and the frozen vectors under tests/golden/.
 *
 * Every function cites the reference file:line it restates.  Precision rules
 * followed throughout (Fortran typing semantics, gfortran -O2 without
 * -ffast-math; compile this file with -ffp-contract=off):
 *   - un-suffixed real literals and DATA constants are fp32;
 *   - real*4 (op) integer is evaluated in fp32;
 *   - real*4 (op) real*8 is evaluated in fp64;
 *   - a statement function declared REAL rounds its result to fp32;
 *   - ifix/int truncate toward zero, nint rounds half away from zero,
 *     amod == fmodf, sign(a,b) == copysign(|a|,b).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t fint;
typedef const int32_t F2Cl;

/* pi.cdk:6-10 -- DATA-initialised default reals, hence fp32 */
static const float PIE_F = 3.1415926535898f;
static const float RDTODG_F = 57.295779513082f;
static const float DGTORD_F = 1.7453292519943e-2f;

/* qqqpar.cdk */
enum { GLOBAL_ = 0, NORD_ = 1, SUD_ = 2, RELATIF_ = 1 };

static inline int f_ifix(float x) { return (int)x; }
static inline int f_nint(float x) { return (int)lroundf(x); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline float f_sign(float a, float b) { return copysignf(fabsf(a), b); }
static inline double d_sign(double a, double b) { return copysign(fabs(a), b); }

static void off_path(const char *name) {
    fprintf(stderr, "oracle/fortran_shim: %s is outside the restated hot path (SURVEY.md 8c)\n", name);
    abort();
}

/* ------------------------------------------------------------------ */
/* src/base helpers                                                    */
/* ------------------------------------------------------------------ */

/* src/base/xyfll.F:62-79 (all fp32, implicit typing) */
static void xyfll(float *x, float *y, float dlat, float dlon, float d60, float dgrw, int nhem) {
    float re = 1.866025f * 6.371e+6f / d60;
    float glon = dlon;
    if (nhem == 2) glon = -dlon;
    float glat = dlat;
    if (nhem == 2) glat = -dlat;
    float rlon = DGTORD_F * (glon + dgrw);
    float rlat = DGTORD_F * glat;
    float sinlat = sinf(rlat);
    float r = re * sqrtf((1.f - sinlat) / (1.f + sinlat));
    *x = r * cosf(rlon);
    *y = r * sinf(rlon);
}

/* src/base/llfxy.F:64-99 (all fp32) */
static void llfxy(float *dlat, float *dlon, float x, float y, float d60, float dgrw, int nhem) {
    float re = 1.866025f * 6.371e+6f / d60;
    float re2 = re * re;
    *dlat = 90.f;
    *dlon = 0.f;
    if (!(x == 0.f && y == 0.f)) {
        if (x == 0.f) *dlon = f_sign(90.f, y);
        if (x != 0.f) *dlon = atanf(y / x) * RDTODG_F;
        if (x < 0.f) *dlon = *dlon + f_sign(180.f, y);
        *dlon = *dlon - dgrw;
        if (*dlon > 180.f) *dlon = *dlon - 360.f;
        if (*dlon < -180.f) *dlon = *dlon + 360.f;
        float r2 = x * x + y * y;
        *dlat = (re2 - r2) / (re2 + r2);
        *dlat = asinf(*dlat) * RDTODG_F;
    }
    if (nhem == 2) *dlat = -*dlat;
    if (nhem == 2) *dlon = -*dlon;
}

/* src/base/igaxg.f90:5-143 */
void cigaxg_(char *cgtyp, float *xg1, float *xg2, float *xg3, float *xg4,
             const fint *ig1, const fint *ig2, const fint *ig3, const fint *ig4, F2Cl l) {
    (void)l;
    char t = cgtyp[0];
    if (t == 'N' || t == 'S') {
        if (*ig4 < 32768) { /* old-style encoding */
            *xg1 = (float)*ig2 * 0.1f;
            *xg2 = (float)*ig1 * 0.1f;
            *xg3 = (float)*ig4 * 100.f;
            *xg4 = (float)*ig3 * 0.01f;
        } else {
            int jg3 = *ig3, jg4 = *ig4;
            jg4 = jg4 - 32768;
            *xg3 = (float)*ig1 * 100.f;
            if (*ig3 > 32767) {
                *xg3 = *xg3 * 10.f;
                jg3 = jg3 - 32768;
            }
            *xg4 = (float)*ig2 * .1f;
            if (jg4 > 16383) {
                *xg4 = 360.f - *xg4;
                jg4 = jg4 - 16384;
            }
            float dlat = 90.f - ((float)jg4 * 180.f / 16383.f);
            float dlon = ((float)jg3 * 360.f / 32767.f);
            int ihem = 1;
            if (t == 'S') ihem = 2;
            xyfll(xg1, xg2, dlat, dlon, *xg3, *xg4, ihem);
            *xg1 = 1.0f - *xg1;
            *xg2 = 1.0f - *xg2;
        }
    } else if (t == 'C') {
        *xg1 = (float)*ig3 * 0.01f - 90.f;
        *xg2 = (float)*ig4 * 0.01f;
        *xg3 = 180.f / (float)*ig1;
        *xg4 = 360.f / (float)*ig2;
    } else if (t == 'A' || t == 'B' || t == 'G') {
        *xg1 = (float)*ig1;
        *xg2 = (float)*ig2;
        *xg3 = 0.f;
        *xg4 = 0.f;
    } else if (t == 'L') {
        *xg1 = (float)*ig3 * 0.01f - 90.f;
        *xg2 = (float)*ig4 * 0.01f;
        *xg3 = (float)*ig1 * 0.01f;
        *xg4 = (float)*ig2 * 0.01f;
    } else if (t == 'H') {
        *xg1 = (float)*ig3;
        *xg2 = .01f * (float)*ig4 - 90.f;
        *xg3 = (float)(500 * *ig2);
        *xg4 = (float)*ig1 * .2f;
    } else if (t == 'E') {
        int i2b = *ig3 & 3;
        int lg3 = (int)((uint32_t)*ig3 >> 2);
        int lg1 = (int)(((uint32_t)*ig1 << 2) | (uint32_t)i2b);
        i2b = *ig4 & 3;
        int lg4 = (int)((uint32_t)*ig4 >> 2);
        int lg2 = (int)(((uint32_t)*ig2 << 2) | (uint32_t)i2b);
        if (lg2 > 3600) lg2 = lg2 - 7201;
        *xg1 = (float)(((double)lg1 - 3600.0) / 40.0);
        if (lg3 < 3559) lg3 = lg3 + 16384;
        *xg2 = (float)(((double)lg3 - 3600.0) / 40.0);
        *xg3 = (float)((double)lg2 / 40.0);
        *xg4 = (float)((double)lg4 / 40.0);
    } else if (t == '+') {
        double xg18 = ((double)*ig3 * .01) - 100.0;
        if (*ig1 != 0) xg18 = xg18 + (double)(*ig1 - 1000) * .00001;
        double xg28 = ((double)*ig4 * .01);
        if (*ig2 != 0) xg28 = xg28 + (double)(*ig2 - 1000) * .00001;
        *xg1 = (float)xg18;
        *xg2 = (float)xg28;
        *xg3 = 0.f;
        *xg4 = 0.f;
    } else {
        fprintf(stderr, "cigaxg: Bad grid pecification (TYPE)\n");
    }
}

static void valide(const char *nom, int ichk, int mn, int mx) {
    if (ichk < mn || ichk > mx)
        printf("MAUVAISE VALEUR POUR%10sVALEUR=%10dMINIMUM=%10dMAXIMUM=%10d\n", nom, ichk, mn, mx);
}

/* src/base/xgaig.f90:4-196 */
void cxgaig_(char *cgtyp, fint *ig1, fint *ig2, fint *ig3, fint *ig4,
             float *xg1, float *xg2, float *xg3, float *xg4, F2Cl l) {
    (void)l;
    char t = cgtyp[0];
    if (t == 'N' || t == 'S') {
        *ig1 = f_nint(*xg2 * 10.f);
        *ig2 = f_nint(*xg1 * 10.f);
        *ig3 = f_nint(*xg4 * 100.f);
        *ig4 = f_nint(*xg3 * 0.01f);
        while (*ig3 < 0) *ig3 = *ig3 + 36000;
        if (*ig1 < 0 || *ig2 < 0 || *ig1 > 2047 || *ig2 > 2047 || *ig4 > 32000) {
            *ig1 = 0; *ig2 = 0; *ig3 = 0; *ig4 = 32768;
            if (*xg3 > 204700.f) {
                *ig3 = 32768;
                *ig1 = f_nint(*xg3 * .001f);
            } else {
                *ig3 = 0;
                *ig1 = f_nint(*xg3 * .01f);
            }
            *ig2 = f_nint(*xg4 * 10.f);
            if (*ig2 < 0) {
                *ig2 = abs(*ig2);
                *ig4 = *ig4 + 16384;
            }
            if (*ig2 > 1800) {
                *ig2 = abs(*ig2 - 3600);
                *ig4 = *ig4 + 16384;
            }
            int ihem = 1;
            if (t == 'S') ihem = 2;
            float dlat, dlon;
            llfxy(&dlat, &dlon, 1.f - *xg1, 1.f - *xg2, *xg3, *xg4, ihem);
            dlat = 90.f - dlat;
            if (dlon < 0.f) dlon = dlon + 360.f;
            *ig3 = *ig3 + f_nint(dlon * 32767.f / 360.f);
            *ig4 = *ig4 + f_nint(dlat * 16383.f / 180.f);
        }
    } else if (t == 'A' || t == 'B' || t == 'G') {
        *ig1 = (int)*xg1;
        *ig2 = (int)*xg2;
        *ig3 = 0;
        *ig4 = 0;
        valide("IG1", *ig1, 0, 2);
        valide("IG2", *ig2, 0, 1);
    } else if (t == 'C') {
        *ig1 = f_nint(180.f / *xg3);
        *ig2 = f_nint(360.f / *xg4);
        *ig3 = f_nint((90.f + *xg1) * 100.f);
        *ig4 = f_nint(*xg2 * 100.f);
        while (*ig4 < 0) *ig4 = *ig4 + 36000;
        if (*ig3 < 0) fprintf(stderr, "cxgaig: Bad specification (LAT0)\n");
    } else if (t == 'H') {
        *ig1 = f_nint(5.f * *xg4);
        while (*ig1 < 0) *ig1 = *ig1 + 1800;
        *ig2 = f_nint(.002f * *xg3);
        *ig3 = f_nint(*xg1);
        *ig4 = f_nint(100.f * (90.f + *xg2));
    } else if (t == 'L') {
        *ig1 = f_nint(*xg3 * 100.f);
        *ig2 = f_nint(*xg4 * 100.f);
        *ig3 = f_nint((90.f + *xg1) * 100.f);
        *ig4 = f_nint(*xg2 * 100.f);
        while (*ig4 < 0) *ig4 = *ig4 + 36000;
        if (*ig3 < 0) fprintf(stderr, "cxgaig: Bad specification (LAT0)\n");
    } else if (t == 'E') {
        valide("XG1", f_nint(*xg1), -90, 90);
        valide("XG3", f_nint(*xg3), -90, 90);
        float xxg2 = *xg2, xxg4 = *xg4;
        while (xxg2 < 0.f) xxg2 = xxg2 + 360.f;
        while (xxg4 < 0.f) xxg4 = xxg4 + 360.f;
        *ig1 = f_nint((*xg1 + 90.f) * 40.f);
        *ig2 = f_nint(*xg3 * 40.f);
        *ig3 = f_nint((xxg2 + 90.f) * 40.f);
        if (*ig3 >= 16384) *ig3 = *ig3 - 16384;
        *ig4 = f_nint(xxg4 * 40.f);
        int i2b = *ig1 & 3;
        *ig1 = (int)((uint32_t)*ig1 >> 2);
        *ig3 = (int)(((uint32_t)*ig3 << 2) | (uint32_t)i2b);
        if (*ig2 < 0) *ig2 = *ig2 + 7201;
        i2b = *ig2 & 3;
        *ig2 = (int)((uint32_t)*ig2 >> 2);
        *ig4 = (int)(((uint32_t)*ig4 << 2) | (uint32_t)i2b);
    } else if (t == '+') {
        double xlat8 = *xg1;
        valide("XG1", (int)lround(xlat8), -90, 90);
        double xlon8 = *xg2;
        if (xlon8 < 0) xlon8 = xlon8 + 360.0f;
        valide("XG2", (int)lround(xlon8), 0, 360);
        *ig3 = (int)lround((xlat8 + 100.f) * 100.f);
        *ig4 = (int)lround(xlon8 * 100.f);
        *ig1 = (int)lround((xlat8 + 100.f) * 100000.f) - *ig3 * 1000;
        *ig1 = *ig1 + 1000;
        *ig2 = (int)lround(xlon8 * 100000.f) - *ig4 * 1000;
        *ig2 = *ig2 + 1000;
    } else {
        fprintf(stderr, "cxgaig: Unknown grid (TYPE)\n");
    }
}

/* src/base/grll.f:24-67 */
void grll_(float *xlat, float *xlon, fint *ni, fint *nj, float *xla0, float *xlo0, float *dla0, float *dlo0) {
    for (int j = 1; j <= *nj; j++) {
        float xla = *xla0 + (float)(j - 1) * *dla0;
        for (int i = 1; i <= *ni; i++) {
            size_t k = (size_t)(j - 1) * *ni + (i - 1);
            xlat[k] = xla;
            xlon[k] = fmodf(*xlo0 + (float)(i - 1) * *dlo0, 360.0f);
        }
    }
}

/* src/base/grps.f:24-75 */
void grps_(float *xlat, float *xlon, fint *ni, fint *nj, float *pi, float *pj, float *d60, float *dgrw, fint *hem) {
    for (int j = 1; j <= *nj; j++) {
        float y = (float)j - *pj;
        for (int i = 1; i <= *ni; i++) {
            float xla, xlo;
            llfxy(&xla, &xlo, (float)i - *pi, y, *d60, *dgrw, *hem);
            size_t k = (size_t)(j - 1) * *ni + (i - 1);
            xlat[k] = xla;
            if (xlo < 0.f) xlo = xlo + 360.f;
            xlon[k] = xlo;
        }
    }
}

/* src/base/permut.f:22-69 */
void permut_(float *z, fint *ni, fint *nj) {
    int ncc = *nj / 2;
    for (int j = 1; j <= ncc; j++) {
        float *a = z + (size_t)(*nj - j) * *ni;
        float *b = z + (size_t)(j - 1) * *ni;
        for (int i = 0; i < *ni; i++) {
            float t = a[i];
            a[i] = b[i];
            b[i] = t;
        }
    }
}

/* src/base/mxm.F90:23-40 -- c(nar,nbc) = a(nar,nac) * b(nac,nbc), column-major, sequential fp32 sums */
void mxm_(const float *a, const fint *nar, const float *b, const fint *nac, float *c, const fint *nbc) {
    int NAR = *nar, NAC = *nac, NBC = *nbc;
    for (int j = 0; j < NBC; j++) {
        for (int i = 0; i < NAR; i++) {
            float s = 0.0f;
            for (int k = 0; k < NAC; k++) {
                s = s + a[(size_t)k * NAR + i] * b[(size_t)j * NAC + k];
            }
            c[(size_t)j * NAR + i] = s;
        }
    }
}

/* ------------------------------------------------------------------ */
/* rotations: ez_lac, ez_cal, ez_crot, ez_gfxyfll, ez_gfllfxy           */
/* ------------------------------------------------------------------ */

/* src/interp/ez_lac.inc:31-47 */
void ez_lac_(float *xyz, float *lon, float *lat, fint *nbpts) {
    float dar = acosf(-1.0f) / 180.0f;
    for (int i = 0; i < *nbpts; i++) {
        float cosdar = cosf(dar * lat[i]);
        xyz[3 * (size_t)i + 0] = cosdar * cosf(dar * lon[i]);
        xyz[3 * (size_t)i + 1] = cosdar * sinf(dar * lon[i]);
        xyz[3 * (size_t)i + 2] = sinf(dar * lat[i]);
    }
}

/* src/interp/ez_cal.inc:22-47 */
void ez_cal_(float *lon, float *lat, float *xyz, fint *n) {
    float rad = 180.f / acosf(-1.00f);
    for (int i = 0; i < *n; i++) {
        lat[i] = asinf(fmaxf(-1.00f, fminf(1.0f, xyz[3 * (size_t)i + 2]))) * rad;
        lon[i] = atan2f(xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 0]) * rad;
        lon[i] = fmodf(lon[i], 360.0f);
        if (lon[i] < 0.0f) lon[i] = lon[i] + 360.0f;
    }
}

/* src/interp/ez_crot.inc:44-97 ; r and ri are column-major 3x3: r(i,j) = r[(j-1)*3+(i-1)] */
void ez_crot_(float *r, float *ri, float *lon1, float *lat1, float *lon2, float *lat2) {
    float xyz1[3], xyz2[3];
    fint one = 1;
    ez_lac_(xyz1, lon1, lat1, &one);
    ez_lac_(xyz2, lon2, lat2, &one);
#define R(i, j) r[((j)-1) * 3 + ((i)-1)]
#define RI(i, j) ri[((j)-1) * 3 + ((i)-1)]
    float a = (xyz1[0] * xyz2[0]) + (xyz1[1] * xyz2[1]) + (xyz1[2] * xyz2[2]);
    float t1 = (xyz1[1] * xyz2[2]) - (xyz2[1] * xyz1[2]);
    float t2 = (xyz2[0] * xyz1[2]) - (xyz1[0] * xyz2[2]);
    float t3 = (xyz1[0] * xyz2[1]) - (xyz2[0] * xyz1[1]);
    float b = sqrtf(t1 * t1 + t2 * t2 + t3 * t3);
    float c = sqrtf(xyz1[0] * xyz1[0] + xyz1[1] * xyz1[1] + xyz1[2] * xyz1[2]);
    float d1 = ((a * xyz1[0]) - xyz2[0]) / b;
    float d2 = ((a * xyz1[1]) - xyz2[1]) / b;
    float d3 = ((a * xyz1[2]) - xyz2[2]) / b;
    float d = sqrtf(d1 * d1 + d2 * d2 + d3 * d3);
    R(1, 1) = -xyz1[0] / c;
    R(1, 2) = -xyz1[1] / c;
    R(1, 3) = -xyz1[2] / c;
    R(2, 1) = (((a * xyz1[0]) - xyz2[0]) / b) / d;
    R(2, 2) = (((a * xyz1[1]) - xyz2[1]) / b) / d;
    R(2, 3) = (((a * xyz1[2]) - xyz2[2]) / b) / d;
    R(3, 1) = ((xyz1[1] * xyz2[2]) - (xyz2[1] * xyz1[2])) / b;
    R(3, 2) = ((xyz2[0] * xyz1[2]) - (xyz1[0] * xyz2[2])) / b;
    R(3, 3) = ((xyz1[0] * xyz2[1]) - (xyz2[0] * xyz1[1])) / b;
    for (int i = 1; i <= 3; i++)
        for (int j = 1; j <= 3; j++) RI(i, j) = R(j, i);
#undef R
#undef RI
}

/* The Fortran wrappers ez_gfxyfll / ez_gfllfxy (ez_gfxyfll.inc:20-60, ez_gfllfxy.inc) call back
 * into the reference's C ezgfxyfll_/ezgfllfxy_ (ez_gfxyfll.c:29-57, ez_gfllfxy.c:29-58) when those are
 * linked in (libezref.so); the stand-alone shim composes the same four leaves itself. */
extern void ezgfxyfll_(float *, float *, float *, float *, float *, float *, fint *, float *, float *, float *, float *)
    __attribute__((weak));
extern void ezgfllfxy_(float *, float *, float *, float *, float *, float *, fint *, float *, float *, float *, float *)
    __attribute__((weak));

void ez_gfxyfll_(float *lonp, float *latp, float *lon, float *lat, fint *n,
                 float *xlat1, float *xlon1, float *xlat2, float *xlon2) {
    float r[9], ri[9];
    if (ezgfxyfll_) {
        ezgfxyfll_(lonp, latp, lon, lat, r, ri, n, xlat1, xlon1, xlat2, xlon2);
        return;
    }
    fint trois = 3;
    float *cart = (float *)malloc(3 * (size_t)*n * sizeof(float));
    float *carot = (float *)malloc(3 * (size_t)*n * sizeof(float));
    ez_crot_(r, ri, xlon1, xlat1, xlon2, xlat2);
    ez_lac_(cart, lonp, latp, n);
    mxm_(r, &trois, cart, &trois, carot, n);
    ez_cal_(lon, lat, carot, n);
    free(cart);
    free(carot);
}

void ez_gfllfxy_(float *lonp, float *latp, float *lon, float *lat, fint *n,
                 float *xlat1, float *xlon1, float *xlat2, float *xlon2) {
    float r[9], ri[9];
    if (ezgfllfxy_) {
        ezgfllfxy_(lonp, latp, lon, lat, r, ri, n, xlat1, xlon1, xlat2, xlon2);
        return;
    }
    fint trois = 3;
    float *cart = (float *)malloc(3 * (size_t)*n * sizeof(float));
    float *carot = (float *)malloc(3 * (size_t)*n * sizeof(float));
    ez_crot_(r, ri, xlon1, xlat1, xlon2, xlat2);
    ez_lac_(cart, lon, lat, n);
    mxm_(ri, &trois, cart, &trois, carot, n);
    ez_cal_(lonp, latp, carot, n);
    free(cart);
    free(carot);
}

/* ------------------------------------------------------------------ */
/* (lat,lon) -> (x,y) in source index space                            */
/* ------------------------------------------------------------------ */

/* src/interp/ez_cherche.inc:22-72 */
static int ez_cherche(float val, const float *tableau, int nbelem) {
    int debut = 1, fin = nbelem;
    int milieu = (int)((float)(debut + fin) * 0.5f);
    while (milieu != debut) {
        if (val <= tableau[milieu - 1])
            fin = milieu;
        else
            debut = milieu;
        milieu = (int)((float)(debut + fin) * 0.5f);
    }
    return milieu;
}

/* src/interp/ez_llll2gd.inc:22-53 -- NB mutates dlon in place */
static void ez_llll2gd(float *x, float *y, float *dlat, float *dlon, int npts,
                       float xlat0, float xlon0, float dellat, float dellon, float lonref) {
    if (lonref == -180.0f) {
        for (int i = 0; i < npts; i++)
            if (dlon[i] > 180.0f) dlon[i] = dlon[i] - 360.0f;
    } else {
        for (int i = 0; i < npts; i++)
            if (dlon[i] < 0.0f) dlon[i] = dlon[i] + 360.0f;
    }
    for (int i = 0; i < npts; i++) {
        x[i] = (dlon[i] - xlon0) / dellon + 1.0f;
        y[i] = (dlat[i] - xlat0) / dellat + 1.0f;
    }
}

/* src/interp/ez_vxyfll.inc:22-98 */
static void ez_vxyfll(float *x, float *y, float *dlat, float *dlon, int npts,
                      float d60, float dgrw, float pi, float pj, int nhem) {
    double re = 1.866025 * 6.371e+6 / (double)d60;
    if (nhem == NORD_) {
        for (int i = 0; i < npts; i++) {
            double rlon = (double)(DGTORD_F * (dlon[i] + dgrw));
            double rlat = (double)(DGTORD_F * dlat[i]);
            double sinlat = sin(rlat);
            double r = re * sqrt((1.0 - sinlat) / (1.0 + sinlat));
            x[i] = (float)(r * cos(rlon) + (double)pi);
            y[i] = (float)(r * sin(rlon) + (double)pj);
        }
        return;
    }
    if (nhem == SUD_) {
        for (int i = 0; i < npts; i++) {
            double rlon = dlon[i];
            if (rlon > 180.0) rlon = rlon - 360.0;
            rlon = (double)DGTORD_F * (-rlon + (double)dgrw);
            double rlat = (double)(DGTORD_F * (-dlat[i]));
            double sinlat = sin(rlat);
            double r = re * sqrt((1.0 - sinlat) / (1.0 + sinlat));
            x[i] = (float)(r * cos(rlon) + (double)pi);
            y[i] = (float)(r * sin(rlon) + (double)pj);
        }
    }
}

/* src/interp/ez_vllfxy.inc:22-103 */
void ez_vllfxy_(float *dlat, float *dlon, float *x, float *y, fint *ni, fint *nj,
                float *d60, float *dgrw, float *pi, float *pj, fint *nhem) {
    double re = 1.866025 * (double)6.371e+6f / (double)*d60;
    double re2 = re * re;
    size_t n = (size_t)*ni * (size_t)*nj;
    double rlat = 0.0, rlon = 0.0;
    for (size_t k = 0; k < n; k++) {
        double x1 = (double)(x[k] - *pi);
        double y1 = (double)(y[k] - *pj);
        if (x1 == 0. && y1 == 0.) {
            rlat = 90.0;
            rlon = 0.0;
        }
        if (x1 == 0.0) rlon = d_sign(90., y1);
        if (x1 != 0.0) rlon = atan(y1 / x1) * (double)RDTODG_F;
        if (x1 < 0.0) rlon = rlon + d_sign(180., y1);
        rlon = rlon - (double)*dgrw;
        if (rlon < 0.0) rlon = rlon + 3.6e2;
        double r2 = x1 * x1 + y1 * y1;
        rlat = (re2 - r2) / (re2 + r2);
        rlat = fmax(-1.0, fmin(rlat, 1.0));
        rlat = asin(rlat) * (double)RDTODG_F;
        if (*nhem == 2) {
            rlat = -rlat;
            rlon = -rlon;
            if (rlon < 0.0) rlon = rlon + 360.0;
        }
        dlat[k] = (float)rlat;
        dlon[k] = (float)rlon;
    }
}

/* src/interp/ez_ggll2gd.inc:23-53 */
static void ez_ggll2gd(float *x, float *y, float *xlat, float *xlon, int npts, int ni, int nj, float *lroots) {
    float dellon = 360.0f / (float)ni;
    float xlon0 = 0.0f;
    for (int i = 0; i < npts; i++) x[i] = (xlon[i] - xlon0) / dellon + 1.0f;
    for (int i = 0; i < npts; i++) {
        int indy = ez_cherche(xlat[i], lroots, nj);
        if (indy >= nj) indy = nj - 1;
        y[i] = (float)indy + (xlat[i] - lroots[indy - 1]) / (lroots[indy] - lroots[indy - 1]);
    }
}

/* src/interp/ez_ll2ergd.inc:22-51 */
static void ez_ll2ergd(float *px, float *py, float *xlat, float *xlon, int npts, int ni, int nj,
                       char *grtyp, fint *ig1, fint *ig2, fint *ig3, fint *ig4) {
    float xlat1, xlon1, xlat2, xlon2;
    float *xlatgf = (float *)malloc((size_t)npts * sizeof(float));
    float *xlongf = (float *)malloc((size_t)npts * sizeof(float));
    fint n = npts;
    cigaxg_(grtyp, &xlat1, &xlon1, &xlat2, &xlon2, ig1, ig2, ig3, ig4, 1);
    ez_gfxyfll_(xlon, xlat, xlongf, xlatgf, &n, &xlat1, &xlon1, &xlat2, &xlon2);
    if (grtyp[0] == 'E') {
        float dellon = 360.0f / (float)(ni - 1);
        float xlon0 = 0.0f;
        float dellat = 180.0f / (float)nj;
        float xlat0 = -90.f + 0.5f * dellat;
        ez_llll2gd(px, py, xlatgf, xlongf, npts, xlat0, xlon0, dellat, dellon, 0.0f);
    }
    free(xlatgf);
    free(xlongf);
}

/* src/interp/ez_ll2rgd.inc:22-166 ('T' and '!' projections are outside the restated path) */
void ez_ll2rgd_(float *px, float *py, float *xlat, float *xlon, fint *npts, fint *ni, fint *nj,
                char *grtyp, fint *ig1, fint *ig2, fint *ig3, fint *ig4, fint *sym, float *lroots, F2Cl l) {
    (void)l; (void)sym;
    char t = grtyp[0];
    float pi, pj, dgrw, d60, dellat = 0, dellon, xlat0 = 0, xlon0;
    if (t == 'N') {
        cigaxg_(grtyp, &pi, &pj, &d60, &dgrw, ig1, ig2, ig3, ig4, 1);
        ez_vxyfll(px, py, xlat, xlon, *npts, d60, dgrw, pi, pj, NORD_);
        return;
    }
    if (t == 'S') {
        cigaxg_(grtyp, &pi, &pj, &d60, &dgrw, ig1, ig2, ig3, ig4, 1);
        ez_vxyfll(px, py, xlat, xlon, *npts, d60, dgrw, pi, pj, SUD_);
        return;
    }
    if (t == 'T' || t == '!') off_path("ez_ll2rgd('T'/'!')");
    if (t == 'A') {
        dellon = 360.0f / (float)*ni;
        xlon0 = 0.0f;
        if (*ig1 == GLOBAL_) { dellat = 180.0f / (float)*nj; xlat0 = -90.0f + dellat * 0.5f; }
        if (*ig1 == NORD_) { dellat = 90.0f / (float)*nj; xlat0 = dellat * 0.5f; }
        if (*ig1 == SUD_) { dellat = 90.0f / (float)*nj; xlat0 = -90.0f + dellat * 0.5f; }
        for (int i = 0; i < *npts; i++)
            if (xlon[i] < 0.0f) xlon[i] = xlon[i] + 360.0f;
        ez_llll2gd(px, py, xlat, xlon, *npts, xlat0, xlon0, dellat, dellon, 0.0f);
        return;
    }
    if (t == 'B') {
        dellon = 360.0f / (float)(*ni - 1);
        xlon0 = 0.0f;
        if (*ig1 == GLOBAL_) { dellat = 180.0f / (float)(*nj - 1); xlat0 = -90.0f; }
        if (*ig1 == NORD_) { dellat = 90.0f / (float)(*nj - 1); xlat0 = 0.0f; }
        if (*ig1 == SUD_) { dellat = 90.0f / (float)(*nj - 1); xlat0 = -90.0f; }
        for (int i = 0; i < *npts; i++)
            if (xlon[i] < 0.0f) xlon[i] = xlon[i] + 360.0f;
        ez_llll2gd(px, py, xlat, xlon, *npts, xlat0, xlon0, dellat, dellon, 0.0f);
        return;
    }
    if (t == 'G') {
        dellon = 360.0f / (float)*ni;
        xlon0 = 0.0f;
        for (int i = 0; i < *npts; i++)
            if (xlon[i] < 0.0f) xlon[i] = xlon[i] + 360.0f;
        if (*ig1 == GLOBAL_) {
            ez_ggll2gd(px, py, xlat, xlon, *npts, *ni, *nj, lroots);
        } else if (*ig1 == NORD_) {
            dellat = 90.0f / (float)*nj;
            xlat0 = dellat * 0.5f;
            ez_llll2gd(px, py, xlat, xlon, *npts, xlat0, xlon0, dellat, dellon, 0.0f);
        } else {
            dellat = 90.0f / (float)*nj;
            xlat0 = -90.0f + dellat * 0.5f;
            ez_llll2gd(px, py, xlat, xlon, *npts, xlat0, xlon0, dellat, dellon, 0.0f);
        }
        return;
    }
    if (t == 'L') {
        cigaxg_(grtyp, &xlat0, &xlon0, &dellat, &dellon, ig1, ig2, ig3, ig4, 1);
        for (int i = 0; i < *npts; i++) {
            if (xlon[i] < xlon0) xlon[i] = xlon[i] + 360.0f;
            if (xlon[i] > (xlon0 + (float)*ni * dellon)) xlon[i] = xlon[i] - 360.0f;
        }
        ez_llll2gd(px, py, xlat, xlon, *npts, xlat0, xlon0, dellat, dellon, 0.0f);
        return;
    }
    if (t == 'E') {
        ez_ll2ergd(px, py, xlat, xlon, *npts, *ni, *nj, grtyp, ig1, ig2, ig3, ig4);
        return;
    }
    fprintf(stderr, "<ez_ll2rgd> bad grid type for type: %c\n", t);
    exit(13);
}

/* src/interp/ez_ll2igd.inc:22-86 */
void ez_ll2igd_(float *px, float *py, float *xlat, float *xlon, fint *npts, fint *ni, fint *nj,
                char *grtyp, char *grref, fint *ig1, fint *ig2, fint *ig3, fint *ig4,
                float *ax, float *ay, fint *coordflag, F2Cl l1, F2Cl l2) {
    (void)l1; (void)l2; (void)grtyp;
    float pi, pj, dgrw, d60, dlat, dlon, xlat0, xlon0, xlat1, xlon1, xlat2, xlon2, lonref;
    char r = grref[0];
    if (r == 'N') {
        cigaxg_(grref, &pi, &pj, &d60, &dgrw, ig1, ig2, ig3, ig4, 1);
        ez_vxyfll(px, py, xlat, xlon, *npts, d60, dgrw, pi, pj, 1);
    }
    if (r == 'S') {
        cigaxg_(grref, &pi, &pj, &d60, &dgrw, ig1, ig2, ig3, ig4, 1);
        ez_vxyfll(px, py, xlat, xlon, *npts, d60, dgrw, pi, pj, 2);
    }
    if (r == 'L') {
        cigaxg_(grref, &xlat0, &xlon0, &dlat, &dlon, ig1, ig2, ig3, ig4, 1);
        if (ax[0] < 0.0f)
            lonref = -180.0f;
        else
            lonref = 0.0f;
        ez_llll2gd(px, py, xlat, xlon, *npts, xlat0, xlon0, dlat, dlon, lonref);
        for (int i = 0; i < *npts; i++) {
            px[i] = px[i] - 1.0f;
            py[i] = py[i] - 1.0f;
        }
    }
    if (r == 'E') {
        cigaxg_(grref, &xlat1, &xlon1, &xlat2, &xlon2, ig1, ig2, ig3, ig4, 1);
        ez_gfxyfll_(xlon, xlat, px, py, npts, &xlat1, &xlon1, &xlat2, &xlon2);
    }
    if (*coordflag == RELATIF_) {
        for (int i = 0; i < *npts; i++) {
            int indx = ez_cherche(px[i], ax, *ni);
            int indy = ez_cherche(py[i], ay, *nj);
            if (indx >= *ni) indx = *ni - 1;
            if (indy >= *nj) indy = *nj - 1;
            px[i] = (float)indx + (px[i] - ax[indx - 1]) / (ax[indx] - ax[indx - 1]);
            py[i] = (float)indy + (py[i] - ay[indy - 1]) / (ay[indy] - ay[indy - 1]);
        }
    }
}

/* ------------------------------------------------------------------ */
/* Gaussian latitudes: ez_glat, dgauss, ordleg                          */
/* ------------------------------------------------------------------ */

/* src/base/ordleg.F:25-89 (all fp32, implicit typing) */
static void ordleg(float *sx, float coa, int ir) {
    float sqr2 = sqrtf(2.f);
    int irpp = ir + 1;
    int irppm = irpp - 1;
    float delta = acosf(coa);
    float theta = delta;
    float c1 = sqr2;
    float fn = 0.f, fn2 = 0.f;
    for (int n = 1; n <= irppm; n++) {
        fn = (float)n;
        fn2 = 2.f * fn;
        float fn2sq = fn2 * fn2;
        c1 = c1 * sqrtf(1.0f - 1.0f / fn2sq);
    }
    int n = irppm;
    float ang = fn * theta;
    float s1 = 0.0f, c4 = 1.0f, a = -1.0f, b = 0.0f;
    int n1 = n + 1;
    for (int kk = 1; kk <= n1; kk += 2) {
        int k = kk - 1;
        if (k == n) c4 = 0.5f * c4;
        s1 = s1 + c4 * cosf(ang);
        a = a + 2.0f;
        b = b + 1.0f;
        float fk = (float)k;
        ang = theta * (fn - fk - 2.0f);
        c4 = (a * (fn - b + 1.0f) / (b * (fn2 - a))) * c4;
    }
    *sx = s1 * c1;
}

/* src/base/dgauss.F:25-132, KASE = GLOBAL (kases.cdk: GLOBAL=0, NORTH=1, SOUTH=2); TOL = 1.0E-6 (non-ALL64) */
static void dgauss(int n, float *roots, int kase) {
    const float tol = 1.0e-6f;
    float normn = sqrtf(2.0f / (2.0f * (float)n + 1.0f));
    float normnm = sqrtf(2.0f / (2.0f * (float)n - 1.0f));
    int l = n / 2;
    int irt = 0;
    for (int i = 1; i <= l; i++) {
        int j = (kase != 2) ? i : i + l + n % 2;
        float t = (float)(4 * j - 1) * PIE_F / (float)(4 * n + 2);
        irt = (kase != 2) ? i : i + n % 2;
        roots[irt - 1] = cosf(t + 1.0f / (8.0f * (float)(n * n) * tanf(t)));
    }
    for (int i = 1; i <= l; i++) {
        float deltav;
        do {
            float g, gm;
            ordleg(&g, roots[i - 1], n);
            ordleg(&gm, roots[i - 1], n - 1);
            float pn = normn * g;
            float pnm = normnm * gm;
            float rdpdx = (roots[i - 1] * roots[i - 1] - 1.0f) / ((float)n * (roots[i - 1] * pn - pnm));
            deltav = -pn * rdpdx;
            roots[i - 1] = roots[i - 1] + deltav;
        } while (fabsf(deltav) > tol);
        roots[n - i] = -roots[i - 1];
    }
    if (n % 2 == 0) return;
    irt = (kase != 2) ? l + 1 : 1;
    roots[irt - 1] = 0.0f;
}

/* src/interp/ez_glat.inc:26-67 */
void ez_glat_(float *latroots, float *groots, fint *nj, fint *hem) {
    int npoly = (*hem != GLOBAL_) ? *nj * 2 : *nj;
    dgauss(npoly, groots, GLOBAL_);
    for (int j = 1; j <= npoly / 2; j++) {
        float temp = groots[j - 1];
        groots[j - 1] = groots[npoly - j];
        groots[npoly - j] = temp;
    }
    if (*hem != NORD_) {
        for (int j = 1; j <= *nj; j++) latroots[j - 1] = 90.f - RDTODG_F * acosf(groots[j - 1]);
    }
    if (*hem == NORD_) {
        for (int j = 1; j <= *nj; j++) latroots[j - 1] = 90.0f - RDTODG_F * acosf(groots[j + *nj - 1]);
    }
}

/* ------------------------------------------------------------------ */
/* Newton coefficient tables                                           */
/* ------------------------------------------------------------------ */

/* src/interp/ez_nwtncof.inc:20-178 ; cx(ni,6), cy(j1:j2,6) column-major */
void ez_nwtncof_(float *cx, float *cy, float *ax, float *ay, fint *ni_, fint *nj_,
                 fint *i1, fint *i2, fint *j1_, fint *j2_, fint *extension) {
    (void)i1; (void)i2;
    int ni = *ni_, nj = *nj_, j1 = *j1_, j2 = *j2_;
    int nyr = j2 - j1 + 1;
#define AX(i) ax[(i)-1]
#define AY(j) ay[(j)-j1]
#define CX(i, k) cx[(size_t)((k)-1) * ni + ((i)-1)]
#define CY(j, k) cy[(size_t)((k)-1) * nyr + ((j)-j1)]
    for (int i = 1; i <= ni - 1; i++) {
        if (AX(i + 1) <= AX(i)) {
            printf(" Probleme detecte dans EZ_NWTNCOF code 998\n (EZ_NWTNCOF) Probleme : x1..x4  : %g %g\n EZ_NWTNCOF CALL EXIT\n",
                   AX(i), AX(i + 1));
            exit(13);
        }
    }
    for (int j = 1; j <= nj - 1; j++) {
        if (AY(j + 1) <= AY(j)) {
            printf(" Probleme detecte dans EZ_NWTNCOF code 999\n (EZ_NWTNCOF) Probleme : y1..y4  : %g %g\n EZ_NWTNCOF CALL EXIT\n",
                   AY(j), AY(j + 1));
            exit(13);
        }
    }
    for (int i = 1; i <= ni; i++)
        for (int k = 1; k <= 6; k++) CX(i, k) = 1.0f;
    for (int k = 1; k <= 6; k++)
        for (int j = j1; j <= j2; j++) CY(j, k) = 1.0f;
    for (int i = 2; i <= ni - 2; i++) {
        CX(i, 1) = 1.f / (AX(i) - AX(i - 1));
        CX(i, 2) = 1.f / (AX(i + 1) - AX(i - 1));
        CX(i, 3) = 1.f / (AX(i + 1) - AX(i));
        CX(i, 4) = 1.f / (AX(i + 2) - AX(i - 1));
        CX(i, 5) = 1.f / (AX(i + 2) - AX(i));
        CX(i, 6) = 1.f / (AX(i + 2) - AX(i + 1));
    }
    for (int j = j1 + 1; j <= j2 - 2; j++) {
        CY(j, 1) = 1.f / (AY(j) - AY(j - 1));
        CY(j, 2) = 1.f / (AY(j + 1) - AY(j - 1));
        CY(j, 3) = 1.f / (AY(j + 1) - AY(j));
        CY(j, 4) = 1.f / (AY(j + 2) - AY(j - 1));
        CY(j, 5) = 1.f / (AY(j + 2) - AY(j));
        CY(j, 6) = 1.f / (AY(j + 2) - AY(j + 1));
    }
    float x1, x2, x3, x4;
#define FILL(i)                         \
    CX(i, 1) = 1.f / (x2 - x1);         \
    CX(i, 2) = 1.f / (x3 - x1);         \
    CX(i, 3) = 1.f / (x3 - x2);         \
    CX(i, 4) = 1.f / (x4 - x1);         \
    CX(i, 5) = 1.f / (x4 - x2);         \
    CX(i, 6) = 1.f / (x4 - x3);
    if (*extension == 1) {
        x1 = AX(1) - (AX(ni) - AX(ni - 1)); x2 = AX(1); x3 = AX(2); x4 = AX(3);
        FILL(1)
        x1 = AX(ni - 2); x2 = AX(ni - 1); x3 = AX(ni); x4 = AX(ni) + (AX(2) - AX(1));
        FILL(ni - 1)
    }
    if (*extension == 2) {
        x1 = AX(1) - (360.0f - AX(ni)); x2 = AX(1); x3 = AX(2); x4 = AX(3);
        FILL(1)
        x1 = AX(ni - 2); x2 = AX(ni - 1); x3 = AX(ni); x4 = AX(1) + 360.0f;
        FILL(ni - 1)
        x1 = AX(ni - 1); x2 = AX(ni); x3 = AX(1) + 360.0f; x4 = AX(2) + 360.0f;
        FILL(ni)
    }
#undef FILL
#undef AX
#undef AY
#undef CX
#undef CY
}

/* ------------------------------------------------------------------ */
/* Regular-grid interpolation kernels                                  */
/* ------------------------------------------------------------------ */

#define Z(i, j) z[(size_t)((j)-j1) * ni + ((i)-1)]

/* zlin8.cdk:1-3 */
static inline double zlin(double zz1, double zz2, double zdx) { return zz1 + (zz2 - zz1) * zdx; }

/* cubic8.cdk:1-3 -- literals are default-real (fp32) promoted to fp64 */
static inline double cubic(double z1, double z2, double z3, double z4, double dx) {
    const double c6 = (double)0.1666666666666f, c3 = (double)0.3333333333333f, h = (double)0.5f;
    return ((((z4 - z1) * c6 + h * (z2 - z3)) * dx + h * (z1 + z3) - z2) * dx + z3 - c6 * z4 - h * z2 - c3 * z1) * dx + z2;
}

/* src/interp/ez_rgdint_0.inc:20-35 */
void ez_rgdint_0_(float *zo, float *px, float *py, fint *npts, float *z, fint *ni_, fint *j1_, fint *j2_) {
    int ni = *ni_, j1 = *j1_, j2 = *j2_;
    for (int n = 0; n < *npts; n++) {
        int i = imin(ni, imax(1, f_nint(px[n])));
        int j = imin(j2, imax(j1, f_nint(py[n])));
        zo[n] = Z(i, j);
    }
}

/* src/interp/ez_rgdint_1_nw.inc:20-44 */
void ez_rgdint_1_nw_(float *zo, float *px, float *py, fint *npts, float *z, fint *ni_, fint *j1_, fint *j2_) {
    int ni = *ni_, j1 = *j1_, j2 = *j2_;
    for (int n = 0; n < *npts; n++) {
        int i = imin(ni - 1, imax(1, f_ifix(px[n])));
        int j = imin(j2 - 1, imax(j1, f_ifix(py[n])));
        double dx = (double)(px[n] - (float)i);
        double dy = (double)(py[n] - (float)j);
        double y2 = zlin((double)Z(i, j), (double)Z(i + 1, j), dx);
        double y3 = zlin((double)Z(i, j + 1), (double)Z(i + 1, j + 1), dx);
        zo[n] = (float)zlin(y2, y3, dy);
    }
}

/* src/interp/ez_rgdint_1_w.inc:20-51 */
void ez_rgdint_1_w_(float *zo, float *px, float *py, fint *npts, float *z, fint *ni_, fint *j1_, fint *j2_, fint *wrap_) {
    int ni = *ni_, j1 = *j1_, j2 = *j2_, wrap = *wrap_;
    int limite = ni + 2 - wrap;
    for (int n = 0; n < *npts; n++) {
        int i = imin(ni - 2 + wrap, imax(1, f_ifix(px[n])));
        int j = imin(j2 - 1, imax(j1, f_ifix(py[n])));
        int iplus1 = i + 1;
        if (wrap > 0 && (i == (ni - 2 + wrap))) iplus1 = (limite + i + 1) % limite;
        double dx = (double)(px[n] - (float)i);
        double dy = (double)(py[n] - (float)j);
        double y2 = zlin((double)Z(i, j), (double)Z(iplus1, j), dx);
        double y3 = zlin((double)Z(i, j + 1), (double)Z(iplus1, j + 1), dx);
        zo[n] = (float)zlin(y2, y3, dy);
    }
}

/* src/interp/ez_rgdint_3_nw.inc:20-77 */
void ez_rgdint_3_nw_(float *zo, float *px, float *py, fint *npts, float *z, fint *ni_, fint *j1_, fint *j2_) {
    int ni = *ni_, j1 = *j1_, j2 = *j2_;
    for (int n = 0; n < *npts; n++) {
        int i = imin(ni - 2, imax(2, f_ifix(px[n])));
        int j = imin(j2 - 2, imax(j1 + 1, f_ifix(py[n])));
        double dx = (double)(px[n] - (float)i);
        double dy = (double)(py[n] - (float)j);
        double y1 = cubic(Z(i - 1, j - 1), Z(i, j - 1), Z(i + 1, j - 1), Z(i + 2, j - 1), dx);
        double y2 = cubic(Z(i - 1, j), Z(i, j), Z(i + 1, j), Z(i + 2, j), dx);
        double y3 = cubic(Z(i - 1, j + 1), Z(i, j + 1), Z(i + 1, j + 1), Z(i + 2, j + 1), dx);
        double y4 = cubic(Z(i - 1, j + 2), Z(i, j + 2), Z(i + 1, j + 2), Z(i + 2, j + 2), dx);
        zo[n] = (float)cubic(y1, y2, y3, y4, dy);
    }
}

/* shared body of ez_rgdint_3_w.inc:72-104 and ez_rgdint_3_wnnc.inc:72-104; `nnc` selects the
 * wnnc branch condition `wrap>0 .and. i<=1 .or. i>=ni-1` (Fortran precedence: .and. binds tighter) */
static void rgdint_3_w_body(float *zo, float *px, float *py, int npts, float *z, int ni, int j1, int j2, int wrap, int nnc) {
    int limite = ni + 2 - wrap;
    for (int n = 0; n < npts; n++) {
        int i = imin(ni - 2 + wrap, imax(1, imax(2 - wrap, f_ifix(px[n]))));
        int j = imin(j2 - 2, imax(j1 + 1, f_ifix(py[n])));
        int imoins1, iplus1, iplus2;
        int wrapped = nnc ? ((wrap > 0 && i <= 1) || i >= (ni - 1)) : (wrap > 0);
        if (wrapped) {
            imoins1 = (limite + i - 1) % limite;
            iplus1 = (limite + i + 1) % limite;
            iplus2 = (limite + i + 2) % limite;
            if (imoins1 == 0) imoins1 = ni;
            if (i == 0) i = ni;
            if (iplus1 == 0) iplus1 = ni;
            if (iplus2 == 0) iplus2 = ni;
            if (wrap == 1) {
                if (iplus2 == ni) iplus2 = 2;
                if (imoins1 == ni) imoins1 = ni - 1;
            }
        } else {
            imoins1 = i - 1;
            iplus1 = i + 1;
            iplus2 = i + 2;
        }
        double dx = (double)(px[n] - (float)i);
        double dy = (double)(py[n] - (float)j);
        double y1 = cubic(Z(imoins1, j - 1), Z(i, j - 1), Z(iplus1, j - 1), Z(iplus2, j - 1), dx);
        double y2 = cubic(Z(imoins1, j), Z(i, j), Z(iplus1, j), Z(iplus2, j), dx);
        double y3 = cubic(Z(imoins1, j + 1), Z(i, j + 1), Z(iplus1, j + 1), Z(iplus2, j + 1), dx);
        double y4 = cubic(Z(imoins1, j + 2), Z(i, j + 2), Z(iplus1, j + 2), Z(iplus2, j + 2), dx);
        zo[n] = (float)cubic(y1, y2, y3, y4, dy);
    }
}

/* src/interp/ez_rgdint_3_w.inc:20-108 */
void ez_rgdint_3_w_(float *zo, float *px, float *py, fint *npts, float *z, fint *ni, fint *j1, fint *j2, fint *wrap) {
    rgdint_3_w_body(zo, px, py, *npts, z, *ni, *j1, *j2, *wrap, 0);
}

/* src/interp/ez_rgdint_3_wnnc.inc:20-107 */
void ez_rgdint_3_wnnc_(float *zo, float *px, float *py, fint *npts, float *z, fint *ni, fint *j1, fint *j2, fint *wrap) {
    rgdint_3_w_body(zo, px, py, *npts, z, *ni, *j1, *j2, *wrap, 1);
}

/* ------------------------------------------------------------------ */
/* Irregular-grid interpolation kernels                                */
/* ------------------------------------------------------------------ */

/* src/interp/ez_irgdint_1_nw.inc:20-50 ; z(ni,nj), ay(nj) 1-based */
void ez_irgdint_1_nw_(float *zo, float *px, float *py, fint *npts, float *ax, float *ay, float *z, fint *ni_, fint *nj_) {
    int ni = *ni_, nj = *nj_;
    const int j1 = 1;
    for (int n = 0; n < *npts; n++) {
        int i = imin(ni - 1, imax(1, f_ifix(px[n])));
        int j = imin(nj - 1, imax(1, f_ifix(py[n])));
        double x1 = ax[i - 1];
        double x2 = ax[i];
        double x = (double)ax[i - 1] + (x2 - x1) * (double)(px[n] - (float)i);
        double y = (double)(ay[j - 1] + (ay[j] - ay[j - 1]) * (py[n] - (float)j));
        double dx = (x - x1) / (x2 - x1);
        double dy = (y - (double)ay[j - 1]) / (double)(ay[j] - ay[j - 1]);
        double y1 = zlin((double)Z(i, j), (double)Z(i + 1, j), dx);
        double y2 = zlin((double)Z(i, j + 1), (double)Z(i + 1, j + 1), dx);
        zo[n] = (float)zlin(y1, y2, dy);
    }
}

/* src/interp/ez_irgdint_1_w.inc:20-64 ; ay(j1:j2) */
void ez_irgdint_1_w_(float *zo, float *px, float *py, fint *npts, float *ax, float *ay, float *z,
                     fint *ni_, fint *j1_, fint *j2_, fint *wrap_) {
    int ni = *ni_, j1 = *j1_, j2 = *j2_, wrap = *wrap_;
#define AY(j) ay[(j)-j1]
    double x2 = 0.0; /* Fortran leaves x2 untouched (stale) when iplus1 > ni and no wrap */
    for (int n = 0; n < *npts; n++) {
        int limite = ni + 2 - wrap;
        int i = imin(ni - 2 + wrap, imax(1, f_ifix(px[n])));
        int j = imin(j2 - 1, imax(j1 + 1, f_ifix(py[n])));
        if (j < 0) j = j - 1;
        int iplus1 = i + 1;
        double x1 = ax[i - 1];
        if (iplus1 <= ni) x2 = ax[iplus1 - 1];
        if (wrap > 0 && (i == (ni - 2 + wrap))) {
            iplus1 = (limite + i + 1) % limite;
            x2 = (double)(ax[1] + ax[ni - 1]);
        }
        double x = x1 + (x2 - x1) * (double)(px[n] - (float)i);
        double y = (double)(AY(j) + (AY(j + 1) - AY(j)) * (py[n] - (float)j));
        double dx = (x - x1) / (x2 - x1);
        double dy = (y - (double)AY(j)) / (double)(AY(j + 1) - AY(j));
        double y1 = zlin((double)Z(i, j), (double)Z(iplus1, j), dx);
        double y2 = zlin((double)Z(i, j + 1), (double)Z(iplus1, j + 1), dx);
        zo[n] = (float)zlin(y1, y2, dy);
    }
#undef AY
}

/* Newton divided-difference statement functions (ez_irgdint_3_nw.inc:84-87 et al.), fp64 form */
static inline double fa_(double a1, double a2, double a3, double a4, double x, double x1, double x2, double x3) {
    return a1 + (x - x1) * (a2 + (x - x2) * (a3 + a4 * (x - x3)));
}
static inline double fa2_(double c1, double a1, double a2) { return c1 * (a2 - a1); }
static inline double fa3_(double c1, double c2, double c3, double a1, double a2, double a3) {
    return c2 * (c3 * (a3 - a2) - c1 * (a2 - a1));
}
static inline double fa4_(double c1, double c2, double c3, double c4, double c5, double c6,
                          double a1, double a2, double a3, double a4) {
    return c4 * (c5 * (c6 * (a4 - a3) - c3 * (a3 - a2)) - c2 * (c3 * (a3 - a2) - c1 * (a2 - a1)));
}

/* src/interp/ez_irgdint_3_nw.inc:20-168 ; fa..fa4 are declared REAL (:32): each result rounds to fp32 */
void ez_irgdint_3_nw_(float *zo, float *px, float *py, fint *npts, float *ax, float *ay, float *cx, float *cy,
                      float *z, fint *i1_, fint *i2_, fint *j1_, fint *j2_) {
    int i1 = *i1_, i2 = *i2_, j1 = *j1_, j2 = *j2_;
    int ni = i2 - i1 + 1, nyr = j2 - j1 + 1;
#define AX(i) ax[(i)-i1]
#define AY(j) ay[(j)-j1]
#define CX(i, k) ((double)cx[(size_t)((k)-1) * ni + ((i)-i1)])
#define CY(j, k) ((double)cy[(size_t)((k)-1) * nyr + ((j)-j1)])
#define ZZ(i, j) ((double)z[(size_t)((j)-j1) * ni + ((i)-i1)])
#define R4(e) ((double)(float)(e))
    for (int n = 0; n < *npts; n++) {
        int i = imin(i2 - 2, imax(i1 + 1, f_ifix(px[n])));
        int j = imin(j2 - 2, imax(j1 + 1, f_ifix(py[n])));
        double x = (double)(AX(i) + (AX(i + 1) - AX(i)) * (px[n] - (float)i));
        double y = (double)(AY(j) + (AY(j + 1) - AY(j)) * (py[n] - (float)j));
        double x1 = AX(i - 1), x2 = AX(i), x3 = AX(i + 1);
        double y1 = AY(j - 1), y2 = AY(j), y3 = AY(j + 1);
        double b[4];
        for (int r = 0; r < 4; r++) {
            int jj = j - 1 + r;
            double z1 = ZZ(i - 1, jj), z2 = ZZ(i, jj), z3 = ZZ(i + 1, jj), z4 = ZZ(i + 2, jj);
            double a1 = z1;
            double a2 = R4(fa2_(CX(i, 1), z1, z2));
            double a3 = R4(fa3_(CX(i, 1), CX(i, 2), CX(i, 3), z1, z2, z3));
            double a4 = R4(fa4_(CX(i, 1), CX(i, 2), CX(i, 3), CX(i, 4), CX(i, 5), CX(i, 6), z1, z2, z3, z4));
            b[r] = R4(fa_(a1, a2, a3, a4, x, x1, x2, x3));
        }
        double b11 = b[0];
        double b12 = R4(fa2_(CY(j, 1), b[0], b[1]));
        double b13 = R4(fa3_(CY(j, 1), CY(j, 2), CY(j, 3), b[0], b[1], b[2]));
        double b14 = R4(fa4_(CY(j, 1), CY(j, 2), CY(j, 3), CY(j, 4), CY(j, 5), CY(j, 6), b[0], b[1], b[2], b[3]));
        zo[n] = (float)fa_(b11, b12, b13, b14, y, y1, y2, y3);
    }
#undef R4
#undef AX
#undef AY
#undef CX
#undef CY
#undef ZZ
}

/* seam handling shared by ez_irgdint_3_w.inc:97-160 and ez_irgdint_3_wnnc.inc (identical except that
 * wnnc sets iplus2 = 1 instead of 2 for wrap==1, i==ni-1).  The x's are computed in fp32 (`ax(k) +- 360.0`). */
static inline void seam_x(int wrap, int ni, const float *ax, int i, int nnc,
                          int *imoins1, int *iplus1, int *iplus2, float *x1, float *x2, float *x3, float *x4) {
#define AX(k) ax[(k)-1]
    *imoins1 = i - 1;
    *iplus1 = i + 1;
    *iplus2 = i + 2;
    if (wrap == 1 && (i <= 1 || i >= (ni - wrap))) {
        if (i == 1) {
            *imoins1 = ni - 1; *iplus1 = 2; *iplus2 = 3;
            *x1 = AX(ni - 1) - 360.0f; *x2 = AX(1); *x3 = AX(2); *x4 = AX(3);
        }
        if (i == (ni - 1)) {
            *imoins1 = ni - 2; *iplus1 = ni; *iplus2 = nnc ? 1 : 2;
            *x1 = AX(ni - 2); *x2 = AX(ni - 1); *x3 = AX(ni); *x4 = AX(2) + 360.0f;
        }
    } else if (wrap == 2 && (i <= 1 || i > (ni - wrap))) {
        if (i == 1) {
            *imoins1 = ni; *iplus1 = 2; *iplus2 = 3;
            *x1 = AX(ni) - 360.0f; *x2 = AX(1); *x3 = AX(2); *x4 = AX(3);
        }
        if (i == (ni - 1)) {
            *imoins1 = ni - 2; *iplus1 = ni; *iplus2 = 1;
            *x1 = AX(ni - 2); *x2 = AX(ni - 1); *x3 = AX(ni); *x4 = AX(1) + 360.0f;
        }
        if (i == ni) {
            *imoins1 = ni - 1; *iplus1 = 1; *iplus2 = 2;
            *x1 = AX(ni - 1); *x2 = AX(ni); *x3 = AX(1) + 360.0f; *x4 = AX(2) + 360.0f;
        }
    } else {
        *x1 = AX(*imoins1); *x2 = AX(i); *x3 = AX(*iplus1); *x4 = AX(*iplus2);
    }
#undef AX
}

/* src/interp/ez_irgdint_3_w.inc:20-235 ; fa* are REAL*8 (:78), x1..x4,y1..y4 are REAL (:75) */
void ez_irgdint_3_w_(float *zo, float *px, float *py, fint *npts, float *ax, float *ay, float *cx, float *cy,
                     float *z, fint *ni_, fint *j1_, fint *j2_, fint *wrap_) {
    int ni = *ni_, j1 = *j1_, j2 = *j2_, wrap = *wrap_;
    int nyr = j2 - j1 + 1;
#define AY(j) ay[(j)-j1]
#define CX(i, k) ((double)cx[(size_t)((k)-1) * ni + ((i)-1)])
#define CY(j, k) ((double)cy[(size_t)((k)-1) * nyr + ((j)-j1)])
#define ZZ(i, j) ((double)z[(size_t)((j)-j1) * ni + ((i)-1)])
    float x1 = 0, x2 = 0, x3 = 0, x4 = 0; /* persist across points like the Fortran locals */
    for (int n = 0; n < *npts; n++) {
        int i = imin(ni - 2 + wrap, imax(1, imax(2 - wrap, f_ifix(px[n]))));
        int j = imin(j2 - 2, imax(j1 + 1, f_ifix(py[n])));
        int imoins1, iplus1, iplus2;
        seam_x(wrap, ni, ax, i, 0, &imoins1, &iplus1, &iplus2, &x1, &x2, &x3, &x4);
        double x = (double)(x2 + (x3 - x2) * (px[n] - (float)i));
        double y = (double)(AY(j) + (AY(j + 1) - AY(j)) * (py[n] - (float)j));
        float y1 = AY(j - 1), y2 = AY(j), y3 = AY(j + 1);
        double b[4];
        for (int r = 0; r < 4; r++) {
            int jj = j - 1 + r;
            double z1 = ZZ(imoins1, jj), z2 = ZZ(i, jj), z3 = ZZ(iplus1, jj), z4 = ZZ(iplus2, jj);
            double a1 = z1;
            double a2 = fa2_(CX(i, 1), z1, z2);
            double a3 = fa3_(CX(i, 1), CX(i, 2), CX(i, 3), z1, z2, z3);
            double a4 = fa4_(CX(i, 1), CX(i, 2), CX(i, 3), CX(i, 4), CX(i, 5), CX(i, 6), z1, z2, z3, z4);
            b[r] = fa_(a1, a2, a3, a4, x, x1, x2, x3);
        }
        double b11 = b[0];
        double b12 = fa2_(CY(j, 1), b[0], b[1]);
        double b13 = fa3_(CY(j, 1), CY(j, 2), CY(j, 3), b[0], b[1], b[2]);
        double b14 = fa4_(CY(j, 1), CY(j, 2), CY(j, 3), CY(j, 4), CY(j, 5), CY(j, 6), b[0], b[1], b[2], b[3]);
        zo[n] = (float)fa_(b11, b12, b13, b14, y, y1, y2, y3);
    }
#undef CX
#undef CY
}

/* src/interp/ez_irgdint_3_wnnc.inc:20-246 ; x1..x4,y1..y4,cx,cy REAL*8, coefficients computed on the fly */
void ez_irgdint_3_wnnc_(float *zo, float *px, float *py, fint *npts, float *ax, float *ay, float *z,
                        fint *ni_, fint *j1_, fint *j2_, fint *wrap_) {
    int ni = *ni_, j1 = *j1_, j2 = *j2_, wrap = *wrap_;
    float fx1 = 0, fx2 = 0, fx3 = 0, fx4 = 0;
    for (int n = 0; n < *npts; n++) {
        int i = imin(ni - 2 + wrap, imax(1, imax(2 - wrap, f_ifix(px[n]))));
        int j = imin(j2 - 2, imax(j1 + 1, f_ifix(py[n])));
        int imoins1, iplus1, iplus2;
        /* each seam value is an fp32 expression (`ax(k) +- 360.0`) stored into a REAL*8 */
        seam_x(wrap, ni, ax, i, 1, &imoins1, &iplus1, &iplus2, &fx1, &fx2, &fx3, &fx4);
        double x1 = fx1, x2 = fx2, x3 = fx3, x4 = fx4;
        double x = x2 + (x3 - x2) * (double)(px[n] - (float)i);
        double y = (double)(AY(j) + (AY(j + 1) - AY(j)) * (py[n] - (float)j));
        double cx1 = 1.0 / (x2 - x1), cx2 = 1.0 / (x3 - x1), cx3 = 1.0 / (x3 - x2);
        double cx4 = 1.0 / (x4 - x1), cx5 = 1.0 / (x4 - x2), cx6 = 1.0 / (x4 - x3);
        double y1 = AY(j - 1), y2 = AY(j), y3 = AY(j + 1), y4 = AY(j + 2);
        double b[4];
        for (int r = 0; r < 4; r++) {
            int jj = j - 1 + r;
            double z1 = ZZ(imoins1, jj), z2 = ZZ(i, jj), z3 = ZZ(iplus1, jj), z4 = ZZ(iplus2, jj);
            double a1 = z1;
            double a2 = fa2_(cx1, z1, z2);
            double a3 = fa3_(cx1, cx2, cx3, z1, z2, z3);
            double a4 = fa4_(cx1, cx2, cx3, cx4, cx5, cx6, z1, z2, z3, z4);
            b[r] = fa_(a1, a2, a3, a4, x, x1, x2, x3);
        }
        double cy1 = 1.0 / (y2 - y1), cy2 = 1.0 / (y3 - y1), cy3 = 1.0 / (y3 - y2);
        double cy4 = 1.0 / (y4 - y1), cy5 = 1.0 / (y4 - y2), cy6 = 1.0 / (y4 - y3);
        double b11 = b[0];
        double b12 = fa2_(cy1, b[0], b[1]);
        double b13 = fa3_(cy1, cy2, cy3, b[0], b[1], b[2]);
        double b14 = fa4_(cy1, cy2, cy3, cy4, cy5, cy6, b[0], b[1], b[2], b[3]);
        zo[n] = (float)fa_(b11, b12, b13, b14, y, y1, y2, y3);
    }
#undef AY
#undef ZZ
}
#undef Z

/* ------------------------------------------------------------------ */
/* pole / zone helpers                                                 */
/* ------------------------------------------------------------------ */

/* src/interp/ez_aminmax.inc:21-28 */
void ez_aminmax_(float *fldmin, float *fldmax, float *fld, fint *ni, fint *nj) {
    size_t n = (size_t)*ni * (size_t)*nj;
    float mn = fld[0], mx = fld[0];
    for (size_t k = 1; k < n; k++) {
        if (fld[k] < mn) mn = fld[k];
        if (fld[k] > mx) mx = fld[k];
    }
    *fldmin = mn;
    *fldmax = mx;
}

/* src/interp/ez_calcpoleval.inc:21-48 -- sequential fp32 accumulation */
void ez_calcpoleval_(float *poleval, float *z, fint *ni, float *ax, char *grtyp, char *grref, F2Cl l1, F2Cl l2) {
    (void)l1; (void)l2;
    if (grtyp[0] == 'Z' && grref[0] == 'E') {
        float p = 0.0f;
        for (int i = 1; i <= *ni - 1; i++) p = p + z[i - 1] * (ax[i] - ax[i - 1]);
        if ((ax[*ni - 1] - ax[0]) != 0.0f) p = p / (ax[*ni - 1] - ax[0]);
        *poleval = p;
        return;
    }
    float p = 0.0f;
    for (int i = 1; i <= *ni; i++) p = p + z[i - 1];
    *poleval = p / (1.0f * (float)*ni);
}

/* src/interp/ez_fillnpole.inc:20-37 */
void ez_fillnpole_(float *zout, float *zin, fint *ni_, fint *j1_, fint *j2_, float *valpole) {
    int ni = *ni_, j1 = *j1_, j2 = *j2_;
    for (int j = 1; j <= 3; j++)
        for (int i = 0; i < ni; i++) zout[(size_t)(j - 1) * ni + i] = zin[(size_t)(j2 - 3 + j - j1) * ni + i];
    for (int i = 0; i < ni; i++) zout[(size_t)3 * ni + i] = *valpole;
}

/* src/interp/ez_fillspole.inc:20-38 */
void ez_fillspole_(float *zout, float *zin, fint *ni_, fint *j1_, fint *j2_, float *valpole) {
    int ni = *ni_, j1 = *j1_;
    (void)j2_;
    for (int j = j1; j <= j1 + 2; j++)
        for (int i = 0; i < ni; i++) zout[(size_t)(j - j1 + 1) * ni + i] = zin[(size_t)(j - j1) * ni + i];
    for (int i = 0; i < ni; i++) zout[i] = *valpole;
}

/* src/interp/ez_corrbgd.inc:20-54 */
void ez_corrbgd_(float *zout, fint *ni_, fint *nj_, fint *hem) {
    int ni = *ni_, nj = *nj_;
    if (*hem == 0 || *hem == 2) {
        float somme = 0.f;
        for (int i = 0; i < ni; i++) somme = somme + zout[i];
        float moyenne = somme / ((float)ni * 1.0f);
        for (int i = 0; i < ni; i++) zout[i] = moyenne;
    }
    if (*hem == 0 || *hem == 1) {
        float somme = 0.f;
        float *row = zout + (size_t)(nj - 1) * ni;
        for (int i = 0; i < ni; i++) somme = somme + row[i];
        float moyenne = somme / ((float)ni * 1.0f);
        for (int i = 0; i < ni; i++) row[i] = moyenne;
    }
}

/* src/interp/ez_xpngdag2.inc:20-70 ; zout(ni,j1:j2), zi(ni,nj) */
void ez_xpngdag2_(float *zout, float *zi, fint *ni_, fint *nj_, fint *j1_, fint *j2_, fint *hem, fint *symetrie) {
    int ni = *ni_, nj = *nj_, j1 = *j1_;
    (void)j2_;
    float sign = (*symetrie == 0) ? -1.0f : 1.0f;
#define ZO(i, j) zout[(size_t)((j)-j1) * ni + (i)]
#define ZI(i, j) zi[(size_t)((j)-1) * ni + (i)]
    if (*hem == NORD_) {
        for (int j = 1; j <= nj; j++)
            for (int i = 0; i < ni; i++) ZO(i, j) = ZI(i, j);
        for (int j = 1; j <= nj; j++)
            for (int i = 0; i < ni; i++) ZO(i, -j + 1) = sign * ZI(i, j);
    }
    if (*hem == SUD_) {
        for (int j = 1; j <= nj; j++)
            for (int i = 0; i < ni; i++) ZO(i, j) = ZI(i, j);
        for (int j = 1; j <= nj; j++)
            for (int i = 0; i < ni; i++) ZO(i, nj + j) = sign * ZI(i, nj - j + 1);
    }
}

/* src/interp/ez_xpngdb2.inc:20-70 */
void ez_xpngdb2_(float *zout, float *zi, fint *ni_, fint *nj_, fint *j1_, fint *j2_, fint *hem, fint *symetrie) {
    int ni = *ni_, nj = *nj_, j1 = *j1_;
    (void)j2_;
    float sign = (*symetrie == 0) ? -1.0f : 1.0f;
    if (*hem == NORD_) {
        for (int j = 1; j <= nj; j++)
            for (int i = 0; i < ni; i++) ZO(i, j) = ZI(i, j);
        for (int j = 2; j <= nj; j++)
            for (int i = 0; i < ni; i++) ZO(i, 2 - j) = sign * ZI(i, j);
    }
    if (*hem == SUD_) {
        for (int j = 1; j <= nj; j++)
            for (int i = 0; i < ni; i++) ZO(i, j) = ZI(i, j);
        for (int j = 2; j <= nj; j++)
            for (int i = 0; i < ni; i++) ZO(i, nj + j - 1) = sign * ZI(i, nj - j + 1);
    }
#undef ZO
#undef ZI
}

/* ------------------------------------------------------------------ */
/* winds                                                               */
/* ------------------------------------------------------------------ */

/* src/interp/ez_llwfgdw.inc:22-170 -- (u,v) on grid -> (speed, meteorological direction); all fp32 */
void ez_llwfgdw_(float *z1, float *z2, float *xlon, fint *li, fint *lj, char *grtyp,
                 fint *ig1, fint *ig2, fint *ig3, fint *ig4, F2Cl l) {
    (void)l;
    size_t n = (size_t)*li * (size_t)*lj;
    char t = grtyp[0];
    float xg1, xg2, xg3, xg4;
    if (t == '!') off_path("ez_llwfgdw('!')");
    if (t == 'N' || t == 'S') {
        cigaxg_(grtyp, &xg1, &xg2, &xg3, &xg4, ig1, ig2, ig3, ig4, 1);
        for (size_t k = 0; k < n; k++) {
            float uu = z1[k], vv = z2[k];
            float spd0 = sqrtf(uu * uu + vv * vv), dir0;
            if (spd0 == 0.0f) {
                dir0 = 0.0f;
            } else if (uu == 0.0f) {
                if (t == 'N')
                    dir0 = (vv >= 0.0f) ? xlon[k] + xg4 - 90.0f : xlon[k] + xg4 + 90.0f;
                else
                    dir0 = (vv >= 0.0f) ? 90.0f - xlon[k] + xg4 : 270.0f - xlon[k] + xg4;
            } else {
                if (t == 'N')
                    dir0 = xlon[k] + xg4 - RDTODG_F * atan2f(vv, uu);
                else
                    dir0 = 180.0f - xlon[k] + xg4 - RDTODG_F * atan2f(vv, uu);
            }
            dir0 = fmodf(fmodf(dir0, 360.0f) + 360.0f, 360.0f);
            z1[k] = spd0;
            z2[k] = dir0;
        }
        return;
    }
    if (t == 'A' || t == 'B' || t == 'G' || t == 'L') {
        for (size_t k = 0; k < n; k++) {
            float uu = z1[k], vv = z2[k];
            float spd0 = sqrtf(uu * uu + vv * vv), dir0;
            if (spd0 == 0.0f) {
                dir0 = 0.0f;
            } else if (uu == 0.0f) {
                dir0 = (vv >= 0.0f) ? 180.0f : 0.0f;
            } else {
                dir0 = 270.0f - RDTODG_F * atan2f(vv, uu);
            }
            dir0 = fmodf(fmodf(dir0, 360.0f) + 360.0f, 360.0f);
            z1[k] = spd0;
            z2[k] = dir0;
        }
        return;
    }
    printf("0 erreur, bad grid type (llwfgdw) - grtyp = %c\n", t);
}

/* src/interp/ez_gdwfllw.inc:23-137 -- (speed, direction) -> (u,v) on grid; all fp32 */
void ez_gdwfllw_(float *z1, float *z2, float *xlon, fint *li, fint *lj, char *grtyp,
                 fint *ig1, fint *ig2, fint *ig3, fint *ig4, F2Cl l) {
    (void)l;
    size_t n = (size_t)*li * (size_t)*lj;
    char t = grtyp[0];
    float xg1, xg2, xg3, xg4;
    if (t == '!') off_path("ez_gdwfllw('!')");
    if (t == 'N' || t == 'S') {
        cigaxg_(grtyp, &xg1, &xg2, &xg3, &xg4, ig1, ig2, ig3, ig4, 1);
        for (size_t k = 0; k < n; k++) {
            float psi = (t == 'N') ? xlon[k] + xg4 - z2[k] : 180.0f - xlon[k] + xg4 - z2[k];
            float u = cosf(psi * DGTORD_F) * z1[k];
            float v = sinf(psi * DGTORD_F) * z1[k];
            z1[k] = u;
            z2[k] = v;
        }
        return;
    }
    if (t == 'A' || t == 'B' || t == 'G' || t == 'L') {
        for (size_t k = 0; k < n; k++) {
            float psi = 270.0f - z2[k];
            float u = cosf(psi * DGTORD_F) * z1[k];
            float v = sinf(psi * DGTORD_F) * z1[k];
            z1[k] = u;
            z2[k] = v;
        }
        return;
    }
    /* any other grtyp: silently returns (ez_gdwfllw.inc:135-136) */
}

/* src/interp/ez_uvacart.inc:24-62 */
void ez_uvacart_(float *xyz, float *u, float *v, float *lon, float *lat, fint *ni, fint *nj) {
    double dar = (double)(acosf(-1.f) / 180.f);
    size_t n = (size_t)*ni * (size_t)*nj;
    for (size_t k = 0; k < n; k++) {
        double a = sin(dar * (double)lon[k]);
        double b = cos(dar * (double)lon[k]);
        double c = sin(dar * (double)lat[k]);
        double d = cos(dar * (double)lat[k]);
        xyz[3 * k + 0] = (float)(-((double)u[k] * a) - ((double)v[k] * b * c));
        xyz[3 * k + 1] = (float)(((double)u[k] * b) - ((double)v[k] * a * c));
        xyz[3 * k + 2] = (float)((double)v[k] * d);
    }
}

/* src/interp/ez_cartauv.inc:23-62 */
void ez_cartauv_(float *u, float *v, float *uvcart, float *lon, float *lat, fint *ni, fint *nj) {
    double dar = (double)(acosf(-1.f) / 180.f);
    size_t n = (size_t)*ni * (size_t)*nj;
    for (size_t k = 0; k < n; k++) {
        double a = cos(dar * (double)lon[k]);
        double b = sin(dar * (double)lon[k]);
        double e = cos(dar * (double)lat[k]);
        double f = sin(dar * (double)lat[k]);
        double c1 = uvcart[3 * k + 0], c2 = uvcart[3 * k + 1];
        u[k] = (float)((c2 * a) - (c1 * b));
        double c = (c1 * a) + (c2 * b);
        float w3 = uvcart[3 * k + 2];
        /* UVCART(3,K)**2 is an fp32 square, then widened for the fp64 sum */
        double d = sqrt(c * c + (double)(w3 * w3));
        v[k] = (float)d_sign(d, ((double)w3 * e) - (c * f));
    }
}

/* ------------------------------------------------------------------ */
/* integer mask transfer                                               */
/* ------------------------------------------------------------------ */

extern fint ezgetopt_(char *option, char *value, F2Cl lenoption, F2Cl lenvalue) __attribute__((weak));

/* src/interp/ezget_mask_zones.inc:61-105.  mask_in(ix+1,iy) etc. are plain column-major address
 * arithmetic in Fortran (no bounds checks); a neighbour that falls outside the array's storage is treated
 * here as non-zero instead of being read (the reference's behaviour there is undefined). */
void qqq_ezsint_mask_(fint *mask_out, float *x, float *y, fint *ni_out_, fint *nj_out_,
                      fint *mask_in, fint *ni_in_, fint *nj_in_) {
    int ni_out = *ni_out_, nj_out = *nj_out_, ni_in = *ni_in_, nj_in = *nj_in_;
    char value[33];
    memset(value, ' ', 32);
    value[32] = 0;
    char opt[] = "cloud_interp_alg";
    if (ezgetopt_) ezgetopt_(opt, value, 16, 32);
    size_t nin = (size_t)ni_in * nj_in;
    for (int j = 0; j < nj_out; j++) {
        for (int i = 0; i < ni_out; i++) {
            size_t k = (size_t)j * ni_out + i;
            mask_out[k] = 1;
            int ix = (int)x[k], iy = (int)y[k];
            int nix = f_nint(x[k]), niy = f_nint(y[k]);
            if (ix < 1 || ix > ni_in || iy < 1 || iy > nj_in) {
                mask_out[k] = 0;
            } else {
                size_t m = (size_t)(niy - 1) * ni_in + (nix - 1);
                if (m < nin && mask_in[m] == 0) mask_out[k] = 0;
            }
        }
    }
    if (strncmp(value, "linear", 6) == 0) {
        for (int j = 0; j < nj_out - 1; j++) {
            for (int i = 0; i < ni_out - 1; i++) {
                size_t k = (size_t)j * ni_out + i;
                if (mask_out[k] == 1) {
                    int ix = (int)x[k], iy = (int)y[k];
                    size_t m1 = (size_t)(iy - 1) * ni_in + ix;       /* (ix+1, iy)   */
                    size_t m2 = (size_t)iy * ni_in + (ix - 1);       /* (ix,   iy+1) */
                    size_t m3 = (size_t)iy * ni_in + ix;             /* (ix+1, iy+1) */
                    if ((m1 < nin && mask_in[m1] == 0) || (m2 < nin && mask_in[m2] == 0) || (m3 < nin && mask_in[m3] == 0))
                        mask_out[k] = 0;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* off-path entry points: present so libezref.so links, abort if reached */
/* ------------------------------------------------------------------ */

/* ------------------------------------------------------------------ */
/* 'Y' cloud SOURCES: ez_calcxy_y (k nearest + inverse-square-distance weights), ez_applywgts */
/* ------------------------------------------------------------------ */
/* src/interp/ez_calcxy_y.inc:1-94.  All arithmetic is default REAL (fp32).  x, y are never written (the call to
 * inside_or_outside is commented out, :60-63).  tmp_idxs(16,2) is write-only.  maxloc(tmpwgts, 1) = FIRST slot of the maximum. */
void ez_calcxy_y_(float *wts, fint *idxs, float *x, float *y, float *gdout_lat, float *gdout_lon, float *gdin_lat, float *gdin_lon,
                  fint *masque, fint *ni_src_, fint *nj_src_, fint *ni_dst_, fint *nj_dst_, fint *num_wts_) {
    (void)x; (void)y;
    const long nsrc = (long)*ni_src_ * *nj_src_, ndst = (long)*ni_dst_ * *nj_dst_;
    const int ni_src = *ni_src_, nw = *num_wts_;
    float lon_min = gdin_lon[0], lon_max = gdin_lon[0], lat_min = gdin_lat[0], lat_max = gdin_lat[0];
    for (long k = 1; k < nsrc; k++) {
        if (gdin_lon[k] < lon_min) lon_min = gdin_lon[k];
        if (gdin_lon[k] > lon_max) lon_max = gdin_lon[k];
        if (gdin_lat[k] < lat_min) lat_min = gdin_lat[k];
        if (gdin_lat[k] > lat_max) lat_max = gdin_lat[k];
    }
    for (long k = 0; k < ndst; k++) {
        masque[k] = 1;
        if (gdout_lon[k] < lon_min || gdout_lon[k] > lon_max || gdout_lat[k] < lat_min || gdout_lat[k] > lat_max) masque[k] = 0;
    }
    for (long k = 0; k < ndst * nw; k++) { wts[k] = 1.0e30f; idxs[k] = 0; }
    float *tmpwgts = (float *)malloc(sizeof(float) * (size_t)nw);
    for (long k = 0; k < ndst; k++) {
        if (masque[k] == 0) continue;
        for (int n = 0; n < nw; n++) tmpwgts[n] = 1.0e30f;
        int locmax = 0;
        int point_found = 0;
        const float lo = gdout_lon[k], la = gdout_lat[k];
        for (long s = 0; s < nsrc; s++) {   /* jj outer, ii inner == increasing linear index */
            float rx = lo - gdin_lon[s];
            const float ry = la - gdin_lat[s];
            if (fabsf(rx) > 180.0f) rx = fabsf(fabsf(rx) - 360.0f);
            if (fabsf(ry) < 10.0f && fabsf(rx) < 10.0f) {
                const float dist = rx * rx + ry * ry;
                point_found = 1;
                if (dist < tmpwgts[locmax]) {
                    tmpwgts[locmax] = dist;
                    idxs[k + (long)locmax * ndst] = (fint)(s + 1);   /* ii + (jj-1)*ni_src */
                    locmax = 0;
                    for (int n = 1; n < nw; n++)
                        if (tmpwgts[n] > tmpwgts[locmax]) locmax = n;
                }
            }
        }
        (void)ni_src;
        if (!point_found) masque[k] = 0;
        if (masque[k] == 1) {
            for (int n = 0; n < nw; n++) wts[k + (long)n * ndst] = tmpwgts[n];
            if (wts[k] > 6371000.0f) {
                for (int n = 0; n < nw; n++) wts[k + (long)n * ndst] = 1.0e30f;
                masque[k] = 0;
            } else {
                float total_wgt = 0.0f;
                for (int n = 0; n < nw; n++) {
                    float w = wts[k + (long)n * ndst];
                    w = fmaxf(w, 1.0e-10f);
                    wts[k + (long)n * ndst] = 1.0f / w;
                }
                for (int n = 0; n < nw; n++) total_wgt = total_wgt + wts[k + (long)n * ndst];
                for (int n = 0; n < nw; n++) wts[k + (long)n * ndst] = wts[k + (long)n * ndst] / total_wgt;
            }
        }
    }
    free(tmpwgts);
}

/* the reference's own C option getters (src/interp/ezgetopt.c, ezgetval.c), present when linked into libezref.so */
extern int32_t c_ezgetopt(char *option, char *value) __attribute__((weak));
extern int32_t c_ezgetval(char *option, float *value) __attribute__((weak));

/* src/interp/ez_applywgts.inc:1-48, quirks kept:
 *   - outfld is DECLARED with ni_src*nj_src elements (:7): the whole-array assignments `outfld = xtrapval` / `outfld = rmin`
 *     (:27,:33) fill only that many; target points beyond keep what the caller's array held when their mask is 0;
 *   - j = 1 + k / ni_dst (:38) is one row too far for the last column (k a multiple of ni_dst): those points read the
 *     weight / index lists of linear element k - 1 + ni_dst of the same plane (running into the next plane on the last row). */
void ez_applywgts_(float *outfld, float *wts, fint *idxs, float *infld, float *x, float *y, fint *masque, fint *ni_src_, fint *nj_src_,
                   fint *ni_dst_, fint *nj_dst_, fint *n_wts_) {
    (void)x; (void)y;
    const long nsrc = (long)*ni_src_ * *nj_src_, ndst = (long)*ni_dst_ * *nj_dst_;
    const int ni_dst = *ni_dst_, nw = *n_wts_;
    char xtrapopt[64] = "maximum";
    float xtrapval = 0.0f;
    if (c_ezgetopt) { char o[] = "extrap_degree"; c_ezgetopt(o, xtrapopt); }
    if (c_ezgetval) { char o[] = "extrap_value"; c_ezgetval(o, &xtrapval); }
    float fill;
    if (strncmp(xtrapopt, "value", 5) == 0) {
        fill = xtrapval;
    } else {
        float rmin = infld[0], rmax = infld[0];
        for (long k = 1; k < nsrc; k++) {
            if (infld[k] < rmin) rmin = infld[k];
            if (infld[k] > rmax) rmax = infld[k];
        }
        rmin = rmin - 0.1f * (rmax - rmin);
        fill = rmin;
    }
    /* (sic) the declared extent, not the target's.  With more source than target points the reference writes past the end of
     * the caller's output array; the restatement stops at the target's size instead of corrupting the heap. */
    for (long k = 0; k < (nsrc < ndst ? nsrc : ndst); k++) outfld[k] = fill;
    for (long k = 1; k <= ndst; k++) {
        const long i = ((k - 1) % ni_dst) + 1, j = 1 + k / ni_dst;
        if (masque[k - 1] == 1) {
            outfld[k - 1] = 0.0f;
            for (int n = 1; n <= nw; n++) {
                const long e = (i - 1) + (j - 1) * (long)ni_dst + (long)(n - 1) * ndst;
                if (e >= ndst * (long)nw) break;   /* (past the end of the arrays: undefined in the reference) */
                if (idxs[e] < 1) break;
                outfld[k - 1] = outfld[k - 1] + wts[e] * infld[idxs[e] - 1];
            }
        }
    }
}
void ez_calcxy_y_m_(void) { off_path("ez_calcxy_y_m"); }
/* ------------------------------------------------------------------------------------------------
 * ez_avg.inc:1-170 and ez_avg_sph.inc:1-232 -- area-weighted box average of the source cells a target cell covers
 * (interp_degree = average / sph_average, gdinterp.c:181-203,283-306).  One routine restates both: they differ only
 * in the x extent of a target cell, widened by 1 / cos(latitude of the target row) in the spherical variant
 * (ez_avg_sph.inc:65-90).  The target cell of point (i, j) runs half-way to its neighbours in the source index space
 * taken from the FIRST row (x) and FIRST column (y) of the gridset coordinates (ez_avg.inc:55-56): the routine assumes
 * axis-aligned grids.  All arithmetic is default real (fp32), sums in loop order.  ztmp(ii, jj) of the reference is the
 * field replicated in x according to `extension` (ez_avg.inc:20-52); an index outside what the reference allocates and
 * fills is undefined there and reads as 0 here. */
static float avg_ztmp(const float *zin, int ni, int nj, int ext, int ii, int jj) {
    if (jj < 1 || jj > nj) return 0.0f;
    int i;
    if (ext == 0) {
        if (ii < 1 || ii > ni) return 0.0f;
        i = ii;
    } else if (ext == 1) {   /* filled: 2-ni .. 2ni-2 from zin(1 .. ni-1) */
        if (ii < 2 - ni || ii > 2 * ni - 2) return 0.0f;
        i = ii >= ni ? ii - ni + 1 : (ii >= 1 ? ii : ii + ni - 1);
    } else {                 /* filled: 1-ni .. 2ni from zin(1 .. ni) */
        if (ii < 1 - ni || ii > 2 * ni) return 0.0f;
        i = ii > ni ? ii - ni : (ii >= 1 ? ii : ii + ni);
    }
    return zin[(long)(jj - 1) * ni + (i - 1)];
}
static void avg_common(float *zout, const float *xx, const float *yy, const float *lats_dst, int ni_dst, int nj_dst, const float *zin,
                       int ni_src, int nj_src, int ext) {
    const float degre_a_radian = 0.017453295199f;
    float *x = (float *)malloc(sizeof(float) * ni_dst), *y = (float *)malloc(sizeof(float) * nj_dst);
    float *y_low = (float *)malloc(sizeof(float) * nj_dst), *y_high = (float *)malloc(sizeof(float) * nj_dst);
    for (int i = 0; i < ni_dst; i++) x[i] = xx[i];                       /* x = xx(:,1) */
    for (int j = 0; j < nj_dst; j++) y[j] = yy[(long)j * ni_dst];        /* y = yy(1,:) */
    if (x[0] > (float)(ni_src - 1)) x[0] = 1.0f;
    y_low[0] = fmaxf(1.0f, y[0] - 0.5f * (y[1] - y[0]));
    for (int j = 1; j < nj_dst; j++) y_low[j] = y[j] - 0.5f * (y[j] - y[j - 1]);
    y_high[nj_dst - 1] = fminf(1.0f * nj_src, y[nj_dst - 1] + 0.5f * (y[nj_dst - 1] - y[nj_dst - 2]));
    for (int j = 0; j < nj_dst - 1; j++) y_high[j] = y[j] + 0.5f * (y[j + 1] - y[j]);
    for (int j = 1; j <= nj_dst; j++) {
        float amp = 1.0f;
        const int sph = lats_dst != NULL;
        if (sph) amp = 1.0f / cosf(lats_dst[j - 1] * degre_a_radian);
        const float yl = y_low[j - 1], yh = y_high[j - 1];
        int jstart, jend;
        if (j == 1) {                      /* "Moyenne 1e rangee" */
            jstart = f_ifix(yl); jend = f_nint(yh);
            if ((float)jstart > yl) jstart = jstart - 1;
        } else if (j == nj_dst) {          /* "Moyenne rangee du haut" */
            jstart = f_ifix(y[nj_dst - 1]); jend = f_nint(yh);
        } else {
            jstart = f_ifix(yl); jend = f_nint(yh);
            if ((0.5f + (float)jstart) < yl) jstart = jstart + 1;
        }
        for (int i = 1; i <= ni_dst; i++) {
            float xl, xh;
            if (i == 1) xl = sph ? x[0] - 0.5f * (x[1] - x[0]) * amp : x[0] - 0.5f * (x[1] - x[0]);
            else xl = sph ? x[i - 1] - 0.5f * (x[i - 1] - x[i - 2]) * amp : x[i - 1] - 0.5f * (x[i - 1] - x[i - 2]);
            if (i == ni_dst) xh = sph ? x[i - 1] + 0.5f * (x[i - 1] - x[i - 2]) * amp : x[i - 1] + 0.5f * (x[i - 1] - x[i - 2]);
            else xh = sph ? x[i - 1] + 0.5f * (x[i] - x[i - 1]) * amp : x[i - 1] + 0.5f * (x[i] - x[i - 1]);
            float acc = 0.0f, total_area = 0.0f;
            int istart = f_ifix(xl), iend = f_nint(xh);
            if ((0.5f + (float)istart) < xl) istart = istart + 1;
            if (j == 1 && (float)iend < xh) iend = iend + 1;
            for (int jj = jstart; jj <= jend; jj++) {
                float ytile_min = (float)jj - 0.5f, ytile_max = (float)jj + 0.5f, yfrac = 1.0f;
                if (j == 1) {
                    if (jj == 1) ytile_min = 1.0f;
                    yfrac = ytile_max - ytile_min;
                } else if (j == nj_dst) {
                    if (jj == nj_src) ytile_max = (float)nj_src;
                    yfrac = ytile_max - ytile_min;
                }
                if (ytile_min < yl) yfrac = ytile_max - yl;
                if (ytile_max > yh) yfrac = yh - ytile_min;
                for (int ii = istart; ii <= iend; ii++) {
                    const float xtile_min = (float)ii - 0.5f, xtile_max = (float)ii + 0.5f;
                    float xfrac = 1.0f;
                    if (xtile_min < xl) xfrac = xtile_max - xl;
                    if (xtile_max > xh) xfrac = xh - xtile_min;
                    const float area = xfrac * yfrac;
                    total_area = total_area + area;
                    acc = acc + avg_ztmp(zin, ni_src, nj_src, ext, ii, jj) * area;
                }
            }
            if (total_area != 0.0f) acc = acc / total_area;
            zout[(long)(j - 1) * ni_dst + (i - 1)] = acc;
        }
    }
    free(x); free(y); free(y_low); free(y_high);
}
void ez_avg_(float *zout, float *xx, float *yy, fint *ni_dst, fint *nj_dst, float *zin, fint *ni_src, fint *nj_src, fint *extension) {
    avg_common(zout, xx, yy, NULL, *ni_dst, *nj_dst, zin, *ni_src, *nj_src, *extension);
}
void ez_avg_sph_(float *zout, float *xx, float *yy, float *lats_dst, fint *ni_dst, fint *nj_dst, float *zin, fint *ni_src, fint *nj_src,
                 fint *extension) {
    avg_common(zout, xx, yy, lats_dst, *ni_dst, *nj_dst, zin, *ni_src, *nj_src, *extension);
}
void ez_genpole_(void) { off_path("ez_genpole"); }
void ez_llflamb_(void) { off_path("ez_llflamb"); }
void ez_vtllfxy_(void) { off_path("ez_vtllfxy"); }
/* lorenzo_mask_fill.inc:1-81 (c_ezsint_mdm / c_ezuvint_mdm, ez_mask.c:127-181, call it with methode = 2) */
void lorenzo_mask_fill_(float *fld, fint *masque, fint *pni, fint *pnj, fint *methode) {
    const int ni = *pni, nj = *pnj;
#define FLD(i, j) fld[(long)((j) - 1) * ni + ((i) - 1)]
#define MSK(i, j) masque[(long)((j) - 1) * ni + ((i) - 1)]
    float rmin = fld[0], rmax = fld[0];
    for (long k = 1; k < (long)ni * nj; k++) { if (fld[k] < rmin) rmin = fld[k]; if (fld[k] > rmax) rmax = fld[k]; }
    switch (*methode) {
        case 1:
            for (int j = 2; j <= nj; j++)
                for (int i = 2; i <= ni; i++)
                    if (MSK(i, j) == 0) {
                        FLD(i, j) = FLD(i - 1, j) + FLD(i, j - 1) - FLD(i - 1, j - 1);
                        if (FLD(i, j) < rmin) FLD(i, j) = rmin;
                        if (FLD(i, j) > rmax) FLD(i, j) = rmax;
                    }
            break;
        case 2:
            for (int j = 1; j <= nj; j++)
                for (int i = 1; i <= ni; i++)
                    if (MSK(i, j) == 0) FLD(i, j) = rmin;
            break;
        case 3:
            for (int j = 1; j <= nj; j++) {
                int i = 1, last_i = 1;
                if (MSK(i, j) == 0) {
                    for (;;) {
                        i = i + 1;
                        if (MSK(i, j) != 0) break;
                        if (i == ni) { for (int k = 1; k <= ni; k++) FLD(k, j) = rmin; break; }
                    }
                    if (i < ni) for (int ii = last_i; ii <= i - 1; ii++) FLD(ii, j) = FLD(i, j);
                }
                for (;;) {
                    i = i + 1;
                    if (i >= ni + 1) break;
                    if (MSK(i, j) == 0) FLD(i, j) = FLD(last_i, j);
                    last_i = i;
                }
            }
            break;
    }
#undef FLD
#undef MSK
}
void qqq_ezget_mask_zones_(void) { off_path("qqq_ezget_mask_zones"); }
/* FST access of ez_lireEnrPositionnels.c (c_fstinf / c_fstsui / c_fstprm, f77 fstluk / fstprm): the reference's file layer
 * (src/fstd98) is not part of the oracle; tests hand in an in-memory record table through oracle_set_fst_hooks (same C
 * signatures as rmn/fstd98.h:42-67) so that the reference's own grid-definition code can be run against it. */
typedef int (*o_fstinf_fn)(int, int *, int *, int *, int, const char *, int, int, int, const char *, const char *);
typedef int (*o_fstsui_fn)(int, int *, int *, int *);
typedef int (*o_fstprm_fn)(int, int *, int *, int *, int *, int *, int *, int *, int *, int *, int *, int *, char *, char *, char *, char *, int *, int *,
                           int *, int *, int *, int *, int *, int *, int *, int *, int *);
typedef int (*o_fstluk_fn)(void *, int, int *, int *, int *);
static o_fstinf_fn h_fstinf; static o_fstsui_fn h_fstsui; static o_fstprm_fn h_fstprm; static o_fstluk_fn h_fstluk;
void oracle_set_fst_hooks(void *inf, void *sui, void *prm, void *luk) {
    h_fstinf = (o_fstinf_fn)inf; h_fstsui = (o_fstsui_fn)sui; h_fstprm = (o_fstprm_fn)prm; h_fstluk = (o_fstluk_fn)luk;
}
int c_fstinf(int iun, int *ni, int *nj, int *nk, int datev, const char *etiket, int ip1, int ip2, int ip3, const char *typvar, const char *nomvar) {
    if (!h_fstinf) { off_path("c_fstinf"); return -1; }
    return h_fstinf(iun, ni, nj, nk, datev, etiket, ip1, ip2, ip3, typvar, nomvar);
}
int c_fstsui(int iun, int *ni, int *nj, int *nk) {
    if (!h_fstsui) { off_path("c_fstsui"); return -1; }
    return h_fstsui(iun, ni, nj, nk);
}
int c_fstprm(int handle, int *dateo, int *deet, int *npas, int *ni, int *nj, int *nk, int *nbits, int *datyp, int *ip1, int *ip2, int *ip3,
             char *typvar, char *nomvar, char *etiket, char *grtyp, int *ig1, int *ig2, int *ig3, int *ig4, int *swa, int *lng, int *dltf,
             int *ubc, int *extra1, int *extra2, int *extra3) {
    if (!h_fstprm) { off_path("c_fstprm"); return -1; }
    return h_fstprm(handle, dateo, deet, npas, ni, nj, nk, nbits, datyp, ip1, ip2, ip3, typvar, nomvar, etiket, grtyp, ig1, ig2, ig3, ig4, swa,
                    lng, dltf, ubc, extra1, extra2, extra3);
}
int fstprm_(int *handle, int *dateo, int *deet, int *npas, int *ni, int *nj, int *nk, int *nbits, int *datyp, int *ip1, int *ip2, int *ip3,
            char *typvar, char *nomvar, char *etiket, char *grtyp, int *ig1, int *ig2, int *ig3, int *ig4, int *swa, int *lng, int *dltf,
            int *ubc, int *extra1, int *extra2, int *extra3, F2Cl l1, F2Cl l2, F2Cl l3, F2Cl l4) {
    (void)l1; (void)l2; (void)l3; (void)l4;
    return c_fstprm(*handle, dateo, deet, npas, ni, nj, nk, nbits, datyp, ip1, ip2, ip3, typvar, nomvar, etiket, grtyp, ig1, ig2, ig3, ig4, swa,
                    lng, dltf, ubc, extra1, extra2, extra3);
}
int fstluk_(void *field, int *handle, int *ni, int *nj, int *nk) {
    if (!h_fstluk) { off_path("fstluk"); return -1; }
    return h_fstluk(field, *handle, ni, nj, nk);
}
