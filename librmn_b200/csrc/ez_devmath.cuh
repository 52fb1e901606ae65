// librmn_b200/csrc/ez_devmath.cuh -- per-point coordinate math shared by the gridset-build and wind kernels.
//
// Compile every translation unit that includes this with -fmad=false: the reference's mixed fp32/fp64
// evaluation sequences are reproduced operation by operation, and indices are
// ifix/nint of the resulting fp32 coordinates, so no multiply-add may be contracted.
//
// Where the reference calls an fp32 libm function (sinf, cosf, asinf, atanf, atan2f, acosf) the device
// evaluates the fp64 function and rounds once: that is the correctly rounded fp32 value in all but ~1e-9
// of cases.  glibc's own fp32 functions are *not* correctly rounded (1-16 % of arguments differ in the
// last bit), so coordinates produced through those paths can differ from a glibc-hosted reference by
// 1 ulp; see DESIGN.md "coordinate parity".
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "ez_types.h"

#define EZ_PIE_F 3.1415926535898f
#define EZ_RDTODG_F 57.295779513082f
#define EZ_DGTORD_F 1.7453292519943e-2f

__device__ __forceinline__ float ez_sinf(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float ez_cosf(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float ez_asinf(float x) { return (float)asin((double)x); }
__device__ __forceinline__ float ez_atanf(float x) { return (float)atan((double)x); }
__device__ __forceinline__ float ez_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ float ez_signf(float a, float b) { return copysignf(fabsf(a), b); }
// acos(-1.0)/180.0 and 180./acos(-1.0) in fp32 (ez_lac.inc:38, ez_cal.inc:37)
__device__ __forceinline__ float ez_dar() { return 3.14159274101257324f / 180.0f; }
__device__ __forceinline__ float ez_rad() { return 180.f / 3.14159274101257324f; }

// ez_lac.inc:39-46
__device__ __forceinline__ void ez_lac1(float lon, float lat, float &c0, float &c1, float &c2) {
    const float dar = ez_dar();
    float cosdar = ez_cosf(dar * lat);
    c0 = cosdar * ez_cosf(dar * lon);
    c1 = cosdar * ez_sinf(dar * lon);
    c2 = ez_sinf(dar * lat);
}
// ez_cal.inc:38-44
__device__ __forceinline__ void ez_cal1(float c0, float c1, float c2, float &lon, float &lat) {
    const float rad = ez_rad();
    lat = ez_asinf(fmaxf(-1.00f, fminf(1.0f, c2))) * rad;
    lon = ez_atan2f(c1, c0) * rad;
    lon = fmodf(lon, 360.0f);
    if (lon < 0.0f) lon = lon + 360.0f;
}
// mxm.F90:30-38 with a 3x3 column-major matrix: c(i) = ((0 + m(i,1) b1) + m(i,2) b2) + m(i,3) b3
__device__ __forceinline__ void ez_mxm3(const float *m, float b0, float b1, float b2, float &c0, float &c1, float &c2) {
    c0 = ((0.0f + m[0] * b0) + m[3] * b1) + m[6] * b2;
    c1 = ((0.0f + m[1] * b0) + m[4] * b1) + m[7] * b2;
    c2 = ((0.0f + m[2] * b0) + m[5] * b1) + m[8] * b2;
}
// ez_gfxyfll.c:44-56: true (lon,lat) -> rotated frame with r ; ez_gfllfxy.c:51-54 uses ri for the inverse
__device__ __forceinline__ void ez_rotate_ll(const float *m, float lon_in, float lat_in, float &lon_out, float &lat_out) {
    float a0, a1, a2, b0, b1, b2;
    ez_lac1(lon_in, lat_in, a0, a1, a2);
    ez_mxm3(m, a0, a1, a2, b0, b1, b2);
    ez_cal1(b0, b1, b2, lon_out, lat_out);
}

// src/base/llfxy.F:64-94 (fp32 throughout)
__device__ __forceinline__ void ez_llfxy(float &dlat, float &dlon, float x, float y, float d60, float dgrw, int nhem) {
    float re = 1.866025f * 6.371e+6f / d60;
    float re2 = re * re;
    dlat = 90.f;
    dlon = 0.f;
    if (!(x == 0.f && y == 0.f)) {
        if (x == 0.f) dlon = ez_signf(90.f, y);
        if (x != 0.f) dlon = ez_atanf(y / x) * EZ_RDTODG_F;
        if (x < 0.f) dlon = dlon + ez_signf(180.f, y);
        dlon = dlon - dgrw;
        if (dlon > 180.f) dlon = dlon - 360.f;
        if (dlon < -180.f) dlon = dlon + 360.f;
        float r2 = x * x + y * y;
        dlat = (re2 - r2) / (re2 + r2);
        dlat = ez_asinf(dlat) * EZ_RDTODG_F;
    }
    if (nhem == 2) dlat = -dlat;
    if (nhem == 2) dlon = -dlon;
}

// ez_vllfxy.inc:64-98 (fp64 core; fp32 subtraction first)
__device__ __forceinline__ void ez_vllfxy1(float &dlat, float &dlon, float x, float y, float d60, float dgrw, float pi, float pj, int nhem) {
    double re = 1.866025 * (double)6.371e+6f / (double)d60;
    double re2 = re * re;
    double x1 = (double)(x - pi);
    double y1 = (double)(y - pj);
    double rlon = 0.0, rlat;
    if (x1 == 0.0) rlon = copysign(90., y1);
    if (x1 != 0.0) rlon = atan(y1 / x1) * (double)EZ_RDTODG_F;
    if (x1 < 0.0) rlon = rlon + copysign(180., y1);
    rlon = rlon - (double)dgrw;
    if (rlon < 0.0) rlon = rlon + 3.6e2;
    double r2 = x1 * x1 + y1 * y1;
    rlat = (re2 - r2) / (re2 + r2);
    rlat = fmax(-1.0, fmin(rlat, 1.0));
    rlat = asin(rlat) * (double)EZ_RDTODG_F;
    if (nhem == 2) {
        rlat = -rlat;
        rlon = -rlon;
        if (rlon < 0.0) rlon = rlon + 360.0;
    }
    dlat = (float)rlat;
    dlon = (float)rlon;
}

// ez_vxyfll.inc:70-94
__device__ __forceinline__ void ez_vxyfll1(float &x, float &y, float dlat, float dlon, float d60, float dgrw, float pi, float pj, int nhem) {
    double re = 1.866025 * 6.371e+6 / (double)d60;
    double rlon, rlat;
    if (nhem == 1) {
        rlon = (double)(EZ_DGTORD_F * (dlon + dgrw));
        rlat = (double)(EZ_DGTORD_F * dlat);
    } else {
        rlon = (double)dlon;
        if (rlon > 180.0) rlon = rlon - 360.0;
        rlon = (double)EZ_DGTORD_F * (-rlon + (double)dgrw);
        rlat = (double)(EZ_DGTORD_F * (-dlat));
    }
    double sinlat = sin(rlat);
    double r = re * sqrt((1.0 - sinlat) / (1.0 + sinlat));
    x = (float)(r * cos(rlon) + (double)pi);
    y = (float)(r * sin(rlon) + (double)pj);
}

// ez_cherche.inc:55-69 -- bisection on an ascending table, returns the lower bracket (1-based)
__device__ __forceinline__ int ez_cherche(float val, const float *__restrict__ tab, int n) {
    int debut = 1, fin = n;
    int milieu = (debut + fin) >> 1;   // (debut+fin)*0.5 truncated; exact for n < 2^24
    while (milieu != debut) {
        if (val <= __ldg(tab + milieu - 1))
            fin = milieu;
        else
            debut = milieu;
        milieu = (debut + fin) >> 1;
    }
    return milieu;
}

// ez_llll2gd.inc:32-50 -- wraps lon in place, then affine map
__device__ __forceinline__ void ez_llll2gd1(float &x, float &y, float lat, float &lon, float xlat0, float xlon0, float dellat, float dellon, float lonref) {
    if (lonref == -180.0f) {
        if (lon > 180.0f) lon = lon - 360.0f;
    } else {
        if (lon < 0.0f) lon = lon + 360.0f;
    }
    x = (lon - xlon0) / dellon + 1.0f;
    y = (lat - xlat0) / dellat + 1.0f;
}

// (lat,lon) -> fractional 1-based (x,y) in the index space of grid g.
// Restates ez_calcxy.c:59-97 + ez_ll2rgd.inc:47-160 + ez_ll2igd.inc:45-85 + gdxyfll.c:141-187.
// `lon` is updated in place exactly where the reference mutates the caller's array.
// `njs` is the length of the ay table searched: gr.nj from ez_calcxy.c:81, gr.j2 from gdxyfll.c:170.
__device__ __forceinline__ void ez_ll2xy(const GridDev &g, int njs, float lat, float &lon, float &x, float &y) {
    switch (g.grtyp) {
        case 'N':
            ez_vxyfll1(x, y, lat, lon, g.xg[2], g.xg[3], g.xg[0], g.xg[1], 1);
            return;
        case 'S':
            ez_vxyfll1(x, y, lat, lon, g.xg[2], g.xg[3], g.xg[0], g.xg[1], 2);
            return;
        case 'A':
        case 'B': {
            // ez_ll2rgd.inc:70-118 ; xyp = (xlat0, xlon0, dellat, dellon) prepared on the host in fp32
            if (lon < 0.0f) lon = lon + 360.0f;
            ez_llll2gd1(x, y, lat, lon, g.xyp[0], g.xyp[1], g.xyp[2], g.xyp[3], 0.0f);
            return;
        }
        case 'L': {
            // ez_ll2rgd.inc:134-149
            float xlat0 = g.xyp[0], xlon0 = g.xyp[1], dellat = g.xyp[2], dellon = g.xyp[3];
            if (lon < xlon0) lon = lon + 360.0f;
            if (lon > (xlon0 + (float)g.ni * dellon)) lon = lon - 360.0f;
            ez_llll2gd1(x, y, lat, lon, xlat0, xlon0, dellat, dellon, 0.0f);
            return;
        }
        case 'E': {
            // ez_ll2ergd.inc:37-48 : rotate, then affine on the rotated lat-lon (local copies, no mutation)
            float lonr, latr;
            ez_rotate_ll(g.r, lon, lat, lonr, latr);
            float dellon = 360.0f / (float)(g.ni - 1);
            float dellat = 180.0f / (float)g.nj;
            float xlat0 = -90.f + 0.5f * dellat;
            ez_llll2gd1(x, y, latr, lonr, xlat0, 0.0f, dellat, dellon, 0.0f);
            return;
        }
        case '#':
        case 'Z':
        case 'G': {
            // ez_ll2igd.inc:45-85 with coordflag = RELATIF ; nj passed is gr.j2 in gdxyfll.c:170, nj in ez_calcxy.c:81
            float px = 0.f, py = 0.f;
            switch (g.grref) {
                case 'N':
                    ez_vxyfll1(px, py, lat, lon, g.xgref[2], g.xgref[3], g.xgref[0], g.xgref[1], 1);
                    break;
                case 'S':
                    ez_vxyfll1(px, py, lat, lon, g.xgref[2], g.xgref[3], g.xgref[0], g.xgref[1], 2);
                    break;
                case 'L': {
                    float lonref = (__ldg(g.ax) < 0.0f) ? -180.0f : 0.0f;
                    ez_llll2gd1(px, py, lat, lon, g.xgref[0], g.xgref[1], g.xgref[2], g.xgref[3], lonref);
                    px = px - 1.0f;
                    py = py - 1.0f;
                    break;
                }
                case 'E':
                    ez_rotate_ll(g.r, lon, lat, px, py);
                    break;
                default:
                    break;
            }
            int indx = ez_cherche(px, g.ax, g.ni);
            int indy = ez_cherche(py, g.ay, njs);
            if (indx >= g.ni) indx = g.ni - 1;
            if (indy >= njs) indy = njs - 1;
            float ax0 = __ldg(g.ax + indx - 1), ax1 = __ldg(g.ax + indx);
            float ay0 = __ldg(g.ay + indy - 1), ay1 = __ldg(g.ay + indy);
            x = (float)indx + (px - ax0) / (ax1 - ax0);
            y = (float)indy + (py - ay0) / (ay1 - ay0);
            // G-north shift, ez_calcxy.c:90-96 / gdxyfll.c:176-180
            if (g.grtyp == 'G' && g.ig[0] == 1) y = y - (float)njs;
            return;
        }
        default:
            x = 0.f;
            y = 0.f;
            return;
    }
}

// (x, y) in a grid's own index space -> true (lat, lon): c_gdllfxy_orig (gdllfxy.c:251-386) ; with newstyle != 0 the
// y inversion of c_gdllfxy_new (gdllfxy.c:185-190) for Z/#/G grids whose ig2 == 1.
__device__ __forceinline__ void ez_xy2ll(const GridDev &g, float xf, float yf, int newstyle, float &lat, float &lon) {
    lat = 0.f;
    lon = 0.f;
    switch (g.grtyp) {
        case 'A': case 'B': case 'L':
            lat = (float)(((double)yf - 1.0) * (double)g.xg[2] + (double)g.xg[0]);
            lon = (float)(((double)xf - 1.0) * (double)g.xg[3] + (double)g.xg[1]);
            lon = (float)fmod((double)lon + 360.0, 360.0);
            break;
        case 'E': {
            const float dlat = (float)(180.0 / (double)g.nj), dlon = (float)(360.0 / (double)(g.ni - 1));
            const float swlat = (float)(-90.0 + 0.5 * (double)dlat), swlon = 0.0f;
            const float tx = (float)(((double)xf - 1.0) * (double)dlon + (double)swlon);
            const float ty = (float)(((double)yf - 1.0) * (double)dlat + (double)swlat);
            ez_rotate_ll(g.ri, tx, ty, lon, lat);
            break;
        }
        case 'N': case 'S':
            ez_vllfxy1(lat, lon, xf, yf, g.xg[2], g.xg[3], g.xg[0], g.xg[1], g.hemisphere);
            lon = (float)fmod((double)lon + 360.0, 360.0);
            break;
        case '#': case 'Z': case 'G': {
            float yt = yf;
            if (newstyle && g.ig[1] == 1) yt = (float)((double)g.nj + 1.0 - (double)yf);
            int indx = (int)xf - 1, indy = (int)yt - 1;
            indx = indx < 0 ? 0 : indx;
            indy = indy < 0 ? 0 : indy;
            indx = indx > g.ni - 2 ? g.ni - 2 : indx;
            indy = indy > g.j2 - 2 ? g.j2 - 2 : indy;
            const float delxx = __ldg(g.ax + indx + 1) - __ldg(g.ax + indx);
            const float tx = (float)((double)__ldg(g.ax + indx) + (((double)xf - 1.0 - (double)indx) * (double)delxx));
            const float delyy = __ldg(g.ay + indy + 1) - __ldg(g.ay + indy);
            const float ty = (float)((double)__ldg(g.ay + indy) + (((double)yt - 1.0 - (double)indy) * (double)delyy));
            if (g.grref == 'E') {
                ez_rotate_ll(g.ri, tx, ty, lon, lat);
            } else if (g.grref == 'N' || g.grref == 'S') {
                ez_vllfxy1(lat, lon, tx, ty, g.xgref[2], g.xgref[3], g.xgref[0], g.xgref[1], g.hemisphere);
                lon = (float)fmod((double)lon + 360.0, 360.0);
            } else if (g.grref == 'L') {
                lat = ty * g.xgref[2] + g.xgref[0];
                lon = tx * g.xgref[3] + g.xgref[1];
                lon = (float)fmod((double)lon + 360.0, 360.0);
            }
            break;
        }
        default: break;
    }
}

