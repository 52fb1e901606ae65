// librmn_b200/csrc/ez_hostmath.h -- host-side (once-per-grid) numerics of the product library: grid
// descriptor encode/decode, wrap "extension" rules, Newton tables, Gaussian latitudes, rotation matrices.
// These are O(ni+nj) set-up computations that the reference also performs on the host when a grid is
// defined (last row); everything per-point runs on the GPU.
//
// Each routine cites the reference file:line whose behaviour it reproduces (same fp32/fp64 sequence, so
// the tables uploaded to the device match the reference's bit for bit on the same libm).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace ezh {

constexpr float PIE = 3.1415926535898f;      // src/base/pi.cdk
constexpr float RDTODG = 57.295779513082f;
constexpr float DGTORD = 1.7453292519943e-2f;

inline int nint_f(float x) { return (int)std::lroundf(x); }
inline float sign_f(float a, float b) { return std::copysign(std::fabs(a), b); }

// src/base/xyfll.F:62-79
inline void xyfll(float &x, float &y, float dlat, float dlon, float d60, float dgrw, int nhem) {
    float re = 1.866025f * 6.371e+6f / d60;
    float glon = (nhem == 2) ? -dlon : dlon;
    float glat = (nhem == 2) ? -dlat : dlat;
    float rlon = DGTORD * (glon + dgrw);
    float rlat = DGTORD * glat;
    float sinlat = sinf(rlat);
    float r = re * sqrtf((1.f - sinlat) / (1.f + sinlat));
    x = r * cosf(rlon);
    y = r * sinf(rlon);
}

// src/base/llfxy.F:64-99
inline void llfxy(float &dlat, float &dlon, float x, float y, float d60, float dgrw, int nhem) {
    float re = 1.866025f * 6.371e+6f / d60;
    float re2 = re * re;
    dlat = 90.f;
    dlon = 0.f;
    if (!(x == 0.f && y == 0.f)) {
        if (x == 0.f) dlon = sign_f(90.f, y);
        if (x != 0.f) dlon = atanf(y / x) * RDTODG;
        if (x < 0.f) dlon = dlon + sign_f(180.f, y);
        dlon = dlon - dgrw;
        if (dlon > 180.f) dlon = dlon - 360.f;
        if (dlon < -180.f) dlon = dlon + 360.f;
        float r2 = x * x + y * y;
        dlat = (re2 - r2) / (re2 + r2);
        dlat = asinf(dlat) * RDTODG;
    }
    if (nhem == 2) {
        dlat = -dlat;
        dlon = -dlon;
    }
}

// src/base/igaxg.f90:5-143 -- ig -> xg.  Returns false for an unknown grid type.
inline bool cigaxg(char t, float xg[4], const int ig[4]) {
    const int ig1 = ig[0], ig2 = ig[1], ig3 = ig[2], ig4 = ig[3];
    switch (t) {
        case 'N':
        case 'S':
            if (ig4 < 32768) {
                xg[0] = (float)ig2 * 0.1f;
                xg[1] = (float)ig1 * 0.1f;
                xg[2] = (float)ig4 * 100.f;
                xg[3] = (float)ig3 * 0.01f;
            } else {
                int jg3 = ig3, jg4 = ig4 - 32768;
                xg[2] = (float)ig1 * 100.f;
                if (ig3 > 32767) {
                    xg[2] = xg[2] * 10.f;
                    jg3 -= 32768;
                }
                xg[3] = (float)ig2 * .1f;
                if (jg4 > 16383) {
                    xg[3] = 360.f - xg[3];
                    jg4 -= 16384;
                }
                float dlat = 90.f - ((float)jg4 * 180.f / 16383.f);
                float dlon = ((float)jg3 * 360.f / 32767.f);
                xyfll(xg[0], xg[1], dlat, dlon, xg[2], xg[3], t == 'S' ? 2 : 1);
                xg[0] = 1.0f - xg[0];
                xg[1] = 1.0f - xg[1];
            }
            return true;
        case 'A':
        case 'B':
        case 'G':
            xg[0] = (float)ig1;
            xg[1] = (float)ig2;
            xg[2] = 0.f;
            xg[3] = 0.f;
            return true;
        case 'L':
            xg[0] = (float)ig3 * 0.01f - 90.f;
            xg[1] = (float)ig4 * 0.01f;
            xg[2] = (float)ig1 * 0.01f;
            xg[3] = (float)ig2 * 0.01f;
            return true;
        case 'C':
            xg[0] = (float)ig3 * 0.01f - 90.f;
            xg[1] = (float)ig4 * 0.01f;
            xg[2] = 180.f / (float)ig1;
            xg[3] = 360.f / (float)ig2;
            return true;
        case 'E': {
            int i2b = ig3 & 3;
            int lg3 = (int)((uint32_t)ig3 >> 2);
            int lg1 = (int)(((uint32_t)ig1 << 2) | (uint32_t)i2b);
            i2b = ig4 & 3;
            int lg4 = (int)((uint32_t)ig4 >> 2);
            int lg2 = (int)(((uint32_t)ig2 << 2) | (uint32_t)i2b);
            if (lg2 > 3600) lg2 -= 7201;
            xg[0] = (float)(((double)lg1 - 3600.0) / 40.0);
            if (lg3 < 3559) lg3 += 16384;
            xg[1] = (float)(((double)lg3 - 3600.0) / 40.0);
            xg[2] = (float)((double)lg2 / 40.0);
            xg[3] = (float)((double)lg4 / 40.0);
            return true;
        }
        default:
            return false;
    }
}

// src/base/xgaig.f90:4-196 -- xg -> ig
inline bool cxgaig(char t, int ig[4], const float xg[4]) {
    const float xg1 = xg[0], xg2 = xg[1], xg3 = xg[2], xg4 = xg[3];
    switch (t) {
        case 'N':
        case 'S': {
            ig[0] = nint_f(xg2 * 10.f);
            ig[1] = nint_f(xg1 * 10.f);
            ig[2] = nint_f(xg4 * 100.f);
            ig[3] = nint_f(xg3 * 0.01f);
            while (ig[2] < 0) ig[2] += 36000;
            if (ig[0] < 0 || ig[1] < 0 || ig[0] > 2047 || ig[1] > 2047 || ig[3] > 32000) {
                ig[0] = 0; ig[1] = 0; ig[2] = 0; ig[3] = 32768;
                if (xg3 > 204700.f) {
                    ig[2] = 32768;
                    ig[0] = nint_f(xg3 * .001f);
                } else {
                    ig[2] = 0;
                    ig[0] = nint_f(xg3 * .01f);
                }
                ig[1] = nint_f(xg4 * 10.f);
                if (ig[1] < 0) {
                    ig[1] = std::abs(ig[1]);
                    ig[3] += 16384;
                }
                if (ig[1] > 1800) {
                    ig[1] = std::abs(ig[1] - 3600);
                    ig[3] += 16384;
                }
                float dlat, dlon;
                llfxy(dlat, dlon, 1.f - xg1, 1.f - xg2, xg3, xg4, t == 'S' ? 2 : 1);
                dlat = 90.f - dlat;
                if (dlon < 0.f) dlon = dlon + 360.f;
                ig[2] += nint_f(dlon * 32767.f / 360.f);
                ig[3] += nint_f(dlat * 16383.f / 180.f);
            }
            return true;
        }
        case 'A':
        case 'B':
        case 'G':
            ig[0] = (int)xg1;
            ig[1] = (int)xg2;
            ig[2] = 0;
            ig[3] = 0;
            return true;
        case 'L':
            ig[0] = nint_f(xg3 * 100.f);
            ig[1] = nint_f(xg4 * 100.f);
            ig[2] = nint_f((90.f + xg1) * 100.f);
            ig[3] = nint_f(xg2 * 100.f);
            while (ig[3] < 0) ig[3] += 36000;
            if (ig[2] < 0) fprintf(stderr, "cxgaig: Bad specification (LAT0)\n");
            return true;
        case 'C':
            ig[0] = nint_f(180.f / xg3);
            ig[1] = nint_f(360.f / xg4);
            ig[2] = nint_f((90.f + xg1) * 100.f);
            ig[3] = nint_f(xg2 * 100.f);
            while (ig[3] < 0) ig[3] += 36000;
            return true;
        case 'E': {
            float xxg2 = xg2, xxg4 = xg4;
            while (xxg2 < 0.f) xxg2 += 360.f;
            while (xxg4 < 0.f) xxg4 += 360.f;
            ig[0] = nint_f((xg1 + 90.f) * 40.f);
            ig[1] = nint_f(xg3 * 40.f);
            ig[2] = nint_f((xxg2 + 90.f) * 40.f);
            if (ig[2] >= 16384) ig[2] -= 16384;
            ig[3] = nint_f(xxg4 * 40.f);
            int i2b = ig[0] & 3;
            ig[0] = (int)((uint32_t)ig[0] >> 2);
            ig[2] = (int)(((uint32_t)ig[2] << 2) | (uint32_t)i2b);
            if (ig[1] < 0) ig[1] += 7201;
            i2b = ig[1] & 3;
            ig[1] = (int)((uint32_t)ig[1] >> 2);
            ig[3] = (int)(((uint32_t)ig[3] << 2) | (uint32_t)i2b);
            return true;
        }
        default:
            return false;
    }
}

// ez_lac.inc:38-46 for one point ; ez_crot.inc:56-92 (column-major 3x3)
inline void lac1(float xyz[3], float lon, float lat) {
    float dar = acosf(-1.0f) / 180.0f;
    float cosdar = cosf(dar * lat);
    xyz[0] = cosdar * cosf(dar * lon);
    xyz[1] = cosdar * sinf(dar * lon);
    xyz[2] = sinf(dar * lat);
}
inline void crot(float r[9], float ri[9], float lon1, float lat1, float lon2, float lat2) {
    float p[3], q[3];
    lac1(p, lon1, lat1);
    lac1(q, lon2, lat2);
    float a = (p[0] * q[0]) + (p[1] * q[1]) + (p[2] * q[2]);
    float t1 = (p[1] * q[2]) - (q[1] * p[2]);
    float t2 = (q[0] * p[2]) - (p[0] * q[2]);
    float t3 = (p[0] * q[1]) - (q[0] * p[1]);
    float b = sqrtf(t1 * t1 + t2 * t2 + t3 * t3);
    float c = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    float d1 = ((a * p[0]) - q[0]) / b, d2 = ((a * p[1]) - q[1]) / b, d3 = ((a * p[2]) - q[2]) / b;
    float d = sqrtf(d1 * d1 + d2 * d2 + d3 * d3);
    auto R = [&](int i, int j) -> float & { return r[(j - 1) * 3 + (i - 1)]; };
    R(1, 1) = -p[0] / c;
    R(1, 2) = -p[1] / c;
    R(1, 3) = -p[2] / c;
    R(2, 1) = (((a * p[0]) - q[0]) / b) / d;
    R(2, 2) = (((a * p[1]) - q[1]) / b) / d;
    R(2, 3) = (((a * p[2]) - q[2]) / b) / d;
    R(3, 1) = t1 / b;
    R(3, 2) = t2 / b;
    R(3, 3) = t3 / b;
    for (int i = 1; i <= 3; i++)
        for (int j = 1; j <= 3; j++) ri[(j - 1) * 3 + (i - 1)] = R(j, i);
}

// src/base/ordleg.F:25-89
inline float ordleg(float coa, int ir) {
    float sqr2 = sqrtf(2.f);
    int irppm = ir;
    float theta = acosf(coa);
    float c1 = sqr2, fn = 0.f, fn2 = 0.f;
    for (int n = 1; n <= irppm; n++) {
        fn = (float)n;
        fn2 = 2.f * fn;
        c1 = c1 * sqrtf(1.0f - 1.0f / (fn2 * fn2));
    }
    int n = irppm;
    float ang = fn * theta, s1 = 0.0f, c4 = 1.0f, a = -1.0f, b = 0.0f;
    for (int kk = 1; kk <= n + 1; kk += 2) {
        int k = kk - 1;
        if (k == n) c4 = 0.5f * c4;
        s1 = s1 + c4 * cosf(ang);
        a = a + 2.0f;
        b = b + 1.0f;
        ang = theta * (fn - (float)k - 2.0f);
        c4 = (a * (fn - b + 1.0f) / (b * (fn2 - a))) * c4;
    }
    return s1 * c1;
}

// src/base/dgauss.F:25-132 with KASE = GLOBAL, TOL = 1e-6
inline void dgauss(int n, float *roots) {
    const float tol = 1.0e-6f;
    float normn = sqrtf(2.0f / (2.0f * (float)n + 1.0f));
    float normnm = sqrtf(2.0f / (2.0f * (float)n - 1.0f));
    int l = n / 2;
    for (int i = 1; i <= l; i++) {
        float t = (float)(4 * i - 1) * PIE / (float)(4 * n + 2);
        roots[i - 1] = cosf(t + 1.0f / (8.0f * (float)(n * n) * tanf(t)));
    }
    for (int i = 1; i <= l; i++) {
        float delta;
        do {
            float pn = normn * ordleg(roots[i - 1], n);
            float pnm = normnm * ordleg(roots[i - 1], n - 1);
            float rdpdx = (roots[i - 1] * roots[i - 1] - 1.0f) / ((float)n * (roots[i - 1] * pn - pnm));
            delta = -pn * rdpdx;
            roots[i - 1] = roots[i - 1] + delta;
        } while (fabsf(delta) > tol);
        roots[n - i] = -roots[i - 1];
    }
    if (n % 2 != 0) roots[l] = 0.0f;
}

// ez_glat.inc:26-67 with hem = GLOBAL (the only way ez_defaxes.c:88-102 calls it): south-to-north latitudes
inline void glat(float *latroots, int nj) {
    std::vector<float> g(nj);
    dgauss(nj, g.data());
    for (int j = 1; j <= nj / 2; j++) std::swap(g[j - 1], g[nj - j]);
    for (int j = 0; j < nj; j++) latroots[j] = 90.f - RDTODG * acosf(g[j]);
}

// ez_nwtncof.inc:20-178.  cx(ni,6), cy(j1:j2,6), column-major.  Returns false where the reference calls exit(13).
inline bool nwtncof(float *cx, float *cy, const float *ax, const float *ay, int ni, int nj, int j1, int j2, int extension) {
    const int nyr = j2 - j1 + 1;
    auto AX = [&](int i) { return ax[i - 1]; };
    auto AY = [&](int j) { return ay[j - j1]; };
    auto CX = [&](int i, int k) -> float & { return cx[(size_t)(k - 1) * ni + (i - 1)]; };
    auto CY = [&](int j, int k) -> float & { return cy[(size_t)(k - 1) * nyr + (j - j1)]; };
    for (int i = 1; i <= ni - 1; i++)
        if (AX(i + 1) <= AX(i)) {
            fprintf(stderr, "EZ_NWTNCOF: x axis not strictly ascending at %d: %g %g\n", i, AX(i), AX(i + 1));
            return false;
        }
    for (int j = 1; j <= nj - 1; j++)
        if (AY(j + 1) <= AY(j)) {
            fprintf(stderr, "EZ_NWTNCOF: y axis not strictly ascending at %d: %g %g\n", j, AY(j), AY(j + 1));
            return false;
        }
    for (size_t k = 0; k < (size_t)ni * 6; k++) cx[k] = 1.0f;
    for (size_t k = 0; k < (size_t)nyr * 6; k++) cy[k] = 1.0f;
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
    auto fill = [&](int i, float x1, float x2, float x3, float x4) {
        CX(i, 1) = 1.f / (x2 - x1);
        CX(i, 2) = 1.f / (x3 - x1);
        CX(i, 3) = 1.f / (x3 - x2);
        CX(i, 4) = 1.f / (x4 - x1);
        CX(i, 5) = 1.f / (x4 - x2);
        CX(i, 6) = 1.f / (x4 - x3);
    };
    if (extension == 1) {
        fill(1, AX(1) - (AX(ni) - AX(ni - 1)), AX(1), AX(2), AX(3));
        fill(ni - 1, AX(ni - 2), AX(ni - 1), AX(ni), AX(ni) + (AX(2) - AX(1)));
    }
    if (extension == 2) {
        fill(1, AX(1) - (360.0f - AX(ni)), AX(1), AX(2), AX(3));
        fill(ni - 1, AX(ni - 2), AX(ni - 1), AX(ni), AX(1) + 360.0f);
        fill(ni, AX(ni - 1), AX(ni), AX(1) + 360.0f, AX(2) + 360.0f);
    }
    return true;
}

}  // namespace ezh
