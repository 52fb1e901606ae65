// librmn_b200/csrc/ez_interp.cu -- gather-interpolation kernels (K0-K10, K19 of SURVEY.md 2.3), sm_100a.
//
// Layout: one thread per target point, fields/levels looped INSIDE the thread so that x, y and the zone
// byte (9 B/point) are read once per batch and the stencil indices/weights are computed once and reused
// for every level.  Consecutive threads own consecutive target points: the output store of each level is
// one fully coalesced 128-B line per warp, and the source gathers of neighbouring points fall in the same
// few 128-B lines (L1-resident; the whole source level set of one wave of CTAs sits in the 126 MB L2).
// Zone handling (DEHORS / AU_NORD / AU_SUD / POLE_*, ez_corrval.c) is fused in as a per-point class, so
// there is no second scatter pass and no compaction.
//
// Arithmetic modes (InterpOpts.arith):
//   0  "reference order": every fp32/fp64 operation in the order the Fortran leaves perform it
//      (ez_rgdint_*.inc, ez_irgdint_*.inc) -- used for bit-parity tests;
//   1  "weight form": the same interpolant written as sum_a sum_b wx_a wy_b z_ab with the 4+4 (or 2+2)
//      weights evaluated once per point in fp64 and applied with fp64 FMAs per level.
// Compiled with -fmad=false; the weight form uses explicit fma().
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <limits.h>
#include <math.h>
#include "ez_kernels.h"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ int f_ifix(float x) { return (int)x; }
__device__ __forceinline__ int f_nint(float x) { return (int)lroundf(x); }

// ---------------------------------------------------------------------------------------------------
// field access
// ---------------------------------------------------------------------------------------------------
// Logical z(i, j), i in 1..ni, j in j1..j2 of the (possibly y-inverted, possibly hemispherically mirrored)
// source field.  Folds permut() (ezsint.c:100-106) and ez_xpngdag2/ez_xpngdb2 (K11) into index math.
struct FieldFetch {
    const float *__restrict__ z;
    int ni, nj, j1, xpn, yinv;
    float sgn;
    __device__ __forceinline__ float operator()(int i, int j) const {
        if (xpn == XPN_NONE && !yinv) return __ldg(z + (long)(j - j1) * ni + (i - 1));
        int jj = j;
        float s = 1.0f;
        switch (xpn) {
            case XPN_AG_NORTH: if (j < 1) { jj = 1 - j; s = sgn; } break;
            case XPN_AG_SOUTH: if (j > nj) { jj = 2 * nj - j + 1; s = sgn; } break;
            case XPN_B_NORTH:  if (j < 1) { jj = 2 - j; s = sgn; } break;
            case XPN_B_SOUTH:  if (j > nj) { jj = 2 * nj - j; s = sgn; } break;
            default: break;
        }
        if (yinv) jj = nj + 1 - jj;
        return s * __ldg(z + (long)(jj - 1) * ni + (i - 1));
    }
};

// temp(ni,4) of ez_fillnpole / ez_fillspole, never materialised: three source rows + a constant pole row.
// jbase = the j index of strip row 0 ; pole_row = 3 (north) or 0 (south) ; jsrc0 = source j of the first data row.
struct StripFetch {
    FieldFetch f;
    int jbase, pole_row, jsrc0;
    float poleval;
    const float *polerow;   // vectors: a synthetic pole ROW (ez_calcnpolarwind / ez_calcspolarwind) instead of one pole value
    __device__ __forceinline__ float operator()(int i, int j) const {
        int r = j - jbase;
        if (r == pole_row) return polerow ? __ldg(polerow + i - 1) : poleval;
        return f(i, jsrc0 + (pole_row == 0 ? r - 1 : r));
    }
};

// ---------------------------------------------------------------------------------------------------
// reference-order evaluators
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double zlin(double a, double b, double dx) { return a + (b - a) * dx; }

// cubic8.cdk:3 with the default-real literals rounded to fp32
__device__ __forceinline__ double cubic(double z1, double z2, double z3, double z4, double dx) {
    const double c6 = (double)0.1666666666666f, c3 = (double)0.3333333333333f, h = 0.5;
    return ((((z4 - z1) * c6 + h * (z2 - z3)) * dx + h * (z1 + z3) - z2) * dx + z3 - c6 * z4 - h * z2 - c3 * z1) * dx + z2;
}

// ez_rgdint_0.inc:27-32
template <class F>
__device__ __forceinline__ float rgd0(const F &z, int ni, int j1, int j2, float px, float py) {
    int i = min(ni, max(1, f_nint(px)));
    int j = min(j2, max(j1, f_nint(py)));
    return z(i, j);
}

// ez_rgdint_1_nw.inc:30-41 (wrap < 0) and ez_rgdint_1_w.inc:31-48 (wrap >= 0)
template <class F>
__device__ __forceinline__ float rgd1(const F &z, int ni, int j1, int j2, int wrap, float px, float py) {
    int i, ip1;
    if (wrap < 0) {
        i = min(ni - 1, max(1, f_ifix(px)));
        ip1 = i + 1;
    } else {
        int limite = ni + 2 - wrap;
        i = min(ni - 2 + wrap, max(1, f_ifix(px)));
        ip1 = i + 1;
        if (wrap > 0 && i == (ni - 2 + wrap)) ip1 = (limite + i + 1) % limite;
    }
    int j = min(j2 - 1, max(j1, f_ifix(py)));
    double dx = (double)(px - (float)i);
    double dy = (double)(py - (float)j);
    double y2 = zlin((double)z(i, j), (double)z(ip1, j), dx);
    double y3 = zlin((double)z(i, j + 1), (double)z(ip1, j + 1), dx);
    return (float)zlin(y2, y3, dy);
}

// column indices of ez_rgdint_3_nw.inc:61 (mode 0), ez_rgdint_3_w.inc:72-93 (mode 1), ez_rgdint_3_wnnc.inc:73-93 (mode 2)
__device__ __forceinline__ void rgd3_cols(int mode, int ni, int wrap, float px, int &i, int &im1, int &ip1, int &ip2) {
    if (mode == 0) {
        i = min(ni - 2, max(2, f_ifix(px)));
        im1 = i - 1; ip1 = i + 1; ip2 = i + 2;
        return;
    }
    int limite = ni + 2 - wrap;
    i = min(ni - 2 + wrap, max(1, max(2 - wrap, f_ifix(px))));
    bool wrapped = (mode == 1) ? (wrap > 0) : ((wrap > 0 && i <= 1) || i >= (ni - 1));
    if (wrapped) {
        im1 = (limite + i - 1) % limite;
        ip1 = (limite + i + 1) % limite;
        ip2 = (limite + i + 2) % limite;
        if (im1 == 0) im1 = ni;
        if (i == 0) i = ni;
        if (ip1 == 0) ip1 = ni;
        if (ip2 == 0) ip2 = ni;
        if (wrap == 1) {
            if (ip2 == ni) ip2 = 2;
            if (im1 == ni) im1 = ni - 1;
        }
    } else {
        im1 = i - 1; ip1 = i + 1; ip2 = i + 2;
    }
}

template <class F>
__device__ __forceinline__ float rgd3(const F &z, int mode, int ni, int j1, int j2, int wrap, float px, float py) {
    int i, im1, ip1, ip2;
    rgd3_cols(mode, ni, wrap, px, i, im1, ip1, ip2);
    int j = min(j2 - 2, max(j1 + 1, f_ifix(py)));
    double dx = (double)(px - (float)i);
    double dy = (double)(py - (float)j);
    double y1 = cubic(z(im1, j - 1), z(i, j - 1), z(ip1, j - 1), z(ip2, j - 1), dx);
    double y2 = cubic(z(im1, j), z(i, j), z(ip1, j), z(ip2, j), dx);
    double y3 = cubic(z(im1, j + 1), z(i, j + 1), z(ip1, j + 1), z(ip2, j + 1), dx);
    double y4 = cubic(z(im1, j + 2), z(i, j + 2), z(ip1, j + 2), z(ip2, j + 2), dx);
    return (float)cubic(y1, y2, y3, y4, dy);
}

// ez_irgdint_1_nw.inc:31-47 ; ay is 1-based over nj
template <class F>
__device__ __forceinline__ float irgd1_nw(const F &z, const float *__restrict__ ax, const float *__restrict__ ay, int ni, int nj, float px, float py) {
    int i = min(ni - 1, max(1, f_ifix(px)));
    int j = min(nj - 1, max(1, f_ifix(py)));
    float axi = __ldg(ax + i - 1), axi1 = __ldg(ax + i), ayj = __ldg(ay + j - 1), ayj1 = __ldg(ay + j);
    double x1 = axi, x2 = axi1;
    double x = (double)axi + (x2 - x1) * (double)(px - (float)i);
    double y = (double)(ayj + (ayj1 - ayj) * (py - (float)j));
    double dx = (x - x1) / (x2 - x1);
    double dy = (y - (double)ayj) / (double)(ayj1 - ayj);
    double y1 = zlin((double)z(i, j), (double)z(i + 1, j), dx);
    double y2 = zlin((double)z(i, j + 1), (double)z(i + 1, j + 1), dx);
    return (float)zlin(y1, y2, dy);
}

// ez_irgdint_1_w.inc:31-61 ; ay(j1:j2)
template <class F>
__device__ __forceinline__ float irgd1_w(const F &z, const float *__restrict__ ax, const float *__restrict__ ay, int ni, int j1, int j2, int wrap,
                                         float px, float py) {
    int limite = ni + 2 - wrap;
    int i = min(ni - 2 + wrap, max(1, f_ifix(px)));
    int j = min(j2 - 1, max(j1 + 1, f_ifix(py)));
    if (j < 0) j = j - 1;
    int ip1 = i + 1;
    double x1 = __ldg(ax + i - 1), x2 = 0.0;
    if (ip1 <= ni) x2 = __ldg(ax + ip1 - 1);
    if (wrap > 0 && i == (ni - 2 + wrap)) {
        ip1 = (limite + i + 1) % limite;
        x2 = (double)(__ldg(ax + 1) + __ldg(ax + ni - 1));
    }
    float ayj = __ldg(ay + (j - j1)), ayj1 = __ldg(ay + (j + 1 - j1));
    double x = x1 + (x2 - x1) * (double)(px - (float)i);
    double y = (double)(ayj + (ayj1 - ayj) * (py - (float)j));
    double dx = (x - x1) / (x2 - x1);
    double dy = (y - (double)ayj) / (double)(ayj1 - ayj);
    double y1 = zlin((double)z(i, j), (double)z(ip1, j), dx);
    double y2 = zlin((double)z(i, j + 1), (double)z(ip1, j + 1), dx);
    return (float)zlin(y1, y2, dy);
}

// Newton divided-difference statement functions, ez_irgdint_3_nw.inc:84-87
__device__ __forceinline__ double nfa(double a1, double a2, double a3, double a4, double x, double x1, double x2, double x3) {
    return a1 + (x - x1) * (a2 + (x - x2) * (a3 + a4 * (x - x3)));
}
__device__ __forceinline__ double nfa2(double c1, double a1, double a2) { return c1 * (a2 - a1); }
__device__ __forceinline__ double nfa3(double c1, double c2, double c3, double a1, double a2, double a3) {
    return c2 * (c3 * (a3 - a2) - c1 * (a2 - a1));
}
__device__ __forceinline__ double nfa4(double c1, double c2, double c3, double c4, double c5, double c6, double a1, double a2, double a3, double a4) {
    return c4 * (c5 * (c6 * (a4 - a3) - c3 * (a3 - a2)) - c2 * (c3 * (a3 - a2) - c1 * (a2 - a1)));
}

// One Newton pass over 4 samples.  R4 = true rounds every statement-function result to fp32, as
// ez_irgdint_3_nw.inc:32 declares fa..fa4 REAL; false keeps fp64 (ez_irgdint_3_w.inc:78, _wnnc:28).
template <bool R4>
__device__ __forceinline__ double newton4(const double c[6], double z1, double z2, double z3, double z4, double x, double x1, double x2, double x3) {
    double a2 = nfa2(c[0], z1, z2);
    double a3 = nfa3(c[0], c[1], c[2], z1, z2, z3);
    double a4 = nfa4(c[0], c[1], c[2], c[3], c[4], c[5], z1, z2, z3, z4);
    if (R4) { a2 = (double)(float)a2; a3 = (double)(float)a3; a4 = (double)(float)a4; }
    double r = nfa(z1, a2, a3, a4, x, x1, x2, x3);
    if (R4) r = (double)(float)r;
    return r;
}

// ez_irgdint_3_nw.inc:89-165 ; ax(i1:i2), ay(j1:j2), cx(i1:i2,6), cy(j1:j2,6)
template <class F>
__device__ __forceinline__ float irgd3_nw(const F &z, const GridDev &g, float px, float py) {
    const int i1 = g.i1, i2 = g.i2, j1 = g.j1, j2 = g.j2;
    const int nx = i2 - i1 + 1, ny = j2 - j1 + 1;
    int i = min(i2 - 2, max(i1 + 1, f_ifix(px)));
    int j = min(j2 - 2, max(j1 + 1, f_ifix(py)));
    const float *ax = g.ax - i1, *ay = g.ay - j1;
    float axm = __ldg(ax + i - 1), ax0 = __ldg(ax + i), axp = __ldg(ax + i + 1);
    float aym = __ldg(ay + j - 1), ay0 = __ldg(ay + j), ayp = __ldg(ay + j + 1);
    double x = (double)(ax0 + (axp - ax0) * (px - (float)i));
    double y = (double)(ay0 + (ayp - ay0) * (py - (float)j));
    double cx[6], cy[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        cx[k] = (double)__ldg(g.cx + (long)k * nx + (i - i1));
        cy[k] = (double)__ldg(g.cy + (long)k * ny + (j - j1));
    }
    double b[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int jj = j - 1 + r;
        b[r] = newton4<true>(cx, z(i - 1, jj), z(i, jj), z(i + 1, jj), z(i + 2, jj), x, axm, ax0, axp);
    }
    double b2 = (double)(float)nfa2(cy[0], b[0], b[1]);
    double b3 = (double)(float)nfa3(cy[0], cy[1], cy[2], b[0], b[1], b[2]);
    double b4 = (double)(float)nfa4(cy[0], cy[1], cy[2], cy[3], cy[4], cy[5], b[0], b[1], b[2], b[3]);
    return (float)nfa(b[0], b2, b3, b4, y, aym, ay0, ayp);
}

// seam logic of ez_irgdint_3_w.inc:97-160 (nnc = false) / ez_irgdint_3_wnnc.inc (nnc = true, iplus2 = 1 at wrap 1, i = ni-1)
__device__ __forceinline__ void seam_x(int wrap, int ni, const float *__restrict__ ax, int i, bool nnc, int &im1, int &ip1, int &ip2,
                                       float &x1, float &x2, float &x3, float &x4) {
#define AX(k) __ldg(ax + (k)-1)
    im1 = i - 1; ip1 = i + 1; ip2 = i + 2;
    if (wrap == 1 && (i <= 1 || i >= (ni - wrap))) {
        if (i == 1) {
            im1 = ni - 1; ip1 = 2; ip2 = 3;
            x1 = AX(ni - 1) - 360.0f; x2 = AX(1); x3 = AX(2); x4 = AX(3);
        }
        if (i == (ni - 1)) {
            im1 = ni - 2; ip1 = ni; ip2 = nnc ? 1 : 2;
            x1 = AX(ni - 2); x2 = AX(ni - 1); x3 = AX(ni); x4 = AX(2) + 360.0f;
        }
    } else if (wrap == 2 && (i <= 1 || i > (ni - wrap))) {
        if (i == 1) {
            im1 = ni; ip1 = 2; ip2 = 3;
            x1 = AX(ni) - 360.0f; x2 = AX(1); x3 = AX(2); x4 = AX(3);
        }
        if (i == (ni - 1)) {
            im1 = ni - 2; ip1 = ni; ip2 = 1;
            x1 = AX(ni - 2); x2 = AX(ni - 1); x3 = AX(ni); x4 = AX(1) + 360.0f;
        }
        if (i == ni) {
            im1 = ni - 1; ip1 = 1; ip2 = 2;
            x1 = AX(ni - 1); x2 = AX(ni); x3 = AX(1) + 360.0f; x4 = AX(2) + 360.0f;
        }
    } else {
        x1 = AX(im1); x2 = AX(i); x3 = AX(ip1); x4 = AX(ip2);
    }
#undef AX
}

// ez_irgdint_3_w.inc:93-229 ; cx(ni,6), ay/cy over j1:j2
template <class F>
__device__ __forceinline__ float irgd3_w(const F &z, const GridDev &g, float px, float py) {
    const int ni = g.ni, j1 = g.j1, j2 = g.j2, wrap = g.ext;
    const int ny = j2 - j1 + 1;
    int i = min(ni - 2 + wrap, max(1, max(2 - wrap, f_ifix(px))));
    int j = min(j2 - 2, max(j1 + 1, f_ifix(py)));
    int im1, ip1, ip2;
    float x1 = 0.f, x2 = 0.f, x3 = 0.f, x4 = 0.f;
    seam_x(wrap, ni, g.ax, i, false, im1, ip1, ip2, x1, x2, x3, x4);
    const float *ay = g.ay - j1;
    float aym = __ldg(ay + j - 1), ay0 = __ldg(ay + j), ayp = __ldg(ay + j + 1);
    double x = (double)(x2 + (x3 - x2) * (px - (float)i));
    double y = (double)(ay0 + (ayp - ay0) * (py - (float)j));
    double cx[6], cy[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        cx[k] = (double)__ldg(g.cx + (long)k * ni + (i - 1));
        cy[k] = (double)__ldg(g.cy + (long)k * ny + (j - j1));
    }
    double b[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int jj = j - 1 + r;
        b[r] = newton4<false>(cx, z(im1, jj), z(i, jj), z(ip1, jj), z(ip2, jj), x, x1, x2, x3);
    }
    double b2 = nfa2(cy[0], b[0], b[1]);
    double b3 = nfa3(cy[0], cy[1], cy[2], b[0], b[1], b[2]);
    double b4 = nfa4(cy[0], cy[1], cy[2], cy[3], cy[4], cy[5], b[0], b[1], b[2], b[3]);
    return (float)nfa(b[0], b2, b3, b4, y, aym, ay0, ayp);
}

// ez_irgdint_3_wnnc.inc:93-243 ; coefficients on the fly, 4-row strip z(ni, j1:j2), ay4 = ay(j1:j2).
// Split into the part that depends on (px, py) only -- computed once per point -- and the part that reads the field,
// so that callers looping over levels do the fp64 reciprocals once; the arithmetic of each part is unchanged.
struct WnncPlan {
    int i, im1, ip1, ip2, j;
    double cx[6], cy[6];
    double x, x1, x2, x3, y, y1, y2, y3;
};
__device__ __forceinline__ WnncPlan wnnc_plan(const float *__restrict__ ax, const float ay4[4], int ni, int j1, int j2, int wrap, float px, float py) {
    WnncPlan p;
    p.i = min(ni - 2 + wrap, max(1, max(2 - wrap, f_ifix(px))));
    p.j = min(j2 - 2, max(j1 + 1, f_ifix(py)));
    float fx1 = 0.f, fx2 = 0.f, fx3 = 0.f, fx4 = 0.f;
    seam_x(wrap, ni, ax, p.i, true, p.im1, p.ip1, p.ip2, fx1, fx2, fx3, fx4);
    const double x1 = fx1, x2 = fx2, x3 = fx3, x4 = fx4;
    const float ay0 = ay4[p.j - j1], ayp = ay4[p.j + 1 - j1];
    p.x = x2 + (x3 - x2) * (double)(px - (float)p.i);
    p.y = (double)(ay0 + (ayp - ay0) * (py - (float)p.j));
    p.cx[0] = 1.0 / (x2 - x1); p.cx[1] = 1.0 / (x3 - x1); p.cx[2] = 1.0 / (x3 - x2);
    p.cx[3] = 1.0 / (x4 - x1); p.cx[4] = 1.0 / (x4 - x2); p.cx[5] = 1.0 / (x4 - x3);
    const double y1 = ay4[p.j - 1 - j1], y2 = ay0, y3 = ayp, y4 = ay4[p.j + 2 - j1];
    p.cy[0] = 1.0 / (y2 - y1); p.cy[1] = 1.0 / (y3 - y1); p.cy[2] = 1.0 / (y3 - y2);
    p.cy[3] = 1.0 / (y4 - y1); p.cy[4] = 1.0 / (y4 - y2); p.cy[5] = 1.0 / (y4 - y3);
    p.x1 = x1; p.x2 = x2; p.x3 = x3; p.y1 = y1; p.y2 = y2; p.y3 = y3;
    return p;
}
template <class F>
__device__ __forceinline__ float wnnc_eval(const WnncPlan &p, const F &z) {
    double b[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int jj = p.j - 1 + r;
        b[r] = newton4<false>(p.cx, z(p.im1, jj), z(p.i, jj), z(p.ip1, jj), z(p.ip2, jj), p.x, p.x1, p.x2, p.x3);
    }
    double b2 = nfa2(p.cy[0], b[0], b[1]);
    double b3 = nfa3(p.cy[0], p.cy[1], p.cy[2], b[0], b[1], b[2]);
    double b4 = nfa4(p.cy[0], p.cy[1], p.cy[2], p.cy[3], p.cy[4], p.cy[5], b[0], b[1], b[2], b[3]);
    return (float)nfa(b[0], b2, b3, b4, p.y, p.y1, p.y2, p.y3);
}
template <class F>
__device__ __forceinline__ float irgd3_wnnc(const F &z, const float *__restrict__ ax, const float ay4[4], int ni, int j1, int j2, int wrap,
                                            float px, float py) {
    const WnncPlan p = wnnc_plan(ax, ay4, ni, j1, j2, wrap, px, py);
    return wnnc_eval(p, z);
}

// c_gdinterp's dispatch (gdinterp.c:133-281) for one point, reference-order arithmetic
template <class F>
__device__ __forceinline__ float interp_point(const F &z, const GridDev &g, int degree, float px, float py) {
    const bool irregular = (g.grtyp == '#' || g.grtyp == 'Z' || g.grtyp == 'G');
    if (degree == EZ_NEAREST) return rgd0(z, g.ni, g.j1, g.j2, px, py);
    if (irregular) {
        if (degree == EZ_LINEAR) {
            if (g.ext == 0) return irgd1_nw(z, g.ax, g.ay, g.ni, g.nj, px, py);
            return irgd1_w(z, g.ax, g.ay, g.ni, g.j1, g.j2, g.ext, px, py);
        }
        if (g.ext == 0) return irgd3_nw(z, g, px, py);
        return irgd3_w(z, g, px, py);
    }
    if (degree == EZ_LINEAR) {
        if (g.ext == 2) return rgd1(z, g.ni, g.j1, g.j2, 2, px, py);
        return rgd1(z, g.ni, g.j1, g.j2, -1, px, py);
    }
    if (g.ext == 0) return rgd3(z, 0, g.ni, g.j1, g.j2, 0, px, py);
    return rgd3(z, 1, g.ni, g.j1, g.j2, g.ext, px, py);
}

__device__ __forceinline__ FieldFetch make_fetch(const GridDev &g, const float *z, int sym_sign) {
    FieldFetch f;
    f.z = z; f.ni = g.ni; f.nj = g.nj; f.j1 = g.j1; f.xpn = g.xpn; f.yinv = g.yinv;
    f.sgn = (float)sym_sign;
    return f;
}

// ez_corrval_aunord.c:39-112 / ez_corrval_ausud.c:46-129 for one point
// (scalars: poleval ; vectors, ez_corrvec_aunord.c:39-133 / ez_corrvec_ausud.c: the same strips with a pole row)
__device__ __noinline__ float polar_point(const FieldFetch &ff, const GridDev &g, int degree, bool north, float poleval, const float *polerow,
                                          float px, float py) {
    StripFetch s;
    s.f = ff;
    s.poleval = poleval;
    s.polerow = polerow;
    const bool irr = (g.grtyp == 'Z' || g.grtyp == 'G');   // ('E' has no axes in the reference: regular strip)
    if (north) {
        int j1s = g.j2 - 2, j2s = j1s + 3;
        s.pole_row = 3; s.jsrc0 = g.j2 - 2;
        if (degree == EZ_CUBIC) {
            s.jbase = j1s;
            if (irr) {
                float ay4[4];
                const float *ay = g.ay;   // C 0-based indices on the raw axis, ez_corrval_aunord.c:61-77
                if (__ldg(ay + g.j2 - 1) == 90.0f) {
                    ay4[0] = __ldg(ay + g.j2 - 4); ay4[1] = __ldg(ay + g.j2 - 3); ay4[2] = __ldg(ay + g.j2 - 2); ay4[3] = __ldg(ay + g.j2 - 1);
                } else {
                    ay4[0] = __ldg(ay + g.j2 - 3); ay4[1] = __ldg(ay + g.j2 - 2); ay4[2] = __ldg(ay + g.j2 - 1); ay4[3] = 90.0f;
                }
                return irgd3_wnnc(s, g.ax, ay4, g.ni, j1s, j2s, g.ext, px, py);
            }
            return rgd3(s, 2, g.ni, j1s, j2s, g.ext, px, py);
        }
        // linear / nearest: strip indexed 1..4 with y shifted by (j2-3) in double, ez_corrval_aunord.c:88-105
        s.jbase = 1;
        float ty = (float)((double)py - (1.0 * (double)(g.j2 - 3)));
        if (degree == EZ_LINEAR) return rgd1(s, g.ni, 1, 4, g.ext, px, ty);
        return rgd0(s, g.ni, 1, 4, px, ty);
    }
    int j1s = g.j1 - 1, j2s = j1s + 3;
    s.pole_row = 0; s.jsrc0 = g.j1; s.jbase = j1s;
    if (degree == EZ_CUBIC) {
        if (irr) {
            float ay4[4];
            const float *ay = g.ay;       // ez_corrval_ausud.c:66-92
            if (__ldg(ay + g.j1 - 1) == -90.0f) {
                ay4[0] = __ldg(ay + 0); ay4[1] = __ldg(ay + 1); ay4[2] = __ldg(ay + 2); ay4[3] = __ldg(ay + 3);
            } else {
                ay4[0] = -90.0f; ay4[1] = __ldg(ay + 0); ay4[2] = __ldg(ay + 1); ay4[3] = __ldg(ay + 2);
            }
            return irgd3_wnnc(s, g.ax, ay4, g.ni, j1s, j2s, g.ext, px, py);
        }
        return rgd3(s, 2, g.ni, j1s, j2s, g.ext, px, py);
    }
    if (degree == EZ_LINEAR) return rgd1(s, g.ni, j1s, j2s, g.ext, px, py);
    return rgd0(s, g.ni, j1s, j2s, px, py);
}

// The level-invariant half of polar_point for the common heavy case (cubic, irregular 'Z'/'G' source): strip geometry
// and the wnnc plan.  polar_strip_eval then costs the reference's own newton4 evaluations per level and nothing else.
struct PolarStrip {
    int jbase, pole_row, jsrc0;
    WnncPlan plan;
};
__device__ __forceinline__ PolarStrip polar_strip_plan(const GridDev &g, bool north, float px, float py) {
    PolarStrip ps;
    float ay4[4];
    const float *ay = g.ay;
    if (north) {
        const int j1s = g.j2 - 2, j2s = j1s + 3;
        ps.pole_row = 3; ps.jsrc0 = g.j2 - 2; ps.jbase = j1s;
        if (__ldg(ay + g.j2 - 1) == 90.0f) {
            ay4[0] = __ldg(ay + g.j2 - 4); ay4[1] = __ldg(ay + g.j2 - 3); ay4[2] = __ldg(ay + g.j2 - 2); ay4[3] = __ldg(ay + g.j2 - 1);
        } else {
            ay4[0] = __ldg(ay + g.j2 - 3); ay4[1] = __ldg(ay + g.j2 - 2); ay4[2] = __ldg(ay + g.j2 - 1); ay4[3] = 90.0f;
        }
        ps.plan = wnnc_plan(g.ax, ay4, g.ni, j1s, j2s, g.ext, px, py);
    } else {
        const int j1s = g.j1 - 1, j2s = j1s + 3;
        ps.pole_row = 0; ps.jsrc0 = g.j1; ps.jbase = j1s;
        if (__ldg(ay + g.j1 - 1) == -90.0f) {
            ay4[0] = __ldg(ay + 0); ay4[1] = __ldg(ay + 1); ay4[2] = __ldg(ay + 2); ay4[3] = __ldg(ay + 3);
        } else {
            ay4[0] = -90.0f; ay4[1] = __ldg(ay + 0); ay4[2] = __ldg(ay + 1); ay4[3] = __ldg(ay + 2);
        }
        ps.plan = wnnc_plan(g.ax, ay4, g.ni, j1s, j2s, g.ext, px, py);
    }
    return ps;
}
__device__ __forceinline__ float polar_strip_eval(const PolarStrip &ps, const FieldFetch &ff, float poleval, const float *polerow) {
    StripFetch s;
    s.f = ff; s.poleval = poleval; s.polerow = polerow;
    s.jbase = ps.jbase; s.pole_row = ps.pole_row; s.jsrc0 = ps.jsrc0;
    return wnnc_eval(ps.plan, s);
}
// everything ez_corrval.c:54-143 does to one point carrying zone bits
__device__ __noinline__ float zone_point(const FieldFetch &ff, const GridDev &g, const InterpOpts &o, uint8_t bits, const FieldScalars *fs,
                                         const float *prow, int f, float px, float py, float normal) {
    float v = normal;
    if (bits & ZB_DEHORS) {
        if (o.extrap >= EZ_MAXIMUM)
            v = (o.vector || !fs) ? 0.0f : fs[f].fudge;
        else
            v = interp_point(ff, g, o.extrap, px, py);
    }
    if (o.vector) {
        // ez_corrvec.c:24-48: only the AU_NORD / AU_SUD lists are rewritten, from strips whose pole row is synthetic
        if (prow) {
            const float *rows = prow + (long)f * 2 * g.ni;
            if (bits & ZB_AU_NORD) v = polar_point(ff, g, o.degree, true, 0.f, rows, px, py);
            if (bits & ZB_AU_SUD) v = polar_point(ff, g, o.degree, false, 0.f, rows + g.ni, px, py);
        }
        return v;
    }
    if (bits & ZB_AU_NORD) v = polar_point(ff, g, o.degree, true, fs[f].pole_n, nullptr, px, py);
    if (bits & ZB_AU_SUD) v = polar_point(ff, g, o.degree, false, fs[f].pole_s, nullptr, px, py);
    if (bits & ZB_POLE_NORD) v = fs[f].pole_n;
    if (bits & ZB_POLE_SUD) v = fs[f].pole_s;
    return v;
}

// ---------------------------------------------------------------------------------------------------
// the interpolation kernel: generic, reference-order arithmetic
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ref_point_body(const GridDev &g, const InterpArgs &a, const float *__restrict__ zbase, int n, float px, float py,
                                               uint8_t zbits) {
    for (int f = 0; f < a.nfields; f++) {
        FieldFetch ff = make_fetch(g, zbase + (long)f * a.in_stride, a.o.sym_sign);
        float v = interp_point(ff, g, a.o.degree, px, py);
        if (zbits) v = zone_point(ff, g, a.o, zbits, a.fs, a.polerow, f, px, py, v);
        a.zout[(long)f * a.out_stride + n] = v;
    }
}

__global__ void __launch_bounds__(TPB) k_interp_ref(const __grid_constant__ InterpArgs a) {
    int n = blockIdx.x * TPB + threadIdx.x;
    if (n >= a.npts) return;
    const float px = a.x[n], py = a.y[n];
    const uint8_t bits = a.zb ? a.zb[n] : (uint8_t)0;
    const uint8_t zbits = a.use_zones ? (uint8_t)(bits & 0x1f) : (uint8_t)0;
    if (a.yy && (bits & ZB_YANG))
        ref_point_body(a.src2, a, a.zin + a.panel_stride, n, px, py, zbits);
    else
        ref_point_body(a.src, a, a.zin, n, px, py, zbits);
}

// ---------------------------------------------------------------------------------------------------
// K10 pre-pass: per-field scalars
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int f2ord(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_minmax_init(unsigned int *scratch, int nfields) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < nfields) {
        scratch[2 * f] = 0xffffffffu;   // min
        scratch[2 * f + 1] = 0u;        // max
    }
}

// ez_aminmax.inc:26-27 over the logical (expanded) field z(ni, j1:j2); grid = (blocks, nfields)
__global__ void __launch_bounds__(TPB) k_minmax(InterpArgs a, unsigned int *scratch) {
    const int f = blockIdx.y;
    const GridDev &g = a.src;
    FieldFetch ff = make_fetch(g, a.zin + (long)f * a.in_stride, a.o.sym_sign);
    const long total = (long)g.ni * (g.j2 - g.j1 + 1);
    float mn = INFINITY, mx = -INFINITY;
    for (long k = (long)blockIdx.x * TPB + threadIdx.x; k < total; k += (long)gridDim.x * TPB) {
        int i = (int)(k % g.ni) + 1, j = (int)(k / g.ni) + g.j1;
        float v = ff(i, j);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(scratch + 2 * f, f2ord(mn));
        atomicMax(scratch + 2 * f + 1, f2ord(mx));
    }
}

// ez_calcpoleval.inc:30-45: strictly sequential fp32 accumulation, so the pole value is bit-identical to the
// reference's; plus the DEHORS fudge value of ez_corrval.c:62-91.  One CTA per (field, pole): the row (and, for
// Z/E sources, its products with the axis spacing -- each an independent, identically rounded operation) is
// fetched by the whole CTA into shared memory in chunks, then ONE thread adds the chunk in index order.
constexpr int FS_CHUNK = 2048;
__global__ void __launch_bounds__(256) k_field_scalars(InterpArgs a, int need_poles, int need_minmax, FieldScalars *fs, const unsigned int *scratch) {
    __shared__ float s_row[FS_CHUNK];
    __shared__ float s_acc;
    const int f = blockIdx.x >> 1, which = blockIdx.x & 1;
    if (f >= a.nfields) return;
    const GridDev &g = a.src;
    if (need_poles) {
        FieldFetch ff = make_fetch(g, a.zin + (long)f * a.in_stride, a.o.sym_sign);
        const int jrow = which ? g.j1 : g.j2;   // north pole uses the last row, south the first (ez_corrval.c:133-140)
        const bool weighted = (g.grtyp == 'Z' && g.grref == 'E');
        const int n = weighted ? g.ni - 1 : g.ni;
        if (threadIdx.x == 0) s_acc = 0.0f;
        for (int i0 = 0; i0 < n; i0 += FS_CHUNK) {
            const int m = min(FS_CHUNK, n - i0);
            __syncthreads();
            for (int k = threadIdx.x; k < m; k += blockDim.x) {
                const int i = i0 + k + 1;
                float v = ff(i, jrow);
                if (weighted) v = v * (__ldg(g.ax + i) - __ldg(g.ax + i - 1));
                s_row[k] = v;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                float p = s_acc;
                for (int k = 0; k < m; k++) p = p + s_row[k];
                s_acc = p;
            }
        }
        if (threadIdx.x == 0) {
            float p = s_acc;
            if (weighted) {
                float span = __ldg(g.ax + g.ni - 1) - __ldg(g.ax);
                if (span != 0.0f) p = p / span;
            } else {
                p = p / (1.0f * (float)g.ni);
            }
            if (which) fs[f].pole_s = p; else fs[f].pole_n = p;
        }
    }
    if (which == 0 && threadIdx.x == 0) {
        float fudge = 0.0f;
        if (need_minmax) {
            float valmin = ord2f(scratch[2 * f]), valmax = ord2f(scratch[2 * f + 1]);
            if (a.o.extrap == EZ_MAXIMUM) fudge = (float)((double)valmax + 0.05 * (double)(valmax - valmin));
            if (a.o.extrap == EZ_MINIMUM) fudge = (float)((double)valmin - 0.05 * (double)(valmax - valmin));
        }
        if (a.o.extrap == EZ_VALUE) fudge = a.o.extrap_value;
        fs[f].fudge = fudge;
    }
}

// ---------------------------------------------------------------------------------------------------
// K19 integer mask transfer, ezget_mask_zones.inc:75-103
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_mask(const float *__restrict__ x, const float *__restrict__ y, const int *__restrict__ mask_in,
                                              int *__restrict__ mask_out, int ni_out, int nj_out, int ni_in, int nj_in, int linear) {
    int k = blockIdx.x * TPB + threadIdx.x;
    int npts = ni_out * nj_out;
    if (k >= npts) return;
    int io = k % ni_out, jo = k / ni_out;
    float px = x[k], py = y[k];
    int ix = (int)px, iy = (int)py, nix = f_nint(px), niy = f_nint(py);
    const long nin = (long)ni_in * nj_in;
    int m = 1;
    if (ix < 1 || ix > ni_in || iy < 1 || iy > nj_in) {
        m = 0;
    } else {
        long q = (long)(niy - 1) * ni_in + (nix - 1);
        if (q >= 0 && q < nin && mask_in[q] == 0) m = 0;
    }
    // second pass only touches i < ni_out, j < nj_out (ezget_mask_zones.inc:91-92); neighbours are plain
    // column-major address arithmetic, storage outside the array is treated as non-zero
    if (linear && m == 1 && io < ni_out - 1 && jo < nj_out - 1) {
        long q1 = (long)(iy - 1) * ni_in + ix, q2 = (long)iy * ni_in + (ix - 1), q3 = (long)iy * ni_in + ix;
        if ((q1 < nin && mask_in[q1] == 0) || (q2 < nin && mask_in[q2] == 0) || (q3 < nin && mask_in[q3] == 0)) m = 0;
    }
    mask_out[k] = m;
}

// ez_xpnsrcgd (ez_xpngdag2 / ez_xpngdb2) + permut, ezsint.c:100-113: the logical field z(ni, j1:j2) of a hemispheric and / or
// y-inverted source, materialised once per call so that every kernel downstream (tile kernels included) sees a plain field
__global__ void __launch_bounds__(TPB) k_expand_field(GridDev g, const float *__restrict__ zin, float *__restrict__ zexp, long in_stride, long exp_stride,
                                                      int sym_sign) {
    const int f = blockIdx.y;
    FieldFetch ff = make_fetch(g, zin + (long)f * in_stride, sym_sign);
    float *o = zexp + (long)f * exp_stride;
    const long total = (long)g.ni * (g.j2 - g.j1 + 1);
    for (long k = (long)blockIdx.x * TPB + threadIdx.x; k < total; k += (long)gridDim.x * TPB) {
        const int i = (int)(k % g.ni) + 1, j = (int)(k / g.ni) + g.j1;
        o[k] = ff(i, j);
    }
}

// ---------------------------------------------------------------------------------------------------
// interp_degree = average / sph_average: ez_avg.inc:55-170, ez_avg_sph.inc:65-230
// ---------------------------------------------------------------------------------------------------
// One thread per target point.  The target cell runs half-way to its neighbours in source index space, taken from the FIRST
// row (x) and the FIRST column (y) of the gridset coordinates (the routines assume axis-aligned grids); the spherical
// variant widens the x extent by 1 / cos(latitude of the target row).  Default-real arithmetic, sums in the reference's loop
// order, no contraction (-fmad=false): bit-equal to the Fortran for ez_avg; ez_avg_sph differs by the rounding of the row
// latitude (computed by ez_calclatlon here, by c_gdllfxy_orig at (1, j) there).  A source index outside what the reference
// allocates and fills (ez_avg.inc:20-52) is undefined there and reads as 0 here.
struct AvgSrc {
    const float *z;
    int ni, nj, ext;
    __device__ __forceinline__ float operator()(int ii, int jj) const {
        if (jj < 1 || jj > nj) return 0.0f;
        int i;
        if (ext == 0) {
            if (ii < 1 || ii > ni) return 0.0f;
            i = ii;
        } else if (ext == 1) {
            if (ii < 2 - ni || ii > 2 * ni - 2) return 0.0f;
            i = ii >= ni ? ii - ni + 1 : (ii >= 1 ? ii : ii + ni - 1);
        } else {
            if (ii < 1 - ni || ii > 2 * ni) return 0.0f;
            i = ii > ni ? ii - ni : (ii >= 1 ? ii : ii + ni);
        }
        return __ldg(z + (long)(jj - 1) * ni + (i - 1));
    }
};
__global__ void __launch_bounds__(TPB) k_avg(const __grid_constant__ InterpArgs a, const float *__restrict__ row_lat, int sph) {
    const int ni_dst = a.out_ni, nj_dst = a.out_nj;
    const int n0 = a.band_j1 > 0 ? a.band_j0 * ni_dst : 0, n1 = a.band_j1 > 0 ? a.band_j1 * ni_dst : a.npts;
    const int n = n0 + blockIdx.x * TPB + threadIdx.x;
    if (n >= n1) return;
    const int j = n / ni_dst + 1, i = n - (j - 1) * ni_dst + 1;
    const int ni_src = a.src.ni, nj_src = a.src.nj;
    // x = xx(:,1), y = yy(1,:) ; if (x(1) > real(ni_src-1)) x(1) = 1.0      (ez_avg.inc:55-57)
    auto X = [&](int k) { float v = a.x[k - 1]; if (k == 1 && v > (float)(ni_src - 1)) v = 1.0f; return v; };
    auto Y = [&](int k) { return a.y[(long)(k - 1) * ni_dst]; };
    float amp = 1.0f;
    if (sph) amp = 1.0f / cosf(row_lat[(long)(j - 1) * ni_dst] * 0.017453295199f);
    float xl, xh, yl, yh;
    const float xi = X(i), yj = Y(j);
    if (i == 1) { const float h = 0.5f * (X(2) - xi); xl = sph ? xi - h * amp : xi - h; }
    else { const float h = 0.5f * (xi - X(i - 1)); xl = sph ? xi - h * amp : xi - h; }
    if (i == ni_dst) { const float h = 0.5f * (xi - X(i - 1)); xh = sph ? xi + h * amp : xi + h; }
    else { const float h = 0.5f * (X(i + 1) - xi); xh = sph ? xi + h * amp : xi + h; }
    if (j == 1) yl = fmaxf(1.0f, yj - 0.5f * (Y(2) - yj));
    else yl = yj - 0.5f * (yj - Y(j - 1));
    if (j == nj_dst) yh = fminf(1.0f * nj_src, yj + 0.5f * (yj - Y(j - 1)));
    else yh = yj + 0.5f * (Y(j + 1) - yj);
    int jstart, jend;
    if (j == 1) {
        jstart = f_ifix(yl); jend = f_nint(yh);
        if ((float)jstart > yl) jstart = jstart - 1;
    } else if (j == nj_dst) {
        jstart = f_ifix(yj); jend = f_nint(yh);
    } else {
        jstart = f_ifix(yl); jend = f_nint(yh);
        if ((0.5f + (float)jstart) < yl) jstart = jstart + 1;
    }
    int istart = f_ifix(xl), iend = f_nint(xh);
    if ((0.5f + (float)istart) < xl) istart = istart + 1;
    if (j == 1 && (float)iend < xh) iend = iend + 1;
    // sph_average on a target row at a pole: 1 / cos(lat) is ~1e7 of either sign and the cell "covers" 1e8 source columns; the
    // reference reads far outside its work array there (undefined).  Bound the walk so that no thread loops for seconds.
    if ((long)iend - (long)istart > 8L * ni_src) { istart = max(istart, 1 - 4 * ni_src); iend = min(iend, 5 * ni_src); }
    AvgSrc src;
    src.ni = ni_src; src.nj = nj_src; src.ext = a.src.ext;
    for (int f = 0; f < a.nfields; f++) {
        src.z = a.zin + (long)f * a.in_stride;
        float acc = 0.0f, total_area = 0.0f;
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
                acc = acc + src(ii, jj) * area;
            }
        }
        if (total_area != 0.0f) acc = acc / total_area;
        a.zout[n + (long)f * a.out_stride] = acc;
    }
}

// lorenzo_mask_fill.inc:15-16,43-51 with methode = 2 (the only one c_ezsint_mdm / c_ezuvint_mdm use, ez_mask.c:127-181):
// rmin = minval(fld) over the whole field, then fld = rmin wherever the transferred mask is 0
__global__ void __launch_bounds__(TPB) k_field_min(const float *__restrict__ fld, long n, unsigned int *__restrict__ ordmin) {
    float mn = INFINITY;
    for (long k = (long)blockIdx.x * TPB + threadIdx.x; k < n; k += (long)gridDim.x * TPB) mn = fminf(mn, fld[k]);
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((threadIdx.x & 31) == 0) atomicMin(ordmin, f2ord(mn));
}
__global__ void __launch_bounds__(TPB) k_mask_fill(float *__restrict__ fld, const int *__restrict__ mask, long n, const unsigned int *__restrict__ ordmin) {
    const long k = (long)blockIdx.x * TPB + threadIdx.x;
    if (k < n && mask[k] == 0) fld[k] = ord2f(*ordmin);
}

// ez_corrbgd.inc:29-52: one thread per (field, pole row), sequential fp32 sum
__global__ void k_corrbgd(float *zout, long stride, int nfields, int ni, int nj, int hem) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int f = t >> 1, which = t & 1;
    if (f >= nfields) return;
    bool doit = which == 0 ? (hem == 0 || hem == 2) : (hem == 0 || hem == 1);
    if (!doit) return;
    float *row = zout + (long)f * stride + (which == 0 ? 0 : (long)(nj - 1) * ni);
    float somme = 0.f;
    for (int i = 0; i < ni; i++) somme = somme + row[i];
    float moyenne = somme / ((float)ni * 1.0f);
    for (int i = 0; i < ni; i++) row[i] = moyenne;
}

// sort key of a target point = the source cell (and panel) its stencil starts in: threads that are neighbours after the
// sort gather from the same 32-byte sectors (scattered 'Y' cloud targets arrive in arbitrary order)
__global__ void k_cell_keys(const float *__restrict__ x, const float *__restrict__ y, const uint8_t *__restrict__ zb, int ni, int nj, int npts,
                            unsigned int *__restrict__ keys, int *__restrict__ vals) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npts) return;
    const int i = min(ni - 1, max(0, (int)x[n] - 1)), j = min(nj - 1, max(0, (int)y[n] - 1));
    // 8 x 4-cell blocks in row-major order keep both stencil rows and columns close
    unsigned int key = (unsigned int)(j >> 2) * (unsigned int)((ni + 7) >> 3) + (unsigned int)(i >> 3);
    key = key * 32u + (unsigned int)((j & 3) * 8 + (i & 7));
    if (zb && (zb[n] & ZB_YANG)) key += 0x80000000u;
    keys[n] = key;
    vals[n] = n;
}

}  // namespace

#include "ez_interp_fast.cuh"
#include "ez_interp_tiled.cuh"
#include "ez_interp_rowsep.cuh"

namespace ezk {

void expand_field(const GridDev &g, const float *zin, float *zexp, long in_stride, long exp_stride, int nfields, int sym_sign, cudaStream_t s) {
    if (nfields <= 0) return;
    const long total = (long)g.ni * (g.j2 - g.j1 + 1);
    int bx = (int)((total + (long)TPB * 4 - 1) / ((long)TPB * 4));
    if (bx < 1) bx = 1;
    if (bx > 1184) bx = 1184;
    k_expand_field<<<dim3(bx, nfields), TPB, 0, s>>>(g, zin, zexp, in_stride, exp_stride, sym_sign);
}

void field_scalars(const InterpArgs &a, int need_poles, int need_minmax, FieldScalars *fs, unsigned int *scratch, cudaStream_t s) {
    if (a.nfields <= 0) return;
    if (need_minmax) {
        k_minmax_init<<<(a.nfields + 127) / 128, 128, 0, s>>>(scratch, a.nfields);
        long total = (long)a.src.ni * (a.src.j2 - a.src.j1 + 1);
        int bx = (int)((total + (long)TPB * 8 - 1) / ((long)TPB * 8));
        if (bx < 1) bx = 1;
        if (bx > 592) bx = 592;   // 4 CTAs x 148 SMs per field is plenty for a streaming min/max
        dim3 grid(bx, a.nfields);
        k_minmax<<<grid, TPB, 0, s>>>(a, scratch);
    }
    k_field_scalars<<<2 * a.nfields, 256, 0, s>>>(a, need_poles, need_minmax, fs, scratch);
}

// permutation of the target points by source cell (device array of npts ints); returns false when scratch memory is not available
bool sort_points_by_cell(const float *x, const float *y, const uint8_t *zb, int ni, int nj, int npts, int *perm, cudaStream_t s) {
    unsigned int *keys_in = nullptr, *keys_out = nullptr;
    int *vals_in = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    bool ok = false;
    if (cudaMalloc((void **)&keys_in, (size_t)npts * 4) == cudaSuccess && cudaMalloc((void **)&keys_out, (size_t)npts * 4) == cudaSuccess &&
        cudaMalloc((void **)&vals_in, (size_t)npts * 4) == cudaSuccess) {
        k_cell_keys<<<(npts + 255) / 256, 256, 0, s>>>(x, y, zb, ni, nj, npts, keys_in, vals_in);
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, perm, npts, 0, 32, s);
        if (cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16) == cudaSuccess) {
            ok = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, perm, npts, 0, 32, s) == cudaSuccess &&
                 cudaStreamSynchronize(s) == cudaSuccess;
        }
    }
    if (!ok) (void)cudaGetLastError();
    cudaFree(keys_in); cudaFree(keys_out); cudaFree(vals_in); cudaFree(tmp);
    return ok;
}

bool direct_eligible(const InterpArgs &a) { return direct_args_ok(a); }

bool tiled_eligible(const InterpArgs &a) {
    if (a.o.arith != 2 || (a.o.degree != EZ_LINEAR && a.o.degree != EZ_CUBIC)) return false;
    if (a.out_ni < 2 || a.out_nj < 2) return false;
    if (a.src.xpn != XPN_NONE || a.src.yinv) return false;
    if (a.yy && (a.src2.xpn != XPN_NONE || a.src2.yinv)) return false;
    // the tile kernels address a pipeline stage (LPS consecutive fields) through signed 32-bit byte offsets
    const long field = a.yy ? 2 * a.panel_stride : (long)a.src.ni * (a.src.j2 - a.src.j1 + 1);
    const long span = (long)LPS * (a.in_stride < 0 ? -a.in_stride : a.in_stride) + field;
    if (span * 4 >= (1L << 31)) return false;
    return true;
}

void tile_dims(int out_ni, int out_nj, int *ntx, int *nty) {
    *ntx = (out_ni + TO - 1) / TO;
    *nty = (out_nj + TO - 1) / TO;
}

void tile_classify(const InterpArgs &a, TileDesc *tiles, uint32_t *flags, unsigned int *rowy, unsigned int *counters, cudaStream_t s) {
    int ntx, nty;
    tile_dims(a.out_ni, a.out_nj, &ntx, &nty);
    dim3 grid(ntx, nty);
    if (a.o.degree == EZ_CUBIC) k_tile_classify<3><<<grid, TT * TT, 0, s>>>(a, tiles, flags, rowy, counters);
    else k_tile_classify<1><<<grid, TT * TT, 0, s>>>(a, tiles, flags, rowy, counters);
}

void tile_list(const InterpArgs &a, const uint32_t *flags, int *list, unsigned int *cursor, int mode, cudaStream_t s) {
    int ntx, nty;
    tile_dims(a.out_ni, a.out_nj, &ntx, &nty);
    dim3 grid(ntx, nty);
    k_tile_list<<<grid, TT * TT, 0, s>>>(a, flags, list, cursor, mode);
}

void plan_cache(const InterpArgs &a, int *idx, float4 *wx, float4 *wy, float *sm1, cudaStream_t s, int mark_split) {
    const int nb = (a.npts + 255) / 256;
    if (a.o.degree == EZ_CUBIC) k_plan_cache<3><<<nb, 256, 0, s>>>(a, idx, wx, wy, sm1, mark_split);
    else k_plan_cache<1><<<nb, 256, 0, s>>>(a, idx, wx, wy, sm1, mark_split);
}

int interp(const InterpArgs &a, cudaStream_t s, const InterpTiming *timing, const InterpFork *fork) {
    if (a.npts <= 0 || a.nfields <= 0) return 0;
    if (a.tiles) {
        // tiled path: the two tile kernels for every point that fits, the per-point kernel for the compacted rest
        int ntx, nty, launches = 0;
        tile_dims(a.out_ni, a.out_nj, &ntx, &nty);
        dim3 grid(ntx, nty);
        static const size_t smem_pad = getenv("EZB200_WIN_SMEM_PAD") ? (size_t)atol(getenv("EZB200_WIN_SMEM_PAD")) : 0;   // occupancy experiments
        const int nstages = (a.nfields + LPS - 1) / LPS;
        const size_t smem = (size_t)(nstages < NSTAGE ? nstages : NSTAGE) * LPS * a.tile_cap * sizeof(float) + smem_pad;   // slots beyond the last stage are never touched
        static const bool few_on = !(getenv("EZB200_TILED_FEW") && atoi(getenv("EZB200_TILED_FEW")) == 0);
        const bool few = few_on && nstages == 1;
        // the window kernel: one CTA per tile of the compacted list
        const int win_ctas = a.ntiles_window;
        const dim3 grid_win(win_ctas > 0 ? win_ctas : 1);
        // row-separable kernel: staged pitch congruent to the source row length mod 4 (16-byte cp.async, ez_interp_rowsep.cuh)
        const int rs_p = rs_pitch(a.src.ni);
        InterpArgs ar = a;
        ar.tile_cap_rs = rs_cap(rs_p, a.tile_cap_rs);   // a.tile_cap_rs arrives as the tallest box height
        const size_t smem_rs = (size_t)RS_NSTAGE * LPS * ar.tile_cap_rs * sizeof(float) + 2 * (size_t)TO * RS_VP * sizeof(float4) + TO * sizeof(double);
        static bool attr_set_dev[64] = {false};   // cudaFuncSetAttribute applies to the current device only
        int cur_dev = 0;
        cudaGetDevice(&cur_dev);
        bool &attr_set = attr_set_dev[cur_dev & 63];
        if (!attr_set) {
            const int mx = 227 * 1024;
            cudaFuncSetAttribute(k_interp_tiled<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            cudaFuncSetAttribute(k_interp_tiled<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            cudaFuncSetAttribute(k_interp_tiled_few<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            cudaFuncSetAttribute(k_interp_tiled_few<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            const int mxr = RS_NSTAGE * LPS * (TILE_CAP + 16) * (int)sizeof(float) + 2 * TO * RS_VP * (int)sizeof(float4) + TO * (int)sizeof(double);
#define EZ_RS_ATTR(P)                                                                                       \
    cudaFuncSetAttribute(k_interp_rowsep<3, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, mxr); \
    cudaFuncSetAttribute(k_interp_rowsep<1, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, mxr);
            EZ_RS_ATTR(35) EZ_RS_ATTR(36) EZ_RS_ATTR(37) EZ_RS_ATTR(38)
#undef EZ_RS_ATTR
            attr_set = true;
        }
        // Concurrent mode (no per-kernel timing): the window kernel, the row-separable kernel and the list kernel write
        // disjoint target points; with two or more of them they go to separate streams between a fork and a join event.
        const int nk = (a.ntiles_rowsep > 0) + (a.ntiles_window > 0) + (a.nlist > 0 || a.nclist > 0);
        const bool conc = fork && !timing && nk >= 2;
        cudaStream_t s_rs = s, s_list = s;
        if (conc) {
            cudaEventRecord(fork->ev_fork, s);
            if (a.ntiles_window > 0 && a.ntiles_rowsep > 0) { s_rs = fork->aux[0]; cudaStreamWaitEvent(s_rs, fork->ev_fork, 0); }
            if (a.nlist > 0 || a.nclist > 0) { s_list = fork->aux[1]; cudaStreamWaitEvent(s_list, fork->ev_fork, 0); }
        }
        if (timing) cudaEventRecord(timing->ev[0], s);
        if (a.ntiles_window > 0 && conc) {   // the largest kernel first
            if (few) {
                if (a.o.degree == EZ_CUBIC) k_interp_tiled_few<3><<<grid_win, NTHR_WIN, smem, s>>>(a);
                else k_interp_tiled_few<1><<<grid_win, NTHR_WIN, smem, s>>>(a);
            } else if (a.o.degree == EZ_CUBIC) k_interp_tiled<3><<<grid_win, NTHR_WIN, smem, s>>>(a);
            else k_interp_tiled<1><<<grid_win, NTHR_WIN, smem, s>>>(a);
            launches++;
        }
        if (a.ntiles_rowsep > 0) {
            const bool cub = a.o.degree == EZ_CUBIC;
#define EZ_RS_LAUNCH(P)                                                               \
    case P:                                                                           \
        if (cub) k_interp_rowsep<3, P><<<grid, RS_NTHR, smem_rs, s_rs>>>(ar);   \
        else k_interp_rowsep<1, P><<<grid, RS_NTHR, smem_rs, s_rs>>>(ar);       \
        break;
            switch (rs_p) { EZ_RS_LAUNCH(35) EZ_RS_LAUNCH(36) EZ_RS_LAUNCH(37) EZ_RS_LAUNCH(38) }
#undef EZ_RS_LAUNCH
            launches++;
        }
        if (timing) cudaEventRecord(timing->ev[1], s);
        if (a.ntiles_window > 0 && !conc) {
            if (few) {
                if (a.o.degree == EZ_CUBIC) k_interp_tiled_few<3><<<grid_win, NTHR_WIN, smem, s>>>(a);
                else k_interp_tiled_few<1><<<grid_win, NTHR_WIN, smem, s>>>(a);
            } else if (a.o.degree == EZ_CUBIC) k_interp_tiled<3><<<grid_win, NTHR_WIN, smem, s>>>(a);
            else k_interp_tiled<1><<<grid_win, NTHR_WIN, smem, s>>>(a);
            launches++;
        }
        if (timing) cudaEventRecord(timing->ev[2], s);
        if (a.nclist > 0) {
            k_store_const<<<(a.nclist + 255) / 256, 256, 0, s_list>>>(a, a.clist, a.nclist);
            launches++;
        }
        if (a.nlist > 0) {
            InterpArgs b = a;
            b.tiles = nullptr;
            b.tile_flags = nullptr;
            launch_fast(b, s_list);
            launches++;
        }
        if (conc) {
            if (s_rs != s) { cudaEventRecord(fork->ev_join[0], s_rs); cudaStreamWaitEvent(s, fork->ev_join[0], 0); }
            if (s_list != s) { cudaEventRecord(fork->ev_join[1], s_list); cudaStreamWaitEvent(s, fork->ev_join[1], 0); }
        }
        if (timing) cudaEventRecord(timing->ev[3], s);
        return launches;
    }
    if (a.o.arith >= 1 && launch_fast(a, s)) return 1;
    int nb = (a.npts + TPB - 1) / TPB;
    k_interp_ref<<<nb, TPB, 0, s>>>(a);
    return 1;
}

int interp_direct(const InterpArgs &a, const int *zlist, int nzlist, cudaStream_t s, const InterpFork *fork, const DirectPre *pre,
                  const InterpTiming *timing) {
    if (a.npts <= 0 || a.nfields <= 0) return 0;
    const int npts = a.band_j1 > 0 ? (a.band_j1 - a.band_j0) * a.out_ni : a.npts;
    int launches = 0;
    // The per-field scalars (pole means, extrapolation value) and the zone points that read them are small, latency-bound
    // launches: with fork resources they run beside the direct kernel (disjoint target points) instead of around it.  A
    // constant extrapolation value is read by the direct kernel itself: then the scalars stay in front of it.
    const bool direct_reads_fs = a.fs && !a.o.vector && a.o.extrap >= EZ_MAXIMUM && (!pre || pre->has_dehors);
    const bool conc = fork && !timing && nzlist > 0 && !direct_reads_fs;   // (per-kernel timing: everything in line on s)
    cudaStream_t s_list = s;
    if (timing) cudaEventRecord(timing->ev[0], s);
    if (conc) {
        cudaEventRecord(fork->ev_fork, s);
        s_list = fork->aux[1];
        cudaStreamWaitEvent(s_list, fork->ev_fork, 0);
    }
    if (pre && pre->run) {
        field_scalars(a, pre->need_poles, pre->need_minmax, pre->fs, pre->scratch, conc ? s_list : s);
        launches += pre->need_minmax ? 3 : 1;
    }
    if (timing) cudaEventRecord(timing->ev[1], s);
    if (a.o.degree == EZ_CUBIC) {
        constexpr int per = TPB * DirectCfg<3>::PPT;
        k_interp_direct<3><<<(npts + per - 1) / per, TPB, 0, s>>>(a);
    } else {
        constexpr int per = TPB * DirectCfg<1>::PPT;
        k_interp_direct<1><<<(npts + per - 1) / per, TPB, 0, s>>>(a);
    }
    launches++;
    if (timing) cudaEventRecord(timing->ev[2], s);
    if (nzlist > 0) {
        InterpArgs b = a;
        b.list = zlist;
        b.nlist = nzlist;
        b.wrot = nullptr;
        launch_fast(b, s_list);
        launches++;
    }
    if (timing) cudaEventRecord(timing->ev[3], s);
    if (conc) { cudaEventRecord(fork->ev_join[1], s_list); cudaStreamWaitEvent(s, fork->ev_join[1], 0); }
    return launches;
}

void zone_list(const uint8_t *zb, int npts, int skip_dehors_only, int *list, unsigned int *cursor, cudaStream_t s, const int *order) {
    if (npts <= 0) return;
    k_zone_list<<<(npts + TPB - 1) / TPB, TPB, 0, s>>>(zb, npts, skip_dehors_only, list, cursor, order);
}

int interp_avg(const InterpArgs &a, const float *row_lat, int sph, const int *zlist, int nzlist, cudaStream_t s) {
    if (a.npts <= 0 || a.nfields <= 0) return 0;
    const int npts = a.band_j1 > 0 ? (a.band_j1 - a.band_j0) * a.out_ni : a.npts;
    k_avg<<<(npts + TPB - 1) / TPB, TPB, 0, s>>>(a, row_lat, sph);
    if (nzlist <= 0) return 1;
    // ez_corrval.c:54-143 on the points with zone bits (k_interp_fast's average branch)
    InterpArgs b = a;
    b.list = zlist;
    b.nlist = nzlist;
    launch_fast(b, s);
    return 2;
}

void mask_fill_min(float *fld, const int *mask, long n, unsigned int *scratch, cudaStream_t s) {
    if (n <= 0) return;
    k_minmax_init<<<1, 32, 0, s>>>(scratch, 1);   // scratch[0] = "no minimum yet"
    int nb = (int)((n + (long)TPB * 8 - 1) / ((long)TPB * 8));
    if (nb > 592) nb = 592;
    k_field_min<<<nb, TPB, 0, s>>>(fld, n, scratch);
    k_mask_fill<<<(unsigned int)((n + TPB - 1) / TPB), TPB, 0, s>>>(fld, mask, n, scratch);
}

void mask_transfer(const float *x, const float *y, const int *mask_in, int *mask_out, int ni_out, int nj_out, int ni_in, int nj_in,
                   int linear, cudaStream_t s) {
    int npts = ni_out * nj_out;
    if (npts <= 0) return;
    k_mask<<<(npts + TPB - 1) / TPB, TPB, 0, s>>>(x, y, mask_in, mask_out, ni_out, nj_out, ni_in, nj_in, linear);
}

void corrbgd(float *zout, long stride, int nfields, int ni, int nj, int hem, cudaStream_t s) {
    int nt = 2 * nfields;
    if (nt <= 0) return;
    k_corrbgd<<<(nt + 63) / 64, 64, 0, s>>>(zout, stride, nfields, ni, nj, hem);
}

}  // namespace ezk
