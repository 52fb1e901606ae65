// librmn_b200/csrc/ez_interp_fast.cuh -- the production interpolation kernel ("arith = 1").
// Included by ez_interp.cu inside its translation unit (uses its evaluators for the rare zone points).
//
// One thread per target point.  A per-point PLAN (element offsets of the stencil and the interpolation
// weights) is computed once from (x, y); the level loop then only gathers, multiplies and stores:
//   nearest : 1 load                                -- bit-identical to ez_rgdint_0
//   linear  : 4 loads + 3 zlin in fp64              -- the reference's own operation order (ez_rgdint_1_*,
//                                                     ez_irgdint_1_*), only hoisted out of the level loop
//   cubic   : 16 loads + 20 fp64 FMAs               -- weight form: zo = sum_b wy_b * (sum_a wx_a * z_ab);
//             wx, wy are the cubic8.cdk polynomial (regular grids) or the Newton basis built from the same
//             cx/cy tables (irregular grids), evaluated in fp64.  Rounding differs from the reference's
//             Horner/Newton order at the 1e-16 (fp64) level -- or 1e-7 of the field magnitude for
//             ez_irgdint_3_nw, whose reference form rounds partial results to fp32 -- far inside the
//             1e-5-of-range tolerance.
// Why fp64 and not fp32: 20 DFMA/point-level keep the FP64 pipe under the L1 gather bound (64 B/point at
// 128 B/clk/SM), so fp32 would buy nothing while costing accuracy on large-offset fields.
#pragma once

namespace {

struct PlanLin {
    int o00, o10, o01, o11;   // element offsets of (i,j) (ip1,j) (i,j+1) (ip1,j+1)
    double dx, dy;
};
struct PlanCub {
    int c[4];                 // element offsets of the four columns within a row (0-based)
    int r0;                   // element offset of row j-1
    int ni;                   // row pitch
    double wx[4], wy[4];
};

__device__ __forceinline__ bool is_irregular(const GridDev &g) { return g.grtyp == '#' || g.grtyp == 'Z' || g.grtyp == 'G'; }

// nearest: ez_rgdint_0.inc:27-32
__device__ __forceinline__ int plan_nearest(const GridDev &g, float px, float py) {
    int i = min(g.ni, max(1, f_nint(px)));
    int j = min(g.j2, max(g.j1, f_nint(py)));
    return (j - g.j1) * g.ni + (i - 1);
}

// linear: index and dx/dy logic of ez_rgdint_1_nw/_w and ez_irgdint_1_nw/_w
__device__ __forceinline__ PlanLin plan_linear(const GridDev &g, float px, float py) {
    PlanLin p;
    const int ni = g.ni, j1 = g.j1, j2 = g.j2;
    int i, ip1, j;
    if (is_irregular(g)) {
        if (g.ext == 0) {
            i = min(ni - 1, max(1, f_ifix(px)));
            j = min(g.nj - 1, max(1, f_ifix(py)));
            ip1 = i + 1;
            float axi = __ldg(g.ax + i - 1), axi1 = __ldg(g.ax + i), ayj = __ldg(g.ay + j - 1), ayj1 = __ldg(g.ay + j);
            double x1 = axi, x2 = axi1;
            double x = (double)axi + (x2 - x1) * (double)(px - (float)i);
            double y = (double)(ayj + (ayj1 - ayj) * (py - (float)j));
            p.dx = (x - x1) / (x2 - x1);
            p.dy = (y - (double)ayj) / (double)(ayj1 - ayj);
        } else {
            const int wrap = g.ext, limite = ni + 2 - wrap;
            i = min(ni - 2 + wrap, max(1, f_ifix(px)));
            j = min(j2 - 1, max(j1 + 1, f_ifix(py)));
            if (j < 0) j = j - 1;
            ip1 = i + 1;
            double x1 = __ldg(g.ax + i - 1), x2 = 0.0;
            if (ip1 <= ni) x2 = __ldg(g.ax + ip1 - 1);
            if (i == (ni - 2 + wrap)) {
                ip1 = (limite + i + 1) % limite;
                x2 = (double)(__ldg(g.ax + 1) + __ldg(g.ax + ni - 1));
            }
            float ayj = __ldg(g.ay + (j - j1)), ayj1 = __ldg(g.ay + (j + 1 - j1));
            double x = x1 + (x2 - x1) * (double)(px - (float)i);
            double y = (double)(ayj + (ayj1 - ayj) * (py - (float)j));
            p.dx = (x - x1) / (x2 - x1);
            p.dy = (y - (double)ayj) / (double)(ayj1 - ayj);
        }
    } else {
        if (g.ext == 2) {
            const int wrap = 2, limite = ni + 2 - wrap;
            i = min(ni - 2 + wrap, max(1, f_ifix(px)));
            ip1 = i + 1;
            if (i == (ni - 2 + wrap)) ip1 = (limite + i + 1) % limite;
        } else {
            i = min(ni - 1, max(1, f_ifix(px)));
            ip1 = i + 1;
        }
        j = min(j2 - 1, max(j1, f_ifix(py)));
        p.dx = (double)(px - (float)i);
        p.dy = (double)(py - (float)j);
    }
    const int r = (j - j1) * ni;
    p.o00 = r + i - 1;
    p.o10 = r + ip1 - 1;
    p.o01 = r + ni + i - 1;
    p.o11 = r + ni + ip1 - 1;
    return p;
}

// Newton-basis weights for nodes with reciprocal tables c[0..5] and offsets t1 = x-x1, t2 = x-x2, t3 = x-x3:
//   b = z1 + t1 (a2 + t2 (a3 + a4 t3))   with a2..a4 as in ez_irgdint_3_nw.inc:85-87
__device__ __forceinline__ void newton_weights(const double c[6], double t1, double t2, double t3, double w[4]) {
    const double P1 = t1, P2 = t1 * t2, P3 = P2 * t3;
    const double c21 = c[1] * c[0], s31 = c[2] + c[0];
    const double k2 = c[1] * s31;                       // c2 (c3 + c1)
    const double k3 = c[1] * c[2];                      // c2 c3
    w[0] = 1.0 - P1 * c[0] + P2 * c21 - P3 * (c[3] * c21);
    w[1] = P1 * c[0] - P2 * k2 + P3 * (c[3] * (c[4] * c[2] + k2));
    w[2] = P2 * k3 - P3 * (c[3] * (c[4] * (c[5] + c[2]) + k3));
    w[3] = P3 * (c[3] * c[4] * c[5]);
}

// Weight form of the same interpolant (the Newton basis expanded in the samples, fp64): 16 samples, 20 DFMA per level
// instead of five Newton evaluations.  Differs from the reference's evaluation order at the 1e-16 level.
struct PolarStripW {
    int jbase, pole_row, jsrc0;
    int i, im1, ip1, ip2, j;
    double wx[4], wy[4];
};
__device__ __forceinline__ PolarStripW polar_strip_weights(const PolarStrip &ps) {
    PolarStripW w;
    w.jbase = ps.jbase; w.pole_row = ps.pole_row; w.jsrc0 = ps.jsrc0;
    const WnncPlan &p = ps.plan;
    w.i = p.i; w.im1 = p.im1; w.ip1 = p.ip1; w.ip2 = p.ip2; w.j = p.j;
    newton_weights(p.cx, p.x - p.x1, p.x - p.x2, p.x - p.x3, w.wx);
    newton_weights(p.cy, p.y - p.y1, p.y - p.y2, p.y - p.y3, w.wy);
    return w;
}
__device__ __forceinline__ float polar_strip_eval_w(const PolarStripW &w, const FieldFetch &ff, float poleval, const float *polerow) {
    StripFetch s;
    s.f = ff; s.poleval = poleval; s.polerow = polerow;
    s.jbase = w.jbase; s.pole_row = w.pole_row; s.jsrc0 = w.jsrc0;
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int jj = w.j - 1 + r;
        double t = w.wx[0] * (double)s(w.im1, jj);
        t = fma(w.wx[1], (double)s(w.i, jj), t);
        t = fma(w.wx[2], (double)s(w.ip1, jj), t);
        t = fma(w.wx[3], (double)s(w.ip2, jj), t);
        acc = fma(w.wy[r], t, acc);
    }
    return (float)acc;
}

// cubic8.cdk:3 expanded in the samples (fp32-rounded literals kept)
__device__ __forceinline__ void cubic_weights(double d, double w[4]) {
    const double c6 = (double)0.1666666666666f, c3 = (double)0.3333333333333f, h = 0.5;
    w[0] = ((-c6 * d + h) * d - c3) * d;
    w[1] = ((h * d - 1.0) * d - h) * d + 1.0;
    w[2] = ((-h * d + h) * d + 1.0) * d;
    w[3] = (c6 * d * d - c6) * d;
}

// PART: 0 = both axes ; 1 = the x side only (c[], wx ; py ignored) ; 2 = the y side only (r0, wy ; px ignored).
// The two sides are independent in the reference too (ez_rgdint_3_*, ez_irgdint_3_*): the split only skips work.
template <int PART = 0>
__device__ __forceinline__ PlanCub plan_cubic(const GridDev &g, float px, float py) {
    PlanCub p;
    constexpr bool DOX = PART != 2, DOY = PART != 1;
    const int ni = g.ni, j1 = g.j1, j2 = g.j2;
    p.ni = ni;
    int i = 1, im1 = 1, ip1 = 1, ip2 = 1;
    const int j = DOY ? min(j2 - 2, max(j1 + 1, f_ifix(py))) : j1 + 1;
#pragma unroll
    for (int k = 0; k < 4; k++) { p.wx[k] = 0.0; p.wy[k] = 0.0; }
    if (is_irregular(g)) {
        if (DOY) {
            const int ny = j2 - j1 + 1;
            const float *ay = g.ay - j1;
            const float aym = __ldg(ay + j - 1), ay0 = __ldg(ay + j), ayp = __ldg(ay + j + 1);
            const double y = (double)(ay0 + (ayp - ay0) * (py - (float)j));
            double cy[6];
#pragma unroll
            for (int k = 0; k < 6; k++) cy[k] = (double)__ldg(g.cy + (long)k * ny + (j - j1));
            newton_weights(cy, y - (double)aym, y - (double)ay0, y - (double)ayp, p.wy);
        }
        if (DOX) {
            double cx[6];
            float x1, x2, x3, x4;
            int nxt;
            if (g.ext == 0) {
                i = min(g.i2 - 2, max(g.i1 + 1, f_ifix(px)));
                im1 = i - 1; ip1 = i + 1; ip2 = i + 2;
                x1 = __ldg(g.ax + im1 - 1); x2 = __ldg(g.ax + i - 1); x3 = __ldg(g.ax + ip1 - 1); x4 = __ldg(g.ax + ip2 - 1);
                nxt = g.i2 - g.i1 + 1;
            } else {
                const int wrap = g.ext;
                i = min(ni - 2 + wrap, max(1, max(2 - wrap, f_ifix(px))));
                seam_x(wrap, ni, g.ax, i, false, im1, ip1, ip2, x1, x2, x3, x4);
                nxt = ni;
            }
            (void)x4;
            const double x = (double)(x2 + (x3 - x2) * (px - (float)i));
#pragma unroll
            for (int k = 0; k < 6; k++) cx[k] = (double)__ldg(g.cx + (long)k * nxt + (i - 1));
            newton_weights(cx, x - (double)x1, x - (double)x2, x - (double)x3, p.wx);
        }
    } else {
        if (DOX) {
            rgd3_cols(g.ext == 0 ? 0 : 1, ni, g.ext, px, i, im1, ip1, ip2);
            cubic_weights((double)(px - (float)i), p.wx);
        }
        if (DOY) cubic_weights((double)(py - (float)j), p.wy);
    }
    p.c[0] = im1 - 1; p.c[1] = i - 1; p.c[2] = ip1 - 1; p.c[3] = ip2 - 1;
    p.r0 = (j - 1 - j1) * ni;
    return p;
}

__device__ __forceinline__ float eval_cubic(const PlanCub &p, const float *__restrict__ z) {
    const float *r = z + p.r0;
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const float z0 = __ldg(r + p.c[0]), z1 = __ldg(r + p.c[1]), z2 = __ldg(r + p.c[2]), z3 = __ldg(r + p.c[3]);
        double s = p.wx[0] * (double)z0;
        s = fma(p.wx[1], (double)z1, s);
        s = fma(p.wx[2], (double)z2, s);
        s = fma(p.wx[3], (double)z3, s);
        acc = fma(p.wy[b], s, acc);
        r += p.ni;
    }
    return (float)acc;
}

// fp32 "difference form" of the same weight-form cubic (arith = 2): with zc = z(i,j) the stencil centre,
//   zo = zc + [ sum_b wy_b sum_a wx_a (z_ab - zc) + zc * ((sum wx)(sum wy) - 1) ]
// The differences are exact or nearly so (Sterbenz), so the fp32 rounding error scales with the LOCAL
// variation of the field (<= its range), not with its magnitude: ~3e-7 of the range, with no fp32->fp64
// conversions (16/clk/SM on this part) and no FP64 pipe use in the level loop.
struct PlanCubF {
    int c[4];
    int r0, ni;
    float wx[4], wy[4];
    float sm1;                // (sum wx)(sum wy) - 1, evaluated in fp64
};
__device__ __forceinline__ PlanCubF to_f32(const PlanCub &p) {
    PlanCubF q;
#pragma unroll
    for (int k = 0; k < 4; k++) { q.c[k] = p.c[k]; q.wx[k] = (float)p.wx[k]; q.wy[k] = (float)p.wy[k]; }
    q.r0 = p.r0; q.ni = p.ni;
    q.sm1 = (float)((p.wx[0] + p.wx[1] + p.wx[2] + p.wx[3]) * (p.wy[0] + p.wy[1] + p.wy[2] + p.wy[3]) - 1.0);
    return q;
}
__device__ __forceinline__ float eval_cubic_f32(const PlanCubF &p, const float *__restrict__ z) {
    const float *r = z + p.r0;
    float v[4][4];
#pragma unroll
    for (int b = 0; b < 4; b++) {
#pragma unroll
        for (int a = 0; a < 4; a++) v[b][a] = __ldg(r + p.c[a]);
        r += p.ni;
    }
    const float zc = v[1][1];
    float acc = 0.f;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        float s = p.wx[0] * (v[b][0] - zc);
        s = fmaf(p.wx[1], v[b][1] - zc, s);
        s = fmaf(p.wx[2], v[b][2] - zc, s);
        s = fmaf(p.wx[3], v[b][3] - zc, s);
        acc = fmaf(p.wy[b], s, acc);
    }
    return zc + fmaf(zc, p.sm1, acc);
}

__device__ __forceinline__ float eval_linear(const PlanLin &p, const float *__restrict__ z) {
    const double z00 = (double)__ldg(z + p.o00), z10 = (double)__ldg(z + p.o10);
    const double z01 = (double)__ldg(z + p.o01), z11 = (double)__ldg(z + p.o11);
    const double y2 = zlin(z00, z10, p.dx);
    const double y3 = zlin(z01, z11, p.dx);
    return (float)zlin(y2, y3, p.dy);
}

// fp32 difference form of the bilinear interpolant about z00 (arith = 2): the four weights sum to exactly one,
// so zo = z00 + w10 (z10 - z00) + w01 (z01 - z00) + w11 (z11 - z00) with fp64-derived weights.
struct PlanLinF {
    int o00, o10, o01, o11;
    float w10, w01, w11;
};
__device__ __forceinline__ PlanLinF to_f32(const PlanLin &p) {
    PlanLinF q;
    q.o00 = p.o00; q.o10 = p.o10; q.o01 = p.o01; q.o11 = p.o11;
    q.w10 = (float)(p.dx * (1.0 - p.dy));
    q.w01 = (float)((1.0 - p.dx) * p.dy);
    q.w11 = (float)(p.dx * p.dy);
    return q;
}
__device__ __forceinline__ float eval_linear_f32(const PlanLinF &p, const float *__restrict__ z) {
    const float z00 = __ldg(z + p.o00), z10 = __ldg(z + p.o10), z01 = __ldg(z + p.o01), z11 = __ldg(z + p.o11);
    float acc = p.w10 * (z10 - z00);
    acc = fmaf(p.w01, z01 - z00, acc);
    acc = fmaf(p.w11, z11 - z00, acc);
    return z00 + acc;
}

// DEG = 0 nearest, 1 linear (fp64, reference order), 3 cubic (fp64 weight form), 4 cubic (fp32 difference form),
// 5 linear (fp32 difference form).  Requires plain field access (no hemispheric mirror, no y inversion).
// The grid descriptor is taken by reference straight from kernel-parameter (constant-bank) space: the Yin/Yang
// choice is a branch around two instantiations rather than a pointer select, which would force a local copy.
#ifndef EZ_LIST_LCH
#define EZ_LIST_LCH 16
#endif
constexpr int LIST_LCH = EZ_LIST_LCH;   // levels per thread when the kernel runs on the compacted list (blockIdx.y picks the chunk)

template <int DEG>
__device__ __forceinline__ void point_body(const GridDev &g, const InterpArgs &a, const float *__restrict__ z, int n, float px, float py,
                                           uint8_t zbits) {
    const long is = a.in_stride, os = a.out_stride;
    // the compacted list is short and its points are expensive (zone points): levels are split over blockIdx.y
    const int lch = a.list_lch > 0 ? a.list_lch : LIST_LCH;
    const int f0 = a.list ? blockIdx.y * lch : 0;
    const int nf = a.list ? min(lch, a.nfields - f0) : a.nfields;
    float *zo = a.zout + n + f0 * os;
    z += f0 * is;
    if (zbits == 0) {
        if (DEG == 0) {
            const int o = plan_nearest(g, px, py);
#pragma unroll 4
            for (int f = 0; f < nf; f++) zo[f * os] = __ldg(z + f * is + o);
        } else if (DEG == 1) {
            const PlanLin p = plan_linear(g, px, py);
#pragma unroll 2
            for (int f = 0; f < nf; f++) zo[f * os] = eval_linear(p, z + f * is);
        } else if (DEG == 3) {
            const PlanCub p = plan_cubic(g, px, py);
#pragma unroll 2
            for (int f = 0; f < nf; f++) zo[f * os] = eval_cubic(p, z + f * is);
        } else if (DEG == 5) {
            const PlanLinF p = to_f32(plan_linear(g, px, py));
#pragma unroll 4
            for (int f = 0; f < nf; f++) zo[f * os] = eval_linear_f32(p, z + f * is);
        } else {
            const PlanCubF p = to_f32(plan_cubic(g, px, py));
#pragma unroll 2
            for (int f = 0; f < nf; f++) zo[f * os] = eval_cubic_f32(p, z + f * is);
        }
        return;
    }
    // interp_degree = average / sph_average: k_avg has written every point; ez_corrval.c:54-143 then rewrites the points outside
    // the source (constant, or interpolated with extrap_degree) and, for scalars, the points at a pole.  Its polar-strip
    // rewrites (ez_corrval_aunord.c:53-112, ez_corrval_ausud.c:64-129) have no case for these degrees and store an uninitialised
    // work array in the reference: here those points keep their average.
    if (a.o.degree == EZ_AVERAGE || a.o.degree == EZ_SPH_AVERAGE) {
        for (int f = 0; f < nf; f++) {
            bool write = false;
            float v = 0.f;
            if (zbits & ZB_DEHORS) {
                if (a.o.extrap >= EZ_MAXIMUM) v = (a.o.vector || !a.fs) ? 0.0f : a.fs[f0 + f].fudge;
                else v = interp_point(make_fetch(g, z + f * is, a.o.sym_sign), g, a.o.extrap, px, py);
                write = true;
            }
            if (!a.o.vector && a.fs) {
                if (zbits & ZB_POLE_NORD) { v = a.fs[f0 + f].pole_n; write = true; }
                if (zbits & ZB_POLE_SUD) { v = a.fs[f0 + f].pole_s; write = true; }
            }
            if (write) zo[f * os] = v;
        }
        return;
    }
    // outside the source with a constant extrapolation (ez_corrval.c:56-93: maximum / minimum / value; 0.0 for vectors, :64-67):
    // a plain store -- a LAM source leaves a third of a larger target here (C3), and DEHORS excludes the polar zones (ez_defzones.c:88-109)
    if (zbits == ZB_DEHORS && a.o.extrap >= EZ_MAXIMUM) {
        for (int f = 0; f < nf; f++) zo[f * os] = (a.o.vector || !a.fs) ? 0.0f : a.fs[f0 + f].fudge;
        return;
    }
    // rare: a point in a polar / outside zone -- reference-order evaluators (ez_corrval semantics)
    const bool au_n = zbits & ZB_AU_NORD, au_s = zbits & ZB_AU_SUD, pole = zbits & (ZB_POLE_NORD | ZB_POLE_SUD);
    const bool irr = (g.grtyp == 'Z' || g.grtyp == 'G');
    // common heavy case: exactly one polar strip, cubic on an irregular source -> strip geometry and fp64 reciprocals once per point
    if ((au_n != au_s) && !(zbits & ZB_DEHORS) && a.o.degree == EZ_CUBIC && irr && (a.o.vector ? a.polerow != nullptr : true)) {
        if (!a.o.vector && pole) {   // ez_corrval.c:130-143: the pole lists are written last
            for (int f = 0; f < nf; f++) zo[f * os] = (zbits & ZB_POLE_SUD) ? a.fs[f0 + f].pole_s : a.fs[f0 + f].pole_n;
            return;
        }
        if (a.o.arith >= 1) {   // weight form (arith 0 keeps the reference's evaluation order)
            const PolarStripW pw = polar_strip_weights(polar_strip_plan(g, au_n, px, py));
            for (int f = 0; f < nf; f++) {
                const FieldFetch ff = make_fetch(g, z + f * is, a.o.sym_sign);
                float pv = 0.f;
                const float *prow = nullptr;
                if (a.o.vector) prow = a.polerow + (long)(f0 + f) * 2 * g.ni + (au_n ? 0 : g.ni);
                else pv = au_n ? a.fs[f0 + f].pole_n : a.fs[f0 + f].pole_s;
                zo[f * os] = polar_strip_eval_w(pw, ff, pv, prow);
            }
            return;
        }
        const PolarStrip ps = polar_strip_plan(g, au_n, px, py);
        for (int f = 0; f < nf; f++) {
            const FieldFetch ff = make_fetch(g, z + f * is, a.o.sym_sign);
            float pv = 0.f;
            const float *prow = nullptr;
            if (a.o.vector) prow = a.polerow + (long)(f0 + f) * 2 * g.ni + (au_n ? 0 : g.ni);
            else pv = au_n ? a.fs[f0 + f].pole_n : a.fs[f0 + f].pole_s;
            zo[f * os] = polar_strip_eval(ps, ff, pv, prow);
        }
        return;
    }
    // the interpolated value is only needed when no zone rewrites it
    const bool need_normal = !(zbits & ZB_DEHORS) && (a.o.vector ? !((au_n || au_s) && a.polerow) : !(au_n || au_s || pole));
    for (int f = 0; f < nf; f++) {
        FieldFetch ff = make_fetch(g, z + f * is, a.o.sym_sign);
        float v = 0.f;
        if (need_normal) v = interp_point(ff, g, a.o.degree, px, py);
        zo[f * os] = zone_point(ff, g, a.o, zbits, a.fs, a.polerow, f0 + f, px, py, v);
    }
}

template <int DEG>
__global__ void __launch_bounds__(TPB) k_interp_fast(const __grid_constant__ InterpArgs a) {
    // with an index list (the points the tile kernel leaves out) thread t owns target point list[t]
    const int t = blockIdx.x * TPB + threadIdx.x;
    if (t >= (a.list ? a.nlist : a.npts)) return;
    const int n = a.list ? a.list[t] : t;
    if (a.band_j1 > 0 && a.list && (n < a.band_j0 * a.out_ni || n >= a.band_j1 * a.out_ni)) return;   // list points outside the output-row band
    const float px = a.x[n], py = a.y[n];
    const uint8_t bits = a.zb ? a.zb[n] : (uint8_t)0;
    const uint8_t zbits = a.use_zones ? (uint8_t)(bits & 0x1f) : (uint8_t)0;
    if (a.yy && (bits & ZB_YANG))
        point_body<DEG>(a.src2, a, a.zin + a.panel_stride, n, px, py, zbits);
    else
        point_body<DEG>(a.src, a, a.zin, n, px, py, zbits);
}

// ---------------------------------------------------------------------------------------------------
// the direct kernel: calls of a few fields (c_ezsint of one field, c_ezuvint of one pair -- C1, C3)
// ---------------------------------------------------------------------------------------------------
// A tile CTA's set-up (descriptor, plans of four points per thread, a copy pipeline with one stage to fill) is amortised over
// the levels of a launch; with one or two levels it is all there is, at 2 CTAs per SM.  Here: one thread per target point,
// no shared memory, few registers (high occupancy hides the gather latency; the source of a single-field call is L2-resident
// after its first touch), the same fp32 difference form as k_interp_fast<4 / 5>.  Zone points are left out: a constant
// extrapolation value outside the source (ez_corrval.c:56-93) is a plain store, the polar zones go to the per-point kernel
// on the gridset's zone list.  With `wrot` (c_ezuvint, one pair = two fields) the cached grid-to-grid wind rotation of
// ez_wind.cu is applied to the pair before it is stored, in k_wind_apply's own operation order.
template <int DEG>
__device__ __forceinline__ void direct_body(const GridDev &g, const InterpArgs &a, const float *__restrict__ z, int n, float px, float py) {
    const long is = a.in_stride, os = a.out_stride;
    const int nf = a.nfields;
    float *zo = a.zout + n;
    if (DEG == 1) {
        const PlanLinF p = to_f32(plan_linear(g, px, py));
        if (a.wrot) {
            const float u = eval_linear_f32(p, z), v = eval_linear_f32(p, z + is);
            const float4 r = __ldg(a.wrot + n);
            zo[0] = fmaf(r.y, v, r.x * u);
            zo[os] = fmaf(r.w, v, r.z * u);
            return;
        }
        for (int f = 0; f < nf; f++) zo[f * os] = eval_linear_f32(p, z + f * is);
    } else {
        // irregular sources, from the second call on a gridset: the plan comes from the gridset's cache (40 B per point,
        // k_plan_cache: the same fp32 weights to the bit) instead of ~30 fp32 <-> fp64 conversions and the fp64 Newton arithmetic
        PlanCubF p;
        int k = -1;
        if (a.pc_idx) k = __ldg(a.pc_idx + n);
        if (k >= 0) {
            const int c0 = k & 0xffff;
            const float4 X = __ldg(a.pc_wx + n), Y = __ldg(a.pc_wy + n);
            p.c[0] = c0; p.c[1] = c0 + 1; p.c[2] = c0 + 2; p.c[3] = c0 + 3;
            p.ni = g.ni; p.r0 = (k >> 16) * g.ni;
            p.wx[0] = X.x; p.wx[1] = X.y; p.wx[2] = X.z; p.wx[3] = X.w;
            p.wy[0] = Y.x; p.wy[1] = Y.y; p.wy[2] = Y.z; p.wy[3] = Y.w;
            p.sm1 = __ldg(a.pc_sm1 + n);
        } else {
            p = to_f32(plan_cubic(g, px, py));
        }
        if (a.wrot) {
            const float u = eval_cubic_f32(p, z), v = eval_cubic_f32(p, z + is);
            const float4 r = __ldg(a.wrot + n);
            zo[0] = fmaf(r.y, v, r.x * u);
            zo[os] = fmaf(r.w, v, r.z * u);
            return;
        }
        for (int f = 0; f < nf; f++) zo[f * os] = eval_cubic_f32(p, z + f * is);
    }
}

#ifndef EZ_DIRECT_MINB
#define EZ_DIRECT_MINB 4
#endif
#ifndef EZ_DIRECT_MINB_LIN
#define EZ_DIRECT_MINB_LIN 4
#endif
#ifndef EZ_DIRECT_PPT
#define EZ_DIRECT_PPT 2
#endif
#ifndef EZ_DIRECT_PPT_LIN
#define EZ_DIRECT_PPT_LIN 4
#endif
// points per thread: the kernel is bound by the latency of its dependent chain (coordinates -> gathers -> store), so a thread
// carries several independent points (TPB apart: every access stays coalesced) and issues all their coordinate loads first
template <int DEG> struct DirectCfg { static constexpr int PPT = (DEG == 1) ? EZ_DIRECT_PPT_LIN : EZ_DIRECT_PPT; };

template <int DEG>
__global__ void __launch_bounds__(TPB, DEG == 1 ? EZ_DIRECT_MINB_LIN : EZ_DIRECT_MINB) k_interp_direct(const __grid_constant__ InterpArgs a) {
    constexpr int PPT = DirectCfg<DEG>::PPT;
    // output-row band: full-grid indexing from the band's first point (zout points band_j0 rows before the band buffer)
    const int n0 = a.band_j1 > 0 ? a.band_j0 * a.out_ni : 0, n1 = a.band_j1 > 0 ? a.band_j1 * a.out_ni : a.npts;
    const int nb = n0 + blockIdx.x * (TPB * PPT) + threadIdx.x;
    float px[PPT], py[PPT];
    uint8_t bits[PPT];
    int pt[PPT];   // target point of slot k (through the walking order of a scattered target), -1 = none
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        const int m = nb + k * TPB;
        pt[k] = m < n1 ? (a.order ? __ldg(a.order + m) : m) : -1;
    }
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        const int n = pt[k];
        const bool in = n >= 0;
        px[k] = in ? a.x[n] : 1.f;
        py[k] = in ? a.y[n] : 1.f;
        bits[k] = (in && a.zb) ? a.zb[n] : (uint8_t)0;
    }
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        const int n = pt[k];
        if (n < 0) continue;
        const uint8_t zbits = a.use_zones ? (uint8_t)(bits[k] & 0x1f) : (uint8_t)0;
        if (zbits) {
            if (zbits == ZB_DEHORS && a.o.extrap >= EZ_MAXIMUM) {
                const long os = a.out_stride;
                for (int f = 0; f < a.nfields; f++) a.zout[n + f * os] = (a.o.vector || !a.fs) ? 0.0f : a.fs[f].fudge;
            }
            continue;   // every other zone point: k_interp_fast on the zone list
        }
        if (a.yy && (bits[k] & ZB_YANG))
            direct_body<DEG>(a.src2, a, a.zin + a.panel_stride, n, px[k], py[k]);
        else
            direct_body<DEG>(a.src, a, a.zin, n, px[k], py[k]);
    }
}

// the points with zone bits (optionally without those that are outside the source and nothing else), compacted
// (order: walk the points of a scattered target in their source-cell order, so that the list comes out roughly sorted as well)
__global__ void __launch_bounds__(TPB) k_zone_list(const uint8_t *__restrict__ zb, int npts, int skip_dehors_only, int *__restrict__ list,
                                                   unsigned int *__restrict__ cursor, const int *__restrict__ order) {
    const int t = blockIdx.x * TPB + threadIdx.x;
    const int n = (t < npts) ? (order ? order[t] : t) : 0;
    const uint8_t z = (t < npts) ? (uint8_t)(zb[n] & 0x1f) : (uint8_t)0;
    const bool take = z != 0 && !(skip_dehors_only && z == ZB_DEHORS);
    const unsigned int ballot = __ballot_sync(0xffffffffu, take);
    if (!ballot) return;
    const int lane = threadIdx.x & 31;
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned int)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) list[base + __popc(ballot & ((1u << lane) - 1u))] = n;
}

bool direct_args_ok(const InterpArgs &a) {
    if (a.o.arith != 2 || (a.o.degree != EZ_LINEAR && a.o.degree != EZ_CUBIC)) return false;
    if (a.src.xpn != XPN_NONE || a.src.yinv) return false;
    if (a.yy && (a.src2.xpn != XPN_NONE || a.src2.yinv)) return false;
    return true;
}

bool launch_fast(const InterpArgs &a0, cudaStream_t s) {
    if (a0.src.xpn != XPN_NONE || a0.src.yinv) return false;
    if (a0.yy && (a0.src2.xpn != XPN_NONE || a0.src2.yinv)) return false;
    const int nbx = ((a0.list ? a0.nlist : a0.npts) + TPB - 1) / TPB;
    if (nbx <= 0) return true;
    // A short list (a few zone points of a many-level call) is one CTA's latency chain: fewer levels per thread, down to one, until
    // there are some 64 K point-levels per chunk; long lists keep LIST_LCH levels per thread (the per-point plan is not free).
    InterpArgs a = a0;
    if (a.list && a.list_lch <= 0) {
        const long per = ((long)a.nlist * a.nfields + 65535) / 65536;
        a.list_lch = (int)(per < 1 ? 1 : (per > LIST_LCH ? LIST_LCH : per));
    }
    const int lch = a.list_lch > 0 ? a.list_lch : LIST_LCH;
    const dim3 nb(nbx, a.list ? (a.nfields + lch - 1) / lch : 1);
    switch (a.o.degree) {
        case EZ_NEAREST: k_interp_fast<0><<<nb, TPB, 0, s>>>(a); return true;
        case EZ_AVERAGE: case EZ_SPH_AVERAGE:   // zone points of an average only (interp_avg): the zone branch does not depend on DEG
            if (!a.list) return false;
            k_interp_fast<0><<<nb, TPB, 0, s>>>(a);
            return true;
        case EZ_LINEAR:
            if (a.o.arith == 2) k_interp_fast<5><<<nb, TPB, 0, s>>>(a);
            else k_interp_fast<1><<<nb, TPB, 0, s>>>(a);
            return true;
        case EZ_CUBIC:
            if (a.o.arith == 2) k_interp_fast<4><<<nb, TPB, 0, s>>>(a);
            else k_interp_fast<3><<<nb, TPB, 0, s>>>(a);
            return true;
        default: return false;
    }
}

}  // namespace
