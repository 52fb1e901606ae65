// librmn_b200/csrc/ez_interp_rowsep.cuh -- the row-separable tile kernel (linear and cubic, "arith = 2").
// Included by ez_interp.cu after ez_interp_tiled.cuh (shares its tile geometry and classification).
//
// Why: when the target rows run along the source grid's y axis -- lat-lon targets fed from lat-lon, Gaussian or
// unrotated/identity-rotated Z grids: BASELINE.json's C5 and the Yin half of C2 -- every point of a target row
// carries the SAME y (bit for bit), hence the same stencil rows and the same y-weights, while x (and the x-weights)
// stay free per point (identity-rotated Z/E grids give x values that differ by an ulp along a column, so the x side
// cannot be shared).  The bicubic sum then factors exactly as the reference evaluates it, only in the other order:
//     zo(i,j) = sum_c wx_ij[c] * V_j[c0_ij + c],      V_j[col] = sum_r wy_j[r] * z[r0_j + r][col]
// Phase 1 builds V_j for the tile's 32 target rows over the columns of the tile's source box: 4 FMAs per (row, col),
// shared by every target point of that row.  Phase 2 is 4 FMAs per output.  7 FMAs per output instead of the window
// kernel's 30, and every shared-memory access is conflict free: lanes read consecutive columns (phase 1) or columns
// that advance by less than one per lane (phase 2; equal addresses broadcast).
//   * levels are processed four at a time (one pipeline stage): V holds float4 (four levels), so phase 2 reads one
//     LDS.128 per stencil column, and both phases run on packed fp32x2 instructions (FFMA2) over level pairs with the
//     weight as the broadcast scalar operand;
//   * fp32 difference form about one reference value per (tile, level), zref = the corner of the staged box:
//     zo = zref (1 + s) + sum wx sum wy (z - zref), s = (sum wx)(sum wy) - 1 from the fp64 plan.  |z - zref| is bounded
//     by the variation of the field over the tile, so the rounding error scales with the field's RANGE, not its
//     magnitude.  A non-finite or absurd corner value (missing-data markers) falls back to zref = 0.
// The tile qualifies (k_tile_classify) when all its eligible quads have one y per target row and row-pair stencil
// origins at most one row apart; points in zones, on the other Yin-Yang panel or on the longitudinal seam go to the
// per-point list exactly as for the window kernel.
#pragma once

namespace {

#ifndef EZ_ROWSEP_MINB
#define EZ_ROWSEP_MINB 2
#endif
#ifndef EZ_RS_NSTAGE
#define EZ_RS_NSTAGE 3
#endif
constexpr int RS_NSTAGE = EZ_RS_NSTAGE;          // staged pipeline stages (4 levels each) per CTA
constexpr int RS_NCW = 16;                       // warps: warp w owns target row pair w in phase 1, rows w and w + 16 in phase 2
constexpr int RS_NCOMP = 32 * RS_NCW;
constexpr int RS_NTHR = RS_NCOMP;                // no dedicated producer: every warp issues its share of the cp.async copies
constexpr int RS_VP = 32;                        // V row pitch (float4): row-separable boxes are at most 32 columns wide
constexpr int RS_NCK = 10;                       // 16-byte chunks that cover one box row: (3 + 32 + 3) / 4 rounded up
// Staged box pitch.  P = rs_pitch(ni) is congruent to the source row length mod 4, so a 16-byte aligned chunk of a
// source row is 16-byte aligned in shared memory for EVERY row of the box (the per-level offset delta absorbs the
// alignment of the level's base): the copies are 16-byte cp.async although callers' row lengths are arbitrary.
__host__ __device__ constexpr int rs_pitch(int ni) { return 35 + (((ni - 35) % 4) + 4) % 4; }
__host__ __device__ constexpr int rs_cap(int P, int h) { return (P * h + 8 + 3) & ~3; }

__device__ __forceinline__ void compute_barrier() { __syncthreads(); }
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32x2(uint32_t addr, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory"); }

__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32x4(uint32_t addr, float2 lo, float2 hi) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(lo.x), "f"(lo.y), "f"(hi.x), "f"(hi.y) : "memory");
}
__device__ __forceinline__ float robust_ref(float z) { return (fabsf(z) < 1e15f) ? z : 0.f; }

static_assert(LPS == 4, "the row-separable kernel processes one 4-level pipeline stage per pass");

// Per-stage reference values (one per level): the corner of the staged box, guarded against missing-data markers.
struct StageRef { float2 z01, z23; };

// Phase 1 of one pipeline stage (four levels, or the 1..3 levels of the last stage when !FULL):
//   V_j[col] = sum_r wy_j[r] (z[r][col] - zref) for this warp's row pair, lanes = box columns (boxes are at most 32 wide).
//   sl[l] : shared address of box element (0, 0) of level l of the stage (pitch P floats)
//   p1    : byte offset of this lane's first window element from (0, 0) ; vst : shared address of V[2 wp][lane]
template <int DEG, int P, bool FULL>
__device__ __forceinline__ StageRef rowsep_phase1(const uint32_t (&sl)[4], int nl, uint32_t p1, uint32_t vst, int w, int lane,
                                                  const float (&wya)[2][Sten<DEG>::W]) {
    constexpr int W = Sten<DEG>::W;
    // levels beyond nl (last stage only) re-read the last valid level; their results are never stored
    const uint32_t sl0 = sl[0], sl1 = (FULL || nl > 1) ? sl[1] : sl0, sl2 = (FULL || nl > 2) ? sl[2] : sl1, sl3 = FULL ? sl[3] : sl2;
    StageRef zr;
    zr.z01 = make_float2(robust_ref(lds_f32(sl0)), robust_ref(lds_f32(sl1)));
    zr.z23 = make_float2(robust_ref(lds_f32(sl2)), robust_ref(lds_f32(sl3)));
    const float2 nz01 = make_float2(-zr.z01.x, -zr.z01.y), nz23 = make_float2(-zr.z23.x, -zr.z23.y);
    if (lane < w) {
        float2 a01, a23, b01, b23;   // a: first row of the pair, b: second ; (levels 0,1), (levels 2,3)
#pragma unroll
        for (int r = 0; r < W; r++) {
            const uint32_t o = p1 + (uint32_t)(r * P) * 4u;
            const float2 d01 = __fadd2_rn(make_float2(lds_f32(sl0 + o), lds_f32(sl1 + o)), nz01);
            const float2 d23 = __fadd2_rn(make_float2(lds_f32(sl2 + o), lds_f32(sl3 + o)), nz23);
            const float2 w0 = make_float2(wya[0][r], wya[0][r]), w1 = make_float2(wya[1][r], wya[1][r]);
            a01 = (r == 0) ? __fmul2_rn(w0, d01) : __ffma2_rn(w0, d01, a01);
            a23 = (r == 0) ? __fmul2_rn(w0, d23) : __ffma2_rn(w0, d23, a23);
            b01 = (r == 0) ? __fmul2_rn(w1, d01) : __ffma2_rn(w1, d01, b01);
            b23 = (r == 0) ? __fmul2_rn(w1, d23) : __ffma2_rn(w1, d23, b23);
        }
        sts_f32x4(vst, a01, a23);
        sts_f32x4(vst + (uint32_t)RS_VP * 16u, b01, b23);
    }
    return zr;
}

// Phase 2: zo = zref (1 + s) + sum_c wx[c] V_j[c0 + c], lanes = target columns ; vld[k] : shared address of V[row_k][c0_k]
template <int DEG, bool FULL>
__device__ __forceinline__ void rowsep_phase2(const StageRef &zr, int nl, const uint32_t (&vld)[2], const float (&wxa)[2][Sten<DEG>::S],
                                              const float (&sm1)[2], unsigned int okmask, char *__restrict__ zo, long osB, long rstepB) {
    constexpr int S = Sten<DEG>::S;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (okmask & (1u << k)) {
            float2 acc01, acc23;
#pragma unroll
            for (int c = 0; c < S; c++) {
                const float4 v = lds_f32x4(vld[k] + 16u * c);
                const float2 wc = make_float2(wxa[k][c], wxa[k][c]);
                acc01 = (c == 0) ? __fmul2_rn(wc, make_float2(v.x, v.y)) : __ffma2_rn(wc, make_float2(v.x, v.y), acc01);
                acc23 = (c == 0) ? __fmul2_rn(wc, make_float2(v.z, v.w)) : __ffma2_rn(wc, make_float2(v.z, v.w), acc23);
            }
            const float2 sk = make_float2(sm1[k], sm1[k]);
            const float2 r01 = __fadd2_rn(zr.z01, __ffma2_rn(zr.z01, sk, acc01));
            const float2 r23 = __fadd2_rn(zr.z23, __ffma2_rn(zr.z23, sk, acc23));
            char *p = zo + k * rstepB;
            *reinterpret_cast<float *>(p) = r01.x;
            if (FULL || nl > 1) *reinterpret_cast<float *>(p + osB) = r01.y;
            if (FULL || nl > 2) *reinterpret_cast<float *>(p + 2 * osB) = r23.x;
            if (FULL) *reinterpret_cast<float *>(p + 3 * osB) = r23.y;
        }
    }
}

template <int DEG, int P>
__device__ __forceinline__ void rowsep_compute(const GridDev &g, const InterpArgs &a, const TileDesc &td, int tile, int bx, int by,
                                               const float *__restrict__ zin, int src_rows, uint32_t smem32, uint32_t vbuf32,
                                               double *__restrict__ sumwy) {
    constexpr int S = Sten<DEG>::S, W = Sten<DEG>::W, VP = RS_VP;
    const int ni = a.out_ni, nj = a.out_nj;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;

    // ---- phase-1 plan: this warp owns target rows 2 wp, 2 wp + 1 of the tile (warp-uniform weights) ----
    float wya[2][W];        // [row of the pair][window row], zero padded
    uint32_t p1off;         // byte offset of the pair's first window element for this lane inside a staged box
    {
        int r0[2] = {INT_MAX, INT_MAX};
        float wy[2][4];
        bool has[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const unsigned int yb = a.tile_rowy[tile * TO + 2 * wp + r];
            has[r] = yb != ROWY_NONE;
            if (has[r]) {
                double sw = 1.0;
                if (DEG == 3) {   // the y side only; sum wy goes to shared memory for the point plans (s = sum wx sum wy - 1)
                    const PlanCub pc = plan_cubic<2>(g, 0.f, __uint_as_float(yb));
                    r0[r] = pc.r0 / pc.ni;
#pragma unroll
                    for (int k = 0; k < 4; k++) wy[r][k] = (float)pc.wy[k];
                    sw = pc.wy[0] + pc.wy[1] + pc.wy[2] + pc.wy[3];
                } else {
                    int c0;
                    float wx[4], s;
                    point_plan<DEG>(g, (float)(g.i1 + 1), __uint_as_float(yb), c0, r0[r], wx, wy[r], s);
                }
                if (lane == 0) sumwy[2 * wp + r] = sw;
            }
        }
        const int base = (has[0] || has[1]) ? min(r0[0], r0[1]) : td.r0;
        p1off = (uint32_t)((base - td.r0) * P + lane) * 4u;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const bool shifted = has[r] && r0[r] != base;   // origin one row below the pair's window origin
#pragma unroll
            for (int k = 0; k < W; k++) {
                const float lo = (k < S) ? wy[r][k] : 0.f, hi = (k > 0) ? wy[r][k - 1] : 0.f;
                wya[r][k] = has[r] ? (shifted ? hi : lo) : 0.f;
            }
        }
    }
    const uint32_t vst0 = vbuf32 + (uint32_t)((2 * wp) * VP + lane) * 16u;
    __syncthreads();   // sumwy of every tile row is in shared memory

    // ---- phase-2 plan: lane = target column of the tile, rows wp and wp + 16 ----
    float wxa[2][S], sm1[2];
    uint32_t vld0[2];
    unsigned int okmask = 0u;
    const int oi = bx * TO + lane;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int jj = wp + RS_NCW * k, oj = by * TO + jj;
        const int quad = (jj >> 1) * TT + (lane >> 1);
        const bool ok = oi < ni && oj < nj && ((a.tile_flags[tile * 8 + (quad >> 5)] >> (quad & 31)) & 1u);
        vld0[k] = vbuf32; sm1[k] = 0.f;
#pragma unroll
        for (int c = 0; c < S; c++) wxa[k][c] = 0.f;
        if (ok) {
            const int n = oj * ni + oi;
            int c0, r0;
            float wx[4], wy[4];
            if (a.pc_idx) {   // plan cache of the tile plan
                cached_plan<DEG, false>(g, a, n, c0, r0, wx, wy, sm1[k]);
            } else if (DEG == 3) {   // the x side only
                const PlanCub pc = plan_cubic<1>(g, a.x[n], 0.f);
                c0 = pc.c[0];
#pragma unroll
                for (int c = 0; c < 4; c++) wx[c] = (float)pc.wx[c];
                sm1[k] = (float)((pc.wx[0] + pc.wx[1] + pc.wx[2] + pc.wx[3]) * sumwy[jj] - 1.0);
            } else {
                point_plan<DEG>(g, a.x[n], a.y[n], c0, r0, wx, wy, sm1[k]);
            }
#pragma unroll
            for (int c = 0; c < S; c++) wxa[k][c] = wx[c];
            vld0[k] = vbuf32 + (uint32_t)(jj * VP + (c0 - td.c0)) * 16u;
            okmask |= 1u << k;
        }
    }

    const int nf = a.nfields;
    const long osB = a.out_stride * 4, rstepB = (long)RS_NCW * ni * 4;
    const int cap = a.tile_cap_rs;
    const uint32_t capB = (uint32_t)cap * 4u;
    constexpr uint32_t VBUF = (uint32_t)(TO * RS_VP) * 16u;
    const int nstages = (nf + LPS - 1) / LPS;

    // ---- the copy side: 16-byte cp.async (LDGSTS.128), one chunk per thread and level.
    // Level l of a stage sits in its slot with box element (r, j) at float offset delta_l + r P + j, where
    // delta_l = (float index in memory of source element (r0, c0) of that level) mod 4.  With P == ni (mod 4) every
    // 16-byte aligned chunk of a source row is 16-byte aligned in shared memory for every row of the box, which is what
    // a 16-byte cp.async needs although callers' row lengths are arbitrary.  delta depends on the level only through
    // l = f mod 4 (f = 4 st + l, and 4 st in_stride == 0 mod 4): fixed per slot position, so a thread's source and
    // destination offsets are computed once.  A chunk may over-read up to 3 floats on either side of the row segment;
    // it never leaves the 16-byte blocks that hold the segment's first and last element.
    // Thread t copies chunk t % RS_NCK of box row t / RS_NCK (boxes are at most 51 rows tall) of each level.
    const int nrow = min(td.h, src_rows - td.r0);
    const long is = a.in_stride;
    uint32_t sloff[LPS];   // byte offset of box element (0, 0) of level l inside a stage slot
    int cj = 0;            // packed: byte l = first column of the chunk relative to c0 (-3 .. w + 2), biased by 3
    unsigned int cvalid = 0u;
    const int crow = threadIdx.x / RS_NCK, ck4 = 4 * (threadIdx.x - crow * RS_NCK);
    {
        const float *lvl0 = zin + (long)td.r0 * g.ni + td.c0;        // source element (r0, c0) of level 0
#pragma unroll
        for (int l = 0; l < LPS; l++) {
            const int dl = (int)((reinterpret_cast<uintptr_t>(lvl0 + l * is) >> 2) & 3u);
            sloff[l] = (uint32_t)l * capB + (uint32_t)dl * 4u;
            const int bm = dl + crow * P;            // float offset of box element (crow, 0) in the level's buffer
            const int m0 = (bm & ~3) + ck4;          // this thread's chunk
            const int j = m0 - bm;
            cj |= (j + 3) << (8 * l);
            if (crow < nrow && j < td.w) cvalid |= 1u << l;
        }
    }
    const float *csrc = zin + (long)(td.r0 + crow) * g.ni + td.c0 - 3;   // (r0 + crow, c0 - 3) of the next stage to issue
    const uint32_t cbase = (uint32_t)(crow * P - 3) * 4u;                // byte offset of box element (crow, -3) from element (0, 0)
    auto issue = [&](int st) {   // called with st = 0, 1, 2, ... : csrc walks the levels
        if (cvalid && st < nstages) {
            const uint32_t sb = smem32 + (uint32_t)(st % RS_NSTAGE) * (LPS * capB);
            const float *p = csrc;
#pragma unroll
            for (int l = 0; l < LPS; l++) {
                if (((cvalid >> l) & 1u) && st * LPS + l < nf) {
                    const uint32_t jb = (uint32_t)(cj >> (8 * l)) & 0xffu;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + sloff[l] + cbase + 4u * jb), "l"(p + jb) : "memory");
                }
                p += is;
            }
        }
        csrc += LPS * is;
        __pipeline_commit();
    };
#pragma unroll 1
    for (int k = 0; k < RS_NSTAGE; k++) issue(k);
    __pipeline_wait_prior(RS_NSTAGE - 1);   // stage 0 has landed (this thread's share)
    __syncthreads();                        // ... and everybody else's

    char *zo = reinterpret_cast<char *>(a.zout + (long)(by * TO + wp) * ni + oi);
    uint32_t vsel = 0;
    uint32_t slot = 0;
    // One CTA barrier per stage: phase 1 (st) | wait for stage st+1 | barrier | refill the slot phase 1 just read | phase 2 (st).
    // V is double buffered, so phase 2 (st) overlaps other warps' phase 1 (st+1).
#pragma unroll 1
    for (int st = 0; st < nstages; st++) {
        const uint32_t sbase = smem32 + slot * (LPS * capB);
        const uint32_t sl[4] = {sbase + sloff[0], sbase + sloff[1], sbase + sloff[2], sbase + sloff[3]};
        const int nl = nf - st * LPS;
        const uint32_t vld[2] = {vld0[0] + vsel, vld0[1] + vsel};
        StageRef zr;
        if (nl >= LPS) zr = rowsep_phase1<DEG, P, true>(sl, LPS, p1off, vst0 + vsel, td.w, lane, wya);
        else zr = rowsep_phase1<DEG, P, false>(sl, nl, p1off, vst0 + vsel, td.w, lane, wya);
        __pipeline_wait_prior(RS_NSTAGE - 2);
        __syncthreads();
        issue(st + RS_NSTAGE);
        if (nl >= LPS) rowsep_phase2<DEG, true>(zr, LPS, vld, wxa, sm1, okmask, zo, osB, rstepB);
        else rowsep_phase2<DEG, false>(zr, nl, vld, wxa, sm1, okmask, zo, osB, rstepB);
        vsel ^= VBUF;
        zo += LPS * osB;
        if (++slot == RS_NSTAGE) slot = 0;
    }
}

template <int DEG, int P>
__global__ void __launch_bounds__(RS_NTHR, EZ_ROWSEP_MINB) k_interp_rowsep(const __grid_constant__ InterpArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int by = remap_tile_row(blockIdx.y, gridDim.y);
    const int tile = by * gridDim.x + blockIdx.x;
    if (a.band_j1 > 0 && (by * TO < a.band_j0 || by * TO >= a.band_j1)) return;   // output-row band: whole tile rows only
    const TileDesc td = a.tiles[tile];
    if (td.nfast == 0 || td.kind != 1) return;
    const int nstage_floats = RS_NSTAGE * LPS * a.tile_cap_rs;
    for (int k = threadIdx.x; k < nstage_floats; k += RS_NTHR) smem[k] = 0.f;   // overhang cells are never copied and must read as zero
    __syncthreads();
    const uint32_t smem32 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t vbuf32 = smem32 + (uint32_t)nstage_floats * 4u;
    double *sumwy = reinterpret_cast<double *>(smem + nstage_floats + 2 * TO * RS_VP * 4);   // 32 doubles after the two V buffers
    if (td.panel) rowsep_compute<DEG, P>(a.src2, a, td, tile, blockIdx.x, by, a.zin + a.panel_stride, a.src2.j2 - a.src2.j1 + 1, smem32, vbuf32, sumwy);
    else rowsep_compute<DEG, P>(a.src, a, td, tile, blockIdx.x, by, a.zin, a.src.j2 - a.src.j1 + 1, smem32, vbuf32, sumwy);
}

}  // namespace
