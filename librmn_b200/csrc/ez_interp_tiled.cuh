// librmn_b200/csrc/ez_interp_tiled.cuh -- the shared-memory tiled, register-blocked gather kernel
// (linear and cubic, "arith = 2" fp32 difference form).  Included by ez_interp.cu after ez_interp_fast.cuh.
//
// Why: the one-thread-per-point kernel issues 16 scalar gathers per point and level; ncu shows it at
// ~2.5 L1 wavefronts per gather and 61 % of the L1 wavefront rate while HBM sits at 8 % (profiles/r1a_*).
// The L1/LSU path, not HBM, is the wall.  This kernel removes most of that traffic:
//   * a CTA owns a 32 x 32 tile of TARGET points; the bounding box of all their stencils in the source
//     level (a few KB: the tile's footprint) is staged in shared memory by 16-byte cp.async, NSTAGE stages of
//     LPS levels deep, so global memory is read in whole rows (coalesced) exactly once per tile and level;
//   * a thread owns a 2 x 2 block of target points.  When the target grid is as fine as or finer than the
//     source (every configuration of BASELINE.json), the four 4x4 stencils fit one 5 x 5 window: 25 LDS feed
//     4 outputs (6.25 loads per output instead of 16).  Each output applies its own weights zero-padded to the
//     window (wx'[5], wy'[5]); no selects, no dynamic register indexing;
//   * the arithmetic is the fp32 difference form of ez_interp_fast.cuh taken about the window centre:
//     zo = zc (1 + s) + sum_r wy'_r sum_c wx'_c (z_rc - zc), weights from the fp64 plan, evaluated on packed
//     fp32x2 instructions whose two halves are two consecutive LEVELS of one point (the weight is the scalar
//     broadcast operand): FADD2 forms the differences, FFMA2 the sums;
//   * the staged row pitch is chosen per tile (k_tile_classify) among the values congruent to the source row
//     length mod 4 (what the 16-byte copies need) so as to minimise shared-memory bank conflicts of the tile's
//     own warps: on a rotated panel the 32 window origins of a warp are a sheared lattice, and the pitch decides
//     which of them collide (scripts/analysis/bank_sim.py: 2.0 wavefronts per LDS at the minimal pitch, 1.4 at
//     the best of 8).
// Points that do not fit (zone points, the longitudinal seam, Yin/Yang panel borders, windows larger than
// 5 x 5 or overhanging the field) are classified ONCE per (gridset, degree) by k_tile_classify into a compacted
// list and handled by the per-point kernel (k_interp_fast with an index list); the tile kernel never branches on them.
#pragma once
#include <cuda_pipeline_primitives.h>

namespace {

constexpr int TT = 16;             // threads per tile side
constexpr int TO = 2 * TT;         // target points per tile side
#ifndef EZ_TILE_CAP
#define EZ_TILE_CAP 2560
#endif
#ifndef EZ_NSTAGE
#define EZ_NSTAGE 3
#endif
#ifndef EZ_LPS
#define EZ_LPS 4
#endif
#ifndef EZ_TILED_MINB
#define EZ_TILED_MINB 2
#endif
#ifndef EZ_WIN_NPITCH
#define EZ_WIN_NPITCH 8
#endif
#ifndef EZ_WIN_CAPLIM
#define EZ_WIN_CAPLIM 2304
#endif
constexpr int TILE_CAP = EZ_TILE_CAP;     // floats per staged level a tile box may take at most
constexpr int LPS = EZ_LPS;               // source levels per pipeline stage (one CTA barrier per stage)
constexpr int NSTAGE = EZ_NSTAGE;         // stages in flight per CTA
constexpr int NTHR_WIN = TT * TT;
constexpr int WIN_NPITCH = EZ_WIN_NPITCH; // candidate pitches per tile: win_pitch(w, ni) + 4 j, j < WIN_NPITCH
constexpr int WIN_CAPLIM = EZ_WIN_CAPLIM; // a wider-than-minimal pitch is only taken while the staged level stays below this (2 CTAs/SM)
// smallest staging pitch >= w + 3 congruent to the source row length mod 4: a 16-byte aligned chunk of a source row is
// then 16-byte aligned in shared memory for every row of the box (see ez_interp_rowsep.cuh)
__host__ __device__ constexpr int win_pitch(int w, int ni) { return (w + 3) + ((((ni - (w + 3)) % 4) + 4) % 4); }
__host__ __device__ constexpr int win_cap(int pitch, int h) { return pitch * h + 8; }

// shared memory through 32-bit shared-window addresses (keeps address arithmetic to one IADD per access group)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ float lds_f32_imm(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}

template <int DEG> struct Sten {
    static constexpr int S = (DEG == 3) ? 4 : 2;   // stencil width
    static constexpr int W = S + 1;                // window width of a 2 x 2 block
};

// stencil origin (0-based column of the first stencil column, 0-based row of the first stencil row in the
// stored field) and the per-axis weights of one target point.  ok = the stencil columns are contiguous.
template <int DEG>
__device__ __forceinline__ bool point_plan(const GridDev &g, float px, float py, int &c0, int &r0, float wx[4], float wy[4], float &sm1) {
    if (DEG == 3) {
        const PlanCub p = plan_cubic(g, px, py);
        c0 = p.c[0];
        r0 = p.r0 / p.ni;
#pragma unroll
        for (int k = 0; k < 4; k++) { wx[k] = (float)p.wx[k]; wy[k] = (float)p.wy[k]; }
        sm1 = (float)((p.wx[0] + p.wx[1] + p.wx[2] + p.wx[3]) * (p.wy[0] + p.wy[1] + p.wy[2] + p.wy[3]) - 1.0);
        return p.c[1] == c0 + 1 && p.c[2] == c0 + 2 && p.c[3] == c0 + 3;
    } else {
        const PlanLin p = plan_linear(g, px, py);
        r0 = p.o00 / g.ni;
        c0 = p.o00 - r0 * g.ni;
        wx[0] = (float)(1.0 - p.dx); wx[1] = (float)p.dx; wx[2] = wx[3] = 0.f;
        wy[0] = (float)(1.0 - p.dy); wy[1] = (float)p.dy; wy[2] = wy[3] = 0.f;
        sm1 = 0.f;
        return p.o10 == p.o00 + 1;
    }
}

// the four target points of a thread: (ox, oy), (ox+1, oy), (ox, oy+1), (ox+1, oy+1)
struct Quad {
    int ox, oy;
    __device__ __forceinline__ bool valid(int p, int ni, int nj) const { return (ox + (p & 1)) < ni && (oy + (p >> 1)) < nj; }
    __device__ __forceinline__ int index(int p, int ni) const { return (oy + (p >> 1)) * ni + ox + (p & 1); }
};
// both polar caps (the rows that carry zone points and ragged work) are scheduled first: 0, last, 1, last-1 ...
__device__ __forceinline__ int remap_tile_row(int by, int nty) { return (by & 1) ? (nty - 1 - (by >> 1)) : (by >> 1); }
__device__ __forceinline__ Quad quad_of_thread(int bx, int by) {
    Quad q;
    q.ox = bx * TO + 2 * (threadIdx.x & (TT - 1));
    q.oy = by * TO + 2 * (threadIdx.x >> 4);
    return q;
}

// ---------------------------------------------------------------------------------------------------
// classification (once per gridset, degree and zone setting)
// ---------------------------------------------------------------------------------------------------
// Per-thread result of the classification: `elig` = every valid point of the quad is zone-free, on the tile's panel and
// has contiguous stencil columns; `fast` = elig and the four stencils fit one W x W window (what the 2x2 window kernel
// needs); `rowpair` = the two target rows of the quad have stencil row origins at most one apart (what the
// row-separable kernel needs).  (cmin..cmax, rmin..rmax) are the extents of the stencil ORIGINS.
struct QuadClass {
    bool elig, fast, rowpair;
    int cmin, cmax, rmin, rmax;
    unsigned int ybits[2];   // bit pattern of y of the quad's two target rows (valid when yok[r])
    bool yok[2];             // row r has at least one valid point and all of them carry the same y
};
template <int DEG>
__device__ __forceinline__ QuadClass classify_thread(const GridDev &g, const InterpArgs &a, const Quad &q, int tile_panel) {
    const int ni = a.out_ni, nj = a.out_nj;
    QuadClass k;
    k.cmin = INT_MAX; k.cmax = INT_MIN; k.rmin = INT_MAX; k.rmax = INT_MIN;
    k.yok[0] = k.yok[1] = false; k.ybits[0] = k.ybits[1] = 0u;
    int nvalid = 0, rrow[2] = {INT_MAX, INT_MAX};
    bool ok = true, ysame = true;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        if (!q.valid(p, ni, nj)) continue;
        nvalid++;
        const int n = q.index(p, ni);
        const uint8_t bits = a.zb ? a.zb[n] : (uint8_t)0;
        const int panel = (a.yy && (bits & ZB_YANG)) ? 1 : 0;
        if ((a.use_zones && (bits & 0x1f)) || panel != tile_panel) { ok = false; continue; }
        int c0, r0;
        float wx[4], wy[4], s;
        const float py = a.y[n];
        if (!point_plan<DEG>(g, a.x[n], py, c0, r0, wx, wy, s)) ok = false;
        k.cmin = min(k.cmin, c0); k.cmax = max(k.cmax, c0);
        k.rmin = min(k.rmin, r0); k.rmax = max(k.rmax, r0);
        const int row = p >> 1;
        const unsigned int yb = __float_as_uint(py);
        if (k.yok[row] && k.ybits[row] != yb) ysame = false;
        k.yok[row] = true; k.ybits[row] = yb;
        rrow[row] = r0;
    }
    k.elig = ok && nvalid > 0;
    // the window kernel reads the whole W x W window: it must lie inside the field (the zero-weight overhang column/row
    // would otherwise read the next row's first element or past the end of the last field)
    constexpr int W = Sten<DEG>::W;
    k.fast = k.elig && (k.cmax - k.cmin) <= 1 && (k.rmax - k.rmin) <= 1 && (k.cmin + W) <= g.ni && (k.rmin + W) <= (g.j2 - g.j1 + 1);
    k.rowpair = k.elig && ysame && (!(k.yok[0] && k.yok[1]) || abs(rrow[1] - rrow[0]) <= 1);
    return k;
}

constexpr unsigned int ROWY_NONE = 0x7fc0dead;   // "this target row of the tile has no point on the row-separable path"

template <int DEG>
__global__ void __launch_bounds__(TT * TT) k_tile_classify(const __grid_constant__ InterpArgs a, TileDesc *__restrict__ tiles,
                                                           uint32_t *__restrict__ flags, unsigned int *__restrict__ rowy,
                                                           unsigned int *__restrict__ counters) {
    __shared__ int s_box[8];          // [0..3] window-kernel box, [4..7] row-separable box
    __shared__ int s_cnt[2];
    __shared__ unsigned int s_rowy[TO];
    __shared__ int s_bad;             // some quad breaks the row-separable structure
    __shared__ int s_conf[WIN_NPITCH];   // window kernel: shared-memory wavefronts per LDS summed over the tile's warps, per candidate pitch
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const Quad q = quad_of_thread(blockIdx.x, blockIdx.y);
    const int ni = a.out_ni, nj = a.out_nj;
    if (threadIdx.x < 8) s_box[threadIdx.x] = (threadIdx.x & 2) ? INT_MIN : INT_MAX;
    if (threadIdx.x < TO) s_rowy[threadIdx.x] = ROWY_NONE;
    if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_bad = 0; }
    if (threadIdx.x < WIN_NPITCH) s_conf[threadIdx.x] = 0;
    // panel of the tile: majority of the threads' first points
    const bool v0 = q.valid(0, ni, nj);
    const bool yang0 = v0 && a.yy && a.zb && (a.zb[q.index(0, ni)] & ZB_YANG);
    const int n_valid = __syncthreads_count(v0);
    const int n_yang = __syncthreads_count(yang0);
    const int tile_panel = (2 * n_yang > n_valid) ? 1 : 0;
    const QuadClass k = tile_panel ? classify_thread<DEG>(a.src2, a, q, 1) : classify_thread<DEG>(a.src, a, q, 0);
    constexpr int S = Sten<DEG>::S, W = Sten<DEG>::W;
    if (k.fast) {
        atomicMin(&s_box[0], k.cmin); atomicMin(&s_box[1], k.rmin);
        atomicMax(&s_box[2], k.cmin + W - 1);
        atomicMax(&s_box[3], k.rmin + W - 1);
    }
    if (k.elig) {
        // row-separable kernel: exact stencil columns, one zero-weight row of overhang for the row-pair window
        atomicMin(&s_box[4], k.cmin); atomicMin(&s_box[5], k.rmin);
        atomicMax(&s_box[6], k.cmax + S - 1);
        atomicMax(&s_box[7], k.rmax + W - 1);
        if (!k.rowpair) s_bad = 1;
        const int row0 = 2 * (threadIdx.x >> 4);
#pragma unroll
        for (int r = 0; r < 2; r++)
            if (k.yok[r]) {
                const unsigned int old = atomicCAS(&s_rowy[row0 + r], ROWY_NONE, k.ybits[r]);
                if (old != ROWY_NONE && old != k.ybits[r]) s_bad = 1;
            }
    }
    __syncthreads();
    // row-separable tile: every eligible quad qualifies (same y along each target row) and the box fits
    const int sc0 = s_box[4], sr0 = s_box[5], sw = s_box[6] - s_box[4] + 1, sh = s_box[7] - s_box[5] + 1;
    const bool rowsep = a.rowsep_enable && !s_bad && sc0 != INT_MAX && sw <= 32 && sh <= 51 && 38L * sh + 8 <= TILE_CAP;
    const int c0 = rowsep ? sc0 : s_box[0], r0 = rowsep ? sr0 : s_box[1];
    const int c1 = rowsep ? s_box[6] : s_box[2], r1 = rowsep ? s_box[7] : s_box[3];
    const int w = c1 - c0 + 1, h = r1 - r0 + 1;
    const int pmin = win_pitch(w, tile_panel ? a.src2.ni : a.src.ni);
    const bool box_ok = (c0 != INT_MAX) && w <= 64 && (rowsep || ((long)win_cap(pmin, h) <= TILE_CAP && h * ((w + 6) >> 2) <= 2 * TT * TT));
    const bool fast = (rowsep ? k.elig : k.fast) && box_ok;
    const unsigned int ballot = __ballot_sync(0xffffffffu, fast);
    if ((threadIdx.x & 31) == 0) flags[tile * 8 + (threadIdx.x >> 5)] = ballot;
    if (threadIdx.x < TO) rowy[tile * TO + threadIdx.x] = rowsep ? s_rowy[threadIdx.x] : ROWY_NONE;
    // window tiles: bank conflicts of this warp's LDS for every candidate pitch (lanes = the threads that will load)
    if (!rowsep && box_ok) {
        const int lane = threadIdx.x & 31;
        const int bc = k.cmin - c0, br = k.rmin - r0;
#pragma unroll 1
        for (int j = 0; j < WIN_NPITCH; j++) {
            const int addr = br * (pmin + 4 * j) + bc;
            const unsigned int mb = __match_any_sync(0xffffffffu, fast ? (addr & 31) : (32 + lane));
            const unsigned int ma = __match_any_sync(0xffffffffu, fast ? addr : (-1 - lane));
            const unsigned int leaders = __ballot_sync(0xffffffffu, fast && lane == (__ffs(ma) - 1));
            const int m = __reduce_max_sync(0xffffffffu, fast ? __popc(mb & leaders) : 0);
            if (lane == 0) atomicAdd(&s_conf[j], m);
        }
    }
    int nslow = 0;
    if (!fast)
        for (int p = 0; p < 4; p++) nslow += q.valid(p, ni, nj) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) nslow += __shfl_xor_sync(0xffffffffu, nslow, o);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], nslow); atomicAdd(&s_cnt[1], __popc(ballot)); }
    __syncthreads();
    if (threadIdx.x == 0) {
        TileDesc td;
        td.c0 = box_ok ? c0 : 0; td.r0 = box_ok ? r0 : 0; td.w = box_ok ? w : 0; td.h = box_ok ? h : 0;
        td.panel = tile_panel; td.nfast = s_cnt[1]; td.kind = (rowsep && box_ok) ? 1 : 0;
        int best = 0;
        if (box_ok && !rowsep)
            for (int j = 1; j < WIN_NPITCH; j++)
                if (win_cap(pmin + 4 * j, h) <= WIN_CAPLIM && s_conf[j] < s_conf[best]) best = j;
        td.pitch = box_ok ? pmin + 4 * best : 0;
        tiles[tile] = td;
        atomicAdd(counters + 0, (unsigned int)s_cnt[0]);
        if (box_ok && s_cnt[1] > 0 && td.kind == 0) {
            atomicMax(counters + 2, (unsigned int)win_cap(td.pitch, h));
            atomicAdd(counters + 6, (unsigned int)s_conf[best]);     // diagnostics: wavefronts per LDS, summed over warps, chosen
            atomicAdd(counters + 7, (unsigned int)s_conf[0]);        // ... and at the minimal pitch
        }
        if (s_cnt[1] > 0) atomicAdd(counters + 1, 1u);
        if (td.kind == 1 && s_cnt[1] > 0) {
            atomicMax(counters + 3, (unsigned int)w);
            atomicAdd(counters + 4, 1u);
            atomicMax(counters + 5, (unsigned int)h);
        }
    }
}

// second pass: the points the tile kernel leaves out, compacted (warp-aggregated append; order is irrelevant)
// mode 0: every point the tile kernels leave out ; 1: those, except the points that are outside the source and in no other zone
// (with a constant extrapolation value they are plain stores: k_store_const on the list mode 2 builds) ; 2: only those points
__global__ void __launch_bounds__(TT * TT) k_tile_list(const __grid_constant__ InterpArgs a, const uint32_t *__restrict__ flags,
                                                       int *__restrict__ list, unsigned int *__restrict__ cursor, int mode) {
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const Quad q = quad_of_thread(blockIdx.x, blockIdx.y);
    const int ni = a.out_ni, nj = a.out_nj;
    const bool fast = (flags[tile * 8 + (threadIdx.x >> 5)] >> (threadIdx.x & 31)) & 1u;
    int k = 0;
    bool take[4] = {false, false, false, false};
    if (!fast)
        for (int p = 0; p < 4; p++) {
            if (!q.valid(p, ni, nj)) continue;
            const bool cst = mode != 0 && a.zb && (a.zb[q.index(p, ni)] & 0x1f) == ZB_DEHORS;
            take[p] = (mode == 0) || (mode == 1 && !cst) || (mode == 2 && cst);
            k += take[p] ? 1 : 0;
        }
    int incl = k;
    const int lane = threadIdx.x & 31;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned int base = 0;
    if (lane == 0 && total > 0) base = atomicAdd(cursor, (unsigned int)total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (k > 0) {
        unsigned int at = base + (unsigned int)(incl - k);
        for (int p = 0; p < 4; p++)
            if (take[p]) list[at++] = q.index(p, ni);
    }
}

// ez_corrval.c:56-93 with extrap_degree = maximum / minimum / value: the points outside the source get one value per field
// (0.0 for vector components, :64-67) -- plain stores on a compacted list
__global__ void __launch_bounds__(256) k_store_const(const __grid_constant__ InterpArgs a, const int *__restrict__ clist, int nclist) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= nclist) return;
    const int n = clist[t];
    if (a.band_j1 > 0 && (n < a.band_j0 * a.out_ni || n >= a.band_j1 * a.out_ni)) return;
    const long os = a.out_stride;
    for (int f = 0; f < a.nfields; f++) a.zout[n + f * os] = (a.o.vector || !a.fs) ? 0.0f : a.fs[f].fudge;
}

// third pass: the plan of every target point, stored for the tile kernels (40 B per point)
// mark_split: a point whose stencil columns are not contiguous (the longitudinal seam) gets idx = -1 -- the direct kernel, which
// has no classification to keep such points away from the cache, then evaluates its plan on the fly
template <int DEG>
__global__ void __launch_bounds__(256) k_plan_cache(const __grid_constant__ InterpArgs a, int *__restrict__ idx, float4 *__restrict__ wxo,
                                                    float4 *__restrict__ wyo, float *__restrict__ sm1o, int mark_split) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= a.npts) return;
    const uint8_t bits = a.zb ? a.zb[n] : (uint8_t)0;
    int c0 = 0, r0 = 0;
    float wx[4] = {0.f, 0.f, 0.f, 0.f}, wy[4] = {0.f, 0.f, 0.f, 0.f}, s = 0.f;
    bool ok;
    if (a.yy && (bits & ZB_YANG)) ok = point_plan<DEG>(a.src2, a.x[n], a.y[n], c0, r0, wx, wy, s);
    else ok = point_plan<DEG>(a.src, a.x[n], a.y[n], c0, r0, wx, wy, s);
    if (mark_split && !ok) { idx[n] = -1; return; }
    idx[n] = (r0 << 16) | (c0 & 0xffff);
    wxo[n] = make_float4(wx[0], wx[1], wx[2], wx[3]);
    wyo[n] = make_float4(wy[0], wy[1], wy[2], wy[3]);
    sm1o[n] = s;
}
// a point's plan: from the cache when there is one, recomputed otherwise (bit-identical values)
template <int DEG, bool NEED_Y = true>
__device__ __forceinline__ void cached_plan(const GridDev &g, const InterpArgs &a, int n, int &c0, int &r0, float wx[4], float wy[4], float &sm1) {
    if (a.pc_idx) {
        const int k = __ldg(a.pc_idx + n);
        c0 = k & 0xffff; r0 = k >> 16;
        const float4 X = __ldg(a.pc_wx + n);
        wx[0] = X.x; wx[1] = X.y; wx[2] = X.z; wx[3] = X.w;
        if (NEED_Y) {   // (the row-separable kernel takes the y side from its row plans)
            const float4 Y = __ldg(a.pc_wy + n);
            wy[0] = Y.x; wy[1] = Y.y; wy[2] = Y.z; wy[3] = Y.w;
        }
        sm1 = __ldg(a.pc_sm1 + n);
    } else {
        point_plan<DEG>(g, a.x[n], a.y[n], c0, r0, wx, wy, sm1);
    }
}

// ---------------------------------------------------------------------------------------------------
// the window tile kernel
// ---------------------------------------------------------------------------------------------------
// Arithmetic of one level PAIR (levels A, B of a pipeline stage) for one thread (2 x 2 target points, W x W window):
//   d_rc = (zA_rc - zcA, zB_rc - zcB)                      FADD2 ; zc = window centre of the level (fp32 difference form)
//   out_p = zc (1 + s_p) + sum_r wy_p[r] * ( sum_c wx_p[c] * d_rc )     FFMA2, the weight is the scalar broadcast operand
// The window is streamed row by row, so only one row of data (W register pairs) is live.
template <int DEG>
struct WinThread {
    static constexpr int S = Sten<DEG>::S, W = Sten<DEG>::W;
    float wx[4][W], wy[4][W], sm1[4];   // per point, zero padded to the window
    uint32_t roff[W];                   // byte offset of window row r from window element (0, 0)
};

template <int DEG, bool HASB>
__device__ __forceinline__ void window_pair(const WinThread<DEG> &t, uint32_t aA, uint32_t aB, float *__restrict__ w0, int ni, long os,
                                            bool full, const bool (&valid)[4]) {
    constexpr int W = Sten<DEG>::W;
    constexpr int CTR = (W / 2) * 4;
    const float2 zc = make_float2(lds_f32_imm<CTR>(aA + t.roff[W / 2]), lds_f32_imm<CTR>(aB + t.roff[W / 2]));
    const float2 nzc = make_float2(-zc.x, -zc.y);
    float2 acc[4];
#pragma unroll
    for (int r = 0; r < W; r++) {
        const uint32_t ar = aA + t.roff[r], br = aB + t.roff[r];
        float2 d[W];
        d[0] = __fadd2_rn(make_float2(lds_f32_imm<0>(ar), lds_f32_imm<0>(br)), nzc);
        d[1] = __fadd2_rn(make_float2(lds_f32_imm<4>(ar), lds_f32_imm<4>(br)), nzc);
        d[2] = __fadd2_rn(make_float2(lds_f32_imm<8>(ar), lds_f32_imm<8>(br)), nzc);
        if (W > 3) {
            d[3] = __fadd2_rn(make_float2(lds_f32_imm<12>(ar), lds_f32_imm<12>(br)), nzc);
            d[W - 1] = __fadd2_rn(make_float2(lds_f32_imm<4 * (W - 1)>(ar), lds_f32_imm<4 * (W - 1)>(br)), nzc);
        }
#pragma unroll
        for (int p = 0; p < 4; p++) {
            float2 rs = __fmul2_rn(make_float2(t.wx[p][0], t.wx[p][0]), d[0]);
#pragma unroll
            for (int c = 1; c < W; c++) rs = __ffma2_rn(make_float2(t.wx[p][c], t.wx[p][c]), d[c], rs);
            const float2 wyr = make_float2(t.wy[p][r], t.wy[p][r]);
            acc[p] = (r == 0) ? __fmul2_rn(wyr, rs) : __ffma2_rn(wyr, rs, acc[p]);
        }
    }
    float2 res[4];
#pragma unroll
    for (int p = 0; p < 4; p++) res[p] = __fadd2_rn(zc, __ffma2_rn(zc, make_float2(t.sm1[p], t.sm1[p]), acc[p]));
    float *w1 = w0 + ni;
    if (full) {
        *reinterpret_cast<float2 *>(w0) = make_float2(res[0].x, res[1].x);
        *reinterpret_cast<float2 *>(w1) = make_float2(res[2].x, res[3].x);
        if (HASB) {
            *reinterpret_cast<float2 *>(w0 + os) = make_float2(res[0].y, res[1].y);
            *reinterpret_cast<float2 *>(w1 + os) = make_float2(res[2].y, res[3].y);
        }
    } else {
        if (valid[0]) w0[0] = res[0].x;
        if (valid[1]) w0[1] = res[1].x;
        if (valid[2]) w1[0] = res[2].x;
        if (valid[3]) w1[1] = res[3].x;
        if (HASB) {
            if (valid[0]) w0[os] = res[0].y;
            if (valid[1]) w0[os + 1] = res[1].y;
            if (valid[2]) w1[os] = res[2].y;
            if (valid[3]) w1[os + 1] = res[3].y;
        }
    }
}

template <int DEG>
__device__ __forceinline__ void window_tile(const GridDev &g, const InterpArgs &a, const TileDesc &td, const Quad &q, bool fast,
                                            const float *__restrict__ zin, int src_rows, const uint32_t smem32) {
    constexpr int S = Sten<DEG>::S, W = Sten<DEG>::W;
    const int ni = a.out_ni, nj = a.out_nj;
    const int pitch = td.pitch;
    const int nf = a.nfields;
    const long os = a.out_stride, is = a.in_stride;
    const uint32_t capB = (uint32_t)a.tile_cap * 4u;
    const int nstages = (nf + LPS - 1) / LPS;

    // ---- copy side (see ez_interp_rowsep.cuh): box element (r, j) of level l of a stage sits at float offset
    // delta_l + r pitch + j of the level's buffer, delta_l = (float index in memory of source element (r0, c0) of that level) mod 4.
    // With pitch == ni (mod 4) every 16-byte aligned chunk of a source row is 16-byte aligned in shared memory.  delta depends on
    // the level only through l = f mod 4 (LPS == 4), so a thread's offsets are computed once.  Thread t copies chunks t and
    // t + 256 of each level (chunk = box row t / nck, 16-byte chunk t % nck of that row): chunk k of row `crow` starts at box
    // column 4 k - phi, phi = (delta_l + crow pitch) mod 4.
    const int nrow = min(td.h, src_rows - td.r0);
    const int nck = (td.w + 6) >> 2;
    unsigned int dpk = 0u;          // 2 bits per level: delta_l
    {
        const float *lvl0 = zin + (long)td.r0 * g.ni + td.c0;
#pragma unroll
        for (int l = 0; l < LPS; l++) dpk |= (unsigned int)((reinterpret_cast<uintptr_t>(lvl0 + l * is) >> 2) & 3u) << (2 * l);
    }
    // slot 0 (every thread): byte offsets of the chunk per level, from the stage's first level in global memory (signed: the
    // fused (u, v) launch may have a negative field stride; range checked by tiled_eligible) / from the stage's slot in shared
    // memory ; OG_NONE = nothing to copy.  Slot 1 (boxes of more than 256 chunks: a few threads of warp 0) is evaluated on the fly.
    constexpr int OG_NONE = INT_MIN;
    int og[LPS];
    uint32_t osm[LPS];
    {
        const int crow = threadIdx.x / nck, ck4 = 4 * (threadIdx.x - crow * nck);
#pragma unroll
        for (int l = 0; l < LPS; l++) {
            const int dl = (int)((dpk >> (2 * l)) & 3u);
            const int phi = (dl + crow * pitch) & 3;
            const bool ok = crow < nrow && ck4 - phi < td.w;
            og[l] = ok ? (int)(((long)l * is + (long)(td.r0 + crow) * g.ni + td.c0 + ck4 - phi) * 4) : OG_NONE;
            osm[l] = (uint32_t)l * capB + (uint32_t)(dl + crow * pitch + ck4 - phi) * 4u;
        }
    }
    const bool slot1 = (int)threadIdx.x + NTHR_WIN < nrow * nck;
    auto issue = [&](int st, uint32_t slot) {
        if (st < nstages) {
            const uint32_t sb = smem32 + slot * (LPS * capB);
            const char *pst = reinterpret_cast<const char *>(zin + (long)(st * LPS) * is);
            const int nl = nf - st * LPS;
#pragma unroll
            for (int l = 0; l < LPS; l++)
                if (og[l] != OG_NONE && l < nl)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + osm[l]), "l"(pst + og[l]) : "memory");
            if (slot1) {
                const int tt = threadIdx.x + NTHR_WIN;
                const int crow = tt / nck, ck4 = 4 * (tt - crow * nck);
#pragma unroll 1
                for (int l = 0; l < min(nl, LPS); l++) {
                    const int dl = (int)((dpk >> (2 * l)) & 3u);
                    const int phi = (dl + crow * pitch) & 3;
                    if (ck4 - phi < td.w) {
                        const float *src = zin + (long)(st * LPS + l) * is + (long)(td.r0 + crow) * g.ni + td.c0 + ck4 - phi;
                        const uint32_t dst = sb + (uint32_t)l * capB + (uint32_t)(dl + crow * pitch + ck4 - phi) * 4u;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    }
                }
            }
        }
        __pipeline_commit();
    };
#pragma unroll 1
    for (int k = 0; k < NSTAGE; k++) issue(k, (uint32_t)k);

    // ---- compute side: the plans are loaded / evaluated while the first stages are in flight ----
    WinThread<DEG> t;
    bool valid[4] = {false, false, false, false};
    uint32_t boff = 0;   // byte offset of window element (0, 0) from box element (0, 0)
#pragma unroll
    for (int r = 0; r < W; r++) t.roff[r] = (uint32_t)(r * pitch) * 4u;
    if (fast) {
        int c0[4], r0[4], wc0 = INT_MAX, wr0 = INT_MAX;
        float tx[4][4], ty[4][4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            valid[p] = q.valid(p, ni, nj);
            c0[p] = INT_MAX; r0[p] = INT_MAX; t.sm1[p] = 0.f;
            if (valid[p]) {
                const int n = q.index(p, ni);
                cached_plan<DEG>(g, a, n, c0[p], r0[p], tx[p], ty[p], t.sm1[p]);
                wc0 = min(wc0, c0[p]); wr0 = min(wr0, r0[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const bool sx = valid[p] && (c0[p] != wc0), sy = valid[p] && (r0[p] != wr0);   // offset 0 or 1 inside the window
#pragma unroll
            for (int k = 0; k < W; k++) {
                const float lo = (k < S) ? tx[p][k] : 0.f, hi = (k > 0) ? tx[p][k - 1] : 0.f;
                t.wx[p][k] = valid[p] ? (sx ? hi : lo) : 0.f;
                const float lo2 = (k < S) ? ty[p][k] : 0.f, hi2 = (k > 0) ? ty[p][k - 1] : 0.f;
                t.wy[p][k] = valid[p] ? (sy ? hi2 : lo2) : 0.f;
            }
        }
        boff = (uint32_t)((wr0 - td.r0) * pitch + (wc0 - td.c0)) * 4u;
    }
    // float2 stores need all four points and 8-byte alignment (even target row length; q.ox is even)
    const bool full = fast && ((ni & 1) == 0) && valid[3] && ((reinterpret_cast<uintptr_t>(a.zout) & 7) == 0) && ((os & 1) == 0);
    float *w0 = a.zout + (long)q.oy * ni + q.ox;
    __pipeline_wait_prior(NSTAGE - 1);
    __syncthreads();

    uint32_t slot = 0;
#pragma unroll 1
    for (int st = 0; st < nstages; st++) {
        const uint32_t sb = smem32 + slot * (LPS * capB) + boff;
        const int nl = nf - st * LPS;
        if (fast) {
            const uint32_t a0 = sb + ((dpk >> 0) & 3u) * 4u, a1 = sb + capB + ((dpk >> 2) & 3u) * 4u;
            const uint32_t a2 = sb + 2u * capB + ((dpk >> 4) & 3u) * 4u, a3 = sb + 3u * capB + ((dpk >> 6) & 3u) * 4u;
            if (nl >= LPS) {
                window_pair<DEG, true>(t, a0, a1, w0, ni, os, full, valid);
                window_pair<DEG, true>(t, a2, a3, w0 + 2 * os, ni, os, full, valid);
            } else {   // last stage: 1 .. 3 levels (an absent second level of a pair re-reads the first and is not stored)
                if (nl >= 2) window_pair<DEG, true>(t, a0, a1, w0, ni, os, full, valid);
                else window_pair<DEG, false>(t, a0, a0, w0, ni, os, full, valid);
                if (nl == 3) window_pair<DEG, false>(t, a2, a2, w0 + 2 * os, ni, os, full, valid);
            }
            w0 += LPS * os;
        }
        __pipeline_wait_prior(NSTAGE - 2);   // stage st + 1 has landed (this thread's share)
        __syncthreads();                     // ... everybody's, and everybody is done reading stage st
        issue(st + NSTAGE, slot);            // refill the slot just read
        if (++slot == NSTAGE) slot = 0;
    }
}

// One CTA per window tile of the compacted list a.win_tiles (k_tile_classify + host compaction: the tiles of the row-separable
// kernel and the empty ones never launch).  A persistent variant whose copy pipeline ran across tile boundaries (descriptor and
// plans of the CTA's next tile prefetched) was measured and lost 5-10 %: the hardware's dynamic CTA scheduling de-phases the
// two resident CTAs of an SM better than a static tile walk does (profiles/r2_window_kernel_log.md).
template <int DEG>
__device__ __forceinline__ void tiled_cta(const InterpArgs &a) {
    extern __shared__ __align__(16) float smem[];
    const int tile = a.win_tiles[blockIdx.x];
    const int ntx = (a.out_ni + TO - 1) / TO;
    if (a.band_j1 > 0) {   // output-row band: whole tile rows only (the host falls back to the per-point kernel otherwise)
        const int row0 = (tile / ntx) * TO;
        if (row0 < a.band_j0 || row0 >= a.band_j1) return;
    }
    const TileDesc td = a.tiles[tile];
    // no zero fill: every cell a fast thread reads lies inside the field (k_tile_classify) and is written by the copies
    const uint32_t smem32 = (uint32_t)__cvta_generic_to_shared(smem);
    const bool fast = (a.tile_flags[tile * 8 + (threadIdx.x >> 5)] >> (threadIdx.x & 31)) & 1u;
    const Quad q = quad_of_thread(tile % ntx, tile / ntx);
    if (td.panel) window_tile<DEG>(a.src2, a, td, q, fast, a.zin + a.panel_stride, a.src2.j2 - a.src2.j1 + 1, smem32);
    else window_tile<DEG>(a.src, a, td, q, fast, a.zin, a.src.j2 - a.src.j1 + 1, smem32);
}
template <int DEG>
__global__ void __launch_bounds__(NTHR_WIN, EZ_TILED_MINB) k_interp_tiled(const __grid_constant__ InterpArgs a) { tiled_cta<DEG>(a); }
// Launches of one pipeline stage (at most LPS fields): the per-tile latency chain (descriptor, plans, copies, barrier) is all there
// is, so CTAs in flight matter more than registers: 3 CTAs per SM at 80 registers (spills in the plan set-up), one stage of shared memory.
#ifndef EZ_TILED_FEW_MINB
#define EZ_TILED_FEW_MINB 3
#endif
template <int DEG>
__global__ void __launch_bounds__(NTHR_WIN, EZ_TILED_FEW_MINB) k_interp_tiled_few(const __grid_constant__ InterpArgs a) { tiled_cta<DEG>(a); }

}  // namespace
