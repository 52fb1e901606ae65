// librmn_b200/csrc/ez_interp_tiled.cuh -- the shared-memory tiled, register-blocked gather kernel
// (linear and cubic, "arith = 2" fp32 difference form).  Included by ez_interp.cu after ez_interp_fast.cuh.
//
// Why: the one-thread-per-point kernel issues 16 scalar gathers per point and level; ncu shows it at
// ~2.5 L1 wavefronts per gather and 61 % of the L1 wavefront rate while HBM sits at 8 % (profiles/r1a_*).
// The L1/LSU path, not HBM, is the wall.  This kernel removes most of that traffic:
//   * a CTA owns a 32 x 32 tile of TARGET points; the bounding box of all their stencils in the source
//     level (a few KB: the tile's footprint) is staged in shared memory by cp.async, double-buffered
//     across levels, so global memory is read in whole rows (coalesced) exactly once per tile and level;
//   * a thread owns a 2 x 2 block of target points.  When the target grid is as fine as or finer than the
//     source (every configuration of BASELINE.json), the four 4x4 stencils fit one 5 x 5 window: 25
//     conflict-free LDS feed 4 outputs (6.25 loads per output instead of 16).  Each output applies its
//     own weights zero-padded to the window (wx'[5], wy'[5]); no selects, no dynamic register indexing;
//   * the arithmetic is the fp32 difference form of ez_interp_fast.cuh taken about the window centre:
//     zo = zc (1 + s) + sum_r wy'_r sum_c wx'_c (z_rc - zc), weights from the fp64 plan.
// Points that do not fit (zone points, the longitudinal seam, Yin/Yang panel borders, windows larger than
// 5 x 5) are classified ONCE per (gridset, degree) by k_tile_classify into a compacted list and handled by
// the per-point kernel (k_interp_fast with an index list); the tile kernel never branches on them.
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
#ifndef EZ_TILED_UNROLL
#define EZ_TILED_UNROLL 1
#endif
constexpr int TILED_UNROLL = EZ_TILED_UNROLL;
#ifndef EZ_TILED_NOPROD
#define EZ_TILED_NOPROD 1
#endif
// EZ_TILED_NOPROD: the window kernel stages with 16-byte cp.async issued by all its threads through a staging pitch
// congruent to the source row length mod 4 (see ez_interp_rowsep.cuh), no producer warp, one barrier per stage.
constexpr bool TILED_NOPROD = EZ_TILED_NOPROD != 0;
__host__ __device__ constexpr int win_pitch(int w, int ni) { return (w + 3) + ((((ni - (w + 3)) % 4) + 4) % 4); }
__host__ __device__ constexpr int win_cap(int w, int h) { return TILED_NOPROD ? (w + 6) * h + 8 : (w | 1) * h; }   // levels of a stage evaluated per loop iteration in the window kernel
constexpr int TILE_CAP = EZ_TILE_CAP;     // floats per shared-memory buffer
constexpr int LPS = EZ_LPS;               // source levels per pipeline stage (one CTA barrier per stage)
constexpr int NSTAGE = EZ_NSTAGE;          // buffers per CTA: the producer warp runs NSTAGE-1 source levels ahead of the compute warps

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
    k.fast = k.elig && (k.cmax - k.cmin) <= 1 && (k.rmax - k.rmin) <= 1;
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
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const Quad q = quad_of_thread(blockIdx.x, blockIdx.y);
    const int ni = a.out_ni, nj = a.out_nj;
    if (threadIdx.x < 8) s_box[threadIdx.x] = (threadIdx.x & 2) ? INT_MIN : INT_MAX;
    if (threadIdx.x < TO) s_rowy[threadIdx.x] = ROWY_NONE;
    if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_bad = 0; }
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
    // the box holds every fast thread's full window; a window may overhang the field by one column/row on the
    // high side (that column/row carries zero weights) -- stage_tile replicates the field edge there
    const int w = c1 - c0 + 1, h = r1 - r0 + 1;
    const bool box_ok = (c0 != INT_MAX) && w <= 64 && (long)win_cap(w, h) <= TILE_CAP && (!TILED_NOPROD || rowsep || h * ((w + 6) >> 2) <= 2 * TT * TT);
    const bool fast = (rowsep ? k.elig : k.fast) && box_ok;
    const unsigned int ballot = __ballot_sync(0xffffffffu, fast);
    if ((threadIdx.x & 31) == 0) flags[tile * 8 + (threadIdx.x >> 5)] = ballot;
    if (threadIdx.x < TO) rowy[tile * TO + threadIdx.x] = rowsep ? s_rowy[threadIdx.x] : ROWY_NONE;
    int nslow = 0;
    if (!fast)
        for (int p = 0; p < 4; p++) nslow += q.valid(p, ni, nj) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) nslow += __shfl_xor_sync(0xffffffffu, nslow, o);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], nslow); atomicAdd(&s_cnt[1], __popc(ballot)); }
    __syncthreads();
    if (threadIdx.x == 0) {
        TileDesc td;
        td.c0 = box_ok ? c0 : 0; td.r0 = box_ok ? r0 : 0; td.w = box_ok ? w : 0; td.h = box_ok ? h : 0;
        td.panel = tile_panel; td.nfast = s_cnt[1]; td.nslow = s_cnt[0]; td.kind = (rowsep && box_ok) ? 1 : 0;
        tiles[tile] = td;
        atomicAdd(counters + 0, (unsigned int)s_cnt[0]);
        if (box_ok && s_cnt[1] > 0 && td.kind == 0) atomicMax(counters + 2, (unsigned int)win_cap(w, h));
        if (s_cnt[1] > 0) atomicAdd(counters + 1, 1u);
        if (td.kind == 1 && s_cnt[1] > 0) {
            atomicMax(counters + 3, (unsigned int)w);
            atomicAdd(counters + 4, 1u);
            atomicMax(counters + 5, (unsigned int)h);
        }
    }
}

// second pass: the points the tile kernel leaves out, compacted (warp-aggregated append; order is irrelevant)
__global__ void __launch_bounds__(TT * TT) k_tile_list(const __grid_constant__ InterpArgs a, const uint32_t *__restrict__ flags,
                                                       int *__restrict__ list, unsigned int *__restrict__ cursor) {
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const Quad q = quad_of_thread(blockIdx.x, blockIdx.y);
    const int ni = a.out_ni, nj = a.out_nj;
    const bool fast = (flags[tile * 8 + (threadIdx.x >> 5)] >> (threadIdx.x & 31)) & 1u;
    int k = 0;
    if (!fast)
        for (int p = 0; p < 4; p++) k += q.valid(p, ni, nj) ? 1 : 0;
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
            if (q.valid(p, ni, nj)) list[at++] = q.index(p, ni);
    }
}

// third pass: the plan of every target point, stored for the tile kernels (40 B per point)
template <int DEG>
__global__ void __launch_bounds__(256) k_plan_cache(const __grid_constant__ InterpArgs a, int *__restrict__ idx, float4 *__restrict__ wxo,
                                                    float4 *__restrict__ wyo, float *__restrict__ sm1o) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= a.npts) return;
    const uint8_t bits = a.zb ? a.zb[n] : (uint8_t)0;
    int c0 = 0, r0 = 0;
    float wx[4] = {0.f, 0.f, 0.f, 0.f}, wy[4] = {0.f, 0.f, 0.f, 0.f}, s = 0.f;
    if (a.yy && (bits & ZB_YANG)) point_plan<DEG>(a.src2, a.x[n], a.y[n], c0, r0, wx, wy, s);
    else point_plan<DEG>(a.src, a.x[n], a.y[n], c0, r0, wx, wy, s);
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
// the tile kernel
// ---------------------------------------------------------------------------------------------------
// CTA-wide barrier shared by the producer warp and the compute warps (they reach it from different code paths)
#ifndef EZ_NPROD
#define EZ_NPROD 1
#endif
constexpr int NPROD = EZ_NPROD;           // producer warps per CTA (each has its own limit of outstanding cp.async requests)
constexpr int NTHR_TILED = TT * TT + 32 * NPROD;
template <int NTHR_BAR = NTHR_TILED>
__device__ __forceinline__ void tile_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NTHR_BAR) : "memory"); }

// The producer warp: stages the tile's box of every source level into shared memory with cp.async (LDGSTS), one
// level ahead of the compute warps.  Lane l copies box columns l and l+32 (boxes are at most 64 wide) of every box
// row; per row that is two predicated 4-byte copies and two pointer bumps.  Box cells that overhang the field (a
// window may stick out by one zero-weight row/column on the high side) are never written and keep the zero fill.
__device__ __forceinline__ void stage_rows(float *__restrict__ dst, const float *__restrict__ src, int nrow, int src_ni, int pitch, bool c0ok,
                                           bool c1ok) {
#pragma unroll 4
    for (int r = 0; r < nrow; r++) {
        if (c0ok) __pipeline_memcpy_async(dst, src, 4);
        if (c1ok) __pipeline_memcpy_async(dst + 32, src + 32, 4);
        src += src_ni;
        dst += pitch;
    }
}

template <int NTHR_BAR = NTHR_TILED>
__device__ __forceinline__ void producer_warp(const InterpArgs &a, const float *__restrict__ zin, const TileDesc &td, int src_ni, int src_rows,
                                              float *__restrict__ smem, const int pitch, const int cap) {
    const int lane = threadIdx.x & 31;
    const int pw = (threadIdx.x - (NTHR_BAR - 32 * NPROD)) >> 5;    // this producer warp copies box rows pw, pw + NPROD, ...
    const int nrow = (min(td.h, src_rows - td.r0) - pw + NPROD - 1) / NPROD;
    const bool c0ok = lane < td.w && (td.c0 + lane) < src_ni;
    const bool c1ok = (lane + 32) < td.w && (td.c0 + lane + 32) < src_ni;
    const float *src = zin + (long)(td.r0 + pw) * src_ni + td.c0 + lane;
    const int nf = a.nfields;
    const long is = a.in_stride;
    // A stage holds LPS consecutive levels; the producer runs NSTAGE-1 stages ahead.  Empty commit groups keep the
    // group count uniform at the tail.
    const int nstages = (nf + LPS - 1) / LPS;
    auto stage = [&](int st) {
        if (st < nstages) {
            float *dst = smem + (st % NSTAGE) * LPS * cap + pw * pitch + lane;
#pragma unroll 1
            for (int l = 0; l < LPS; l++) {
                const int f = st * LPS + l;
                if (f < nf) stage_rows(dst + l * cap, src + f * is, nrow, NPROD * src_ni, NPROD * pitch, c0ok, c1ok);
            }
        }
        __pipeline_commit();
    };
#pragma unroll
    for (int k = 0; k < NSTAGE - 1; k++) stage(k);
    for (int st = 0; st < nstages; st++) {
        __pipeline_wait_prior(NSTAGE - 2);   // stage st has landed
        tile_barrier<NTHR_BAR>();            // ... and every compute warp is done with stage st-1, whose buffers are refilled next
        stage(st + NSTAGE - 1);
    }
}

// The compute warps.  Arithmetic of one level for one thread (2 x 2 target points p0 p1 / p2 p3, W x W window):
//   d_rc = z_rc - zc                                  (zc = window centre; fp32 difference form)
//   out_p = zc (1 + s_p) + sum_r wy_p[r] * ( sum_c wx_p[c] * d_rc )
// evaluated with packed fp32x2 instructions (FFMA2/FMUL2: sm_100 issues two fp32 FMAs per instruction, measured
// 114 vs 72 FMA/clk/SM for scalar FFMA, scripts/ubench/ffma2.cu): the two points of a row pair share the data
// operand (d, d) and carry their own weights in the two halves.  The window is streamed row by row, so only one
// row of data is live in registers.
template <int DEG>
__device__ __forceinline__ void compute_warps(const GridDev &g, const InterpArgs &a, const TileDesc &td, const Quad &q, bool fast,
                                              const float *__restrict__ smem) {
    constexpr int S = Sten<DEG>::S, W = Sten<DEG>::W;
    const int ni = a.out_ni, nj = a.out_nj;
    const int pitch = td.w | 1;
    float2 wxp[2][W], wyp[2][W], smp[2];   // [pair][k] = (weight of point 2*pair, weight of point 2*pair+1)
    bool valid[4] = {false, false, false, false};
    int base = 0;
    if (fast) {
        int c0[4], r0[4], wc0 = INT_MAX, wr0 = INT_MAX;
        float tx[4][4], ty[4][4], sm1[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            valid[p] = q.valid(p, ni, nj);
            c0[p] = INT_MAX; r0[p] = INT_MAX; sm1[p] = 0.f;
            if (valid[p]) {
                const int n = q.index(p, ni);
                point_plan<DEG>(g, a.x[n], a.y[n], c0[p], r0[p], tx[p], ty[p], sm1[p]);
                wc0 = min(wc0, c0[p]); wr0 = min(wr0, r0[p]);
            }
        }
        float wx[4][W], wy[4][W];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const bool sx = valid[p] && (c0[p] != wc0), sy = valid[p] && (r0[p] != wr0);   // offset 0 or 1 inside the window
#pragma unroll
            for (int k = 0; k < W; k++) {
                const float lo = (k < S) ? tx[p][k] : 0.f, hi = (k > 0) ? tx[p][k - 1] : 0.f;
                wx[p][k] = valid[p] ? (sx ? hi : lo) : 0.f;
                const float lo2 = (k < S) ? ty[p][k] : 0.f, hi2 = (k > 0) ? ty[p][k - 1] : 0.f;
                wy[p][k] = valid[p] ? (sy ? hi2 : lo2) : 0.f;
            }
        }
#pragma unroll
        for (int pp = 0; pp < 2; pp++) {
#pragma unroll
            for (int k = 0; k < W; k++) {
                wxp[pp][k] = make_float2(wx[2 * pp][k], wx[2 * pp + 1][k]);
                wyp[pp][k] = make_float2(wy[2 * pp][k], wy[2 * pp + 1][k]);
            }
            smp[pp] = make_float2(sm1[2 * pp], sm1[2 * pp + 1]);
        }
        base = (wr0 - td.r0) * pitch + (wc0 - td.c0);
    }
    const int nf = a.nfields;
    const long os = a.out_stride;
    // float2 stores need all four points and 8-byte alignment (even target row length; q.ox is even)
    const bool full = fast && ((ni & 1) == 0) && valid[3] && ((reinterpret_cast<uintptr_t>(a.zout) & 7) == 0) && ((os & 1) == 0);
    float *w0 = a.zout + (long)q.oy * ni + q.ox;
    const int cap = a.tile_cap;
    const int nstages = (nf + LPS - 1) / LPS;
    for (int st = 0; st < nstages; st++) {
        tile_barrier();
        const float *sbase = smem + (st % NSTAGE) * LPS * cap + base;
        const int nl = min(LPS, nf - st * LPS);
        if (fast) {
#pragma unroll TILED_UNROLL
            for (int l = 0; l < nl; l++) {
                const float *s = sbase + l * cap;
                const float zc = s[(W / 2) * pitch + W / 2];
                const float2 zc2 = make_float2(zc, zc);
                float2 acc[2];
#pragma unroll
                for (int r = 0; r < W; r++) {
                    float2 d[W];
#pragma unroll
                    for (int c = 0; c < W; c++) {
                        const float v = s[r * pitch + c] - zc;
                        d[c] = make_float2(v, v);
                    }
#pragma unroll
                    for (int pp = 0; pp < 2; pp++) {
                        float2 rs = __fmul2_rn(wxp[pp][0], d[0]);
#pragma unroll
                        for (int c = 1; c < W; c++) rs = __ffma2_rn(wxp[pp][c], d[c], rs);
                        acc[pp] = (r == 0) ? __fmul2_rn(wyp[pp][0], rs) : __ffma2_rn(wyp[pp][r], rs, acc[pp]);
                    }
                }
                const float2 r01 = __fadd2_rn(zc2, __ffma2_rn(zc2, smp[0], acc[0]));
                const float2 r23 = __fadd2_rn(zc2, __ffma2_rn(zc2, smp[1], acc[1]));
                float *w1 = w0 + ni;
                if (full) {
                    *reinterpret_cast<float2 *>(w0) = r01;
                    *reinterpret_cast<float2 *>(w1) = r23;
                } else {
                    if (valid[0]) w0[0] = r01.x;
                    if (valid[1]) w0[1] = r01.y;
                    if (valid[2]) w1[0] = r23.x;
                    if (valid[3]) w1[1] = r23.y;
                }
                w0 += os;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// EZ_TILED_NOPROD variant: same arithmetic, staging by every thread (16-byte cp.async), one barrier per stage
// ---------------------------------------------------------------------------------------------------
constexpr int NTHR_WIN = TILED_NOPROD ? TT * TT : NTHR_TILED;

template <int DEG>
__device__ __forceinline__ void window_all(const GridDev &g, const InterpArgs &a, const TileDesc &td, const Quad &q, bool fast,
                                           const float *__restrict__ zin, int src_rows, float *__restrict__ smem) {
    constexpr int S = Sten<DEG>::S, W = Sten<DEG>::W;
    const int ni = a.out_ni, nj = a.out_nj;
    const int pitch = win_pitch(td.w, g.ni);
    float2 wxp[2][W], wyp[2][W], smp[2];   // [pair][k] = (weight of point 2*pair, weight of point 2*pair+1)
    bool valid[4] = {false, false, false, false};
    int base = 0;
    if (fast) {
        int c0[4], r0[4], wc0 = INT_MAX, wr0 = INT_MAX;
        float tx[4][4], ty[4][4], sm1[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            valid[p] = q.valid(p, ni, nj);
            c0[p] = INT_MAX; r0[p] = INT_MAX; sm1[p] = 0.f;
            if (valid[p]) {
                const int n = q.index(p, ni);
                cached_plan<DEG>(g, a, n, c0[p], r0[p], tx[p], ty[p], sm1[p]);
                wc0 = min(wc0, c0[p]); wr0 = min(wr0, r0[p]);
            }
        }
        float wx[4][W], wy[4][W];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const bool sx = valid[p] && (c0[p] != wc0), sy = valid[p] && (r0[p] != wr0);   // offset 0 or 1 inside the window
#pragma unroll
            for (int k = 0; k < W; k++) {
                const float lo = (k < S) ? tx[p][k] : 0.f, hi = (k > 0) ? tx[p][k - 1] : 0.f;
                wx[p][k] = valid[p] ? (sx ? hi : lo) : 0.f;
                const float lo2 = (k < S) ? ty[p][k] : 0.f, hi2 = (k > 0) ? ty[p][k - 1] : 0.f;
                wy[p][k] = valid[p] ? (sy ? hi2 : lo2) : 0.f;
            }
        }
#pragma unroll
        for (int pp = 0; pp < 2; pp++) {
#pragma unroll
            for (int k = 0; k < W; k++) {
                wxp[pp][k] = make_float2(wx[2 * pp][k], wx[2 * pp + 1][k]);
                wyp[pp][k] = make_float2(wy[2 * pp][k], wy[2 * pp + 1][k]);
            }
            smp[pp] = make_float2(sm1[2 * pp], sm1[2 * pp + 1]);
        }
        base = (wr0 - td.r0) * pitch + (wc0 - td.c0);
    }
    const int nf = a.nfields;
    const long os = a.out_stride, is = a.in_stride;
    // float2 stores need all four points and 8-byte alignment (even target row length; q.ox is even)
    const bool full = fast && ((ni & 1) == 0) && valid[3] && ((reinterpret_cast<uintptr_t>(a.zout) & 7) == 0) && ((os & 1) == 0);
    float *w0 = a.zout + (long)q.oy * ni + q.ox;
    const int cap = a.tile_cap;
    const int nstages = (nf + LPS - 1) / LPS;

    // ---- copy side (see ez_interp_rowsep.cuh): box element (r, j) of level l of a stage at float offset delta_l + r pitch + j
    // of the level's buffer, delta_l = (float index in memory of source element (r0, c0)) mod 4; thread t copies chunks t and
    // t + 256 of each level (chunk = row t / nck, 16-byte chunk t % nck of that row).
    const uint32_t smem32 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t capB = (uint32_t)cap * 4u;
    const int nrow = min(td.h, src_rows - td.r0);
    const int nck = (td.w + 6) >> 2;
    unsigned int dpk = 0u;
    int cj[2] = {0, 0}, crowoff[2] = {0, 0};
    unsigned int cvalid = 0u;     // bit 4 k + l: slot k, level l
    uint32_t cbase[2] = {0u, 0u};
    {
        const float *lvl0 = zin + (long)td.r0 * g.ni + td.c0;
#pragma unroll
        for (int l = 0; l < LPS; l++) dpk |= (unsigned int)((reinterpret_cast<uintptr_t>(lvl0 + l * is) >> 2) & 3u) << (2 * l);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int t = threadIdx.x + k * NTHR_WIN;
            const int crow = t / nck, ck4 = 4 * (t - crow * nck);
            crowoff[k] = (td.r0 + crow) * g.ni + td.c0 - 3;
            cbase[k] = (uint32_t)(crow * pitch - 3) * 4u;
#pragma unroll
            for (int l = 0; l < LPS; l++) {
                const int bm = (int)((dpk >> (2 * l)) & 3u) + crow * pitch;
                const int j = ((bm & ~3) + ck4) - bm;
                cj[k] |= (j + 3) << (8 * l);
                if (crow < nrow && j < td.w) cvalid |= 1u << (4 * k + l);
            }
        }
    }
    auto sloff = [&](int l) { return (uint32_t)l * capB + ((dpk >> (2 * l)) & 3u) * 4u; };
    auto issue = [&](int st) {
        if (cvalid && st < nstages) {
            const uint32_t sb = smem32 + (uint32_t)(st % NSTAGE) * (LPS * capB);
#pragma unroll
            for (int k = 0; k < 2; k++) {
                if ((cvalid >> (4 * k)) & 0xfu) {
                    const float *p = zin + (long)(st * LPS) * is + crowoff[k];
#pragma unroll
                    for (int l = 0; l < LPS; l++) {
                        if (((cvalid >> (4 * k + l)) & 1u) && st * LPS + l < nf) {
                            const uint32_t jb = (uint32_t)(cj[k] >> (8 * l)) & 0xffu;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + cbase[k] + sloff(l) + 4u * jb), "l"(p + jb) : "memory");
                        }
                        p += is;
                    }
                }
            }
        }
        __pipeline_commit();
    };
#pragma unroll 1
    for (int k = 0; k < NSTAGE; k++) issue(k);
    __pipeline_wait_prior(NSTAGE - 1);
    __syncthreads();

    int slot = 0;
#pragma unroll 1
    for (int st = 0; st < nstages; st++) {
        const float *sbase = smem + slot * LPS * cap + base;
        const int nl = min(LPS, nf - st * LPS);
        if (fast) {
#pragma unroll TILED_UNROLL
            for (int l = 0; l < nl; l++) {
                const float *s = sbase + l * cap + ((dpk >> (2 * l)) & 3u);
                const float zc = s[(W / 2) * pitch + W / 2];
                const float2 zc2 = make_float2(zc, zc);
                float2 acc[2];
#pragma unroll
                for (int r = 0; r < W; r++) {
                    float2 d[W];
#pragma unroll
                    for (int c = 0; c < W; c++) {
                        const float v = s[r * pitch + c] - zc;
                        d[c] = make_float2(v, v);
                    }
#pragma unroll
                    for (int pp = 0; pp < 2; pp++) {
                        float2 rs = __fmul2_rn(wxp[pp][0], d[0]);
#pragma unroll
                        for (int c = 1; c < W; c++) rs = __ffma2_rn(wxp[pp][c], d[c], rs);
                        acc[pp] = (r == 0) ? __fmul2_rn(wyp[pp][0], rs) : __ffma2_rn(wyp[pp][r], rs, acc[pp]);
                    }
                }
                const float2 r01 = __fadd2_rn(zc2, __ffma2_rn(zc2, smp[0], acc[0]));
                const float2 r23 = __fadd2_rn(zc2, __ffma2_rn(zc2, smp[1], acc[1]));
                float *w1 = w0 + ni;
                if (full) {
                    *reinterpret_cast<float2 *>(w0) = r01;
                    *reinterpret_cast<float2 *>(w1) = r23;
                } else {
                    if (valid[0]) w0[0] = r01.x;
                    if (valid[1]) w0[1] = r01.y;
                    if (valid[2]) w1[0] = r23.x;
                    if (valid[3]) w1[1] = r23.y;
                }
                w0 += os;
            }
        }
        __pipeline_wait_prior(NSTAGE - 2);   // stage st + 1 has landed (this thread's share)
        __syncthreads();                     // ... everybody's, and everybody is done reading stage st
        issue(st + NSTAGE);                  // refill the slot just read
        if (++slot == NSTAGE) slot = 0;
    }
}

template <int DEG>
__global__ void __launch_bounds__(NTHR_WIN, EZ_TILED_MINB) k_interp_tiled(const __grid_constant__ InterpArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int by = remap_tile_row(blockIdx.y, gridDim.y);
    const int tile = by * gridDim.x + blockIdx.x;
    const TileDesc td = a.tiles[tile];
    if (td.nfast == 0 || td.kind != 0) return;   // kind 1: the row-separable kernel owns the tile
    for (int k = threadIdx.x; k < NSTAGE * LPS * a.tile_cap; k += NTHR_WIN) smem[k] = 0.f;
    __syncthreads();
    if (TILED_NOPROD) {
        const bool fast = (a.tile_flags[tile * 8 + (threadIdx.x >> 5)] >> (threadIdx.x & 31)) & 1u;
        const Quad q = quad_of_thread(blockIdx.x, by);
        if (td.panel) window_all<DEG>(a.src2, a, td, q, fast, a.zin + a.panel_stride, a.src2.j2 - a.src2.j1 + 1, smem);
        else window_all<DEG>(a.src, a, td, q, fast, a.zin, a.src.j2 - a.src.j1 + 1, smem);
        return;
    }
    if (threadIdx.x >= TT * TT) {
        if (td.panel) producer_warp(a, a.zin + a.panel_stride, td, a.src2.ni, a.src2.j2 - a.src2.j1 + 1, smem, td.w | 1, a.tile_cap);
        else producer_warp(a, a.zin, td, a.src.ni, a.src.j2 - a.src.j1 + 1, smem, td.w | 1, a.tile_cap);
        return;
    }
    const bool fast = (a.tile_flags[tile * 8 + (threadIdx.x >> 5)] >> (threadIdx.x & 31)) & 1u;
    const Quad q = quad_of_thread(blockIdx.x, by);
    if (td.panel) compute_warps<DEG>(a.src2, a, td, q, fast, smem);
    else compute_warps<DEG>(a.src, a, td, q, fast, smem);
}

}  // namespace
