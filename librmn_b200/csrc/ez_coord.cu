// librmn_b200/csrc/ez_coord.cu -- gridset-build kernels (K12-K15 in SURVEY.md 2.3), sm_100a.
//
// One thread per target point; every kernel is a single streaming pass (coalesced loads/stores of the
// per-point arrays, axis tables read through the read-only path and resident in L1/L2).  These run once
// per (source,target) pair; the results (lat, lon, x, y, zone bits) stay in HBM for every later call.
//
// Compiled with -fmad=false (see ez_devmath.cuh).
#include "ez_devmath.cuh"
#include "ez_kernels.h"

namespace {

constexpr int TPB = 256;
inline int nblocks(long n) { return (int)((n + TPB - 1) / TPB); }

// ---------------------------------------------------------------------------------------------------
// K12: latitude/longitude of every point of a grid.  Restates ez_calclatlon.c:47-251.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_calclatlon(GridDev g, float *__restrict__ lat, float *__restrict__ lon, int npts) {
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= npts) return;
    int i = k % g.ni, j = k / g.ni;
    float la = 0.f, lo = 0.f;
    // permut() of the latitudes for A/B/G with ig2 == 1 (ez_calclatlon.c:253-266): row j takes row nj-1-j
    int jl = j;
    if ((g.grtyp == 'G' || g.grtyp == 'B' || g.grtyp == 'A') && g.ig[1] == 1) jl = g.nj - 1 - j;
    switch (g.grtyp) {
        case 'A':
        case 'B':
        case 'L':
            // grll.f:57-64
            la = g.llp[0] + (float)jl * g.llp[2];
            lo = fmodf(g.llp[1] + (float)i * g.llp[3], 360.0f);
            break;
        case 'E': {
            // ez_calclatlon.c:63-82: regular lat-lon in the rotated frame, then back to true lat-lon (ri)
            float lar = g.llp[0] + (float)j * g.llp[2];
            float lor = fmodf(g.llp[1] + (float)i * g.llp[3], 360.0f);
            ez_rotate_ll(g.ri, lor, lar, lo, la);
            break;
        }
        case 'N':
        case 'S': {
            // grps.f:63-73
            float y = (float)(j + 1) - g.xg[1];
            float x = (float)(i + 1) - g.xg[0];
            ez_llfxy(la, lo, x, y, g.xg[2], g.xg[3], g.grtyp == 'N' ? 1 : 2);
            if (lo < 0.f) lo = lo + 360.f;
            break;
        }
        case 'Y':
            // ez_calclatlon.c:136-147
            lo = __ldg(g.ax + k);
            la = __ldg(g.ay + k);
            if (lo < 0.0f) lo = lo + 360.0f;
            break;
        case '#':
        case 'Z':
        case 'G': {
            // ez_calclatlon.c:160-226
            la = __ldg(g.ay + jl);
            lo = __ldg(g.ax + i);
            if (g.grtyp == 'G' && g.ig[0] == 1) la = __ldg(g.ay + jl + g.nj);
            switch (g.grref) {
                case 'N':
                case 'S': {
                    float la2, lo2;
                    ez_vllfxy1(la2, lo2, lo, la, g.xgref[2], g.xgref[3], g.xgref[0], g.xgref[1], g.hemisphere);
                    if (lo2 < 0.0f) lo2 = lo2 + 360.0f;
                    la = la2;
                    lo = lo2;
                    break;
                }
                case 'L': {
                    // "+= 1.0" then c_llfgr (ez_llfgr.c:27-35): double arithmetic on fp32 operands
                    float y1 = (float)((double)la + 1.0);
                    float x1 = (float)((double)lo + 1.0);
                    float l = (float)((double)g.xgref[1] + (double)g.xgref[3] * ((double)x1 - 1.0));
                    lo = (float)fmod(fmod((double)l, 360.0) + 360.0, 360.0);
                    la = (float)((double)g.xgref[0] + (double)g.xgref[2] * ((double)y1 - 1.0));
                    break;
                }
                case 'E': {
                    float la2, lo2;
                    ez_rotate_ll(g.ri, lo, la, lo2, la2);
                    la = la2;
                    lo = lo2;
                    break;
                }
                default:
                    break;
            }
            break;
        }
        default:
            break;
    }
    lat[k] = la;
    lon[k] = lo;
}

// ---------------------------------------------------------------------------------------------------
// K13 + K14: x,y of every target point in source index space, and its zone bits.
// Restates ez_calcxy.c:59-97 and ez_defzones.c:83-109 + ez_defzone_*.c.
// ---------------------------------------------------------------------------------------------------
struct ZoneParams {
    int enabled;        // build zone bits at all
    int extrap;         // EZ_EXTRAP (1) -> DEHORS only ; EZ_NO_EXTRAP (0) -> the four polar zones
    int ni, nj, j1, j2; // source dims
    float ypolenord, ypolesud;
};

__device__ __forceinline__ uint8_t zone_bits(const ZoneParams &z, float x, float y) {
    uint8_t b = 0;
    if (z.extrap) {
        // ez_defzone_dehors.c:65-67 : (int)(x + 0.5) with the sum in double
        int ix = (int)((double)x + 0.5);
        int iy = (int)((double)y + 0.5);
        if (ix < 1 || iy < 1 || ix > z.ni || iy > z.nj) b |= ZB_DEHORS;
    } else {
        if (fabs((double)(y - z.ypolenord)) < 1.0e-3) b |= ZB_POLE_NORD;   // ez_defzone_polenord.c:53
        if (fabs((double)(y - z.ypolesud)) < 1.0e-3) b |= ZB_POLE_SUD;     // ez_defzone_polesud.c:57
        if ((int)y < z.j1 + 1) b |= ZB_AU_SUD;                             // ez_defzone_sud.c:42-44
        if ((int)y > z.j2 - 2) b |= ZB_AU_NORD;                            // ez_defzone_nord.c:41-43
    }
    return b;
}

__global__ void __launch_bounds__(TPB) k_calcxy(GridDev src, int njs, const float *__restrict__ lat, float *__restrict__ lon,
                                                 int writeback_lon, float *__restrict__ x, float *__restrict__ y,
                                                 uint8_t *__restrict__ zb, ZoneParams zp, unsigned int *__restrict__ counts, int npts) {
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= npts) return;
    float la = lat[k], lo = lon[k], lo0 = lo, px, py;
    ez_ll2xy(src, njs, la, lo, px, py);
    x[k] = px;
    y[k] = py;
    if (writeback_lon && lo != lo0) lon[k] = lo;
    if (zb) {
        uint8_t b = zp.enabled ? zone_bits(zp, px, py) : 0;
        zb[k] = b;
        if (b && counts) {
            if (b & ZB_DEHORS) atomicAdd(counts + 0, 1u);
            if (b & ZB_AU_NORD) atomicAdd(counts + 1, 1u);
            if (b & ZB_AU_SUD) atomicAdd(counts + 2, 1u);
            if (b & ZB_POLE_NORD) atomicAdd(counts + 3, 1u);
            if (b & ZB_POLE_SUD) atomicAdd(counts + 4, 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K15: Yin-Yang panel selection + per-panel x,y, without the reference's compaction/merge.
// The reference (ezyymint.c:42-98, ezyy_calcxy.c:89-111) nearest-interpolates an all-zero field from
// Yin's "mask grid" with extrap_degree=value 1.0, i.e. a target point goes to Yang iff it is in the
// DEHORS zone of the mask grid; then each compacted list gets x,y on its own panel via c_gdxyfll_orig.
// Here: one pass, panel bit + x,y per point, in place.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_yy_calcxy(GridDev maskgrid, GridDev yin, GridDev yan, int use_mask,
                                                    const float *__restrict__ lat, const float *__restrict__ lon,
                                                    float *__restrict__ x, float *__restrict__ y, uint8_t *__restrict__ zb,
                                                    unsigned int *__restrict__ counts, int npts) {
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= npts) return;
    float la = lat[k], lo = lon[k];
    int yang = 0;
    if (use_mask) {
        float mx, my, lom = lo;
        ez_ll2xy(maskgrid, maskgrid.nj, la, lom, mx, my);          // ez_calcxy route (njs = nj)
        int ix = (int)((double)mx + 0.5);
        int iy = (int)((double)my + 0.5);
        yang = (ix < 1 || iy < 1 || ix > maskgrid.ni || iy > maskgrid.nj);
    }
    float px, py, lop = lo;
    if (yang)
        ez_ll2xy(yan, yan.j2, la, lop, px, py);                    // c_gdxyfll_orig route (njs = j2)
    else
        ez_ll2xy(yin, yin.j2, la, lop, px, py);
    x[k] = px;
    y[k] = py;
    zb[k] = yang ? ZB_YANG : 0;
    if (yang && counts) atomicAdd(counts + 5, 1u);
}

// generic caller-supplied points: (lat,lon) -> (x,y)   [c_gdxyfll_orig, gdxyfll.c:141-187]
__global__ void __launch_bounds__(TPB) k_ll2xy_points(GridDev g, int njs, const float *__restrict__ lat, const float *__restrict__ lon,
                                                       float *__restrict__ x, float *__restrict__ y, int npts) {
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= npts) return;
    float lo = lon[k], px, py;
    ez_ll2xy(g, njs, lat[k], lo, px, py);
    x[k] = px;
    y[k] = py;
}

// caller-supplied points: (x, y) -> (lat, lon)   [c_gdllfxy, gdllfxy.c:30-92: 'U' grids split by y > nj of the Yin panel]
__global__ void __launch_bounds__(TPB) k_xy2ll_points(GridDev g, GridDev g2, int is_u, const float *__restrict__ x, const float *__restrict__ y,
                                                       float *__restrict__ lat, float *__restrict__ lon, int npts) {
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= npts) return;
    float la, lo;
    const float xf = x[k], yf = y[k];
    if (!is_u) {
        ez_xy2ll(g, xf, yf, 1, la, lo);
    } else if (yf > (float)g.nj) {
        ez_xy2ll(g2, xf, yf - (float)g.nj, 0, la, lo);
    } else {
        ez_xy2ll(g, xf, yf, 0, la, lo);
    }
    lat[k] = la;
    lon[k] = lo;
}

// (lat, lon) -> (x, y) on a 'U' supergrid: both panels, then the selection of gdxyfll.c:199-236 (the mask-core bounds
// are the reference's 0-based indices compared against 1-based coordinates, as written there)
__global__ void __launch_bounds__(TPB) k_ll2xy_u(GridDev yin, GridDev yan, int4 myin, int4 myan, const float *__restrict__ lat,
                                                  const float *__restrict__ lon, float *__restrict__ x, float *__restrict__ y, int npts) {
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= npts) return;
    const float la = lat[k];
    float lo1 = lon[k], lo2 = lo1, xi, yi, xa, ya;
    ez_ll2xy(yin, yin.j2, la, lo1, xi, yi);
    ez_ll2xy(yan, yan.j2, la, lo2, xa, ya);
    const float maxni = (float)yin.ni, maxnj = (float)yin.nj;
    float xo, yo;
    if (xi > maxni || xi < 1.0f || yi > maxnj || yi < 1.0f) { xo = xa; yo = ya + maxnj; }
    else { xo = xi; yo = yi; }
    if (xa >= (float)myan.x && xa <= (float)myan.y && ya >= (float)myan.z && ya <= (float)myan.w) { xo = xa; yo = ya + maxnj; }
    if (xi >= (float)myin.x && xi <= (float)myin.y && yi >= (float)myin.z && yi <= (float)myin.w) { xo = xi; yo = yi; }
    x[k] = xo;
    y[k] = yo;
}

// split 'U' coordinates into a panel bit and the y inside the panel (gdxysval.c:70-81)
__global__ void __launch_bounds__(TPB) k_split_u(const float *__restrict__ y, int njyin, float *__restrict__ ty, uint8_t *__restrict__ zb, int npts) {
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= npts) return;
    const float yf = y[k];
    const bool yang = yf > (float)njyin;
    ty[k] = yang ? yf - (float)njyin : yf;
    zb[k] = yang ? ZB_YANG : 0;
}

}  // namespace

namespace ezk {
void xy2ll_points(const GridDev &g, const GridDev &g2, int is_u, const float *x, const float *y, float *lat, float *lon, int npts, cudaStream_t s) {
    if (npts <= 0) return;
    k_xy2ll_points<<<nblocks(npts), TPB, 0, s>>>(g, g2, is_u, x, y, lat, lon, npts);
}
void ll2xy_u(const GridDev &yin, const GridDev &yan, const int myin[4], const int myan[4], const float *lat, const float *lon, float *x, float *y,
             int npts, cudaStream_t s) {
    if (npts <= 0) return;
    k_ll2xy_u<<<nblocks(npts), TPB, 0, s>>>(yin, yan, make_int4(myin[0], myin[1], myin[2], myin[3]), make_int4(myan[0], myan[1], myan[2], myan[3]),
                                            lat, lon, x, y, npts);
}
void split_u(const float *y, int njyin, float *ty, uint8_t *zb, int npts, cudaStream_t s) {
    if (npts <= 0) return;
    k_split_u<<<nblocks(npts), TPB, 0, s>>>(y, njyin, ty, zb, npts);
}

void calclatlon(const GridDev &g, float *lat, float *lon, int npts, cudaStream_t s) {
    if (npts <= 0) return;
    k_calclatlon<<<nblocks(npts), TPB, 0, s>>>(g, lat, lon, npts);
}

void calcxy(const GridDev &src, int njs, const float *lat, float *lon, int writeback_lon, float *x, float *y, uint8_t *zb,
            int zones_enabled, int extrap, float ypolenord, float ypolesud, unsigned int *counts, int npts, cudaStream_t s) {
    if (npts <= 0) return;
    ZoneParams zp;
    zp.enabled = zones_enabled;
    zp.extrap = extrap;
    zp.ni = src.ni;
    zp.nj = src.nj;
    zp.j1 = src.j1;
    zp.j2 = src.j2;
    zp.ypolenord = ypolenord;
    zp.ypolesud = ypolesud;
    k_calcxy<<<nblocks(npts), TPB, 0, s>>>(src, njs, lat, lon, writeback_lon, x, y, zb, zp, counts, npts);
}

void yy_calcxy(const GridDev &maskgrid, const GridDev &yin, const GridDev &yan, int use_mask, const float *lat, const float *lon,
               float *x, float *y, uint8_t *zb, unsigned int *counts, int npts, cudaStream_t s) {
    if (npts <= 0) return;
    k_yy_calcxy<<<nblocks(npts), TPB, 0, s>>>(maskgrid, yin, yan, use_mask, lat, lon, x, y, zb, counts, npts);
}

void ll2xy_points(const GridDev &g, int njs, const float *lat, const float *lon, float *x, float *y, int npts, cudaStream_t s) {
    if (npts <= 0) return;
    k_ll2xy_points<<<nblocks(npts), TPB, 0, s>>>(g, njs, lat, lon, x, y, npts);
}

}  // namespace ezk
