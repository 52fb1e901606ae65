// librmn_b200/csrc/ez_wind.cu -- wind-vector rotation (K17 of SURVEY.md 2.3), sm_100a.
//
// The reference turns grid-frame (u,v) into (speed, meteorological direction) on the SOURCE grid's frame
// (c_gdwdfuv_orig, gdwdfuv.c:43-97) and then into (u,v) on the TARGET grid's frame (c_gduvfwd_orig,
// gduvfwd.c:43-93) in about ten full passes over the output with four temporaries malloc'd per call
// (ezuvint.c:81-86, ez_llwfgfw.c:58-69).  Here both stages are fused into one pass: one thread per target
// point, (u,v) of every field updated in place, the per-point trigonometry (rotated lat/lon, sines and
// cosines) evaluated once and reused for all fields.  Compiled with -fmad=false.
#include "ez_devmath.cuh"
#include "ez_kernels.h"

namespace {

constexpr int TPB = 256;

struct Rot {          // per-point constants of a rotated ('E' / ref 'E') frame
    double sa, ca, sc, cc;   // sin/cos of DAR*lon_rot, DAR*lat_rot   (ez_uvacart on the rotated frame)
    double st, ct, se, ce;   // sin/cos of DAR*lon_true, DAR*lat_true (ez_cartauv on the true frame)
};

__device__ __forceinline__ bool is_rotated(const GridDev &g) {
    if (g.grtyp == 'E') return true;
    if ((g.grtyp == '#' || g.grtyp == 'Y' || g.grtyp == 'Z') && g.grref == 'E') return true;
    return false;
}
// grid-type letter and dgrw that ez_llwfgdw / ez_gdwfllw see (gdwdfuv.c:80-95, gduvfwd.c:81-92)
__device__ __forceinline__ char frame_type(const GridDev &g) {
    if (g.grtyp == '#' || g.grtyp == 'Y' || g.grtyp == 'Z') return g.grref;
    return g.grtyp;
}
__device__ __forceinline__ float frame_dgrw(const GridDev &g) {
    if (g.grtyp == '#' || g.grtyp == 'Y' || g.grtyp == 'Z') return g.xgref[3];
    return g.xg[3];
}

__device__ __forceinline__ float wrapdir(float d) { return fmodf(fmodf(d, 360.0f) + 360.0f, 360.0f); }

// ez_llwfgdw.inc:49-164 for one point; t = frame letter
__device__ __forceinline__ void llwfgdw1(char t, float dgrw, float xlon, float &z1, float &z2) {
    float uu = z1, vv = z2;
    if (!(t == 'N' || t == 'S' || t == 'A' || t == 'B' || t == 'G' || t == 'L')) return;   // "bad grid type": untouched
    float spd0 = sqrtf(uu * uu + vv * vv), dir0;
    if (spd0 == 0.0f) {
        dir0 = 0.0f;
    } else if (uu == 0.0f) {
        if (t == 'N') dir0 = (vv >= 0.0f) ? xlon + dgrw - 90.0f : xlon + dgrw + 90.0f;
        else if (t == 'S') dir0 = (vv >= 0.0f) ? 90.0f - xlon + dgrw : 270.0f - xlon + dgrw;
        else dir0 = (vv >= 0.0f) ? 180.0f : 0.0f;
    } else {
        float at = EZ_RDTODG_F * ez_atan2f(vv, uu);
        if (t == 'N') dir0 = xlon + dgrw - at;
        else if (t == 'S') dir0 = 180.0f - xlon + dgrw - at;
        else dir0 = 270.0f - at;
    }
    z1 = spd0;
    z2 = wrapdir(dir0);
}

// ez_gdwfllw.inc:89-133 for one point
__device__ __forceinline__ void gdwfllw1(char t, float dgrw, float xlon, float &z1, float &z2) {
    float psi;
    if (t == 'N') psi = xlon + dgrw - z2;
    else if (t == 'S') psi = 180.0f - xlon + dgrw - z2;
    else if (t == 'A' || t == 'B' || t == 'G' || t == 'L') psi = 270.0f - z2;
    else return;
    float u = ez_cosf(psi * EZ_DGTORD_F) * z1;
    float v = ez_sinf(psi * EZ_DGTORD_F) * z1;
    z1 = u;
    z2 = v;
}

__device__ __forceinline__ Rot make_rot(const GridDev &g, float lat, float lon) {
    Rot r;
    float lonr, latr;
    ez_rotate_ll(g.r, lon, lat, lonr, latr);     // ez_gfxyfll: true -> rotated
    const double dar = (double)ez_dar();         // DAR = ACOS(-1.)/180. evaluated in fp32, ez_uvacart.inc:45
    sincos(dar * (double)lonr, &r.sa, &r.ca);
    sincos(dar * (double)latr, &r.sc, &r.cc);
    sincos(dar * (double)lon, &r.st, &r.ct);
    sincos(dar * (double)lat, &r.se, &r.ce);
    return r;
}

// ez_uvacart.inc:49-58 ; (lon,lat) sines/cosines passed in
__device__ __forceinline__ void uvacart1(float u, float v, double a, double b, double c, double d, float &x0, float &x1, float &x2) {
    x0 = (float)(-((double)u * a) - ((double)v * b * c));
    x1 = (float)(((double)u * b) - ((double)v * a * c));
    x2 = (float)((double)v * d);
}
// ez_cartauv.inc:48-58
__device__ __forceinline__ void cartauv1(float c0, float c1, float c2, double a, double b, double e, double f, float &u, float &v) {
    u = (float)(((double)c1 * a) - ((double)c0 * b));
    double c = ((double)c0 * a) + ((double)c1 * b);
    double d = sqrt(c * c + (double)(c2 * c2));
    v = (float)copysign(fabs(d), ((double)c2 * e) - (c * f));
}

__global__ void __launch_bounds__(TPB) k_wind(WindArgs a) {
    int n = blockIdx.x * TPB + threadIdx.x;
    if (n >= a.npts) return;
    const float lat = a.lat[n], lon = a.lon[n];
    const bool yang = a.yy && a.zb && (a.zb[n] & ZB_YANG);
    const GridDev &gs = yang ? a.src2 : a.src;
    const GridDev &gd = a.dst;
    const bool do1 = (a.mode != 2), do2 = (a.mode != 1);
    const bool rot_s = do1 && is_rotated(gs), rot_d = do2 && is_rotated(gd);
    Rot rs, rd;
    if (rot_s) rs = make_rot(gs, lat, lon);
    if (rot_d) rd = make_rot(gd, lat, lon);
    const char ts = frame_type(gs), td = frame_type(gd);
    const float dgs = frame_dgrw(gs), dgd = frame_dgrw(gd);
    for (int f = 0; f < a.nfields; f++) {
        long k = (long)f * a.stride + n;
        float u = a.u[k], v = a.v[k];
        if (do1) {
            if (rot_s) {
                // c_ezllwfgfw, ez_llwfgfw.c:62-69: rotated-frame winds -> true-frame winds -> (spd, dir)
                float x0, x1, x2, c0, c1, c2;
                uvacart1(u, v, rs.sa, rs.ca, rs.sc, rs.cc, x0, x1, x2);
                ez_mxm3(gs.ri, x0, x1, x2, c0, c1, c2);
                cartauv1(c0, c1, c2, rs.ct, rs.st, rs.ce, rs.se, u, v);
                llwfgdw1('L', 0.f, lon, u, v);
            } else {
                llwfgdw1(ts, dgs, lon, u, v);
            }
        }
        if (do2) {
            if (rot_d) {
                // c_ezgfwfllw, ez_gfwfllw.c:66-72: (spd, dir) -> true-frame winds -> rotated-frame winds
                float x0, x1, x2, c0, c1, c2;
                gdwfllw1('L', 0.f, lon, u, v);
                uvacart1(u, v, rd.st, rd.ct, rd.se, rd.ce, x0, x1, x2);
                ez_mxm3(gd.r, x0, x1, x2, c0, c1, c2);
                cartauv1(c0, c1, c2, rd.ca, rd.sa, rd.cc, rd.sc, u, v);
            } else {
                gdwfllw1(td, dgd, lon, u, v);
            }
        }
        a.u[k] = u;
        a.v[k] = v;
    }
}

// ---------------------------------------------------------------------------------------------------
// K18: the synthetic pole row of a vector field, ez_calcnpolarwind.c:59-125 / ez_calcspolarwind.c.
// One CTA per (field, pole).  The reference's sequence, per source column i of the row next to the pole:
//   (lat, lon) of the grid point (c_gdllfxy_orig; rotated-frame coordinates for Z/E sources, :70-83)
//   (u, v) -> (speed, direction) on an 'A' grid ; -> (u, v) on a polar-stereographic grid centred on the pole (:85-95)
//   pole value = row mean of those (u, v) (ez_calcpoleval: sequential fp32 sum) ; -> (speed, direction) at the pole (:97-103)
//   direction_i = direction_pole +/- lon_i, same speed ; -> (u, v) on the 'A' grid = the pole row (:106-114)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_point_lonlat(const GridDev &g, int i, int y, float &plon, float &plat) {
    // c_gdllfxy_orig(gdin, lat, lon, x = i + 1, y) ; then ez_gfxyfll for Z/E (ez_calcnpolarwind.c:70-83)
    float lon, lat;
    ez_xy2ll(g, (float)(i + 1), (float)y, 0, lat, lon);
    if (g.grtyp == 'Z' && g.grref == 'E') {
        float lr, ar;
        ez_rotate_ll(g.r, lon, lat, lr, ar);
        lon = lr;
        lat = ar;
    }
    plon = lon;
    plat = lat;
}

constexpr int PW_CHUNK = 1024;
__global__ void __launch_bounds__(256) k_polar_wind(PolarWindArgs a) {
    __shared__ float s_u[PW_CHUNK], s_v[PW_CHUNK];
    __shared__ float s_acc[2], s_pole[2];
    const int f = blockIdx.x >> 1, south = blockIdx.x & 1;
    const GridDev &g = a.src;
    const int ni = g.ni;
    // The row next to the pole of the LOGICAL source field (j2 north, j1 south), through the same index folding as the gather
    // kernels' FieldFetch: hemispheric sources are mirrored about the equator (u symmetric, v antisymmetric, ezuvint.c:65-71),
    // y-inverted sources are stored north to south.  (The reference hands its pole-wind routines the caller's raw, unexpanded
    // array together with the expanded row count, ez_corrvec_aunord.c:45-62 -> ez_calcnpolarwind.c:85: it reads past the end of a
    // hemispheric field and the wrong rows of a y-inverted one; this is the evaluation its own expanded field would give.)
    int jj = south ? g.j1 : g.j2;
    float sgn_v = 1.0f;
    switch (g.xpn) {
        case XPN_AG_NORTH: if (jj < 1) { jj = 1 - jj; sgn_v = -1.0f; } break;
        case XPN_AG_SOUTH: if (jj > g.nj) { jj = 2 * g.nj - jj + 1; sgn_v = -1.0f; } break;
        case XPN_B_NORTH:  if (jj < 1) { jj = 2 - jj; sgn_v = -1.0f; } break;
        case XPN_B_SOUTH:  if (jj > g.nj) { jj = 2 * g.nj - jj; sgn_v = -1.0f; } break;
        default: break;
    }
    if (g.yinv) jj = g.nj + 1 - jj;
    const int row = 1;   // (only the longitude of the column enters below)
    const int srow = (g.xpn == XPN_NONE && !g.yinv) ? jj - g.j1 : jj - 1;   // storage row (plain fields start at logical row j1)
    const float *u = a.u + (long)f * a.stride + (long)srow * ni;
    const float *v = a.v + (long)f * a.stride + (long)srow * ni;
    const char t = south ? 'S' : 'N';
    const float dgrw = south ? a.dgrw_s : a.dgrw_n;
    const bool weighted = (g.grtyp == 'Z' && g.grref == 'E');
    const int n = weighted ? ni - 1 : ni;
    if (threadIdx.x == 0) { s_acc[0] = 0.f; s_acc[1] = 0.f; }
    for (int i0 = 0; i0 < n; i0 += PW_CHUNK) {
        const int m = min(PW_CHUNK, n - i0);
        __syncthreads();
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const int i = i0 + k;
            float plon, plat;
            src_point_lonlat(g, i, row, plon, plat);
            float z1 = u[i], z2 = sgn_v * v[i];
            llwfgdw1('A', 0.f, plon, z1, z2);       // c_gdwdfuv on the auxiliary 'A' grid
            gdwfllw1(t, dgrw, plon, z1, z2);        // c_gduvfwd on the auxiliary polar-stereographic grid
            if (weighted) {
                const float dx = __ldg(g.ax + i + 1) - __ldg(g.ax + i);
                z1 = z1 * dx;
                z2 = z2 * dx;
            }
            s_u[k] = z1;
            s_v[k] = z2;
        }
        __syncthreads();
        if (threadIdx.x < 2) {   // ez_calcpoleval.inc:30-45: sequential fp32 accumulation
            const float *sv = threadIdx.x ? s_v : s_u;
            float p = s_acc[threadIdx.x];
            for (int k = 0; k < m; k++) p = p + sv[k];
            s_acc[threadIdx.x] = p;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float up = s_acc[0], vp = s_acc[1];
        if (weighted) {
            const float span = __ldg(g.ax + ni - 1) - __ldg(g.ax);
            if (span != 0.0f) { up = up / span; vp = vp / span; }
        } else {
            up = up / (1.0f * (float)ni);
            vp = vp / (1.0f * (float)ni);
        }
        llwfgdw1(t, dgrw, 0.0f, up, vp);            // (speed, direction) at the pole, lon = 0
        s_pole[0] = up;
        s_pole[1] = vp;
    }
    __syncthreads();
    const float spd0 = s_pole[0], wd0 = s_pole[1];
    float *pu = a.prow_u + ((long)f * 2 + south) * ni, *pv = a.prow_v + ((long)f * 2 + south) * ni;
    for (int i = threadIdx.x; i < ni; i += blockDim.x) {
        float plon, plat;
        src_point_lonlat(g, i, row, plon, plat);
        // the i = 0 entry adds the longitude for BOTH poles (ez_calcspolarwind.c:108-113 subtracts it only for i >= 1)
        float wd = (i == 0 || !south) ? wd0 + plon : wd0 - plon;
        float sp = spd0;
        gdwfllw1('A', 0.f, plon, sp, wd);
        pu[i] = sp;
        pv[i] = wd;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// Cached rotation.  For (u, v) -> (u, v) the chain above is, at every target point, a rotation by an angle that depends
// on the point only: linear in (u, v).  Its matrix is obtained by pushing the unit vectors (1, 0) and (0, 1) through
// k_wind itself (so every grid-type case is covered by construction) and kept per gridset; later calls stream
// 16 B of matrix per point instead of redoing the trigonometry.  Differs from the per-call evaluation by a few 1e-7
// of the wind speed; the reference-order mode (arith 0) does not use it.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_wind_unit(float *__restrict__ u, float *__restrict__ v, long stride, int npts) {
    const int n = blockIdx.x * TPB + threadIdx.x;
    if (n >= npts) return;
    u[n] = 1.f; v[n] = 0.f;
    u[stride + n] = 0.f; v[stride + n] = 1.f;
}
__global__ void __launch_bounds__(TPB) k_wind_pack(const float *__restrict__ u, const float *__restrict__ v, long stride, int npts, float4 *__restrict__ m) {
    const int n = blockIdx.x * TPB + threadIdx.x;
    if (n >= npts) return;
    m[n] = make_float4(u[n], u[stride + n], v[n], v[stride + n]);   // u' = x u + y v ; v' = z u + w v
}
__global__ void __launch_bounds__(TPB) k_wind_apply(const float4 *__restrict__ m, float *__restrict__ u, float *__restrict__ v, long stride, int nfields,
                                                    int npts) {
    const int n = blockIdx.x * TPB + threadIdx.x;
    if (n >= npts) return;
    const float4 r = m[n];
#pragma unroll 4
    for (int f = 0; f < nfields; f++) {
        const long k = (long)f * stride + n;
        const float a = u[k], b = v[k];
        u[k] = fmaf(r.y, b, r.x * a);
        v[k] = fmaf(r.w, b, r.z * a);
    }
}

// the same rotation on a compacted list of target points (the zone points of a direct-kernel launch, ez_interp_fast.cuh)
__global__ void __launch_bounds__(TPB) k_wind_apply_list(const float4 *__restrict__ m, float *__restrict__ u, float *__restrict__ v, long stride,
                                                         int nfields, const int *__restrict__ list, int nlist, int n0, int n1) {
    const int t = blockIdx.x * TPB + threadIdx.x;
    if (t >= nlist) return;
    const int n = list[t];
    if (n < n0 || n >= n1) return;   // output-row band
    const float4 r = m[n];
    for (int f = 0; f < nfields; f++) {
        const long k = (long)f * stride + n;
        const float a = u[k], b = v[k];
        u[k] = fmaf(r.y, b, r.x * a);
        v[k] = fmaf(r.w, b, r.z * a);
    }
}

namespace ezk {
bool wind_matrix(const WindArgs &a0, float4 *m, cudaStream_t s) {
    float *tmp = nullptr;
    const size_t n = (size_t)a0.npts;
    if (cudaMalloc((void **)&tmp, 4 * n * sizeof(float)) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    WindArgs a = a0;
    a.u = tmp; a.v = tmp + 2 * n; a.stride = (long)n; a.nfields = 2; a.mode = 0;
    const int nb = (a.npts + TPB - 1) / TPB;
    k_wind_unit<<<nb, TPB, 0, s>>>(a.u, a.v, a.stride, a.npts);
    k_wind<<<nb, TPB, 0, s>>>(a);
    k_wind_pack<<<nb, TPB, 0, s>>>(a.u, a.v, a.stride, a.npts, m);
    const bool ok = cudaStreamSynchronize(s) == cudaSuccess;
    cudaFree(tmp);
    return ok;
}
void wind_apply(const float4 *m, float *u, float *v, long stride, int nfields, int npts, cudaStream_t s) {
    if (npts <= 0 || nfields <= 0) return;
    k_wind_apply<<<(npts + TPB - 1) / TPB, TPB, 0, s>>>(m, u, v, stride, nfields, npts);
}
void wind_apply_list(const float4 *m, float *u, float *v, long stride, int nfields, const int *list, int nlist, int n0, int n1, cudaStream_t s) {
    if (nlist <= 0 || nfields <= 0) return;
    k_wind_apply_list<<<(nlist + TPB - 1) / TPB, TPB, 0, s>>>(m, u, v, stride, nfields, list, nlist, n0, n1);
}
void polar_wind_rows(const PolarWindArgs &a, cudaStream_t s) {
    if (a.nfields <= 0) return;
    k_polar_wind<<<2 * a.nfields, 256, 0, s>>>(a);
}
void wind(const WindArgs &a, cudaStream_t s) {
    if (a.npts <= 0 || a.nfields <= 0) return;
    k_wind<<<(a.npts + TPB - 1) / TPB, TPB, 0, s>>>(a);
}
}  // namespace ezk
