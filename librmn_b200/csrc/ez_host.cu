// librmn_b200/csrc/ez_host.cu -- host logic of libezb200.so: grid registry, per-thread session and options,
// the device-resident gridset cache, and the C-ABI entry points declared in include/ezscint_b200.h.
//
// This is the L3 layer of the reference (src/interp/*.c: Grille[][] registry, gset cache, options, dispatch)
// re-designed for a GPU: the registry holds *descriptors*; everything per-point (lat/lon of the target, x/y
// in source index space, zone bits, Yin-Yang panel) is built by kernels once per (source,target) pair and
// kept in HBM; c_ezsint / c_ezuvint then cost one (batched) gather kernel plus, when poles or outside
// points exist, a tiny per-field pre-pass.  There is no CPU fallback: if no CUDA device is usable the
// compute entry points print the CUDA error and return -1.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/ezscint_b200.h"
#include "ez_hostmath.h"
#include "ez_kernels.h"

namespace {

// ---------------------------------------------------------------------------------------------------
// data model
// ---------------------------------------------------------------------------------------------------
struct Options {   // _groptions, ez_def.h:225-243 ; defaults ez_statics.c:17
    int degre_interp = EZ_CUBIC;
    int degre_extrap = EZ_MAXIMUM;
    int use_1subgrid = 0;
    int valeur_1subgrid = -1;
    int verbose = 0;
    int polar_correction = 1;
    int wgt_num = 16;
    int msg_pt_tol = 0;
    int cld_interp_alg = 4;   // DISTANCE
    float msg_gridpt_dist = 0.5f;
    float msg_dist_thresh = 3.0f;
    float valeur_extrap = 0.0f;
};

struct TilePlan {   // classification of the target points for the tiled kernel (ez_interp_tiled.cuh), per degree and zone setting
    bool built = false, suitable = false;
    TileDesc *d_tiles = nullptr;
    uint32_t *d_flags = nullptr;
    int *d_list = nullptr;
    int *d_wintiles = nullptr;        // compacted list of the window kernel's tiles
    unsigned int *d_rowy = nullptr;   // row-separable tiles: y of each target row (ez_interp_rowsep.cuh)
    int nlist = 0, ntiles_fast = 0, cap = 0;
    int ntiles_rowsep = 0, vw = 0, cap_rs = 0;
    // constant extrapolation value (extrap_degree = maximum / minimum / value): d_list split into the points outside the source
    // (plain stores, d_clist) and the rest (d_list2); built on the first launch that wants it
    int *d_list2 = nullptr, *d_clist = nullptr;
    int nlist2 = 0, nclist = 0;
    bool const_tried = false, const_ok = false;
    // plan cache (ez_interp_tiled.cuh: k_plan_cache): stencil origin, fp32 weights and s of every target point
    int *d_pc_idx = nullptr;
    float4 *d_pc_wx = nullptr, *d_pc_wy = nullptr;
    float *d_pc_sm1 = nullptr;
    bool pc_tried = false;
};

struct Gset {   // _gridset, ez_def.h:148-166, device-resident
    TilePlan plans[4];          // [cubic ? 2 : 0) + use_zones]
    int gdin = -1;
    bool built = false;
    bool yy = false;            // source is a 'U' supergrid: zb carries the panel bit, no zones
    float *d_x = nullptr, *d_y = nullptr;
    uint8_t *d_zb = nullptr;
    float4 *d_wrot = nullptr;   // c_ezuvint: per-point rotation matrix of the (u,v) -> (u,v) wind conversion, built on first use
    bool wrot_tried = false;
    int *d_perm = nullptr;      // 'Y' cloud targets: the points sorted by source cell (gather locality), built on first use
    bool perm_tried = false;
    // direct kernel (calls of a few fields): the points with zone bits, compacted; [1] leaves out the points that are outside the
    // source and nothing else (a constant extrapolation value is stored by the direct kernel itself)
    int *d_zlist[2] = {nullptr, nullptr};
    int nzlist[2] = {0, 0};
    bool zlist_built[2] = {false, false};
    // direct kernel, cubic on an irregular source: per-point plans (k_plan_cache), built at the second direct call on the gridset
    int *d_dc_idx = nullptr;
    float4 *d_dc_wx = nullptr, *d_dc_wy = nullptr;
    float *d_dc_sm1 = nullptr;
    bool dc_tried = false;
    int dc_calls = 0;
    int counts[6] = {0, 0, 0, 0, 0, 0};   // DEHORS, AU_NORD, AU_SUD, POLE_NORD, POLE_SUD, yang
    // 'Y' cloud SOURCE (ez_cloud.cu; the reference's _ygrid, ez_def.h:136-146): k nearest source points and weights per target point
    bool cloud = false;
    int cloud_nw = 0;
    int *d_cidx = nullptr, *d_cmask = nullptr;
    float *d_cwts = nullptr;
};

struct Grid {   // _Grille, ez_def.h:196-222
    int id = -1;
    bool alive = true;
    int access_count = 1;
    char grtyp = 0, grref = 0;
    int ni = 0, nj = 0;
    int ig[4] = {0, 0, 0, 0}, igref[4] = {0, 0, 0, 0};
    float xg[4] = {0, 0, 0, 0}, xgref[4] = {0, 0, 0, 0};
    float llp[4] = {0, 0, 0, 0}, xyp[4] = {0, 0, 0, 0};
    float r[9] = {0}, ri[9] = {0};
    int i1 = 1, i2 = 0, j1 = 1, j2 = 0, extension = 0, needs_expansion = 0, axe_y_inverse = 0, hemisphere = 0;
    std::vector<float> ax, ay, ncx, ncy;   // host copies (the library owns them, ez_defaxes.c:49-53)
    int nsubgrids = 0;
    int subgrid[2] = {-1, -1};
    int mymaskgrid = -1, mi0 = 0, mi1 = 0, mj0 = 0, mj1 = 0;
    std::vector<int> mask;
    bool has_mask = false;
    // device side: one replica per CUDA device that has used the grid (the registry and the descriptors above are shared; a call
    // works on the replica of the calling thread's current device -- one process per GPU sees exactly one, a process that fans
    // a multi-level call out over the GPUs of a box (c_ezb200_set_devices) builds one per device, deterministically identical)
    struct DevState {
        bool ready = false;   // axes and Newton tables uploaded
        float *d_ax = nullptr, *d_ay = nullptr, *d_cx = nullptr, *d_cy = nullptr;
        float *d_lat = nullptr, *d_lon = nullptr;   // flag LAT
        std::map<int, std::unique_ptr<Gset>> gsets; // keyed by gdin (this grid is the target)
    };
    static constexpr int MAX_DEV = 16;
    DevState dv[MAX_DEV];
    DevState &ds();
    const DevState &ds() const;
};

// One mutex guards the registry, the gridsets and everything built into them (tile plans, plan caches, rotations).  The
// reference leaves its tables unguarded and asks callers to define grids serially and then call c_ezdefset / c_ezsint from many
// threads with their own __thread session (ez_statics.c:12-17).  Here an entry point holds the mutex while it
// looks grids up, builds what is missing and ENQUEUES its kernels on the calling thread's stream; it lets go of it for
// everything that blocks on the device or on the host link -- host<->device copies of caller arrays and the final stream
// synchronisation -- so that the per-thread streams of concurrent callers overlap (Unlocked below).
std::recursive_mutex g_mtx;
thread_local int t_lock_depth = 0;
thread_local int t_dev = 0;   // the calling thread's current CUDA device, refreshed whenever an entry point takes the lock
Grid::DevState &Grid::ds() { return dv[t_dev]; }
const Grid::DevState &Grid::ds() const { return dv[t_dev]; }
struct Lock {
    Lock() {
        g_mtx.lock();
        if (t_lock_depth++ == 0) {
            int d = 0;
            if (cudaGetDevice(&d) != cudaSuccess) { (void)cudaGetLastError(); d = 0; }
            t_dev = (d >= 0 && d < Grid::MAX_DEV) ? d : 0;
        }
    }
    ~Lock() { t_lock_depth--; g_mtx.unlock(); }
    Lock(const Lock &) = delete;
    Lock &operator=(const Lock &) = delete;
};
struct Unlocked {   // releases every level this thread holds for the duration of a blocking call, then takes them back
    int depth;
    Unlocked() : depth(t_lock_depth) {
        for (int k = 0; k < depth; k++) g_mtx.unlock();
        t_lock_depth = 0;
    }
    ~Unlocked() {
        for (int k = 0; k < depth; k++) g_mtx.lock();
        t_lock_depth = depth;
    }
    Unlocked(const Unlocked &) = delete;
    Unlocked &operator=(const Unlocked &) = delete;
};
std::vector<std::unique_ptr<Grid>> g_grids;

// Per-thread device resources are released when a caller's worker thread ends (the CUDA runtime is alive then).  The main
// thread's are left to process teardown: its thread_local destructors run inside exit(), when the runtime may be gone.
static const std::thread::id g_main_thread = std::this_thread::get_id();
inline bool thread_may_release() { return std::this_thread::get_id() != g_main_thread; }

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p && thread_may_release()) { cudaFree(p); (void)cudaGetLastError(); } }
    void *get(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFree(p);
            p = nullptr;
            cap = 0;
            // rounded up (+32 B, 256-byte granular): fields staged here end inside the allocation even after the tile kernels'
            // 16-byte block rounding (source_blocks_inside_allocation), whatever the field length mod 4
            const size_t want = (bytes + 32 + 255) & ~(size_t)255;
            if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; (void)cudaGetLastError(); return nullptr; }
            cap = want;
        }
        return p;
    }
};

inline int env_devices() {
    const char *e = getenv("EZB200_DEVICES");
    const int n = e ? atoi(e) : 1;
    return n < 1 ? 1 : n;
}

struct Session {   // the reference's __thread statics (ez_statics.c:12-17) plus per-thread CUDA state
    int gdin = -1, gdout = -1;
    Options opt;
    cudaStream_t stream = 0;    // the stream every launch of this thread goes to (c_ezb200_set_stream; default: the legacy default stream)
    bool stream_user = false;   // set by the caller: never replaced
    cudaStream_t own_stream = nullptr;   // a private non-blocking stream for host-pointer calls of threads that never set one (HostCallStream)
    int devices = env_devices();   // c_ezb200_set_devices (initial value: EZB200_DEVICES, default 1): GPUs a host-pointer multi-field call of this thread is fanned out over
    int direct_max = getenv("EZB200_DIRECT_MAX_FIELDS") ? atoi(getenv("EZB200_DIRECT_MAX_FIELDS")) : 4;   // c_ezb200_set_direct_max_fields
    int call_fields = 0;        // fields of the API call in progress (a call cut into level chunks or device blocks picks its kernels by the whole call's count)
    int row0 = -1, row1 = -1;   // c_ezb200_set_rows: output-row band [row0, row1) of the target (multi-GPU split of one field); -1 = whole target
    int arith = 2;   // 0 reference order, 1 fp64 weight form, 2 fp32 difference form + tiled kernel (default)
    long launches = 0;
    bool timing = false;        // c_ezb200_set_timing: record events around the tile kernels of each interpolation launch
    bool timing_valid = false;
    ezk::InterpTiming tev = {{nullptr, nullptr, nullptr, nullptr}};
    ezk::InterpFork fork = {{nullptr, nullptr}, nullptr, {nullptr, nullptr}};   // created on the first tiled launch
    bool fork_ok = false, fork_tried = false;
    int last_tiles[3] = {0, 0, 0};   // tiles on the row-separable kernel, on the window kernel, active tiles
    int last_path = 0;          // kernels used by the last interpolation launch: 1 per-point, 2 window tile kernel, 4 per-point on the compacted list, 8 row-separable tile kernel
    int last_nlist = 0;
    DevBuf st_in[2], st_out[2], st_fs, st_scratch, st_misc[4], st_xpn;
    // host-pointer pipeline (ezsint_pipelined): two copy streams beside the compute stream, double-buffered chunks
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    DevBuf st_pts[10];          // staging of caller arrays in the scattered-point API
    DevBuf prow[2];             // synthetic pole rows of (u, v) for the polar vector correction (K18)
    ~Session() {
        if (!thread_may_release()) return;
        cudaStream_t streams[] = {own_stream, s_h2d, s_d2h, fork.aux[0], fork.aux[1]};
        for (cudaStream_t st : streams) if (st) cudaStreamDestroy(st);
        cudaEvent_t events[] = {tev.ev[0], tev.ev[1], tev.ev[2], tev.ev[3], fork.ev_fork, fork.ev_join[0], fork.ev_join[1],
                                ev_in[0], ev_in[1], ev_comp[0], ev_comp[1], ev_out[0], ev_out[1]};
        for (cudaEvent_t e : events) if (e) cudaEventDestroy(e);
        (void)cudaGetLastError();
    }
};
thread_local Session S;

#define CUDA_OK(call)                                                                               \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            fprintf(stderr, "<ezb200> CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
            return -1;                                                                              \
        }                                                                                           \
    } while (0)

Grid *grid_of(int gdid) {
    if (gdid < 0 || gdid >= (int)g_grids.size()) return nullptr;
    Grid *g = g_grids[gdid].get();
    return (g && g->alive) ? g : nullptr;
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

template <class T>
int dev_alloc(T **p, size_t n) {
    CUDA_OK(cudaMalloc((void **)p, n * sizeof(T)));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// grid definition (host)
// ---------------------------------------------------------------------------------------------------

// ez_xpncof.c:48-226
void xpncof(Grid &g) {
    g.i1 = 1;
    g.i2 = g.ni;
    const int ni = g.ni, nj = g.nj;
    switch (g.grtyp) {
        case '!': case 'N': case 'S': case 'T':
            g.j1 = 1; g.j2 = nj; g.extension = 0;
            break;
        case 'L': {
            g.extension = 0; g.j1 = 1; g.j2 = nj;
            float x[4];
            ezh::cigaxg('L', x, g.ig);
            float first_lat = x[0], first_lon = x[1], dlat = x[2], dlon = x[3];
            if ((first_lat - dlat) > (-90.0 + 0.01 * dlat)) break;
            float last_lat = first_lat + (nj - 1) * dlat;
            if ((last_lat + dlat) < (90.0 - 0.01 * dlat)) break;
            if (first_lon < 0.0) first_lon += 360.0;
            float last_lon = first_lon + (dlon * (ni - 1));
            if ((last_lon - first_lon) > (360 - 0.01 * dlon)) { g.extension = 1; break; }
            float extra_lon = last_lon + dlon;
            if ((extra_lon - first_lon) > (360.0 - dlon * 0.01)) g.extension = 2;
            break;
        }
        case 'A': case 'G':
            g.extension = 2;
            if (g.ig[0] == 0) { g.j1 = 1; g.j2 = nj; }
            else if (g.ig[0] == 1) { g.j1 = -nj + 1; g.j2 = nj; }
            else if (g.ig[0] == 2) { g.j1 = 1; g.j2 = 2 * nj; }
            break;
        case 'B':
            g.extension = 1;
            if (g.ig[0] == 0) { g.j1 = 1; g.j2 = nj; }
            else if (g.ig[0] == 1) { g.j1 = -nj + 2; g.j2 = nj; }
            else if (g.ig[0] == 2) { g.j1 = 1; g.j2 = 2 * nj - 1; }
            break;
        case 'E':
            g.j1 = 1; g.j2 = nj;   // `extension` is never assigned for 'E' (ez_xpncof.c:152-155): stays 0
            break;
        case '#': case 'Z':
            if (g.ax.empty() || g.ay.empty()) { g.j1 = 1; g.j2 = nj; g.extension = 0; break; }
            switch (g.grref) {
                case 'E':
                    g.j1 = 1; g.j2 = nj;
                    g.extension = ((g.ax[ni - 1] - g.ax[0]) < 359.0) ? 0 : 1;
                    break;
                case 'L': {
                    g.extension = 0; g.j1 = 1; g.j2 = nj;
                    float first_lat = g.ay[0], last_lat = g.ay[nj - 1];
                    float dlat = g.ay[1] - g.ay[0];
                    if ((first_lat - dlat) > (-90.0 + 0.01 * dlat)) break;
                    dlat = g.ay[nj - 1] - g.ay[nj - 2];
                    if ((last_lat + dlat) < (90.0 - 0.01 * dlat)) break;
                    float first_lon = g.ax[0], last_lon = g.ax[ni - 1];
                    float dlon = g.ax[ni - 1] - g.ax[ni - 2];
                    if ((last_lon - first_lon) > (360.0 - 0.01 * dlon)) {
                        g.extension = 1;
                    } else {
                        float extra_lon = last_lon + dlon;
                        if ((extra_lon - first_lon) > (360.0 - 0.01 * dlon)) g.extension = 2;
                    }
                    break;
                }
                default:
                    g.j1 = 1; g.j2 = nj; g.extension = 0;
                    break;
            }
            break;
        default:
            g.j1 = 1; g.j2 = nj; g.extension = 0;
            break;
    }
}

// c_ezdefxg, ez_defxg.c:28-165 (+ the fp32 parameter sets ez_ll2rgd.inc:70-118 derives per call)
void defxg(Grid &g) {
    const int ni = g.ni, nj = g.nj;
    switch (g.grtyp) {
        case 'A': case 'G':
            g.xg[3] = 360. / ni;   // DLON (double division, stored fp32)
            g.xg[1] = 0.0;         // SWLON
            switch (g.ig[0]) {
                case 0: g.xg[2] = 180. / nj; g.xg[0] = -90. + 0.5 * g.xg[2]; break;
                case 1: g.xg[2] = 90. / nj; g.xg[0] = 0.5 * g.xg[2]; g.needs_expansion = 1; break;
                case 2: g.xg[2] = 90. / nj; g.xg[0] = -90. + 0.5 * g.xg[2]; g.needs_expansion = 1; break;
                default: fprintf(stderr, "<ez_gdef_fmem> 'A' grid has to be Global/North/South\n"); break;
            }
            if (g.ig[1] == 1) g.axe_y_inverse = 1;
            break;
        case 'B':
            g.xg[3] = 360. / (ni - 1);
            g.xg[1] = 0.0;
            switch (g.ig[0]) {
                case 0: g.xg[2] = 180. / (nj - 1); g.xg[0] = -90.; break;
                case 1: g.xg[2] = 90. / (nj - 1); g.xg[0] = 0.; g.needs_expansion = 1; break;
                case 2: g.xg[2] = 90. / (nj - 1); g.xg[0] = -90.; g.needs_expansion = 1; break;
                default: fprintf(stderr, "<ezgdef_fmem> 'B' grid has to be Global/North/South\n"); break;
            }
            if (g.ig[1] == 1) g.axe_y_inverse = 1;
            break;
        case 'E':
            ezh::cigaxg('E', g.xg, g.ig);
            break;
        case '#': case 'Z':
            if (g.grref == 'N') g.hemisphere = 1;
            if (g.grref == 'S') g.hemisphere = 2;
            if (g.grref == 'E') ezh::cigaxg('E', g.xgref, g.igref);
            break;
        case 'L':
            ezh::cigaxg('L', g.xg, g.ig);
            break;
        case 'N':
            ezh::cigaxg('N', g.xg, g.ig);
            g.hemisphere = 1;
            break;
        case 'S':
            ezh::cigaxg('S', g.xg, g.ig);
            g.hemisphere = 2;
            break;
        default:
            break;
    }
}

// parameters the kernels need beyond xg/xgref
void derive_params(Grid &g) {
    const int ni = g.ni, nj = g.nj;
    // ez_calclatlon's grll arguments
    if (g.grtyp == 'A' || g.grtyp == 'B' || g.grtyp == 'L') {
        for (int k = 0; k < 4; k++) g.llp[k] = g.xg[k];
    } else if (g.grtyp == 'E') {
        float dlon = 360. / (ni - 1);            // ez_calclatlon.c:64-67
        float dlat = 180. / (nj);
        g.llp[0] = -90. + 0.5 * dlat;
        g.llp[1] = 0.0f;
        g.llp[2] = dlat;
        g.llp[3] = dlon;
    }
    // ez_ll2rgd's fp32 (xlat0, xlon0, dellat, dellon)
    if (g.grtyp == 'L') {
        for (int k = 0; k < 4; k++) g.xyp[k] = g.xg[k];
    } else if (g.grtyp == 'A') {
        float dellon = 360.0f / (float)ni, dellat = 0.f, xlat0 = 0.f;
        if (g.ig[0] == 0) { dellat = 180.0f / (float)nj; xlat0 = -90.0f + dellat * 0.5f; }
        if (g.ig[0] == 1) { dellat = 90.0f / (float)nj; xlat0 = dellat * 0.5f; }
        if (g.ig[0] == 2) { dellat = 90.0f / (float)nj; xlat0 = -90.0f + dellat * 0.5f; }
        g.xyp[0] = xlat0; g.xyp[1] = 0.0f; g.xyp[2] = dellat; g.xyp[3] = dellon;
    } else if (g.grtyp == 'B') {
        float dellon = 360.0f / (float)(ni - 1), dellat = 0.f, xlat0 = 0.f;
        if (g.ig[0] == 0) { dellat = 180.0f / (float)(nj - 1); xlat0 = -90.0f; }
        if (g.ig[0] == 1) { dellat = 90.0f / (float)(nj - 1); xlat0 = 0.0f; }
        if (g.ig[0] == 2) { dellat = 90.0f / (float)(nj - 1); xlat0 = -90.0f; }
        g.xyp[0] = xlat0; g.xyp[1] = 0.0f; g.xyp[2] = dellat; g.xyp[3] = dellon;
    }
    // rotation matrices: ez_crot(r, ri, xlon1, xlat1, xlon2, xlat2) with (xlat1,xlon1,xlat2,xlon2) = xg / xgref
    if (g.grtyp == 'E') ezh::crot(g.r, g.ri, g.xg[1], g.xg[0], g.xg[3], g.xg[2]);
    if ((g.grtyp == 'Z' || g.grtyp == '#') && g.grref == 'E') ezh::crot(g.r, g.ri, g.xgref[1], g.xgref[0], g.xgref[3], g.xgref[2]);
}

int calcntncof(Grid &g) {   // ez_calcntncof.c:29-54
    if (g.grtyp == 'Y' || !g.ncx.empty()) return 0;
    int nnj = g.j2 - g.j1 + 1;
    g.ncx.assign((size_t)g.ni * 6, 1.0f);
    g.ncy.assign((size_t)nnj * 6, 1.0f);
    if (!ezh::nwtncof(g.ncx.data(), g.ncy.data(), g.ax.data(), g.ay.data(), g.ni, g.nj, g.j1, g.j2, g.extension)) return -1;
    return 0;
}

// c_ezdefaxes, ez_defaxes.c:31-114
int defaxes(Grid &g, const float *ax, const float *ay) {
    switch (g.grtyp) {
        case '#': case 'Z':
            ezh::cigaxg(g.grref, g.xgref, g.igref);
            g.ax.assign(ax, ax + g.ni);
            g.ay.assign(ay, ay + g.nj);
            xpncof(g);
            if (calcntncof(g) < 0) return -1;
            break;
        case 'Y':
            g.ax.assign(ax, ax + (size_t)g.ni * g.nj);
            g.ay.assign(ay, ay + (size_t)g.ni * g.nj);
            xpncof(g);
            break;
        case 'G': {
            g.grref = 'L';
            g.xgref[0] = 0.0f; g.xgref[1] = 0.0f; g.xgref[2] = 1.0f; g.xgref[3] = 1.0f;
            ezh::cxgaig('L', g.igref, g.xgref);
            g.ax.resize(g.ni);
            float dlon = 360. / (float)g.ni;
            for (int i = 0; i < g.ni; i++) g.ax[i] = (float)i * dlon;
            xpncof(g);
            int nlat = (g.ig[0] == 0) ? g.nj : 2 * g.nj;
            g.ay.resize(nlat);
            ezh::glat(g.ay.data(), nlat);
            if (calcntncof(g) < 0) return -1;
            break;
        }
        default:
            xpncof(g);
            break;
    }
    return 0;
}

bool same_axes(const std::vector<float> &a, const float *b, size_t n) {
    if (a.size() < n) return false;
    for (size_t i = 0; i < n; i++)
        if (a[i] != b[i]) return false;
    return true;
}

// registry lookup with the reference's identity rules (ez_findgrid.c:49-130)
int find_grid(char grtyp, char grref, int ni, int nj, const int ig[4], const int igref[4], const float *ax, const float *ay, bool irregular) {
    for (auto &up : g_grids) {
        Grid *g = up.get();
        if (!g || !g->alive || g->nsubgrids > 0) continue;
        if (g->grtyp != grtyp || g->ni != ni || g->nj != nj) continue;
        if (memcmp(g->ig, ig, sizeof(int) * 4) != 0) continue;
        if (grtyp == 'G') return g->id;
        if (!irregular) return g->id;
        if (g->grref != grref || memcmp(g->igref, igref, sizeof(int) * 4) != 0) continue;
        size_t nx = (grtyp == 'Y') ? (size_t)ni * nj : (size_t)ni, ny = (grtyp == 'Y') ? (size_t)ni * nj : (size_t)nj;
        if (ax && ay && same_axes(g->ax, ax, nx) && same_axes(g->ay, ay, ny)) return g->id;
    }
    return -1;
}

int add_grid(std::unique_ptr<Grid> g) {
    g->id = (int)g_grids.size();
    g_grids.push_back(std::move(g));
    return g_grids.back()->id;
}

// keys: what c_ezgprm reports as ig1..ig4 when it differs from the reference grid's descriptors -- the ip1, ip2, ip3 (and the '#'
// window origin) of the positional records a file-backed definition was read from (ez_lireEnrPositionnels.c:289-316)
int gdef_fmem(int ni, int nj, const char *grtyp, const char *grref, int ig1, int ig2, int ig3, int ig4, const float *ax, const float *ay,
              const int *keys = nullptr) {
    Lock lk;
    const char t = grtyp[0];
    const int ig[4] = {ig1, ig2, ig3, ig4};
    if (ni <= 0 || nj <= 0) {
        fprintf(stderr, "<c_ezgdef_fmem> invalid dimensions %d x %d\n", ni, nj);
        return -1;
    }
    const bool irregular = (t == '#' || t == 'Y' || t == 'Z' || t == 'G');
    if (irregular) {
        if (t != 'G' && (!ax || !ay)) {
            fprintf(stderr, "<c_ezgdef_fmem> grid type '%c' needs ax and ay\n", t);
            return -1;
        }
        const char rr = grref ? grref[0] : ' ';
        // 'G' grids are matched on (ni,nj,ig) alone (ez_findgrid.c:53-57)
        int found = find_grid(t, (t == 'G') ? 'L' : rr, ni, nj, keys ? keys : ig, ig, ax, ay, t != 'G');
        if (found >= 0) {
            g_grids[found]->access_count++;
            return found;
        }
        auto g = std::make_unique<Grid>();
        g->grtyp = t;
        g->grref = rr;
        g->ni = ni;
        g->nj = nj;
        for (int k = 0; k < 4; k++) { g->ig[k] = keys ? keys[k] : ig[k]; g->igref[k] = ig[k]; }
        g->i1 = 1; g->i2 = ni; g->j1 = 1; g->j2 = nj;
        defxg(*g);
        if (defaxes(*g, ax, ay) < 0) return -1;
        xpncof(*g);
        derive_params(*g);
        return add_grid(std::move(g));
    }
    if (!(t == 'A' || t == 'B' || t == 'E' || t == 'L' || t == 'N' || t == 'S')) {
        fprintf(stderr, "<c_ezgdef_fmem> grid type '%c' is outside the B200 hot-path scope (L/N/S/E/A/B/G/Z/#/Y/U)\n", t);
        return -1;
    }
    const int zero[4] = {0, 0, 0, 0};
    int found = find_grid(t, 0, ni, nj, ig, zero, nullptr, nullptr, false);
    if (found >= 0) {
        g_grids[found]->access_count++;
        return found;
    }
    auto g = std::make_unique<Grid>();
    g->grtyp = t;
    g->grref = 0;   // c_ezidentify_reg_grid, ez_identifygrid.c:63
    g->ni = ni;
    g->nj = nj;
    for (int k = 0; k < 4; k++) g->ig[k] = ig[k];
    defxg(*g);
    xpncof(*g);
    derive_params(*g);
    return add_grid(std::move(g));
}

// c_ezgdef_yymask, ezgdef_yymask.c:27-69
void gdef_yymask(Grid &sub) {
    int k = 0, i0 = 0, i1 = 0;
    for (int i = 0; i < sub.ni; i++)
        if (sub.ax[i] >= 45.0 && sub.ax[i] <= 315.0) { k++; if (k == 1) i0 = i; i1 = i; }
    int yni = k, j0 = 0, j1 = 0;
    k = 0;
    for (int j = 0; j < sub.nj; j++)
        if (sub.ay[j] >= -45.0 && sub.ay[j] <= 45.0) { k++; if (k == 1) j0 = j; j1 = j; }
    int ynj = k;
    char t[2] = {sub.grtyp, 0}, r[2] = {sub.grref, 0};
    std::vector<float> mx(sub.ax.begin() + i0, sub.ax.begin() + i0 + yni), my(sub.ay.begin() + j0, sub.ay.begin() + j0 + ynj);
    int sid = sub.id;
    int mid = gdef_fmem(yni, ynj, t, r, sub.igref[0], sub.igref[1], sub.igref[2], sub.igref[3], mx.data(), my.data());
    Grid *s = grid_of(sid);   // g_grids may have been reallocated
    s->mymaskgrid = mid;
    s->mi0 = i0; s->mi1 = i1; s->mj0 = j0; s->mj1 = j1;
}

// ---------------------------------------------------------------------------------------------------
// device mirrors
// ---------------------------------------------------------------------------------------------------
int ensure_device(Grid &g) {
    Grid::DevState &d = g.ds();
    if (d.ready) return 0;
    auto up = [&](float **dp, const std::vector<float> &h) -> int {
        if (h.empty()) return 0;
        CUDA_OK(cudaMalloc((void **)dp, h.size() * sizeof(float)));
        CUDA_OK(cudaMemcpy(*dp, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    };
    if (up(&d.d_ax, g.ax) < 0 || up(&d.d_ay, g.ay) < 0 || up(&d.d_cx, g.ncx) < 0 || up(&d.d_cy, g.ncy) < 0) return -1;
    d.ready = true;
    return 0;
}

GridDev to_dev(const Grid &g) {
    GridDev d;
    memset(&d, 0, sizeof(d));
    d.grtyp = g.grtyp; d.grref = g.grref;
    d.ni = g.ni; d.nj = g.nj; d.i1 = g.i1; d.i2 = g.i2; d.j1 = g.j1; d.j2 = g.j2;
    d.ext = g.extension;
    d.xpn = XPN_NONE;
    if (g.needs_expansion) {
        if (g.grtyp == 'A' || g.grtyp == 'G') d.xpn = (g.ig[0] == 1) ? XPN_AG_NORTH : XPN_AG_SOUTH;
        if (g.grtyp == 'B') d.xpn = (g.ig[0] == 1) ? XPN_B_NORTH : XPN_B_SOUTH;
    }
    d.yinv = g.axe_y_inverse;
    d.hemisphere = g.hemisphere;
    for (int k = 0; k < 4; k++) {
        d.ig[k] = g.ig[k]; d.igref[k] = g.igref[k]; d.xg[k] = g.xg[k]; d.xgref[k] = g.xgref[k]; d.llp[k] = g.llp[k]; d.xyp[k] = g.xyp[k];
    }
    for (int k = 0; k < 9; k++) { d.r[k] = g.r[k]; d.ri[k] = g.ri[k]; }
    d.ax = g.ds().d_ax; d.ay = g.ds().d_ay; d.cx = g.ds().d_cx; d.cy = g.ds().d_cy;
    return d;
}

// ez_calclatlon, ez_calclatlon.c:30 (flag LAT): device-resident lat/lon of every grid point
int ensure_latlon(Grid &g) {
    if (g.ds().d_lat) return 0;
    if (g.grtyp == 'Y' && (g.grref == 'N' || g.grref == 'S' || g.grref == 'E')) {
        fprintf(stderr, "<ez_calclatlon> Operation not supported - Y grid on %c Grid\n", g.grref);
        return -1;
    }
    if (ensure_device(g) < 0) return -1;
    size_t n = (size_t)g.ni * g.nj;
    if (dev_alloc(&g.ds().d_lat, n) < 0) { g.ds().d_lat = nullptr; return -1; }
    if (dev_alloc(&g.ds().d_lon, n) < 0) { cudaFree(g.ds().d_lat); g.ds().d_lat = g.ds().d_lon = nullptr; return -1; }
    ezk::calclatlon(to_dev(g), g.ds().d_lat, g.ds().d_lon, (int)n, S.stream);
    S.launches++;
    // cached state shared by every session: complete before it is visible to other threads' streams
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(S.stream);
    if (e != cudaSuccess) {
        fprintf(stderr, "<ezb200> lat/lon build failed: %s\n", cudaGetErrorString(e));
        cudaFree(g.ds().d_lat); cudaFree(g.ds().d_lon);
        g.ds().d_lat = g.ds().d_lon = nullptr;
        return -1;
    }
    return 0;
}

// ez_defzones.c:44-81
int zone_extrap(const Grid &s) {
    int extrap = 0;
    bool through = false;
    switch (s.grtyp) {
        case 'N': case 'S': case '!': return 1;
        case 'L':
            extrap = (s.extension == 0) ? 1 : 0;
            through = true;   // the reference's `case 'L'` has no break
        // fallthrough
        case '#': case 'Z': case 'Y':
            (void)through;
            switch (s.grref) {
                case 'N': case 'S': extrap = 1; break;
                case 'E': case 'L':
                    if (!s.ax.empty() && 358.0 > (s.ax[s.ni - 1] - s.ax[0])) extrap = 1;
                    break;
                default: break;
            }
            break;
        default: break;
    }
    return extrap;
}

// one (lat,lon) -> (x,y) on the device, c_gdxyfll_orig semantics; used for the pole coordinates of the zones
int point_xy(Grid &s, float lat, float lon, float *x, float *y) {
    float *d = (float *)S.st_misc[0].get(4 * sizeof(float));
    if (!d) return -1;
    float h[4] = {lat, lon, 0.f, 0.f};
    CUDA_OK(cudaMemcpyAsync(d, h, 2 * sizeof(float), cudaMemcpyHostToDevice, S.stream));
    ezk::ll2xy_points(to_dev(s), s.j2, d, d + 1, d + 2, d + 3, 1, S.stream);
    S.launches++;
    CUDA_OK(cudaMemcpyAsync(h, d, 4 * sizeof(float), cudaMemcpyDeviceToHost, S.stream));
    CUDA_OK(cudaStreamSynchronize(S.stream));
    *x = h[2];
    *y = h[3];
    return 0;
}

void free_gset(Gset &gs);

// The gridset build: ez_calcxy (ez_calcxy.c:28) + ez_defzones (ez_defzones.c:25) for a plain source,
// c_ezyy_calcxy (ezyy_calcxy.c:28) for a 'U' source.  Runs once per (gdout, gdin); results stay in HBM.
Gset *ensure_gset(Grid &out, Grid &in) {
    auto &slot = out.ds().gsets[in.id];
    if (!slot) {
        slot = std::make_unique<Gset>();
        slot->gdin = in.id;
    }
    Gset *gs = slot.get();
    if (gs->built) return gs;
    if (ensure_latlon(out) < 0) return nullptr;
    size_t n = (size_t)out.ni * out.nj;
    if (cudaMalloc((void **)&gs->d_x, n * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&gs->d_y, n * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void **)&gs->d_zb, n) != cudaSuccess) {
        fprintf(stderr, "<ezb200> cudaMalloc failed for the gridset (%zu points): %s\n", n, cudaGetErrorString(cudaGetLastError()));
        free_gset(*gs);
        return nullptr;
    }
    struct Undo {   // a failed build leaves no half-built gridset behind (and leaks nothing on retry)
        Gset *g;
        ~Undo() { if (g && !g->built) free_gset(*g); }
    } undo{gs};
    unsigned int *d_counts = (unsigned int *)S.st_misc[1].get(8 * sizeof(unsigned int));
    if (!d_counts) return nullptr;
    cudaMemsetAsync(d_counts, 0, 8 * sizeof(unsigned int), S.stream);
    if (in.nsubgrids > 0) {
        Grid *yin = grid_of(in.subgrid[0]), *yan = grid_of(in.subgrid[1]);
        Grid *mk = yin ? grid_of(yin->mymaskgrid) : nullptr;
        if (!yin || !yan || !mk) return nullptr;
        if (ensure_device(*yin) < 0 || ensure_device(*yan) < 0 || ensure_device(*mk) < 0) return nullptr;
        // the mask is produced under the caller's current polar_correction (ezyymint.c:56-64)
        ezk::yy_calcxy(to_dev(*mk), to_dev(*yin), to_dev(*yan), S.opt.polar_correction == 1, out.ds().d_lat, out.ds().d_lon, gs->d_x, gs->d_y, gs->d_zb,
                       d_counts, (int)n, S.stream);
        S.launches++;
        gs->yy = true;
    } else {
        if (in.grtyp == 'Y') {
            // ez_calcxy.c:99-128: no x, y for a cloud source (ez_calcxy_y never writes them): nearest points + weights instead
            if (in.has_mask) {
                fprintf(stderr, "<ezb200> masked 'Y' cloud sources (ez_calcxy_y_m) are not supported\n");
                return nullptr;
            }
            if (ensure_latlon(in) < 0) return nullptr;
            const int nw = S.opt.wgt_num;
            if (nw < 1 || nw > 32) { fprintf(stderr, "<ezb200> weight_number %d is outside 1..32\n", nw); return nullptr; }
            const size_t nsrc = (size_t)in.ni * in.nj;
            if (cudaMalloc((void **)&gs->d_cidx, n * nw * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&gs->d_cwts, n * nw * sizeof(float)) != cudaSuccess ||
                cudaMalloc((void **)&gs->d_cmask, n * sizeof(int)) != cudaSuccess) {
                fprintf(stderr, "<ezb200> cudaMalloc failed for the cloud weights (%zu points x %d): %s\n", n, nw, cudaGetErrorString(cudaGetLastError()));
                return nullptr;
            }
            cudaMemsetAsync(gs->d_x, 0, n * sizeof(float), S.stream);
            cudaMemsetAsync(gs->d_y, 0, n * sizeof(float), S.stream);
            cudaMemsetAsync(gs->d_zb, 0, n, S.stream);
            if (!ezk::cloud_build(in.ds().d_lat, in.ds().d_lon, (int)nsrc, out.ds().d_lat, out.ds().d_lon, (int)n, nw, gs->d_cidx, gs->d_cwts, gs->d_cmask, S.stream)) {
                fprintf(stderr, "<ezb200> cloud gridset build failed: %s\n", cudaGetErrorString(cudaGetLastError()));
                return nullptr;
            }
            S.launches += 8;
            gs->cloud = true;
            gs->cloud_nw = nw;
            gs->built = true;
            return gs;
        }
        if (ensure_device(in) < 0) return nullptr;
        int extrap = zone_extrap(in);
        float ypn = 0.f, yps = 0.f;
        if (!extrap) {
            if (in.grtyp == 'Z' && in.grref == 'E') {
                ypn = in.nj + 0.5;   // ez_defzone_polenord.c:46-48, ez_defzone_polesud.c:47-49
                yps = 0.5;
            } else {
                float xp;
                if (point_xy(in, 90.0f, 0.0f, &xp, &ypn) < 0) return nullptr;
                if (point_xy(in, -90.0f, 0.0f, &xp, &yps) < 0) return nullptr;
            }
        }
        // lon is mutated in place by the A/B/L and Z-on-L coordinate routes
        int wb = (in.grtyp == 'A' || in.grtyp == 'B' || in.grtyp == 'L' || ((in.grtyp == 'Z' || in.grtyp == '#' || in.grtyp == 'G') && in.grref == 'L'));
        ezk::calcxy(to_dev(in), in.nj, out.ds().d_lat, out.ds().d_lon, wb, gs->d_x, gs->d_y, gs->d_zb, 1, extrap, ypn, yps, d_counts, (int)n, S.stream);
        S.launches++;
    }
    unsigned int hc[8];
    if (cudaMemcpyAsync(hc, d_counts, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost, S.stream) != cudaSuccess ||
        cudaStreamSynchronize(S.stream) != cudaSuccess) {
        fprintf(stderr, "<ezb200> gridset build failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    for (int k = 0; k < 6; k++) gs->counts[k] = (int)hc[k];
    gs->built = true;
    return gs;
}

// The tile kernels stage source rows in whole 16-byte blocks: up to 12 bytes before the first and after the last
// element of the source fields are read.  That is inside the caller's allocation for every CUDA allocator (256-byte
// granularity) unless the fields are a tight sub-range at the very edge of it; ask the driver for the allocation's
// range and keep the per-point kernel when the 16-byte blocks would leave it (or when the driver cannot tell).
bool source_blocks_inside_allocation(const float *d_zin, size_t nbytes) {
    typedef int (*get_range_fn)(unsigned long long *, size_t *, unsigned long long);
    static get_range_fn fn = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (get_range_fn)p;
        else
            (void)cudaGetLastError();
    }
    if (!fn) return false;
    unsigned long long base = 0;
    size_t size = 0;
    const unsigned long long lo = (unsigned long long)(uintptr_t)d_zin & ~15ull;
    const unsigned long long hi = ((unsigned long long)(uintptr_t)d_zin + nbytes + 15ull) & ~15ull;
    if (fn(&base, &size, (unsigned long long)(uintptr_t)d_zin) != 0) return false;
    return lo >= base && hi <= base + size;
}

// Tile plan for the current (degree, zones) of `a` on this gridset; built on first use, kept in HBM.
TilePlan *ensure_tileplan(Gset &gs, const InterpArgs &a) {
    TilePlan &tp = gs.plans[(a.o.degree == EZ_CUBIC ? 2 : 0) + (a.use_zones ? 1 : 0)];
    if (tp.built) return &tp;
    int ntx, nty;
    ezk::tile_dims(a.out_ni, a.out_nj, &ntx, &nty);
    const size_t ntiles = (size_t)ntx * nty;
    unsigned int *d_counters = (unsigned int *)S.st_misc[3].get(8 * sizeof(unsigned int));
    if (!d_counters) return nullptr;
    auto free_plan = [&]() {
        cudaFree(tp.d_tiles); cudaFree(tp.d_flags); cudaFree(tp.d_rowy); cudaFree(tp.d_list); cudaFree(tp.d_wintiles);
        tp.d_tiles = nullptr; tp.d_flags = nullptr; tp.d_rowy = nullptr; tp.d_list = nullptr; tp.d_wintiles = nullptr;
    };
    if (cudaMalloc((void **)&tp.d_tiles, ntiles * sizeof(TileDesc)) != cudaSuccess ||
        cudaMalloc((void **)&tp.d_flags, ntiles * 8 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc((void **)&tp.d_rowy, ntiles * 32 * sizeof(unsigned int)) != cudaSuccess) {
        fprintf(stderr, "<ezb200> cudaMalloc failed for the tile plan: %s\n", cudaGetErrorString(cudaGetLastError()));
        free_plan();
        return nullptr;
    }
    cudaMemsetAsync(d_counters, 0, 8 * sizeof(unsigned int), S.stream);
    InterpArgs ca = a;
    static const bool rowsep_on = !(getenv("EZB200_ROWSEP") && atoi(getenv("EZB200_ROWSEP")) == 0);
    ca.rowsep_enable = (rowsep_on && (!a.yy || ((a.src.ni - a.src2.ni) & 3) == 0)) ? 1 : 0;   // one staged pitch for both panels
    ezk::tile_classify(ca, tp.d_tiles, tp.d_flags, tp.d_rowy, d_counters, S.stream);
    S.launches++;
    unsigned int hc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyAsync(hc, d_counters, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost, S.stream) != cudaSuccess ||
        cudaStreamSynchronize(S.stream) != cudaSuccess) {
        fprintf(stderr, "<ezb200> tile classification failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        free_plan();
        return nullptr;
    }
    tp.nlist = (int)hc[0];
    tp.ntiles_fast = (int)hc[1];
    tp.cap = ((int)hc[2] + 3) & ~3;   // floats per staged level: the largest tile box, 16-byte granular
    tp.vw = (int)hc[3];
    tp.ntiles_rowsep = (int)hc[4];
    tp.cap_rs = (int)hc[5];   // tallest row-separable box; ezk::interp turns it into floats per staged level
    // the tile kernel pays off when most points take it (target as fine as the source); otherwise keep the per-point kernel
    tp.suitable = tp.ntiles_fast > 0 && (size_t)tp.nlist * 2 <= (size_t)a.npts;
    if (tp.suitable && tp.ntiles_fast > tp.ntiles_rowsep) {
        // the window kernel's tiles, compacted; both polar caps first (the rows that carry ragged work): 0, last, 1, last-1 ...
        std::vector<TileDesc> ht(ntiles);
        if (cudaMemcpy(ht.data(), tp.d_tiles, ntiles * sizeof(TileDesc), cudaMemcpyDeviceToHost) != cudaSuccess) { free_plan(); return nullptr; }
        std::vector<int> wt;
        for (int b = 0; b < nty; b++) {
            const int by = (b & 1) ? (nty - 1 - (b >> 1)) : (b >> 1);
            for (int bx = 0; bx < ntx; bx++) {
                const TileDesc &t = ht[(size_t)by * ntx + bx];
                if (t.nfast > 0 && t.kind == 0) wt.push_back(by * ntx + bx);
            }
        }
        if ((int)wt.size() != tp.ntiles_fast - tp.ntiles_rowsep) { fprintf(stderr, "<ezb200> tile plan: window tile count mismatch\n"); free_plan(); return nullptr; }
        if (cudaMalloc((void **)&tp.d_wintiles, wt.size() * sizeof(int)) != cudaSuccess ||
            cudaMemcpy(tp.d_wintiles, wt.data(), wt.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
            (void)cudaGetLastError();
            free_plan();
            return nullptr;
        }
    }
    if (tp.suitable && tp.nlist > 0) {
        if (cudaMalloc((void **)&tp.d_list, (size_t)tp.nlist * sizeof(int)) != cudaSuccess) { (void)cudaGetLastError(); tp.d_list = nullptr; free_plan(); return nullptr; }
        cudaMemsetAsync(d_counters, 0, sizeof(unsigned int), S.stream);
        ezk::tile_list(a, tp.d_flags, tp.d_list, d_counters, 0, S.stream);
        S.launches++;
        if (cudaStreamSynchronize(S.stream) != cudaSuccess) {   // the plan is shared by every session: complete before it is published
            fprintf(stderr, "<ezb200> tile list build failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            free_plan();
            return nullptr;
        }
    }
    if (!tp.suitable) free_plan();
    if (S.opt.verbose > 0)
        fprintf(stderr, "<ezb200> tile plan: %d of %d points on the per-point list, %d of %zu tiles active (%d row-separable), %s; window tiles: "
                "%.2f shared-memory wavefronts per LDS at the chosen pitch (%.2f at the minimal pitch), %d floats per staged level\n", tp.nlist,
                a.npts, tp.ntiles_fast, ntiles, tp.ntiles_rowsep, tp.suitable ? "tiled kernels" : "per-point kernel",
                hc[6] / (8.0 * (tp.ntiles_fast - tp.ntiles_rowsep > 0 ? tp.ntiles_fast - tp.ntiles_rowsep : 1)),
                hc[7] / (8.0 * (tp.ntiles_fast - tp.ntiles_rowsep > 0 ? tp.ntiles_fast - tp.ntiles_rowsep : 1)), tp.cap);
    tp.built = true;
    return &tp;
}

// The per-point list of a tile plan without its constant-extrapolation points, and those points (see TilePlan).
void ensure_const_split(TilePlan &tp, const InterpArgs &a) {
    if (tp.const_tried) return;
    tp.const_tried = true;
    if (!tp.suitable || tp.nlist <= 0) return;
    unsigned int *d_counters = (unsigned int *)S.st_misc[3].get(8 * sizeof(unsigned int));
    if (!d_counters) return;
    if (cudaMalloc((void **)&tp.d_list2, (size_t)tp.nlist * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&tp.d_clist, (size_t)tp.nlist * sizeof(int)) != cudaSuccess) {
        (void)cudaGetLastError();
        cudaFree(tp.d_list2); cudaFree(tp.d_clist);
        tp.d_list2 = tp.d_clist = nullptr;
        return;
    }
    unsigned int n[2] = {0, 0};
    cudaMemsetAsync(d_counters, 0, 2 * sizeof(unsigned int), S.stream);
    ezk::tile_list(a, tp.d_flags, tp.d_list2, d_counters, 1, S.stream);
    ezk::tile_list(a, tp.d_flags, tp.d_clist, d_counters + 1, 2, S.stream);
    S.launches += 2;
    if (cudaMemcpyAsync(n, d_counters, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, S.stream) != cudaSuccess ||
        cudaStreamSynchronize(S.stream) != cudaSuccess || (long)n[0] + (long)n[1] != (long)tp.nlist) {
        (void)cudaGetLastError();
        cudaFree(tp.d_list2); cudaFree(tp.d_clist);
        tp.d_list2 = tp.d_clist = nullptr;
        return;
    }
    tp.nlist2 = (int)n[0];
    tp.nclist = (int)n[1];
    tp.const_ok = true;
}

// Plan cache of a tile plan, built on the first launch that wants it (see sint_dev): 40 B per target point.
void ensure_plan_cache(TilePlan &tp, const InterpArgs &a) {
    if (tp.pc_tried) return;
    tp.pc_tried = true;
    static const bool pc_on = !(getenv("EZB200_PLAN_CACHE") && atoi(getenv("EZB200_PLAN_CACHE")) == 0);
    if (tp.suitable && pc_on && a.src.ni < 65536 && (a.src.j2 - a.src.j1 + 1) < 32768 && (!a.yy || (a.src2.ni < 65536 && (a.src2.j2 - a.src2.j1 + 1) < 32768))) {
        const size_t n = (size_t)a.npts;
        if (cudaMalloc((void **)&tp.d_pc_idx, n * sizeof(int)) == cudaSuccess && cudaMalloc((void **)&tp.d_pc_wx, n * sizeof(float4)) == cudaSuccess &&
            cudaMalloc((void **)&tp.d_pc_wy, n * sizeof(float4)) == cudaSuccess && cudaMalloc((void **)&tp.d_pc_sm1, n * sizeof(float)) == cudaSuccess) {
            ezk::plan_cache(a, tp.d_pc_idx, tp.d_pc_wx, tp.d_pc_wy, tp.d_pc_sm1, S.stream);
            S.launches++;
            (void)cudaStreamSynchronize(S.stream);   // shared by every session: complete before other streams may read it
        } else {   // not enough memory: the tile kernels recompute the plans
            (void)cudaGetLastError();
            cudaFree(tp.d_pc_idx); cudaFree(tp.d_pc_wx); cudaFree(tp.d_pc_wy); cudaFree(tp.d_pc_sm1);
            tp.d_pc_idx = nullptr; tp.d_pc_wx = tp.d_pc_wy = nullptr; tp.d_pc_sm1 = nullptr;
        }
    }
}

void free_gset(Gset &gs) {
    for (auto &tp : gs.plans) {
        if (tp.d_tiles) cudaFree(tp.d_tiles);
        if (tp.d_flags) cudaFree(tp.d_flags);
        if (tp.d_list) cudaFree(tp.d_list);
        if (tp.d_list2) cudaFree(tp.d_list2);
        if (tp.d_clist) cudaFree(tp.d_clist);
        if (tp.d_wintiles) cudaFree(tp.d_wintiles);
        if (tp.d_rowy) cudaFree(tp.d_rowy);
        if (tp.d_pc_idx) cudaFree(tp.d_pc_idx);
        if (tp.d_pc_wx) cudaFree(tp.d_pc_wx);
        if (tp.d_pc_wy) cudaFree(tp.d_pc_wy);
        if (tp.d_pc_sm1) cudaFree(tp.d_pc_sm1);
        tp = TilePlan();
    }
    if (gs.d_x) cudaFree(gs.d_x);
    if (gs.d_y) cudaFree(gs.d_y);
    if (gs.d_zb) cudaFree(gs.d_zb);
    if (gs.d_perm) cudaFree(gs.d_perm);
    for (int k = 0; k < 2; k++)
        if (gs.d_zlist[k]) cudaFree(gs.d_zlist[k]);
    if (gs.d_wrot) cudaFree(gs.d_wrot);
    if (gs.d_cidx) cudaFree(gs.d_cidx);
    if (gs.d_cwts) cudaFree(gs.d_cwts);
    if (gs.d_cmask) cudaFree(gs.d_cmask);
    gs.d_cidx = gs.d_cmask = nullptr; gs.d_cwts = nullptr; gs.cloud = false;
    gs.d_wrot = nullptr; gs.wrot_tried = false;
    gs.d_perm = nullptr; gs.perm_tried = false;
    for (int k = 0; k < 2; k++) { gs.d_zlist[k] = nullptr; gs.nzlist[k] = 0; gs.zlist_built[k] = false; }
    cudaFree(gs.d_dc_idx); cudaFree(gs.d_dc_wx); cudaFree(gs.d_dc_wy); cudaFree(gs.d_dc_sm1);
    gs.d_dc_idx = nullptr; gs.d_dc_wx = gs.d_dc_wy = nullptr; gs.d_dc_sm1 = nullptr; gs.dc_tried = false; gs.dc_calls = 0;
    gs.d_x = gs.d_y = nullptr;
    gs.d_zb = nullptr;
    gs.built = false;
}

// ---------------------------------------------------------------------------------------------------
// host <-> device staging of caller arrays
// ---------------------------------------------------------------------------------------------------
struct Staged {
    float *dev = nullptr;
    float *host = nullptr;   // non-null when a copy back is due
    size_t n = 0;
};

int stage_in(Staged &s, DevBuf &buf, const float *p, size_t n) {
    s.n = n;
    if (is_device_ptr(p)) { s.dev = const_cast<float *>(p); return 0; }
    s.dev = (float *)buf.get(n * sizeof(float));
    if (!s.dev) { fprintf(stderr, "<ezb200> cannot allocate %zu bytes of device staging\n", n * sizeof(float)); return -1; }
    Unlocked u;   // a copy from pageable memory blocks the calling thread until it has been staged
    CUDA_OK(cudaMemcpyAsync(s.dev, p, n * sizeof(float), cudaMemcpyHostToDevice, S.stream));
    return 0;
}
int stage_out(Staged &s, DevBuf &buf, float *p, size_t n) {
    s.n = n;
    if (is_device_ptr(p)) { s.dev = p; return 0; }
    s.dev = (float *)buf.get(n * sizeof(float));
    if (!s.dev) { fprintf(stderr, "<ezb200> cannot allocate %zu bytes of device staging\n", n * sizeof(float)); return -1; }
    s.host = p;
    return 0;
}
int finish_out(Staged &s) {
    if (!s.host) return 0;
    Unlocked u;
    CUDA_OK(cudaMemcpyAsync(s.host, s.dev, s.n * sizeof(float), cudaMemcpyDeviceToHost, S.stream));
    CUDA_OK(cudaStreamSynchronize(S.stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// the scalar path on device pointers: c_ezsint_orig (ezsint.c:75-135) for nfields fields
// ---------------------------------------------------------------------------------------------------
InterpOpts snapshot_opts(int vector, int sym_sign) {
    InterpOpts o;
    o.degree = S.opt.degre_interp;
    o.extrap = S.opt.degre_extrap;
    o.extrap_value = S.opt.valeur_extrap;
    o.vector = vector;
    o.sym_sign = sym_sign;
    o.polar = S.opt.polar_correction;
    o.arith = S.arith;
    return o;
}

int check_degree(bool grid_call = false) {
    int d = S.opt.degre_interp;
    if (d == EZ_NEAREST || d == EZ_LINEAR || d == EZ_CUBIC) return 0;
    if (grid_call && (d == EZ_AVERAGE || d == EZ_SPH_AVERAGE)) return 0;   // c_ezsint / c_ezuvint: ez_avg / ez_avg_sph (sint_dev)
    fprintf(stderr, "<ezb200> interp_degree average/sph_average applies to grid-to-grid calls only (ez_avg.inc works on the whole target grid)\n");
    return -1;
}

// Output-row band of the current session for target `out` (GPU g of G owns rows [g nj / G, (g + 1) nj / G) of a call
// with fewer fields than GPUs; the caller's output arrays then hold (j1 - j0) * ni points per field).
struct Band { int j0, j1; bool on; };
Band band_of(const Grid &out) {
    Band b{0, out.nj, false};
    if (S.row0 >= 0 && S.row1 > S.row0) {
        b.j0 = S.row0 < out.nj ? S.row0 : out.nj;
        b.j1 = S.row1 < out.nj ? S.row1 : out.nj;
        b.on = !(b.j0 == 0 && b.j1 == out.nj);
    }
    return b;
}
size_t out_points(const Grid &out) {
    const Band b = band_of(out);
    return (size_t)out.ni * (b.j1 - b.j0);
}

// zin/zout are DEVICE pointers here.  Returns 0, 2 (extrapolated) or -1.
// fork / join resources of the session: kernels of one launch that write disjoint target points run on separate streams
bool ensure_fork() {
    if (!S.fork_tried) {
        S.fork_tried = true;
        static const bool conc_on = !(getenv("EZB200_CONCURRENT") && atoi(getenv("EZB200_CONCURRENT")) == 0);
        S.fork_ok = conc_on && cudaStreamCreateWithFlags(&S.fork.aux[0], cudaStreamNonBlocking) == cudaSuccess &&
                    cudaStreamCreateWithFlags(&S.fork.aux[1], cudaStreamNonBlocking) == cudaSuccess &&
                    cudaEventCreateWithFlags(&S.fork.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&S.fork.ev_join[0], cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&S.fork.ev_join[1], cudaEventDisableTiming) == cudaSuccess;
        if (!S.fork_ok) (void)cudaGetLastError();
    }
    return S.fork_ok;
}

// Scattered ('Y' cloud) targets arrive in arbitrary order: the kernels walk them sorted by source cell (a permutation built once
// per gridset) so that neighbouring threads gather from the same sectors.  nullptr when it does not apply or cannot be built.
const int *cloud_target_order(Gset &gs, const InterpArgs &a, const Grid &out) {
    if (out.grtyp != 'Y' || a.o.arith < 1 || a.npts < 32768 || band_of(out).on) return nullptr;
    if (!gs.perm_tried) {
        gs.perm_tried = true;
        if (cudaMalloc((void **)&gs.d_perm, (size_t)a.npts * sizeof(int)) == cudaSuccess) {
            const GridDev &sg = a.src;
            if (!ezk::sort_points_by_cell(gs.d_x, gs.d_y, a.yy ? gs.d_zb : nullptr, sg.ni, sg.j2 - sg.j1 + 1, a.npts, gs.d_perm, S.stream)) {
                cudaFree(gs.d_perm);
                gs.d_perm = nullptr;
            }
            S.launches += 2;
        } else {
            gs.d_perm = nullptr;
            (void)cudaGetLastError();
        }
    }
    return gs.d_perm;
}

// Zone list of a gridset for the direct kernel, built on first use (shared by every session: complete before it is published).
bool ensure_zonelist(Gset &gs, int skip_dehors_only, int npts, const int *order = nullptr) {
    const int k = skip_dehors_only ? 1 : 0;
    if (gs.zlist_built[k]) return true;
    const int *c = gs.counts;
    const long bound = std::min<long>(npts, (long)(skip_dehors_only ? 0 : c[0]) + c[1] + c[2] + c[3] + c[4]);
    gs.nzlist[k] = 0;
    if (bound > 0) {
        unsigned int *d_cursor = (unsigned int *)S.st_misc[3].get(8 * sizeof(unsigned int));
        if (!d_cursor) return false;
        if (cudaMalloc((void **)&gs.d_zlist[k], (size_t)bound * sizeof(int)) != cudaSuccess) { (void)cudaGetLastError(); gs.d_zlist[k] = nullptr; return false; }
        unsigned int n = 0;
        cudaMemsetAsync(d_cursor, 0, sizeof(unsigned int), S.stream);
        ezk::zone_list(gs.d_zb, npts, skip_dehors_only, gs.d_zlist[k], d_cursor, S.stream, order);
        S.launches++;
        if (cudaMemcpyAsync(&n, d_cursor, sizeof(unsigned int), cudaMemcpyDeviceToHost, S.stream) != cudaSuccess ||
            cudaStreamSynchronize(S.stream) != cudaSuccess || (long)n > bound) {
            fprintf(stderr, "<ezb200> zone list build failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            cudaFree(gs.d_zlist[k]);
            gs.d_zlist[k] = nullptr;
            return false;
        }
        gs.nzlist[k] = (int)n;
    }
    gs.zlist_built[k] = true;
    return true;
}

// wrot / fused: c_ezuvint of one pair hands in the cached wind rotation; *fused reports whether the launch applied it
int sint_dev(Grid &out, Grid &in, float *d_zout, const float *d_zin, int nfields, long in_stride, long out_stride, int vector, int sym_sign,
             const float *polerow = nullptr, const float4 *wrot = nullptr, bool *fused = nullptr) {
    if (fused) *fused = false;
    if (check_degree(true) < 0) return -1;
    Gset *gs = ensure_gset(out, in);
    if (!gs) return -1;
    if (gs->cloud) {
        // gdinterp.c:207-235.  cloud_interp_alg = linear on a 2-D cloud would run ez_rgdint_1_nw on x, y that the reference never
        // computes (ez_calcxy_y.inc:60-63 is commented out): undefined there, refused here.  Zones / polar correction do not apply:
        // with polar_correction = yes the reference classifies the (uninitialised) x, y and, where they happen to be zero,
        // overwrites the whole result with the extrapolation value -- callers of cloud sources switch it off.
        if (in.ni > 1 && in.nj > 1 && S.opt.cld_interp_alg == EZ_LINEAR) {
            fprintf(stderr, "<ezb200> cloud_interp_alg = linear on a 2-D 'Y' source reads coordinates the reference never computes; use distance\n");
            return -1;
        }
        if (band_of(out).on) { fprintf(stderr, "<ezb200> output-row bands are not supported for 'Y' sources\n"); return -1; }
        unsigned int *scr = (unsigned int *)S.st_scratch.get((size_t)nfields * 2 * sizeof(unsigned int));
        if (!scr) return -1;
        S.launches += ezk::cloud_apply(gs->d_cidx, gs->d_cwts, gs->d_cmask, gs->cloud_nw, d_zin, d_zout, in_stride, out_stride, nfields, in.ni * in.nj,
                                       out.ni * out.nj, out.ni, S.opt.degre_extrap == EZ_VALUE, S.opt.valeur_extrap, scr, S.stream);
        S.last_path = 1; S.last_nlist = 0;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    InterpArgs a;
    memset(&a, 0, sizeof(a));
    a.src = to_dev(in);
    a.x = gs->d_x; a.y = gs->d_y; a.zb = gs->d_zb;
    a.zin = d_zin; a.zout = d_zout;
    a.in_stride = in_stride; a.out_stride = out_stride;
    a.nfields = nfields;
    a.npts = out.ni * out.nj;
    a.o = snapshot_opts(vector, sym_sign);
    a.polerow = polerow;
    if (!gs->yy && (in.needs_expansion || in.axe_y_inverse)) {
        // ezsint.c:100-113: hemispheric sources are mirrored about the equator (with the sign of the field's symmetry) and
        // y-inverted ones flipped, once per call, into the logical layout z(ni, j1:j2); every kernel downstream -- the tile
        // kernels included -- then sees a plain field (xpn / yinv cleared in the descriptor it is given)
        const long exp_stride = (long)in.ni * (in.j2 - in.j1 + 1);
        float *d_exp = (float *)S.st_xpn.get((size_t)exp_stride * nfields * sizeof(float));
        if (!d_exp) { fprintf(stderr, "<ezb200> cannot allocate the expanded source field\n"); return -1; }
        ezk::expand_field(a.src, d_zin, d_exp, in_stride, exp_stride, nfields, sym_sign, S.stream);
        S.launches++;
        a.src.xpn = XPN_NONE;
        a.src.yinv = 0;
        a.zin = d_exp;
        a.in_stride = exp_stride;
        d_zin = d_exp;
        in_stride = exp_stride;
    }
    int rc = 0;
    ezk::DirectPre pre;   // per-field scalars pre-pass (pole means, extrapolation value), launched once the kernel path is known
    memset(&pre, 0, sizeof(pre));
    if (gs->yy) {
        Grid *yin = grid_of(in.subgrid[0]), *yan = grid_of(in.subgrid[1]);
        a.src = to_dev(*yin);
        a.src2 = to_dev(*yan);
        a.yy = 1;
        a.panel_stride = (long)yin->ni * yin->nj;
        a.use_zones = 0;   // c_gdxysint calls c_gdinterp directly: no zone correction on this route (gdxysint.c:30-54)
    } else {
        const int *c = gs->counts;
        bool any = (c[0] + c[1] + c[2] + c[3] + c[4]) > 0;
        a.use_zones = (S.opt.polar_correction == 1 && any) ? 1 : 0;
        if (a.use_zones) {
            if (c[0] > 0) {
                rc = 2;
                if (S.opt.degre_extrap == EZ_ABORT) {
                    fprintf(stderr, "<ez_corrval> There are points on the destination grid that lie outside the source grid\n");
                    fprintf(stderr, "<ez_corrval> aborting at your request!\n\n\n");
                    a.use_zones = 0;
                    rc = -1;
                }
            }
        }
        if (a.use_zones) {
            FieldScalars *fs = (FieldScalars *)S.st_fs.get((size_t)nfields * sizeof(FieldScalars));
            unsigned int *scr = (unsigned int *)S.st_scratch.get((size_t)nfields * 2 * sizeof(unsigned int));
            if (!fs || !scr) return -1;
            // (vector components never read them: their pole values come from the synthetic pole rows, their outside value is 0)
            pre.run = !vector;
            pre.need_poles = (!vector) && (c[1] + c[2] + c[3] + c[4] > 0);
            pre.need_minmax = (!vector) && c[0] > 0 && (S.opt.degre_extrap == EZ_MAXIMUM || S.opt.degre_extrap == EZ_MINIMUM);
            pre.fs = fs; pre.scratch = scr;
            a.fs = vector ? nullptr : fs;
        }
    }
    a.out_ni = out.ni;
    a.out_nj = out.nj;
    // (fields of one call may come from different allocations -- the fused (u, v) launch --, so the first and the last are checked)
    const size_t field_bytes = (gs->yy ? 2 * (size_t)a.panel_stride : (size_t)a.src.ni * (a.src.j2 - a.src.j1 + 1)) * sizeof(float);
    if (a.o.degree == EZ_AVERAGE || a.o.degree == EZ_SPH_AVERAGE) {
        // gdinterp.c:32-75,118-131: c_gdcompatible_grids decides on the TARGET type alone ; otherwise the call is a linear one
        if (!(out.grtyp == 'L' || out.grtyp == 'A' || out.grtyp == 'B' || out.grtyp == 'G')) {
            fprintf(stderr, "(gdinterp) input and output grids are not compatible for average computation\n");
            fprintf(stderr, "(gdinterp) interpolaton level set to linear\n");
            a.o.degree = EZ_LINEAR;
        } else {
            if (gs->yy || in.needs_expansion || in.axe_y_inverse) {
                fprintf(stderr, "<ezb200> interp_degree average / sph_average is not supported for 'U', hemispheric or y-inverted sources\n");
                return -1;
            }
            const Band band = band_of(out);
            if (band.on && band.j1 <= band.j0) return rc;
            const int sph = a.o.degree == EZ_SPH_AVERAGE;
            if (sph && ensure_latlon(out) < 0) return -1;
            int nz = 0;
            const int *zl = nullptr;
            if (a.use_zones) {
                if (!ensure_zonelist(*gs, 0, a.npts)) return -1;
                zl = gs->d_zlist[0]; nz = gs->nzlist[0];
            }
            if (band.on) {   // full-grid indexing lands in the band buffer
                a.band_j0 = band.j0; a.band_j1 = band.j1;
                a.zout = d_zout - (long)band.j0 * out.ni;
            }
            if (pre.run) {
                ezk::field_scalars(a, pre.need_poles, pre.need_minmax, pre.fs, pre.scratch, S.stream);
                S.launches += pre.need_minmax ? 3 : 1;
            }
            S.launches += ezk::interp_avg(a, sph ? out.ds().d_lat : nullptr, sph, zl, nz, S.stream);
            S.last_path = 32 | (nz > 0 ? 4 : 0);
            S.last_nlist = nz;
            S.last_tiles[0] = S.last_tiles[1] = S.last_tiles[2] = 0;
            S.timing_valid = false;
            if (S.opt.polar_correction == 1 && !vector && rc >= 0 && (in.grtyp == 'Z' || in.grtyp == '#') && in.grref == 'E' && out.grtyp == 'B') {
                if (band.on) {
                    fprintf(stderr, "<ezb200> output-row bands are not supported for 'B' targets fed from a rotated source (pole rows are averaged over the whole row set)\n");
                    return -1;
                }
                ezk::corrbgd(d_zout, out_stride, nfields, out.ni, out.nj, out.ig[0], S.stream);   // ez_corrval.c:144-147
                S.launches++;
            }
            CUDA_OK(cudaGetLastError());
            return rc;
        }
    }
    // Few fields (c_ezsint of one field, c_ezuvint of one pair): a tile CTA's set-up (descriptor, plans of four points per thread,
    // a copy pipeline with one stage to fill) is all there is to such a launch; the direct kernel (ez_interp_fast.cuh) takes it.
    // The choice follows the field count of the API CALL, not of the level chunk or device block at hand, so that a call's
    // result does not depend on how it was cut.  c_ezb200_set_direct_max_fields / EZB200_DIRECT_MAX_FIELDS override (0 = never).
    // ('Y' cloud targets have no tile path: for them the direct kernel replaces the per-point kernel at any field count)
    const bool direct_few = (S.call_fields > 0 ? S.call_fields : nfields) <= S.direct_max && nfields <= 8 && out.grtyp != 'Y';
    const bool direct_cloud = out.grtyp == 'Y' && S.direct_max > 0 && !band_of(out).on;
    if ((direct_few || direct_cloud) && ezk::direct_eligible(a)) {
        const Band band = band_of(out);
        if (band.on && band.j1 <= band.j0) return rc;
        int nz = 0;
        const int *zl = nullptr;
        a.order = direct_cloud ? cloud_target_order(*gs, a, out) : nullptr;
        if (a.use_zones) {
            const int k = (a.o.extrap >= EZ_MAXIMUM) ? 1 : 0;
            if (!ensure_zonelist(*gs, k, a.npts, a.order)) return -1;
            zl = gs->d_zlist[k]; nz = gs->nzlist[k];
        }
        pre.has_dehors = gs->counts[0] > 0;
        float *zo = d_zout;
        if (band.on) {   // full-grid indexing lands in the band buffer
            a.band_j0 = band.j0; a.band_j1 = band.j1;
            zo = d_zout - (long)band.j0 * out.ni;
            a.zout = zo;
        }
        a.wrot = (wrot && nfields == 2) ? wrot : nullptr;
        // cubic on an irregular source: from the second call on, the per-point plans come from a cache (bit-identical weights)
        if (a.o.degree == EZ_CUBIC && (a.yy || in.grtyp == 'Z' || in.grtyp == '#' || in.grtyp == 'G') && a.src.ni < 65536 &&
            (a.src.j2 - a.src.j1 + 1) < 32768 && (!a.yy || (a.src2.ni < 65536 && (a.src2.j2 - a.src2.j1 + 1) < 32768))) {
            static const bool dc_on = !(getenv("EZB200_DIRECT_PLAN_CACHE") && atoi(getenv("EZB200_DIRECT_PLAN_CACHE")) == 0);
            if (dc_on && !gs->dc_tried && ++gs->dc_calls >= 2) {
                gs->dc_tried = true;
                const size_t n = (size_t)a.npts;
                if (cudaMalloc((void **)&gs->d_dc_idx, n * sizeof(int)) == cudaSuccess && cudaMalloc((void **)&gs->d_dc_wx, n * sizeof(float4)) == cudaSuccess &&
                    cudaMalloc((void **)&gs->d_dc_wy, n * sizeof(float4)) == cudaSuccess && cudaMalloc((void **)&gs->d_dc_sm1, n * sizeof(float)) == cudaSuccess) {
                    InterpArgs pa = a;
                    pa.band_j0 = pa.band_j1 = 0;
                    ezk::plan_cache(pa, gs->d_dc_idx, gs->d_dc_wx, gs->d_dc_wy, gs->d_dc_sm1, S.stream, 1);
                    S.launches++;
                    (void)cudaStreamSynchronize(S.stream);   // shared by every session: complete before other streams may read it
                } else {
                    (void)cudaGetLastError();
                    cudaFree(gs->d_dc_idx); cudaFree(gs->d_dc_wx); cudaFree(gs->d_dc_wy); cudaFree(gs->d_dc_sm1);
                    gs->d_dc_idx = nullptr; gs->d_dc_wx = gs->d_dc_wy = nullptr; gs->d_dc_sm1 = nullptr;
                }
            }
            if (gs->d_dc_idx) { a.pc_idx = gs->d_dc_idx; a.pc_wx = gs->d_dc_wx; a.pc_wy = gs->d_dc_wy; a.pc_sm1 = gs->d_dc_sm1; }
        }
        const bool timed_direct = S.timing && S.tev.ev[0];
        S.launches += ezk::interp_direct(a, zl, nz, S.stream, ensure_fork() ? &S.fork : nullptr, &pre, timed_direct ? &S.tev : nullptr);
        if (a.wrot) {
            if (nz > 0) {
                ezk::wind_apply_list(a.wrot, zo, zo + out_stride, 0, 1, zl, nz, band.j0 * out.ni, band.j1 * out.ni, S.stream);
                S.launches++;
            }
            if (fused) *fused = true;
        }
        S.last_path = 16 | (nz > 0 ? 4 : 0);
        S.last_nlist = nz;
        S.last_tiles[0] = S.last_tiles[1] = S.last_tiles[2] = 0;
        S.timing_valid = timed_direct;
        if (!gs->yy && S.opt.polar_correction == 1 && !vector && rc >= 0 && (in.grtyp == 'Z' || in.grtyp == '#') && in.grref == 'E' && out.grtyp == 'B') {
            if (band.on) {
                fprintf(stderr, "<ezb200> output-row bands are not supported for 'B' targets fed from a rotated source (pole rows are averaged over the whole row set)\n");
                return -1;
            }
            ezk::corrbgd(d_zout, out_stride, nfields, out.ni, out.nj, out.ig[0], S.stream);   // ez_corrval.c:144-147
            S.launches++;
        }
        CUDA_OK(cudaGetLastError());
        return rc;
    }
    if (pre.run) {
        ezk::field_scalars(a, pre.need_poles, pre.need_minmax, pre.fs, pre.scratch, S.stream);
        S.launches += pre.need_minmax ? 3 : 1;
    }
    static const int tiled_min_fields = getenv("EZB200_TILED_MIN_FIELDS") ? atoi(getenv("EZB200_TILED_MIN_FIELDS")) : 1;
    if (nfields >= tiled_min_fields && ezk::tiled_eligible(a) && source_blocks_inside_allocation(d_zin, field_bytes) &&
        (nfields == 1 || source_blocks_inside_allocation(d_zin + (long)(nfields - 1) * in_stride, field_bytes))) {
        TilePlan *tp = ensure_tileplan(*gs, a);
        if (!tp) return -1;
        if (tp->suitable) {
            a.tiles = tp->d_tiles;
            a.tile_flags = tp->d_flags;
            a.tile_cap = tp->cap;
            a.tile_rowy = tp->d_rowy;
            // The cached plans pay when the plan arithmetic is heavy (Newton weights of irregular Z / # / G sources) or the
            // launch has many levels to amortise nothing else over; a single field from a regular source recomputes them
            // faster than it would read 40 B per point.
            const bool irregular = a.yy || in.grtyp == 'Z' || in.grtyp == '#' || in.grtyp == 'G';
            if (irregular || nfields >= 8) {
                ensure_plan_cache(*tp, a);
                a.pc_idx = tp->d_pc_idx; a.pc_wx = tp->d_pc_wx; a.pc_wy = tp->d_pc_wy; a.pc_sm1 = tp->d_pc_sm1;
            }
            a.tile_vw = tp->vw;
            a.tile_cap_rs = tp->cap_rs;
            a.ntiles_rowsep = tp->ntiles_rowsep;
            a.ntiles_window = tp->ntiles_fast - tp->ntiles_rowsep;
            a.win_tiles = tp->d_wintiles;
            a.list = tp->d_list;
            a.nlist = tp->nlist;
            if (a.use_zones && a.o.extrap >= EZ_MAXIMUM && gs->counts[0] > 0 && tp->nlist > 0) {
                ensure_const_split(*tp, a);
                if (tp->const_ok) { a.list = tp->d_list2; a.nlist = tp->nlist2; a.clist = tp->d_clist; a.nclist = tp->nclist; }
            }
        }
    }
    // scattered ('Y' cloud) targets arrive in arbitrary order: the per-point kernel walks them sorted by source cell
    if (!a.tiles) {
        const int *perm = cloud_target_order(*gs, a, out);
        if (perm) { a.list = perm; a.nlist = a.npts; }
    }
    const Band band = band_of(out);
    if (band.on) {
        if (band.j1 <= band.j0) return rc;
        const long off = (long)band.j0 * out.ni;
        if (a.tiles && (band.j0 % 32) == 0 && ((band.j1 % 32) == 0 || band.j1 == out.nj)) {
            // whole tile rows: the tile kernels and the list kernel skip what lies outside ; full-grid indexing lands in the band buffer
            a.band_j0 = band.j0; a.band_j1 = band.j1;
            a.zout = d_zout - off;
        } else {
            // any other band: the per-point kernel on the band's points (same plans, same arithmetic class as a full per-point call)
            a.tiles = nullptr; a.tile_flags = nullptr; a.list = nullptr; a.nlist = 0; a.clist = nullptr; a.nclist = 0;
            a.ntiles_rowsep = a.ntiles_window = 0;
            a.pc_idx = nullptr; a.pc_wx = a.pc_wy = nullptr; a.pc_sm1 = nullptr;
            a.x += off; a.y += off;
            if (a.zb) a.zb += off;
            a.npts = (band.j1 - band.j0) * out.ni;
            a.out_nj = band.j1 - band.j0;
        }
    }
    S.last_path = a.tiles ? ((a.ntiles_window > 0 ? 2 : 0) | ((a.nlist > 0 || a.nclist > 0) ? 4 : 0) | (a.ntiles_rowsep > 0 ? 8 : 0)) : 1;
    S.last_nlist = a.tiles ? a.nlist + a.nclist : 0;
    S.last_tiles[0] = a.tiles ? a.ntiles_rowsep : 0;
    S.last_tiles[1] = a.tiles ? a.ntiles_window : 0;
    S.last_tiles[2] = S.last_tiles[0] + S.last_tiles[1];
    const bool timed = S.timing && a.tiles && S.tev.ev[0];
    S.timing_valid = timed;
    S.launches += ezk::interp(a, S.stream, timed ? &S.tev : nullptr, (a.tiles && ensure_fork()) ? &S.fork : nullptr);
    // ez_corrval.c:144-147: B-grid targets fed from a rotated Z/# source get their pole rows averaged
    if (!gs->yy && S.opt.polar_correction == 1 && !vector && rc >= 0 && (in.grtyp == 'Z' || in.grtyp == '#') && in.grref == 'E' && out.grtyp == 'B') {
        if (band.on) {
            fprintf(stderr, "<ezb200> output-row bands are not supported for 'B' targets fed from a rotated source (pole rows are averaged over the whole row set)\n");
            return -1;
        }
        ezk::corrbgd(d_zout, out_stride, nfields, out.ni, out.nj, out.ig[0], S.stream);
        S.launches++;
    }
    CUDA_OK(cudaGetLastError());
    return rc;
}

// c_ezyysint (ezyysint.c:26-242) + c_ezsint_orig, on device pointers
int sint_any_dev(Grid &out, Grid &in, float *d_zout, const float *d_zin, int nfields, long in_stride, long out_stride, int vector, int sym_sign) {
    const bool yyin = in.nsubgrids > 0, yyout = out.nsubgrids > 0;
    if (yyout && S.row0 >= 0) {
        fprintf(stderr, "<ezb200> output-row bands (c_ezb200_set_rows) are not supported for 'U' targets: split by sub-grid instead\n");
        return -1;
    }
    if (!yyin && !yyout) return sint_dev(out, in, d_zout, d_zin, nfields, in_stride, out_stride, vector, sym_sign);
    Grid *yin_out = yyout ? grid_of(out.subgrid[0]) : &out;
    Grid *yan_out = yyout ? grid_of(out.subgrid[1]) : nullptr;
    Grid *yin_in = yyin ? grid_of(in.subgrid[0]) : &in;
    Grid *yan_in = yyin ? grid_of(in.subgrid[1]) : nullptr;
    const long half_out = (long)yin_out->ni * yin_out->nj;
    const long half_in = (long)yin_in->ni * yin_in->nj;
    if (!yyin && yyout) {   // ezyysint.c:86-95
        int r1 = sint_dev(*yin_out, in, d_zout, d_zin, nfields, in_stride, out_stride, vector, sym_sign);
        int r2 = sint_dev(*yan_out, in, d_zout + half_out, d_zin, nfields, in_stride, out_stride, vector, sym_sign);
        if (r1 < 0 || r2 < 0) return -1;
        return (r1 == 2 || r2 == 2) ? 2 : 0;
    }
    // a source sub-grid IS the target (ezyysint.c:99-108)
    if (yin_in->id == out.id) {
        for (int f = 0; f < nfields; f++)
            cudaMemcpyAsync(d_zout + f * out_stride, d_zin + f * in_stride, half_in * sizeof(float), cudaMemcpyDeviceToDevice, S.stream);
        return 1;
    }
    if (yan_in && yan_in->id == out.id) {
        for (int f = 0; f < nfields; f++)
            cudaMemcpyAsync(d_zout + f * out_stride, d_zin + f * in_stride + half_in, half_in * sizeof(float), cudaMemcpyDeviceToDevice, S.stream);
        return 1;
    }
    if (S.opt.use_1subgrid == 1) {   // ezyysint.c:113-138
        if (yyout) {
            fprintf(stderr, "<c_ezyysint> cannot use 1 subgrid to interpolate to a Yin-Yang grid  Aborting...\n");
            return -1;
        }
        if (S.opt.valeur_1subgrid != yin_in->id && S.opt.valeur_1subgrid != yan_in->id) {
            fprintf(stderr, "<c_ezyysint> define src subgridid in ezsetival(subgridid)! Aborting...\n");
            return -1;
        }
        if (S.opt.valeur_1subgrid == yin_in->id) return sint_dev(*yin_out, *yin_in, d_zout, d_zin, nfields, in_stride, out_stride, vector, sym_sign);
        return sint_dev(*yin_out, *yan_in, d_zout, d_zin + half_in, nfields, in_stride, out_stride, vector, sym_sign);
    }
    // both panels, hard Yin-priority selection (ezyysint.c:142-239)
    int rc = sint_dev(*yin_out, in, d_zout, d_zin, nfields, in_stride, out_stride, vector, sym_sign);
    if (rc < 0) return rc;
    if (yyout) {
        rc = sint_dev(*yan_out, in, d_zout + half_out, d_zin, nfields, in_stride, out_stride, vector, sym_sign);
        if (rc < 0) return rc;
    }
    return 0;
}

// A call whose arrays all live in host memory is synchronous by construction (the results are back in the caller's arrays when
// it returns), so it does not have to be ordered with anything the caller queued on the legacy default stream: threads that
// never chose a stream run such calls on a private non-blocking stream, and concurrent host threads overlap on the device
// instead of serialising on stream 0.  Device-pointer calls keep the caller's (or the default) stream semantics.
struct HostCallStream {
    cudaStream_t saved;
    bool active = false;
    explicit HostCallStream(bool all_host) : saved(S.stream) {
        if (!all_host || S.stream_user) return;
        if (!S.own_stream && cudaStreamCreateWithFlags(&S.own_stream, cudaStreamNonBlocking) != cudaSuccess) {
            (void)cudaGetLastError();
            S.own_stream = nullptr;
            return;
        }
        S.stream = S.own_stream;
        active = true;
    }
    ~HostCallStream() { if (active) S.stream = saved; }
};

bool session_ok(const char *who) {
    if (S.gdin == -1 || S.gdout == -1 || !grid_of(S.gdin) || !grid_of(S.gdout)) {
        fprintf(stderr, "<%s> Source or target grid undefined! Aborting...\n", who);
        return false;
    }
    return true;
}

int ensure_pipe() {
    if (S.s_h2d) return 0;
    CUDA_OK(cudaStreamCreateWithFlags(&S.s_h2d, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&S.s_d2h, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        CUDA_OK(cudaEventCreateWithFlags(&S.ev_in[k], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&S.ev_comp[k], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&S.ev_out[k], cudaEventDisableTiming));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// host-pointer calls: the level-chunk pipeline
// ---------------------------------------------------------------------------------------------------
// Host-side copy pool for PAGEABLE caller arrays.  cudaMemcpyAsync from / to plain malloc'ed memory is staged by the driver
// through one bounce buffer and blocks the calling thread (measured: 16 GB/s, 4x below the pinned rate, and it serialises the
// three-stream pipeline).  Here pageable chunks go through a pinned ring owned by the session: a few helper threads (plus the
// caller) memcpy the chunk into / out of the ring while the DMA engines and the kernels work on the neighbouring chunks.
class CopyPool {
    struct Batch {   // lives on the waiter's stack: `left` only changes under `m`, so the waiter cannot leave while a worker still holds it
        int left = 0;
        std::mutex m;
        std::condition_variable cv;
    };
    struct Job { char *dst; const char *src; size_t n; Batch *b; };
    std::vector<std::thread> threads;
    std::deque<Job> q;
    std::mutex m;
    std::condition_variable cv;
    bool started = false;
    int nthreads = 0;

    void worker() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return !q.empty(); });
                j = q.front();
                q.pop_front();
            }
            memcpy(j.dst, j.src, j.n);
            {
                std::lock_guard<std::mutex> lk(j.b->m);
                if (--j.b->left == 0) j.b->cv.notify_all();
            }
        }
    }

public:
    int helpers() {
        std::lock_guard<std::mutex> lk(m);
        if (!started) {
            started = true;
            const char *e = getenv("EZB200_COPY_THREADS");
            unsigned hw = std::thread::hardware_concurrency();
            nthreads = e ? atoi(e) : (int)(hw >= 16 ? 7 : (hw >= 8 ? 3 : 1));
            if (nthreads < 0) nthreads = 0;
            if (nthreads > 31) nthreads = 31;
            for (int k = 0; k < nthreads; k++) {
                threads.emplace_back([this] { worker(); });
                threads.back().detach();   // the pool lives as long as the process
            }
        }
        return nthreads;
    }
    // blocking parallel memcpy: the caller copies one slice itself
    void copy(void *dst, const void *src, size_t bytes) {
        const int nh = helpers();
        const size_t slice_min = (size_t)1 << 20;
        int parts = (int)std::min<size_t>((size_t)nh + 1, (bytes + slice_min - 1) / slice_min);
        if (parts <= 1) { memcpy(dst, src, bytes); return; }
        Batch b;
        const size_t per = ((bytes / parts) + 4095) & ~(size_t)4095;
        int njobs = 0;
        for (int k = 1; k < parts; k++)
            if (per * k < bytes) njobs++;
        b.left = njobs;
        {
            std::lock_guard<std::mutex> lk(m);
            for (int k = 1; k < parts; k++) {
                const size_t o = per * k;
                if (o >= bytes) continue;
                q.push_back(Job{(char *)dst + o, (const char *)src + o, std::min(per, bytes - o), &b});
            }
        }
        cv.notify_all();
        memcpy(dst, src, std::min(per, bytes));
        std::unique_lock<std::mutex> lk(b.m);
        b.cv.wait(lk, [&] { return b.left <= 0; });
    }
};
// never destroyed: its helper threads wait on its condition variable until the process ends (destroying a condition variable
// with waiters blocks in pthread_cond_destroy)
CopyPool &g_copy_pool = *new CopyPool;

struct PinBuf {   // pinned host ring slot
    void *p = nullptr;
    size_t cap = 0;
    ~PinBuf() { if (p && thread_may_release()) { cudaFreeHost(p); (void)cudaGetLastError(); } }
    void *get(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr; cap = 0;
            if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); p = nullptr; return nullptr; }
            cap = bytes;
        }
        return p;
    }
};
thread_local PinBuf t_pin_in[2][2], t_pin_out[2][2];
thread_local DevBuf t_pipe_in[2][2], t_pipe_out[2][2];

bool is_pinned_host(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// Host arrays in, host arrays out, several fields: the fields are cut into chunks of levels and the three legs
// (host->device copy, interpolation, device->host copy) run on three streams, so that the PCIe link carries
// inputs of chunk k+1 and outputs of chunk k-1 (full duplex) while chunk k is being interpolated.  One or two input
// arrays and one or two output arrays (c_ezsint / c_ezuvint).  compute(d_in0, d_in1, d_out0, d_out1, nf) enqueues the
// interpolation of nf fields on S.stream and returns the reference's return code.
struct PipeSpec {
    int narr;                    // 1 (scalar) or 2 (u, v)
    const float *in[2];
    float *out[2];
    size_t nin, nout;            // floats per field
    int nfields;
};
template <class F>
int host_pipeline(const PipeSpec &p, F compute) {
    if (ensure_pipe() < 0) return -1;
    const int nfields = p.nfields;
    const int chunk = nfields >= 16 ? 4 : (nfields >= 4 ? 2 : 1);
    bool in_pg = false, out_pg = false;
    for (int a = 0; a < p.narr; a++) { in_pg |= !is_pinned_host(p.in[a]); out_pg |= !is_pinned_host(p.out[a]); }
    static const bool ring_on = !(getenv("EZB200_PINNED_RING") && atoi(getenv("EZB200_PINNED_RING")) == 0);
    if (!ring_on) in_pg = out_pg = false;
    float *d_in[2][2], *d_out[2][2], *h_in[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}, *h_out[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    for (int a = 0; a < p.narr; a++)
        for (int k = 0; k < 2; k++) {
            d_in[a][k] = (float *)t_pipe_in[a][k].get(p.nin * chunk * sizeof(float));
            d_out[a][k] = (float *)t_pipe_out[a][k].get(p.nout * chunk * sizeof(float));
            if (!d_in[a][k] || !d_out[a][k]) { fprintf(stderr, "<ezb200> cannot allocate device staging for the host-pointer pipeline\n"); return -1; }
            if (in_pg) h_in[a][k] = (float *)t_pin_in[a][k].get(p.nin * chunk * sizeof(float));
            if (out_pg) h_out[a][k] = (float *)t_pin_out[a][k].get(p.nout * chunk * sizeof(float));
            if ((in_pg && !h_in[a][k]) || (out_pg && !h_out[a][k])) { in_pg = out_pg = false; }   // no pinned memory: plain copies
        }
    // the copy streams must not run ahead of work already queued on the caller's stream
    CUDA_OK(cudaEventRecord(S.ev_comp[0], S.stream));
    CUDA_OK(cudaStreamWaitEvent(S.s_h2d, S.ev_comp[0], 0));
    int rc = 0, c = 0;
    int prev_f0 = 0, prev_nf = 0;
    auto fail = [&](int r) { Unlocked u; cudaStreamSynchronize(S.s_h2d); cudaStreamSynchronize(S.stream); cudaStreamSynchronize(S.s_d2h); return r; };
    auto drain = [&](int b, int f0, int nf) -> int {   // pageable outputs: ring slot b -> the caller's arrays
        Unlocked u;
        CUDA_OK(cudaEventSynchronize(S.ev_out[b]));
        for (int a = 0; a < p.narr; a++) g_copy_pool.copy(p.out[a] + (size_t)f0 * p.nout, h_out[a][b], p.nout * nf * sizeof(float));
        return 0;
    };
    for (int f0 = 0; f0 < nfields; f0 += chunk, c++) {
        const int nf = (nfields - f0 < chunk) ? (nfields - f0) : chunk, b = c & 1;
        {
            Unlocked u;   // nothing below touches the registry; copies from pageable memory block
            if (in_pg) {
                if (c >= 2) CUDA_OK(cudaEventSynchronize(S.ev_in[b]));     // the DMA that read this ring slot is done
                for (int a = 0; a < p.narr; a++) g_copy_pool.copy(h_in[a][b], p.in[a] + (size_t)f0 * p.nin, p.nin * nf * sizeof(float));
            }
            if (c >= 2) CUDA_OK(cudaStreamWaitEvent(S.s_h2d, S.ev_comp[b], 0));      // the kernel that read this input buffer is done
            for (int a = 0; a < p.narr; a++)
                CUDA_OK(cudaMemcpyAsync(d_in[a][b], in_pg ? h_in[a][b] : p.in[a] + (size_t)f0 * p.nin, p.nin * nf * sizeof(float), cudaMemcpyHostToDevice, S.s_h2d));
        }
        CUDA_OK(cudaEventRecord(S.ev_in[b], S.s_h2d));
        CUDA_OK(cudaStreamWaitEvent(S.stream, S.ev_in[b], 0));
        if (c >= 2) CUDA_OK(cudaStreamWaitEvent(S.stream, S.ev_out[b], 0));      // the copy that drained this output buffer is done
        int r = compute(d_in[0][b], d_in[1][b], d_out[0][b], d_out[1][b], nf);
        if (r < 0) return fail(r);
        if (r > rc) rc = r;
        CUDA_OK(cudaEventRecord(S.ev_comp[b], S.stream));
        CUDA_OK(cudaStreamWaitEvent(S.s_d2h, S.ev_comp[b], 0));
        {
            Unlocked u;
            for (int a = 0; a < p.narr; a++)
                CUDA_OK(cudaMemcpyAsync(out_pg ? h_out[a][b] : p.out[a] + (size_t)f0 * p.nout, d_out[a][b], p.nout * nf * sizeof(float), cudaMemcpyDeviceToHost, S.s_d2h));
        }
        CUDA_OK(cudaEventRecord(S.ev_out[b], S.s_d2h));
        if (out_pg && c >= 1 && drain(b ^ 1, prev_f0, prev_nf) < 0) return fail(-1);   // the previous chunk leaves the ring while this one is in flight
        prev_f0 = f0; prev_nf = nf;
    }
    if (out_pg && c >= 1 && drain((c - 1) & 1, prev_f0, prev_nf) < 0) return fail(-1);
    Unlocked u;
    CUDA_OK(cudaStreamSynchronize(S.s_d2h));
    CUDA_OK(cudaStreamSynchronize(S.stream));
    return rc;
}

int ezsint_pipelined(Grid &out, Grid &in, float *zout, const float *zin, int nfields, size_t nin, size_t nout) {
    PipeSpec p;
    p.narr = 1; p.in[0] = zin; p.in[1] = nullptr; p.out[0] = zout; p.out[1] = nullptr;
    p.nin = nin; p.nout = nout; p.nfields = nfields;
    return host_pipeline(p, [&](float *di, float *, float *dout, float *, int nf) {
        return sint_any_dev(out, in, dout, di, nf, (long)nin, (long)nout, 0, 1);
    });
}

// ---------------------------------------------------------------------------------------------------
// one process, several GPUs: a host-pointer call over many fields is cut into contiguous blocks of fields, one per device
// ---------------------------------------------------------------------------------------------------
// BASELINE.json north_star: "multi-field and multi-level calls are partitioned across the 8 GPUs of one box by level ... each GPU
// holding a replicated gset".  One persistent worker thread per extra device (its thread-local Session owns that device's
// streams, staging buffers and pinned ring); the calling thread keeps the block of its own current device.  Every block runs
// the ordinary single-device path (the level-chunk pipeline above), so each device's PCIe link carries its own share.
class DeviceWorkers {
    struct Worker {
        int dev = 0;
        std::thread th;
        std::mutex m;
        std::condition_variable cv;
        std::function<int()> job;
        bool has_job = false, done = true;
        int rc = 0;
    };
    std::vector<std::unique_ptr<Worker>> w;   // index = CUDA device ordinal
    std::mutex m;

    static void run(Worker *k) {
        cudaSetDevice(k->dev);
        for (;;) {
            std::function<int()> job;
            {
                std::unique_lock<std::mutex> lk(k->m);
                k->cv.wait(lk, [&] { return k->has_job; });
                job = std::move(k->job);
                k->has_job = false;
            }
            const int rc = job();
            {
                std::lock_guard<std::mutex> lk(k->m);
                k->rc = rc;
                k->done = true;
            }
            k->cv.notify_all();
        }
    }

public:
    void submit(int dev, std::function<int()> job) {
        Worker *k;
        {
            std::lock_guard<std::mutex> lk(m);
            if ((int)w.size() <= dev) w.resize(dev + 1);
            if (!w[dev]) {
                w[dev] = std::make_unique<Worker>();
                w[dev]->dev = dev;
                w[dev]->th = std::thread(run, w[dev].get());
                w[dev]->th.detach();
            }
            k = w[dev].get();
        }
        {
            std::lock_guard<std::mutex> lk(k->m);
            k->job = std::move(job);
            k->has_job = true;
            k->done = false;
        }
        k->cv.notify_all();
    }
    int wait(int dev) {
        Worker *k;
        {
            std::lock_guard<std::mutex> lk(m);
            k = w[dev].get();
        }
        std::unique_lock<std::mutex> lk(k->m);
        k->cv.wait(lk, [&] { return k->done; });
        return k->rc;
    }
};
DeviceWorkers &g_workers = *new DeviceWorkers;   // never destroyed (threads wait on its condition variables until the process ends)

int visible_devices() {
    static int n = -1;
    if (n < 0) {
        if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); n = 1; }
        if (n > Grid::MAX_DEV) n = Grid::MAX_DEV;
        if (n < 1) n = 1;
    }
    return n;
}

int ezsint_n(float *zout, float *zin, int nfields);
int ezuvint_n(float *uuout, float *vvout, float *uuin, float *vvin, int nfields, int mode);

// Returns -2 when the call is not fanned out (the caller then runs the single-device path), else the reference's return code.
// block(f0, nf) runs fields [f0, f0 + nf) on the CURRENT thread and device.
// the field count of the outermost API call in progress (see sint_dev: kernel choice) -- nested entries (device blocks) keep it
struct CallFields {
    bool owner;
    explicit CallFields(int n) : owner(S.call_fields == 0) { if (owner) S.call_fields = n; }
    ~CallFields() { if (owner) S.call_fields = 0; }
};

int fan_out(int nfields, bool all_host, const std::function<int(int, int)> &block) {
    const int ndev = std::min(S.devices, visible_devices());
    if (ndev <= 1 || !all_host || nfields < 2 * ndev || S.row0 >= 0) return -2;
    // this thread's session, as the workers must see it
    const int gdin = S.gdin, gdout = S.gdout, arith = S.arith, call_fields = S.call_fields, direct_max = S.direct_max;
    const Options opt = S.opt;
    const int dev0 = t_dev;
    std::vector<int> devs;
    devs.push_back(dev0);
    for (int d = 0; d < visible_devices() && (int)devs.size() < ndev; d++)
        if (d != dev0) devs.push_back(d);
    Unlocked u;   // the blocks take the registry lock themselves
    std::vector<std::pair<int, int>> span(devs.size());
    for (size_t k = 0; k < devs.size(); k++) {
        const int base = nfields / (int)devs.size(), rem = nfields % (int)devs.size();
        const int lo = (int)k * base + std::min((int)k, rem);
        span[k] = {lo, base + ((int)k < rem ? 1 : 0)};
    }
    for (size_t k = 1; k < devs.size(); k++) {
        const int lo = span[k].first, nf = span[k].second;
        g_workers.submit(devs[k], [=, &block]() -> int {
            S.gdin = gdin; S.gdout = gdout; S.arith = arith; S.opt = opt; S.devices = 1; S.row0 = S.row1 = -1;
            S.call_fields = call_fields; S.direct_max = direct_max;
            const int r = block(lo, nf);
            S.call_fields = 0;
            return r;
        });
    }
    const int saved = S.devices;
    S.devices = 1;
    int rc = block(span[0].first, span[0].second);
    S.devices = saved;
    for (size_t k = 1; k < devs.size(); k++) {
        const int r = g_workers.wait(devs[k]);
        if (r < 0 || rc < 0) rc = -1;
        else if (r > rc) rc = r;
    }
    return rc;
}

int ezsint_n(float *zout, float *zin, int nfields) {
    Lock lk;
    if (!session_ok("c_ezsint")) return -1;
    if (nfields < 1) return -1;
    CallFields cf(nfields);
    Grid &in = *grid_of(S.gdin), &out = *grid_of(S.gdout);
    const size_t nin = (size_t)in.ni * in.nj, nout = out_points(out);
    if (S.gdin == S.gdout) {   // ezsint.c:56-60
        Unlocked u;
        CUDA_OK(cudaMemcpyAsync(zout, zin, nin * nfields * sizeof(float), cudaMemcpyDefault, S.stream));
        CUDA_OK(cudaStreamSynchronize(S.stream));
        return 1;
    }
    const bool all_host = !is_device_ptr(zin) && !is_device_ptr(zout);
    const bool cloud_src = in.grtyp == 'Y';
    if (S.devices > 1 && !cloud_src) {
        const int r = fan_out(nfields, all_host, [&](int f0, int nf) { return ezsint_n(zout + (size_t)f0 * nout, zin + (size_t)f0 * nin, nf); });
        if (r != -2) return r;
    }
    HostCallStream hs(all_host);
    if (nfields >= 2 && all_host && !cloud_src) return ezsint_pipelined(out, in, zout, zin, nfields, nin, nout);
    Staged si, so;
    if (stage_in(si, S.st_in[0], zin, nin * nfields) < 0) return -1;
    if (stage_out(so, S.st_out[0], zout, nout * nfields) < 0) return -1;
    if (cloud_src && so.host && nout > nin) {
        // ez_applywgts leaves target points beyond the source's size untouched when they are masked out: they must read back
        // as the caller had them
        Unlocked u;
        CUDA_OK(cudaMemcpyAsync(so.dev, zout, nout * nfields * sizeof(float), cudaMemcpyHostToDevice, S.stream));
    }
    int rc = sint_any_dev(out, in, so.dev, si.dev, nfields, (long)nin, (long)nout, 0, 1);
    if (finish_out(so) < 0) return -1;
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// the vector path: c_ezuvint_orig (ezuvint.c:51-94) / c_ezyyuvint (ezyyuvint.c:28-319)
// ---------------------------------------------------------------------------------------------------
// everything of a wind-rotation launch but the fields: frames, target lat/lon, panel bits (band-shifted when a row band is set)
int wind_args(Grid &out, Grid &in, WindArgs &w, long &boff) {
    if (ensure_latlon(out) < 0) return -1;
    memset(&w, 0, sizeof(w));
    if (in.nsubgrids > 0) {
        Gset *gs = ensure_gset(out, in);
        if (!gs) return -1;
        w.src = to_dev(*grid_of(in.subgrid[0]));
        w.src2 = to_dev(*grid_of(in.subgrid[1]));
        w.yy = 1;
        w.zb = gs->d_zb;
    } else {
        w.src = to_dev(in);
    }
    w.dst = to_dev(out);
    const Band band = band_of(out);
    boff = band.on ? (long)band.j0 * out.ni : 0;   // output-row band: the per-point arrays of the target start at its first row
    w.lat = out.ds().d_lat + boff; w.lon = out.ds().d_lon + boff;
    if (w.zb) w.zb += boff;
    w.npts = out.ni * (band.j1 - band.j0);
    return 0;
}

// Cached per-point rotation of the (u,v) -> (u,v) conversion between the frames of `in` and `out` (ez_wind.cu), whole target,
// built on first use; nullptr when it does not apply (arith 0 keeps the reference's per-call evaluation) or cannot be built.
const float4 *wind_rotation(Grid &out, Grid &in, const WindArgs &w, long boff) {
    if (S.arith < 1 || in.id == out.id) return nullptr;
    Gset *gs = ensure_gset(out, in);
    if (!gs) return nullptr;
    if (!gs->wrot_tried) {
        gs->wrot_tried = true;
        WindArgs wf = w;   // the cached rotation always covers the whole target
        wf.lat = out.ds().d_lat; wf.lon = out.ds().d_lon;
        if (wf.zb) wf.zb -= boff;
        wf.npts = out.ni * out.nj;
        if (cudaMalloc((void **)&gs->d_wrot, (size_t)wf.npts * sizeof(float4)) == cudaSuccess) {
            if (!ezk::wind_matrix(wf, gs->d_wrot, S.stream)) { cudaFree(gs->d_wrot); gs->d_wrot = nullptr; }
            S.launches += 3;
        } else {
            gs->d_wrot = nullptr;
            (void)cudaGetLastError();
        }
    }
    return gs->d_wrot;
}

int wind_dev(Grid &out, Grid &in, float *d_u, float *d_v, int nfields, long stride, int mode) {
    WindArgs w;
    long boff = 0;
    if (wind_args(out, in, w, boff) < 0) return -1;
    w.u = d_u; w.v = d_v;
    w.stride = stride;
    w.nfields = nfields;
    w.mode = mode;
    if (mode == 0) {
        const float4 *m = wind_rotation(out, in, w, boff);
        if (m) {
            ezk::wind_apply(m + boff, d_u, d_v, stride, nfields, w.npts, S.stream);
            S.launches++;
            CUDA_OK(cudaGetLastError());
            return 0;
        }
    }
    ezk::wind(w, S.stream);
    S.launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// uv on device pointers for one plain target (a sub-grid when the caller's target is 'U')
int uvint_dev(Grid &out, Grid &in, float *d_uo, float *d_vo, const float *d_ui, const float *d_vi, int nfields, long in_stride, long out_stride, int mode) {
    const bool yyin = in.nsubgrids > 0;
    Grid *eff_in = &in;
    const float *ui = d_ui, *vi = d_vi;
    if (yyin) {
        Grid *yin_in = grid_of(in.subgrid[0]), *yan_in = grid_of(in.subgrid[1]);
        const long half_in = (long)yin_in->ni * yin_in->nj;
        if (yin_in->id == out.id) { eff_in = yin_in; }
        else if (yan_in->id == out.id) { eff_in = yan_in; ui += half_in; vi += half_in; }
        else if (S.opt.use_1subgrid == 1) {
            if (S.opt.valeur_1subgrid == yin_in->id) eff_in = yin_in;
            else if (S.opt.valeur_1subgrid == yan_in->id) { eff_in = yan_in; ui += half_in; vi += half_in; }
            else {
                fprintf(stderr, "<c_ezyyuvint> define src subgridid in ezsetival(subgridid)! Aborting...\n");
                return -1;
            }
        }
    }
    int r1, r2;
    bool wind_done = false;
    if (eff_in->id == out.id) {   // c_ezsint's gdin == gdout copy, then the wind conversions still run
        for (int f = 0; f < nfields; f++) {
            cudaMemcpyAsync(d_uo + f * out_stride, ui + f * in_stride, (size_t)out.ni * out.nj * sizeof(float), cudaMemcpyDeviceToDevice, S.stream);
            cudaMemcpyAsync(d_vo + f * out_stride, vi + f * in_stride, (size_t)out.ni * out.nj * sizeof(float), cudaMemcpyDeviceToDevice, S.stream);
        }
        r1 = r2 = 1;
    } else {
        // ez_corrvec (ezuvint.c:77-79): points next to a pole are re-interpolated from strips whose pole row is synthetic
        const float *prow_u = nullptr, *prow_v = nullptr;
        if (eff_in->nsubgrids == 0 && S.opt.polar_correction == 1) {
            Gset *gs = ensure_gset(out, *eff_in);
            if (!gs) return -1;
            if ((gs->counts[1] + gs->counts[2] + gs->counts[3] + gs->counts[4]) > 0) {
                {
                    const size_t nrow = (size_t)nfields * 2 * eff_in->ni;
                    float *pu = (float *)S.prow[0].get(nrow * sizeof(float)), *pv = (float *)S.prow[1].get(nrow * sizeof(float));
                    if (!pu || !pv) return -1;
                    PolarWindArgs pw;
                    memset(&pw, 0, sizeof(pw));
                    pw.src = to_dev(*eff_in);
                    pw.u = ui; pw.v = vi;
                    pw.stride = in_stride;
                    pw.nfields = nfields;
                    // auxiliary polar-stereographic grids: cxgaig(pi 0, pj 0, d60 1000, dgrw 0) then the decode c_ezqkdef performs
                    for (int h = 0; h < 2; h++) {
                        float xg[4] = {0.0f, 0.0f, 1000.0f, 0.0f};
                        int ig[4];
                        ezh::cxgaig(h ? 'S' : 'N', ig, xg);
                        ezh::cigaxg(h ? 'S' : 'N', xg, ig);
                        (h ? pw.dgrw_s : pw.dgrw_n) = xg[3];
                    }
                    pw.prow_u = pu; pw.prow_v = pv;
                    ezk::polar_wind_rows(pw, S.stream);
                    S.launches++;
                    prow_u = pu; prow_v = pv;
                }
            }
        }
        const long uv_in = (long)(vi - ui), uv_out = (long)(d_vo - d_uo);
        const bool uv_near = (uv_in < 0 ? -uv_in : uv_in) < (1L << 26) && (uv_out < 0 ? -uv_out : uv_out) < (1L << 40);
        // (the tile kernels address a stage through 32-bit offsets, hence uv_near; the direct kernel, which takes a pair unless it
        //  is switched off, has no such limit -- callers' u and v arrays may lie anywhere)
        const bool direct_likely = (S.call_fields > 0 ? S.call_fields : 2) <= S.direct_max && S.direct_max >= 2 && S.arith == 2 &&
                                   (S.opt.degre_interp == EZ_LINEAR || S.opt.degre_interp == EZ_CUBIC) && out.grtyp != 'Y';
        if (nfields == 1 && !prow_u && !prow_v && !eff_in->needs_expansion && (uv_near || direct_likely)) {
            // one (u, v) pair, no synthetic pole rows, no hemispheric mirror (the only place the two components are treated
            // differently): ONE launch over two "fields" whose stride is the distance between the caller's arrays, so the
            // gridset, the plans and the tile setup are read and built once for both components
            // (and, on the direct kernel, the cached wind rotation applied to the pair before it is stored)
            const float4 *m = nullptr;
            if (mode == 0) {
                WindArgs w;
                long boff = 0;
                if (wind_args(out, *eff_in, w, boff) < 0) return -1;
                m = wind_rotation(out, *eff_in, w, boff);
            }
            r1 = r2 = sint_dev(out, *eff_in, d_uo, ui, 2, (long)(vi - ui), (long)(d_vo - d_uo), 1, +1, nullptr, m, &wind_done);
        } else {
            r1 = sint_dev(out, *eff_in, d_uo, ui, nfields, in_stride, out_stride, 1, +1, prow_u);
            r2 = sint_dev(out, *eff_in, d_vo, vi, nfields, in_stride, out_stride, 1, -1, prow_v);
        }
        if (r1 < 0 || r2 < 0) return -1;
    }
    if (!wind_done && wind_dev(out, *eff_in, d_uo, d_vo, nfields, out_stride, mode) < 0) return -1;
    return (r1 == 2 || r2 == 2) ? 2 : 0;
}

int ezuvint_n(float *uuout, float *vvout, float *uuin, float *vvin, int nfields, int mode) {
    Lock lk;
    if (!session_ok("c_ezuvint")) return -1;
    if (nfields < 1 || check_degree(true) < 0) return -1;
    CallFields cf(2 * nfields);
    Grid &in = *grid_of(S.gdin), &out = *grid_of(S.gdout);
    const size_t nin = (size_t)in.ni * in.nj, nout = out_points(out);
    if (out.nsubgrids > 0 && in.nsubgrids > 0 && S.opt.use_1subgrid == 1) {
        fprintf(stderr, "<c_ezyyuvint> cannot use 1 subgrid to interpolate to a Yin-Yang grid  Aborting...\n");
        return -1;
    }
    // device part: one plain target, or the two sub-grids of a 'U' target (ezyyuvint.c:28-319)
    auto on_device = [&](float *d_uo, float *d_vo, const float *d_ui, const float *d_vi, int nf) -> int {
        if (out.nsubgrids > 0) {
            Grid *yin_out = grid_of(out.subgrid[0]), *yan_out = grid_of(out.subgrid[1]);
            if (!yin_out || !yan_out) return -1;
            const long half = (long)yin_out->ni * yin_out->nj;
            int r1 = uvint_dev(*yin_out, in, d_uo, d_vo, d_ui, d_vi, nf, (long)nin, (long)nout, mode);
            int r2 = uvint_dev(*yan_out, in, d_uo + half, d_vo + half, d_ui, d_vi, nf, (long)nin, (long)nout, mode);
            return (r1 < 0 || r2 < 0) ? -1 : ((r1 == 2 || r2 == 2) ? 2 : 0);
        }
        return uvint_dev(out, in, d_uo, d_vo, d_ui, d_vi, nf, (long)nin, (long)nout, mode);
    };
    const bool all_host = !is_device_ptr(uuin) && !is_device_ptr(vvin) && !is_device_ptr(uuout) && !is_device_ptr(vvout);
    if (S.devices > 1 && in.grtyp != 'Y') {
        const int r = fan_out(nfields, all_host, [&](int f0, int nf) {
            return ezuvint_n(uuout + (size_t)f0 * nout, vvout + (size_t)f0 * nout, uuin + (size_t)f0 * nin, vvin + (size_t)f0 * nin, nf, mode);
        });
        if (r != -2) return r;
    }
    HostCallStream hs(all_host);
    if (all_host && nfields >= 2 && in.grtyp != 'Y') {   // the level-chunk pipeline of the scalar path, two arrays each way
        PipeSpec p;
        p.narr = 2; p.in[0] = uuin; p.in[1] = vvin; p.out[0] = uuout; p.out[1] = vvout;
        p.nin = nin; p.nout = nout; p.nfields = nfields;
        return host_pipeline(p, [&](float *du, float *dv, float *duo, float *dvo, int nf) { return on_device(duo, dvo, du, dv, nf); });
    }
    Staged su, sv, ou, ov;
    if (stage_in(su, S.st_in[0], uuin, nin * nfields) < 0 || stage_in(sv, S.st_in[1], vvin, nin * nfields) < 0) return -1;
    if (stage_out(ou, S.st_out[0], uuout, nout * nfields) < 0 || stage_out(ov, S.st_out[1], vvout, nout * nfields) < 0) return -1;
    if (in.grtyp == 'Y' && nout > nin) {   // cloud sources leave masked-out target points beyond the source's size as the caller had them
        Unlocked u;
        if (ou.host) CUDA_OK(cudaMemcpyAsync(ou.dev, uuout, nout * nfields * sizeof(float), cudaMemcpyHostToDevice, S.stream));
        if (ov.host) CUDA_OK(cudaMemcpyAsync(ov.dev, vvout, nout * nfields * sizeof(float), cudaMemcpyHostToDevice, S.stream));
    }
    int rc = on_device(ou.dev, ov.dev, su.dev, sv.dev, nfields);
    if (rc < 0) return rc;
    if (finish_out(ou) < 0 || finish_out(ov) < 0) return -1;
    return rc;
}

void lower(char *dst, const char *src) {
    size_t n = strlen(src);
    if (n > 31) n = 31;
    for (size_t i = 0; i < n; i++) dst[i] = (char)tolower((int)src[i]);
    dst[n] = 0;
}

unsigned int ftn_strlen(const char *str, uint32_t maxLen) {   // src/base/ftnStrLen.c:5-19
    if (maxLen == 0) return 0;
    uint32_t len = maxLen - 1;
    while (len > 0 && (str[len] == ' ' || str[len] == '\0')) len--;
    if (str[len] != ' ' && str[len] != '\0') len++;
    return len;
}
void ftn_copy(char dst[32], const char *src, int32_t l) {
    int n = (int)ftn_strlen(src, (uint32_t)l);
    if (n > 31) n = 31;
    memcpy(dst, src, n);
    dst[n] = 0;
}

}  // namespace

// ===================================================================================================
// C ABI
// ===================================================================================================
extern "C" {

int32_t c_ezb200_set_devices(int32_t n) {
    if (n < 1) n = 1;
    S.devices = std::min(n, visible_devices());
    return S.devices;
}
int32_t c_ezb200_set_direct_max_fields(int32_t n) {
    const int prev = S.direct_max;
    S.direct_max = n < 0 ? 0 : n;
    return prev;
}
int32_t c_ezb200_set_rows(int32_t j0, int32_t j1) {
    if (j0 < 0 || j1 <= j0) { S.row0 = S.row1 = -1; return 0; }
    S.row0 = j0; S.row1 = j1;
    return 0;
}
const char *c_ezb200_version(void) { return "ezb200 0.1 (sm_100a)"; }

int32_t c_ezb200_set_stream(void *stream) {
    S.stream = (cudaStream_t)stream;
    S.stream_user = true;
    return 0;
}
int32_t c_ezb200_set_arith(int32_t mode) {
    if (mode < 0 || mode > 2) return -1;
    S.arith = mode;
    return 0;
}
int32_t c_ezb200_get_arith(void) { return S.arith; }
int64_t c_ezb200_launch_count(int32_t reset) {
    int64_t v = S.launches;
    if (reset) S.launches = 0;
    return v;
}
int32_t c_ezb200_last_path(int32_t *nlist) {
    if (nlist) *nlist = S.last_nlist;
    return S.last_path;
}
int32_t c_ezb200_set_timing(int32_t on) {
    S.timing = on != 0;
    S.timing_valid = false;
    if (S.timing && !S.tev.ev[0])
        for (int k = 0; k < 4; k++) CUDA_OK(cudaEventCreate(&S.tev.ev[k]));
    return 0;
}
int32_t c_ezb200_last_kernel_ms(float *ms, int32_t *tiles) {
    if (tiles) { tiles[0] = S.last_tiles[0]; tiles[1] = S.last_tiles[1]; tiles[2] = S.last_nlist; }
    if (!ms) return 0;
    ms[0] = ms[1] = ms[2] = 0.f;
    if (!S.timing_valid) return -1;
    CUDA_OK(cudaEventSynchronize(S.tev.ev[3]));
    for (int k = 0; k < 3; k++) CUDA_OK(cudaEventElapsedTime(&ms[k], S.tev.ev[k], S.tev.ev[k + 1]));
    return 0;
}
int32_t c_ezb200_sync(void) {
    CUDA_OK(cudaStreamSynchronize(S.stream));
    return 0;
}

// ---- descriptors ----
void cigaxg_(char *grtyp, float *xg1, float *xg2, float *xg3, float *xg4, const int32_t *ig1, const int32_t *ig2, const int32_t *ig3,
             const int32_t *ig4, F2Cl l) {
    (void)l;
    int ig[4] = {*ig1, *ig2, *ig3, *ig4};
    float xg[4] = {*xg1, *xg2, *xg3, *xg4};
    if (!ezh::cigaxg(grtyp[0], xg, ig)) fprintf(stderr, "cigaxg: Bad grid pecification (TYPE)\n");
    *xg1 = xg[0]; *xg2 = xg[1]; *xg3 = xg[2]; *xg4 = xg[3];
}
void cxgaig_(char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, float *xg1, float *xg2, float *xg3, float *xg4, F2Cl l) {
    (void)l;
    int ig[4] = {0, 0, 0, 0};
    float xg[4] = {*xg1, *xg2, *xg3, *xg4};
    if (!ezh::cxgaig(grtyp[0], ig, xg)) fprintf(stderr, "cxgaig: Unknown grid (TYPE)\n");
    *ig1 = ig[0]; *ig2 = ig[1]; *ig3 = ig[2]; *ig4 = ig[3];
}

// ---- grid definition ----
int32_t c_ezgdef_fmem(int32_t ni, int32_t nj, char *grtyp, char *grref, int32_t ig1, int32_t ig2, int32_t ig3, int32_t ig4, float *ax, float *ay) {
    return gdef_fmem(ni, nj, grtyp, grref, ig1, ig2, ig3, ig4, ax, ay);
}
int32_t ezgdef_fmem_(int32_t *ni, int32_t *nj, char *grtyp, char *grref, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, float *ax,
                     float *ay, F2Cl l1, F2Cl l2) {
    (void)l1; (void)l2;
    char t[2] = {grtyp[0], 0}, r[2] = {grref[0], 0};
    return c_ezgdef_fmem(*ni, *nj, t, r, *ig1, *ig2, *ig3, *ig4, ax, ay);
}
// ezgdef_fll.c:34-49: a 'Y' cloud of (lat, lon) points on an 'L' reference with lat0 = lon0 = 0, dlat = dlon = 1
int32_t c_ezgdef_fll(int32_t ni, int32_t nj, float *lat, float *lon) {
    int ig[4];
    float xg[4] = {0.0f, 0.0f, 1.0f, 1.0f};
    ezh::cxgaig('L', ig, xg);
    char t[2] = {'Y', 0}, r[2] = {'L', 0};
    return c_ezgdef_fmem(ni, nj, t, r, ig[0], ig[1], ig[2], ig[3], lon, lat);
}
int32_t ezgdef_fll_(int32_t *ni, int32_t *nj, float *lat, float *lon) { return c_ezgdef_fll(*ni, *nj, lat, lon); }

// gdgaxes.c:31-57: the positional axes the library copied at definition time (Z, #, G, Y grids)
int32_t c_gdgaxes(int32_t gdid, float *ax, float *ay) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g) return -1;
    if (g->ax.empty() || g->ay.empty()) {
        fprintf(stderr, "(gdgaxes) Erreur! A l'aide! Descripteurs manquants!\n");
        return -1;
    }
    const size_t nix = g->grtyp == 'Y' ? (size_t)g->ni * g->nj : (size_t)g->ni;
    const size_t njy = g->grtyp == 'Y' ? (size_t)g->ni * g->nj : (size_t)g->nj;
    memcpy(ax, g->ax.data(), std::min(nix, g->ax.size()) * sizeof(float));
    memcpy(ay, g->ay.data(), std::min(njy, g->ay.size()) * sizeof(float));
    return 0;
}
int32_t gdgaxes_(int32_t *gdid, float *ax, float *ay) { c_gdgaxes(*gdid, ax, ay); return 0; }

// gdxpncf.c:33-48: logical (expanded) index range of a grid
int32_t c_gdxpncf(int32_t gdin, int32_t *i1, int32_t *i2, int32_t *j1, int32_t *j2) {
    Lock lk;
    Grid *g = grid_of(gdin);
    if (!g) return -1;
    if (g->nsubgrids > 0) {
        fprintf(stderr, "<gdxpncf> This operation is not supported for 'U' grids.\n");
        return -1;
    }
    *i1 = g->i1; *i2 = g->i2; *j1 = g->j1; *j2 = g->j2;
    return 0;
}
int32_t gdxpncf_(int32_t *gdin, int32_t *i1, int32_t *i2, int32_t *j1, int32_t *j2) { return c_gdxpncf(*gdin, i1, i2, j1, j2); }

// ---------------------------------------------------------------------------------------------------
// N4: grids defined by positional records of an RPN standard file (ezgdef_ffile.c:29-140, ez_lireEnrPositionnels.c:12-349)
// ---------------------------------------------------------------------------------------------------
// The records are read through the caller's own FST layer: a drop-in replacement of the interpolator lives next to the librmn
// that opened the file, so c_fstinf / c_fstsui / c_fstprm / c_fstluk are looked up in the process at the first
// file-backed definition (dlsym), or handed over with c_ezb200_set_fst_hooks.  This library contains no file format code.
typedef int (*fstinf_fn)(int iun, int *ni, int *nj, int *nk, int datev, const char *etiket, int ip1, int ip2, int ip3, const char *typvar,
                         const char *nomvar);
typedef int (*fstsui_fn)(int iun, int *ni, int *nj, int *nk);
typedef int (*fstprm_fn)(int handle, int *dateo, int *deet, int *npas, int *ni, int *nj, int *nk, int *nbits, int *datyp, int *ip1, int *ip2,
                         int *ip3, char *typvar, char *nomvar, char *etiket, char *grtyp, int *ig1, int *ig2, int *ig3, int *ig4, int *swa,
                         int *lng, int *dltf, int *ubc, int *extra1, int *extra2, int *extra3);
typedef int (*fstluk_fn)(void *field, int handle, int *ni, int *nj, int *nk);
static struct { fstinf_fn inf; fstsui_fn sui; fstprm_fn prm; fstluk_fn luk; } g_fst = {nullptr, nullptr, nullptr, nullptr};

static bool fst_ready() {
    if (g_fst.inf && g_fst.sui && g_fst.prm && g_fst.luk) return true;
    fstinf_fn a = (fstinf_fn)dlsym(RTLD_DEFAULT, "c_fstinf");
    fstsui_fn b = (fstsui_fn)dlsym(RTLD_DEFAULT, "c_fstsui");
    fstprm_fn c = (fstprm_fn)dlsym(RTLD_DEFAULT, "c_fstprm");
    fstluk_fn d = (fstluk_fn)dlsym(RTLD_DEFAULT, "c_fstluk");
    if (a && b && c && d) { g_fst.inf = a; g_fst.sui = b; g_fst.prm = c; g_fst.luk = d; return true; }
    return false;
}
int32_t c_ezb200_set_fst_hooks(void *fstinf, void *fstsui, void *fstprm, void *fstluk) {
    Lock lk;
    g_fst.inf = (fstinf_fn)fstinf; g_fst.sui = (fstsui_fn)fstsui; g_fst.prm = (fstprm_fn)fstprm; g_fst.luk = (fstluk_fn)fstluk;
    return (g_fst.inf && g_fst.sui && g_fst.prm && g_fst.luk) ? 0 : -1;
}

struct FstRec { int key, ni, nj, nk, ip1, ip2, ip3, ig[4]; char grtyp; };
static int fst_prm(int key, FstRec &r) {
    int dateo, deet, npas, nbits, bidon, swa, lng, dltf, ubc, e1, e2, e3;
    char typvar[8] = "  ", nomvar[8] = "    ", etiket[16] = "            ", grtyp[4] = " ";
    r.key = key;
    const int ier = g_fst.prm(key, &dateo, &deet, &npas, &r.ni, &r.nj, &r.nk, &nbits, &bidon, &r.ip1, &r.ip2, &r.ip3, typvar, nomvar, etiket, grtyp,
                              &r.ig[0], &r.ig[1], &r.ig[2], &r.ig[3], &swa, &lng, &dltf, &ubc, &e1, &e2, &e3);
    r.grtyp = grtyp[0];
    return ier;
}

// Lire_enrUvercode1, ez_lireEnrPositionnels.c:12-46: the '^>' record of a Yin-Yang grid
static int gdef_u_record(const float *yy, int nix, const int keys[4]) {
    if (nix < 16) return -1;
    const int ndiv = (int)yy[2], ni = (int)yy[5], nj = (int)yy[6];
    if (ndiv != 2 || ni <= 0 || nj <= 0 || nix < 2 * (15 + ni + nj) - 5) {
        fprintf(stderr, "<LireEnrPositionnels>: malformed ^> record (ndiv %d, ni %d, nj %d, length %d)\n", ndiv, ni, nj, nix);
        return -1;
    }
    const int yinsize = 15 + ni + nj;
    std::vector<float> ax(yy + 15, yy + 15 + ni), ay(yy + 15 + ni, yy + 15 + ni + nj);
    char z[2] = {'Z', 0}, e[2] = {'E', 0}, u[2] = {'U', 0}, f[2] = {'F', 0};
    int sub[2];
    for (int k = 0; k < 2; k++) {
        const float *rot = k ? yy + yinsize + 6 : yy + 11;
        float xg[4] = {rot[0], rot[1], rot[2], rot[3]};
        int ig[4];
        ezh::cxgaig('E', ig, xg);
        sub[k] = gdef_fmem(ni, nj, z, e, ig[0], ig[1], ig[2], ig[3], ax.data(), ay.data());
        if (sub[k] < 0) return -1;
    }
    const int gd = c_ezgdef_supergrid(ni, nj * ndiv, u, f, 1, 2, sub);
    if (gd >= 0 && keys) {
        Lock lk;
        Grid *g = grid_of(gd);
        if (g) for (int k = 0; k < 4; k++) g->ig[k] = keys[k];
    }
    return gd;
}

// c_ezgdef_ffile + LireEnrPositionnels for '#', 'Y', 'Z', 'U' and ' '
static int gdef_ffile(int ni, int nj, char t, int ig1, int ig2, int ig3, int ig4, int iunit) {
    if (!fst_ready()) {
        fprintf(stderr, "<c_ezgdef_ffile> grid type '%c' needs positional records from an RPN standard file, and this process exports no "
                        "c_fstinf / c_fstsui / c_fstprm / c_fstluk (link the librmn that opened unit %d, or call c_ezb200_set_fst_hooks)\n", t, iunit);
        return -1;
    }
    const char blank12[] = "            ", blank2[] = "  ";
    int n1, n2, n3;
    if (t == 'U') {
        int key = g_fst.inf(iunit, &n1, &n2, &n3, -1, blank12, ig1, ig2, ig3, blank2, "^>  ");
        FstRec r;
        bool found = false;
        while (!found && key > 0) {
            const int ier = fst_prm(key, r);
            if (r.ni > 0 && r.nj == 1 && ier >= 0) found = true;
            else key = g_fst.sui(iunit, &n1, &n2, &n3);
        }
        if (!found) { fprintf(stderr, "<LireEnrPositionnels>: Positional records ^> not found. Exiting...\n\n"); return -1; }
        if (r.grtyp != 'F') { fprintf(stderr, "<LireEnrPositionnels>: Unknown reference grid %c for grid type U. Exiting...\n", r.grtyp); return -1; }
        if (r.ig[0] != 1) { fprintf(stderr, "<LireEnrPositionnels>: ^> record with vercode %d: only vercode 1 is defined\n", r.ig[0]); return -1; }
        std::vector<float> yy((size_t)r.ni);
        if (g_fst.luk(yy.data(), r.key, &n1, &n2, &n3) < 0) return -1;
        const int keys[4] = {r.ip1, r.ip2, r.ip3, ig4};
        return gdef_u_record(yy.data(), r.ni, keys);
    }
    const int tmpip3 = (t == '#') ? -1 : ig3;
    // '>>'
    FstRec rx, ry;
    int key = g_fst.inf(iunit, &n1, &n2, &n3, -1, blank12, ig1, ig2, tmpip3, blank2, ">>  ");
    if ((ni == 0 || ni == -1) && key > 0) ni = n1;
    bool fx = false, fy = false;
    while (!fx && key > 0) {
        fst_prm(key, rx);
        if (rx.ni == ni || t == '#') fx = true;
        else key = g_fst.sui(iunit, &n1, &n2, &n3);
    }
    // '^^'
    key = g_fst.inf(iunit, &n1, &n2, &n3, -1, blank12, ig1, ig2, tmpip3, blank2, "^^  ");
    if ((nj == 0 || nj == -1) && key > 0) nj = n2;
    while (!fy && key > 0) {
        fst_prm(key, ry);
        if (ry.nj == nj || t == '#') fy = true;
        else key = g_fst.sui(iunit, &n1, &n2, &n3);
    }
    if (!fx || !fy) { fprintf(stderr, "<LireEnrPositionnels>: Positional records ^^ and >> not found. Exiting...\n\n"); return -1; }
    char gt = t;
    if (ry.ni == rx.ni && ry.nj == rx.nj) gt = 'Y';
    else if (t != '#') gt = 'Z';
    const char grref = ry.grtyp;   // (the check looks at the '^^' record's reference, the descriptors kept are the '>>' record's)
    if (grref != 'N' && grref != 'S' && grref != 'L' && grref != 'E' && grref != 'O') {
        fprintf(stderr, "<LireEnrPositionnels>: Unknown reference grid. Exiting...\n");
        return -1;
    }
    std::vector<float> ay((size_t)ry.ni * ry.nj), ax((size_t)rx.ni * rx.nj);
    if (g_fst.luk(ay.data(), ry.key, &n1, &n2, &n3) < 0 || g_fst.luk(ax.data(), rx.key, &n1, &n2, &n3) < 0) return -1;
    std::vector<float> gax, gay;
    if (gt == '#') {   // Lire_enrTicTac: the (ig3, ig4) window of the full axes
        if (ig3 < 1 || ig4 < 1 || (size_t)(ig3 - 1 + ni) > ax.size() || (size_t)(ig4 - 1 + nj) > ay.size()) {
            fprintf(stderr, "<LireEnrPositionnels>: '#' window (%d, %d) + %d x %d leaves the positional records\n", ig3, ig4, ni, nj);
            return -1;
        }
        gax.assign(ax.begin() + (ig3 - 1), ax.begin() + (ig3 - 1) + ni);
        gay.assign(ay.begin() + (ig4 - 1), ay.begin() + (ig4 - 1) + nj);
    } else {
        gax = ax; gay = ay;
    }
    if (grref == 'L')
        for (int i = 0; i < ni && i < (int)gax.size(); i++)
            if (gax[i] < 0.0f) gax[i] += 360.0f;
    const int keys[4] = {rx.ip1, rx.ip2, gt == '#' ? ig3 : rx.ip3, gt == '#' ? ig4 : 0};
    char tt[2] = {gt, 0}, rr[2] = {grref, 0};
    return gdef_fmem(ni, nj, tt, rr, rx.ig[0], rx.ig[1], rx.ig[2], rx.ig[3], gax.data(), gay.data(), keys);
}
int32_t c_ezgdef_ffile(int32_t ni, int32_t nj, char *grtyp, int32_t ig1, int32_t ig2, int32_t ig3, int32_t ig4, int32_t iunit) {
    const char t = grtyp[0];
    if (t == '#' || t == 'Y' || t == 'Z' || t == 'U' || t == ' ') return gdef_ffile(ni, nj, t, ig1, ig2, ig3, ig4, iunit);   // ezgdef_ffile.c:44-50
    char r[2] = {' ', 0};
    return c_ezgdef_fmem(ni, nj, grtyp, r, ig1, ig2, ig3, ig4, nullptr, nullptr);
}
int32_t ezgdef_ffile_(int32_t *ni, int32_t *nj, char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, int32_t *iunit, F2Cl l) {
    (void)l;
    char t[2] = {grtyp[0], 0};
    return c_ezgdef_ffile(*ni, *nj, t, *ig1, *ig2, *ig3, *ig4, *iunit);
}
// ezqkdef.c: c_ezqkdef is c_ezgdef_ffile
int32_t c_ezqkdef(int32_t ni, int32_t nj, char *grtyp, int32_t ig1, int32_t ig2, int32_t ig3, int32_t ig4, int32_t iunit) {
    return c_ezgdef_ffile(ni, nj, grtyp, ig1, ig2, ig3, ig4, iunit);
}
int32_t ezqkdef_(int32_t *ni, int32_t *nj, char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, int32_t *iunit, F2Cl l) {
    (void)l;
    char t[2] = {grtyp[0], 0};
    return c_ezqkdef(*ni, *nj, t, *ig1, *ig2, *ig3, *ig4, *iunit);
}
int32_t c_ezgdef(int32_t ni, int32_t nj, char *grtyp, char *grref, int32_t ig1, int32_t ig2, int32_t ig3, int32_t ig4, float *ax, float *ay) {
    if (grtyp[0] == '#') {
        fprintf(stderr, "The '#' grid type is not supported with ezgdef.\nPlease use ezgdef_ffile or ezgdef_fmem\n");
        return -1;
    }
    if ((grtyp[0] == 'Y' || grtyp[0] == 'Z') && grref && (0 == strcmp(grref, "FILE") || 0 == strcmp(grref, "file")))
        return c_ezgdef_ffile(ni, nj, grtyp, ig1, ig2, ig3, ig4, ig4);   // ezgdef.c:59-90 (the unit travels in ig4)
    return c_ezgdef_fmem(ni, nj, grtyp, grref, ig1, ig2, ig3, ig4, ax, ay);
}
int32_t ezgdef_(int32_t *ni, int32_t *nj, char *grtyp, char *grref, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, float *ax, float *ay,
                F2Cl l1, F2Cl l2) {
    (void)l1; (void)l2;
    char t[2] = {grtyp[0], 0}, r[2] = {grref[0], 0};
    return c_ezgdef(*ni, *nj, t, r, *ig1, *ig2, *ig3, *ig4, ax, ay);
}

int32_t c_ezgdef_supergrid(int32_t ni, int32_t nj, char *grtyp, char *grref, int32_t vercode, int32_t nsubgrids, int32_t *subgrid) {
    Lock lk;
    if (nsubgrids <= 1) {
        fprintf(stderr, "<c_ezgdef_supergrid> nsubgrids given is less than 2! Aborting...\n");
        return -1;
    }
    if (vercode != 1) {
        fprintf(stderr, "<c_ezgdef_supergrid> invalid vercode! Aborting...\n");
        return -1;
    }
    if (nsubgrids != 2 || !grid_of(subgrid[0]) || !grid_of(subgrid[1])) {
        fprintf(stderr, "<c_ezgdef_supergrid> exactly two defined sub-grids (Yin, Yang) are supported\n");
        return -1;
    }
    Grid *s0 = grid_of(subgrid[0]);
    // identity rule for 'U' grids (ez_findgrid.c:33-46): type, ref, ig of the first sub-grid
    for (auto &up : g_grids) {
        Grid *g = up.get();
        if (g && g->alive && g->nsubgrids > 0 && g->grtyp == grtyp[0] && g->grref == grref[0] && g->ni == ni && g->nj == nj &&
            memcmp(g->ig, s0->ig, sizeof(int) * 4) == 0 && g->subgrid[0] == subgrid[0] && g->subgrid[1] == subgrid[1]) {
            g->access_count++;
            return g->id;
        }
    }
    auto g = std::make_unique<Grid>();
    g->grtyp = grtyp[0];
    g->grref = grref[0];
    g->ni = ni;
    g->nj = nj;
    for (int k = 0; k < 4; k++) g->ig[k] = s0->ig[k];
    g->igref[0] = vercode;
    g->nsubgrids = nsubgrids;
    g->subgrid[0] = subgrid[0];
    g->subgrid[1] = subgrid[1];
    g->i1 = 1; g->i2 = ni; g->j1 = 1; g->j2 = nj;
    int id = add_grid(std::move(g));
    for (int k = 0; k < 2; k++) {
        gdef_yymask(*grid_of(subgrid[k]));
        grid_of(subgrid[k])->access_count++;   // the supergrid holds its sub-grids: c_gdrls on them alone cannot pull them from under it
    }
    return id;
}
int32_t ezgdef_supergrid_(int32_t *ni, int32_t *nj, char *grtyp, char *grref, int32_t *vercode, int32_t *nsubgrids, int32_t *subgrid, F2Cl l1,
                          F2Cl l2) {
    (void)l1; (void)l2;
    char t[2] = {grtyp[0], 0}, r[2] = {grref[0], 0};
    return c_ezgdef_supergrid(*ni, *nj, t, r, *vercode, *nsubgrids, subgrid);
}

int32_t c_gdrls(int32_t gdin) {   // gdrls.c:34-90
    Lock lk;
    Grid *g = grid_of(gdin);
    if (!g) return 0;
    if (g->access_count > 0) g->access_count--;
    if (g->access_count == 0) {
        for (int dvi = 0; dvi < Grid::MAX_DEV; dvi++) {   // every device replica (cudaFree takes pointers of any device)
            Grid::DevState &d = g->dv[dvi];
            for (auto &kv : d.gsets) free_gset(*kv.second);
            d.gsets.clear();
            for (auto &up : g_grids)   // gridsets that use this grid as their source
                if (up && up->alive) {
                    auto it = up->dv[dvi].gsets.find(gdin);
                    if (it != up->dv[dvi].gsets.end()) { free_gset(*it->second); up->dv[dvi].gsets.erase(it); }
                }
            float **ptrs[] = {&d.d_ax, &d.d_ay, &d.d_cx, &d.d_cy, &d.d_lat, &d.d_lon};
            for (auto p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
            d.ready = false;
        }
        g->alive = false;
        // what the grid held: the sub-grids of a supergrid, the mask grid of a Yin / Yang panel
        for (int k = 0; k < g->nsubgrids && k < 2; k++)
            if (g->subgrid[k] >= 0) c_gdrls(g->subgrid[k]);
        if (g->mymaskgrid >= 0) { const int m = g->mymaskgrid; g->mymaskgrid = -1; c_gdrls(m); }
    }
    return 0;
}
int32_t gdrls_(int32_t *gdin) { return c_gdrls(*gdin); }

// ---- queries ----
int32_t c_ezgprm(int32_t gdid, char *grtyp, int32_t *ni, int32_t *nj, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g) return -1;
    *grtyp = g->grtyp; *ni = g->ni; *nj = g->nj;
    *ig1 = g->ig[0]; *ig2 = g->ig[1]; *ig3 = g->ig[2]; *ig4 = g->ig[3];
    return 0;
}
int32_t ezgprm_(int32_t *gdid, char *grtyp, int32_t *ni, int32_t *nj, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, F2Cl l) {
    (void)l;
    char t[2] = {' ', 0};
    int32_t rc = c_ezgprm(*gdid, t, ni, nj, ig1, ig2, ig3, ig4);
    grtyp[0] = t[0];
    return rc;
}
int32_t c_ezgxprm(int32_t gdid, int32_t *ni, int32_t *nj, char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, char *grref,
                  int32_t *ig1ref, int32_t *ig2ref, int32_t *ig3ref, int32_t *ig4ref) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g) return -1;
    *ni = g->ni; *nj = g->nj;
    grtyp[0] = g->grtyp; grtyp[1] = 0;
    grref[0] = g->grref; grref[1] = 0;
    *ig1 = g->ig[0]; *ig2 = g->ig[1]; *ig3 = g->ig[2]; *ig4 = g->ig[3];
    *ig1ref = g->igref[0]; *ig2ref = g->igref[1]; *ig3ref = g->igref[2]; *ig4ref = g->igref[3];
    return 0;
}
int32_t ezgxprm_(int32_t *gdid, int32_t *ni, int32_t *nj, char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, char *grref,
                 int32_t *ig1ref, int32_t *ig2ref, int32_t *ig3ref, int32_t *ig4ref, F2Cl l1, F2Cl l2) {
    (void)l1; (void)l2;
    char t[2] = {' ', 0}, r[2] = {' ', 0};
    int32_t rc = c_ezgxprm(*gdid, ni, nj, t, ig1, ig2, ig3, ig4, r, ig1ref, ig2ref, ig3ref, ig4ref);
    grtyp[0] = t[0];
    grref[0] = r[0];
    return rc;
}
int32_t c_ezget_nsubgrids(int32_t gdid) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g) return -1;
    return g->nsubgrids == 0 ? 1 : g->nsubgrids;
}
int32_t ezget_nsubgrids_(int32_t *gdid) { return c_ezget_nsubgrids(*gdid); }
int32_t c_ezget_subgridids(int32_t gdid, int32_t *subgrid) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g) return -1;
    if (g->nsubgrids == 0) { *subgrid = gdid; return 1; }
    for (int i = 0; i < g->nsubgrids; i++) subgrid[i] = g->subgrid[i];
    return g->nsubgrids;
}
int32_t ezget_subgridids_(int32_t *gdid, int32_t *subgrid) { return c_ezget_subgridids(*gdid, subgrid); }

int32_t c_gdll(int32_t gdid, float *lat, float *lon) {   // gdll.c:25-73
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g) return -1;
    if (g->nsubgrids > 0) {
        Grid *a = grid_of(g->subgrid[0]);
        size_t half = (size_t)a->ni * a->nj;
        if (c_gdll(g->subgrid[0], lat, lon) < 0) return -1;
        return c_gdll(g->subgrid[1], lat + half, lon + half);
    }
    if (ensure_latlon(*g) < 0) return -1;
    size_t n = (size_t)g->ni * g->nj;
    CUDA_OK(cudaMemcpyAsync(lat, g->ds().d_lat, n * sizeof(float), cudaMemcpyDefault, S.stream));
    CUDA_OK(cudaMemcpyAsync(lon, g->ds().d_lon, n * sizeof(float), cudaMemcpyDefault, S.stream));
    CUDA_OK(cudaStreamSynchronize(S.stream));
    return 0;
}
int32_t gdll_(int32_t *gdid, float *lat, float *lon) { return c_gdll(*gdid, lat, lon); }

// ---- scattered-point API: the same kernels on caller-supplied points -------------------------
namespace {
// stages up to 10 caller arrays (host or device) for one call
struct PtStage {
    Staged st[10];
    int n = 0;
    bool ok = true;
    float *in(const float *p, size_t cnt) {
        if (n >= 10 || stage_in(st[n], S.st_pts[n], p, cnt) < 0) { ok = false; return nullptr; }
        return st[n++].dev;
    }
    float *out(float *p, size_t cnt) {
        if (n >= 10 || stage_out(st[n], S.st_pts[n], p, cnt) < 0) { ok = false; return nullptr; }
        return st[n++].dev;
    }
    float *tmp(size_t cnt) {
        if (n >= 10) { ok = false; return nullptr; }
        float *d = (float *)S.st_pts[n].get(cnt * sizeof(float));
        if (!d) ok = false;
        n++;
        return d;
    }
    int finish() {
        for (int k = 0; k < n; k++)
            if (finish_out(st[k]) < 0) return -1;
        return 0;
    }
};

// (lat, lon) -> (x, y) on device pointers: c_gdxyfll (gdxyfll.c:190-243)
int xyfll_dev(Grid &g, float *d_x, float *d_y, const float *d_lat, const float *d_lon, int n) {
    if (g.grtyp == 'Y') {
        fprintf(stderr, "<c_gdxyfll> 'Y' grids are not supported by this entry point\n");
        return -1;
    }
    if (g.nsubgrids > 0) {
        Grid *yin = grid_of(g.subgrid[0]), *yan = grid_of(g.subgrid[1]);
        if (!yin || !yan || ensure_device(*yin) < 0 || ensure_device(*yan) < 0) return -1;
        const int myin[4] = {yin->mi0, yin->mi1, yin->mj0, yin->mj1}, myan[4] = {yan->mi0, yan->mi1, yan->mj0, yan->mj1};
        ezk::ll2xy_u(to_dev(*yin), to_dev(*yan), myin, myan, d_lat, d_lon, d_x, d_y, n, S.stream);
    } else {
        if (ensure_device(g) < 0) return -1;
        ezk::ll2xy_points(to_dev(g), g.j2, d_lat, d_lon, d_x, d_y, n, S.stream);
    }
    S.launches++;
    return 0;
}

// c_gdxysint (gdxysint.c:30-54) on device pointers, one field; 'U' sources are split by y (gdxysval.c:50-100)
int xysint_dev(Grid &g, float *d_zout, const float *d_zin, const float *d_x, const float *d_y, int n, int vector, int sym_sign) {
    if (check_degree() < 0) return -1;
    if (g.grtyp == 'Y') {
        fprintf(stderr, "<c_gdxysval> 'Y' cloud sources are outside the hot-path scope (SURVEY.md 8f N2)\n");
        return -1;
    }
    InterpArgs a;
    memset(&a, 0, sizeof(a));
    a.zin = d_zin; a.zout = d_zout;
    a.in_stride = (long)g.ni * g.nj; a.out_stride = n;
    a.nfields = 1; a.npts = n;
    a.o = snapshot_opts(vector, sym_sign);
    a.x = d_x; a.y = d_y;
    if (g.nsubgrids > 0) {
        Grid *yin = grid_of(g.subgrid[0]), *yan = grid_of(g.subgrid[1]);
        if (!yin || !yan || ensure_device(*yin) < 0 || ensure_device(*yan) < 0) return -1;
        float *ty = (float *)S.st_misc[2].get((size_t)n * sizeof(float));
        uint8_t *zb = (uint8_t *)S.st_scratch.get((size_t)n);
        if (!ty || !zb) return -1;
        ezk::split_u(d_y, yin->nj, ty, zb, n, S.stream);
        S.launches++;
        a.src = to_dev(*yin); a.src2 = to_dev(*yan);
        a.yy = 1;
        a.panel_stride = (long)yin->ni * yin->nj;
        a.y = ty; a.zb = zb;
    } else {
        if (ensure_device(g) < 0) return -1;
        a.src = to_dev(g);
    }
    S.last_path = 1;
    S.launches += ezk::interp(a, S.stream);
    return 0;
}

int wind_points_dev(Grid &g, float *d_u, float *d_v, const float *d_lat, const float *d_lon, const uint8_t *zb, int n, int mode) {
    WindArgs w;
    memset(&w, 0, sizeof(w));
    if (g.nsubgrids > 0) {
        w.src = to_dev(*grid_of(g.subgrid[0]));
        w.src2 = to_dev(*grid_of(g.subgrid[1]));
        w.yy = 1;
        w.zb = zb;
    } else {
        w.src = to_dev(g);
    }
    w.dst = w.src;   // mode 1 uses src only, mode 2 uses dst only
    w.lat = d_lat; w.lon = d_lon;
    w.u = d_u; w.v = d_v;
    w.stride = n; w.nfields = 1; w.npts = n; w.mode = mode;
    ezk::wind(w, S.stream);
    S.launches++;
    return 0;
}
}  // namespace

int32_t c_gdxyfll(int32_t gdid, float *x, float *y, float *lat, float *lon, int32_t n) {   // gdxyfll.c:190-243
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g || n < 0) return -1;
    if (n == 0) return 0;
    PtStage p;
    float *dla = p.in(lat, n), *dlo = p.in(lon, n), *dx = p.out(x, n), *dy = p.out(y, n);
    if (!p.ok || xyfll_dev(*g, dx, dy, dla, dlo, n) < 0) return -1;
    return p.finish();
}
int32_t gdxyfll_(int32_t *gdid, float *x, float *y, float *lat, float *lon, int32_t *n) { return c_gdxyfll(*gdid, x, y, lat, lon, *n); }

int32_t c_gdllfxy(int32_t gdid, float *lat, float *lon, float *x, float *y, int32_t n) {   // gdllfxy.c:30-92
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g || n < 0) return -1;
    if (g->grtyp == 'Y') {
        fprintf(stderr, "<gdllfxy>: This operation is not supported for 'Y' grids\n");
        return -1;
    }
    if (n == 0) return 0;
    Grid *a = g->nsubgrids > 0 ? grid_of(g->subgrid[0]) : g, *b = g->nsubgrids > 0 ? grid_of(g->subgrid[1]) : g;
    if (!a || !b || ensure_device(*a) < 0 || ensure_device(*b) < 0) return -1;
    PtStage p;
    float *dx = p.in(x, n), *dy = p.in(y, n), *dla = p.out(lat, n), *dlo = p.out(lon, n);
    if (!p.ok) return -1;
    ezk::xy2ll_points(to_dev(*a), to_dev(*b), g->nsubgrids > 0, dx, dy, dla, dlo, n, S.stream);
    S.launches++;
    return p.finish();
}
int32_t gdllfxy_(int32_t *gdid, float *lat, float *lon, float *x, float *y, int32_t *n) { return c_gdllfxy(*gdid, lat, lon, x, y, *n); }

int32_t c_gdxysval(int32_t gdin, float *zout, float *zin, float *x, float *y, int32_t n) {   // gdxysval.c:50-116
    Lock lk;
    Grid *g = grid_of(gdin);
    if (!g || n < 0) return -1;
    if (n == 0) return 0;
    PtStage p;
    float *dz = p.in(zin, (size_t)g->ni * g->nj), *dx = p.in(x, n), *dy = p.in(y, n), *dout = p.out(zout, n);
    if (!p.ok || xysint_dev(*g, dout, dz, dx, dy, n, 0, 1) < 0) return -1;
    return p.finish();
}
int32_t gdxysval_(int32_t *gdin, float *zout, float *zin, float *x, float *y, int32_t *n) { return c_gdxysval(*gdin, zout, zin, x, y, *n); }

int32_t c_gdllsval(int32_t gdid, float *zout, float *zin, float *lat, float *lon, int32_t n) {   // gdllsval.c:35-55
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g || n < 0) return -1;
    if (n == 0) return 0;
    PtStage p;
    float *dz = p.in(zin, (size_t)g->ni * g->nj), *dla = p.in(lat, n), *dlo = p.in(lon, n), *dout = p.out(zout, n);
    float *dx = p.tmp(n), *dy = p.tmp(n);
    if (!p.ok || xyfll_dev(*g, dx, dy, dla, dlo, n) < 0 || xysint_dev(*g, dout, dz, dx, dy, n, 0, 1) < 0) return -1;
    if (p.finish() < 0) return -1;
    return 0;
}
int32_t gdllsval_(int32_t *gdid, float *zout, float *zin, float *lat, float *lon, int32_t *n) { return c_gdllsval(*gdid, zout, zin, lat, lon, *n); }

// c_gdxyvval (gdxyvval.c:38-116): both components at caller-supplied x, y, as GRID components (no rotation)
int32_t c_gdxyvval(int32_t gdin, float *uuout, float *vvout, float *uuin, float *vvin, float *x, float *y, int32_t n) {
    Lock lk;
    Grid *g = grid_of(gdin);
    if (!g || n < 0) return -1;
    if (n == 0) return 0;
    const size_t nin = (size_t)g->ni * g->nj;
    PtStage p;
    float *du = p.in(uuin, nin), *dv = p.in(vvin, nin), *dx = p.in(x, n), *dy = p.in(y, n), *ou = p.out(uuout, n), *ov = p.out(vvout, n);
    if (!p.ok || xysint_dev(*g, ou, du, dx, dy, n, 1, +1) < 0 || xysint_dev(*g, ov, dv, dx, dy, n, 1, -1) < 0) return -1;
    return p.finish();
}
int32_t gdxyvval_(int32_t *gdin, float *uuout, float *vvout, float *uuin, float *vvin, float *x, float *y, int32_t *n) {
    return c_gdxyvval(*gdin, uuout, vvout, uuin, vvin, x, y, *n);
}

int32_t c_gdllvval(int32_t gdid, float *uuout, float *vvout, float *uuin, float *vvin, float *lat, float *lon, int32_t n) {   // gdllvval.c:37-58
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g || n < 0) return -1;
    if (g->nsubgrids > 0) {
        fprintf(stderr, "<gdllvval>: This operation is not supported for 'U' grids\n");
        return -1;
    }
    if (n == 0) return 0;
    const size_t nin = (size_t)g->ni * g->nj;
    PtStage p;
    float *du = p.in(uuin, nin), *dv = p.in(vvin, nin), *dla = p.in(lat, n), *dlo = p.in(lon, n), *ou = p.out(uuout, n), *ov = p.out(vvout, n);
    float *dx = p.tmp(n), *dy = p.tmp(n);
    if (!p.ok || xyfll_dev(*g, dx, dy, dla, dlo, n) < 0) return -1;
    if (xysint_dev(*g, ou, du, dx, dy, n, 1, +1) < 0 || xysint_dev(*g, ov, dv, dx, dy, n, 1, -1) < 0) return -1;
    if (p.finish() < 0) return -1;
    return 0;
}
int32_t gdllvval_(int32_t *gdid, float *uuout, float *vvout, float *uuin, float *vvin, float *lat, float *lon, int32_t *n) {
    return c_gdllvval(*gdid, uuout, vvout, uuin, vvin, lat, lon, *n);
}

// c_gdwdfuv (gdwdfuv.c:100-113) / c_gduvfwd (gduvfwd.c:112-125): wind conversions at caller-supplied lat/lon
int32_t c_gdwdfuv(int32_t gdid, float *spd_out, float *wd_out, float *uuin, float *vvin, float *latin, float *lonin, int32_t npts) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g || npts < 0) return -1;
    if (g->nsubgrids > 0) {
        fprintf(stderr, "<gdwdfuv>: This operation is not supported for 'U' grids\n");
        return -1;
    }
    if (npts == 0) return 0;
    PtStage p;
    float *dla = p.in(latin, npts), *dlo = p.in(lonin, npts), *du = p.in(uuin, npts), *dv = p.in(vvin, npts);
    float *os = p.out(spd_out, npts), *ow = p.out(wd_out, npts);
    if (!p.ok) return -1;
    if (os != du) CUDA_OK(cudaMemcpyAsync(os, du, (size_t)npts * sizeof(float), cudaMemcpyDeviceToDevice, S.stream));
    if (ow != dv) CUDA_OK(cudaMemcpyAsync(ow, dv, (size_t)npts * sizeof(float), cudaMemcpyDeviceToDevice, S.stream));
    if (wind_points_dev(*g, os, ow, dla, dlo, nullptr, npts, 1) < 0) return -1;
    return p.finish();
}
int32_t gdwdfuv_(int32_t *gdid, float *spd_out, float *wd_out, float *uuin, float *vvin, float *latin, float *lonin, int32_t *npts) {
    return c_gdwdfuv(*gdid, spd_out, wd_out, uuin, vvin, latin, lonin, *npts);
}
int32_t c_gduvfwd(int32_t gdid, float *uugdout, float *vvgdout, float *uullin, float *vvllin, float *latin, float *lonin, int32_t npts) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g || npts < 0) return -1;
    if (g->nsubgrids > 0) {
        fprintf(stderr, "<gduvfwd>: This operation is not supported for 'U' grids\n");
        return -1;
    }
    if (npts == 0) return 0;
    PtStage p;
    float *dla = p.in(latin, npts), *dlo = p.in(lonin, npts), *ds = p.in(uullin, npts), *dw = p.in(vvllin, npts);
    float *ou = p.out(uugdout, npts), *ov = p.out(vvgdout, npts);
    if (!p.ok) return -1;
    if (ou != ds) CUDA_OK(cudaMemcpyAsync(ou, ds, (size_t)npts * sizeof(float), cudaMemcpyDeviceToDevice, S.stream));
    if (ov != dw) CUDA_OK(cudaMemcpyAsync(ov, dw, (size_t)npts * sizeof(float), cudaMemcpyDeviceToDevice, S.stream));
    if (wind_points_dev(*g, ou, ov, dla, dlo, nullptr, npts, 2) < 0) return -1;
    return p.finish();
}
int32_t gduvfwd_(int32_t *gdid, float *uugdout, float *vvgdout, float *uullin, float *vvllin, float *latin, float *lonin, int32_t *npts) {
    return c_gduvfwd(*gdid, uugdout, vvgdout, uullin, vvllin, latin, lonin, *npts);
}

// c_gdxywdval (gdxywdval.c:38) / c_gdllwdval (gdllwdval.c:36-80): grid components at the points, then (speed, direction)
namespace {
int wdval_common(Grid &g, float *ou, float *ov, const float *du, const float *dv, const float *dx, const float *dy, const float *dla, const float *dlo,
                 int n) {
    if (xysint_dev(g, ou, du, dx, dy, n, 1, +1) < 0 || xysint_dev(g, ov, dv, dx, dy, n, 1, -1) < 0) return -1;
    const uint8_t *zb = nullptr;
    if (g.nsubgrids > 0) zb = (const uint8_t *)S.st_scratch.p;   // the panel bits xysint_dev just produced from y
    return wind_points_dev(g, ou, ov, dla, dlo, zb, n, 1);
}
}  // namespace
int32_t c_gdllwdval(int32_t gdid, float *uuout, float *vvout, float *uuin, float *vvin, float *lat, float *lon, int32_t n) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g || n < 0) return -1;
    if (n == 0) return 0;
    const size_t nin = (size_t)g->ni * g->nj;
    PtStage p;
    float *du = p.in(uuin, nin), *dv = p.in(vvin, nin), *dla = p.in(lat, n), *dlo = p.in(lon, n), *ou = p.out(uuout, n), *ov = p.out(vvout, n);
    float *dx = p.tmp(n), *dy = p.tmp(n);
    if (!p.ok || xyfll_dev(*g, dx, dy, dla, dlo, n) < 0 || wdval_common(*g, ou, ov, du, dv, dx, dy, dla, dlo, n) < 0) return -1;
    if (p.finish() < 0) return -1;
    return 0;
}
int32_t gdllwdval_(int32_t *gdid, float *uuout, float *vvout, float *uuin, float *vvin, float *lat, float *lon, int32_t *n) {
    return c_gdllwdval(*gdid, uuout, vvout, uuin, vvin, lat, lon, *n);
}
int32_t c_gdxywdval(int32_t gdin, float *uuout, float *vvout, float *uuin, float *vvin, float *x, float *y, int32_t n) {
    Lock lk;
    Grid *g = grid_of(gdin);
    if (!g || n < 0) return -1;
    if (g->grtyp == 'Y') return -1;
    if (n == 0) return 0;
    const size_t nin = (size_t)g->ni * g->nj;
    Grid *a = g->nsubgrids > 0 ? grid_of(g->subgrid[0]) : g, *b = g->nsubgrids > 0 ? grid_of(g->subgrid[1]) : g;
    if (!a || !b || ensure_device(*a) < 0 || ensure_device(*b) < 0) return -1;
    PtStage p;
    float *du = p.in(uuin, nin), *dv = p.in(vvin, nin), *dx = p.in(x, n), *dy = p.in(y, n), *ou = p.out(uuout, n), *ov = p.out(vvout, n);
    float *dla = p.tmp(n), *dlo = p.tmp(n);
    if (!p.ok) return -1;
    ezk::xy2ll_points(to_dev(*a), to_dev(*b), g->nsubgrids > 0, dx, dy, dla, dlo, n, S.stream);   // c_gdllfxy (gdxywdval.c)
    S.launches++;
    if (wdval_common(*g, ou, ov, du, dv, dx, dy, dla, dlo, n) < 0) return -1;
    if (p.finish() < 0) return -1;
    return 0;
}
int32_t gdxywdval_(int32_t *gdin, float *uuout, float *vvout, float *uuin, float *vvin, float *x, float *y, int32_t *n) {
    return c_gdxywdval(*gdin, uuout, vvout, uuin, vvin, x, y, *n);
}

int c_gdsetmask(int gdid, int *mask) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g) return -1;
    if (g->nsubgrids > 0) { fprintf(stderr, "<gdsetmask> This operation is not supported for 'U' grids.\n"); return -1; }
    g->mask.assign(mask, mask + (size_t)g->ni * g->nj);
    g->has_mask = true;
    return 0;
}
int gdsetmask_(int *gdid, int *mask) { return c_gdsetmask(*gdid, mask); }
int c_gdgetmask(int gdid, int *mask) {
    Lock lk;
    Grid *g = grid_of(gdid);
    if (!g) return -1;
    if (g->nsubgrids > 0) { fprintf(stderr, "<gdgetmask> This operation is not supported for 'U' grids.\n"); return -1; }
    if (!g->has_mask) return -1;
    memcpy(mask, g->mask.data(), g->mask.size() * sizeof(int));
    return 0;
}
int gdgetmask_(int *gdid, int *mask) { return c_gdgetmask(*gdid, mask); }

// ---- gridset selection ----
int32_t c_ezdefset(int32_t gdout, int32_t gdin) {   // ezdefset.c:38-170 ; the x/y arrays are built lazily by the first compute call
    Lock lk;
    if (gdout == -1 && gdin != -1) gdout = gdin;
    if (!grid_of(gdin) || !grid_of(gdout)) {
        fprintf(stderr, "<c_ezdefset> undefined grid id (gdout=%d gdin=%d)\n", gdout, gdin);
        return -1;
    }
    S.gdin = gdin;
    S.gdout = gdout;
    if (gdin != gdout) {
        auto &slot = grid_of(gdout)->ds().gsets[gdin];
        if (!slot) {
            slot = std::make_unique<Gset>();
            slot->gdin = gdin;
        }
    }
    return 1;
}
int32_t ezdefset_(int32_t *gdout, int32_t *gdin) { return c_ezdefset(*gdout, *gdin); }
int32_t c_ezgetgdin(void) { return S.gdin; }
int32_t c_ezgetgdout(void) {
    if (S.gdout == -1 && S.gdin != -1) S.gdout = S.gdin;
    return S.gdout;
}

// ---- hot path ----
int32_t c_ezsint(float *zout, float *zin) { return ezsint_n(zout, zin, 1); }
int32_t ezsint_(float *zout, float *zin) { return c_ezsint(zout, zin); }
int32_t c_ezsint_batch(float *zout, float *zin, int32_t nfields) { return ezsint_n(zout, zin, nfields); }
int32_t c_ezuvint(float *uuout, float *vvout, float *uuin, float *vvin) { return ezuvint_n(uuout, vvout, uuin, vvin, 1, 0); }
int32_t ezuvint_(float *uuout, float *vvout, float *uuin, float *vvin) { return c_ezuvint(uuout, vvout, uuin, vvin); }
int32_t c_ezuvint_batch(float *uuout, float *vvout, float *uuin, float *vvin, int32_t nfields) {
    return ezuvint_n(uuout, vvout, uuin, vvin, nfields, 0);
}
int32_t c_ezwdint(float *spdout, float *wdout, float *uuin, float *vvin) { return ezuvint_n(spdout, wdout, uuin, vvin, 1, 1); }
int32_t ezwdint_(float *spdout, float *wdout, float *uuin, float *vvin) { return c_ezwdint(spdout, wdout, uuin, vvin); }

int c_ezsint_mask(int *mask_out, int *mask_in) {   // ez_mask.c:184-223
    Lock lk;
    if (!session_ok("ezsint_mask")) return -1;
    Grid &in = *grid_of(S.gdin), &out = *grid_of(S.gdout);
    if (in.nsubgrids > 0 || out.nsubgrids > 0) {
        fprintf(stderr, "<ezsint_mask> This operation is not supported for 'U' grids.\n");
        return -1;
    }
    if (in.id == out.id || in.grtyp == 'Y') return -1;
    Gset *gs = ensure_gset(out, in);
    if (!gs) return -1;
    size_t nin = (size_t)in.ni * in.nj, nout = (size_t)out.ni * out.nj;
    bool din = is_device_ptr(mask_in), dout = is_device_ptr(mask_out);
    int *d_in = din ? mask_in : (int *)S.st_in[0].get(nin * sizeof(int));
    int *d_out = dout ? mask_out : (int *)S.st_out[0].get(nout * sizeof(int));
    if (!d_in || !d_out) return -1;
    if (!din) CUDA_OK(cudaMemcpyAsync(d_in, mask_in, nin * sizeof(int), cudaMemcpyHostToDevice, S.stream));
    ezk::mask_transfer(gs->d_x, gs->d_y, d_in, d_out, out.ni, out.nj, in.ni, in.nj, S.opt.cld_interp_alg == EZ_LINEAR, S.stream);
    S.launches++;
    if (!dout) {
        CUDA_OK(cudaMemcpyAsync(mask_out, d_out, nout * sizeof(int), cudaMemcpyDeviceToHost, S.stream));
        CUDA_OK(cudaStreamSynchronize(S.stream));
    }
    return 0;
}
int ezsint_mask_(int *mask_out, int *mask_in) { return c_ezsint_mask(mask_out, mask_in); }

// lorenzo_mask_fill(fld, mask, ni, nj, 2) on a caller array (host or device pointers)
static int mask_fill_any(float *fld, int *mask, size_t n) {
    Lock lk;
    unsigned int *scr = (unsigned int *)S.st_scratch.get(2 * sizeof(unsigned int));
    if (!scr) return -1;
    const bool df = is_device_ptr(fld), dm = is_device_ptr(mask);
    float *d_f = df ? fld : (float *)S.st_out[0].get(n * sizeof(float));
    int *d_m = dm ? mask : (int *)S.st_in[0].get(n * sizeof(int));
    if (!d_f || !d_m) return -1;
    if (!df) CUDA_OK(cudaMemcpyAsync(d_f, fld, n * sizeof(float), cudaMemcpyHostToDevice, S.stream));
    if (!dm) CUDA_OK(cudaMemcpyAsync(d_m, mask, n * sizeof(int), cudaMemcpyHostToDevice, S.stream));
    ezk::mask_fill_min(d_f, d_m, (long)n, scr, S.stream);
    S.launches += 3;
    CUDA_OK(cudaGetLastError());
    if (!df) {
        CUDA_OK(cudaMemcpyAsync(fld, d_f, n * sizeof(float), cudaMemcpyDeviceToHost, S.stream));
        CUDA_OK(cudaStreamSynchronize(S.stream));
    }
    return 0;
}
static int mdm_grids(const char *who, size_t *nout) {
    Lock lk;
    if (!session_ok(who)) return -1;
    Grid &in = *grid_of(S.gdin), &out = *grid_of(S.gdout);
    if (in.nsubgrids > 0 || out.nsubgrids > 0) {
        fprintf(stderr, "<%s> This operation is not supported for 'U' grids.\n", who);
        return -1;
    }
    *nout = (size_t)out.ni * out.nj;
    return 0;
}
// ez_mask.c:127-152: c_ezsint, the mask transfer, then the masked-out target points take the field's minimum
int32_t c_ezsint_mdm(float *zout, int *mask_out, float *zin, int *mask_in) {
    size_t nout = 0;
    if (mdm_grids("ezsint_mdm", &nout) < 0) return -1;
    if (ezsint_n(zout, zin, 1) < 0 || c_ezsint_mask(mask_out, mask_in) < 0) return -1;
    return mask_fill_any(zout, mask_out, nout) < 0 ? -1 : 0;
}
// ez_mask.c:155-181
int32_t c_ezuvint_mdm(float *uuout, float *vvout, int *mask_out, float *uuin, float *vvin, int *mask_in) {
    size_t nout = 0;
    if (mdm_grids("ezuvint_mdm", &nout) < 0) return -1;
    if (c_ezsint_mask(mask_out, mask_in) < 0 || ezuvint_n(uuout, vvout, uuin, vvin, 1, 0) < 0) return -1;
    if (mask_fill_any(uuout, mask_out, nout) < 0 || mask_fill_any(vvout, mask_out, nout) < 0) return -1;
    return 0;
}
int32_t ezsint_mdm_(float *zout, int *mask_out, float *zin, int *mask_in) { return c_ezsint_mdm(zout, mask_out, zin, mask_in); }
int32_t ezuvint_mdm_(float *uuout, float *vvout, int *mask_out, float *uuin, float *vvin, int *mask_in) {
    return c_ezuvint_mdm(uuout, vvout, mask_out, uuin, vvin, mask_in);
}
// ez_mask.c:114-124: announced as not implemented by the reference itself
int32_t c_ezsint_m(float *zout, float *zin) { (void)zout; (void)zin; fprintf(stderr, "<ezsint_m> This operation is currently not implemented.\n"); return 0; }
int32_t c_ezuvint_m(float *uuout, float *vvout, float *uuin, float *vvin) {
    (void)uuout; (void)vvout; (void)uuin; (void)vvin;
    fprintf(stderr, "<ezuvint_m> This operation is currently not implemented.\n");
    return 0;
}

int32_t c_ezb200_get_gridset(int32_t gdout, int32_t gdin, float *x, float *y, uint8_t *zonebits, int32_t *counts) {
    Lock lk;
    Grid *o = grid_of(gdout), *i = grid_of(gdin);
    if (!o || !i || o->nsubgrids > 0) return -1;
    Gset *gs = ensure_gset(*o, *i);
    if (!gs) return -1;
    size_t n = (size_t)o->ni * o->nj;
    if (x) CUDA_OK(cudaMemcpyAsync(x, gs->d_x, n * sizeof(float), cudaMemcpyDefault, S.stream));
    if (y) CUDA_OK(cudaMemcpyAsync(y, gs->d_y, n * sizeof(float), cudaMemcpyDefault, S.stream));
    if (zonebits) CUDA_OK(cudaMemcpyAsync(zonebits, gs->d_zb, n, cudaMemcpyDefault, S.stream));
    CUDA_OK(cudaStreamSynchronize(S.stream));
    if (counts) for (int k = 0; k < 6; k++) counts[k] = gs->counts[k];
    return 0;
}

// ---- options (ezsetopt.c:59-215, ezsetval.c:86-148, ezgetopt.c:53-234, ezgetval.c:55-120) ----
int32_t c_ezsetopt(char *option, char *value) {
    char o[32], v[32];
    lower(o, option);
    lower(v, value);
    if (!strcmp(o, "correction_polaire")) strcpy(o, "polar_correction");
    if (!strcmp(o, "degre_interp")) strcpy(o, "interp_degree");
    if (!strcmp(o, "degre_extrap")) strcpy(o, "extrap_degree");
    if (!strcmp(o, "use_1sousgrille")) strcpy(o, "use_1subgrid");
    if (!strcmp(v, "oui")) strcpy(v, "yes");
    if (!strcmp(v, "ouiouioui")) strcpy(v, "yesyesyes");
    if (!strcmp(v, "non")) strcpy(v, "no");
    if (!strcmp(v, "voisin")) strcpy(v, "nearest");
    if (!strcmp(v, "lineair") || !strcmp(v, "lineaire")) strcpy(v, "linear");
    if (!strcmp(v, "cubique")) strcpy(v, "cubic");
    if (!strcmp(v, "neutre")) strcpy(v, "neutral");
    if (!strcmp(v, "valeur")) strcpy(v, "value");
    int option_ok = 0, value_ok = 0;
    Options &g = S.opt;
    if (!strcmp(o, "use_1subgrid")) {
        option_ok = value_ok = 1;
        if (!strcmp(v, "yes")) g.use_1subgrid = 1;
        else if (!strcmp(v, "no")) g.use_1subgrid = 0;
        else value_ok = 0;
    }
    if (!strcmp(o, "verbose")) {
        option_ok = value_ok = 1;
        if (!strcmp(v, "yes")) g.verbose = 1;
        else if (!strcmp(v, "yesyesyes")) g.verbose = 2;
        else if (!strcmp(v, "no")) g.verbose = 0;
        else value_ok = 0;
    }
    if (!strcmp(o, "polar_correction")) {
        option_ok = value_ok = 1;
        if (!strcmp(v, "yes")) g.polar_correction = 1;
        else if (!strcmp(v, "no")) g.polar_correction = 0;
        else value_ok = 0;
    }
    if (!strcmp(o, "interp_degree")) {
        option_ok = value_ok = 1;
        if (!strcmp(v, "nearest")) g.degre_interp = 0;
        else if (!strcmp(v, "linear")) g.degre_interp = 1;
        else if (!strcmp(v, "cubic")) g.degre_interp = 3;
        else if (!strcmp(v, "average")) g.degre_interp = 4;
        else if (!strcmp(v, "sph_average")) g.degre_interp = 5;
        else value_ok = 0;
    }
    if (!strcmp(o, "extrap_degree")) {
        option_ok = value_ok = 1;
        if (!strcmp(v, "neutral")) g.degre_extrap = g.degre_interp;
        else if (!strcmp(v, "nearest")) g.degre_extrap = 0;
        else if (!strcmp(v, "linear")) g.degre_extrap = 1;
        else if (!strcmp(v, "cubic")) g.degre_extrap = 3;
        else if (!strcmp(v, "maximum")) g.degre_extrap = EZ_MAXIMUM;
        else if (!strcmp(v, "minimum")) g.degre_extrap = EZ_MINIMUM;
        else if (!strcmp(v, "value")) g.degre_extrap = EZ_VALUE;
        else if (!strcmp(v, "abort")) g.degre_extrap = EZ_ABORT;
        else value_ok = 0;
    }
    if (!strcmp(o, "cloud_interp_alg")) {
        if (!strcmp(v, "linear")) { value_ok = option_ok = 1; g.cld_interp_alg = EZ_LINEAR; }
        if (!strcmp(v, "distance")) { value_ok = option_ok = 1; g.cld_interp_alg = 4; }
    }
    if (value_ok + option_ok != 2) {
        if (!option_ok) fprintf(stderr, "ezsetopt : option not recognized : %s\n", option);
        if (!value_ok) fprintf(stderr, "ezsetopt : value not recognized : %s\n", value);
        return -1;
    }
    return 0;
}
int32_t ezsetopt_(char *option, char *value, F2Cl lo, F2Cl lv) {
    char o[32], v[32];
    ftn_copy(o, option, lo);
    ftn_copy(v, value, lv);
    return c_ezsetopt(o, v);
}
int32_t c_ezsetval(char *option, float fvalue) {
    char o[32];
    lower(o, option);
    if (!strcmp(o, "extrap_value")) S.opt.valeur_extrap = fvalue;
    if (!strcmp(o, "missing_gridpt_distance")) S.opt.msg_gridpt_dist = fvalue;
    if (!strcmp(o, "missing_distance_threshold")) S.opt.msg_dist_thresh = fvalue;
    return 0;
}
int32_t ezsetval_(char *option, float *fvalue, F2Cl l) {
    char o[32];
    ftn_copy(o, option, l);
    return c_ezsetval(o, *fvalue);
}
int32_t c_ezsetival(char *option, int32_t ivalue) {
    char o[32];
    lower(o, option);
    if (!strcmp(o, "weight_number")) S.opt.wgt_num = ivalue;
    if (!strcmp(o, "missing_points_tolerance")) S.opt.msg_pt_tol = ivalue;
    if (!strcmp(o, "subgridid")) S.opt.valeur_1subgrid = ivalue;
    return 0;
}
int32_t ezsetival_(char *option, int32_t *ivalue, F2Cl l) {
    char o[32];
    ftn_copy(o, option, l);
    return c_ezsetival(o, *ivalue);
}
int32_t c_ezgetopt(char *option, char *value) {
    char o[32];
    lower(o, option);
    const Options &g = S.opt;
    auto deg = [](int d, bool fr, bool ex) -> const char * {
        switch (d) {
            case 0: return fr ? "voisin" : "nearest";
            case 1: return fr ? "lineaire" : "linear";
            case 3: return fr ? "cubique" : "cubic";
            default: break;
        }
        if (ex) {
            if (d == EZ_MAXIMUM) return "maximum";
            if (d == EZ_MINIMUM) return "minimum";
            if (d == EZ_VALUE) return fr ? "valeur" : "value";
            if (d == EZ_ABORT) return "abort";
        }
        return fr ? "erreur" : "error";
    };
    if (!strcmp(o, "verbose")) strcpy(value, g.verbose == 1 ? "yes" : (g.verbose == 2 ? "yesyesyes" : "no"));
    if (!strcmp(o, "use_1subgrid")) strcpy(value, g.use_1subgrid == 1 ? "yes" : "no");
    if (!strcmp(o, "polar_correction")) strcpy(value, g.polar_correction == 1 ? "yes" : "no");
    if (!strcmp(o, "interp_degree")) strcpy(value, deg(g.degre_interp, false, false));
    if (!strcmp(o, "degre_interp")) strcpy(value, deg(g.degre_interp, true, false));
    if (!strcmp(o, "extrap_degree")) strcpy(value, deg(g.degre_extrap, false, true));
    if (!strcmp(o, "degre_extrap")) strcpy(value, deg(g.degre_extrap, true, true));
    if (!strcmp(o, "cloud_interp_alg")) strcpy(value, g.cld_interp_alg == EZ_LINEAR ? "linear" : "distance");
    return 0;
}
int32_t ezgetopt_(char *option, char *value, F2Cl lo, F2Cl lv) {
    char o[32], v[32];
    v[0] = 0;
    ftn_copy(o, option, lo);
    int32_t rc = c_ezgetopt(o, v);
    int len = (int)strlen(v);
    if (len > lv) len = lv;
    memcpy(value, v, len);
    if (len < lv) value[len] = 0;
    return rc;
}
int32_t c_ezgetval(char *option, float *value) {
    char o[32];
    lower(o, option);
    if (!strcmp(o, "extrap_value")) *value = S.opt.valeur_extrap;
    if (!strcmp(o, "missing_distance_threshold")) *value = S.opt.msg_dist_thresh;
    if (!strcmp(o, "weight_number")) *value = (float)S.opt.wgt_num;
    if (!strcmp(o, "missing_points_tolerance")) *value = (float)S.opt.msg_pt_tol;
    return 0;
}
int32_t ezgetval_(char *option, float *value, F2Cl l) {
    char o[32];
    ftn_copy(o, option, l);
    return c_ezgetval(o, value);
}
int32_t c_ezgetival(char *option, int32_t *value) {
    char o[32];
    lower(o, option);
    if (!strcmp(o, "subgridid")) *value = S.opt.valeur_1subgrid;
    if (!strcmp(o, "weight_number")) *value = S.opt.wgt_num;
    if (!strcmp(o, "missing_points_tolerance")) *value = S.opt.msg_pt_tol;
    return 0;
}
int32_t ezgetival_(char *option, int32_t *value, F2Cl l) {
    char o[32];
    ftn_copy(o, option, l);
    return c_ezgetival(o, value);
}

}  // extern "C"
