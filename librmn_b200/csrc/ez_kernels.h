// librmn_b200/csrc/ez_kernels.h -- host-side launchers of the CUDA kernels (internal interface between the
// host logic in ez_host.cu and the kernel translation units).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ez_types.h"

// Everything a gather-interpolation launch needs, by value.
struct InterpArgs {
    GridDev src;            // source grid (Yin panel when the source is a 'U' supergrid)
    GridDev src2;           // Yang panel (only read when yy != 0)
    int yy;                 // source is Yin-Yang: ZB_YANG in zb picks src2 and adds panel_stride to the field base
    long panel_stride;      // ni*nj of one panel
    const float *x, *y;     // device, npts each (fractional 1-based source coordinates)
    const uint8_t *zb;      // device zone/panel bits, or nullptr
    int use_zones;          // apply zone handling (polar_correction on, zones built)
    const float *zin;       // device source fields
    float *zout;            // device target fields
    long in_stride, out_stride;   // elements between consecutive fields
    int nfields;
    int npts;
    InterpOpts o;
    const FieldScalars *fs; // device per-field scalars (pole means, fudge) or nullptr
    const float *polerow;   // vectors: synthetic pole rows of this component, [nfields][north, south][ni], or nullptr (K18)
    // tiled kernel (ez_interp_tiled.cuh): target grid shape, per-tile descriptors and thread masks
    int out_ni, out_nj;
    const TileDesc *tiles;
    const uint32_t *tile_flags;   // 8 words per tile: bit t = thread t takes the tile path
    int tile_cap;                 // floats reserved per staged source level (>= the largest tile box)
    // row-separable kernel (ez_interp_rowsep.cuh): y of each target row of each tile (32 words per tile), width of the
    // widest row-separable box, how many tiles take each of the two tile kernels
    const unsigned int *tile_rowy;
    int tile_vw;                  // widest row-separable box (<= 32: pitch-33 kernel, else pitch-65 kernel)
    int tile_cap_rs;              // floats per staged source level in the row-separable kernel (pitch x tallest box)
    int ntiles_rowsep, ntiles_window;
    const int *win_tiles;         // compacted list of the tiles the (persistent) window kernel owns, ntiles_window entries
    int rowsep_enable;            // classification only: allow tiles to be marked row-separable
    // output-row band (multi-GPU split of one field): only target rows [band_j0, band_j1) are produced; zout then
    // points band_j0 rows BEFORE the caller's band buffer, so that full-grid indexing lands in it.  0, 0 = the whole target.
    int band_j0, band_j1;
    // plan cache of the tile plan (per target point: stencil origin, fp32 x / y weights, s = sum wx sum wy - 1), or nullptr:
    // the tile kernels then load 40 B per point instead of redoing the fp64 plan arithmetic in every launch
    const int *pc_idx;            // (r0 << 16) | c0
    const float4 *pc_wx, *pc_wy;
    const float *pc_sm1;
    // per-point kernel restricted to a compacted list of target points (nullptr = all npts points)
    const int *list;
    int nlist;
    int list_lch;           // levels per thread when the kernel runs on a list (set by the launcher; 0 = the default chunk)
    // tiled path with a constant extrapolation value: the points outside the source (and in no other zone), taken off `list`
    const int *clist;
    int nclist;
    // direct kernel on a scattered ('Y' cloud) target: the order in which the points are walked (sorted by source cell), or nullptr
    const int *order;
    // direct kernel, c_ezuvint of one pair: cached grid-to-grid wind rotation (ez_wind.cu), applied to fields (0, 1) before the store
    const float4 *wrot;
};

struct WindArgs {
    GridDev src;            // grid whose frame (u,v) are currently expressed in
    GridDev src2;           // Yang panel
    int yy;
    GridDev dst;            // grid whose frame they must be expressed in
    const float *lat, *lon; // device lat/lon of the target points
    const uint8_t *zb;      // panel bits (yy) or nullptr
    float *u, *v;           // device, in place
    long stride;
    int nfields;
    int npts;
    int mode;               // 0: uv(src frame) -> uv(dst frame) ; 1: -> (speed,direction) only ; 2: (speed,direction) -> uv(dst)
};

// K18 synthetic pole rows for vectors (ez_calcnpolarwind.c / ez_calcspolarwind.c): prow_u, prow_v are [nfields][2][ni]
struct PolarWindArgs {
    GridDev src;
    const float *u, *v;     // device source components, nfields fields `stride` apart
    long stride;
    int nfields;
    float dgrw_n, dgrw_s;   // dgrw of the auxiliary 'N' / 'S' grids after the cxgaig -> cigaxg round trip
    float *prow_u, *prow_v;
};

namespace ezk {
void calclatlon(const GridDev &g, float *lat, float *lon, int npts, cudaStream_t s);
void calcxy(const GridDev &src, int njs, const float *lat, float *lon, int writeback_lon, float *x, float *y, uint8_t *zb,
            int zones_enabled, int extrap, float ypolenord, float ypolesud, unsigned int *counts, int npts, cudaStream_t s);
void yy_calcxy(const GridDev &maskgrid, const GridDev &yin, const GridDev &yan, int use_mask, const float *lat, const float *lon,
               float *x, float *y, uint8_t *zb, unsigned int *counts, int npts, cudaStream_t s);
void ll2xy_points(const GridDev &g, int njs, const float *lat, const float *lon, float *x, float *y, int npts, cudaStream_t s);
// scattered-point API: inverse transform, 'U' supergrid selection, 'U' panel split
void xy2ll_points(const GridDev &g, const GridDev &g2, int is_u, const float *x, const float *y, float *lat, float *lon, int npts, cudaStream_t s);
void ll2xy_u(const GridDev &yin, const GridDev &yan, const int myin[4], const int myan[4], const float *lat, const float *lon, float *x, float *y,
             int npts, cudaStream_t s);
void split_u(const float *y, int njyin, float *ty, uint8_t *zb, int npts, cudaStream_t s);

// K11 hemispheric mirror expansion / y inversion of a source field into its logical layout z(ni, j1:j2) (ez_xpnsrcgd, permut)
void expand_field(const GridDev &g, const float *zin, float *zexp, long in_stride, long exp_stride, int nfields, int sym_sign, cudaStream_t s);
// K10 pre-pass: per-field pole-row means (sequential fp32 sums, bit-faithful to ez_calcpoleval) and min/max.
void field_scalars(const InterpArgs &a, int need_poles, int need_minmax, FieldScalars *fs, unsigned int *scratch, cudaStream_t s);
// K0-K9: the gather-interpolation kernel.  Returns the number of kernel launches issued.
// optional per-kernel timing of the tiled path: events recorded on `s` around [row-separable | window | list] launches
struct InterpTiming {
    cudaEvent_t ev[4];   // before row-separable, before window, before list, after list
};
// optional fork/join resources: the kernels of one tiled launch touch disjoint target points, so they may run concurrently
// (the smaller kernels fill the SMs the largest one leaves idle in its last wave)
struct InterpFork {
    cudaStream_t aux[2];
    cudaEvent_t ev_fork, ev_join[2];
};
int interp(const InterpArgs &a, cudaStream_t s, const InterpTiming *timing = nullptr, const InterpFork *fork = nullptr);
// calls of a few fields: one thread per target point, zone points on `zlist` (nzlist entries) through the per-point kernel
bool direct_eligible(const InterpArgs &a);
// pre: the per-field scalars pre-pass of this launch (field_scalars), issued by interp_direct so that it can run beside the kernel
struct DirectPre {
    bool run;
    bool has_dehors;   // the gridset has points outside the source (only then does the direct kernel itself read the scalars)
    int need_poles, need_minmax;
    FieldScalars *fs;
    unsigned int *scratch;
};
// timing (c_ezb200_set_timing): events around [scalars pre-pass | direct kernel | zone-list kernel], all in line on s
int interp_direct(const InterpArgs &a, const int *zlist, int nzlist, cudaStream_t s, const InterpFork *fork = nullptr, const DirectPre *pre = nullptr,
                  const InterpTiming *timing = nullptr);
void zone_list(const uint8_t *zb, int npts, int skip_dehors_only, int *list, unsigned int *cursor, cudaStream_t s, const int *order = nullptr);
// interp_degree = average / sph_average (ez_avg.inc, ez_avg_sph.inc): every target point, then ez_corrval on the zone list.
// row_lat: latitudes of the target points (first column read), sph_average only
int interp_avg(const InterpArgs &a, const float *row_lat, int sph, const int *zlist, int nzlist, cudaStream_t s);
// scattered targets: permutation of the target points by source cell, for gather locality in the per-point kernel
bool sort_points_by_cell(const float *x, const float *y, const uint8_t *zb, int ni, int nj, int npts, int *perm, cudaStream_t s);
// tiled path: is (degree, arith, source) eligible ; classification passes ; tile geometry
bool tiled_eligible(const InterpArgs &a);
void tile_dims(int out_ni, int out_nj, int *ntx, int *nty);
void tile_classify(const InterpArgs &a, TileDesc *tiles, uint32_t *flags, unsigned int *rowy, unsigned int *counters, cudaStream_t s);
void tile_list(const InterpArgs &a, const uint32_t *flags, int *list, unsigned int *cursor, int mode, cudaStream_t s);
void plan_cache(const InterpArgs &a, int *idx, float4 *wx, float4 *wy, float *sm1, cudaStream_t s, int mark_split = 0);
// 'Y' point clouds as a SOURCE (ez_cloud.cu): k nearest source points + inverse-square-distance weights per target point
// ([nw][ndst] plane-major, 1-based indices, 0 = empty slot), and their application to nfields fields
bool cloud_build(const float *src_lat, const float *src_lon, int nsrc, const float *dst_lat, const float *dst_lon, int ndst, int nw,
                 int *idx, float *wts, int *mask, cudaStream_t s);
int cloud_apply(const int *idx, const float *wts, const int *mask, int nw, const float *zin, float *zout, long in_stride, long out_stride, int nfields,
                int nsrc, int ndst, int ni_dst, int use_value, float value, unsigned int *scratch, cudaStream_t s);
// K19 integer mask transfer
void mask_transfer(const float *x, const float *y, const int *mask_in, int *mask_out, int ni_out, int nj_out, int ni_in, int nj_in,
                   int linear, cudaStream_t s);
// lorenzo_mask_fill (methode 2): fld = minval(fld) wherever mask == 0 ; scratch: 2 words
void mask_fill_min(float *fld, const int *mask, long n, unsigned int *scratch, cudaStream_t s);
// K17 wind rotation
void wind(const WindArgs &a, cudaStream_t s);
// cached per-point rotation of the (u,v) -> (u,v) conversion: build once per gridset, apply per call
bool wind_matrix(const WindArgs &a, float4 *m, cudaStream_t s);
void wind_apply(const float4 *m, float *u, float *v, long stride, int nfields, int npts, cudaStream_t s);
void wind_apply_list(const float4 *m, float *u, float *v, long stride, int nfields, const int *list, int nlist, int n0, int n1, cudaStream_t s);
// K18 synthetic pole rows for vectors (ez_calcnpolarwind.c / ez_calcspolarwind.c)
void polar_wind_rows(const PolarWindArgs &a, cudaStream_t s);
// B-grid pole-row averaging, ez_corrbgd.inc:29-52
void corrbgd(float *zout, long stride, int nfields, int ni, int nj, int hem, cudaStream_t s);
}  // namespace ezk
