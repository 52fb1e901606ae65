// librmn_b200/csrc/ez_cloud.cu -- 'Y' point clouds as the SOURCE of an interpolation.
//
// Reference: ez_calcxy.c:99-128 -> ez_calcxy_y.inc:1-94 builds, for every target point, the `weight_number` (default 16)
// nearest source points -- nearest in plain (lon, lat) degrees, within a +/-10 degree window, longitudes wrapped -- and their
// normalised inverse-square-distance weights; gdinterp.c:207-235 -> ez_applywgts.inc:19-48 applies them.  The search is a
// brute-force double loop, O(N_out N_src): 10^12 distance evaluations for a million points on either side.
//
// Here the source points are binned once into lon/lat cells (sorted cell keys + binary search, longitude periodic) and every
// target point walks the rings of cells around its own until the k-th best distance is closer than anything a further
// ring could hold: the same k points -- ties between equal distances go to the lower source index, as the reference's scan
// order does -- at O(k) cells per target.  The weights use the reference's fp32 operation sequence; the order in which the k
// slots are filled differs from the reference's replacement history, so the weighted sum is the same set of terms added
// in another order (fp32 rounding of the sum only).
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "ez_kernels.h"

namespace {

constexpr int TPB = 256;
constexpr int MAXW = 32;   // upper bound on weight_number (the reference's own scratch table is 16 deep, ez_calcxy_y.inc:11)

__device__ __forceinline__ unsigned int f2ord(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

struct CellGrid {
    double c;        // cell size in degrees, 360 / ncx exactly
    int ncx, ncy;
    __device__ __forceinline__ int cx(float lon) const {
        int k = (int)floor((double)lon / c);
        k %= ncx;
        return k < 0 ? k + ncx : k;
    }
    __device__ __forceinline__ int cy(float lat) const {
        int k = (int)floor(((double)lat + 90.0) / c);
        return k < 0 ? 0 : (k >= ncy ? ncy - 1 : k);
    }
};

__global__ void k_cloud_minmax_init(unsigned int *s, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) s[k] = (k & 1) ? 0u : 0xffffffffu;   // even = min, odd = max
}
// min / max of `nseries` arrays of n floats `stride` apart: s[2 q] = min, s[2 q + 1] = max (ordered-uint encoding)
__global__ void __launch_bounds__(TPB) k_cloud_minmax(const float *__restrict__ v, long stride, int n, unsigned int *__restrict__ s) {
    const int q = blockIdx.y;
    const float *p = v + (long)q * stride;
    float mn = INFINITY, mx = -INFINITY;
    for (long k = (long)blockIdx.x * TPB + threadIdx.x; k < n; k += (long)gridDim.x * TPB) {
        const float x = __ldg(p + k);
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(s + 2 * q, f2ord(mn));
        atomicMax(s + 2 * q + 1, f2ord(mx));
    }
}
__global__ void k_cloud_decode(const unsigned int *s, float *out, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = ord2f(s[k]);
}

// occupancy of a coarse 1 x 1 degree raster (360 x 180 bytes): lets the ring walk skip empty regions without a binary search
constexpr int OCC_NX = 360, OCC_NY = 180;
__device__ __forceinline__ int occ_x(double lon) { int k = (int)floor(lon); k %= OCC_NX; return k < 0 ? k + OCC_NX : k; }
__device__ __forceinline__ int occ_y(double lat) { int k = (int)floor(lat + 90.0); return k < 0 ? 0 : (k >= OCC_NY ? OCC_NY - 1 : k); }

__global__ void __launch_bounds__(TPB) k_cloud_keys(const float *__restrict__ lat, const float *__restrict__ lon, int n, CellGrid g,
                                                   unsigned int *__restrict__ keys, int *__restrict__ vals, unsigned char *__restrict__ occ) {
    const int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= n) return;
    keys[k] = (unsigned int)g.cy(lat[k]) * (unsigned int)g.ncx + (unsigned int)g.cx(lon[k]);
    vals[k] = k;
    // mark through the fine cell's own extent, so that the test in the ring walk (which sees cells, not points) is conservative
    const double x0 = g.cx(lon[k]) * g.c, y0 = g.cy(lat[k]) * g.c - 90.0;
    for (int yy = occ_y(y0); yy <= occ_y(y0 + g.c); yy++)
        for (double xx = x0; ; xx += 1.0) {
            occ[yy * OCC_NX + occ_x(xx)] = 1;
            if (xx >= x0 + g.c) break;
        }
}

// first position of `key` in the sorted key array, or -1
__device__ __forceinline__ int cell_begin(const unsigned int *__restrict__ keys, int n, unsigned int key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n && __ldg(keys + lo) == key) ? lo : -1;
}

// ez_calcxy_y.inc:30-91 for one target point per thread
__global__ void __launch_bounds__(128) k_cloud_knn(const float *__restrict__ slat, const float *__restrict__ slon, int nsrc,
                                                   const unsigned int *__restrict__ keys, const int *__restrict__ order, CellGrid g,
                                                   const float *__restrict__ dlat, const float *__restrict__ dlon, int ndst, int nw,
                                                   const float *__restrict__ bbox, const unsigned char *__restrict__ occ, int *__restrict__ idx,
                                                   float *__restrict__ wts, int *__restrict__ mask) {
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= ndst) return;
    const float lo = dlon[p], la = dlat[p];
    for (int n = 0; n < nw; n++) { idx[(long)n * ndst + p] = 0; wts[(long)n * ndst + p] = 1.0e30f; }
    // masque: inside the bounding box of the source coordinates (:21-29)
    if (lo < bbox[0] || lo > bbox[1] || la < bbox[2] || la > bbox[3]) { mask[p] = 0; return; }
    float bd[MAXW];
    int bi[MAXW];
    int cnt = 0;
    bool found = false;
    const int cx0 = g.cx(lo), cy0 = g.cy(la);
    const int rmax = (int)ceil(10.0 / g.c) + 1;
    for (int r = 0; r <= rmax; r++) {
        for (int dy = -r; dy <= r; dy++) {
            const int cy = cy0 + dy;
            if (cy < 0 || cy >= g.ncy) continue;
            const bool edge_row = (dy == -r || dy == r);
            for (int dx = -r; dx <= r; dx += (edge_row ? 1 : (r > 0 ? 2 * r : 1))) {
                int cx = (cx0 + dx) % g.ncx;
                if (cx < 0) cx += g.ncx;
                if (g.c <= 1.0 && !occ[occ_y((cy + 0.5) * g.c - 90.0) * OCC_NX + occ_x((cx + 0.5) * g.c)]) continue;   // nothing anywhere near this cell
                const unsigned int key = (unsigned int)cy * (unsigned int)g.ncx + (unsigned int)cx;
                int q = cell_begin(keys, nsrc, key);
                if (q < 0) continue;
                for (; q < nsrc && __ldg(keys + q) == key; q++) {
                    const int s = __ldg(order + q);
                    float rx = lo - __ldg(slon + s);
                    const float ry = la - __ldg(slat + s);
                    if (fabsf(rx) > 180.0f) rx = fabsf(fabsf(rx) - 360.0f);
                    if (fabsf(ry) < 10.0f && fabsf(rx) < 10.0f) {
                        const float dist = rx * rx + ry * ry;
                        found = true;
                        // keep the nw smallest by (distance, source index): the reference's scan keeps the earlier index on ties
                        if (cnt < nw || dist < bd[cnt - 1] || (dist == bd[cnt - 1] && s < bi[cnt - 1])) {
                            int k = (cnt < nw) ? cnt : nw - 1;
                            while (k > 0 && (bd[k - 1] > dist || (bd[k - 1] == dist && bi[k - 1] > s))) {
                                bd[k] = bd[k - 1];
                                bi[k] = bi[k - 1];
                                k--;
                            }
                            bd[k] = dist;
                            bi[k] = s;
                            if (cnt < nw) cnt++;
                        }
                    }
                }
            }
        }
        // every point not examined yet is at least r cells away along one axis
        const double reach = (double)r * g.c * (1.0 - 1.0e-5);
        if (cnt == nw && (double)bd[nw - 1] < reach * reach) break;
        if (reach > 10.0 + g.c) break;
    }
    if (!found) { mask[p] = 0; return; }
    mask[p] = 1;
    // weights (:70-86): w = 1 / max(dist, 1e-10), normalised by the sum over ALL nw slots (empty slots hold 1e30 -> 1e-30)
    float total = 0.0f;
    for (int n = 0; n < nw; n++) {
        float w = (n < cnt) ? bd[n] : 1.0e30f;
        w = fmaxf(w, 1.0e-10f);
        w = 1.0f / w;
        bd[n] = w;
        total = total + w;
    }
    for (int n = 0; n < nw; n++) {
        wts[(long)n * ndst + p] = bd[n] / total;
        idx[(long)n * ndst + p] = (n < cnt) ? bi[n] + 1 : 0;   // 1-based, 0 = empty (the reference's idxs)
    }
}

// ez_applywgts.inc:19-48 for one target point per thread, every field (see oracle/fortran_shim.c: ez_applywgts_ for the two
// reference quirks reproduced here: the fill covers the first N_src output elements only, and the last column of every
// target row reads the lists of the element one row further)
__global__ void __launch_bounds__(TPB) k_cloud_apply(const int *__restrict__ idx, const float *__restrict__ wts, const int *__restrict__ mask, int nw,
                                                     const float *__restrict__ zin, float *__restrict__ zout, long in_stride, long out_stride,
                                                     int nfields, int nsrc, int ndst, int ni_dst, int use_value, float value,
                                                     const unsigned int *__restrict__ mm) {
    const int p = blockIdx.x * TPB + threadIdx.x;
    if (p >= ndst) return;
    const int m = mask[p];
    if (m != 1 && p >= nsrc) return;   // beyond the declared extent of outfld: left as the caller had it
    const long e0 = ((p + 1) % ni_dst == 0) ? (long)p + ni_dst : (long)p;
    const long total = (long)ndst * nw;
    for (int f = 0; f < nfields; f++) {
        const float *z = zin + (long)f * in_stride;
        float v;
        if (m == 1) {
            v = 0.0f;
            for (int n = 0; n < nw; n++) {
                const long e = e0 + (long)n * ndst;
                if (e >= total) break;
                const int id = __ldg(idx + e);
                if (id < 1) break;
                v = v + __ldg(wts + e) * __ldg(z + id - 1);
            }
        } else if (use_value) {
            v = value;
        } else {
            const float rmin = ord2f(mm[2 * f]), rmax = ord2f(mm[2 * f + 1]);
            v = rmin - 0.1f * (rmax - rmin);
        }
        zout[(long)f * out_stride + p] = v;
    }
}

}  // namespace

namespace ezk {

// Builds idx / wts ([nw][ndst], plane-major like the reference's wts(ni, nj, n)) and mask [ndst] on the device.
bool cloud_build(const float *src_lat, const float *src_lon, int nsrc, const float *dst_lat, const float *dst_lon, int ndst, int nw,
                 int *idx, float *wts, int *mask, cudaStream_t s) {
    if (nw < 1 || nw > MAXW || nsrc < 1 || ndst < 1) return false;
    unsigned int *d_mm = nullptr, *keys_in = nullptr, *keys_out = nullptr;
    unsigned char *d_occ = nullptr;
    float *d_bbox = nullptr;
    int *vals_in = nullptr, *vals_out = nullptr;
    void *tmp = nullptr;
    bool ok = false;
    do {
        if (cudaMalloc((void **)&d_mm, 4 * sizeof(unsigned int)) != cudaSuccess || cudaMalloc((void **)&d_bbox, 4 * sizeof(float)) != cudaSuccess) break;
        // bounding box of the source coordinates: [lon_min, lon_max, lat_min, lat_max]
        k_cloud_minmax_init<<<1, 32, 0, s>>>(d_mm, 4);
        const int nb = nsrc < 592 * TPB ? (nsrc + TPB - 1) / TPB : 592;
        k_cloud_minmax<<<dim3(nb, 1), TPB, 0, s>>>(src_lon, 0, nsrc, d_mm);
        k_cloud_minmax<<<dim3(nb, 1), TPB, 0, s>>>(src_lat, 0, nsrc, d_mm + 2);
        k_cloud_decode<<<1, 32, 0, s>>>(d_mm, d_bbox, 4);
        float bbox[4];
        if (cudaMemcpyAsync(bbox, d_bbox, sizeof(bbox), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) break;
        // cells sized for a handful of source points each, 360 / ncx degrees wide
        const double area = fmax(1e-6, (double)(bbox[1] - bbox[0])) * fmax(1e-6, (double)(bbox[3] - bbox[2]));
        double c = sqrt(4.0 * area / nsrc);
        c = fmin(5.0, fmax(0.02, c));
        CellGrid g;
        g.ncx = (int)(360.0 / c);
        if (g.ncx < 72) g.ncx = 72;
        g.c = 360.0 / g.ncx;
        g.ncy = (int)ceil(180.0 / g.c);
        if (cudaMalloc((void **)&keys_in, (size_t)nsrc * 4) != cudaSuccess || cudaMalloc((void **)&keys_out, (size_t)nsrc * 4) != cudaSuccess ||
            cudaMalloc((void **)&vals_in, (size_t)nsrc * 4) != cudaSuccess || cudaMalloc((void **)&vals_out, (size_t)nsrc * 4) != cudaSuccess)
            break;
        if (cudaMalloc((void **)&d_occ, OCC_NX * OCC_NY) != cudaSuccess) break;
        cudaMemsetAsync(d_occ, 0, OCC_NX * OCC_NY, s);
        k_cloud_keys<<<(nsrc + TPB - 1) / TPB, TPB, 0, s>>>(src_lat, src_lon, nsrc, g, keys_in, vals_in, d_occ);
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, nsrc, 0, 32, s);
        if (cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16) != cudaSuccess) break;
        if (cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, nsrc, 0, 32, s) != cudaSuccess) break;   // stable: indices ascend within a cell
        k_cloud_knn<<<(ndst + 127) / 128, 128, 0, s>>>(src_lat, src_lon, nsrc, keys_out, vals_out, g, dst_lat, dst_lon, ndst, nw, d_bbox, d_occ, idx, wts, mask);
        ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess;
    } while (0);
    if (!ok) (void)cudaGetLastError();
    cudaFree(d_mm); cudaFree(d_bbox); cudaFree(d_occ); cudaFree(keys_in); cudaFree(keys_out); cudaFree(vals_in); cudaFree(vals_out); cudaFree(tmp);
    return ok;
}

// scratch: 2 * nfields unsigned ints (min / max of every source field, used when the extrapolation is not `value`)
int cloud_apply(const int *idx, const float *wts, const int *mask, int nw, const float *zin, float *zout, long in_stride, long out_stride, int nfields,
                int nsrc, int ndst, int ni_dst, int use_value, float value, unsigned int *scratch, cudaStream_t s) {
    int launches = 0;
    if (!use_value) {
        k_cloud_minmax_init<<<(2 * nfields + 127) / 128, 128, 0, s>>>(scratch, 2 * nfields);
        const int nb = nsrc < 592 * TPB ? (nsrc + TPB - 1) / TPB : 592;
        k_cloud_minmax<<<dim3(nb, nfields), TPB, 0, s>>>(zin, in_stride, nsrc, scratch);
        launches += 2;
    }
    k_cloud_apply<<<(ndst + TPB - 1) / TPB, TPB, 0, s>>>(idx, wts, mask, nw, zin, zout, in_stride, out_stride, nfields, nsrc, ndst, ni_dst, use_value, value, scratch);
    return launches + 1;
}

}  // namespace ezk
