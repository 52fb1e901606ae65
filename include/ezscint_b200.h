/*
 * include/ezscint_b200.h -- C ABI of libezb200.so, the B200-native drop-in for librmn's EZ interpolator
 * hot path (c_ezdefset / c_ezsint / c_ezuvint / c_ezsetopt and the grid-definition calls they need).
 *
 * Every entry point below keeps the name, argument order, argument meaning and return codes of the
 * reference declaration it replaces; the reference line is cited beside it
 * (ref = /root/reference/src/PUBLIC_INCLUDES/rmn/ezscint.h unless another file is named).
 * Fortran bindings use the reference's mangling: trailing underscore, all arguments by reference, hidden
 * string lengths last as int32 (rpnmacros.h:21,27-45).
 *
 * Pointer convention (additive to the reference): every field / coordinate pointer may be a HOST pointer
 * (pageable or pinned; the library stages it through HBM, as a stock caller expects) or a DEVICE pointer
 * on the current CUDA device (used in place, no copies).  Grids, gridsets (x, y, zone bits), axes and
 * Newton tables stay resident in HBM between calls.  No torch types cross this boundary.
 */
#ifndef EZSCINT_B200_H
#define EZSCINT_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef const int32_t F2Cl;

/* ---- gridset selection ------------------------------------------------------------------------ */
int32_t c_ezdefset(int32_t gdout, int32_t gdin);                                   /* ref :11  (ezdefset.c:38)  -> 1 */
int32_t ezdefset_(int32_t *gdout, int32_t *gdin);                                  /* ref :10 */
int32_t c_ezgetgdin(void);                                                         /* ref :186 */
int32_t c_ezgetgdout(void);                                                        /* ref :187 */

/* ---- the hot path ----------------------------------------------------------------------------- */
int32_t c_ezsint(float *zout, float *zin);                                         /* ref :87  (ezsint.c:38)   -> 0 | 2 | 1 | -1 */
int32_t ezsint_(float *zout, float *zin);                                          /* ref :86 */
int32_t c_ezuvint(float *uuout, float *vvout, float *uuin, float *vvin);           /* ref :90  (ezuvint.c:32)  -> 0 | 2 | -1 */
int32_t ezuvint_(float *uuout, float *vvout, float *uuin, float *vvin);            /* ref :89 */
int32_t c_ezwdint(float *spdout, float *wdout, float *uuin, float *vvin);          /* ref :93  (ezwdint.c:36) */
int32_t ezwdint_(float *spdout, float *wdout, float *uuin, float *vvin);           /* ref :92 */
int     c_ezsint_mask(int *mask_out, int *mask_in);                                /* ref :164 (ez_mask.c:184) */
/* masked interpolation (ez_mask.c:127-181): c_ezsint / c_ezuvint, the mask transfer, then every target point whose transferred mask
 * is 0 takes the minimum of the interpolated field (lorenzo_mask_fill.inc, methode 2).  Not for 'U' grids.  c_ezsint_m / c_ezuvint_m
 * only print "currently not implemented", as in the reference (ez_mask.c:114-124). */
int     c_ezsint_mdm(float *zout, int *mask_out, float *zin, int *mask_in);                                   /* ref :158 */
int     c_ezuvint_mdm(float *uuout, float *vvout, int *mask_out, float *uuin, float *vvin, int *mask_in);     /* ref :161 */
int     ezsint_mdm_(float *zout, int *mask_out, float *zin, int *mask_in);                                    /* ref :157 */
int     ezuvint_mdm_(float *uuout, float *vvout, int *mask_out, float *uuin, float *vvin, int *mask_in);      /* ref :160 */
int     c_ezsint_m(float *zout, float *zin);                                                                  /* ref :152 */
int     c_ezuvint_m(float *uuout, float *vvout, float *uuin, float *vvin);                                    /* ref :155 */
int     ezsint_mask_(int *mask_out, int *mask_in);                                 /* ref :163 */

/* ---- options ---------------------------------------------------------------------------------- */
int32_t c_ezsetopt(char *option, char *value);                                     /* ref :78  (ezsetopt.c:59) -> 0 | -1 */
int32_t ezsetopt_(char *option, char *value, F2Cl lenoption, F2Cl lenvalue);       /* ref :77 */
int32_t c_ezsetval(char *option, float fvalue);                                    /* ref :84  (ezsetval.c:115) */
int32_t ezsetval_(char *option, float *fvalue, F2Cl lenoption);                    /* ref :83 */
int32_t c_ezsetival(char *option, int32_t ivalue);                                 /* ref :81  (ezsetval.c:121) */
int32_t ezsetival_(char *option, int32_t *ivalue, F2Cl lenoption);                 /* ref :80 */
int32_t c_ezgetopt(char *option, char *value);                                     /* ref :35  (ezgetopt.c:53) */
int32_t ezgetopt_(char *option, char *value, F2Cl lenoption, F2Cl lenvalue);       /* ref :34 */
int32_t c_ezgetval(char *option, float *value);                                    /* ref :41  (ezgetval.c:55) */
int32_t ezgetval_(char *option, float *value, F2Cl lenoption);                     /* ref :40 */
int32_t c_ezgetival(char *option, int32_t *value);                                 /* ref :38  (ezgetval.c:89) */
int32_t ezgetival_(char *option, int32_t *value, F2Cl lenoption);                  /* ref :37 */

/* ---- grid definition -------------------------------------------------------------------------- */
int32_t c_ezgdef_fmem(int32_t ni, int32_t nj, char *grtyp, char *grref, int32_t ig1, int32_t ig2, int32_t ig3, int32_t ig4,
                      float *ax, float *ay);                                       /* ref :28  (ezgdef_fmem.c:37) -> gdid */
int32_t ezgdef_fmem_(int32_t *ni, int32_t *nj, char *grtyp, char *grref, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4,
                     float *ax, float *ay, F2Cl lengrtyp, F2Cl lengrref);          /* ref :26 */
int32_t c_ezgdef_fll(int32_t ni, int32_t nj, float *lat, float *lon);              /* ref :24  (ezgdef_fll.c:34) 'Y' cloud from lat/lon -> gdid */
int32_t ezgdef_fll_(int32_t *ni, int32_t *nj, float *lat, float *lon);             /* ref :23 */
int32_t c_gdgaxes(int32_t gdid, float *ax, float *ay);                             /* ref :96  (gdgaxes.c:31) */
int32_t gdgaxes_(int32_t *gdid, float *ax, float *ay);                             /* ref :95 */
int32_t c_gdxpncf(int32_t gdin, int32_t *i1, int32_t *i2, int32_t *j1, int32_t *j2);   /* ref :117 (gdxpncf.c:33) logical index range */
int32_t gdxpncf_(int32_t *gdin, int32_t *i1, int32_t *i2, int32_t *j1, int32_t *j2);   /* ref :116 */
int32_t c_ezgdef(int32_t ni, int32_t nj, char *grtyp, char *grref, int32_t ig1, int32_t ig2, int32_t ig3, int32_t ig4,
                 float *ax, float *ay);                                            /* ref :15  (ezgdef.c:50); "FILE" sources -> c_ezgdef_ffile */
int32_t ezgdef_(int32_t *ni, int32_t *nj, char *grtyp, char *grref, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4,
                float *ax, float *ay, F2Cl lengrtyp, F2Cl lengrref);               /* ref :13 */
int32_t c_ezqkdef(int32_t ni, int32_t nj, char *grtyp, int32_t ig1, int32_t ig2, int32_t ig3, int32_t ig4, int32_t iunit);
                                                                                   /* ref :66  (ezqkdef.c:25) = c_ezgdef_ffile */
/* Grids defined by positional records ('>>', '^^', '^>') of an RPN standard file: grid types '#', 'Y', 'Z', 'U', ' '
 * (ezgdef_ffile.c:29-140, ez_lireEnrPositionnels.c:12-349).  The records are located and read through the FST layer of the process
 * -- c_fstinf / c_fstsui / c_fstprm / c_fstluk of the librmn that opened `iunit`, found with dlsym at the first such call, or
 * handed over with c_ezb200_set_fst_hooks (same signatures as rmn/fstd98.h:42-67).  Without them these grid types fail with -1. */
int32_t c_ezgdef_ffile(int32_t ni, int32_t nj, char *grtyp, int32_t ig1, int32_t ig2, int32_t ig3, int32_t ig4, int32_t iunit);   /* ref :19 */
int32_t ezgdef_ffile_(int32_t *ni, int32_t *nj, char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, int32_t *iunit,
                      F2Cl lengrtyp);                                              /* ref :17 */
int32_t c_ezb200_set_fst_hooks(void *c_fstinf, void *c_fstsui, void *c_fstprm, void *c_fstluk);
int32_t ezqkdef_(int32_t *ni, int32_t *nj, char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, int32_t *iunit,
                 F2Cl lengrtyp);                                                   /* ref :64 */
int32_t c_ezgdef_supergrid(int32_t ni, int32_t nj, char *grtyp, char *grref, int32_t vercode, int32_t nsubgrids, int32_t *subgrid);
                                                                                   /* ref :32  (ezgdef_supergrid.c:40) */
int32_t ezgdef_supergrid_(int32_t *ni, int32_t *nj, char *grtyp, char *grref, int32_t *vercode, int32_t *nsubgrids, int32_t *subgrid,
                          F2Cl lengrtyp, F2Cl lengrref);                           /* ref :31 */
int32_t c_gdrls(int32_t gdin);                                                     /* ref :75  (gdrls.c:88) */
int32_t gdrls_(int32_t *gdin);                                                     /* ref :74 */

/* ---- grid queries ----------------------------------------------------------------------------- */
int32_t c_ezgprm(int32_t gdid, char *grtyp, int32_t *ni, int32_t *nj, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4);  /* ref :51 */
int32_t ezgprm_(int32_t *gdid, char *grtyp, int32_t *ni, int32_t *nj, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4, F2Cl l);
int32_t c_ezgxprm(int32_t gdid, int32_t *ni, int32_t *nj, char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4,
                  char *grref, int32_t *ig1ref, int32_t *ig2ref, int32_t *ig3ref, int32_t *ig4ref);                             /* ref :57 */
int32_t ezgxprm_(int32_t *gdid, int32_t *ni, int32_t *nj, char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4,
                 char *grref, int32_t *ig1ref, int32_t *ig2ref, int32_t *ig3ref, int32_t *ig4ref, F2Cl l1, F2Cl l2);
int32_t c_ezget_nsubgrids(int32_t gdid);                                           /* ref :169 */
int32_t ezget_nsubgrids_(int32_t *gdid);
int32_t c_ezget_subgridids(int32_t gdid, int32_t *subgrid);                        /* ref :172 */
int32_t ezget_subgridids_(int32_t *gdid, int32_t *subgrid);
int32_t c_gdll(int32_t gdid, float *lat, float *lon);                              /* ref :62  (gdll.c:47) */
int32_t gdll_(int32_t *gdid, float *lat, float *lon);
int32_t c_gdxyfll(int32_t gdid, float *x, float *y, float *lat, float *lon, int32_t n);   /* ref :138 (gdxyfll.c:190) */
int32_t gdxyfll_(int32_t *gdid, float *x, float *y, float *lat, float *lon, int32_t *n);
int32_t c_gdllfxy(int32_t gdid, float *lat, float *lon, float *x, float *y, int32_t n);   /* ref :102 (gdllfxy.c:30) */
int32_t gdllfxy_(int32_t *gdid, float *lat, float *lon, float *x, float *y, int32_t *n);  /* ref :101 */

/* ---- scattered-point API: the same kernels at caller-supplied points (no zone correction, as in the reference) ---- */
int32_t c_gdxysval(int32_t gdin, float *zout, float *zin, float *x, float *y, int32_t n); /* ref :120 (gdxysval.c:50) */
int32_t gdxysval_(int32_t *gdin, float *zout, float *zin, float *x, float *y, int32_t *n); /* ref :119 */
int32_t c_gdllsval(int32_t gdid, float *zout, float *zin, float *lat, float *lon, int32_t n);   /* ref :108 (gdllsval.c:35) */
int32_t gdllsval_(int32_t *gdid, float *zout, float *zin, float *lat, float *lon, int32_t *n);  /* ref :107 */
int32_t c_gdxyvval(int32_t gdin, float *uuout, float *vvout, float *uuin, float *vvin, float *x, float *y, int32_t n);      /* ref :126 (gdxyvval.c:38) */
int32_t gdxyvval_(int32_t *gdin, float *uuout, float *vvout, float *uuin, float *vvin, float *x, float *y, int32_t *n);     /* ref :125 */
int32_t c_gdllvval(int32_t gdid, float *uuout, float *vvout, float *uuin, float *vvin, float *lat, float *lon, int32_t n);  /* ref :111 (gdllvval.c:37) */
int32_t gdllvval_(int32_t *gdid, float *uuout, float *vvout, float *uuin, float *vvin, float *lat, float *lon, int32_t *n); /* ref :110 */
int32_t c_gdxywdval(int32_t gdin, float *uuout, float *vvout, float *uuin, float *vvin, float *x, float *y, int32_t n);     /* ref :123 (gdxywdval.c:38) */
int32_t gdxywdval_(int32_t *gdin, float *uuout, float *vvout, float *uuin, float *vvin, float *x, float *y, int32_t *n);    /* ref :122 */
int32_t c_gdllwdval(int32_t gdid, float *uuout, float *vvout, float *uuin, float *vvin, float *lat, float *lon, int32_t n); /* ref :114 (gdllwdval.c:36) */
int32_t gdllwdval_(int32_t *gdid, float *uuout, float *vvout, float *uuin, float *vvin, float *lat, float *lon, int32_t *n);/* ref :113 */
int32_t c_gdwdfuv(int32_t gdid, float *uullout, float *vvllout, float *uuin, float *vvin, float *latin, float *lonin, int32_t npts);   /* ref :132 (gdwdfuv.c:100) */
int32_t gdwdfuv_(int32_t *gdid, float *uullout, float *vvllout, float *uuin, float *vvin, float *latin, float *lonin, int32_t *npts);  /* ref :131 */
int32_t c_gduvfwd(int32_t gdid, float *uugdout, float *vvgdout, float *uullin, float *vvllin, float *latin, float *lonin, int32_t npts);  /* ref :129 (gduvfwd.c:112) */
int32_t gduvfwd_(int32_t *gdid, float *uugdout, float *vvgdout, float *uullin, float *vvllin, float *latin, float *lonin, int32_t *npts); /* ref :128 */
int     c_gdsetmask(int gdid, int *mask);                                          /* ref :146 (ez_mask.c:67) */
int     gdsetmask_(int *gdid, int *mask);
int     c_gdgetmask(int gdid, int *mask);                                          /* ref :149 (ez_mask.c:89) */
int     gdgetmask_(int *gdid, int *mask);

/* ---- descriptor helpers callers use to build ig1..ig4 (Fortran ABI only in the reference) ------- */
void cxgaig_(char *grtyp, int32_t *ig1, int32_t *ig2, int32_t *ig3, int32_t *ig4,
             float *xg1, float *xg2, float *xg3, float *xg4, F2Cl l);              /* src/base/xgaig.f90:4 */
void cigaxg_(char *grtyp, float *xg1, float *xg2, float *xg3, float *xg4,
             const int32_t *ig1, const int32_t *ig2, const int32_t *ig3, const int32_t *ig4, F2Cl l);   /* src/base/igaxg.f90:5 */

/* ---- additive entry points (not in the reference) ---------------------------------------------- */
/* nfields fields laid out back to back (ni*nj of gdin / gdout each); x, y and zone bits are read once for the batch */
int32_t c_ezsint_batch(float *zout, float *zin, int32_t nfields);
int32_t c_ezuvint_batch(float *uuout, float *vvout, float *uuin, float *vvin, int32_t nfields);
/* CUDA stream (cudaStream_t as void*) used by the calling thread's launches; NULL = default stream */
int32_t c_ezb200_set_stream(void *stream);
/* 0 = reference evaluation order (bit-parity), 1 = fp64 weight form, 2 = fp32 difference form with the
   shared-memory tiled kernel where the target is a 2-D grid at least as fine as the source (default); per calling thread */
int32_t c_ezb200_set_arith(int32_t mode);
int32_t c_ezb200_get_arith(void);
/* copy the cached gridset of (gdout, gdin) to host arrays (any may be NULL): x, y [npts], zone bits [npts], counts[6] */
int32_t c_ezb200_get_gridset(int32_t gdout, int32_t gdin, float *x, float *y, uint8_t *zonebits, int32_t *counts);
/* number of CUDA kernels launched by the calling thread since the last call with reset != 0 */
int64_t c_ezb200_launch_count(int32_t reset);
/* kernels used by the calling thread's last interpolation launch: bit 0 per-point kernel, bit 1 window tile kernel,
   bit 2 per-point kernel on the compacted list of points the tiles leave out (*nlist = its length),
   bit 3 row-separable tile kernel, bit 4 direct kernel (calls of a few fields, c_ezb200_set_direct_max_fields),
   bit 5 averaging kernel (interp_degree = average / sph_average) */
int32_t c_ezb200_last_path(int32_t *nlist);
/* block until the calling thread's stream is idle ; returns 0 or -1 with the CUDA error on stderr */
/* Measurement aid: with timing on, every tiled interpolation launch records CUDA events on the session stream around its
 * kernels; c_ezb200_last_kernel_ms returns ms[3] = {row-separable tile kernel, window tile kernel, per-point list kernel}
 * of the LAST launch and tiles[3] = {tiles on the row-separable kernel, tiles on the window kernel, points on the list}.
 * After a direct-kernel launch (c_ezb200_last_path bit 4): ms[3] = {per-field scalars pre-pass, direct kernel, zone-list kernel}.
 * Returns -1 when the last launch was not timed. */
int32_t c_ezb200_set_timing(int32_t on);
int32_t c_ezb200_last_kernel_ms(float *ms, int32_t *tiles);
/* Output-row band (multi-GPU split of a call with fewer fields than GPUs): until reset with (-1, -1) the calling
 * thread's c_ezsint / c_ezuvint / *_batch calls produce only target rows [j0, j1) of the current set's target grid; the output
 * arrays hold (j1 - j0) * ni points per field.  Bands made of whole 32-row tile rows (librmn_b200.sharding.row_band) keep the
 * tile kernels and, stitched together, equal the full call bit for bit; any other band runs on the per-point kernel (same plans,
 * fp32 difference form about the stencil centre: equal to the full call within rounding of the local variation).
 * Not for 'U' targets. */
int32_t c_ezb200_set_rows(int32_t j0, int32_t j1);
/* One process, several GPUs: from now on the calling thread's c_ezsint_batch / c_ezuvint_batch calls on HOST arrays with at least
 * 2 n fields are cut into n contiguous blocks of fields, one per CUDA device (the current device and the next visible ones); each
 * device holds its own replica of the gridset (built on its first call) and moves its own block over its own PCIe link.
 * Returns the number of devices that will be used (clamped to the visible ones); 1 = single device (default).
 * The environment variable EZB200_DEVICES sets the initial value for every thread. */
int32_t c_ezb200_set_devices(int32_t n);
/* Kernel choice by field count: a call of at most n fields (c_ezsint = 1, c_ezuvint = 2 per pair, *_batch = their count) runs on
 * the direct kernel (one thread per target point, no per-tile set-up; for one (u, v) pair the grid-to-grid wind rotation is
 * applied before the store); larger calls run on the tile kernels, whose set-up is amortised over the levels.  Both evaluate the
 * same fp32 difference form from the same plans and agree within rounding of the field's local variation (~1e-7 of its range).
 * Scattered ('Y' cloud) targets, which have no tile path, take the direct kernel at any field count (in source-cell order).
 * Default 4 (environment: EZB200_DIRECT_MAX_FIELDS); 0 = never the direct kernel.  Per calling thread; returns the previous value. */
int32_t c_ezb200_set_direct_max_fields(int32_t n);
int32_t c_ezb200_sync(void);
const char *c_ezb200_version(void);

#ifdef __cplusplus
}
#endif
#endif
