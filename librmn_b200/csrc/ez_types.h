// librmn_b200/csrc/ez_types.h -- plain-data descriptors shared by the host logic and the CUDA kernels.
//
// These mirror the *information content* of the reference's _Grille / _gridset / _groptions
// (src/interp/ez_def.h:114-243) but are laid out for a device-resident gridset: every array a kernel
// touches lives in HBM, and every scalar a kernel needs travels by value in a small POD descriptor.
#pragma once
#include <stdint.h>

// zone classes: one byte per target point instead of the reference's five compacted index lists
// (ez_defzones.c:83-109).  The value is the *last writer* in ez_corrval's overwrite order
// (ez_corrval.c:54,121,125,130-143): DEHORS, AU_NORD, AU_SUD, POLE_NORD, POLE_SUD.
// Stored as membership bits (a point can sit in AU_NORD and POLE_NORD at once; scalars take the last
// writer, vectors are corrected by AU_* membership, ez_corrvec.c:31-45); bit 7 carries the Yin-Yang panel.
enum EzZoneBits : uint8_t { ZB_DEHORS = 1, ZB_AU_NORD = 2, ZB_AU_SUD = 4, ZB_POLE_NORD = 8, ZB_POLE_SUD = 16, ZB_YANG = 128 };

// interp_degree / extrap_degree codes (ez_def.h:37-57)
enum { EZ_NEAREST = 0, EZ_LINEAR = 1, EZ_CUBIC = 3, EZ_AVERAGE = 4, EZ_SPH_AVERAGE = 5, EZ_MAXIMUM = 4, EZ_MINIMUM = 5, EZ_VALUE = 6, EZ_ABORT = 13 };

// hemispheric expansion folded into index math (replaces ez_xpngdag2/ez_xpngdb2 copies, K11)
enum { XPN_NONE = 0, XPN_AG_NORTH = 1, XPN_AG_SOUTH = 2, XPN_B_NORTH = 3, XPN_B_SOUTH = 4 };

// One grid, as a kernel sees it.
struct GridDev {
    char grtyp, grref;
    int ni, nj;                 // as given by the caller
    int i1, i2, j1, j2;         // logical (expanded) index range, ez_xpncof.c:48-226
    int ext;                    // wrap "extension" 0/1/2
    int xpn;                    // XPN_* (needs_expansion)
    int yinv;                   // fst.axe_y_inverse
    int hemisphere;             // 1 north / 2 south for N,S and Z-on-N/S
    int ig[4], igref[4];
    float xg[4];                // decoded descriptors of grtyp (cigaxg)
    float xgref[4];             // decoded descriptors of grref
    float llp[4];               // (lat0, lon0, dlat, dlon) used by ez_calclatlon's grll call (A/B/L: xg ; E: ez_calclatlon.c:64-67)
    float xyp[4];               // (xlat0, xlon0, dellat, dellon) in fp32 as ez_ll2rgd computes them for A/B/L
    float r[9], ri[9];          // ez_crot rotation (column-major 3x3) for 'E' / ref 'E'
    const float *ax, *ay;       // device axes (Z,#,G: ni / nj(+nj) values; Y: ni*nj each)
    const float *cx, *cy;       // device Newton tables cx(ni,6), cy(j1:j2,6)
};

// Per-call interpolation options (snapshot of the thread-local _groptions, passed by value).
struct InterpOpts {
    int degree;        // interp_degree
    int extrap;        // extrap_degree
    float extrap_value;
    int vector;        // groptions.vecteur
    int sym_sign;      // +1 SYM / -1 ANTISYM (hemispheric mirror sign)
    int polar;         // polar_correction
    int arith;         // 0 = reference evaluation order, 1 = fp64 weight form
};

// Per-field scalars produced by the pre-pass (K10): pole-row means and the DEHORS fudge value.
struct FieldScalars {
    float pole_n, pole_s, fudge, pad;
};

// One 32 x 32 tile of target points for the tiled gather kernel (ez_interp_tiled.cuh): the box of the source
// field its fast threads read (first column/row, width, height), the Yin-Yang panel it reads from, and counts.
struct TileDesc {
    int c0, r0, w, h;
    int panel, nfast;
    int pitch;  // window kernel: staged row pitch in floats (chosen per tile against bank conflicts)
    int kind;   // 0: 2x2 window kernel (ez_interp_tiled.cuh) ; 1: row-separable kernel (ez_interp_rowsep.cuh)
};
