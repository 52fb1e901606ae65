"""Reduced-size instances of the five BASELINE.json configurations used by the parity tests and the golden
fixtures (same grid types, wrap modes, zone handling and options as the full-size cases, a few thousand points)."""
import numpy as np

from librmn_b200 import configs

DEGREES = ("nearest", "linear", "cubic")


def small_cases(lib, degree):
    """Returns a list of Case objects (grids defined in `lib`) for one interpolation degree."""
    cs = [
        configs.c1(lib, n=96, degree=degree),
        configs.c2(lib, res_src=3.0, res_dst=2.5, degree=degree),
        configs.c3(lib, scale=32, degree=degree),
        configs.c4(lib, ni=96, nj=48, npts=4000, degree=degree),
        configs.c5(lib, res_src=3.0, res_dst=2.5, degree=degree),
    ]
    # the LAM case again as scalars with every extrapolation mode that ez_corrval distinguishes
    for ex in ("maximum", "minimum", "value", "nearest", "linear"):
        c = configs.c3(lib, scale=32, degree=degree, extrap=ex)
        c.vector = False
        c.name += "_scalar_extrap_" + ex
        c.opts["extrap_value"] = -5.0
        cs.append(c)
    # 'N' -> 'L' (polar-stereographic SOURCE: ez_vxyfll route, DEHORS over the southern hemisphere)
    c = configs.Case("PS_N64_to_L_%s" % degree, lib)
    c.gdin = configs.def_PS(lib, 64, 64, 32.5, 32.5, 400000.0, 10.0, "N")
    c.gdout = configs.def_L(lib, 72, 37, -90.0, 0.0, 5.0, 5.0)
    c.src_shape, c.dst_shape = (64, 64), (37, 72)
    jj, ii = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    c.src_lat, c.src_lon = jj * 2.0 - 60.0, ii * 5.0
    c.opts = {"interp_degree": degree, "extrap_degree": "value", "extrap_value": 1.5}
    cs.append(c)
    # global 'L' (wrap extension 2) -> 'S' polar stereographic
    c = configs.Case("L144x73_to_S48_%s" % degree, lib)
    c.gdin = configs.def_L(lib, 144, 73, -90.0, 0.0, 2.5, 2.5)
    c.gdout = configs.def_PS(lib, 48, 48, 24.5, 24.5, 300000.0, 0.0, "S")
    c.src_shape, c.dst_shape = (73, 144), (48, 48)
    c.src_lon, c.src_lat = configs._ll_grid(144, 73, -90.0, 0.0, 2.5, 2.5)
    c.opts = {"interp_degree": degree}
    cs.append(c)
    # global Gaussian 'G' -> global 'L' : irregular axes with wrap 2 and both polar caps
    c = configs.Case("G96x48_to_L_%s" % degree, lib)
    c.gdin = lib.ezqkdef(96, 48, "G", 0, 0, 0, 0)
    c.gdout = configs.def_L(lib, 90, 46, -90.0, 0.0, 4.0, 4.0)
    c.src_shape, c.dst_shape = (48, 96), (46, 90)
    c.src_lon, c.src_lat = configs._ll_grid(96, 48, -88.0, 0.0, 3.75, 3.75)
    c.opts = {"interp_degree": degree}
    cs.append(c)
    # winds from GLOBAL sources: the polar vector correction (ez_corrvec, ez_calcnpolarwind / ez_calcspolarwind)
    c = configs.c5(lib, res_src=3.0, res_dst=2.5, degree=degree)      # Z/E global: ax-weighted pole means, rotated-frame longitudes
    c.vector = True
    c.name += "_winds"
    cs.append(c)
    c = configs.Case("L144x73_to_L_winds_%s" % degree, lib)           # regular 'L' global, wrap extension 2
    c.gdin = configs.def_L(lib, 144, 73, -90.0, 0.0, 2.5, 2.5)
    c.gdout = configs.def_L(lib, 100, 91, -90.0, 0.0, 2.0, 3.6)
    c.src_shape, c.dst_shape = (73, 144), (91, 100)
    c.src_lon, c.src_lat = configs._ll_grid(144, 73, -90.0, 0.0, 2.5, 2.5)
    c.opts = {"interp_degree": degree}
    c.vector = True
    cs.append(c)
    c = configs.Case("G96x48_to_L_winds_%s" % degree, lib)            # Gaussian source: irregular strip at both caps
    c.gdin = lib.ezqkdef(96, 48, "G", 0, 0, 0, 0)
    c.gdout = configs.def_L(lib, 90, 46, -90.0, 0.0, 4.0, 4.0)
    c.src_shape, c.dst_shape = (48, 96), (46, 90)
    c.src_lon, c.src_lat = configs._ll_grid(96, 48, -88.0, 0.0, 3.75, 3.75)
    c.opts = {"interp_degree": degree}
    c.vector = True
    cs.append(c)
    return cs


def run_case(lib, case, k=0):
    """One c_ezsint / c_ezuvint call on host arrays. Returns (rc, stacked float32 outputs)."""
    case.defset()
    if case.vector:
        u, v = case.source_uv(k)
        uo = np.zeros(case.dst_shape, np.float32)
        vo = np.zeros(case.dst_shape, np.float32)
        rc = lib.ezuvint(uo, vo, u, v)
        return rc, np.stack([uo, vo])
    z = case.source(k)
    out = np.zeros(case.dst_shape, np.float32)
    rc = lib.ezsint(out, z)
    return rc, out[None]


def field_range(case):
    if case.vector:
        u, v = case.source_uv(0)
        return float(max(u.max() - u.min(), v.max() - v.min()))
    z = case.source(0)
    return float(z.max() - z.min())
