"""Synthetic instantiations of BASELINE.json's five configurations, plus reduced
variants of the same shapes for parity tests.  Everything here only talks to an ``EzLib``-style object (the
reference's C API by name), so the same builders define the grids in libezb200.so and -- in tests and in the
CPU-baseline leg of bench.py -- in the oracle library.

Fields are analytic and deterministic:
    f_k = 280 + 30 sin(2 lat) cos(3 lon + 0.1 k) + 5 cos(7 lat)
evaluated at the source grid's own (lat, lon) in ITS frame (rotated frame for Z/E grids) -- any smooth
function of the source indices serves; what matters is that both libraries are fed the same array.
"""
import numpy as np


def field(lat_deg, lon_deg, k=0, dtype=np.float32):
    la = np.deg2rad(np.asarray(lat_deg, np.float64))
    lo = np.deg2rad(np.asarray(lon_deg, np.float64))
    return (280.0 + 30.0 * np.sin(2 * la) * np.cos(3 * lo + 0.1 * k) + 5.0 * np.cos(7 * la)).astype(dtype)


def wind_fields(lat_deg, lon_deg, k=0):
    """Smooth, non-degenerate (u, v) pair."""
    la = np.deg2rad(np.asarray(lat_deg, np.float64))
    lo = np.deg2rad(np.asarray(lon_deg, np.float64))
    u = 10.0 * np.cos(la) * np.sin(2 * lo + 0.05 * k) + 3.0
    v = 8.0 * np.sin(2 * la) * np.cos(lo) - 2.0 + 0.1 * k
    return u.astype(np.float32), v.astype(np.float32)


class Case:
    """One (source grid, target grid, options) configuration defined in a given library."""

    def __init__(self, name, lib):
        self.name = name
        self.lib = lib
        self.gdin = self.gdout = -1
        self.src_shape = self.dst_shape = None     # (nj, ni) C-order shapes of one field
        self.src_lat = self.src_lon = None         # coordinates used to synthesise the source field
        self.opts = {}
        self.vector = False

    @property
    def n_src(self):
        return int(np.prod(self.src_shape))

    @property
    def n_out(self):
        return int(np.prod(self.dst_shape))

    def apply_options(self):
        lib = self.lib
        # back to the reference defaults first (ez_statics.c:17), then this case's options
        lib.ezsetopt("interp_degree", "cubic")
        lib.ezsetopt("extrap_degree", "maximum")
        lib.ezsetopt("polar_correction", "yes")
        lib.ezsetopt("use_1subgrid", "no")
        lib.ezsetopt("cloud_interp_alg", "distance")
        lib.ezsetval("extrap_value", 0.0)
        for k, v in self.opts.items():
            if k == "extrap_value":
                lib.ezsetval(k, v)
            else:
                lib.ezsetopt(k, v)

    def defset(self):
        self.apply_options()
        return self.lib.ezdefset(self.gdout, self.gdin)

    def source(self, k=0):
        return field(self.src_lat, self.src_lon, k).reshape(self.src_shape)

    def source_uv(self, k=0):
        u, v = wind_fields(self.src_lat, self.src_lon, k)
        return u.reshape(self.src_shape), v.reshape(self.src_shape)


def _ll_grid(ni, nj, lat0, lon0, dlat, dlon):
    lon = lon0 + dlon * np.arange(ni)
    lat = lat0 + dlat * np.arange(nj)
    return np.meshgrid(lon, lat)   # each (nj, ni)


def def_L(lib, ni, nj, lat0, lon0, dlat, dlon):
    ig = lib.cxgaig("L", lat0, lon0, dlat, dlon)
    return lib.ezqkdef(ni, nj, "L", *ig)


def def_PS(lib, ni, nj, pi, pj, d60, dgrw, hem="N"):
    ig = lib.cxgaig(hem, pi, pj, d60, dgrw)
    return lib.ezqkdef(ni, nj, hem, *ig)


def def_ZE(lib, ax, ay, xlat1, xlon1, xlat2, xlon2):
    ig = lib.cxgaig("E", xlat1, xlon1, xlat2, xlon2)
    ax = np.asarray(ax, np.float32)
    ay = np.asarray(ay, np.float32)
    return lib.ezgdef_fmem(len(ax), len(ay), "Z", "E", *ig, ax, ay)


# ------------------------------------------------------------------------------------------------
# C1: 'L' 0.5 deg global (720x361) -> polar-stereographic 'N', bilinear
# ------------------------------------------------------------------------------------------------
def c1(lib, n=2048, degree="linear"):
    c = Case("C1_L720x361_to_N%d_%s" % (n, degree), lib)
    c.gdin = lib.ezqkdef(720, 361, "L", 50, 50, 0, 0)
    # pole-centred; d60 chosen so the n x n square reaches ~ the equator at mid-edge, as in SURVEY.md (2048 -> 11.6 km)
    d60 = 11610.0 * 2048.0 / n
    c.gdout = def_PS(lib, n, n, (n + 1) / 2.0, (n + 1) / 2.0, d60, 0.0, "N")
    c.src_shape, c.dst_shape = (361, 720), (n, n)
    c.src_lon, c.src_lat = _ll_grid(720, 361, -90.0, 0.0, 0.5, 0.5)
    c.opts = {"interp_degree": degree}
    return c


# ------------------------------------------------------------------------------------------------
# C2: GEM Yin-Yang 'U' (2 x Z/E panels) -> 'L' lat-lon, cubic
# ------------------------------------------------------------------------------------------------
YIN_ROT = (0.0, 180.0, 0.0, 270.0)
YANG_ROT = (0.0, 0.0, 90.0, 0.0)


def yy_axes(res):
    """Panel axes: ax 42 ... 318, ay -48 ... 48 (a 3-degree halo around the 45..315 x -45..45 core; res=0.15 gives 1841 x 641)."""
    ni = int(round(276.0 / res)) + 1
    nj = int(round(96.0 / res)) + 1
    ax = 42.0 + res * np.arange(ni)
    ay = -48.0 + res * np.arange(nj)
    return ax.astype(np.float32), ay.astype(np.float32)


def def_U(lib, res):
    ax, ay = yy_axes(res)
    yin = def_ZE(lib, ax, ay, *YIN_ROT)
    yan = def_ZE(lib, ax, ay, *YANG_ROT)
    u = lib.ezgdef_supergrid(len(ax), 2 * len(ay), "U", "F", 1, [yin, yan])
    return u, yin, yan, ax, ay


def c2(lib, res_src=0.15, res_dst=0.1, degree="cubic"):
    c = Case("C2_U%g_to_L%g_%s" % (res_src, res_dst, degree), lib)
    u, yin, yan, ax, ay = def_U(lib, res_src)
    c.gdin = u
    c.subgrids = (yin, yan)
    ni = int(round(360.0 / res_dst))
    nj = int(round(180.0 / res_dst)) + 1
    c.gdout = def_L(lib, ni, nj, -90.0, 0.0, res_dst, res_dst)
    c.src_shape, c.dst_shape = (2 * len(ay), len(ax)), (nj, ni)
    lon, lat = np.meshgrid(ax.astype(np.float64), ay.astype(np.float64))
    # Yin then Yang, each synthesised on its own rotated frame (with an offset so the panels differ)
    c.src_lon = np.concatenate([lon, lon + 30.0], axis=0)
    c.src_lat = np.concatenate([lat, lat], axis=0)
    c.opts = {"interp_degree": degree}
    return c


# ------------------------------------------------------------------------------------------------
# C3: rotated 'Z'/E LAM (HRDPS-like) -> 'L', winds
# ------------------------------------------------------------------------------------------------
LAM_ROT = (36.0, 250.0, 45.0, 290.0)


def c3(lib, scale=1, degree="cubic", extrap=None):
    ni, nj = 2576 // scale, 1456 // scale
    d = 0.0225 * scale
    c = Case("C3_ZE%dx%d_to_L_%s" % (ni, nj, degree), lib)
    ax = (180.0 - d * (ni - 1) / 2.0) + d * np.arange(ni)
    ay = (-d * (nj - 1) / 2.0) + d * np.arange(nj)
    c.gdin = def_ZE(lib, ax, ay, *LAM_ROT)
    no, mo = 3200 // scale, 1600 // scale
    dd = 0.02 * scale
    c.gdout = def_L(lib, no, mo, 36.0 - dd * mo / 2.0, 250.0 - dd * no / 2.0, dd, dd)
    c.src_shape, c.dst_shape = (nj, ni), (mo, no)
    c.src_lon, c.src_lat = np.meshgrid(ax, ay)
    c.opts = {"interp_degree": degree}
    if extrap:
        c.opts["extrap_degree"] = extrap
    c.vector = True
    return c


# ------------------------------------------------------------------------------------------------
# C4: Gaussian 'G' -> 'Y' cloud of scattered points
# ------------------------------------------------------------------------------------------------
def c4(lib, ni=1600, nj=800, npts=1_000_000, degree="linear", seed=1234):
    c = Case("C4_G%dx%d_to_Y%d_%s" % (ni, nj, npts, degree), lib)
    c.gdin = lib.ezqkdef(ni, nj, "G", 0, 0, 0, 0)
    rng = np.random.default_rng(seed)
    lon = rng.uniform(0.0, 360.0, npts).astype(np.float32)
    lat = np.rad2deg(np.arcsin(rng.uniform(-1.0, 1.0, npts))).astype(np.float32)
    igl = lib.cxgaig("L", 0.0, 0.0, 1.0, 1.0)
    c.gdout = lib.ezgdef_fmem(npts, 1, "Y", "L", *igl, lon, lat)
    c.cloud = (lat, lon)
    c.src_shape, c.dst_shape = (nj, ni), (1, npts)
    # Gaussian latitudes are not needed to synthesise a field: use equally spaced stand-ins
    c.src_lon, c.src_lat = _ll_grid(ni, nj, -90.0 + 90.0 / nj, 0.0, 180.0 / nj, 360.0 / ni)
    c.opts = {"interp_degree": degree, "extrap_degree": "value", "extrap_value": -999.0}
    return c


# ------------------------------------------------------------------------------------------------
# C5: global unrotated 'Z'/E (repeated meridian) -> 'L', cubic
# ------------------------------------------------------------------------------------------------
def c5(lib, res_src=0.15, res_dst=0.1, degree="cubic"):
    c = Case("C5_Zglobal%g_to_L%g_%s" % (res_src, res_dst, degree), lib)
    ni = int(round(360.0 / res_src)) + 1
    nj = int(round(180.0 / res_src))
    ax = res_src * np.arange(ni)
    ay = (-90.0 + res_src / 2.0) + res_src * np.arange(nj)
    c.gdin = def_ZE(lib, ax, ay, *YIN_ROT)   # (0,180,0,270): identity rotation
    no = int(round(360.0 / res_dst))
    mo = int(round(180.0 / res_dst)) + 1
    c.gdout = def_L(lib, no, mo, -90.0, 0.0, res_dst, res_dst)
    c.src_shape, c.dst_shape = (nj, ni), (mo, no)
    c.src_lon, c.src_lat = np.meshgrid(ax, ay)
    c.opts = {"interp_degree": degree}
    return c
