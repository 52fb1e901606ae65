"""CPU tests that pin the ORACLE (oracle/_ref/libezref.so).

The reference ships no tests, golden vectors or fixtures for ezscint, so these are the
pins we can create: analytic identities every correct implementation of the reference's algorithms must
satisfy, known-answer descriptor encodings, inverse round trips, cross-checks between independent kernels
(regular vs irregular axes), and frozen vectors (tests/golden) that detect drift of the oracle itself.
"""
import os

import numpy as np
import pytest

import cases
from librmn_b200 import configs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ezref_small.npz")


def _xy(lib, gdin, gdout, shape):
    """Fractional source coordinates of every target point, through the reference's own public calls."""
    lat = np.zeros(shape, np.float32)
    lon = np.zeros(shape, np.float32)
    assert lib.gdll(gdout, lat, lon) == 0
    x = np.zeros(shape, np.float32)
    y = np.zeros(shape, np.float32)
    assert lib.gdxyfll(gdin, x, y, lat, lon.copy()) == 0
    return x, y, lat, lon


def _reset(lib):
    c = configs.Case("reset", lib)
    c.apply_options()


# ---------------------------------------------------------------------------------------------------
# descriptors: known answers and round trips (src/base/xgaig.f90, igaxg.f90)
# ---------------------------------------------------------------------------------------------------
def test_cxgaig_known_answers(ezref):
    assert ezref.cxgaig("L", -90.0, 0.0, 0.5, 0.5) == (50, 50, 0, 0)
    assert ezref.cxgaig("L", 0.0, 0.0, 1.0, 1.0) == (100, 100, 9000, 0)
    assert ezref.cxgaig("N", 200.5, 200.5, 40000.0, 0.0) == (2005, 2005, 0, 400)   # old-style: pj*10, pi*10, dgrw*100, d60/100
    assert ezref.cxgaig("A", 0.0, 0.0, 0.0, 0.0) == (0, 0, 0, 0)


@pytest.mark.parametrize("rot", [(0.0, 180.0, 0.0, 270.0), (0.0, 0.0, 90.0, 0.0), (36.0, 250.0, 45.0, 290.0), (-12.5, 33.75, 10.25, 359.5)])
def test_E_descriptor_roundtrip(ezref, rot):
    ig = ezref.cxgaig("E", *rot)
    back = ezref.cigaxg("E", *ig)
    assert np.allclose(back, rot, atol=1.0 / 80.0)   # 1/40 degree quantisation


def test_PS_new_style_roundtrip(ezref):
    # pi,pj outside the old-style range force the new encoding; decode reconstructs pi,pj through xyfll (fp32)
    ig = ezref.cxgaig("N", 1024.5, 1024.5, 11610.0, 0.0)
    assert ig[3] >= 32768
    pi, pj, d60, dgrw = ezref.cigaxg("N", *ig)
    assert d60 == 11600.0 and abs(dgrw) < 0.1          # d60 is stored in hectometres
    # the encoding keeps the lat/lon of grid point (1,1); with d60 quantised the pole moves to 1 + 1023.5 * 11610/11600
    expect = 1.0 + 1023.5 * 11610.0 / 11600.0
    # (1,1) itself is stored at 360/32767 x 180/16383 degree resolution: ~0.1 grid unit at this d60
    assert abs(pi - expect) < 0.25 and abs(pj - expect) < 0.25


# ---------------------------------------------------------------------------------------------------
# interpolation identities
# ---------------------------------------------------------------------------------------------------
def _poly_case(lib, kind):
    """A source grid of the given kind and an interior target; returns (case, x, y)."""
    if kind == "L":
        gdin = configs.def_L(lib, 60, 40, 10.0, 20.0, 1.0, 1.0)
        ax, ay = 20.0 + np.arange(60), 10.0 + np.arange(40)
    else:   # irregular axes, identity rotation ('E' ref with (0,180,0,270)), stretched spacing
        ax = 100.0 + np.cumsum(0.5 + 0.01 * np.arange(60))
        ay = -20.0 + np.cumsum(0.4 + 0.02 * np.arange(40))
        gdin = configs.def_ZE(lib, ax, ay, *configs.YIN_ROT)
    c = configs.Case("poly_" + kind, lib)
    c.gdin = gdin
    if kind == "L":
        c.gdout = configs.def_L(lib, 37, 23, 14.3, 25.7, 0.77, 1.13)
    else:
        c.gdout = configs.def_L(lib, 37, 23, -14.0, 105.0, 0.53, 0.71)
    c.src_shape, c.dst_shape = (40, 60), (23, 37)
    c.axes = (np.asarray(ax, np.float32).astype(np.float64), np.asarray(ay, np.float32).astype(np.float64))
    return c


@pytest.mark.parametrize("kind", ["L", "Z"])
def test_bilinear_reproduces_linear_functions(ezref, kind):
    _reset(ezref)
    c = _poly_case(ezref, kind)
    ax, ay = c.axes
    X, Y = np.meshgrid(ax, ay)
    zin = (3.0 + 0.25 * X - 0.5 * Y).astype(np.float32)
    ezref.ezsetopt("interp_degree", "linear")
    ezref.ezdefset(c.gdout, c.gdin)
    zout = np.zeros(c.dst_shape, np.float32)
    assert ezref.ezsint(zout, zin) == 0
    x, y, lat, lon = _xy(ezref, c.gdin, c.gdout, c.dst_shape)
    # position of each target point in axis units (piecewise-linear map from index space)
    xs = np.interp(x, 1 + np.arange(len(ax)), ax)
    ys = np.interp(y, 1 + np.arange(len(ay)), ay)
    expect = 3.0 + 0.25 * xs - 0.5 * ys
    assert np.abs(zout - expect).max() < 2e-5


@pytest.mark.parametrize("kind", ["L", "Z"])
def test_cubic_reproduces_cubic_polynomials(ezref, kind):
    _reset(ezref)
    c = _poly_case(ezref, kind)
    ax, ay = c.axes
    X, Y = np.meshgrid(ax - ax[0], ay - ay[0])
    f = lambda u, v: 1.0 + 0.1 * u - 0.2 * v + 0.01 * u * v + 0.003 * u ** 2 - 0.002 * v ** 3 + 0.0005 * u ** 3   # noqa: E731
    zin = f(X, Y).astype(np.float32)
    ezref.ezsetopt("interp_degree", "cubic")
    ezref.ezdefset(c.gdout, c.gdin)
    zout = np.zeros(c.dst_shape, np.float32)
    assert ezref.ezsint(zout, zin) == 0
    x, y, lat, lon = _xy(ezref, c.gdin, c.gdout, c.dst_shape)
    xs = np.interp(x, 1 + np.arange(len(ax)), ax) - ax[0]
    ys = np.interp(y, 1 + np.arange(len(ay)), ay) - ay[0]
    rng = float(zin.max() - zin.min())
    assert np.abs(zout - f(xs, ys)).max() < 2e-6 * rng + 1e-4


def test_nearest_returns_source_samples(ezref):
    _reset(ezref)
    c = _poly_case(ezref, "L")
    rs = np.random.default_rng(7)
    zin = rs.standard_normal(c.src_shape).astype(np.float32)
    ezref.ezsetopt("interp_degree", "nearest")
    ezref.ezdefset(c.gdout, c.gdin)
    zout = np.zeros(c.dst_shape, np.float32)
    ezref.ezsint(zout, zin)
    x, y, _, _ = _xy(ezref, c.gdin, c.gdout, c.dst_shape)
    ii = np.clip(np.floor(x + 0.5).astype(int), 1, 60) - 1
    jj = np.clip(np.floor(y + 0.5).astype(int), 1, 40) - 1
    assert np.array_equal(zout, zin[jj, ii])


@pytest.mark.parametrize("degree", ["linear", "cubic"])
def test_regular_and_irregular_kernels_agree_on_a_uniform_grid(ezref, degree):
    """The same uniform lat-lon grid defined as 'L' (ez_rgdint_*) and as 'Z' on an identity-rotated 'E' frame
    (ez_irgdint_*, Newton form): two independent code paths that must give the same interpolant."""
    _reset(ezref)
    ni, nj = 80, 50
    ax, ay = 150.0 + 0.5 * np.arange(ni), -12.0 + 0.5 * np.arange(nj)
    g_l = configs.def_L(ezref, ni, nj, -12.0, 150.0, 0.5, 0.5)
    g_z = configs.def_ZE(ezref, ax, ay, *configs.YIN_ROT)
    gout = configs.def_L(ezref, 41, 29, -8.0, 155.0, 0.37, 0.61)
    X, Y = np.meshgrid(np.deg2rad(ax), np.deg2rad(ay))
    zin = (280 + 30 * np.sin(6 * Y) * np.cos(5 * X)).astype(np.float32)
    ezref.ezsetopt("interp_degree", degree)
    outs = []
    for g in (g_l, g_z):
        ezref.ezdefset(gout, g)
        z = np.zeros((29, 41), np.float32)
        assert ezref.ezsint(z, zin) == 0
        outs.append(z)
    rng = float(zin.max() - zin.min())
    assert np.abs(outs[0] - outs[1]).max() < 2e-5 * rng


def test_constant_field_stays_constant_everywhere(ezref):
    """Weights sum to one, and the polar / outside fix-ups (pole mean, max/min fudge) keep a constant constant
    -- except 'maximum'/'minimum' extrapolation, which by definition adds 0 * 5 % of a zero range."""
    for deg in cases.DEGREES:
        for c in cases.small_cases(ezref, deg):
            if c.vector or c.opts.get("extrap_degree") == "value":
                continue
            c.defset()
            zin = np.full(c.src_shape, 7.25, np.float32)
            zout = np.zeros(c.dst_shape, np.float32)
            assert ezref.ezsint(zout, zin) >= 0
            assert np.abs(zout - 7.25).max() < 1e-5, c.name


def test_yinyang_panel_selection_geometry(ezref):
    """Field = 1 on Yin, 2 on Yang, nearest: the output must be 1 inside Yin's 45..315 x -45..45 core and 2 outside."""
    _reset(ezref)
    c = configs.c2(ezref, res_src=3.0, res_dst=2.5, degree="nearest")
    c.defset()
    nj2, ni = c.src_shape
    zin = np.ones(c.src_shape, np.float32)
    zin[nj2 // 2:] = 2.0
    zout = np.zeros(c.dst_shape, np.float32)
    assert ezref.ezsint(zout, zin) == 0
    lat = np.zeros(c.dst_shape, np.float32)
    lon = np.zeros(c.dst_shape, np.float32)
    ezref.gdll(c.gdout, lat, lon)
    core = (lon > 48) & (lon < 312) & (np.abs(lat) < 42)
    far = (np.abs(lat) > 48) | (lon < 42) | (lon > 318)
    assert set(np.unique(zout)) == {1.0, 2.0}
    assert (zout[core] == 1.0).all()
    assert (zout[far] == 2.0).all()


def test_wind_rotation_preserves_speed(ezref):
    """c_ezuvint on a rotated LAM: |(u,v)| after interpolation+rotation equals the speed of the scalar-interpolated
    components away from the extrapolated region (rotation is orthogonal)."""
    _reset(ezref)
    c = configs.c3(ezref, scale=16, degree="linear")
    rc, uv = cases.run_case(ezref, c)
    assert rc == 2
    cs = configs.c3(ezref, scale=16, degree="linear")
    cs.vector = False
    cs.defset()
    u, v = c.source_uv()
    us = np.zeros(c.dst_shape, np.float32)
    vs = np.zeros(c.dst_shape, np.float32)
    ezref.ezsint(us, u)
    ezref.ezsint(vs, v)
    inside = (uv[0] != 0) | (uv[1] != 0)
    spd_rot = np.hypot(uv[0], uv[1])[inside]
    # scalar DEHORS points were set to max+5%; compare only where the vector path did not extrapolate
    spd_sca = np.hypot(us, vs)[inside]
    assert np.abs(spd_rot - spd_sca).max() < 2e-4 * spd_sca.max()


def test_return_codes(ezref):
    _reset(ezref)
    c = configs.c3(ezref, scale=32, degree="linear", extrap="abort")
    c.vector = False
    c.defset()
    z = np.zeros(c.dst_shape, np.float32)
    assert ezref.ezsint(z, c.source()) == -1          # ez_corrval.c:56-60
    ezref.ezsetopt("extrap_degree", "value")
    assert ezref.ezsint(z, c.source()) == 2           # extrapolated
    assert ezref.ezdefset(c.gdin, c.gdin) == 1
    zz = np.zeros(c.src_shape, np.float32)
    assert ezref.ezsint(zz, c.source()) == 1          # gdin == gdout: plain copy (ezsint.c:56-60)
    assert np.array_equal(zz, c.source())
    assert ezref.ezsetopt("interp_degree", "quintic") == -1
    assert ezref.ezsetopt("no_such_option", "yes") == -1


# ---------------------------------------------------------------------------------------------------
# frozen vectors
# ---------------------------------------------------------------------------------------------------
def test_oracle_matches_golden_vectors(ezref):
    g = np.load(GOLDEN)
    worst = 0.0
    for deg in cases.DEGREES:
        for c in cases.small_cases(ezref, deg):
            rc, out = cases.run_case(ezref, c)
            assert rc == int(g[c.name + "__rc"]), c.name
            ref = g[c.name + "__out"]
            d = np.abs(out.astype(np.float64) - ref).max() / cases.field_range(c)
            worst = max(worst, d)
            assert d < 2e-6, "%s drifted by %g of range" % (c.name, d)
    assert worst >= 0.0


# ---------------------------------------------------------------------------------------------------
# interp_degree = average / sph_average (ez_avg.inc, ez_avg_sph.inc restated in oracle/fortran_shim.c)
# ---------------------------------------------------------------------------------------------------
def test_average_preserves_constants_and_means(ezref):
    """Pins of the box-average leaf: a constant field stays constant; at an exact 4:1 coarsening whose target cells are made of
    whole source cells plus half cells on their rim, the result is that weighted mean (computed here independently in numpy);
    sph_average at the equator row equals average (1 / cos(0) = 1)."""
    from librmn_b200 import configs
    try:
        ezref.ezsetopt("interp_degree", "average")
        gdin = configs.def_L(ezref, 360, 181, -90.0, 0.0, 1.0, 1.0)
        gdout = configs.def_L(ezref, 40, 21, -40.0, 20.0, 4.0, 4.0)
        assert ezref.ezdefset(gdout, gdin) == 1
        z = np.full((181, 360), 3.25, np.float32)
        o = np.zeros((21, 40), np.float32)
        assert ezref.ezsint(o, z) >= 0
        assert np.array_equal(o, np.full_like(o, 3.25))
        lon, lat = np.meshgrid(np.arange(360.0), np.arange(181.0) - 90.0)
        z = configs.field(lat, lon, k=1).astype(np.float32)
        assert ezref.ezsint(o, z) >= 0
        # interior target point (i, j): centre at source index (x, y) = (21 + 4 i, 51 + 4 j) (1-based), cell = +-2 source cells:
        # three whole cells and two half cells per axis
        w = np.array([0.5, 1.0, 1.0, 1.0, 0.5])
        zz = z.astype(np.float64)
        for (i, j) in ((3, 4), (17, 10), (30, 15)):
            ci, cj = 20 + 4 * i, 50 + 4 * j      # 0-based centre
            box = zz[cj - 2:cj + 3, ci - 2:ci + 3]
            want = (box * np.outer(w, w)).sum() / 16.0
            assert abs(o[j, i] - want) < 2e-5 * abs(want), (i, j, o[j, i], want)
        ezref.ezsetopt("interp_degree", "sph_average")
        o2 = np.zeros_like(o)
        assert ezref.ezsint(o2, z) >= 0
        assert np.array_equal(o2[10], o[10])          # target row 10 lies on the equator
        assert not np.array_equal(o2[18], o[18])      # 32 N: cells 1 / cos(32 deg) wider
    finally:
        ezref.ezsetopt("interp_degree", "cubic")
