"""An INDEPENDENT pin of the three Fortran leaves that carry the bicubic arithmetic (VERDICT r1, next #2d).

oracle/fortran_shim.c (C) and the CUDA reference-order kernels were restated by the same hand; a shared misreading
of an .inc file would pass every oracle-vs-GPU test.  This file transcribes the three leaves a second time, in numpy,
straight from the Fortran DECLARATIONS (which variable is real, which real*8, which statement function is typed
`real` and therefore rounds its result to fp32), vectorised over the points, and demands BIT equality with
libezshim.so on adversarial inputs: seam columns i in {1, ni-1, ni} in both wrap modes, clamped rows, points outside
the grid (extrapolating dx / dy), large-offset / small-range fields, irregular axes.

    ez_rgdint_3_w    /root/reference/src/interp/ez_rgdint_3_w.inc:64-108 + cubic8.cdk:1-3
    ez_irgdint_3_nw  /root/reference/src/interp/ez_irgdint_3_nw.inc:30-33,73-87 (typing), :89-165
    ez_irgdint_3_w   /root/reference/src/interp/ez_irgdint_3_w.inc:71-79 (typing), :86-229
"""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "libezshim.so")
f32, f64 = np.float32, np.float64


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(SHIM):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "shim"])
    return C.CDLL(SHIM)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i(v):
    return C.byref(C.c_int32(v))


def ifix(p):                       # Fortran IFIX: truncation toward zero
    return np.trunc(p).astype(np.int64)


# --------------------------------------------------------------------------------------------------
# cubic8.cdk:1-3 -- `real*8 cubic, dx, dy, z1..z4`; the literals are DEFAULT-REAL (fp32) constants promoted to real*8
# --------------------------------------------------------------------------------------------------
C6 = f64(f32(0.1666666666666))
C3 = f64(f32(0.3333333333333))
H = f64(f32(0.5))


def cubic(z1, z2, z3, z4, dx):
    return ((((z4 - z1) * C6 + H * (z2 - z3)) * dx + H * (z1 + z3) - z2) * dx + z3 - C6 * z4 - H * z2 - C3 * z1) * dx + z2


def np_rgdint_3_w(px, py, z, ni, j1, j2, wrap):
    """z[j - j1, i - 1] = z(i, j).  ez_rgdint_3_w.inc:70-104."""
    limite = ni + 2 - wrap
    i = np.minimum(ni - 2 + wrap, np.maximum(1, np.maximum(2 - wrap, ifix(px))))
    j = np.minimum(j2 - 2, np.maximum(j1 + 1, ifix(py)))
    if wrap > 0:
        im1 = (limite + i - 1) % limite
        ip1 = (limite + i + 1) % limite
        ip2 = (limite + i + 2) % limite
        im1 = np.where(im1 == 0, ni, im1)
        i = np.where(i == 0, ni, i)
        ip1 = np.where(ip1 == 0, ni, ip1)
        ip2 = np.where(ip2 == 0, ni, ip2)
        if wrap == 1:
            ip2 = np.where(ip2 == ni, 2, ip2)
            im1 = np.where(im1 == ni, ni - 1, im1)
    else:
        im1, ip1, ip2 = i - 1, i + 1, i + 2
    # dx = px(n) - i : real - integer is a REAL (fp32) subtraction, then stored in the real*8 dx
    dx = (px - i.astype(f32)).astype(f32).astype(f64)
    dy = (py - j.astype(f32)).astype(f32).astype(f64)
    Z = lambda ii, jj: z[jj - j1, ii - 1].astype(f64)            # dble(z(i, j))
    rows = [cubic(Z(im1, j + k), Z(i, j + k), Z(ip1, j + k), Z(ip2, j + k), dx) for k in (-1, 0, 1, 2)]
    return cubic(rows[0], rows[1], rows[2], rows[3], dy).astype(f32)   # zo is real


# --------------------------------------------------------------------------------------------------
# the Newton statement functions.  rnd = the type the functions are DECLARED with:
#   ez_irgdint_3_nw.inc:32  `real fa, fa2, fa3, fa4`    -> every result is rounded to fp32 (then widened again on assignment)
#   ez_irgdint_3_w.inc:78   `real*8 fa, fa2, fa3, fa4`  -> no rounding
# arguments are real*8 in both (dummy arguments a1..a4, x, c1..c6 real*8; x1..x3 real*8 in _nw, REAL in _w: the mixed
# expression (x - x1) is evaluated in real*8 either way)
# --------------------------------------------------------------------------------------------------
def newton(rnd):
    r = (lambda v: v.astype(f32).astype(f64)) if rnd else (lambda v: v)
    fa = lambda a1, a2, a3, a4, x, x1, x2, x3: r(a1 + (x - x1) * (a2 + (x - x2) * (a3 + a4 * (x - x3))))
    fa2 = lambda c1, a1, a2: r(c1 * (a2 - a1))
    fa3 = lambda c1, c2, c3, a1, a2, a3: r(c2 * (c3 * (a3 - a2) - c1 * (a2 - a1)))
    fa4 = lambda c1, c2, c3, c4, c5, c6, a1, a2, a3, a4: r(c4 * (c5 * (c6 * (a4 - a3) - c3 * (a3 - a2)) - c2 * (c3 * (a3 - a2) - c1 * (a2 - a1))))
    return fa, fa2, fa3, fa4


def newton_2d(fns, cxr, cyr, zrows, x, xs, y, ys):
    fa, fa2, fa3, fa4 = fns
    c = [cxr[:, k].astype(f64) for k in range(6)]
    b = []
    for z1, z2, z3, z4 in zrows:
        a1 = z1
        a2 = fa2(c[0], z1, z2)
        a3 = fa3(c[0], c[1], c[2], z1, z2, z3)
        a4 = fa4(c[0], c[1], c[2], c[3], c[4], c[5], z1, z2, z3, z4)
        b.append(fa(a1, a2, a3, a4, x, xs[0], xs[1], xs[2]))
    d = [cyr[:, k].astype(f64) for k in range(6)]
    b1, b2, b3, b4 = b
    b12 = fa2(d[0], b1, b2)
    b13 = fa3(d[0], d[1], d[2], b1, b2, b3)
    b14 = fa4(d[0], d[1], d[2], d[3], d[4], d[5], b1, b2, b3, b4)
    return fa(b1, b12, b13, b14, y, ys[0], ys[1], ys[2]).astype(f32)


def np_irgdint_3_nw(px, py, ax, ay, cx, cy, z, i1, i2, j1, j2):
    """ax[i - i1], ay[j - j1], cx[i - i1, k], z[j - j1, i - i1].  ez_irgdint_3_nw.inc:89-165."""
    i = np.minimum(i2 - 2, np.maximum(i1 + 1, ifix(px)))
    j = np.minimum(j2 - 2, np.maximum(j1 + 1, ifix(py)))
    AX = lambda ii: ax[ii - i1]
    AY = lambda jj: ay[jj - j1]
    # x = ax(i) + (ax(i+1)-ax(i))*(px(n)-i) : every operand is real -> evaluated in fp32, then stored in the real*8 x
    x = (AX(i) + ((AX(i + 1) - AX(i)).astype(f32) * (px - i.astype(f32)).astype(f32)).astype(f32)).astype(f32).astype(f64)
    y = (AY(j) + ((AY(j + 1) - AY(j)).astype(f32) * (py - j.astype(f32)).astype(f32)).astype(f32)).astype(f32).astype(f64)
    xs = [AX(i + k).astype(f64) for k in (-1, 0, 1, 2)]     # x1..x4 real*8 (:78)
    ys = [AY(j + k).astype(f64) for k in (-1, 0, 1, 2)]
    Z = lambda ii, jj: z[jj - j1, ii - i1].astype(f64)      # z1..z4 real*8 (:80)
    zrows = [(Z(i - 1, j + k), Z(i, j + k), Z(i + 1, j + k), Z(i + 2, j + k)) for k in (-1, 0, 1, 2)]
    return newton_2d(newton(True), cx[i - i1], cy[j - j1], zrows, x, xs, y, ys)


def np_irgdint_3_w(px, py, ax, ay, cx, cy, z, ni, j1, j2, wrap):
    """ez_irgdint_3_w.inc:93-229: x1..x4, y1..y4 are REAL (:75), the statement functions real*8 (:78)."""
    i = np.minimum(ni - 2 + wrap, np.maximum(1, np.maximum(2 - wrap, ifix(px))))
    j = np.minimum(j2 - 2, np.maximum(j1 + 1, ifix(py)))
    AX = lambda ii: ax[ii - 1]
    AY = lambda jj: ay[jj - j1]
    im1, ip1, ip2 = i - 1, i + 1, i + 2
    K360 = f32(360.0)
    # interior values first (the `else` branch, :151-154); out-of-range indices of seam points are overwritten below
    safe = lambda ii: np.clip(ii, 1, ni)
    x1, x2, x3, x4 = AX(safe(im1)), AX(safe(i)), AX(safe(ip1)), AX(safe(ip2))

    def put(mask, a_im1, a_ip1, a_ip2, v1, v2, v3, v4):
        nonlocal im1, ip1, ip2, x1, x2, x3, x4
        im1 = np.where(mask, a_im1, im1); ip1 = np.where(mask, a_ip1, ip1); ip2 = np.where(mask, a_ip2, ip2)
        x1 = np.where(mask, v1, x1); x2 = np.where(mask, v2, x2); x3 = np.where(mask, v3, x3); x4 = np.where(mask, v4, x4)

    one = np.ones_like(px)
    if wrap == 1:      # :101-120 -- repeated meridian: column ni duplicates column 1
        put(i == 1, ni - 1, 2, 3, (ax[ni - 2] - K360) * one, ax[0] * one, ax[1] * one, ax[2] * one)
        put(i == ni - 1, ni - 2, ni, 2, ax[ni - 3] * one, ax[ni - 2] * one, ax[ni - 1] * one, (ax[1] + K360) * one)
    elif wrap == 2:    # :121-150
        put(i == 1, ni, 2, 3, (ax[ni - 1] - K360) * one, ax[0] * one, ax[1] * one, ax[2] * one)
        put(i == ni - 1, ni - 2, ni, 1, ax[ni - 3] * one, ax[ni - 2] * one, ax[ni - 1] * one, (ax[0] + K360) * one)
        put(i == ni, ni - 1, 1, 2, ax[ni - 2] * one, ax[ni - 1] * one, (ax[0] + K360) * one, (ax[1] + K360) * one)
    x1, x2, x3, x4 = (v.astype(f32) for v in (x1, x2, x3, x4))
    # x = x2 + (x3-x2)*(px(n)-i) : all REAL -> fp32, stored in the real*8 x (:156) ; y likewise (:157)
    x = (x2 + ((x3 - x2).astype(f32) * (px - i.astype(f32)).astype(f32)).astype(f32)).astype(f32).astype(f64)
    y = (AY(j) + ((AY(j + 1) - AY(j)).astype(f32) * (py - j.astype(f32)).astype(f32)).astype(f32)).astype(f32).astype(f64)
    xs = [v.astype(f64) for v in (x1, x2, x3, x4)]           # REAL values entering real*8 expressions
    ys = [AY(j + k).astype(f32).astype(f64) for k in (-1, 0, 1, 2)]
    Z = lambda ii, jj: z[jj - j1, ii - 1].astype(f64)
    zrows = [(Z(im1, j + k), Z(i, j + k), Z(ip1, j + k), Z(ip2, j + k)) for k in (-1, 0, 1, 2)]
    # the table row is cx(i, .) of the stencil's own i, also on the seam (:166-169)
    return newton_2d(newton(False), cx[i - 1], cy[j - j1], zrows, x, xs, y, ys)


# --------------------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------------------
def recip_tables(nodes_of):
    """c(., 1..6) as the header comments define them (ez_irgdint_3_nw.inc:62-67), in fp32, from four fp32 nodes per row."""
    x1, x2, x3, x4 = (np.asarray(v, f32) for v in nodes_of)
    one = f32(1.0)
    return np.stack([one / (x2 - x1), one / (x3 - x1), one / (x3 - x2), one / (x4 - x1), one / (x4 - x2), one / (x4 - x3)], axis=1).astype(f32)


def fields(rs, nj, ni):
    jj, ii = np.meshgrid(np.arange(nj), np.arange(ni), indexing="ij")
    smooth = (280.0 + 30.0 * np.sin(0.37 * jj) * np.cos(0.29 * ii + 0.3)).astype(f32)
    offset = (1.0e5 + 0.5 * np.sin(0.11 * ii * jj)).astype(f32)          # large offset, small range
    noise = rs.standard_normal((nj, ni)).astype(f32)
    return {"smooth": smooth, "offset": offset, "noise": noise}


def points(rs, ni, j1, j2, n=4000):
    px = rs.uniform(-1.5, ni + 2.5, n).astype(f32)
    py = rs.uniform(j1 - 1.5, j2 + 1.5, n).astype(f32)
    # every seam column, exact nodes, just below / above integers
    ex = np.array([1.0, 1.5, 1.999999, 2.0, ni - 2.0, ni - 1.000001, ni - 1.0, ni - 0.5, ni, ni + 0.25, ni + 0.999, 0.5, 0.999999], f32)
    ey = np.array([j1, j1 + 0.5, j1 + 1, j1 + 1.5, j2 - 2, j2 - 1.5, j2 - 1, j2], f32)
    gx, gy = np.meshgrid(ex, ey)
    return np.concatenate([px, gx.ravel()]).astype(f32), np.concatenate([py, gy.ravel()]).astype(f32)


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


@pytest.mark.parametrize("wrap", [0, 1, 2])
def test_rgdint_3_w_bit_for_bit(shim, wrap):
    rs = np.random.default_rng(100 + wrap)
    ni, j1, j2 = 37, 1, 23
    px, py = points(rs, ni, j1, j2)
    for name, z in fields(rs, j2 - j1 + 1, ni).items():
        zo = np.zeros(px.size, f32)
        shim.ez_rgdint_3_w_(_p(zo), _p(px), _p(py), _i(px.size), _p(z), _i(ni), _i(j1), _i(j2), _i(wrap))
        ref = np_rgdint_3_w(px, py, z, ni, j1, j2, wrap)
        assert np.array_equal(bits(zo), bits(ref)), "%s wrap=%d: %d of %d differ" % (name, wrap, int((bits(zo) != bits(ref)).sum()), zo.size)


def test_irgdint_3_nw_bit_for_bit(shim):
    rs = np.random.default_rng(7)
    i1, i2, j1, j2 = 1, 41, 1, 29
    ni, nj = i2 - i1 + 1, j2 - j1 + 1
    ax = np.cumsum(rs.uniform(0.4, 1.7, ni)).astype(f32)
    ay = (-40.0 + np.cumsum(rs.uniform(0.3, 2.1, nj))).astype(f32)
    idx = lambda a, n: [np.concatenate([a[:1], a[:-1]]), a, np.concatenate([a[1:], a[-1:]]), np.concatenate([a[2:], a[-2:]])]
    # rows without a full stencil hold 1.0 (ez_nwtncof.inc:83-90): they are never addressed by the clamped i, j
    cx = np.ones((ni, 6), f32); cy = np.ones((nj, 6), f32)
    cx[1:ni - 2] = recip_tables([ax[0:ni - 3], ax[1:ni - 2], ax[2:ni - 1], ax[3:ni]])
    cy[1:nj - 2] = recip_tables([ay[0:nj - 3], ay[1:nj - 2], ay[2:nj - 1], ay[3:nj]])
    px, py = points(rs, ni, j1, j2)
    cxf, cyf = np.asfortranarray(cx), np.asfortranarray(cy)      # Fortran cx(i1:i2, 6)
    for name, z in fields(rs, nj, ni).items():
        zo = np.zeros(px.size, f32)
        shim.ez_irgdint_3_nw_(_p(zo), _p(px), _p(py), _i(px.size), _p(ax), _p(ay), cxf.ctypes.data_as(C.c_void_p), cyf.ctypes.data_as(C.c_void_p),
                              _p(z), _i(i1), _i(i2), _i(j1), _i(j2))
        ref = np_irgdint_3_nw(px, py, ax, ay, cx, cy, z, i1, i2, j1, j2)
        assert np.array_equal(bits(zo), bits(ref)), "%s: %d of %d differ" % (name, int((bits(zo) != bits(ref)).sum()), zo.size)


@pytest.mark.parametrize("wrap", [1, 2])
def test_irgdint_3_w_bit_for_bit(shim, wrap):
    rs = np.random.default_rng(40 + wrap)
    j1, j2 = 1, 27
    nj = j2 - j1 + 1
    if wrap == 1:     # global axis with the repeated meridian: ax(ni) = ax(1) + 360
        ni = 49
        ax = np.concatenate([np.sort(rs.uniform(0.0, 359.0, ni - 2)), [360.0]])
        ax = np.concatenate([[0.0], ax]).astype(f32)
    else:             # global axis, no repeated column
        ni = 48
        ax = np.sort(rs.uniform(0.0, 359.5, ni)).astype(f32)
    ay = (-80.0 + np.cumsum(rs.uniform(1.0, 9.0, nj))).astype(f32)
    K = f32(360.0)
    cx = np.ones((ni, 6), f32); cy = np.ones((nj, 6), f32)
    cx[1:ni - 2] = recip_tables([ax[0:ni - 3], ax[1:ni - 2], ax[2:ni - 1], ax[3:ni]])
    cy[1:nj - 2] = recip_tables([ay[0:nj - 3], ay[1:nj - 2], ay[2:nj - 1], ay[3:nj]])
    # seam rows of the table, from the same four nodes the leaf substitutes on the seam (ez_irgdint_3_w.inc:101-150)
    if wrap == 1:
        cx[0] = recip_tables([[ax[ni - 2] - K], [ax[0]], [ax[1]], [ax[2]]])[0]
        cx[ni - 2] = recip_tables([[ax[ni - 3]], [ax[ni - 2]], [ax[ni - 1]], [ax[1] + K]])[0]
    else:
        cx[0] = recip_tables([[ax[ni - 1] - K], [ax[0]], [ax[1]], [ax[2]]])[0]
        cx[ni - 2] = recip_tables([[ax[ni - 3]], [ax[ni - 2]], [ax[ni - 1]], [ax[0] + K]])[0]
        cx[ni - 1] = recip_tables([[ax[ni - 2]], [ax[ni - 1]], [ax[0] + K], [ax[1] + K]])[0]
    px, py = points(rs, ni, j1, j2)
    cxf, cyf = np.asfortranarray(cx), np.asfortranarray(cy)
    for name, z in fields(rs, nj, ni).items():
        zo = np.zeros(px.size, f32)
        shim.ez_irgdint_3_w_(_p(zo), _p(px), _p(py), _i(px.size), _p(ax), _p(ay), cxf.ctypes.data_as(C.c_void_p), cyf.ctypes.data_as(C.c_void_p),
                             _p(z), _i(ni), _i(j1), _i(j2), _i(wrap))
        ref = np_irgdint_3_w(px, py, ax, ay, cx, cy, z, ni, j1, j2, wrap)
        assert np.array_equal(bits(zo), bits(ref)), "%s wrap=%d: %d of %d differ" % (name, wrap, int((bits(zo) != bits(ref)).sum()), zo.size)
