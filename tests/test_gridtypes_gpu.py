"""GPU parity on the grid types, source layouts and Yin-Yang directions tests/cases.py does not reach:
'A' / 'B' / 'G' sources (global, hemispheric = needs_expansion, y-inverted = permut), 'E' grids as source and target, '#' sources,
'A' / 'B' / 'G' / 'E' / 'U' TARGETS, the B-grid pole-row averaging (ez_corrbgd), winds through every one of them (c_ezuvint,
c_ezyyuvint, c_ezwdint) and use_1subgrid.  Every case runs in the oracle (the reference's own C + restated Fortran leaves) and in
libezb200.so on the same inputs; values must agree to 1e-5 of the field's range away from coordinate ties."""
import numpy as np
import pytest

import cases
import cases_extra as ce
from test_parity_gpu import TOL, coord_eps, near_half, near_int, oracle_xy

pytestmark = pytest.mark.gpu

FAMILIES = {"regular_sources": ce.regular_source_cases, "rotated": ce.rotated_cases, "yinyang": ce.yinyang_cases}


def tie_mask(ezref, c_r, c_b, degree, shape):
    """Points whose oracle coordinates sit within rounding noise of a decision boundary (nearest pick, zone class)."""
    yy_in, yy_out = c_b.lib.ezget_nsubgrids(c_b.gdin) > 1, c_b.lib.ezget_nsubgrids(c_b.gdout) > 1
    if yy_in and not yy_out and degree == "nearest":
        # 'U' source: per-panel coordinates come from the product's own gridset (the reference's c_gdxyfll on a 'U' grid
        # answers a different question, gdxyfll.c:199-236); exact .5 ties of the nearest pick may flip
        gdin = getattr(c_b, "subgridid", c_b.gdin)
        xb, yb, _, _ = c_b.lib.get_gridset(c_b.gdout, gdin, c_b.n_out)
        ex, ey = coord_eps(c_b, xb, yb)
        return np.broadcast_to((near_half(xb, ex) | near_half(yb, ey)).reshape(c_b.dst_shape), shape)
    if yy_in or yy_out:
        return None
    if hasattr(c_b, "equiv"):
        # ties located with the product's own coordinates (valid for both oracle runs of these cases)
        xb, yb, _, _ = c_b.lib.get_gridset(c_b.gdout, c_b.gdin, c_b.n_out)
        ex, ey = coord_eps(c_b, xb, yb)
        tie = near_int(yb, ey)
        if degree == "nearest":
            tie |= near_half(xb, ex) | near_half(yb, ey)
        return np.broadcast_to(tie.reshape(c_b.dst_shape), shape)
    xr, yr = oracle_xy(ezref, c_r)
    ex, ey = coord_eps(c_b, xr, yr)
    nj, ni = c_b.src_shape
    tie = np.zeros(xr.shape, bool)
    if degree == "nearest" or c_b.opts.get("extrap_degree") == "nearest":
        tie |= near_half(xr, ex) | near_half(yr, ey)
    tie |= (near_half(xr, ex) | near_half(yr, ey)) & ((xr < 1.5) | (yr < 1.5) | (xr > ni - 0.5) | (yr > nj - 0.5))
    # polar-cap rows of the LOGICAL (expanded) source: any integer y can be a cap boundary of a hemispheric grid
    tie |= near_int(yr, ey)
    return np.broadcast_to(tie.reshape(c_b.dst_shape), shape)


def compare_case(ezb, ezref, c_b, c_r, degree, wd=False, expanded=False):
    """expanded=True (winds from hemispheric / y-inverted sources): the oracle runs the explicitly expanded global grid."""
    rc_b, out_b = ce.run(ezb, c_b, wd=wd)
    rc_r, out_r = ce.run(ezref, c_r, wd=wd, expanded=expanded)
    rng = cases.field_range(c_b)
    if wd:
        # direction is an angle: compare on the circle, in degrees, scaled like a field of range 360
        d = np.abs(out_b.astype(np.float64) - out_r)
        d[1] = np.minimum(d[1], 360.0 - d[1]) * (rng / 360.0)
        # the direction of a vanishing wind is ill-conditioned
        d[1][out_r[0] < 1e-3 * rng] = 0.0
        d = d / rng
    else:
        d = np.abs(out_b.astype(np.float64) - out_r) / rng
    tie = tie_mask(ezref, c_r, c_b, degree, out_b.shape)
    if hasattr(c_b, "equiv") and not expanded:
        # oracle on the hemispheric / y-inverted grid itself: its polar-cap vector rewrite reads out of bounds (see cases_extra),
        # everything else is c_ezsint + rotation and must agree.  Polar-cap points are left to the expanded comparison.
        _, _, zb, _ = ezb.get_gridset(c_b.gdout, c_b.gdin, c_b.n_out)
        d = np.where(np.broadcast_to(((zb & 0x1e) != 0).reshape(c_b.dst_shape), d.shape), 0.0, d)
    bad = d > TOL
    nbad_tie = 0
    if tie is not None:
        nbad_tie = int((bad & tie).sum())
        bad = bad & ~tie
    return rc_b, rc_r, int(bad.sum()), nbad_tie, float(d.max()), out_b.size


@pytest.mark.parametrize("arith", [0, 1, 2])
@pytest.mark.parametrize("degree", cases.DEGREES)
@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_grid_types_against_oracle(ezb, ezref, cuda_ok, family, degree, arith):
    ezb.set_arith(arith)
    problems = []
    try:
        todo = []
        for c_b, c_r in zip(FAMILIES[family](ezb, degree), FAMILIES[family](ezref, degree)):
            todo.append((c_b, c_r, False))
            if hasattr(c_b, "equiv"):
                # Polar caps of hemispheric / y-inverted wind sources: the oracle on the expanded global grid is the yardstick
                # where that grid is truly equivalent.  It is not for NORTH-hemispheric sources (their mirrored half has negative
                # logical indices, where the reference's ifix() truncation picks the cell above and extrapolates,
                # ez_rgdint_1_nw.inc:30-34) nor for hemispheric Gaussian grids (pole-row means over different latitude sets).
                grtyp, _, _, ig1, ig2 = c_b.equiv
                if ig1 == 0 or (ig1 == 2 and grtyp != "G"):
                    todo.append((c_b, c_r, True))
        for c_b, c_r, expanded in todo:
            rc_b, rc_r, nbad, nbad_tie, dmax, n = compare_case(ezb, ezref, c_b, c_r, degree, expanded=expanded)
            yy = ezb.ezget_nsubgrids(c_b.gdin) > 1 or ezb.ezget_nsubgrids(c_b.gdout) > 1
            # Yin-Yang routes have no oracle-side coordinates to locate ties with: panel choices and nearest picks on exact ties
            # may flip at a handful of points (the reduced grids are commensurate on purpose)
            allow = max(4, int(0.004 * n)) if yy else 0
            print("%-44s a%d%s rc=%d/%d off-tol=%d (allowed %d) on-ties=%d max|d|/range=%.2e" % (
                c_b.name, arith, " [expanded]" if expanded else "", rc_b, rc_r, nbad, allow, nbad_tie, dmax))
            if rc_b != rc_r:
                problems.append("%s: rc %d vs oracle %d" % (c_b.name, rc_b, rc_r))
            if nbad > allow:
                problems.append("%s: %d of %d values off by more than %g of range (max %.3g)" % (c_b.name, nbad, n, TOL, dmax))
            if nbad_tie > max(4, 0.03 * n):
                problems.append("%s: %d near-tie points flipped" % (c_b.name, nbad_tie))
    finally:
        ezb.set_arith(2)
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("degree", ["linear", "cubic"])
def test_ezwdint_against_oracle(ezb, ezref, cuda_ok, degree):
    """c_ezwdint (ezwdint.c:36): interpolated winds returned as (speed, meteorological direction)."""
    problems = []
    def mk(lib):
        cs = [c for c in cases.small_cases(lib, degree) if c.vector]
        cs += [c for c in ce.rotated_cases(lib, degree) + ce.yinyang_cases(lib, degree) if c.vector]
        return cs
    for c_b, c_r in zip(mk(ezb), mk(ezref)):
        rc_b, rc_r, nbad, nbad_tie, dmax, n = compare_case(ezb, ezref, c_b, c_r, degree, wd=True)
        yy = ezb.ezget_nsubgrids(c_b.gdin) > 1 or ezb.ezget_nsubgrids(c_b.gdout) > 1
        allow = max(4, int(0.004 * n)) if yy else 0
        print("%-44s wd rc=%d/%d off-tol=%d on-ties=%d max=%.2e" % (c_b.name, rc_b, rc_r, nbad, nbad_tie, dmax))
        if rc_b != rc_r or nbad > allow:
            problems.append("%s: rc %d/%d, %d of %d off (max %.3g)" % (c_b.name, rc_b, rc_r, nbad, n, dmax))
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("degree", ["linear", "cubic"])
def test_e_target_winds_equal_ze_target_winds(ezb, ezref, cuda_ok, degree):
    """Winds into an 'E' TARGET.  The reference's c_gduvfwd_orig is broken for grtyp 'E' (gduvfwd.c:45-49 feeds uninitialised
    memory to ez_gfxyfll), while its 'Z' on 'E' branch (:60-75) is sound and means the same thing: the product's 'E' target
    must give the winds of the equivalent 'Z'/'E' target (same rotation, axes = the 'E' grid's own lon/lat), and that 'Z'/'E'
    result must match the oracle."""
    rot = ce.E_ROT
    ni, nj = 60, 31
    outs = {}
    for lib, tag in ((ezb, "b"), (ezref, "r")):
        src = ce.def_L(lib, 144, 73, -90.0, 0.0, 2.5, 2.5)
        ige = lib.cxgaig("E", *rot)
        # axes of an 'E' grid (ez_calclatlon.c:64-67 / ez_ll2rgd: dlon = 360 / (ni - 1), dlat = 180 / nj, first row at -90 + dlat / 2)
        ax = (360.0 / (ni - 1)) * np.arange(ni)
        ay = -90.0 + (180.0 / nj) * (np.arange(nj) + 0.5)
        gz = ce.def_ZE(lib, ax, ay, *rot)
        ge = lib.ezqkdef(ni, nj, "E", *ige)
        lon, lat = ce.configs._ll_grid(144, 73, -90.0, 0.0, 2.5, 2.5)
        u, v = ce.configs.wind_fields(lat, lon)
        for g, nm in ((gz, "z"), (ge, "e")):
            if tag == "r" and nm == "e":
                continue
            c = ce.Case("e_equiv", lib)
            c.opts = {"interp_degree": degree}
            c.apply_options()
            assert lib.ezdefset(g, src) == 1
            uo, vo = np.zeros((nj, ni), np.float32), np.zeros((nj, ni), np.float32)
            assert lib.ezuvint(uo, vo, u.reshape(73, 144), v.reshape(73, 144)) >= 0
            outs[tag + nm] = np.stack([uo, vo])
    rng = float(max(u.max() - u.min(), v.max() - v.min()))
    assert np.abs(outs["bz"].astype(np.float64) - outs["rz"]).max() / rng < TOL
    assert np.abs(outs["be"].astype(np.float64) - outs["bz"]).max() / rng < TOL
