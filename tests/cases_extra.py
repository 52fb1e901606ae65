"""Grid types, hemispheric / y-inverted sources, 'U' targets and options that tests/cases.py does not reach
(VERDICT r1, weak #3): every case is defined through the reference's public API in whichever library it is given."""
import numpy as np

from librmn_b200 import configs
from librmn_b200.configs import Case, def_L, def_PS, def_ZE, def_U


def _idx_field_coords(nj, ni, lat0=-88.0, lat1=88.0, lon0=0.0, lon1=355.0):
    lon = np.linspace(lon0, lon1, ni)
    lat = np.linspace(lat0, lat1, nj)
    return np.meshgrid(lon, lat)


def _L_global(lib, ni=90, nj=46):
    return def_L(lib, ni, nj, -90.0, 0.0, 180.0 / (nj - 1), 360.0 / ni), (nj, ni)


def regular_source_cases(lib, degree):
    """'A', 'B', 'G' sources: global, hemispheric (needs_expansion: ez_xpnsrcgd / ez_xpngdag2 / ez_xpngdb2) and y-inverted
    (axe_y_inverse + permut), scalars and winds.  The target is a global 'L' so both polar caps and the mirrored half are read."""
    cs = []
    for grtyp, ni, nj in (("A", 72, 36), ("B", 73, 37), ("G", 96, 48)):
        for ig1, hem in ((0, "glb"), (1, "north"), (2, "south")):
            for ig2 in (0, 1):
                njs = nj if ig1 == 0 else nj // 2
                if grtyp == "B" and ig1 != 0:
                    njs = (nj + 1) // 2
                for vec in (False, True):
                    c = Case("%s_%s%s_to_L_%s%s" % (grtyp, hem, "_yinv" if ig2 else "", degree, "_winds" if vec else ""), lib)
                    c.gdin = lib.ezqkdef(ni, njs, grtyp, ig1, ig2, 0, 0)
                    c.gdout, c.dst_shape = _L_global(lib)
                    c.src_shape = (njs, ni)
                    la0, la1 = (-88.0, 88.0) if ig1 == 0 else ((2.0, 88.0) if ig1 == 1 else (-88.0, -2.0))
                    c.src_lon, c.src_lat = _idx_field_coords(njs, ni, la0, la1)
                    c.opts = {"interp_degree": degree}
                    c.vector = vec
                    if vec and (ig1 != 0 or ig2 != 0):
                        # The reference's own polar wind correction cannot be the yardstick here: it hands ez_calcnpolarwind the
                        # caller's raw array with the EXPANDED row count (ez_corrvec_aunord.c:45-62, ez_calcnpolarwind.c:85), i.e. it
                        # reads past the end of a hemispheric field and the wrong rows of a y-inverted one.  These cases are checked
                        # against the oracle run on the explicitly expanded / un-inverted GLOBAL grid instead (expand_to_global).
                        c.equiv = (grtyp, ni, njs, ig1, ig2)
                    cs.append(c)
    return cs


def expand_to_global(lib, case):
    """The global, south-to-north grid and fields a hemispheric / y-inverted 'A' / 'B' / 'G' source stands for
    (ez_xpngdag2.inc:42-68, ez_xpngdb2.inc, permut.f:55-66).  Returns (gdin, src_shape, expand(array, sign))."""
    grtyp, ni, njs, ig1, ig2 = case.equiv
    if ig1 == 0:
        nj_glb = njs
    else:
        nj_glb = 2 * njs - 1 if grtyp == "B" else 2 * njs
    gd = lib.ezqkdef(ni, nj_glb, grtyp, 0, 0, 0, 0)

    def expand(a, sign):
        d = a[::-1] if ig2 else a
        if ig1 == 0:
            out = d
        elif grtyp == "B":
            out = np.concatenate([sign * d[:0:-1], d]) if ig1 == 1 else np.concatenate([d, sign * d[-2::-1]])
        else:
            out = np.concatenate([sign * d[::-1], d]) if ig1 == 1 else np.concatenate([d, sign * d[::-1]])
        return np.ascontiguousarray(out, np.float32)

    return gd, (nj_glb, ni), expand


E_ROT = (30.0, 200.0, 50.0, 250.0)


def rotated_cases(lib, degree):
    """'E' grids as source and as target, '#' source, 'B' target fed from a rotated 'Z' source (ez_corrbgd)."""
    cs = []
    ige = lib.cxgaig("E", *E_ROT)
    # 'E' sources only below cubic: with cubic + polar correction the reference itself dereferences the (absent) ay axis of a
    # regular 'E' grid (ez_corrval_aunord.c:57-63, `lgd->ay[lgd->j2-1]` with ay == NULL) and crashes
    e_src = degree != "cubic"
    # 'E' global source -> 'L'
    c = Case("E72x37_to_L_%s" % degree, lib)
    c.gdin = lib.ezqkdef(72, 37, "E", *ige)
    c.gdout, c.dst_shape = _L_global(lib, 80, 41)
    c.src_shape = (37, 72)
    c.src_lon, c.src_lat = _idx_field_coords(37, 72, -90.0, 90.0, 0.0, 355.0)
    c.opts = {"interp_degree": degree}
    if e_src:
        cs.append(c)
    # 'L' -> 'E' target
    c = Case("L144x73_to_E60x31_%s" % degree, lib)
    c.gdin = def_L(lib, 144, 73, -90.0, 0.0, 2.5, 2.5)
    c.gdout = lib.ezqkdef(60, 31, "E", *ige)
    c.src_shape, c.dst_shape = (73, 144), (31, 60)
    c.src_lon, c.src_lat = configs._ll_grid(144, 73, -90.0, 0.0, 2.5, 2.5)
    c.opts = {"interp_degree": degree}
    cs.append(c)
    for vec in (False,):
        # (winds INTO an 'E' target are not comparable: c_gduvfwd_orig's 'E' branch passes freshly malloc'ed, uninitialised
        # arrays as the INPUT of ez_gfxyfll and overwrites the caller's lat/lon with the result, gduvfwd.c:45-49; see
        # test_e_target_winds_equal_ze_target_winds for what is checked instead)
        c = Case("E72x37_to_E60x31_%s%s" % (degree, "_winds" if vec else ""), lib)
        c.gdin = lib.ezqkdef(72, 37, "E", *lib.cxgaig("E", 10.0, 100.0, 40.0, 150.0))
        c.gdout = lib.ezqkdef(60, 31, "E", *ige)
        c.src_shape, c.dst_shape = (37, 72), (31, 60)
        c.src_lon, c.src_lat = _idx_field_coords(37, 72, -90.0, 90.0, 0.0, 355.0)
        c.opts = {"interp_degree": degree}
        c.vector = vec
        if e_src:
            cs.append(c)
    # '#' source (from memory it is a 'Z' with its own axes) on a rotated frame, LAM
    ax = (150.0 + 0.5 * np.arange(80)).astype(np.float32)
    ay = (-20.0 + 0.5 * np.arange(60)).astype(np.float32)
    for vec in (False, True):
        c = Case("hash80x60_to_L_%s%s" % (degree, "_winds" if vec else ""), lib)
        c.gdin = lib.ezgdef_fmem(80, 60, "#", "E", *lib.cxgaig("E", *configs.LAM_ROT), ax, ay)
        c.gdout = def_L(lib, 70, 50, 20.0, 225.0, 0.6, 0.7)
        c.src_shape, c.dst_shape = (60, 80), (50, 70)
        c.src_lon, c.src_lat = np.meshgrid(ax.astype(np.float64), ay.astype(np.float64))
        c.opts = {"interp_degree": degree, "extrap_degree": "value", "extrap_value": -3.0}
        c.vector = vec
        cs.append(c)
    # global rotated 'Z' -> 'B' target: pole rows averaged (ez_corrval.c:144-147, ez_corrbgd)
    axg = (2.5 * np.arange(145)).astype(np.float32)
    ayg = (-88.75 + 2.5 * np.arange(72)).astype(np.float32)
    c = Case("Zrot145x72_to_B73x37_%s" % degree, lib)
    c.gdin = def_ZE(lib, axg, ayg, *E_ROT)
    c.gdout = lib.ezqkdef(73, 37, "B", 0, 0, 0, 0)
    c.src_shape, c.dst_shape = (72, 145), (37, 73)
    c.src_lon, c.src_lat = np.meshgrid(axg.astype(np.float64), ayg.astype(np.float64))
    c.opts = {"interp_degree": degree}
    cs.append(c)
    # 'L' -> 'A' and 'L' -> 'G' targets (grll / Gaussian latitudes in ez_calclatlon)
    for g, ni, nj in (("A", 72, 36), ("G", 64, 32), ("B", 73, 37)):
        c = Case("L144x73_to_%s_%s" % (g, degree), lib)
        c.gdin = def_L(lib, 144, 73, -90.0, 0.0, 2.5, 2.5)
        c.gdout = lib.ezqkdef(ni, nj, g, 0, 0, 0, 0)
        c.src_shape, c.dst_shape = (73, 144), (nj, ni)
        c.src_lon, c.src_lat = configs._ll_grid(144, 73, -90.0, 0.0, 2.5, 2.5)
        c.opts = {"interp_degree": degree}
        cs.append(c)
    return cs


def yinyang_cases(lib, degree):
    """'U' as a TARGET (L -> U, U -> U), winds through c_ezyyuvint in the three directions, use_1subgrid."""
    cs = []
    for vec in (False, True):
        sfx = "_winds" if vec else ""
        # L -> U
        c = Case("L144x73_to_U4_%s%s" % (degree, sfx), lib)
        c.gdin = def_L(lib, 144, 73, -90.0, 0.0, 2.5, 2.5)
        u, yin, yan, ax, ay = def_U(lib, 4.0)
        c.gdout = u
        c.src_shape, c.dst_shape = (73, 144), (2 * len(ay), len(ax))
        c.src_lon, c.src_lat = configs._ll_grid(144, 73, -90.0, 0.0, 2.5, 2.5)
        c.opts = {"interp_degree": degree}
        c.vector = vec
        cs.append(c)
        # U -> U (different resolutions)
        c = Case("U3_to_U4_%s%s" % (degree, sfx), lib)
        us, yin_s, yan_s, axs, ays = def_U(lib, 3.0)
        ud, yin_d, yan_d, axd, ayd = def_U(lib, 4.0)
        c.gdin, c.gdout = us, ud
        c.subgrids = (yin_s, yan_s)
        c.src_shape, c.dst_shape = (2 * len(ays), len(axs)), (2 * len(ayd), len(axd))
        lon, lat = np.meshgrid(axs.astype(np.float64), ays.astype(np.float64))
        c.src_lon = np.concatenate([lon, lon + 30.0], axis=0)
        c.src_lat = np.concatenate([lat, lat], axis=0)
        c.opts = {"interp_degree": degree}
        c.vector = vec
        cs.append(c)
    # U -> L winds (c_ezyyuvint, per-panel rotation then merge)
    c = configs.c2(lib, 3.0, 2.5, degree)
    c.vector = True
    c.name += "_winds"
    cs.append(c)
    # use_1subgrid: only the Yin (then only the Yang) panel feeds the target
    for which in (0, 1):
        c = configs.c2(lib, 3.0, 2.5, degree)
        c.name += "_1subgrid_%s" % ("yin", "yang")[which]
        c.opts["use_1subgrid"] = "yes"
        c.subgridid = c.subgrids[which]
        cs.append(c)
    return cs


def defset(case):
    """case.defset() plus the integer option the Case class does not know about."""
    lib = case.lib
    case.apply_options()
    if hasattr(case, "subgridid"):
        lib.ezsetival("subgridid", case.subgridid)
    return lib.ezdefset(case.gdout, case.gdin)


def run(lib, case, k=0, wd=False, expanded=False):
    """One call on host arrays.  expanded=True: run the case's global equivalent (expand_to_global) instead of the grid as defined."""
    if expanded:
        gd, shape, expand = expand_to_global(lib, case)
        case.apply_options()
        lib.ezdefset(case.gdout, gd)
        u, v = case.source_uv(k)
        uo = np.zeros(case.dst_shape, np.float32)
        vo = np.zeros(case.dst_shape, np.float32)
        rc = lib.ezuvint(uo, vo, expand(u, 1.0), expand(v, -1.0))
        return rc, np.stack([uo, vo])
    defset(case)
    try:
        if case.vector:
            u, v = case.source_uv(k)
            uo = np.zeros(case.dst_shape, np.float32)
            vo = np.zeros(case.dst_shape, np.float32)
            rc = (lib.ezwdint if wd else lib.ezuvint)(uo, vo, u, v)
            return rc, np.stack([uo, vo])
        z = case.source(k)
        out = np.zeros(case.dst_shape, np.float32)
        rc = lib.ezsint(out, z)
        return rc, out[None]
    finally:
        lib.ezsetopt("use_1subgrid", "no")
