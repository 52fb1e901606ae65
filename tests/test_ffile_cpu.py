"""File-backed grid definitions: c_ezqkdef / c_ezgdef_ffile on '#', 'Z', 'Y' and 'U' grids whose positional
records ('>>', '^^', '^>') come from an FST layer -- here tests/fake_fst.py, served to the product and to the oracle alike, so
that the product's restatement of ezgdef_ffile.c / ez_lireEnrPositionnels.c is compared with the reference's own C code.
Host logic only: no GPU needed (the compute on such grids is covered by test_ffile_gpu.py)."""
import numpy as np
import pytest

from fake_fst import FakeFst
from librmn_b200 import configs


def _table(lib):
    f = FakeFst()
    ige = lib.cxgaig("E", 31.0, 264.0, 45.0, 12.0)          # a rotated reference
    ax = (150.0 + 0.25 * np.arange(120)).astype(np.float32)
    ay = (-12.0 + 0.25 * np.arange(90)).astype(np.float32)
    f.add(">>", ax, 1001, 1002, 0, "E", ige)
    f.add_column("^^", ay, 1001, 1002, 0, "E", ige)
    igl = lib.cxgaig("L", 0.0, 0.0, 1.0, 1.0)
    axl = (-100.0 + 0.5 * np.arange(200)).astype(np.float32)  # negative longitudes (-100 .. -0.5): + 360 on an L reference
    ayl = (10.0 + 0.5 * np.arange(100)).astype(np.float32)
    f.add(">>", axl, 2001, 2002, 7, "L", igl)
    f.add_column("^^", ayl, 2001, 2002, 7, "L", igl)
    # a second '>>' under the same keys with another length: the search must walk on to the one that matches ni (c_fstsui)
    f.add(">>", axl[:50], 3001, 3002, 0, "L", igl)
    f.add(">>", axl, 3001, 3002, 0, "L", igl)
    f.add_column("^^", ayl, 3001, 3002, 0, "L", igl)
    # 'Y' cloud: both records ni x nj
    lon2, lat2 = np.meshgrid(200.0 + 0.3 * np.arange(40), 30.0 + 0.2 * np.arange(25))
    f.add(">>", lon2.astype(np.float32), 4001, 4002, 0, "L", igl)
    f.add("^^", lat2.astype(np.float32), 4001, 4002, 0, "L", igl)
    # '^>' of a Yin-Yang grid, vercode 1 (ez_lireEnrPositionnels.c:12-46)
    ni, nj = 61, 21
    yax = (45.0 + 4.5 * np.arange(ni)).astype(np.float32)
    yay = (-45.0 + 4.5 * np.arange(nj)).astype(np.float32)
    def panel(rot):
        return np.concatenate([np.array([ni, nj, yax[0], yax[-1], yay[0], yay[-1], *rot], np.float32), yax, yay])
    yy = np.concatenate([np.array([0.0, 0.0, 2.0, 1.0, 0.0], np.float32), panel((0.0, 180.0, 0.0, 270.0)), panel((0.0, 0.0, 90.0, 0.0))])
    f.add("^>", yy, 5001, 5002, 0, "F", (1, 0, 0, 0))
    return f


def _describe(lib, gd):
    rc, ni, nj, grtyp, ig, grref, igref = lib.ezgxprm(gd)
    assert rc >= 0
    out = {"prm": {"ni": ni, "nj": nj, "grtyp": grtyp, "ig": ig, "grref": grref, "igref": igref}}
    if grtyp in ("Z", "#", "Y"):
        n = ni * nj if grtyp == "Y" else None
        ax = np.zeros(n or ni, np.float32); ay = np.zeros(n or nj, np.float32)
        assert lib.gdgaxes(gd, ax, ay) == 0
        out["ax"], out["ay"] = ax, ay
    return out


def _same(a, b, what):
    for k in a["prm"]:
        assert a["prm"][k] == b["prm"][k], (what, k, a["prm"][k], b["prm"][k])
    for k in ("ax", "ay"):
        if k in a:
            assert np.array_equal(a[k], b[k]), (what, k)


@pytest.fixture
def both(ezb, ezref):
    fb, fr = _table(ezb), _table(ezref)
    assert fb.install(ezb) == 0
    fr.install(ezref)
    yield ezb, ezref, fb, fr
    ezb.lib.c_ezb200_set_fst_hooks(None, None, None, None)
    ezref.lib.oracle_set_fst_hooks(None, None, None, None)


def test_z_hash_and_y_grids_from_positional_records(both):
    ezb, ezref, fb, fr = both
    cases = [("Z rotated", (120, 90, "Z", 1001, 1002, 0, 0, 1)),
             ("Z on L, negative longitudes", (200, 100, "Z", 2001, 2002, 7, 0, 1)),
             ("Z, second '>>' matches", (200, 100, "Z", 3001, 3002, 0, 0, 1)),
             ("Z, dimensions from the records", (-1, -1, "Z", 1001, 1002, 0, 0, 1)),
             ("# window", (40, 30, "#", 2001, 2002, 21, 11, 1)),
             ("Y cloud", (40, 25, "Y", 4001, 4002, 0, 0, 1))]
    for what, args in cases:
        gb, gr = ezb.ezqkdef(*args), ezref.ezqkdef(*args)
        assert gb >= 0 and gr >= 0, (what, gb, gr)
        _same(_describe(ezb, gb), _describe(ezref, gr), what)
    # the same definition again returns the same grid (ezgdef_ffile.c:84-96) ; missing records fail
    assert ezb.ezqkdef(120, 90, "Z", 1001, 1002, 0, 0, 1) == ezb.ezqkdef(120, 90, "Z", 1001, 1002, 0, 0, 1)
    assert ezb.ezqkdef(120, 90, "Z", 9001, 9002, 0, 0, 1) < 0
    assert ezb.ezqkdef(77, 90, "Z", 1001, 1002, 0, 0, 1) < 0           # no '>>' of that length


def test_u_grid_from_the_yy_record_equals_the_supergrid_built_from_memory(both):
    ezb, ezref, fb, fr = both
    res = []
    for lib in (ezb, ezref):
        gd = lib.ezqkdef(0, 0, "U", 5001, 5002, 0, 0, 1)
        assert gd >= 0
        assert lib.ezget_nsubgrids(gd) == 2
        subs = lib.ezget_subgridids(gd)
        d = [_describe(lib, s) for s in subs]
        rc, grtyp, v = lib.ezgprm(gd)
        res.append(((grtyp, v[0], v[1]), d))
    (pb, db), (pr, dr) = res
    assert pb == pr == ("U", 61, 42)
    for a, b in zip(db, dr):
        _same(a, b, "U sub-grid")
    # ... and is the grid c_ezgdef_fmem x 2 + c_ezgdef_supergrid define from the same numbers
    yax = (45.0 + 4.5 * np.arange(61)).astype(np.float32); yay = (-45.0 + 4.5 * np.arange(21)).astype(np.float32)
    yin = ezb.ezgdef_fmem(61, 21, "Z", "E", *ezb.cxgaig("E", 0.0, 180.0, 0.0, 270.0), yax, yay)
    yan = ezb.ezgdef_fmem(61, 21, "Z", "E", *ezb.cxgaig("E", 0.0, 0.0, 90.0, 0.0), yax, yay)
    assert [yin, yan] == list(ezb.ezget_subgridids(ezb.ezqkdef(0, 0, "U", 5001, 5002, 0, 0, 1)))


def test_without_an_fst_layer_file_backed_types_fail_loudly(ezb):
    ezb.lib.c_ezb200_set_fst_hooks(None, None, None, None)
    assert ezb.ezqkdef(10, 10, "Z", 1, 2, 3, 0, 1) == -1
    assert ezb.ezqkdef(10, 10, "L", *ezb.cxgaig("L", 0.0, 0.0, 1.0, 1.0), 0) >= 0     # regular types never touch the file
