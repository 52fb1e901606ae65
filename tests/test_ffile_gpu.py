"""Interpolation on grids defined from positional records: the 'U' grid of a '^>' record and a rotated 'Z'
grid of '>>' / '^^' records as SOURCES, against the oracle running the reference's own ezgdef_ffile.c on the same records."""
import numpy as np
import pytest

from librmn_b200 import configs
from test_ffile_cpu import _table
from test_parity_gpu import TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("degree", ["linear", "cubic"])
def test_interpolation_from_file_defined_grids(ezb, ezref, cuda_ok, degree):
    fb, fr = _table(ezb), _table(ezref)
    assert fb.install(ezb) == 0
    fr.install(ezref)
    try:
        outs = {}
        fields = {}
        for lib in (ezref, ezb):   # (the oracle first: both sides interpolate the field built from ITS lat / lon)
            lib.ezsetopt("interp_degree", degree)
            lib.ezsetopt("extrap_degree", "maximum")
            gu = lib.ezqkdef(0, 0, "U", 5001, 5002, 0, 0, 1)
            gz = lib.ezqkdef(120, 90, "Z", 1001, 1002, 0, 0, 1)
            assert gu >= 0 and gz >= 0
            res = []
            for gdin, shape in ((gu, (42, 61)), (gz, (90, 120))):
                lat = np.zeros(shape, np.float32); lon = np.zeros(shape, np.float32)
                assert lib.gdll(gdin, lat, lon) >= 0
                if lib is ezref:
                    fields[shape] = configs.field(lat.astype(np.float64), lon.astype(np.float64), k=1).astype(np.float32)
                else:   # c_gdll of the two libraries agrees to rounding (libm), not to the bit
                    la_r = fields[shape + ("lat",)]
                    assert np.abs(lat - la_r).max() < 1e-4
                fields[shape + ("lat",)] = lat if lib is ezref else fields[shape + ("lat",)]
                zin = fields[shape]
                gdout = configs.def_L(lib, 90, 45, -88.0, 0.0, 4.0, 4.0)
                assert lib.ezdefset(gdout, gdin) == 1
                z = np.zeros((45, 90), np.float32)
                rc = lib.ezsint(z, zin)
                res.append((rc, z, zin))
            outs[lib is ezb] = res
        for (rc_b, zb_, zin), (rc_r, zr, zin_r) in zip(outs[True], outs[False]):
            assert rc_b == rc_r
            rng = float(zin.max() - zin.min())
            d = np.abs(zb_.astype(np.float64) - zr) / rng
            assert int((d > TOL).sum()) <= 4, float(d.max())   # (a point exactly on the Yin / Yang mask border may flip panel)
    finally:
        ezb.lib.c_ezb200_set_fst_hooks(None, None, None, None)
        ezref.lib.oracle_set_fst_hooks(None, None, None, None)
        for lib in (ezb, ezref):
            lib.ezsetopt("interp_degree", "cubic")
