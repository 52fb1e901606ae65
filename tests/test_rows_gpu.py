"""Output-row bands: a call with fewer fields than GPUs is split by target rows.
Bands of whole tile rows stitched together must equal the full call BIT FOR BIT (same kernels, same tiles); arbitrary
bands run on the per-point kernel and must agree with the full call to rounding of the local variation."""
import numpy as np
import pytest

import cases
from librmn_b200 import configs, sharding

pytestmark = pytest.mark.gpu


def _cases(lib):
    cs = [configs.c1(lib, 416, "linear"), configs.c5(lib, 0.6, 0.4, "cubic"), configs.c2(lib, 0.6, 0.4, "cubic"), configs.c3(lib, 4, "cubic"),
          configs.c3(lib, 4, "linear", extrap="value")]
    cs[4].vector = False
    cs[4].opts["extrap_value"] = -5.0
    return cs


def _call(lib, c, j0=None, j1=None):
    nj, ni = c.dst_shape
    rows = nj if j0 is None else j1 - j0
    if j0 is not None:
        lib.set_rows(j0, j1)
    try:
        if c.vector:
            u, v = c.source_uv()
            uo = np.full((rows, ni), -77.0, np.float32); vo = np.full((rows, ni), -77.0, np.float32)
            rc = lib.ezuvint(uo, vo, u, v)
            return rc, np.stack([uo, vo])
        o = np.full((rows, ni), -77.0, np.float32)
        rc = lib.ezsint(o, c.source())
        return rc, o[None]
    finally:
        lib.set_rows()


@pytest.mark.parametrize("direct_max", [0, 4])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_tile_row_bands_stitch_to_the_full_call_bit_for_bit(ezb, cuda_ok, world, direct_max):
    """direct_max = 0: the tile kernels on whole tile rows ; 4: the direct kernel (one field / one pair per call), which takes any band."""
    ezb.set_arith(2)
    prev = ezb.set_direct_max_fields(direct_max)
    try:
        _stitch(ezb, world, 16 if direct_max else 0)
    finally:
        ezb.set_direct_max_fields(prev)


def _stitch(ezb, world, want_path):
    for c in _cases(ezb):
        c.defset()
        rc_full, full = _call(ezb, c)
        nj, ni = c.dst_shape
        parts = []
        for r in range(world):
            j0, j1 = sharding.row_band(nj, r, world)
            if j1 > j0:
                rc, band = _call(ezb, c, j0, j1)
                assert rc == rc_full or rc >= 0, c.name
                path, _ = ezb.last_path()
                assert not want_path or (path & want_path), (c.name, path)
                parts.append(band)
        stitched = np.concatenate(parts, axis=1)
        assert stitched.shape == full.shape, c.name
        assert np.array_equal(stitched, full), "%s world=%d: %d values differ" % (c.name, world, int((stitched != full).sum()))


def test_arbitrary_bands_match_the_full_call(ezb, cuda_ok):
    ezb.set_arith(2)
    for c in _cases(ezb):
        c.defset()
        _, full = _call(ezb, c)
        nj, ni = c.dst_shape
        rng = cases.field_range(c)
        for j0, j1 in ((5, 77), (0, 1), (nj - 3, nj), (33, 64)):
            rc, band = _call(ezb, c, j0, j1)
            assert rc >= 0
            d = np.abs(band.astype(np.float64) - full[:, j0:j1]).max() / rng
            assert d < 3e-6, (c.name, j0, j1, d)
    # device pointers and several fields
    import torch
    c = configs.c5(ezb, 0.6, 0.4, "cubic")
    c.defset()
    nj, ni = c.dst_shape
    z = torch.from_numpy(np.stack([c.source(k) for k in range(5)])).cuda()
    full = torch.empty((5, nj, ni), device="cuda")
    assert ezb.ezsint_batch(full, z, 5) == 0
    j0, j1 = sharding.row_band(nj, 1, 4)
    band = torch.empty((5, j1 - j0, ni), device="cuda")
    ezb.set_rows(j0, j1)
    try:
        assert ezb.ezsint_batch(band, z, 5) == 0
    finally:
        ezb.set_rows()
    torch.cuda.synchronize()
    assert torch.equal(band, full[:, j0:j1])
