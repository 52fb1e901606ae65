"""The direct kernel (ez_interp_fast.cuh: k_interp_direct): calls of 1-4 fields -- c_ezsint of one field, c_ezuvint of one
(u, v) pair, C1 and C3 of BASELINE.json -- skip the tile kernels' per-tile set-up.  Checked here: against the oracle at
full size, against the tile kernels on the same call, that a call's result does not depend on how many of its fields share
the call (within the direct range), that the fused wind rotation equals the separate pass bit for bit, and constant
extrapolation outside the source."""
import numpy as np
import pytest

import cases
from librmn_b200 import configs
from test_parity_gpu import TOL

pytestmark = pytest.mark.gpu


def _cases(lib):
    cs = [configs.c1(lib, 416, "linear"), configs.c5(lib, 0.6, 0.4, "cubic"), configs.c2(lib, 0.6, 0.4, "cubic"), configs.c3(lib, 4, "cubic"),
          configs.c3(lib, 4, "linear"), configs.c3(lib, 4, "cubic", extrap="value"), configs.c5(lib, 0.6, 0.4, "linear")]
    cs[5].vector = False
    cs[5].opts["extrap_value"] = -5.0
    return cs


def test_direct_kernel_against_oracle_and_tile_kernels(ezb, ezref, cuda_ok):
    ezb.set_arith(2)
    for c_b, c_r in zip(_cases(ezb), _cases(ezref)):
        rc_d, out_d = cases.run_case(ezb, c_b)
        path, nz = ezb.last_path()
        assert path & 16, "%s: direct kernel not taken (path=%d)" % (c_b.name, path)
        prev = ezb.set_direct_max_fields(0)
        try:
            rc_t, out_t = cases.run_case(ezb, c_b)
            assert not (ezb.last_path()[0] & 16)
        finally:
            ezb.set_direct_max_fields(prev)
        rc_r, out_r = cases.run_case(ezref, c_r)
        assert rc_d == rc_r == rc_t, c_b.name
        rng = cases.field_range(c_b)
        err_o = float(np.abs(out_d.astype(np.float64) - out_r).max()) / rng
        err_t = float(np.abs(out_d.astype(np.float64) - out_t).max()) / rng
        print("%-40s rc=%d zone-list=%d  |direct-oracle|=%.2e  |direct-tiles|=%.2e of range" % (c_b.name, rc_d, nz, err_o, err_t))
        if not c_b.name.startswith("C2"):   # (C2: a point exactly on the Yin/Yang mask border may flip panel against the oracle)
            assert err_o < TOL, c_b.name
        else:
            assert (np.abs(out_d.astype(np.float64) - out_r) / rng > TOL).sum() <= 8
        assert err_t < 3e-6, c_b.name


@pytest.mark.parametrize("name", ["c1", "c3"])
def test_direct_kernel_full_size(ezb, ezref, cuda_ok, name):
    """BASELINE.json's single-call configurations at full size, one c_ezsint / c_ezuvint call, against the oracle."""
    from test_fullsize_gpu import FULL, oracle_cached
    c_r, rc_r, ref = oracle_cached(ezref, name)
    c_b = FULL[name](ezb)
    ezb.set_arith(2)
    rc_b, out = cases.run_case(ezb, c_b)
    assert ezb.last_path()[0] & 16
    assert rc_b == rc_r
    rng = cases.field_range(c_b)
    d = np.abs(out.astype(np.float64) - ref[0]) / rng
    bad = int((d > TOL).sum())
    print("%s full size: max|direct-oracle|/range=%.2e, %d over tolerance" % (name, float(d.max()), bad))
    assert bad <= 4, (name, bad, float(d.max()))


def test_result_does_not_depend_on_the_field_count_of_the_call(ezb, cuda_ok):
    """1, 2, 3 or 4 fields per call: the same kernel evaluates every field the same way."""
    ezb.set_arith(2)
    for c in (configs.c5(ezb, 3.0, 2.5, "cubic"), configs.c2(ezb, 3.0, 2.5, "cubic"), configs.c1(ezb, 96, "linear")):
        c.defset()
        zin = np.stack([c.source(k) for k in range(4)])
        single = np.zeros((4,) + c.dst_shape, np.float32)
        for k in range(4):
            assert ezb.ezsint(single[k], zin[k]) >= 0
        for nf in (2, 3, 4):
            batch = np.zeros((nf,) + c.dst_shape, np.float32)
            assert ezb.ezsint_batch(batch, zin[:nf], nf) >= 0
            assert ezb.last_path()[0] & 16
            assert np.array_equal(batch, single[:nf]), (c.name, nf)
        # a larger call takes the tile kernels: same plans, same form, another centre of the differences
        zin9 = np.stack([c.source(k % 4) for k in range(9)])
        big = np.zeros((9,) + c.dst_shape, np.float32)
        assert ezb.ezsint_batch(big, zin9, 9) >= 0
        rng = float(zin[0].max() - zin[0].min())
        assert np.abs(big[:4].astype(np.float64) - single).max() < 3e-6 * rng, c.name


@pytest.mark.parametrize("degree", ["linear", "cubic"])
def test_fused_wind_rotation_equals_the_separate_pass(ezb, cuda_ok, degree):
    """One (u, v) pair: the direct kernel rotates the pair before the store.  Two pairs in one call: separate interpolation
    launches per component and the k_wind_apply pass.  Same operations in the same order: bit-equal."""
    ezb.set_arith(2)
    for c in (configs.c3(ezb, 4, degree), configs.c3(ezb, 4, degree, extrap="value")):
        c.defset()
        u0, v0 = c.source_uv(0)
        u1, v1 = c.source_uv(1)
        one = [np.zeros(c.dst_shape, np.float32) for _ in range(4)]
        assert ezb.ezuvint(one[0], one[1], u0, v0) >= 0
        assert ezb.last_path()[0] & 16
        assert ezb.ezuvint(one[2], one[3], u1, v1) >= 0
        uo = np.zeros((2,) + c.dst_shape, np.float32); vo = np.zeros_like(uo)
        assert ezb.ezuvint_batch(uo, vo, np.stack([u0, u1]), np.stack([v0, v1]), 2) >= 0
        assert np.array_equal(uo[0], one[0]) and np.array_equal(vo[0], one[1]), c.name
        assert np.array_equal(uo[1], one[2]) and np.array_equal(vo[1], one[3]), c.name
        # and against the per-call evaluation of the rotation (arith 1: fp64 weight form, cached rotation too; arith 0: reference order)
        ezb.set_arith(0)
        try:
            r = [np.zeros(c.dst_shape, np.float32) for _ in range(2)]
            assert ezb.ezuvint(r[0], r[1], u0, v0) >= 0
        finally:
            ezb.set_arith(2)
        rng = float(max(u0.max() - u0.min(), v0.max() - v0.min()))
        assert max(np.abs(r[0].astype(np.float64) - one[0]).max(), np.abs(r[1].astype(np.float64) - one[1]).max()) < TOL * rng
