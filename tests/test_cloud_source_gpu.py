"""'Y' point clouds as the SOURCE grid: ez_calcxy_y (k nearest + inverse-square-distance weights) and
ez_applywgts, spatial hashing on the GPU against the reference's brute-force search in the oracle.

Both libraries run with polar_correction = no: with it on, the reference classifies zones from x, y it never computes for a
cloud source (ez_calcxy_y.inc:60-63) and -- where that uninitialised memory happens to be zero -- overwrites the whole result with
the extrapolation value."""
import numpy as np
import pytest

from librmn_b200 import configs

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _field(lat, lon, k=0):
    return (280.0 + 30.0 * np.sin(np.radians(lat) * 2) * np.cos(np.radians(lon) * 3 + 0.1 * k) + 5.0 * np.cos(np.radians(lat) * 7)).astype(np.float32)


def _cloud(lib, lon, lat):
    igl = lib.cxgaig("L", 0.0, 0.0, 1.0, 1.0)
    return lib.ezgdef_fmem(lon.size, 1, "Y", "L", *igl, lon.astype(np.float32), lat.astype(np.float32))


def _opts(lib, extrap="value", nw=16):
    c = configs.Case("cloud", lib)
    c.opts = {"interp_degree": "linear", "polar_correction": "no", "extrap_degree": extrap, "extrap_value": -1.0}
    c.apply_options()
    lib.ezsetival("weight_number", nw)


SCEN = {
    # regional cloud, target larger than its bounding box: masked-out points, and N_dst > N_src (points the reference leaves untouched)
    "regional": dict(n=2000, lon=(200.0, 300.0), lat=(20.0, 70.0), target=lambda L: (configs.def_L(L, 60, 40, 15.0, 190.0, 1.5, 2.0), (40, 60))),
    # global cloud: longitude wrap of the search window on both sides of the seam
    "global": dict(n=6000, lon=(0.0, 360.0), lat=(-89.0, 89.0), target=lambda L: (configs.def_L(L, 120, 61, -90.0, 0.0, 3.0, 3.0), (61, 120))),
    # sparse cloud: fewer than weight_number points inside the +/-10 degree window of many targets
    "sparse": dict(n=150, lon=(100.0, 220.0), lat=(-40.0, 40.0), target=lambda L: (configs.def_L(L, 50, 40, -45.0, 95.0, 2.0, 2.5), (40, 50))),
}


@pytest.mark.parametrize("extrap", ["value", "maximum"])
@pytest.mark.parametrize("nw", [16, 4])
@pytest.mark.parametrize("scen", sorted(SCEN))
def test_cloud_source_against_oracle(ezb, ezref, cuda_ok, scen, nw, extrap):
    sc = SCEN[scen]
    # (one cloud per weight_number: a gridset keeps the number of weights it was BUILT with, ez_calcxy.c:99-113 -- the oracle
    #  starts from scratch in every test, the product's registry lives for the session, so the two must build it under the same option)
    rs = np.random.default_rng(42 + nw)
    lon = rs.uniform(*sc["lon"], sc["n"]).astype(np.float32)
    lat = rs.uniform(*sc["lat"], sc["n"]).astype(np.float32)
    z = _field(lat, lon)
    outs = []
    for lib in (ezb, ezref):
        gy = _cloud(lib, lon, lat)
        gl, shape = sc["target"](lib)
        _opts(lib, extrap, nw)
        assert lib.ezdefset(gl, gy) == 1
        o = np.full(shape, -7.0, np.float32)            # -7: what the caller had there (some points are never written)
        rc = lib.ezsint(o, z.reshape(1, -1))
        outs.append((rc, o))
        lib.ezsetival("weight_number", 16)
    (rc_b, ob), (rc_r, orf) = outs
    assert rc_b == rc_r
    rng = float(z.max() - z.min())
    d = np.abs(ob.astype(np.float64) - orf) / rng
    # a tie between two source points at exactly the same distance for the last slot may pick either one: bounded, not excluded silently
    assert int((d > TOL).sum()) <= max(1, int(2e-4 * d.size)), "%s nw=%d: %d values differ (max %.3g)" % (scen, nw, int((d > TOL).sum()), d.max())
    assert np.array_equal(ob == -7.0, orf == -7.0), "untouched points differ"
    fill = -1.0 if extrap == "value" else None
    if fill is not None:
        assert np.array_equal(ob == fill, orf == fill), "masked-out points differ"
    print("%s nw=%d %s: max|d|/range=%.2e untouched=%d" % (scen, nw, extrap, d.max(), int((orf == -7.0).sum())))


def test_cloud_source_to_cloud_target_winds_and_device_pointers(ezb, ezref, cuda_ok):
    import torch
    rs = np.random.default_rng(7)
    n, m = 5000, 20000
    lon, lat = rs.uniform(0, 360, n).astype(np.float32), rs.uniform(-80, 80, n).astype(np.float32)
    tlon, tlat = rs.uniform(0, 360, m).astype(np.float32), rs.uniform(-75, 75, m).astype(np.float32)
    u, v = configs.wind_fields(lat, lon)
    res = []
    for lib in (ezb, ezref):
        gy, gt = _cloud(lib, lon, lat), _cloud(lib, tlon, tlat)
        _opts(lib, "value")
        assert lib.ezdefset(gt, gy) == 1
        uo, vo = np.full((1, m), -7.0, np.float32), np.full((1, m), -7.0, np.float32)
        assert lib.ezuvint(uo, vo, u.reshape(1, -1), v.reshape(1, -1)) >= 0
        res.append(np.stack([uo, vo]))
    rng = float(max(u.max() - u.min(), v.max() - v.min()))
    d = np.abs(res[0].astype(np.float64) - res[1]) / rng
    assert int((d > TOL).sum()) <= 4, (int((d > TOL).sum()), d.max())
    # several fields, device pointers: same numbers as field-by-field host calls
    z = np.stack([_field(lat, lon, k) for k in range(5)])
    gy, gt = _cloud(ezb, lon, lat), _cloud(ezb, tlon, tlat)
    _opts(ezb, "value")
    assert ezb.ezdefset(gt, gy) == 1
    host = np.full((5, m), -7.0, np.float32)
    for k in range(5):
        assert ezb.ezsint(host[k:k + 1], z[k:k + 1]) == 0
    dz = torch.from_numpy(z).cuda()
    do = torch.full((5, m), -7.0, device="cuda")
    assert ezb.ezsint_batch(do, dz, 5) == 0
    torch.cuda.synchronize()
    assert np.array_equal(do.cpu().numpy(), host)
