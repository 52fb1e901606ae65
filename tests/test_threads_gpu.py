"""Threading contract: grids are defined serially, then several host threads call
c_ezsetopt / c_ezdefset / c_ezsint / c_ezuvint concurrently, each with its own thread-local session (the reference's
__thread iset_gdin / iset_gdout / groptions, ez_statics.c:12-17).  Results must be bit-equal to the serial ones, whether the
threads share one prebuilt gridset or build different gridsets at the same time."""
import threading

import numpy as np
import pytest

from librmn_b200 import configs

pytestmark = pytest.mark.gpu


def _cases(lib):
    cs = [configs.c5(lib, 1.5, 1.0, "cubic"), configs.c2(lib, 1.5, 1.0, "cubic"), configs.c1(lib, 300, "linear"), configs.c3(lib, 8, "cubic")]
    cs[3].vector = True
    return cs


def _run(lib, c, k, nf):
    """One thread's work: its own options, its own set, nf fields in one batched call on host arrays."""
    c.apply_options()
    assert lib.ezdefset(c.gdout, c.gdin) == 1
    if c.vector:
        uv = [c.source_uv(k + f) for f in range(nf)]
        u = np.stack([a for a, _ in uv]); v = np.stack([b for _, b in uv])
        uo = np.zeros((nf,) + c.dst_shape, np.float32); vo = np.zeros_like(uo)
        rc = lib.ezuvint_batch(uo, vo, u, v, nf)
        return rc, np.stack([uo, vo])
    z = np.stack([c.source(k + f) for f in range(nf)])
    o = np.zeros((nf,) + c.dst_shape, np.float32)
    rc = lib.ezsint_batch(o, z, nf)
    return rc, o


def _threaded(lib, jobs):
    res = [None] * len(jobs)
    err = []

    def work(i):
        try:
            c, k, nf = jobs[i]
            res[i] = _run(lib, c, k, nf)
        except Exception as ex:   # pragma: no cover
            err.append(repr(ex))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not err, err
    return res


def test_threads_share_a_prebuilt_gridset(ezb, cuda_ok):
    c = configs.c5(ezb, 1.5, 1.0, "cubic")
    jobs = [(c, 3 * i, 5) for i in range(4)]
    serial = [_run(ezb, *j) for j in jobs]          # also builds the gridset, the tile plan and the plan cache
    for rep in range(3):
        got = _threaded(ezb, jobs)
        for (rc_s, o_s), (rc_t, o_t) in zip(serial, got):
            assert rc_s == rc_t and np.array_equal(o_s, o_t)


def test_threads_build_different_gridsets_concurrently(ezb, cuda_ok):
    cs = _cases(ezb)                                 # grids defined serially (the reference's rule) ...
    jobs = [(c, i, 3) for i, c in enumerate(cs)]
    got = _threaded(ezb, jobs)                       # ... gridsets, tile plans and wind rotations built by four threads at once
    serial = [_run(ezb, *j) for j in jobs]
    for j, (rc_s, o_s), (rc_t, o_t) in zip(jobs, serial, got):
        assert rc_s == rc_t, j[0].name
        assert np.array_equal(o_s, o_t), j[0].name


def test_thread_options_are_thread_local(ezb, cuda_ok):
    """A thread that switches to nearest-neighbour must not change what another thread's cubic call returns."""
    c_cub = configs.c5(ezb, 1.5, 1.0, "cubic")
    c_nn = configs.c5(ezb, 1.5, 1.0, "nearest")
    ref_cub = _run(ezb, c_cub, 0, 2)[1]
    ref_nn = _run(ezb, c_nn, 0, 2)[1]
    assert not np.array_equal(ref_cub, ref_nn)
    got = _threaded(ezb, [(c_cub, 0, 2), (c_nn, 0, 2), (c_cub, 0, 2), (c_nn, 0, 2)])
    assert np.array_equal(got[0][1], ref_cub) and np.array_equal(got[2][1], ref_cub)
    assert np.array_equal(got[1][1], ref_nn) and np.array_equal(got[3][1], ref_nn)


def test_pageable_and_pinned_host_arrays_agree(ezb, cuda_ok, tiles_only):
    """The pinned ring that carries pageable caller arrays (host_pipeline) must not change a bit."""
    import torch
    c = configs.c2(ezb, 1.5, 1.0, "cubic")
    c.defset()
    nf = 11
    z = np.stack([c.source(k) for k in range(nf)])
    o_pg = np.zeros((nf,) + c.dst_shape, np.float32)
    assert ezb.ezsint_batch(o_pg, z, nf) == 0
    zp = torch.from_numpy(z).pin_memory()
    op = torch.zeros((nf,) + c.dst_shape, dtype=torch.float32).pin_memory()
    assert ezb.ezsint_batch(op, zp, nf) == 0
    assert np.array_equal(op.numpy(), o_pg)
    one = np.zeros(c.dst_shape, np.float32)
    assert ezb.ezsint(one, z[7]) == 0
    assert np.array_equal(one, o_pg[7])
