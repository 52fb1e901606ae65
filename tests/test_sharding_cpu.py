"""world_size-2 `gloo` tests of the multi-GPU host logic (level partition + gather of unequal shares)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from librmn_b200 import sharding  # noqa: E402


def test_level_partition_covers_everything_once():
    for n in (1, 2, 5, 80, 81, 1000):
        for world in (1, 2, 3, 8):
            spans = [sharding.level_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeLib:
    """Stands in for libezb200 on CPU: 'interpolation' = every output point is 2 * level mean + point index."""

    def ezsint_batch(self, out, zin, n):
        out.copy_(2.0 * zin.mean(dim=1, keepdim=True) + torch.arange(out.shape[1], dtype=torch.float32)[None])
        return 0


def _worker(rank, world, port, nfields, n_src, n_out, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.level_range(nfields, rank, world)
        full_in = torch.arange(nfields * n_src, dtype=torch.float32).reshape(nfields, n_src)
        out = sharding.interpolate_sharded(_FakeLib(), full_in[lo:hi].clone(), n_out, nfields, gather=True)
        expect = 2.0 * full_in.mean(dim=1, keepdim=True) + torch.arange(n_out, dtype=torch.float32)[None]
        ok = torch.equal(out, expect)
        local = sharding.interpolate_sharded(_FakeLib(), full_in[lo:hi].clone(), n_out, nfields, gather=False)
        ok = ok and torch.equal(local, expect[lo:hi])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather_of_unequal_shares():
    world, nfields = 2, 5
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, nfields, 7, 11, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_row_bands_are_whole_tile_rows_and_cover_the_target():
    for nj in (1, 31, 32, 33, 1600, 1801, 2048):
        for world in (1, 2, 3, 8):
            spans = [sharding.row_band(nj, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nj
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(j0 % sharding.TILE_ROWS == 0 for j0, _ in spans if j0 < nj)


class _FakeRowsLib:
    """Stands in for libezb200 on CPU: output value = 1000 * row + column (+ field mean), restricted to the set band."""

    def __init__(self):
        self.rows = None

    def set_rows(self, j0=-1, j1=-1):
        self.rows = None if j0 < 0 else (j0, j1)
        return 0

    def ezsint(self, out, zin):
        j0, j1 = self.rows
        out.copy_(1000.0 * torch.arange(j0, j1, dtype=torch.float32)[:, None] + torch.arange(out.shape[1], dtype=torch.float32)[None] + zin.mean())
        return 0


def _rows_worker(rank, world, port, nj, ni, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        zin = torch.full((4, 5), 3.0)
        full = sharding.interpolate_rows(_FakeRowsLib(), zin, ni, nj, gather=True)
        expect = 1000.0 * torch.arange(nj, dtype=torch.float32)[:, None] + torch.arange(ni, dtype=torch.float32)[None] + 3.0
        ret[rank] = bool(torch.equal(full, expect))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather_of_row_bands():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_rows_worker, args=(2, port, 97, 13, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
