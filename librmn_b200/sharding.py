"""Multi-GPU partitioning of the hot path: one process per GPU, every rank holds a replicated
gridset and interpolates its own slice of the levels / fields; there is NO data-path collective.  An optional
gather (torch.distributed all_gather over NCCL/NVLink on GPU tensors, gloo on CPU tensors in tests) assembles
one contiguous [nfields, n_out] array when the caller asks for it.
"""
import torch
import torch.distributed as dist


def level_range(nfields, rank, world):
    """Contiguous block partition of `nfields` levels: rank r owns [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(nfields, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


TILE_ROWS = 32   # the tile kernels work on 32 x 32 target tiles: bands made of whole tile rows keep them (c_ezb200_set_rows)


def row_band(nj, rank, world, align=TILE_ROWS):
    """Output-row band partition for calls with fewer fields than GPUs: rank r owns target rows [j0, j1), a whole
    number of `align`-row tile rows (the last band takes the ragged rest).  Bands may be empty when nj < align * world."""
    lo, hi = level_range((nj + align - 1) // align, rank, world)
    return min(nj, lo * align), min(nj, hi * align)


def interpolate_rows(lib, zin, ni, nj, gather=False, group=None, vector=None):
    """One field (or one (u, v) pair: zin = (u, v), vector=True) with the target rows split over the ranks.

    Every rank holds the whole source field and a replicated gridset, produces its own band of target rows
    (c_ezb200_set_rows) and -- with gather=True -- receives everybody's bands (all_gather over NCCL / NVLink)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    j0, j1 = row_band(nj, rank, world)
    dev = (zin[0] if vector else zin).device
    outs = [torch.empty(((j1 - j0), ni), dtype=torch.float32, device=dev) for _ in range(2 if vector else 1)]
    lib.set_rows(j0, j1)
    try:
        if j1 > j0:
            rc = lib.ezuvint(outs[0], outs[1], zin[0], zin[1]) if vector else lib.ezsint(outs[0], zin)
            if rc < 0:
                raise RuntimeError("row-band call failed on rank %d (rc=%d)" % (rank, rc))
    finally:
        lib.set_rows()
    if not gather or world == 1:
        return outs if vector else outs[0]
    full = [gather_rows(o, nj, group) for o in outs]
    return full if vector else full[0]


def gather_rows(band, nj, group=None):
    """all_gather of per-rank row bands of unequal height into one [nj, ni] tensor (bands are contiguous in row-major order)."""
    world = dist.get_world_size(group)
    ni = band.shape[1]
    spans = [row_band(nj, r, world) for r in range(world)]
    biggest = max(j1 - j0 for j0, j1 in spans)
    pad = torch.zeros((biggest, ni), dtype=band.dtype, device=band.device)
    pad[: band.shape[0]] = band
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    full = torch.empty((nj, ni), dtype=band.dtype, device=band.device)
    for r, (j0, j1) in enumerate(spans):
        full[j0:j1] = bufs[r][: j1 - j0]
    return full


def interpolate_sharded(lib, zin_local, n_out, nfields_total, gather=False, group=None):
    """Interpolate this rank's levels (zin_local: [n_local, n_src] CUDA tensor) with the current gridset.

    Returns the local output [n_local, n_out]; with gather=True returns the full [nfields_total, n_out] array on
    every rank (ranks may own different numbers of levels, so the gather pads to the largest share)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = level_range(nfields_total, rank, world)
    n_local = hi - lo
    assert zin_local.shape[0] == n_local, "rank %d owns levels [%d,%d) but got %d" % (rank, lo, hi, zin_local.shape[0])
    out = torch.empty((n_local, n_out), dtype=torch.float32, device=zin_local.device)
    if n_local:
        rc = lib.ezsint_batch(out, zin_local.contiguous(), n_local)
        if rc < 0:
            raise RuntimeError("c_ezsint_batch failed on rank %d (rc=%d)" % (rank, rc))
    if not gather or world == 1:
        return out
    return gather_levels(out, nfields_total, group)


def gather_levels(out_local, nfields_total, group=None):
    """all_gather of per-rank level blocks of unequal size into one [nfields_total, n_out] tensor."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    n_out = out_local.shape[1]
    shares = [level_range(nfields_total, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in shares)
    pad = torch.zeros((biggest, n_out), dtype=out_local.dtype, device=out_local.device)
    pad[: out_local.shape[0]] = out_local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    full = torch.empty((nfields_total, n_out), dtype=out_local.dtype, device=out_local.device)
    for r, (lo, hi) in enumerate(shares):
        full[lo:hi] = bufs[r][: hi - lo]
    assert shares[rank][1] - shares[rank][0] == out_local.shape[0]
    return full
