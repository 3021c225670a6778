"""Multi-GPU partitioning of the hot path: one process per GPU, no data-path collective.

* mapping on a regular source (EASY search): every target is independent -> contiguous blocks of target ROWS per rank
  (``row_shard``); each rank passes its block to ``Sosie.bilin_set_shard``.
* apply / drown / Akima evaluation: slabs (records x levels) are independent -> contiguous slab ranges per rank
  (``slab_shard``); the mapping is replicated, broadcast once from the rank that built it (``broadcast_mapping``,
  NCCL over NVLink on GPUs, gloo in the CPU tests) -- the only collective of the path.
* outputs: a single host gather (each rank copies its rows/slabs into the caller's arrays).
"""
from __future__ import annotations


def row_shard(ny: int, rank: int, world: int):
    """1-based inclusive target rows (j0, j1) of `rank`.  More ranks than rows would leave ranks without work while the
    prelude exchange of a sharded first call is a collective that all ranks or none must enter: refused here, with the
    same arguments on every rank, i.e. before any rank reaches a collective."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"row_shard: rank {rank} of {world}")
    if world > ny:
        raise ValueError(f"row_shard: {world} ranks for {ny} target rows (every rank needs at least one row)")
    j0 = (ny * rank) // world + 1
    j1 = (ny * (rank + 1)) // world
    return j0, j1


def slab_shard(nslab: int, rank: int, world: int):
    """0-based half-open slab range [s0, s1) of `rank`."""
    return (nslab * rank) // world, (nslab * (rank + 1)) // world


def broadcast_mapping(tensors, src: int = 0, group=None):
    """Broadcast the mapping arrays (metrics i32, alphabeta f64, ipb i16) from `src` to every rank, in place.
    `tensors` is a list of torch tensors already allocated with the same shapes on every rank."""
    import torch
    import torch.distributed as dist
    # byte views: the payload is opaque to the collective (int16 is not a gloo/NCCL reduction type)
    works = [dist.broadcast(t.view(torch.uint8), src=src, group=group, async_op=True) for t in tensors]
    for w in works:
        w.wait()
    return tensors


def gather_rows(local_rows, ny: int, world: int, group=None):
    """All-gather per-rank row blocks (torch tensors shaped (rows_r, nx)) into the full (ny, nx) array."""
    import torch
    import torch.distributed as dist
    nx = local_rows.shape[1]
    sizes = [row_shard(ny, r, world) for r in range(world)]
    maxr = max(j1 - j0 + 1 for j0, j1 in sizes)
    pad = torch.zeros((maxr, nx), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([out[r][: sizes[r][1] - sizes[r][0] + 1] for r in range(world)], dim=0)
