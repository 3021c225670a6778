"""Multi-GPU host logic on CPU: the partitioning helpers, and a world_size-2 gloo run in which each rank computes its
shard of the hot path (with the oracle standing in for the device) and the gathered result equals the unsharded one."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sosie_b200 import grids as G
from sosie_b200.shard import row_shard, slab_shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_exactly():
    for n in (1, 7, 480, 4729):
        for world in (1, 2, 3, 8):
            if world > n:   # more ranks than rows: refused identically on every rank, before any collective
                for r in range(world):
                    with pytest.raises(ValueError):
                        row_shard(n, r, world)
                continue
            rows = [row_shard(n, r, world) for r in range(world)]
            covered = [j for j0, j1 in rows for j in range(j0, j1 + 1)]
            assert covered == list(range(1, n + 1))
            slabs = [slab_shard(n, r, world) for r in range(world)]
            assert [s for a, b in slabs for s in range(a, b)] == list(range(n))
            assert max(b - a for a, b in slabs) - min(b - a for a, b in slabs) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from sosie_b200.shard import broadcast_mapping, gather_rows
    O.set_libm(O.PMATH)
    cfg = G.config("D", 0.04)
    Xs, Ys = G.src_2d(cfg)
    S = 6
    Z1 = G.field_slabs(Xs, Ys, S, seed=5)
    ny2, nx2 = cfg["nyt"], cfg["nxt"]
    # (1) mapping: rank 0 builds it, the mapping arrays are broadcast (the path's only collective)
    met = torch.zeros((3, ny2, nx2), dtype=torch.int32); ab = torch.zeros((2, ny2, nx2), dtype=torch.float64)
    ipb = torch.zeros((ny2, nx2), dtype=torch.int16)
    if rank == 0:
        _, m = O.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1[:1], cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
        met.copy_(torch.from_numpy(m["metrics"])); ab.copy_(torch.from_numpy(m["alphabeta"])); ipb.copy_(torch.from_numpy(m["ipb"]))
    broadcast_mapping([met, ab, ipb], src=0)
    mapping = dict(metrics=met.numpy(), alphabeta=ab.numpy(), ipb=ipb.numpy(), dist_np=np.zeros((ny2, nx2), np.float32))
    # (2) apply: slabs sharded, mapping replicated
    s0, s1 = slab_shard(S, rank, world)
    Z2, _ = O.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1[s0:s1], cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True,
                       mapping=mapping)
    parts = [None] * world
    dist.all_gather_object(parts, Z2)
    # (3) target-row sharding of a single slab (config E style): gather rows
    j0, j1 = row_shard(ny2, rank, world)
    rows = torch.from_numpy(np.ascontiguousarray(Z2[0][j0 - 1:j1])) if s1 > s0 else torch.zeros((j1 - j0 + 1, nx2))
    full_rows = gather_rows(torch.from_numpy(np.ascontiguousarray(parts[0][0][j0 - 1:j1])), ny2, world)
    if rank == 0:
        q.put((np.concatenate(parts, 0), full_rows.numpy(), mapping["metrics"].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_sharded_equals_unsharded(orc):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, rows, met = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    orc.set_libm(orc.PMATH)
    cfg = G.config("D", 0.04)
    Xs, Ys = G.src_2d(cfg)
    Z1 = G.field_slabs(Xs, Ys, 6, seed=5)
    ref, m = orc.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
    assert np.array_equal(got, ref)
    assert np.array_equal(rows, ref[0])
    assert np.array_equal(met, m["metrics"])
