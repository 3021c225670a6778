"""SURVEY 8e on real GPUs: records sharded over the ranks, the mapping built ONCE on rank 0 and broadcast GPU -> GPU with NCCL
(the only collective of the path), every rank applies its own slab range; the result of each rank is checked against a
single-GPU run of the same slabs.   torchrun --nproc-per-node N tools/multi_gpu_records.py [records]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import sosie_b200
from sosie_b200 import grids as G
from sosie_b200.shard import slab_shard, broadcast_mapping

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1460
cfg = G.config("D", 1.0)
nx1, ny1, nx2, ny2 = cfg["nxs"], cfg["nys"], cfg["nxt"], cfg["nyt"]
ctx = sosie_b200.Sosie(local)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
met = torch.empty((3, ny2, nx2), dtype=torch.int32, device=dev)
ab = torch.empty((2, ny2, nx2), dtype=torch.float64, device=dev)
ipb = torch.empty((ny2, nx2), dtype=torch.int16, device=dev)
ctx.bilin_set_grids(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
t_map = t_bc = 0.0
if rank == 0:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.bilin_map()
    ctx.bilin_get_map_dev(met, ab, ipb)
    torch.cuda.synchronize(); t_map = time.perf_counter() - t0
if world > 1:
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    broadcast_mapping([met, ab, ipb], src=0)
    torch.cuda.synchronize(); t_bc = time.perf_counter() - t0
    if rank != 0:
        ctx.bilin_load_map_dev(nx2, ny2, met, ab, ipb)
s0, s1 = slab_shard(S, rank, world)
ns = s1 - s0
g = torch.Generator(device=dev); g.manual_seed(1234)
# every rank generates the records of ITS range from a per-record seed, so that a single-GPU check can regenerate them
def records(a, b):
    out = torch.empty((b - a, ny1, nx1), device=dev, dtype=torch.float32)
    for r in range(a, b):
        g.manual_seed(1000 + r)
        out[r - a] = torch.rand((ny1, nx1), device=dev, generator=g) * 80 + 230
    return out
Z1 = records(s0, s1)
Z2 = torch.empty((ns, ny2, nx2), device=dev, dtype=torch.float32)
ctx.bilin_apply_dev(ns, Z1, Z2, first_call=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ctx.bilin_apply_dev(ns, Z1, Z2)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# check: a checksum of this rank's slabs against the same slabs applied with a mapping built locally
chk = float(Z2.double().sum().item())
ctx2 = sosie_b200.Sosie(local); ctx2.set_stream(st.cuda_stream)
ctx2.bilin_set_grids(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
ctx2.bilin_map()
k = min(ns, 8)
Zc = torch.empty((k, ny2, nx2), device=dev, dtype=torch.float32)
ctx2.bilin_apply_dev(k, Z1[:k], Zc, first_call=True)
torch.cuda.synchronize()
same = bool(torch.equal(Zc, Z2[:k]))
flags = [same]
if world > 1:
    allf = [None] * world
    dist.all_gather_object(allf, same)
    flags = allf
if rank == 0:
    nt = nx2 * ny2
    print(json.dumps({"workload": f"D: {S} records sharded over {world} GPU(s), mapping built on rank 0 and broadcast with NCCL",
                      "n_gpus": world, "apply_ms_max_over_ranks": float(t.item()), "points_per_s": S * nt / (float(t.item()) * 1e-3),
                      "map_ms_rank0": t_map * 1e3, "nccl_broadcast_ms": t_bc * 1e3, "mapping_bytes": int(met.numel() * 4 + ab.numel() * 8 + ipb.numel() * 2),
                      "broadcast_gbs": (met.numel() * 4 + ab.numel() * 8 + ipb.numel() * 2) / max(t_bc, 1e-9) / 1e9 if world > 1 else None,
                      "every_rank_equals_local_mapping": all(flags)}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
