#!/usr/bin/env python
"""Work per target-row shard of the E mapping on ONE GPU: the 8 equal-row shards of an 8-GPU run mapped one after the other
(is the N = 8 step bounded by an unlucky shard?).  Prints ms per shard for equal-row shards."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sosie_b200
from sosie_b200 import grids as G
from sosie_b200.shard import row_shard

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = G.config("E", 1.0)
dev = torch.device("cuda", 0)
ctx = sosie_b200.Sosie(0)
ctx.profile(True)
ny2 = cfg["nyt"]
ctx.bilin_set_grids(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], cfg["trg_lon"], cfg["trg_lat"], l_reg_src=cfg["l_reg_src"])
out = []
for r in range(W):
    j0, j1 = row_shard(ny2, r, W)
    ctx.bilin_set_shard(j0, j1)
    ts = []
    for _ in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.bilin_map(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out.append((j0, j1, min(ts[1:])))
    print(f"shard {r}: rows {j0}-{j1}  map call {min(ts[1:]):.3f} ms", flush=True)
tot = sum(o[2] for o in out)
print(f"sum {tot:.3f} ms, max {max(o[2] for o in out):.3f} ms, mean {tot / W:.3f} ms -> balance efficiency {tot / W / max(o[2] for o in out):.3f}")
