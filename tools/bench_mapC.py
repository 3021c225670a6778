import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sosie_b200
from sosie_b200 import grids as G
name = sys.argv[1] if len(sys.argv) > 1 else "C"
cfg = G.config(name, 1.0)
ctx = sosie_b200.Sosie(0)
ctx.bilin_set_grids(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], cfg["trg_lon"], cfg["trg_lat"], l_reg_src=cfg["l_reg_src"], i_orca_trg=cfg["i_orca_trg"])
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); ctx.bilin_map(); ctx.sync(); print("map %s: %.3f ms" % (name, (time.perf_counter() - t) * 1e3))
print(ctx.bilin_get_map()["info"])
