"""One rank's share of the N-GPU strong-scaling step of config E, run alone: step time for a 1/N row shard (fixed per-step overhead)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sosie_b200
from sosie_b200 import grids as G
from sosie_b200.shard import row_shard
dev = torch.device("cuda", 0)
cfg = G.config("E", 1.0)
nx1, ny1, nx2, ny2 = cfg["nxs"], cfg["nys"], cfg["nxt"], cfg["nyt"]
d_lon, d_lat = torch.from_numpy(cfg["src_lon"]).to(dev), torch.from_numpy(cfg["src_lat"]).to(dev)
d_Xt, d_Yt = torch.from_numpy(cfg["trg_lon"]).to(dev), torch.from_numpy(cfg["trg_lat"]).to(dev)
d_Z1 = torch.rand((ny1, nx1), device=dev, dtype=torch.float32)
d_Z2 = torch.empty((ny2, nx2), device=dev, dtype=torch.float32)
ctx = sosie_b200.Sosie(0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
out = {}
for world in (1, 2, 4, 8):
    r0, r1 = row_shard(ny2, 0, world)
    ctx.bilin_set_shard(r0, r1) if world > 1 else ctx.bilin_set_shard(0, 0)
    def step():
        ctx.bilin_set_grids_dev(cfg["ewper_src"], nx1, ny1, d_lon, d_lat, 1, nx2, ny2, d_Xt, d_Yt, 0, None, l_reg_src=True)
        ctx.bilin_map()
        ctx.bilin_apply_dev(1, d_Z1, d_Z2, first_call=True)
    for _ in range(3): step()
    torch.cuda.synchronize()
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 10
    ms = e0.elapsed_time(e1) / 10
    km, kp, ka = ctx.timing_ms(ctx.T_MAP), ctx.timing_ms(ctx.T_PACK), ctx.timing_ms(ctx.T_APPLY)
    ctx.profile(False)
    out[f"1/{world}"] = {"step_ms": ms, "wall_ms": wall * 1e3, "map_ms": km, "pack_ms": kp, "apply_ms": ka, "overhead_ms": ms - km - max(kp, 0.0) - ka}
print(json.dumps(out))
