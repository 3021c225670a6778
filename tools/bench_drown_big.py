"""BDROWN / SMOOTHER on HBM-sized slab stacks (general multi-launch path): algorithmic GB/s of the sweeps and smoothing passes.
    python tools/bench_drown_big.py [nx ny nslab nb_inc nb_smooth]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sosie_b200
from sosie_b200 import grids as G

nx, ny, S, nb_inc, nb_smooth = (int(a) for a in (sys.argv[1:6] + [2560, 1280, 64, 100, 50][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
lon, lat = G.regular_latlon(nx, ny)
X2, Y2 = G.expand_2d(lon, lat)
mask = G.land_mask(X2, Y2)
base = G.field_slabs(X2, Y2, 1)[0]
ctx = sosie_b200.Sosie(0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
X0 = torch.from_numpy(base).to(dev)[None].repeat(S, 1, 1).contiguous()
X0 = X0 + torch.rand_like(X0)
mk = torch.from_numpy(mask).to(dev)
X0 = torch.where(mk[None] == 0, torch.full_like(X0, -9999.0), X0)
X = X0.clone()
ctx.profile(True)
tb, ts, tot = [], [], []
for it in range(4):
    X.copy_(X0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nsw = ctx.bdrown_dev(0, nx, ny, S, X, mk, 0, nb_inc, nb_smooth, -1.0e6, 1.0e6, 0.0, want_nsweeps=True)
    e1.record(); torch.cuda.synchronize()
    if it:
        tb.append(ctx.timing_ms(ctx.T_BDROWN)); ts.append(ctx.timing_ms(ctx.T_SMOOTH)); tot.append(e0.elapsed_time(e1))
n = nx * ny
sweeps = int(nsw.sum())
out = {"shape": [nx, ny, S], "land_frac": float((mask == 0).mean()), "sweeps_executed_total": sweeps, "sweeps_per_slab": int(nsw[0]),
       "sweeps_ms": float(np.mean(tb)), "smooth_ms": float(np.mean(ts)), "call_ms": float(np.mean(tot)),
       "sweeps_algorithmic_gbs": 10 * n * sweeps / (np.mean(tb) * 1e-3) / 1e9,
       "smooth_algorithmic_gbs": 10 * n * S * nb_smooth / (np.mean(ts) * 1e-3) / 1e9 if nb_smooth else None,
       "launches": ctx.launches}
print(json.dumps(out))
