#!/usr/bin/env python
"""Kernel micro-benchmarks used while tuning (not the driver's bench): D-shaped batched apply, C-shaped BDROWN."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sosie_b200  # noqa: E402
from sosie_b200 import grids as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="applyD")
ap.add_argument("--slabs", type=int, default=292)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--tma", type=int, default=1, help="0: batched apply on the cp.async-staged kernel only")
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = sosie_b200.Sosie(0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx.set_stream(st.cuda_stream)
ctx.profile(True)
HBM = 6539.2
if a.what == "applyD":
    cfg = G.config("D", 1.0)
    S = a.slabs
    n1, nt = cfg["nxs"] * cfg["nys"], cfg["nxt"] * cfg["nyt"]
    ctx.bilin_set_grids(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
    ctx.bilin_map()
    m = ctx.bilin_get_map()
    jP = m["metrics"][1]
    print("source rows touched:", jP.min(), jP.max(), "of", cfg["nys"])
    Z1 = torch.rand((S, cfg["nys"], cfg["nxs"]), device=dev) * 80 + 230
    Z2 = torch.empty((S, cfg["nyt"], cfg["nxt"]), device=dev)
    ctx.apply_tma(int(a.tma))
    ctx.bilin_apply_dev(S, Z1, Z2, first_call=True)
    print("tile classes (tma, other):", ctx.apply_tile_classes(), "tma", a.tma)
    ts = []
    for _ in range(a.reps):
        ctx.bilin_apply_dev(S, Z1, Z2)
        ts.append(ctx.timing_ms(ctx.T_APPLY))
    t = float(np.mean(ts))
    algo = S * (4 * n1 + 4 * nt) + 32 * nt
    touched = S * (4 * cfg["nxs"] * (int(jP.max()) - int(jP.min()) + 2) + 4 * nt) + 32 * nt
    print(f"apply D S={S}: {t:.3f} ms  algorithmic {algo/t/1e6:.0f} GB/s ({algo/t/1e6/HBM:.3f} of HBM)  touched {touched/t/1e6:.0f} GB/s ({touched/t/1e6/HBM:.3f})")
elif a.what == "bdrownC":
    cfg = G.config("C", 1.0)
    S = cfg["nslab"]
    n1 = cfg["nxs"] * cfg["nys"]
    base = G.field_slabs(cfg["src_lon"], cfg["src_lat"], 1)[0]
    masks = np.stack([G.land_mask(cfg["src_lon"], cfg["src_lat"], thresh=0.25 - 0.4 * k / 31) for k in range(31)])
    X0 = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(base, (S,) + base.shape))).to(dev)
    X0 = X0 + torch.rand_like(X0)
    mk = torch.from_numpy(np.ascontiguousarray(np.tile(masks, (12, 1, 1)))).to(dev)
    X0 = torch.where(mk == 0, torch.full_like(X0, -9999.0), X0)
    X = X0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(a.reps):
        X.copy_(X0)
        e0.record()
        nsw = ctx.bdrown_dev(cfg["ewper_src"], cfg["nxs"], cfg["nys"], S, X, mk, 1, cfg["nb_inc"], cfg["nb_smooth"], -1e3, 1e3, 0.0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.mean(ts[1:]))
    nsw = ctx.bdrown_dev(cfg["ewper_src"], cfg["nxs"], cfg["nys"], S, X, mk, 1, cfg["nb_inc"], cfg["nb_smooth"], -1e3, 1e3, 0.0, want_nsweeps=True)
    algo = 10 * n1 * (int(nsw.sum()) + S * cfg["nb_smooth"]) + 8 * n1 * S
    print(f"bdrown C ({S} slabs): {t:.3f} ms; sweeps {ctx.timing_ms(ctx.T_BDROWN):.3f} ms smooth {ctx.timing_ms(ctx.T_SMOOTH):.3f} ms; algorithmic {algo/t/1e6:.0f} GB/s ({algo/t/1e6/HBM:.2f} of HBM)")
