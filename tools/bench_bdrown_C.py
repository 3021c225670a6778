import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sosie_b200
from sosie_b200 import grids as G
dev = torch.device("cuda", 0)
ctx = sosie_b200.Sosie(0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
cfg = G.config("C", 1.0); S = cfg["nslab"]
base = G.field_slabs(cfg["src_lon"], cfg["src_lat"], 1)[0]
masks = np.stack([G.land_mask(cfg["src_lon"], cfg["src_lat"], thresh=0.25 - 0.4 * k / 31) for k in range(31)])
X0 = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(base, (S,) + base.shape))).to(dev)
X0 = X0 + torch.rand_like(X0)
mk = torch.from_numpy(np.ascontiguousarray(np.tile(masks, (12, 1, 1)))).to(dev)
X0 = torch.where(mk == 0, torch.full_like(X0, -9999.0), X0)
X = X0.clone()
ts = []
for it in range(12):
    X.copy_(X0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.bdrown_dev(cfg["ewper_src"], cfg["nxs"], cfg["nys"], S, X, mk, 1, 20, 50, -1e3, 1e3, 0.0)
    e1.record(); torch.cuda.synchronize()
    if it >= 2: ts.append(e0.elapsed_time(e1))
n1 = cfg["nxs"] * cfg["nys"]
print(json.dumps({"bdrown_C_ms_min": float(np.min(ts)), "median": float(np.median(ts)), "algorithmic_gbs": 10 * n1 * S * 70 / (np.min(ts) * 1e-3) / 1e9}))
