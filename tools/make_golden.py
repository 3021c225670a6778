#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the CPU oracle (glibc libm mode).

The reference ships no golden vectors for this path and cannot be compiled here, so these
fixtures pin the ORACLE against regressions: tests/test_golden.py re-runs the oracle (both libm modes) and the
GPU tests re-run the CUDA path against the same files.  Regenerate with:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from sosie_b200 import grids as G  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = [("A", 0.12), ("B", 0.12), ("C", 0.15), ("D", 0.03), ("E", 0.006)]


def make_case(name, scale):
    cfg = G.config(name, scale)
    Xs, Ys = G.src_2d(cfg)
    Z1 = G.field_slabs(Xs, Ys, 2, seed=123)
    Z2, mp = O.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"],
                        l_reg_src=cfg["l_reg_src"], i_orca_trg=cfg["i_orca_trg"])
    out = dict(scale=scale, Z1=Z1, Z2=Z2, metrics=mp["metrics"].astype(np.int16) if mp["metrics"].max() < 32000 else mp["metrics"],
               alphabeta=mp["alphabeta"], ipb=mp["ipb"].astype(np.int8), dist_np=mp["dist_np"], info=mp["info"])
    mask = G.land_mask(Xs, Ys)
    Zm = Z1.copy()
    Zm[:, mask == 0] = -9999.0
    Xd, nsw = O.bdrown(cfg["ewper_src"], Zm, mask, 12, 6, -1.0e3, 1.0e3, 0.0)
    out.update(bd_mask=mask.astype(np.int8), bd_out=Xd, bd_nsw=nsw)
    if cfg["l_reg_src"]:
        Za, ixy = O.akima_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"])
        out.update(ak_out=Za, ak_ixy=ixy.astype(np.int16))
    return out


if __name__ == "__main__":
    O.build()
    O.set_libm(O.GLIBC)
    os.makedirs(OUT, exist_ok=True)
    for name, scale in CASES:
        d = make_case(name, scale)
        p = os.path.join(OUT, f"case_{name}.npz")
        np.savez_compressed(p, **d)
        print(p, os.path.getsize(p) // 1024, "KiB", {k: getattr(v, "shape", v) for k, v in d.items() if k in ("Z2", "metrics")})
