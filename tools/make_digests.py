#!/usr/bin/env python
"""Full-size parity pins: run the CPU oracle in BOTH libm modes (glibc = the reference platform, pmath = the portable libm
the CUDA kernels use) over EVERY target of configs E (39.5 M targets) and D (1.2 M), and store SHA-256 digests of
metrics / ID_problem / alpha,beta / dist_np / the interpolated field per block of 64 target rows in tests/golden/.

    python tools/make_digests.py [E] [D] [--threads N]

tests/test_gpu_fullsize.py::test_config_E_all_rows_vs_digests compares the COMPLETE GPU mapping with them (all 4729 rows);
tests/test_golden.py checks the file's own consistency on the CPU.  The inputs are the portable-bit grids of
tests/portable.py (same shapes and construction as sosie_b200.grids).  The blocks are mapped independently: on a regular
source the per-target computation (EASY search + MAPPING_BL body) does not depend on the other targets, and the search
prelude keeps every row of these targets (both facts are re-checked by the GPU test, which maps the whole grid at once).

Recorded per config: the digests of the pmath run (what the GPU must reproduce bit for bit), the digests of the glibc run,
and per array the number of elements on which the two libm modes differ.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from sosie_b200 import grids as G  # noqa: E402
from tests import portable as P  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BLOCK = 64


def inputs(name):
    """(cfg-like dict, Z1 (nslab, ny1, nx1) f32) with portable bits"""
    if name == "E":
        nx1, ny1, nx2, ny2 = 21600, 10800, 8354, 4729
        lon, lat = G.regular_latlon(nx1, ny1)
        tl, tt = P.enatl_like(nx2, ny2)
        return dict(name="E", src_lon=lon, src_lat=lat, trg_lon=tl, trg_lat=tt, ewper_src=0, nslab=1), P.integer_field(nx1, ny1, 1)
    if name == "D":
        nx1, ny1, nx2, ny2 = 2560, 1280, 2560, 480
        lon, lat = G.regular_latlon(nx1, ny1)
        tl, tt = P.polar_stereo(nx2, ny2)
        return dict(name="D", src_lon=lon, src_lat=lat, trg_lon=tl, trg_lat=tt, ewper_src=0, nslab=2), P.integer_field(nx1, ny1, 2, lo=230, span=81)
    raise ValueError(name)


def run(name, nthreads):
    cfg, Z1 = inputs(name)
    tl, tt = cfg["trg_lon"], cfg["trg_lat"]
    ny2, nx2 = tl.shape
    rec = {"config": name, "block_rows": BLOCK, "nx_trg": nx2, "ny_trg": ny2, "nx_src": int(cfg["src_lon"].size), "ny_src": int(cfg["src_lat"].size),
           "inputs": {"trg_lon": P.sha(tl), "trg_lat": P.sha(tt), "src_lon": P.sha(cfg["src_lon"]), "src_lat": P.sha(cfg["src_lat"]), "Z1": P.sha(Z1)},
           "blocks": [], "libm_differences": {k: 0 for k in ("metrics", "ipb", "alphabeta", "dist_np", "Z2")}, "targets": nx2 * ny2}
    t0 = time.time()
    for j0 in range(0, ny2, BLOCK):
        j1 = min(ny2, j0 + BLOCK)
        Xb, Yb = np.ascontiguousarray(tl[j0:j1]), np.ascontiguousarray(tt[j0:j1])
        res = {}
        for mode, tag in ((O.PMATH, "pmath"), (O.GLIBC, "glibc")):
            O.set_libm(mode)
            Z2, mp = O.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, Xb, Yb, l_reg_src=True, nthreads=nthreads, virtual_src=True)
            res[tag] = dict(metrics=mp["metrics"], ipb=mp["ipb"], alphabeta=mp["alphabeta"], dist_np=mp["dist_np"], Z2=Z2)
        blk = {"rows": [j0, j1]}
        for tag in ("pmath", "glibc"):
            blk[tag] = {k: P.sha(v) for k, v in res[tag].items()}
        for k in rec["libm_differences"]:
            a, b = res["pmath"][k], res["glibc"][k]
            rec["libm_differences"][k] += int(np.count_nonzero(a != b))
        rec["blocks"].append(blk)
        done = j1 * nx2
        sys.stderr.write(f"\r{name}: rows {j1}/{ny2}  {done / max(time.time() - t0, 1e-9):.0f} targets/s (2 modes)   ")
    sys.stderr.write("\n")
    rec["seconds"] = round(time.time() - t0, 1)
    rec["threads"] = nthreads
    rec["note"] = ("libm_differences: elements on which the glibc-mode and pmath-mode oracle disagree over the whole grid "
                   "(metrics/ipb/alphabeta/Z2 must be 0; dist_np differs in its last REAL(4) bit where acos amplifies a 1-ulp "
                   "difference of sin/cos)")
    return rec


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in ("E", "D")] or ["D", "E"]
    nthr = O.max_threads()
    if "--threads" in sys.argv:
        nthr = int(sys.argv[sys.argv.index("--threads") + 1])
    O.build()
    os.makedirs(OUT, exist_ok=True)
    for n in names:
        rec = run(n, nthr)
        p = os.path.join(OUT, f"digest_{n}.json")
        with open(p, "w") as f:
            json.dump(rec, f, indent=0)
        print(p, os.path.getsize(p) // 1024, "KiB", rec["libm_differences"], f"{rec['seconds']} s")
