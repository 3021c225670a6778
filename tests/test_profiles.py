"""The machine-readable ncu captures bench.py reads (profiles/*_ncu_*.json) are well formed: a capture of the batched apply
without its slab count would silently turn `others.D.frac_hbm_dram` into null."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ncu_twins_are_well_formed():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_*.json")))
    assert files, "no profiles/*_ncu_*.json"
    for f in files:
        d = json.load(open(f))
        assert d.get("csrc_files") and d.get("csrc_digest") and d.get("kernels"), f
        for name, m in d["kernels"].items():
            assert m.get("gpu__time_duration.sum", 0) > 0 and "dram__bytes_read.sum" in m and "dram__bytes_write.sum" in m, (f, name)
            if name.startswith("k_apply_grp"):
                assert m.get("slabs", 0) >= 1, f"{f}: {name} has no slab count (tools/ncu_summary.py json ... files,slabs=N)"
