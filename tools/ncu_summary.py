#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small tracked text files under profiles/.
    python tools/ncu_summary.py rep  gpurun_out/prof.ncu-rep  profiles/r01_xxx.md  "title"
    python tools/ncu_summary.py list gpurun_out/launches.csv   profiles/r01_launches.md "title"
    python tools/ncu_summary.py json gpurun_out/prof.ncu-rep  profiles/r02_ncu_E.json bilin_map.cu,geom.cuh,pmath.h,common.cuh
        (machine-readable twin read by bench.py: per kernel the KEYS below + the digest of the named csrc files, so that a
         capture of an older kernel is recognised as stale)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def rep(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `{path}` (ncu --set full --clock-control none; per-launch, cold cache, serialised)\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']].split('(')[0]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(r[idx[h]] or 0))
                  for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
            st.sort(key=lambda x: -x[1])
            f.write("\ntop stall reasons (warps per issue-active cycle): " + ", ".join(f"{a} {b:.2f}" for a, b in st[:6]) + "\n\n")


def lst(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "us" else v / 1e6 if u == "ns" else v * 1e3 if u == "s" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: "
                f"compare SHARES)\n\ntotal {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k[:80]} | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f} % |\n")


def js(path, out, files):
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from sosie_b200 import build as B
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    kernels = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        d = {}
        for k in KEYS:
            if k in idx and r[idx[k]] != "":
                v = float(r[idx[k]].replace(",", ""))
                u = units[idx[k]]
                if k.startswith("dram__bytes"):
                    v *= scale.get(u, 1.0); u = "byte"
                if k == "gpu__time_duration.sum":
                    v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(u, 1.0); u = "ms"
                d[k] = v
        kernels.setdefault(name, d)      # first launch of each kernel
    extra = {}
    names = []
    for f in files.split(","):       # "slabs=292"-style items are copied into every kernel's record
        if "=" in f:
            k_, v_ = f.split("=", 1)
            extra[k_] = float(v_) if v_.replace(".", "", 1).isdigit() else v_
        elif f:
            names.append(f)
    for d in kernels.values():
        d.update(extra)
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    with open(out, "w") as f:
        json.dump({"source": path, "how": "ncu --set full --clock-control none, one launch per kernel (cold cache, serialised)",
                   "git_head_at_summary": head, "csrc_files": names, "csrc_digest": B.source_digest(names), "kernels": kernels}, f, indent=1)


if __name__ == "__main__":
    {"rep": rep, "list": lst, "json": js}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
