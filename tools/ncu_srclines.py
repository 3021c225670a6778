#!/usr/bin/env python
"""Per-source-line shares of executed warp-instructions and stall samples of the kernels in an ncu report that was taken with
--import-source on (needs -lineinfo):   python tools/ncu_srclines.py REPORT.ncu-rep [top] [kernel-substring]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
h, out, fn, keep = None, [], "", True
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "Function Name":
        fn = r[1]; keep = want in fn
    if r and r[0] == "Line No":
        h = r; continue
    if h and keep and len(r) == len(h) and r[0].isdigit():
        g = lambda k: r[h.index(k)] if k in h else "0"   # (a kernel without shared memory has no such column)
        out.append((int(g("Instructions Executed") or 0), int(g("# Samples") or 0), r[0], r[1][:120], g("L1 Wavefronts Shared Excessive"), fn.split("(")[0]))
tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print(f"total warp-instructions {tot}, stall samples {ts}")
for o in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{100 * o[0] / tot:5.1f}% inst {100 * o[1] / ts:5.1f}% smp  L{o[2]:>4} excess_smem_wavefronts {o[4]:>9} | {o[3]}")
