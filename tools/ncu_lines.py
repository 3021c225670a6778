#!/usr/bin/env python
"""Attribute the warp-stall samples / executed instructions of one kernel in an ncu report to source lines.

    python tools/ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [top]

ncu's CSV source page is SASS-only, so the line table comes from `nvdisasm -gi` on the cubin extracted from the
SAME library build that was profiled (compiled with -lineinfo); the join is by instruction offset and is checked
by comparing opcodes.  Samples are reported per outermost line in the kernel's own file (who called the inlined
code) and per innermost line (where the instruction lives)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    recs = None
    for f in sorted(os.listdir(tmp)):
        sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        lines = sass.split("\n")
        starts = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")]
        if not starts:
            continue
        recs, chain, pending = {}, [], []
        for l in lines[starts[0] + 1:]:
            if l.startswith("//-----"):
                break
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
            if m:
                pending.append((os.path.basename(m.group(1)), int(m.group(2))))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", l)
            if m:
                if pending:
                    chain, pending = pending, []
                recs[int(m.group(1), 16)] = (m.group(2).strip(), chain)
        break
    if recs is None:
        sys.exit("kernel not found in " + lib)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + os.environ.get("NCU_KERNEL", kern)], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    ix = {h: k for k, h in enumerate(hdr)}
    body = rows[hi + 1:]
    base = int(body[0][0], 16)
    outer_s, outer_i, inner_s, inner_i = (collections.Counter() for _ in range(4))
    tot_s = tot_i = miss = bad = 0
    kfile = None
    for r in body:
        if len(r) < len(hdr):
            continue
        a = int(r[0], 16) - base
        s, n = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        tot_s += s
        tot_i += n
        if a not in recs:
            miss += 1
            continue
        op, ch = recs[a]
        if op.split()[-1 if False else 0].lstrip("@!P0123456789U ") [:4] != r[1].strip().split()[0].lstrip("@!P0123456789U ")[:4]:
            bad += 1
        if kfile is None and ch:
            kfile = ch[-1][0]
        own = [c for c in ch if c[0] == kfile]
        o = own[-1] if own else (ch[-1] if ch else ("?", 0))
        outer_s[o] += s
        outer_i[o] += n
        LEVEL2 = os.environ.get("NCU_LEVEL2")
        if LEVEL2 and o == (kfile, int(LEVEL2)) and len(ch) >= 2:
            outer_s[("  under " + LEVEL2 + ": " + ch[-2][0], ch[-2][1])] += s
            outer_i[("  under " + LEVEL2 + ": " + ch[-2][0], ch[-2][1])] += n
        inn = ch[0] if ch else ("?", 0)
        inner_s[inn] += s
        inner_i[inn] += n
    print(f"samples {tot_s}  warp-instructions {tot_i}  unmatched {miss}  opcode-mismatch {bad}  of {len(body)}")
    for title, cs, ci in (("outermost line in " + str(kfile), outer_s, outer_i), ("innermost line", inner_s, inner_i)):
        print(f"--- by {title}")
        for (f, ln), s in cs.most_common(top):
            print(f"{100 * s / max(1, tot_s):5.1f}% samples {100 * ci[(f, ln)] / max(1, tot_i):5.1f}% inst   {f}:{ln}")


if __name__ == "__main__":
    main()
