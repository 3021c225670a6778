#!/usr/bin/env python
"""Executed warp-instructions and stall samples of one kernel per source line, from `ncu --page source --csv` joined by
instruction offset with `nvdisasm -gi` of the cubin of the SAME library build.

    python tools/ncu_lines2.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [top]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile


def main():
    rep, lib, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # several kernels may be in the report: blocks start with a "Kernel Name" row
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    blk = [b for b in blocks if kern in b["name"]][0]
    h, data = blk["rows"][0], blk["rows"][1:]
    ia, iex, ismp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
    base = int(data[0][ia], 16)
    ex = {int(r[ia], 16) - base: (int(r[iex]), int(r[ismp] or 0)) for r in data}
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    mangled = re.sub(r"[^A-Za-z0-9_]", "", kern)
    inner, outer, fn = collections.Counter(), collections.Counter(), collections.Counter()
    sin, sout = collections.Counter(), collections.Counter()
    found = False
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
        # template instantiations share the substring: pick the one the report names (k_map_easy<(bool)1> -> ...ILb1...)
        inst = re.search(r"<\(bool\)([01])>", blk["name"])
        want = os.environ.get("NCU_SASS_NAME", mangled + ("ILb" + inst.group(1) if inst else ""))
        starts = [i for i, l in enumerate(sass) if l.startswith(".text.") and want in l and l.rstrip().endswith(":")]
        if not starts:
            continue
        found = True
        chain, fresh = [], True
        for l in sass[starts[0] + 1:]:
            if l.startswith(".text.") or l.startswith("//-----"):
                break
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
            if m:
                if fresh:
                    chain, fresh = [], False
                chain.append(os.path.basename(m.group(1)) + ":" + m.group(2))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
            if m and chain:
                fresh = True
                off = int(m.group(1), 16)
                e, s = ex.get(off, (0, 0))
                inner[chain[0]] += e; sin[chain[0]] += s
                outer[chain[-1]] += e; sout[chain[-1]] += s
                if len(chain) >= 2:
                    fn[chain[-2]] += e
        break
    assert found, "kernel not found in the library's cubins (set NCU_SASS_NAME to a substring of the mangled name)"
    tot, stot = sum(inner.values()), sum(sin.values())
    print(f"kernel {blk['name'][:60]}: {tot} warp-instructions, {stot} samples")
    for title, c, sc in (("outermost line", outer, sout), ("innermost line", inner, sin)):
        print(f"--- by {title}")
        for k, v in c.most_common(top):
            print(f"  {100 * v / tot:5.1f}% inst  {100 * sc[k] / max(stot, 1):5.1f}% samples  {k}")
    print("--- by call site one level below the kernel line")
    for k, v in fn.most_common(top):
        print(f"  {100 * v / tot:5.1f}% inst  {k}")


main()
