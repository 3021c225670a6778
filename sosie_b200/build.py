"""Build libsosie_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m sosie_b200.build [--force]

-fmad=false / -ffp-contract=off: the reference is compiled by gfortran -O3 for baseline x86-64
(no FMA, no reassociation); parity-critical arithmetic must round the same way.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libsosie_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def source_digest(names):
    """sha256 over the named csrc files + the nvcc flags: profiles/*.json record it at capture time, bench.py compares it with
    the tree it runs from and refuses a `roofline.traffic` taken from an older build of the kernel"""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for n in sorted(names):
        with open(os.path.join(CSRC, n), "rb") as f:
            h.update(n.encode()); h.update(f.read())
    return h.hexdigest()[:16]


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in sources() + headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libsosie_b200.so")
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    env = dict(os.environ)
    for src in sources():
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-ccbin", "/usr/bin/g++", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)))
    log = []
    for src, p in procs:
        out = p.communicate()[0].decode()
        log.append(f"==== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(OUT_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart"]
    subprocess.check_call(cmd)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
