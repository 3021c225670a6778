#!/usr/bin/env python
"""Experiment builds: recompile ONE csrc file with extra -D flags and link it with the other objects of the regular build into
variants/libsosie_NAME.so (git-ignored, travels to the GPU box); select it with SOSIE_B200_LIB=variants/libsosie_NAME.so.

    python tools/build_variant.py NAME FILE.cu -DMACRO=VALUE [...]
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sosie_b200 import build as B

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
out = os.path.join(ROOT, "variants"); os.makedirs(out, exist_ok=True)
obj = os.path.join(out, f"{name}_{src[:-3]}.o")
subprocess.check_call(["nvcc"] + B.NVCC_FLAGS + flags + ["-ccbin", "/usr/bin/g++", "-c", os.path.join(B.CSRC, src), "-o", obj],
                      stdout=open(os.path.join(out, f"{name}.ptxas.log"), "w"), stderr=subprocess.STDOUT)
objs = [obj if os.path.basename(s)[:-3] == src[:-3] else os.path.join(B.OUT_DIR, os.path.basename(s)[:-3] + ".o") for s in B.sources()]
lib = os.path.join(out, f"libsosie_{name}.so")
subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart"])
print(lib)
