"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol include/sosie_b200.h declares,
and refuses to work (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

import sosie_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sosie_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sosie_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(sosie_b200.lib_path()):
        from sosie_b200 import build
        build.build()
    return ctypes.CDLL(sosie_b200.lib_path())


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sosie_b200.h but not exported"
    assert sorted(sosie_b200.EXPORTS) == names


def test_version_string(lib):
    lib.sosie_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.sosie_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the refusal path cannot be exercised")
    with pytest.raises(sosie_b200.SosieError):
        sosie_b200.Sosie(0)


def test_product_does_not_touch_the_oracle():
    """the product package must not import / link the test oracle"""
    for dp, _, files in os.walk(os.path.join(ROOT, "sosie_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".f90")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "from oracle" not in txt and "import oracle" not in txt and "orc_" not in txt, os.path.join(dp, f)
    out = os.popen(f"ldd {sosie_b200.lib_path()}").read()
    assert "oracle" not in out


def test_sass_is_sm100a():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    out = subprocess.run([cuobjdump, "-lelf", sosie_b200.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_mapping_binary_twin_roundtrip(tmp_path):
    """the raw-binary twin of the mapping file (csrc/mapfile.cu) -- host-side I/O, no GPU needed.
    A mapping built by the oracle (as stock CPU SOSIE would) survives the round trip bit for bit; the header is checked."""
    import numpy as np
    import sosie_b200
    from oracle import oracle as O
    from sosie_b200 import grids as G
    O.build(); O.set_libm(O.PMATH)
    cfg = G.config("E", 0.01)
    Xt, Yt = G.trg_2d(cfg)
    Xs, Ys = G.src_2d(cfg)
    Z1 = G.field_slabs(Xs, Ys, 1)
    _, mo = O.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, Xt, Yt, l_reg_src=True)
    f = str(tmp_path / "sosie_mapping_test.bin")
    sosie_b200.mapping_write_bin(f, Xt, Yt, mo)
    ny, nx = Xt.shape
    import os
    assert os.path.getsize(f) == 64 + nx * ny * (8 + 8 + 12 + 16 + 4 + 4)
    back = sosie_b200.mapping_read_bin(f)
    assert np.array_equal(back["lon"], Xt) and np.array_equal(back["lat"], Yt)
    for k in ("metrics", "alphabeta", "ipb", "dist_np"):
        assert np.array_equal(back[k], mo[k]), k
    with open(f, "r+b") as fh:
        fh.write(b"NOTAMAPF")
    import pytest
    with pytest.raises(sosie_b200.SosieError):
        sosie_b200.mapping_read_bin(f)


def test_header_is_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: include/sosie_b200.h compiles as C99 (what cgo / a Fortran interoperability checker /
    any FFI generator would feed on) and as C++11, with warnings as errors"""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "sosie_b200.h"\nint main(void) { return 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)])
