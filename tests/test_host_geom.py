"""The device geometry functions of sosie_b200/csrc/geom.cuh compiled for the HOST by g++ (same source, same IEEE flags,
pmath libm) and exercised on millions of random inputs -- CPU-side evidence for the exact shortcuts of the EASY mapping
kernel: `d_l_inpoly_rect` against the general `d_l_inpoly`, and the sign rule for the first-guess
quadrant against the literal five-heading decision.  Test infrastructure; nothing here is part of the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def hg(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    out = str(tmp_path_factory.mktemp("hg") / "libhostgeom.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-w",
                           "-I", CUDA_INC, "-I", os.path.join(ROOT, "sosie_b200", "csrc"),
                           os.path.join(HERE, "host_geom", "harness.cpp"), "-o", out])
    lib = C.CDLL(out)
    lib.h_l_inpoly.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double]
    lib.h_l_inpoly_rect.argtypes = [C.c_double] * 4 + [C.c_int] + [C.c_double] * 2
    lib.h_heading.argtypes = [C.c_double] * 4
    lib.h_heading.restype = C.c_double
    lib.h_degE_to_degWE.argtypes = [C.c_double]
    lib.h_degE_to_degWE.restype = C.c_double
    lib.h_local_coord.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.c_int)]
    return lib


def _coord_pool(rng, n, lon=True):
    """coordinates of the kinds a regular source holds: fine / coarse spacing, integer and half-integer values (the +0.001
    nudge of L_InPoly), values near the 0/360 seam, exactly 360 (a longitude of 0 after WHERE(loni <= 0) + 360)"""
    kind = rng.integers(0, 6, n)
    base = rng.uniform(0.0, 360.0 if lon else 180.0, n) - (0.0 if lon else 90.0)
    v = np.where(kind == 0, np.round(base * 60.0) / 60.0, base)                 # 1/60 degree grid (every 60th node an integer)
    v = np.where(kind == 1, np.round(base), v)                                 # integers
    v = np.where(kind == 2, np.round(base * 2.0) / 2.0, v)                     # half-integers
    v = np.where(kind == 3, np.round(base * 7.0) / 7.0, v)
    if lon:
        v = np.where(kind == 4, np.where(rng.random(n) < 0.5, rng.uniform(0.0, 0.05, n), rng.uniform(359.95, 360.0, n)), v)
        v = np.where((kind == 5) & (rng.random(n) < 0.2), 360.0, v)
        v = np.where(v <= 0.0, v + 360.0, v)
    return v


def test_inpoly_rect_equals_general_inpoly(hg):
    rng = np.random.default_rng(2026)
    n = 1000000
    xa = _coord_pool(rng, n); ya = _coord_pool(rng, n, lon=False)
    step = rng.choice([1.0 / 60.0, 0.25, 1.0, 1.0e-6, 5.0], n) * rng.choice([-1.0, 1.0], n)
    xb = xa + step
    xb = np.where(xb <= 0.0, xb + 360.0, np.where(xb > 360.0, xb - 360.0, xb))   # the neighbour across the seam
    xb = np.where(rng.random(n) < 0.02, xa, xb)                                 # degenerate: equal abscissae
    yb = ya + rng.choice([1.0 / 60.0, 0.25, 1.0, 1.0e-6, 5.0], n) * rng.choice([-1.0, 1.0], n)
    yb = np.where(rng.random(n) < 0.02, ya, yb)
    pat = rng.integers(0, 2, n)
    t = rng.uniform(-0.3, 1.3, n); u = rng.uniform(-0.3, 1.3, n)
    dx = xb - xa
    dx = np.where(dx > 180.0, dx - 360.0, np.where(dx < -180.0, dx + 360.0, dx))
    px = xa + t * dx
    py = ya + u * (yb - ya)
    snap = rng.integers(0, 8, n)                                                # points ON edges and corners
    px = np.where(snap == 0, xa, np.where(snap == 1, xb, px))
    py = np.where(snap == 2, ya, np.where(snap == 3, yb, py))
    px = np.where(snap == 4, xa, px); py = np.where(snap == 4, ya, py)
    px = np.mod(px, 360.0)                                                      # the un-moved xP of Q8: in [0, 360)
    px = np.where(rng.random(n) < 0.01, 0.0, px)
    neg = rng.random(n) < 0.002                                                 # the STOP of mod_poly.f90:79
    px = np.where(neg, -px, px)
    vlon = (C.c_double * 4)(); vlat = (C.c_double * 4)()
    nin = nstop = 0
    for k in range(n):
        a, b, c, d = float(xa[k]), float(xb[k]), float(ya[k]), float(yb[k])
        if pat[k]:
            vlon[:] = (a, b, b, a); vlat[:] = (c, c, d, d)
        else:
            vlon[:] = (a, a, b, b); vlat[:] = (c, d, d, c)
        g = hg.h_l_inpoly(vlon, vlat, float(px[k]), float(py[k]))
        r = hg.h_l_inpoly_rect(a, b, c, d, int(pat[k]), float(px[k]), float(py[k]))
        assert g == r, (k, a, b, c, d, int(pat[k]), float(px[k]), float(py[k]), g, r)
        nin += g == 1; nstop += g == -1
    assert nin > 0.2 * n and nstop > 0


def test_first_guess_quadrant_sign_rule(hg):
    """with neighbour headings exactly 0 / 90 / 180 / 270 the quadrant tests of src/mod_bilin_2d.f90:531-540
    only ask in which quarter ATAN2 falls; inside the guard of the kernel the literal decision equals the sign rule."""
    rng = np.random.default_rng(7)
    n = 600000
    conv = float.fromhex("0x1.921fb54442d18p+1") / 180.0
    twopi = 2.0 * float.fromhex("0x1.921fb54442d18p+1")
    lonP = _coord_pool(rng, n)
    lonP = np.where(lonP >= 360.0, 0.0, lonP)
    latP = np.clip(_coord_pool(rng, n, lon=False), -89.0, 89.0)
    mag = rng.choice([1.0e-7, 1.0000001e-7, 1.1459156e-7, 1.146e-7, 3.0e-7, 1.0e-5, 1.0e-3, 0.01, 0.3, 5.0, 44.9], (2, n))
    sgn = rng.choice([-1.0, 1.0], (2, n))
    xP = np.mod(lonP + sgn[0] * mag[0], 360.0)
    yP = latP + sgn[1] * mag[1]
    nfast = 0
    quads = set()
    for k in range(n):
        lp, la, x, y = float(lonP[k]), float(latP[k]), float(xP[k]), float(yP[k])
        dxr = x * conv - lp * conv                       # the kernel's guard, statement for statement
        if abs(dxr) >= twopi:
            continue
        if dxr >= twopi / 2:
            dxr -= twopi
        if dxr <= -twopi / 2:
            dxr += twopi
        dphi = y - la
        if not (abs(dxr) >= 2.0e-9 and abs(dxr) <= 0.8 and abs(dphi) >= 1.0e-7 and abs(dphi) <= 45.0 and abs(y) <= 89.0 and abs(la) <= 89.0):
            continue
        fast = (1 if dphi > 0.0 else 2) if dxr > 0.0 else (4 if dphi > 0.0 else 3)
        hP = hg.h_heading(lp, x, la, y)
        hN, hE, hS, hW = 0.0, 90.0, 180.0, 270.0
        iq = 4
        hPp = hg.h_degE_to_degWE(hP)
        if hPp > hN and hPp <= hE: iq = 1
        if hP > hE and hP <= hS: iq = 2
        if hP > hS and hP <= hW: iq = 3
        if hP > hW and hPp <= hN: iq = 4
        assert iq == fast, (lp, la, x, y, hP, iq, fast)
        nfast += 1
        quads.add(iq)
    assert nfast > 0.5 * n and quads == {1, 2, 3, 4}
    # the table values the kernel checks for: due east / south / west / north of a point are exactly 90 / 180 / 270 / 0
    for la in (-60.0, 0.0, 33.3, 88.0):
        assert hg.h_heading(10.0, 10.25, la, la) == 90.0 and hg.h_heading(10.0, 9.75, la, la) == 270.0
        assert hg.h_heading(10.0, 10.0, la, la + 0.25) == 0.0 and hg.h_heading(10.0, 10.0, la, la - 0.25) == 180.0


def test_local_coord_without_sqrt_equals_literal(hg, orc):
    """the kernel's LOCAL_COORD decides "zres > zresmax" on the squared step (no SQRT) and skips the
    divides whose numerator is zero; the oracle's LOCAL_COORD is the literal statement sequence (src/mod_bilin_2d.f90:
    697-799).  Same (alpha, beta, ipb), bit for bit, on random cells: rectangles of a regular grid, skewed quads, cells
    across the 0/360 seam (the -180..180 re-framing), collapsed cells (ipb = 12) and targets far outside (ipb = 11 / big steps)."""
    rng = np.random.default_rng(99)
    n = 200000
    xl = (C.c_double * 5)(); xp = (C.c_double * 5)()
    a, b, ip = C.c_double(), C.c_double(), C.c_int()
    seen = set()
    for k in range(n):
        kind = k % 5
        lon0 = rng.uniform(0.0, 360.0) if kind != 2 else rng.choice([359.99, 0.004, 359.5])
        lat0 = rng.uniform(-85.0, 85.0)
        dx, dy = rng.choice([1.0 / 60.0, 0.25, 1.0, 2.5]), rng.choice([1.0 / 60.0, 0.25, 1.0, 2.5])
        sx, sy = rng.choice([-1.0, 1.0]), rng.choice([-1.0, 1.0])
        corners = [(0, 0), (sx * dx, 0), (sx * dx, sy * dy), (0, sy * dy)]
        if kind == 1:      # skewed
            corners = [(cx + rng.normal(0, 0.15) * dx, cy + rng.normal(0, 0.15) * dy) for cx, cy in corners]
        if kind == 3:      # all four latitudes equal
            corners = [(cx, 0.0) for cx, cy in corners]
        t, u = rng.uniform(-0.2, 1.2, 2) if kind != 4 else rng.uniform(-40.0, 40.0, 2)
        px = lon0 + t * sx * dx; py = lat0 + u * sy * dy
        lons = [px] + [lon0 + cx for cx, _ in corners]
        lats = [py] + [lat0 + cy for _, cy in corners]
        lons = [v % 360.0 for v in lons]
        lons = [v + 360.0 if v <= 0.0 else v for v in lons]
        xl[:] = lons; xp[:] = lats
        hg.h_local_coord(xl, xp, C.byref(a), C.byref(b), C.byref(ip))
        ao, bo, ipo = orc.local_coord(np.array(lons), np.array(lats))
        same = (a.value == ao or (a.value != a.value and ao != ao)) and (b.value == bo or (b.value != b.value and bo != bo))
        assert same and ip.value == ipo, (k, lons, lats, a.value, ao, b.value, bo, ip.value, ipo)
        seen.add(ipo)
    assert {0, 12} <= seen
