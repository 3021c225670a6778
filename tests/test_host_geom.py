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
    lib.h_pack_record.argtypes = [C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 4 + [C.c_int] * 4 + [C.c_double] * 4 + \
        [C.POINTER(C.c_int), C.POINTER(C.c_float)]
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


@pytest.mark.parametrize("kind", ["regular_1d", "regular_pole_rows", "curvilinear_2d"])
def test_apply_record_equals_interp_bl(hg, orc, kind):
    """pack.cuh builds the 32-byte record the apply kernels read (four source offsets + four REAL(4) weights, or "copy the
    nearest point", or one of INTERP_BL's error constants).  The value the record yields -- the kernels' statement
    (((z1*w1 + z2*w2) + z3*w3) + z4*w4) / wup in REAL(4) -- against the apply loop of BILIN_2D transliterated in tests/pyref.py
    (index clamp, identical-point copy, INTERP_BL with its periodic wrap and error codes), on random and edge mapping entries."""
    from tests import pyref as P
    from sosie_b200 import grids as G
    rng = np.random.default_rng(31)
    f4 = np.float32
    if kind == "curvilinear_2d":
        Xs, Ys = G.orca_like(40, 30, "F", 15.0)
        sep, lon1, lat1 = 0, None, None
        X1w, Y1w = Xs, Ys
        Z1w = rng.normal(5.0, 3.0, Xs.shape).astype(f4)
    else:
        lon1, lat1 = G.regular_latlon(36, 24)
        sep = 1
        X1w, Y1w = G.expand_2d(lon1, lat1)
        Z1w = rng.normal(5.0, 3.0, X1w.shape).astype(f4)
        if kind == "regular_pole_rows":       # the working arrays BILIN_2D applies on: two rows beyond the last latitude
            X1w, Y1w, Z1w = orc.ext_north_to_90_regg(X1w, Y1w, Z1w)
            lat1 = np.ascontiguousarray(Y1w[:, 0])
    nyw, nx = X1w.shape
    dp = lambda a: None if a is None else np.ascontiguousarray(a, np.float64).ctypes.data_as(C.POINTER(C.c_double))
    keep = [np.ascontiguousarray(a, np.float64) for a in (lon1 if sep else X1w[0], lat1 if sep else Y1w[:, 0], X1w, Y1w)]
    plon1, plat1, pX, pY = (a.ctypes.data_as(C.POINTER(C.c_double)) for a in keep)
    id4 = (C.c_int * 4)(); w4 = (C.c_float * 4)()
    Zflat = Z1w.reshape(-1)
    n, ncopy, nconst = 60000, 0, 0
    for k in range(n):
        kew = int(rng.choice([-1, 0, 2]))
        iP = int(rng.choice([-1, 1, 2, nx - 1, nx, int(rng.integers(1, nx + 1))]))
        jP = int(rng.choice([-1, 1, 2, nyw - 1, nyw, int(rng.integers(1, nyw + 1))]))
        iq = int(rng.integers(1, 5))
        xa, xb = (float(v) for v in rng.choice([0.0, 1.0, 0.5, rng.random(), rng.random(), 1.0e-12], 2))
        iPc, jPc = max(iP, 1), max(jP, 1)
        lon_t, lat_t = float(rng.uniform(0, 360)), float(rng.uniform(-89, 89))
        r = rng.random()
        if r < 0.15:          # the target ON the nearest point, possibly 360 degrees apart or within / beyond 1e-5
            lon_t = float(X1w[jPc - 1, iPc - 1]) + float(rng.choice([0.0, 360.0, -360.0, 5.0e-6, -5.0e-6, 2.0e-5]))
            lat_t = float(Y1w[jPc - 1, iPc - 1]) + float(rng.choice([0.0, 5.0e-6, -5.0e-6, 2.0e-5]))
        hg.h_pack_record(nx, nyw, sep, plon1, plat1, pX, pY, kew, iP, jP, iq, xa, xb, lon_t, lat_t, id4, w4)
        w = [f4(v) for v in w4]
        if id4[0] >= 0:
            z = [Zflat[id4[q]] for q in range(4)]
            wup = f4(f4(f4(w[0] + w[1]) + w[2]) + w[3])
            val = f4(f4(f4(f4(z[0] * w[0]) + f4(z[1] * w[1])) + f4(z[2] * w[2])) + f4(z[3] * w[3])) / wup
        elif id4[0] == -1:
            val = Zflat[id4[1]]; ncopy += 1
        else:
            val = w[0]; nconst += 1
        # BILIN_2D's loop body for this target (src/mod_bilin_2d.f90:217-234)
        eps5 = float(f4(1.0e-5))
        if abs(P.degE_to_degWE(float(X1w[jPc - 1, iPc - 1])) - P.degE_to_degWE(lon_t)) < eps5 and abs(float(Y1w[jPc - 1, iPc - 1]) - lat_t) < eps5:
            ref = Z1w[jPc - 1, iPc - 1]
        else:
            with np.errstate(all="ignore"):
                ref = P.interp_bl_np(kew, iPc, jPc, iq, xa, xb, Z1w)
        assert f4(val) == f4(ref) or (val != val and ref != ref), (k, kew, iP, jP, iq, xa, xb, lon_t, lat_t, list(id4), list(w4), val, ref)
    assert ncopy > 0.05 * n and nconst > 0.05 * n
