"""TEST INFRASTRUCTURE -- inputs that come out bit-identical on every machine.

numpy's float64 sin/cos/arcsin/arctan2 pick SIMD code paths by CPU (AVX-512 SVML vs scalar libm) and may differ in the
last bit between this container and the GPU box, so digests of full-size results (tests/golden/digest_*.json, made by
tools/make_digests.py) cannot be anchored on sosie_b200.grids.  Here every transcendental goes through pmath (the portable
libm of sosie_b200/csrc/pmath.h, evaluated by the oracle library: IEEE + - * / sqrt in a fixed order, built with
-ffp-contract=off) and everything else is a single IEEE operation per element, so the arrays hash the same everywhere.
"""
import hashlib

import numpy as np

from oracle import oracle as O

D2R = np.pi / 180.0


def _pm(fn, x, y=None):
    return O.pmath_eval(fn, np.ascontiguousarray(x, np.float64), None if y is None else np.ascontiguousarray(y, np.float64))[0]


def psin(x):
    return _pm(0, x).reshape(np.shape(x))


def pcos(x):
    return _pm(1, x).reshape(np.shape(x))


def patan2(y, x):
    return _pm(5, x, y).reshape(np.shape(x))


def enatl_like(nx, ny, lonr0=-97.0, latr0=4.0, dres=1.0 / 60.0, rot_deg=4.0, axis_lon=-30.0):
    """eNATL60-shaped regional curvilinear target (same construction and parameters as grids.rotated_regional: a regular grid in
    a frame rotated by rot_deg about the equatorial axis at axis_lon), portable bits.  Returns lon, lat (ny, nx) float64."""
    lonr = (lonr0 + np.arange(nx, dtype=np.float64) * dres) * D2R
    latr = (latr0 + np.arange(ny, dtype=np.float64) * dres) * D2R
    a = lonr - axis_lon * D2R
    cl, sl = pcos(a), psin(a)
    cp, sp = pcos(latr), psin(latr)
    ca, sa = float(pcos(np.array([rot_deg * D2R]))[0]), float(psin(np.array([rot_deg * D2R]))[0])
    lon = np.empty((ny, nx), np.float64)
    lat = np.empty((ny, nx), np.float64)
    for j0 in range(0, ny, 256):
        j1 = min(ny, j0 + 256)
        x = cp[j0:j1, None] * cl[None, :]
        y = cp[j0:j1, None] * sl[None, :]
        z = np.ascontiguousarray(np.broadcast_to(sp[j0:j1, None], x.shape))
        y2 = y * ca - z * sa
        z2 = y * sa + z * ca
        lat[j0:j1] = patan2(z2, np.sqrt(x * x + y2 * y2)) / D2R
        lon[j0:j1] = np.mod(patan2(y2, x) / D2R + axis_lon, 360.0)
    return lon, lat


def polar_stereo(nx, ny, lat_edge_x=35.0, lon0=315.0):
    """config D target (grids.polar_stereo) with portable bits: atan2 through pmath, rho = sqrt(x*x + y*y)."""
    pole_i, pole_j = 0.5 * nx + 0.3, 0.5 * ny + 0.7
    t = np.array([np.pi / 4 - lat_edge_x * D2R / 2])
    L = float(psin(t)[0] / pcos(t)[0])
    dx = 2.0 * L / nx
    x = (np.arange(1, nx + 1, dtype=np.float64) - pole_i) * dx
    y = (np.arange(1, ny + 1, dtype=np.float64) - pole_j) * dx
    X, Y = np.meshgrid(x, y)
    rho = np.sqrt(X * X + Y * Y)
    lat = 90.0 - 2.0 * patan2(rho, np.ones_like(rho)) / D2R
    lon = np.mod(patan2(X, -Y) / D2R + lon0, 360.0)
    return np.ascontiguousarray(lon), np.ascontiguousarray(lat)


def integer_field(nx, ny, nslab=1, lo=-6000, span=10007):
    """(nslab, ny, nx) float32 slabs made of integer arithmetic only (exact in REAL(4)): a bathymetry-like saw-tooth"""
    i = np.arange(nx, dtype=np.int64)[None, :]
    j = np.arange(ny, dtype=np.int64)[:, None]
    out = np.empty((nslab, ny, nx), np.float32)
    for s in range(nslab):
        out[s] = (((i * 7 + j * 13 + s * 101) * 37) % span + lo).astype(np.float32)
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def block_digests(arrs, j0, j1):
    """arrs: dict name -> array whose LAST two axes are (rows, nx); digest of rows j0:j1 of each"""
    return {k: sha(v[..., j0:j1, :]) for k, v in arrs.items()}
