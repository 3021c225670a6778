#!/usr/bin/env python
"""Randomised parity check: the CUDA path through the C ABI against the CPU oracle on many small random cases
(grids, periodicity, masks, sizes, options).  Bit-exact results are required; where the oracle reports that the
reference would STOP, the library must raise too.

    python tools/fuzz_parity.py [ncases] [seed]

Test infrastructure (it loads oracle/): not part of the product."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import sosie_b200
from sosie_b200 import grids as G
from oracle import oracle as orc


def rnd_regular_axes(rng):
    nx, ny = int(rng.integers(8, 90)), int(rng.integers(7, 70))
    kind = rng.integers(0, 4)
    if kind == 0:      # global cell-centred (pole extension)
        lon = (np.arange(nx) + 0.5) * 360.0 / nx
        lat = -90.0 + (np.arange(ny) + 0.5) * 180.0 / ny
        kew = int(rng.choice([0, -1]))
    elif kind == 1:    # global with overlap columns
        ov = 2
        d = 360.0 / (nx - ov)
        lon = np.mod(rng.uniform(0, 40) + np.arange(nx) * d, 360.0)
        lon = np.sort(lon) if rng.random() < 0.5 else lon
        lat = np.linspace(-80.0, 85.0, ny)
        kew = ov
    elif kind == 2:    # regional, positive longitudes
        l0 = rng.uniform(5.0, 200.0)
        lon = l0 + np.arange(nx) * rng.uniform(0.2, 1.5)
        lat = rng.uniform(-60.0, 0.0) + np.arange(ny) * rng.uniform(0.2, 1.2)
        lat = lat[lat < 89.5]
        kew = -1
    else:              # node-centred global reaching the pole
        lon = np.arange(nx) * 360.0 / nx
        lat = np.linspace(-90.0, 90.0, ny)
        kew = int(rng.choice([0, -1]))
    return lon, lat, kew


def rnd_target(rng, lon_s, lat_s):
    nx, ny = int(rng.integers(3, 60)), int(rng.integers(3, 50))
    kind = rng.integers(0, 4)
    lo0, lo1 = float(np.min(lon_s)), float(np.max(lon_s))
    la0, la1 = float(np.min(lat_s)), float(np.max(lat_s))
    cx, cy = rng.uniform(lo0, lo1), rng.uniform(min(max(la0, -85.0), 84.0), max(min(la1, 85.0), min(max(la0, -85.0), 84.0) + 0.5))
    w, h = rng.uniform(2.0, max(2.5, 0.6 * (lo1 - lo0 + 1.0))), rng.uniform(2.0, max(2.5, 0.6 * (la1 - la0 + 1.0)))
    if kind == 0:      # regular 1-D axes
        return np.mod(np.linspace(cx - w / 2, cx + w / 2, nx), 360.0), np.clip(np.linspace(cy - h / 2, cy + h / 2, ny), -89.9, 89.9)
    u = np.linspace(-0.5, 0.5, nx)[None, :].repeat(ny, 0)
    v = np.linspace(-0.5, 0.5, ny)[:, None].repeat(nx, 1)
    if kind == 1:      # rotated + sheared
        a = rng.uniform(-0.6, 0.6)
        X = cx + w * (u * np.cos(a) - v * np.sin(a)) + rng.uniform(-0.2, 0.2) * w * u * v
        Y = cy + h * (u * np.sin(a) + v * np.cos(a))
    elif kind == 2:    # polar-cap like: rows are circles around the pole
        r = 3.0 + 25.0 * (v + 0.5)
        th = 2 * np.pi * (u + 0.5)
        X = np.degrees(th) + rng.uniform(0, 90)
        Y = (90.0 - r) * (1 if rng.random() < 0.5 else -1)
    else:              # regular but given as 2-D arrays
        X = cx + w * u
        Y = cy + h * v
    return np.mod(X, 360.0), np.clip(Y, -89.95, 89.95)


ORACLE_ONLY = bool(os.environ.get("FUZZ_ORACLE_ONLY"))
ONLY = {int(x) for x in os.environ.get("FUZZ_ONLY", "").split(",") if x}


class _NoGpu:
    """FUZZ_ORACLE_ONLY=1: exercise the case generator and the oracle alone (no GPU needed)"""
    def __getattr__(self, name):
        return lambda *a, **k: None


def both(tag, fg, fo):
    """run the GPU call and the oracle call; either both fail or both succeed"""
    if ORACLE_ONLY:
        print("  ", tag, flush=True)
        try:
            fo()
        except RuntimeError as e:
            print("   oracle STOP:", e)
        return None, None
    eg = eo = None
    try:
        rg = fg()
    except sosie_b200.SosieError as e:
        eg = e
    try:
        ro = fo()
    except RuntimeError as e:
        eo = e
    if (eg is None) != (eo is None):
        raise AssertionError(f"{tag}: error behaviour differs: gpu={eg!r} oracle={eo!r}")
    return (None, None) if eg is not None else (rg, ro)


def eq(tag, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        same = np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    else:
        same = np.array_equal(a, b)
    if not same:
        if os.environ.get("FUZZ_VERBOSE"):
            idx = np.argwhere(a != b)
            for ix in idx[:12]:
                print("    at", tuple(ix), "gpu", a[tuple(ix)], "oracle", b[tuple(ix)], flush=True)
        raise AssertionError(f"{tag}: {(a != b).sum()} of {a.size} differ")


def case_bilin(gpu, rng, k):
    curvi = rng.random() < 0.3
    if curvi:
        Xs, Ys = G.orca_like(int(rng.integers(24, 100)), int(rng.integers(24, 80)), str(rng.choice(["F", "T"])), float(rng.uniform(0.0, 300.0)))
        kew, lreg = 2, False
        lon_s, lat_s = Xs, Ys
        X2, Y2 = Xs, Ys
    else:
        lon_s, lat_s, kew = rnd_regular_axes(rng)
        lreg = True
        X2, Y2 = G.expand_2d(lon_s, lat_s)
        if rng.random() < 0.25:
            lon_s, lat_s, lreg = X2, Y2, bool(rng.random() < 0.5)   # 2-D arrays of a regular grid
    Xt, Yt = rnd_target(rng, X2, Y2)
    iorca = 0
    if rng.random() < 0.2:
        piv = str(rng.choice(["F", "T"]))
        Xt, Yt = G.orca_like(int(rng.integers(24, 70)), int(rng.integers(24, 60)), piv, float(rng.uniform(0.0, 300.0)))
        iorca = int(rng.choice([0, 6 if piv == "F" else 4]))
    nslab = int(rng.integers(1, 4))
    Z1 = G.field_slabs(X2, Y2, nslab, seed=int(rng.integers(1 << 30)))
    mask = None
    if rng.random() < 0.4 and Xt.ndim == 2:
        mask = (rng.random(Xt.shape) > 0.2).astype(np.int32)
    first = rng.random() < 0.5 and mask is None
    gpu.set_pipeline_min_targets(0 if rng.random() < 0.5 else 1 << 30)
    tag = f"bilin#{k} curvi={curvi} src={np.shape(lon_s)} trg={np.shape(Xt)} kew={kew} lreg={lreg} mask={mask is not None} first={first} iorca={iorca}"

    gpu.l_first_call_interp_routine = True

    def fg():
        if first:
            return gpu.bilin_2d_first(kew, lon_s, lat_s, Z1, Xt, Yt, l_reg_src=lreg, i_orca_trg=iorca)
        Z2 = gpu.bilin_2d(kew, lon_s, lat_s, Z1, Xt, Yt, mask_domain_trg=mask, l_reg_src=lreg, i_orca_trg=iorca)
        return Z2, gpu.bilin_get_map()
    with np.errstate(all="ignore"):
        rg, ro = both(tag, fg, lambda: orc.bilin_2d(kew, lon_s, lat_s, Z1, Xt, Yt, mask_domain_trg=mask, l_reg_src=lreg, i_orca_trg=iorca))
    if rg is None:
        return tag + " (both STOP)"
    if os.environ.get("FUZZ_VERBOSE"):
        np.set_printoptions(precision=17, linewidth=200)
        print(tag, "\n  lon_s", np.asarray(lon_s).ravel()[:100], "\n  lat_s", np.asarray(lat_s).ravel()[:100] if np.ndim(lat_s) == 1 else np.asarray(lat_s)[:, 0],
              "\n  Xt", np.asarray(Xt).ravel()[:60], "\n  Yt", np.asarray(Yt).ravel()[:60], "\n  info gpu", rg[1].get("info"), "oracle", ro[1]["info"], flush=True)
        try:
            eq(tag + " metrics", rg[1]["metrics"], ro[1]["metrics"])
        except AssertionError:
            idx = np.argwhere(np.any(rg[1]["metrics"] != ro[1]["metrics"], axis=0))
            X2t, Y2t = (G.expand_2d(Xt, Yt) if np.ndim(Xt) == 1 else (Xt, Yt))
            for j, i in idx[:10]:
                print("    target", (j, i), "lon/lat", X2t[j, i], Y2t[j, i], "gpu", rg[1]["metrics"][:, j, i], rg[1]["alphabeta"][:, j, i], "oracle", ro[1]["metrics"][:, j, i], ro[1]["alphabeta"][:, j, i])
    for key in ("metrics", "alphabeta", "ipb", "dist_np"):
        eq(tag + " " + key, rg[1][key], ro[1][key])
    eq(tag + " field", rg[0], ro[0])
    return tag


def case_akima(gpu, rng, k):
    iorca = 0
    if rng.random() < 0.25:   # ORCA-shaped source below its north cap (AKIMA_2D wants longitudes increasing along i): fold options still run
        piv = str(rng.choice(["F", "T"]))
        lon_s, lat_s = G.orca_like(int(rng.integers(24, 90)), int(rng.integers(24, 70)), piv, float(rng.uniform(0.0, 200.0)))
        kew, iorca = 2, int(rng.choice([0, 6 if piv == "F" else 4]))
        X2, Y2 = lon_s, lat_s
    else:
        lon_s, lat_s, kew = rnd_regular_axes(rng)
        if lon_s.size < 6 or lat_s.size < 6 or np.any(np.diff(lon_s) <= 0):
            return None
        X2, Y2 = G.expand_2d(lon_s, lat_s)
    Xt, Yt = rnd_target(rng, X2, Y2)
    Z1 = G.field_slabs(X2, Y2, int(rng.integers(1, 4)), seed=int(rng.integers(1 << 30)))
    tag = f"akima#{k} src={X2.shape} trg={np.shape(Xt)} kew={kew} iorca={iorca}"
    gpu.akima_untiled_slopes(int(rng.choice([0, 1, 3])))   # tiled slope arrays / per-point slope arrays / fused
    with np.errstate(all="ignore"):
        rg, ro = both(tag, lambda: gpu.akima_2d(kew, lon_s, lat_s, Z1, Xt, Yt, i_orca_src=iorca), lambda: orc.akima_2d(kew, lon_s, lat_s, Z1, Xt, Yt, i_orca_src=iorca)[0])
    if rg is None:
        return tag + " (both STOP)"
    eq(tag, rg, ro)
    return tag


def case_drown(gpu, rng, k):
    ni, nj = int(rng.integers(4, 120)), int(rng.integers(4, 90))
    nslab = int(rng.integers(1, 5))
    X = rng.normal(10.0, 5.0, (nslab, nj, ni)).astype(np.float32)
    per_slab = rng.random() < 0.4
    shape = (nslab, nj, ni) if per_slab else (nj, ni)
    land = rng.random(shape) < rng.uniform(0.0, 0.9)
    if rng.random() < 0.5:   # blobs rather than salt and pepper
        yy, xx = np.mgrid[0:nj, 0:ni]
        land = (np.sin(xx * rng.uniform(0.05, 0.5)) * np.cos(yy * rng.uniform(0.05, 0.5)) > rng.uniform(-0.5, 0.8))
        land = np.broadcast_to(land, shape).copy()
    if rng.random() < 0.1:
        land[...] = True     # everything land
    mask = (~land).astype(np.int32)
    X[np.broadcast_to(mask, X.shape) == 0] = -9999.0
    if rng.random() < 0.2:
        X.flat[rng.integers(0, X.size, 5)] = np.float32(-0.0)
    kew = int(rng.choice([-1, 0, 2]))
    ninc, nsm = int(rng.integers(0, 40)), int(rng.integers(0, 12))
    pmin, pmax = (-1.0e3, 1.0e3) if rng.random() < 0.7 else (5.0, 12.0)
    pval = float(rng.choice([0.0, -9999.0, 1.0e-13]))
    mode = int(rng.integers(0, 3))
    gpu.bdrown_force_general(mode)
    which = rng.integers(0, 3)
    tag = f"drown#{k} kind={which} {ni}x{nj}x{nslab} per_slab={per_slab} kew={kew} inc={ninc} sm={nsm} mode={mode}"
    try:
        with np.errstate(all="ignore"):
            if which == 0:
                rg, ro = both(tag, lambda: gpu.bdrown(kew, X, mask, ninc, nsm, pmin, pmax, pval)[0], lambda: orc.bdrown(kew, X, mask, ninc, nsm, pmin, pmax, pval)[0])
            elif which == 1:
                m8 = mask if mask.ndim == 2 else mask[0]
                rg, ro = both(tag, lambda: gpu.drown(kew, X, m8.astype(np.int8), min(ninc, 12))[0], lambda: orc.drown(kew, X, m8.astype(np.int8), min(ninc, 12))[0])
            else:
                m2 = mask if mask.ndim == 2 else mask[0]
                use_mask = rng.random() < 0.6
                lemp = bool(use_mask and rng.random() < 0.5)
                rg, ro = both(tag, lambda: gpu.smoother(kew, X, nsm, m2 if use_mask else None, lemp), lambda: orc.smoother(kew, X, nsm, m2 if use_mask else None, lemp))
    finally:
        gpu.bdrown_force_general(0)
    if rg is None:
        return tag + " (both STOP)"
    eq(tag, rg, ro)
    return tag


def case_vert(gpu, rng, k):
    """vertical stage: AKIMA_1D_3D, AKIMA_1D_1D (missing values, persistence), the sigma branch, lbc_lnk, post-ops, extrp_hl,
    rotation"""
    which = int(rng.integers(0, 7))
    tag = f"vert#{k} kind={which}"
    with np.errstate(all="ignore"):
        if which == 0:
            nk1, nk2, ny, nx = int(rng.integers(3, 40)), int(rng.integers(1, 50)), int(rng.integers(1, 30)), int(rng.integers(1, 40))
            vz1 = np.cumsum(rng.uniform(1.0, 300.0, nk1)).astype(np.float32)
            if rng.random() < 0.15:
                vz1[int(rng.integers(1, nk1))] = vz1[0]                  # repeated / non-monotone depth
            vz2 = np.sort(rng.uniform(0.0, float(vz1.max()) * 1.1, nk2)).astype(np.float32)
            if nk2 > 2 and rng.random() < 0.5:
                vz2[int(rng.integers(0, nk2))] = vz1[int(rng.integers(0, nk1))]
            xf1 = rng.normal(5.0, 3.0, (nk1, ny, nx)).astype(np.float32)
            rg, ro = both(tag, lambda: gpu.akima_1d_3d(vz1, xf1, vz2, -9999.0), lambda: orc.akima_1d_3d(vz1, xf1, vz2, -9999.0))
            if rg is None:
                return tag
            eq(tag, rg, ro)
        elif which == 1:
            ncol, nk1, nk2 = int(rng.integers(1, 300)), int(rng.integers(3, 40)), int(rng.integers(1, 40))
            vz1 = np.cumsum(rng.uniform(1.0, 300.0, (ncol, nk1)), axis=1)
            vf1 = rng.normal(5.0, 3.0, (ncol, nk1))
            vz2 = np.sort(rng.uniform(-10.0, vz1.max() * 1.1, (ncol, nk2)), axis=1)
            vz2[:, 0] = vz1[:, 0] if rng.random() < 0.5 else vz2[:, 0]
            rmv = -9999.0 if rng.random() < 0.5 else None
            if rmv is not None:
                for c in range(ncol):
                    a, b = (int(rng.integers(0, 3)), int(rng.integers(1, 4))) if rng.random() < 0.7 else (0, 0)
                    if a:
                        vf1[c, :a] = rmv
                    if b and (a or rng.random() < 0.8):
                        vf1[c, nk1 - b:] = rmv
            opts = dict(rmask_val=rmv, l_surf_persist=bool(rng.random() < 0.5), l_botm_persist=bool(rng.random() < 0.5))
            if ORACLE_ONLY:
                orc.akima_1d_1d(vz1, vf1, vz2, **opts)
                return tag
            sg = None
            try:
                a, sg = gpu.akima_1d_1d(vz1, vf1, vz2, **opts)
            except sosie_b200.SosieError:
                a, sg = gpu.last_vf2, gpu.last_status
                assert sg.min() < 0, tag
            b, sb = orc.akima_1d_1d(vz1, vf1, vz2, **opts)
            eq(tag + " status", sg, sb)
            ok = sb >= 0
            eq(tag, a[ok], b[ok])
        elif which == 2:
            nk_src, nk_trg, nj, ni = int(rng.integers(3, 30)), int(rng.integers(1, 30)), int(rng.integers(1, 20)), int(rng.integers(1, 30))
            ss, ts, lm = bool(rng.random() < 0.5), bool(rng.random() < 0.5), bool(rng.random() < 0.5)
            H = rng.uniform(100.0, 4000.0, (nj, ni))
            ds = (np.sort(rng.uniform(0.01, 1.0, (nk_src, nj, ni)), axis=0) * H).astype(np.float32)
            dt = (np.sort(rng.uniform(0.0, 1.2, (nk_trg, nj, ni)), axis=0) * H).astype(np.float32)
            da = rng.normal(10.0, 4.0, ds.shape).astype(np.float32)
            if ss:
                ds, da = ds[::-1].copy(), da[::-1].copy()
            if ts:
                dt = dt[::-1].copy()
            nwet = rng.integers(0, nk_trg + 1, (nj, ni))
            mask = (np.arange(nk_trg)[:, None, None] < nwet[None]).astype(np.int32)
            prev = rng.normal(size=dt.shape).astype(np.float32)
            rg, ro = both(tag, lambda: gpu.interp3d_sigma(ds, da, dt, prev, mask, lm, ss, ts), lambda: orc.interp3d_sigma(ds, da, dt, prev, mask, lm, ss, ts)[0])
            if rg is None:
                return tag
            eq(tag, rg, ro)
        elif which == 3:
            jpj, jpi, ns = int(rng.integers(4, 40)), int(rng.integers(4, 60)), int(rng.integers(1, 4))
            X = rng.normal(size=(ns, jpj, jpi)).astype(np.float32)
            nperio, cd = int(rng.integers(0, 7)), str(rng.choice(["T", "U", "V", "F", "W"]))
            psgn, pval = float(rng.choice([1.0, -1.0])), (float(rng.normal()) if rng.random() < 0.3 else None)
            tag += f" nperio={nperio} cd={cd} {jpi}x{jpj}"
            rg, ro = both(tag, lambda: gpu.lbc_lnk(nperio, X, cd, psgn, pval), lambda: orc.lbc_lnk(nperio, X, cd, psgn, pval))
            if rg is None:
                return tag
            eq(tag, rg, ro)
        elif which == 4:
            nk, nj, ni = int(rng.integers(1, 20)), int(rng.integers(4, 30)), int(rng.integers(4, 40))
            X = rng.normal(10.0, 8.0, (nk, nj, ni)).astype(np.float32)
            X[rng.random(X.shape) < 0.05] = -9999.0
            nwet = rng.integers(0, nk + 1, (nj, ni))
            mask = (np.arange(nk)[:, None, None] < nwet[None]).astype(np.int32)
            args = (mask, int(rng.integers(0, 2)), 2.0, 18.0, -9999.0, int(rng.choice([0, 0, 4, 6])), str(rng.choice(["T", "U", "V"])), bool(rng.random() < 0.5))
            rg, ro = both(tag, lambda: gpu.interp3d_post(X, *args), lambda: orc.interp3d_post(X, *args))
            if rg is None:
                return tag
            eq(tag, rg, ro)
        elif which == 5:
            ns, nj, ni = int(rng.integers(1, 4)), int(rng.integers(6, 40)), int(rng.integers(2, 30))
            X = rng.normal(size=(ns, nj, ni)).astype(np.float32)
            icr = int(rng.choice([1, -1]))
            # as INTERP_2D/3D set them (src/mod_interp.f90:56-65): the "top" rows lie beyond the "bottom" rows in the direction icr
            hi_rows, lo_rows = int(rng.integers(nj // 2 + 1, nj + 1)), int(rng.integers(1, nj // 2))
            top, btm = (hi_rows, lo_rows) if icr == 1 else (lo_rows, hi_rows)
            if rng.random() < 0.3:
                top = 0
            if rng.random() < 0.3:
                btm = 0
            if icr == 1:
                top, btm = (max(top, 2) if top else 0), btm
            else:
                top, btm = (min(top, nj - 1) if top else 0), btm
            rg, ro = both(tag, lambda: gpu.extrp_hl(X, top, btm, icr), lambda: orc.extrp_hl(X, top, btm, icr))
            if rg is None:
                return tag
            eq(tag, rg, ro)
        else:
            ns, n = int(rng.integers(1, 4)), int(rng.integers(1, 3000))
            U, V = rng.normal(size=(ns, n)).astype(np.float32), rng.normal(size=(ns, n)).astype(np.float32)
            th = rng.uniform(0, 2 * np.pi, n)
            inv = bool(rng.random() < 0.5)
            rg, ro = both(tag, lambda: gpu.rotate_uv(U, V, np.cos(th), np.sin(th), inv), lambda: orc.rotate_uv(U, V, np.cos(th), np.sin(th), inv))
            if rg is None:
                return tag
            eq(tag + " U", rg[0], ro[0]); eq(tag + " V", rg[1], ro[1])
    return tag


def case_track(gpu, rng, k):
    nk = int(rng.choice([1, 1, int(rng.integers(2, 9))]))
    nj, ni, n = int(rng.integers(3, 60)), int(rng.integers(3, 80)), int(rng.integers(1, 4000))
    F1 = rng.normal(size=(nk, nj, ni)).astype(np.float32)
    timed = nk == 1 and rng.random() < 0.7
    F2 = rng.normal(size=(nk, nj, ni)).astype(np.float32) if timed else None
    if rng.random() < 0.3:
        F1[rng.random(F1.shape) < 0.02] = -9999.0
    imask = (rng.random((nk, nj, ni)) > rng.uniform(0.0, 0.2)).astype(np.int8)
    iP, jP, iq = rng.integers(0, ni + 1, n), rng.integers(0, nj + 1, n), rng.integers(1, 5, n)
    xa, xb = rng.uniform(-0.1, 1.1, n), rng.uniform(-0.1, 1.1, n)
    r_obs = rng.uniform(-30, 60, n)
    vt1 = float(rng.uniform(1.0e4, 3.0e4)); vt2 = vt1 + float(rng.uniform(0.01, 30.0))
    rt = rng.uniform(vt1, vt2, n) if timed else None
    lo, hi, clip = (-20.0, 20.0, True) if timed else (-20.0, 50.0, bool(rng.random() < 0.5))
    tag = f"track#{k} nk={nk} {ni}x{nj} n={n} timed={timed}"
    args = (F1, F2, vt1, vt2, imask, rt, iP, jP, iq, xa, xb, r_obs, lo, hi, clip)
    rg, ro = both(tag, lambda: gpu.track_interp(*args), lambda: orc.track_interp(*args))
    if rg is None:
        return tag
    for a_, b_, nm in zip(rg, ro, ("f_bl", "f_np", "fmask")):
        eq(tag + " " + nm, a_, b_)
    return tag


CASES = (case_bilin, case_bilin, case_akima, case_drown, case_vert, case_track)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    orc.build()
    orc.set_libm(orc.PMATH)   # the bit-exact checker: same portable libm as the kernels (glibc differs in the last ulp)
    gpu = _NoGpu() if ORACLE_ONLY else sosie_b200.Sosie(0)
    fails, done = [], 0
    for k in range(n):
        sub = np.random.default_rng(rng.integers(1 << 62))
        fn = CASES[k % len(CASES)]
        if ONLY and k not in ONLY:
            continue
        try:
            tag = fn(gpu, sub, k)
            done += tag is not None
        except AssertionError as e:
            fails.append(str(e))
            print("FAIL", e, flush=True)
        except Exception as e:  # noqa: BLE001
            fails.append(f"case {k} ({fn.__name__}): {e!r}")
            print("ERROR", k, fn.__name__, repr(e), flush=True)
            traceback.print_exc()
            gpu = _NoGpu() if ORACLE_ONLY else sosie_b200.Sosie(0)
    print(f"fuzz: {done} cases compared, {len(fails)} failures (seed {seed})")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
