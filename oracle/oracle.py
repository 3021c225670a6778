"""TEST INFRASTRUCTURE -- ctypes front end of the CPU oracle (oracle/orc_*.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this.  The product package (sosie_b200/) never does.  "parity unpinned": the reference pins no
numbers for this path, so the oracle is a statement-by-statement restatement.

Array convention: Fortran (ni,nj) column-major == numpy shape (nj,ni) C-contiguous.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsosie_oracle.so")
_lib = None


def build(force: bool = False):
    """Compile the oracle with the committed Makefile (g++ -O3 -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("orc_bilin.cpp", "orc_akima_drown.cpp", "orc_vert.cpp", "orc_common.hpp")]
    srcs.append(os.path.join(_HERE, "..", "sosie_b200", "csrc", "pmath.h"))
    if not force and os.path.exists(_LIB_PATH):
        t = os.path.getmtime(_LIB_PATH)
        if all(os.path.getmtime(s) <= t for s in srcs if os.path.exists(s)):
            return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_distance.restype = C.c_double
        _lib.orc_distance.argtypes = [C.c_double] * 4
        _lib.orc_heading.restype = C.c_double
        _lib.orc_heading.argtypes = [C.c_double] * 4
        _lib.orc_earth_radius.restype = C.c_double
        _lib.orc_interp_bl.restype = C.c_float
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


GLIBC, PMATH = 0, 1


def set_libm(mode: int):
    """0 = glibc (reference platform), 1 = pmath (portable; bit-identical to the CUDA kernels)."""
    lib().orc_set_libm(int(mode))


def get_error() -> int:
    return lib().orc_get_error()


def max_threads() -> int:
    return lib().orc_max_threads()


def distance(lona, lonb, lata, latb):
    return lib().orc_distance(lona, lonb, lata, latb)


def heading(lona, lonb, lata, latb):
    return lib().orc_heading(lona, lonb, lata, latb)


def earth_radius():
    return lib().orc_earth_radius()


def l_inpoly(vplon, vplat, px, py) -> bool:
    a, b = _f64(vplon), _f64(vplat)
    return bool(lib().orc_l_inpoly(_p(a, C.c_double), _p(b, C.c_double), C.c_double(px), C.c_double(py)))


def local_coord(xlam, xphi):
    a, b = _f64(xlam), _f64(xphi)
    xa, xb, ipb = C.c_double(), C.c_double(), C.c_int()
    lib().orc_local_coord(_p(a, C.c_double), _p(b, C.c_double), C.byref(xa), C.byref(xb), C.byref(ipb))
    return xa.value, xb.value, ipb.value


def pmath_eval(fn: int, x, y=None):
    x = _f64(x)
    y = _f64(y) if y is not None else np.zeros_like(x)
    a, b = np.empty_like(x), np.empty_like(x)
    lib().orc_pmath_eval(fn, C.c_int64(x.size), _p(x, C.c_double), _p(y, C.c_double), _p(a, C.c_double), _p(b, C.c_double))
    return a, b


def is_grid_regular(X, Y) -> bool:
    X, Y = _f64(X), _f64(Y)
    ny, nx = X.shape
    return bool(lib().orc_is_grid_regular(nx, ny, _p(X, C.c_double), _p(Y, C.c_double)))


def find_nearest_point(Xtrg, Ytrg, Xsrc, Ysrc, mask=None, elide_fill=True, nthreads=1):
    Xtrg, Ytrg, Xsrc, Ysrc = _f64(Xtrg), _f64(Ytrg), _f64(Xsrc), _f64(Ysrc)
    nyt, nxt = Xtrg.shape
    nys, nxs = Xsrc.shape
    JI = np.empty((nyt, nxt), np.int32)
    JJ = np.empty((nyt, nxt), np.int32)
    m = np.ones((nyt, nxt), np.int32) if mask is None else _i32(mask).copy()
    info = np.zeros(5, np.int32)
    used = lib().orc_find_nearest_point(nxt, nyt, _p(Xtrg, C.c_double), _p(Ytrg, C.c_double), nxs, nys,
                                        _p(Xsrc, C.c_double), _p(Ysrc, C.c_double), _p(JI, C.c_int32), _p(JJ, C.c_int32),
                                        _p(m, C.c_int32), int(elide_fill), int(nthreads), _p(info, C.c_int32))
    return JI, JJ, m, used, info


def mapping_bl(k_ew, X1, Y1, lon_trg, lat_trg, mask=None, elide_fill=True, nthreads=1):
    """MAPPING_BL: returns dict(metrics (3,nj,ni) i32, alphabeta (2,nj,ni) f64, ipb i16, dist_np f32, mask)."""
    X1, Y1, lon_trg, lat_trg = _f64(X1), _f64(Y1), _f64(lon_trg), _f64(lat_trg)
    nyi, nxi = X1.shape
    nyo, nxo = lon_trg.shape
    m = np.ones((nyo, nxo), np.int32) if mask is None else _i32(mask).copy()
    met = np.empty((3, nyo, nxo), np.int32)
    ab = np.empty((2, nyo, nxo), np.float64)
    ipb = np.empty((nyo, nxo), np.int16)
    dnp = np.empty((nyo, nxo), np.float32)
    info = np.zeros(5, np.int32)
    lib().orc_mapping_bl(int(k_ew), nxi, nyi, _p(X1, C.c_double), _p(Y1, C.c_double), nxo, nyo, _p(lon_trg, C.c_double),
                         _p(lat_trg, C.c_double), _p(m, C.c_int32), _p(met, C.c_int32), _p(ab, C.c_double),
                         _p(ipb, C.c_int16), _p(dnp, C.c_float), int(elide_fill), int(nthreads), _p(info, C.c_int32))
    return dict(metrics=met, alphabeta=ab, ipb=ipb, dist_np=dnp, mask=m, info=info)


def ext_north_to_90_regg(XX, YY, XF):
    XX, YY, XF = _f64(XX), _f64(YY), _f32(XF)
    ny, nx = XX.shape
    XP = np.empty((ny + 2, nx)); YP = np.empty((ny + 2, nx)); FP = np.empty((ny + 2, nx), np.float32)
    lib().orc_ext_north_to_90_regg(nx, ny, _p(XX, C.c_double), _p(YY, C.c_double), _p(XF, C.c_float), _p(XP, C.c_double),
                                   _p(YP, C.c_double), _p(FP, C.c_float))
    return XP, YP, FP


def interp_bl(k_ew, iP, jP, iqd, xa, xb, Z_in):
    Z = _f32(Z_in)
    ny, nx = Z.shape
    return lib().orc_interp_bl(int(k_ew), nx, ny, int(iP), int(jP), int(iqd), C.c_double(xa), C.c_double(xb), _p(Z, C.c_float))


def track_interp(F1, F2, vt1, vt2, imask, rt, jiP, jjP, iqd, xa, xb, r_obs, obs_lo=-20.0, obs_hi=20.0, clip_missing=True, rmissval=-9999.0):
    """per-point block of interp_to_ground_track.f90:718-779 / interp_to_hydro_section.f90:567-607.  F1 (and F2 or None):
    (nk, nj, ni) or (nj, ni) f32; imask same shape i8.  Returns (f_bl, f_np, fmask), each (n, nk), pre-filled with rmissval / 0."""
    F1 = _f32(F1)
    if F1.ndim == 2:
        F1 = F1[None]
    nk, nj, ni = F1.shape
    F2 = None if F2 is None else _f32(F2).reshape(F1.shape)
    imask = np.ascontiguousarray(imask, np.int8).reshape(F1.shape)
    jiP, jjP, iqd = _i32(jiP), _i32(jjP), _i32(iqd)
    xa, xb, r_obs = _f64(xa), _f64(xb), _f64(r_obs)
    n = jiP.size
    rt = _f64(rt) if rt is not None else np.zeros(n)
    f_bl = np.full((n, nk), rmissval, np.float64); f_np = np.full((n, nk), rmissval, np.float64); fmask = np.zeros((n, nk), np.int8)
    lib().orc_track_interp(ni, nj, nk, _p(F1, C.c_float), _p(F2, C.c_float), C.c_double(vt1), C.c_double(vt2), _p(imask, C.c_int8),
                           C.c_int64(n), _p(rt, C.c_double), _p(jiP, C.c_int32), _p(jjP, C.c_int32), _p(iqd, C.c_int32),
                           _p(xa, C.c_double), _p(xb, C.c_double), _p(r_obs, C.c_double), C.c_double(obs_lo), C.c_double(obs_hi),
                           int(clip_missing), C.c_double(rmissval), _p(f_bl, C.c_double), _p(f_np, C.c_double), _p(fmask, C.c_int8))
    return f_bl, f_np, fmask


def bilin_2d(k_ew, X10, Y10, Z1, X20, Y20, mask_domain_trg=None, l_reg_src=False, i_orca_trg=0, mapping=None,
             elide_fill=True, nthreads=1, virtual_src=False):
    """BILIN_2D over nslab source slabs Z1 (nslab, ny1, nx1).  X10/Y10 (and X20/Y20) are either 1-D axes or
    2-D (nj,ni) arrays.  Returns (Z2 (nslab, ny2, nx2), mapping dict)."""
    Z1 = _f32(Z1)
    if Z1.ndim == 2:
        Z1 = Z1[None]
    nslab, ny1, nx1 = Z1.shape
    X10, Y10, X20, Y20 = _f64(X10), _f64(Y10), _f64(X20), _f64(Y20)
    src_1d = X10.ndim == 1
    trg_1d = X20.ndim == 1
    if trg_1d:
        nx2, ny2 = X20.size, Y20.size
    else:
        ny2, nx2 = X20.shape
    Z2 = np.empty((nslab, ny2, nx2), np.float32)
    have = mapping is not None
    if have:
        met = _i32(mapping["metrics"]).copy(); ab = _f64(mapping["alphabeta"]).copy()
        ipb = np.ascontiguousarray(mapping["ipb"], np.int16).copy(); dnp = _f32(mapping["dist_np"]).copy()
    else:
        met = np.zeros((3, ny2, nx2), np.int32); ab = np.zeros((2, ny2, nx2)); ipb = np.zeros((ny2, nx2), np.int16)
        dnp = np.zeros((ny2, nx2), np.float32)
    m = None if mask_domain_trg is None else _i32(mask_domain_trg)
    lrow = C.c_int(0)
    info = np.zeros(5, np.int32)
    rc = lib().orc_bilin_2d(int(k_ew), nx1, ny1, _p(X10, C.c_double), _p(Y10, C.c_double), int(src_1d), nslab,
                            _p(Z1, C.c_float), nx2, ny2, _p(X20, C.c_double), _p(Y20, C.c_double), int(trg_1d),
                            _p(Z2, C.c_float), _p(m, C.c_int32), int(l_reg_src), int(i_orca_trg), 1, int(have),
                            _p(met, C.c_int32), _p(ab, C.c_double), _p(ipb, C.c_int16), _p(dnp, C.c_float), C.byref(lrow),
                            int(elide_fill), int(nthreads), _p(info, C.c_int32), int(virtual_src))
    if rc:
        raise RuntimeError(f"oracle BILIN_2D: reference would STOP (code {rc})")
    return Z2, dict(metrics=met, alphabeta=ab, ipb=ipb, dist_np=dnp, l_last_y_row_missing=lrow.value, info=info)


def fill_extra_bands(k_ew, XX, YY, XF, iorca=0):
    XX, YY, XF = _f64(XX), _f64(YY), _f64(XF)
    ny, nx = XX.shape
    o = [np.empty((ny + 4, nx + 4)) for _ in range(3)]
    lib().orc_fill_extra_bands(int(k_ew), nx, ny, _p(XX, C.c_double), _p(YY, C.c_double), _p(XF, C.c_double),
                               _p(o[0], C.c_double), _p(o[1], C.c_double), _p(o[2], C.c_double), int(iorca))
    return o


def slopes_akima(k_ew, XX, XY, XF):
    XX, XY, XF = _f64(XX), _f64(XY), _f64(XF)
    ny, nx = XX.shape
    o = [np.empty((ny, nx)) for _ in range(3)]
    lib().orc_slopes_akima(int(k_ew), nx, ny, _p(XX, C.c_double), _p(XY, C.c_double), _p(XF, C.c_double),
                           _p(o[0], C.c_double), _p(o[1], C.c_double), _p(o[2], C.c_double))
    return o


def akima_2d(k_ew, X10, Y10, Z1, X20, Y20, i_orca_src=0, nthreads=1):
    Z1 = _f32(Z1)
    if Z1.ndim == 2:
        Z1 = Z1[None]
    nslab, ny1, nx1 = Z1.shape
    X10, Y10, X20, Y20 = _f64(X10), _f64(Y10), _f64(X20), _f64(Y20)
    src_1d = X10.ndim == 1
    trg_1d = X20.ndim == 1
    if trg_1d:
        nx2, ny2 = X20.size, Y20.size
    else:
        ny2, nx2 = X20.shape
    Z2 = np.empty((nslab, ny2, nx2), np.float32)
    ixy = np.zeros((2, ny2, nx2), np.int32)
    lib().orc_akima_2d(int(k_ew), nx1, ny1, _p(X10, C.c_double), _p(Y10, C.c_double), int(src_1d), nslab, _p(Z1, C.c_float),
                       nx2, ny2, _p(X20, C.c_double), _p(Y20, C.c_double), int(trg_1d), _p(Z2, C.c_float), int(i_orca_src), 1,
                       _p(ixy, C.c_int32), int(nthreads))
    return Z2, ixy


def bdrown(k_ew, X, mask, nb_inc=200, nb_smooth=0, pmin=-1.0e24, pmax=1.0e24, pval_land=0.0, nthreads=1):
    """BDROWN on X (nslab, nj, ni) or (nj, ni) (copy returned).  mask (nj,ni) shared or (nslab,nj,ni)."""
    X = _f32(X).copy()
    sq = X.ndim == 2
    if sq:
        X = X[None]
    nslab, nj, ni = X.shape
    mask = _i32(mask)
    per = mask.ndim == 3
    nsw = np.zeros(nslab, np.int32)
    lib().orc_bdrown(int(k_ew), ni, nj, nslab, _p(X, C.c_float), _p(mask, C.c_int32), int(per), int(nb_inc), int(nb_smooth),
                     C.c_float(pmin), C.c_float(pmax), C.c_float(pval_land), _p(nsw, C.c_int32), int(nthreads))
    return (X[0] if sq else X), nsw


def smoother(k_ew, X, nb_smooth=10, msk=None, l_exclude_mask_points=False):
    X = _f32(X).copy()
    sq = X.ndim == 2
    if sq:
        X = X[None]
    nslab, nj, ni = X.shape
    m = None if msk is None else _i32(msk)
    rc = lib().orc_smoother(int(k_ew), ni, nj, nslab, _p(X, C.c_float), int(nb_smooth), _p(m, C.c_int32), int(l_exclude_mask_points))
    if rc:
        raise RuntimeError("oracle SMOOTHER: reference would STOP")
    return X[0] if sq else X


def drown(k_ew, X, mask, nb_inc=200, nthreads=1):
    X = _f32(X).copy()
    sq = X.ndim == 2
    if sq:
        X = X[None]
    nslab, nj, ni = X.shape
    mask = np.ascontiguousarray(mask, np.int8)
    nsw = np.zeros(nslab, np.int32)
    rc = lib().orc_drown(int(k_ew), ni, nj, nslab, _p(X, C.c_float), _p(mask, C.c_int8), int(nb_inc), _p(nsw, C.c_int32), int(nthreads))
    if rc:
        raise RuntimeError(f"oracle DROWN: the reference would index out of bounds (code {rc})")
    return (X[0] if sq else X), nsw


def rotate_uv(U, V, xcos, xsin, inverse=False):
    U, V = _f32(U), _f32(V)
    xcos, xsin = _f64(xcos), _f64(xsin)
    n = xcos.size
    nslab = U.size // n
    Uo, Vo = np.empty_like(U), np.empty_like(V)
    lib().orc_rotate_uv(C.c_int64(n), nslab, _p(U, C.c_float), _p(V, C.c_float), _p(xcos, C.c_double), _p(xsin, C.c_double),
                        _p(Uo, C.c_float), _p(Vo, C.c_float), int(inverse))
    return Uo, Vo


# ---- vertical stage of INTERP_3D (orc_vert.cpp)
def akima_1d_3d(vz1, xf1, vz2, rfill_val=-9999.0):
    """AKIMA_1D_3D, src/mod_akima_1d.f90:243-389: xf1 (nk1, ny, nx) -> (nk2, ny, nx)"""
    vz1, vz2, xf1 = _f32(vz1), _f32(vz2), _f32(xf1)
    nk1, ny, nx = xf1.shape
    xf2 = np.empty((vz2.size, ny, nx), np.float32)
    lib().orc_akima_1d_3d(nx, ny, nk1, _p(vz1, C.c_float), _p(xf1, C.c_float), vz2.size, _p(vz2, C.c_float), _p(xf2, C.c_float),
                          C.c_float(rfill_val))
    return xf2


def akima_1d_1d(vz1, vf1, vz2, rmask_val=None, l_surf_persist=False, l_botm_persist=False):
    """AKIMA_1D_1D, src/mod_akima_1d.f90:26-232, over columns: vz1 (nk1,) or (ncol, nk1), vf1 (ncol, nk1) or (nk1,),
    vz2 (nk2,) or (ncol, nk2) -> (vf2 (ncol, nk2) f32, status (ncol,) i32)"""
    vf1 = np.atleast_2d(_f64(vf1))
    ncol, nk1 = vf1.shape
    vz1 = _f64(vz1)
    vz2 = _f64(vz2)
    z1cs = nk1 if vz1.ndim == 2 else 0
    nk2 = vz2.shape[-1]
    z2cs = nk2 if vz2.ndim == 2 else 0
    vf2 = np.empty((ncol, nk2), np.float32)
    status = np.zeros(ncol, np.int32)
    lib().orc_akima_1d_1d(C.c_int64(ncol), nk1, _p(vz1, C.c_double), C.c_int64(z1cs), C.c_int64(1), _p(vf1, C.c_double), C.c_int64(nk1),
                          C.c_int64(1), nk2, _p(vz2, C.c_double), C.c_int64(z2cs), C.c_int64(1), _p(vf2, C.c_float), C.c_int64(nk2),
                          C.c_int64(1), int(rmask_val is not None), C.c_double(0.0 if rmask_val is None else rmask_val),
                          int(l_surf_persist), int(l_botm_persist), _p(status, C.c_int32))
    return vf2, status


def interp3d_sigma(depth_src, data3d_tmp, depth_trg, data3d_trg, mask_trg=None, lmout=False, src_sigma=False, trg_sigma=False):
    """generic-coordinate branch of INTERP_3D, src/mod_interp.f90:374-438: cubes (nk, nj, ni); returns (data3d_trg, worst status)"""
    depth_src, data3d_tmp, depth_trg = _f32(depth_src), _f32(data3d_tmp), _f32(depth_trg)
    out = _f32(data3d_trg).copy()
    nk_src, nj, ni = data3d_tmp.shape
    nk_trg = depth_trg.shape[0]
    mk = _i32(mask_trg) if mask_trg is not None else None
    st = lib().orc_interp3d_sigma(ni, nj, nk_src, _p(depth_src, C.c_float), _p(data3d_tmp, C.c_float), nk_trg, _p(depth_trg, C.c_float),
                                  _p(mk, C.c_int32), int(lmout), int(src_sigma), int(trg_sigma), _p(out, C.c_float))
    return out, st


def lbc_lnk(nperio, X, cd_type="T", psgn=1.0, pval=None):
    """lbc_lnk_2d_r4, src/mod_nemotools.f90:154: slabs X (nslab, jpj, jpi) or (jpj, jpi)"""
    X = _f32(X).copy()
    sq = X.ndim == 2
    if sq:
        X = X[None]
    nslab, jpj, jpi = X.shape
    lib().orc_lbc_lnk(int(nperio), jpi, jpj, nslab, _p(X, C.c_float), ord(cd_type), C.c_double(psgn), int(pval is not None),
                      C.c_double(0.0 if pval is None else pval))
    return X[0] if sq else X


def interp3d_post(X, mask_trg=None, ixtrpl_bot=0, vmin=-1.0e30, vmax=1.0e30, rmiss_val=-9999.0, i_orca_trg=0, c_orca_trg="T", lmout=False):
    """post-interpolation block of INTERP_3D, src/mod_interp.f90:447-511: X (nk, nj, ni)"""
    X = _f32(X).copy()
    nk, nj, ni = X.shape
    m = None if mask_trg is None else _i32(mask_trg)
    lib().orc_interp3d_post(ni, nj, nk, _p(X, C.c_float), _p(m, C.c_int32), int(ixtrpl_bot), C.c_float(vmin), C.c_float(vmax),
                            C.c_float(rmiss_val), int(i_orca_trg), ord(c_orca_trg), int(lmout))
    return X


def extrp_hl(X, jj_ex_top, jj_ex_btm, nlat_icr_trg=1):
    """extrp_hl, src/mod_interp.f90:521-542: X (nslab, nj, ni) or (nj, ni)"""
    X = _f32(X).copy()
    sq = X.ndim == 2
    if sq:
        X = X[None]
    nslab, nj, ni = X.shape
    lib().orc_extrp_hl(ni, nj, nslab, _p(X, C.c_float), int(jj_ex_top), int(jj_ex_btm), int(nlat_icr_trg))
    return X[0] if sq else X


def angle2(nperio, lamt, phit, lamu, phiu, lamv, phiv, lamf, phif):
    """angle2, src/mod_nemotools.f90:607-823: returns (cost, sint, cosu, sinu, cosv, sinv, cosf, sinf), each (ny, nx) f64"""
    ins = [_f64(a) for a in (lamt, phit, lamu, phiu, lamv, phiv, lamf, phif)]
    ny, nx = ins[0].shape
    outs = [np.empty((ny, nx), np.float64) for _ in range(8)]
    lib().orc_angle2(int(nperio), nx, ny, *[_p(a, C.c_double) for a in ins], *[_p(a, C.c_double) for a in outs])
    return tuple(outs)
