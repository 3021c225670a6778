"""TEST INFRASTRUCTURE -- an independent, slow numpy/Python transliteration of parts of the reference, used only to
pin the C++ oracle ("two independent transcriptions must agree bit-for-bit").  Array-syntax
statements of the Fortran source are written as numpy slicing (a different route from the C++ scalar loops).

Arrays here are Fortran-shaped (ni, nj) numpy arrays indexed [i, j] 0-based; helpers convert from the (nj, ni)
C-order convention of the rest of the repo.
"""
import math

import numpy as np

F32 = np.float32
RIS2 = F32(1.0) / np.sqrt(F32(2.0))


def to_f(a):        # (nj, ni) C-order -> (ni, nj) view
    return np.ascontiguousarray(a).T


def from_f(a):
    return np.ascontiguousarray(a.T)


# ---------------------------------------------------------------------------------------------------------------
# scalar geometry (src/mod_manip.f90:1443-1504, src/mod_bilin_2d.f90:697-870, src/mod_poly.f90:54-266), glibc libm
ZR = float((F32(6378.137) + F32(6356.7523)) / F32(2.0))
REPS = float(F32(1.0e-9))


def distance(lona, lonb, lata, latb):
    zconv = math.acos(-1.0) / 180.0
    ux = math.cos(lona * zconv) * math.cos(lata * zconv)
    uy = math.sin(lona * zconv) * math.cos(lata * zconv)
    uz = math.sin(lata * zconv)
    vx = math.cos(lonb * zconv) * math.cos(latb * zconv)
    vy = math.sin(lonb * zconv) * math.cos(latb * zconv)
    vz = math.sin(latb * zconv)
    pds = ux * vx + uy * vy + uz * vz
    return 0.0 if pds >= 1.0 else ZR * math.acos(pds)


def degE_to_degWE(x):
    return math.copysign(1.0, 180.0 - x) * min(x, abs(x - 360.0))


def heading(plona, plonb, plata, platb):
    zpi = math.acos(-1.0)
    zconv = zpi / 180.0
    xa, xb = plona * zconv, plonb * zconv
    ya = -math.log(max(abs(math.tan(zpi / 4.0 - zconv * plata / 2.0)), REPS))
    yb = -math.log(max(abs(math.tan(zpi / 4.0 - zconv * platb / 2.0)), REPS))
    d = math.fmod(xb - xa, 2 * zpi)
    if d >= zpi:
        d -= 2 * zpi
    if d <= -zpi:
        d += 2 * zpi
    h = math.atan2(d, yb - ya) * 180.0 / zpi
    return h + 360.0 if h < 0 else h


def l_inpoly(vplon, vplat, px, py):
    vplon = [float(v) for v in vplon]
    vplat = [float(v) for v in vplat]
    rmaxx, rminx = F32(max(vplon)), F32(min(vplon))
    rmaxy, rminy = F32(max(vplat)), F32(min(vplat))
    if rminx < 0 or px < 0:
        raise ValueError("negative longitude")
    if rminx < F32(10.0) and rmaxx > F32(350.0):
        zx = math.copysign(1.0, 180.0 - px) * min(px, abs(px - 360.0))
        vx = [F32(math.copysign(1.0, 180.0 - v) * min(v, abs(v - 360.0))) for v in vplon]
        rmaxx, rminx = max(vx), min(vx)
    else:
        zx = px
        vx = [F32(v) for v in vplon]
    zy = py
    vy = [F32(v) for v in vplat]
    vx.append(vx[0])
    vy.append(vy[0])
    for k in range(5):
        if vx[k] - F32(int(vx[k])) == 0:
            vx[k] = F32(vx[k] + F32(0.001))
        if vy[k] - F32(int(vy[k])) == 0:
            vy[k] = F32(vy[k] + F32(0.001))
    if not (zx <= float(rmaxx) and zx >= float(rminx) and zy <= float(rmaxy) and zy >= float(rminy)):
        return False
    icross = 0
    for ji in range(4):
        jn = 0 if ji == 3 else ji + 1
        xa, xb, ya, yb = float(vx[ji]), float(vx[jn]), float(vy[ji]), float(vy[jn])
        rise, run = yb - ya, xb - xa
        slope = 9999.0 if run == 0 else rise / run
        if abs(slope) <= float(F32(0.001)):
            slope = 0.0
        if slope == 0.0:
            continue
        if slope == 9999.0:
            if zx >= xa and ((zy <= ya and zy > yb) or (zy >= ya and zy < yb)):
                icross += 1
        else:
            rc = slope * xa - ya
            zxpt = (rc + zy) / slope
            if ((zxpt <= xa and zxpt > xb) or (zxpt >= xa and zxpt < xb)) and zx >= zxpt:
                icross += 1
    return icross % 2 == 1


def local_coord(xlam, xphi):
    zx = [float(v) for v in xlam]
    xp = [float(v) for v in xphi]
    if abs(zx[1] - zx[4]) >= 180.0 or abs(zx[1] - zx[2]) >= 180.0 or abs(zx[1] - zx[3]) >= 180.0:
        zx = [degE_to_degWE(v) for v in zx]
    zres, dlam, dphi, a, b, n = 1000.0, 0.5, 0.5, 0.0, 0.0, 0
    zresmax = float(F32(0.1))
    while zres > zresmax and n < 200:
        z1, z2, z3 = zx[2] - zx[1], zx[1] - zx[4], zx[3] - zx[2]
        a11 = z1 + (z2 + z3) * b
        a12 = -z2 + (z2 + z3) * a
        a21 = xp[2] - xp[1] + (xp[1] - xp[4] + xp[3] - xp[2]) * b
        a22 = xp[4] - xp[1] + (xp[1] - xp[4] + xp[3] - xp[2]) * a
        det = a11 * a22 - a12 * a21
        det = math.copysign(1.0, det) * max(abs(det), REPS)
        da = (dlam * a22 - a12 * dphi) / det
        db = (a11 * dphi - dlam * a21) / det
        zres = math.sqrt(da * da + db * db)
        a, b = a + da, b + db
        dlam = zx[0] - ((1.0 - a) * (1 - b) * zx[1] + a * (1 - b) * zx[2] + a * b * zx[3] + (1 - a) * b * zx[4])
        dphi = xp[0] - ((1.0 - a) * (1 - b) * xp[1] + a * (1 - b) * xp[2] + a * b * xp[3] + (1 - a) * b * xp[4])
        n += 1
    ipb = 0
    if n >= 200:
        a, b, ipb = 0.5, 0.5, 11
    if xp[1] == xp[2] == xp[3] == xp[4]:
        a, b, ipb = 0.5, 0.5, 12
    return a, b, ipb


def find_nearest_easy_point(rlon, rlat, lon1d, lat1d):
    """EASY search for one target on a regular source given by its 1-D axes (src/mod_manip.f90:1028-1042); 1-based result."""
    ip = int(np.argmin(np.abs(lon1d - rlon))) + 1
    jp = int(np.argmin(np.abs(lat1d - rlat))) + 1
    i1, i2 = max(ip - 4, 1), min(ip + 4, lon1d.size)
    j1, j2 = max(jp - 4, 1), min(jp + 4, lat1d.size)
    best, bi, bj = None, 0, 0
    zconv = math.acos(-1.0) / 180.0
    ux = math.cos(rlon * zconv) * math.cos(rlat * zconv)
    uy = math.sin(rlon * zconv) * math.cos(rlat * zconv)
    uz = math.sin(rlat * zconv)
    for jj in range(j1, j2 + 1):
        for ji in range(i1, i2 + 1):
            lo, la = lon1d[ji - 1] * zconv, lat1d[jj - 1] * zconv
            pds = ux * (math.cos(lo) * math.cos(la)) + uy * (math.sin(lo) * math.cos(la)) + uz * math.sin(la)
            d = ZR * math.acos(min(pds, 1.0))
            if best is None or d < best:
                best, bi, bj = d, ji, jj
    return bi, bj


def mapping_point_regular(k_ew, lon1d, lat1d, xP, yP, iP, jP):
    """One target of MAPPING_BL (src/mod_bilin_2d.f90:459-634) on a regular source; returns (iqdrn, alpha, beta, ipb) or None if skipped."""
    nxi, nyi = lon1d.size, lat1d.size
    if not (jP < nyi):
        return None
    iPm1, iPp1, jPm1, jPp1 = iP - 1, iP + 1, jP - 1, jP + 1
    if iPm1 == 0 and k_ew >= 0:
        iPm1 = nxi - k_ew
    if iPp1 == nxi + 1 and k_ew >= 0:
        iPp1 = 1 + k_ew
    if iPm1 < 1 or jPm1 < 1 or iPp1 > nxi:
        return None
    X = lambda i: math.fmod(lon1d[i - 1], 360.0)  # noqa: E731
    Y = lambda j: float(lat1d[j - 1])             # noqa: E731
    lonP, latP = X(iP), Y(jP)
    lonN, latN, lonE, latE = X(iP), Y(jPp1), X(iPp1), Y(jP)
    lonS, latS, lonW, latW = X(iP), Y(jPm1), X(iPm1), Y(jP)
    xP = math.fmod(xP, 360.0)
    hP = heading(lonP, xP, latP, yP)
    hN, hE = heading(lonP, lonN, latP, latN), heading(lonP, lonE, latP, latE)
    hS, hW = heading(lonP, lonS, latP, latS), heading(lonP, lonW, latP, latW)
    iq = 4
    hPp = degE_to_degWE(hP)
    if hN > hE:
        hN -= 360.0
    if hPp > hN and hPp <= hE:
        iq = 1
    if hP > hE and hP <= hS:
        iq = 2
    if hP > hS and hP <= hW:
        iq = 3
    if hP > hW and hPp <= hN:
        iq = 4
    icpt = 0
    while True:
        if iq == 1:
            quad = [(lonE, latE), (X(iPp1), Y(jPp1)), (lonN, latN)]
        elif iq == 2:
            quad = [(lonS, latS), (X(iPp1), Y(jPm1)), (lonE, latE)]
        elif iq == 3:
            quad = [(lonW, latW), (X(iPm1), Y(jPm1)), (lonS, latS)]
        elif iq == 4:
            quad = [(lonN, latN), (X(iPm1), Y(jPp1)), (lonW, latW)]
        loni = [xP, lonP] + [q[0] for q in quad]
        lati = [yP, latP] + [q[1] for q in quad]
        loni = [v + 360.0 if v <= 0.0 else v for v in loni]
        ok = l_inpoly(loni[1:], lati[1:], xP, yP)
        if (not ok) and yP < 88.0:
            icpt += 1
            iq = icpt
        else:
            break
        if icpt == 5:
            iq = 0
            break
    a, b, ipb = local_coord(loni, lati)
    return iq, a, b, ipb


# ---------------------------------------------------------------------------------------------------------------
# BDROWN / SMOOTHER in array syntax (src/mod_bdrown.f90), X and mask are (ni, nj) [i, j]
def smoother(k_ew, X, nsmooth, msk=None, l_emp=False):
    X = X.astype(F32).copy()
    ni, nj = X.shape
    w0 = F32(0.35)
    omw = F32(1.0) - w0
    c4 = F32(4.0) * (F32(1.0) + RIS2)
    xorig = X.copy() if msk is not None else None
    if l_emp:
        nbp = np.zeros((ni, nj), np.int32)
        rdnm = np.full((ni, nj), F32(0.25))
    I = slice(1, ni - 1)   # 2:ni-1
    J = slice(1, nj - 1)
    Ip, Im = slice(2, ni), slice(0, ni - 2)
    Jp, Jm = slice(2, nj), slice(0, nj - 2)
    if k_ew >= 0:
        vim = [ni - k_ew - 1, ni - 2]   # 0-based of (ni-k_ew, ni-1)
        vip = [1, k_ew]                 # 0-based of (2, 1+k_ew)
        ivi = [0, ni - 1]
    for _ in range(nsmooth):
        xt = X.copy()
        if l_emp:
            xt = (xt * msk.astype(F32)).astype(F32)
        s4 = ((xt[Ip, J] + xt[I, Jp]) + xt[Im, J]) + xt[I, Jm]
        sd = ((xt[Ip, Jp] + xt[Ip, Jm]) + xt[Im, Jm]) + xt[Im, Jp]
        if l_emp:
            m = msk
            nbp[I, J] = m[Ip, J] + m[I, Jp] + m[Im, J] + m[I, Jm] + m[Ip, Jp] + m[Ip, Jm] + m[Im, Jm] + m[Im, Jp]
            i4 = (m[Ip, J] + m[I, Jp] + m[Im, J] + m[I, Jm]).astype(F32)
            idg = (m[Ip, Jp] + m[Ip, Jm] + m[Im, Jm] + m[Im, Jp]).astype(F32)
            rdnm[I, J] = F32(1.0) / np.maximum((i4 + RIS2 * idg).astype(F32), F32(0.01))
            X[I, J] = w0 * xt[I, J] + (omw * (s4 + RIS2 * sd)) * rdnm[I, J]
            X[I, J] = np.where(nbp[I, J] == 0, xorig[I, J], X[I, J])
        else:
            X[I, J] = w0 * xt[I, J] + (omw * (s4 + RIS2 * sd)) / c4
        if k_ew >= 0:
            for c in range(2):
                ji, jim, jip = ivi[c], vim[c], vip[c]
                s4 = ((xt[jip, J] + xt[ji, Jp]) + xt[jim, J]) + xt[ji, Jm]
                sd = ((xt[jip, Jp] + xt[jip, Jm]) + xt[jim, Jm]) + xt[jim, Jp]
                if l_emp:
                    m = msk
                    nbp[ji, J] = m[jip, J] + m[ji, Jp] + m[jim, J] + m[ji, Jm] + m[jip, Jp] + m[jip, Jm] + m[jim, Jm] + m[jim, Jp]
                    i4 = (m[jip, J] + m[ji, Jp] + m[jim, J] + m[ji, Jm]).astype(F32)
                    idg = (m[jip, Jp] + m[jip, Jm] + m[jim, Jm] + m[jim, Jp]).astype(F32)
                    rdnm[ji, J] = F32(1.0) / np.maximum((i4 + RIS2 * idg).astype(F32), F32(0.01))
                    X[ji, J] = w0 * xt[ji, J] + (omw * (s4 + RIS2 * sd)) * rdnm[ji, J]
                    X[ji, J] = np.where(nbp[ji, J] == 0, xorig[ji, J], X[ji, J])
                else:
                    X[ji, J] = w0 * xt[ji, J] + (omw * (s4 + RIS2 * sd)) / c4
        if msk is not None:
            mf = msk[:, J].astype(F32)
            X[:, J] = mf * X[:, J] - (msk[:, J] - 1).astype(F32) * xorig[:, J]
    return X


def bdrown(k_ew, X, mask, ninc_max, nsmooth, zmin, zmax, zval_land):
    X = X.astype(F32).copy()
    ni, nj = X.shape
    zmin, zmax = F32(zmin), F32(zmax)
    maskv = mask.astype(np.int32).copy()
    for jj in range(nj):
        nbp = int(mask[:, jj].sum())
        if nbp > 0:
            s = F32(0.0)
            for v, m in zip(X[:, jj], mask[:, jj]):      # sequential REAL(4) SUM
                s = F32(s + F32(v * F32(m)))
            zmv = F32(s / F32(nbp))
            X[mask[:, jj] == 0, jj] = zmv
    if k_ew >= 0:
        vim, vip, ivi = [ni - k_ew - 1, ni - 2], [1, k_ew], [0, ni - 1]
    nsw = 0
    E = lambda a: a   # noqa: E731
    for _ in range(ninc_max):
        if not (maskv == 0).any():
            break
        nsw += 1
        d = X.copy()
        mv = maskv
        mc = np.zeros((ni, nj), np.int32)
        I, J = slice(1, ni - 1), slice(1, nj - 1)
        Ip, Im, Jp, Jm = slice(2, ni), slice(0, ni - 2), slice(2, nj), slice(0, nj - 2)
        mc[I, J] = (mv[Ip, J] + mv[I, Jp] + mv[Im, J] + mv[I, Jm]) * (-(mv[I, J] - 1))
        if k_ew >= 0:
            for c in range(2):
                ji, jim, jip = ivi[c], vim[c], vip[c]
                mc[ji, J] = (mv[jip, J] + mv[ji, Jp] + mv[jim, J] + mv[ji, Jm]) * (-(mv[ji, J] - 1))
        else:
            mc[0, J] = (mv[1, J] + mv[0, Jp] + mv[0, Jm]) * (-(mv[0, J] - 1))
            mc[ni - 1, J] = (mv[ni - 1, Jp] + mv[ni - 2, J] + mv[ni - 1, Jm]) * (-(mv[ni - 1, J] - 1))
        mc[I, 0] = (mv[Ip, 0] + mv[I, 1] + mv[Im, 0]) * (-(mv[I, 0] - 1))
        if k_ew >= 0:
            mc[0, 0] = (mv[1, 0] + mv[0, 1] + mv[ni - k_ew - 1, 0]) * (-(mv[0, 0] - 1))
            mc[ni - 1, 0] = (mv[k_ew, 0] + mv[ni - 1, 1] + mv[ni - 1, 0]) * (-(mv[ni - 1, 0] - 1))
        else:
            mc[0, 0] = (mv[1, 0] + mv[0, 1]) * (-(mv[0, 0] - 1))
            mc[ni - 1, 0] = (mv[ni - 1, 1] + mv[ni - 1, 0]) * (-(mv[ni - 1, 0] - 1))
        mc[I, nj - 1] = (mv[Ip, nj - 1] + mv[Im, nj - 1] + mv[I, nj - 2]) * (-(mv[I, nj - 1] - 1))
        if k_ew >= 0:
            mc[0, nj - 1] = (mv[1, nj - 1] + mv[ni - k_ew - 1, nj - 1] + mv[0, nj - 2]) * (-(mv[0, nj - 1] - 1))
            mc[ni - 1, nj - 1] = (mv[k_ew, nj - 1] + mv[ni - 2, nj - 1] + mv[ni - 2, nj - 2]) * (-(mv[ni - 1, nj - 1] - 1))
        else:
            mc[0, nj - 1] = (mv[1, nj - 1] + mv[0, nj - 2]) * (-(mv[0, nj - 1] - 1))
            mc[ni - 1, nj - 1] = (mv[ni - 2, nj - 1] + mv[ni - 2, nj - 2]) * (-(mv[ni - 1, nj - 1] - 1))
        mc = (mc > 0).astype(np.int32)

        def T(i, j):
            return F32(F32(mv[i, j]) * d[i, j])

        def TD(i, j):
            return F32(F32(RIS2 * F32(mv[i, j])) * d[i, j])

        def MD(i, j):
            return F32(RIS2 * F32(mv[i, j]))

        def upd(directs, diags, grouped):
            den = F32(sum(int(mv[i, j]) for i, j in directs))
            if grouped:
                den = F32(den + F32(RIS2 * F32(sum(int(mv[i, j]) for i, j in diags))))
            else:
                for i, j in diags:
                    den = F32(den + MD(i, j))
            num = None
            for i, j in directs:
                t = T(i, j)
                num = t if num is None else F32(num + t)
            for i, j in diags:
                num = F32(num + TD(i, j))
            return F32(F32(F32(1.0) / den) * num)

        for ji, jj in zip(*np.nonzero(mc)):
            per = k_ew >= 0
            if 1 <= jj <= nj - 2:
                if 1 <= ji <= ni - 2:
                    v = upd([(ji + 1, jj), (ji, jj + 1), (ji - 1, jj), (ji, jj - 1)],
                            [(ji + 1, jj + 1), (ji - 1, jj + 1), (ji - 1, jj - 1), (ji + 1, jj - 1)], True)
                elif per:
                    c = 0 if ji == 0 else 1
                    jim, jip = vim[c], vip[c]
                    v = upd([(jip, jj), (ji, jj + 1), (jim, jj), (ji, jj - 1)],
                            [(jip, jj + 1), (jim, jj + 1), (jim, jj - 1), (jip, jj - 1)], True)
                elif ji == 0:
                    v = upd([(1, jj), (0, jj + 1), (0, jj - 1)], [(1, jj + 1), (1, jj - 1)], False)
                else:
                    v = upd([(ji, jj + 1), (ni - 2, jj), (ji, jj - 1)], [(ni - 2, jj + 1), (ni - 2, jj - 1)], False)
            elif jj == nj - 1:
                if 1 <= ji <= ni - 2:
                    v = upd([(ji + 1, jj), (ji - 1, jj), (ji, jj - 1)], [(ji - 1, jj - 1), (ji + 1, jj - 1)], False)
                elif per:
                    c = 0 if ji == 0 else 1
                    jim, jip = vim[c], vip[c]
                    v = upd([(jip, jj), (jim, jj), (ji, jj - 1)], [(jim, jj - 1), (jip, jj - 1)], False)
                elif ji == 0:
                    v = upd([(1, jj), (0, jj - 1)], [(1, jj - 1)], False)
                else:
                    v = upd([(ni - 2, jj), (ji, jj - 1)], [(ni - 2, jj - 1)], False)
            else:
                if 1 <= ji <= ni - 2:
                    v = upd([(ji + 1, jj), (ji, jj + 1), (ji - 1, jj)], [(ji + 1, jj + 1), (ji - 1, jj + 1)], False)
                elif per:
                    c = 0 if ji == 0 else 1
                    jim, jip = vim[c], vip[c]
                    v = upd([(jip, jj), (ji, jj + 1), (jim, jj)], [(jip, jj + 1), (jim, jj + 1)], False)
                elif ji == 0:
                    v = upd([(1, jj), (0, jj + 1)], [(1, jj + 1)], False)
                else:
                    v = upd([(ji, jj + 1), (ni - 2, jj)], [(ni - 2, jj + 1)], False)
            X[ji, jj] = v
        maskv = maskv + mc
        X = np.minimum(np.maximum(X, zmin), zmax)
    if nsmooth > 0:
        X = smoother(k_ew, X, nsmooth, msk=(1 - mask).astype(np.int32))
    if abs(F32(zval_land)) > F32(1.0e-12):
        X[maskv == 0] = F32(zval_land)
    return X, nsw


# ---------------------------------------------------------------------------------------------------------------
# FILL_EXTRA_BANDS in array syntax (src/mod_manip.f90:69-305), iorca = 0 only; arrays (ni, nj)
def _extra_2_east(x1, x2, x3, x4, x5, y1, y2, y3):
    A, B, C, D = x2 - x1, x3 - x2, x4 - x3, x5 - x4
    ALF, BET = y2 - y1, y3 - y2
    with np.errstate(divide="ignore", invalid="ignore"):
        y4 = C * (2 * BET / B - ALF / A) + y3
        y5 = y4 + y4 * D / C + BET * D / B - ALF * D / A - y3 * D / C
    deg = (A == 0.0) | (B == 0.0) | (C == 0.0)
    return np.where(deg, y3, y4), np.where(deg, y3, y5)


def fill_extra_bands(k_ew, XX, YY, XF):
    nx, ny = XX.shape
    nxp4, nyp4 = nx + 4, ny + 4
    XP4 = np.zeros((nxp4, nyp4)); YP4 = np.zeros((nxp4, nyp4)); FP4 = np.zeros((nxp4, nyp4))
    C = (slice(2, nxp4 - 2), slice(2, nyp4 - 2))
    XP4[C], YP4[C], FP4[C] = XX, YY, XF
    R = slice(2, nyp4 - 2)
    if k_ew > -1:
        XP4[0, R] = XX[nx - 2 - k_ew, :] - 360.0
        XP4[1, R] = XX[nx - 1 - k_ew, :] - 360.0
        XP4[nxp4 - 1, R] = XX[1 + k_ew, :] + 360.0
        XP4[nxp4 - 2, R] = XX[k_ew, :] + 360.0
    else:
        XP4[1, R] = XX[1, :] - (XX[2, :] - XX[0, :])
        XP4[0, R] = XX[0, :] - (XX[2, :] - XX[0, :])
        XP4[nxp4 - 2, R] = XX[nx - 2, :] + XX[nx - 1, :] - XX[nx - 3, :]
        XP4[nxp4 - 1, R] = XX[nx - 1, :] + XX[nx - 1, :] - XX[nx - 3, :]
    XP4[:, 1] = XP4[:, 3] - (XP4[:, 4] - XP4[:, 2])
    XP4[:, 0] = XP4[:, 2] - (XP4[:, 4] - XP4[:, 2])
    XP4[:, nyp4 - 2] = XP4[:, nyp4 - 4] + XP4[:, nyp4 - 3] - XP4[:, nyp4 - 5]
    XP4[:, nyp4 - 1] = XP4[:, nyp4 - 3] + XP4[:, nyp4 - 3] - XP4[:, nyp4 - 5]
    Cc = slice(2, nxp4 - 2)
    YP4[Cc, nyp4 - 2] = YY[:, ny - 2] + YY[:, ny - 1] - YY[:, ny - 3]
    YP4[Cc, nyp4 - 1] = YY[:, ny - 1] + YY[:, ny - 1] - YY[:, ny - 3]
    YP4[Cc, 1] = YY[:, 1] - (YY[:, 2] - YY[:, 0])
    YP4[Cc, 0] = YY[:, 0] - (YY[:, 2] - YY[:, 0])
    if k_ew > -1:
        YP4[0, :] = YP4[nx - 2 - k_ew + 2, :]
        YP4[1, :] = YP4[nx - 1 - k_ew + 2, :]
        YP4[nxp4 - 1, :] = YP4[1 + k_ew + 2, :]
        YP4[nxp4 - 2, :] = YP4[k_ew + 2, :]
    else:
        YP4[1, :] = YP4[3, :] - (YP4[4, :] - YP4[2, :])
        YP4[0, :] = YP4[2, :] - (YP4[4, :] - YP4[2, :])
        YP4[nxp4 - 2, :] = YP4[nxp4 - 4, :] + YP4[nxp4 - 3, :] - YP4[nxp4 - 5, :]
        YP4[nxp4 - 1, :] = YP4[nxp4 - 3, :] + YP4[nxp4 - 3, :] - YP4[nxp4 - 5, :]
    if k_ew > -1:
        FP4[0, R] = XF[nx - 2 - k_ew, :]
        FP4[1, R] = XF[nx - 1 - k_ew, :]
        FP4[nxp4 - 1, R] = XF[1 + k_ew, :]
        FP4[nxp4 - 2, R] = XF[k_ew, :]
    else:
        y4, y5 = _extra_2_east(XP4[nxp4 - 5, R], XP4[nxp4 - 4, R], XP4[nxp4 - 3, R], XP4[nxp4 - 2, R], XP4[nxp4 - 1, R],
                               FP4[nxp4 - 5, R], FP4[nxp4 - 4, R], FP4[nxp4 - 3, R])
        FP4[nxp4 - 2, R], FP4[nxp4 - 1, R] = y4, y5
        y2, y1 = _extra_2_east(XP4[4, R], XP4[3, R], XP4[2, R], XP4[1, R], XP4[0, R], FP4[4, R], FP4[3, R], FP4[2, R])
        FP4[1, R], FP4[0, R] = y2, y1
    y4, y5 = _extra_2_east(YP4[:, nyp4 - 5], YP4[:, nyp4 - 4], YP4[:, nyp4 - 3], YP4[:, nyp4 - 2], YP4[:, nyp4 - 1],
                           FP4[:, nyp4 - 5], FP4[:, nyp4 - 4], FP4[:, nyp4 - 3])
    FP4[:, nyp4 - 2], FP4[:, nyp4 - 1] = y4, y5
    y2, y1 = _extra_2_east(YP4[:, 4], YP4[:, 3], YP4[:, 2], YP4[:, 1], YP4[:, 0], FP4[:, 4], FP4[:, 3], FP4[:, 2])
    FP4[:, 1], FP4[:, 0] = y2, y1
    return XP4, YP4, FP4


# ---------------------------------------------------------------------------------------------------
# vertical stage of INTERP_3D: independent numpy (float32, array syntax) / pure-Python-loop transliterations
def akima_1d_3d_np(vz1, xf1, vz2, rfill_val):
    """AKIMA_1D_3D, src/mod_akima_1d.f90:243-389, written with whole-array float32 numpy statements like the
    Fortran source (extended cube + slope cube materialised); xf1 (nk1, ny, nx) -> (nk2, ny, nx)."""
    f4 = np.float32
    vz1 = np.asarray(vz1, f4); vz2 = np.asarray(vz2, f4); xf1 = np.asarray(xf1, f4)
    nk1, nk2 = vz1.size, vz2.size
    nk1e = nk1 + 4
    vz1e = np.zeros(nk1e + 1, f4)                      # 1-based
    xf1e = np.zeros((nk1e + 1,) + xf1.shape[1:], f4)
    vz1e[3:nk1 + 3] = vz1
    xf1e[3:nk1 + 3] = xf1
    vz1e[2] = vz1e[4] - (vz1e[5] - vz1e[3])
    vz1e[1] = vz1e[3] - (vz1e[5] - vz1e[3])
    vz1e[nk1e - 1] = vz1e[nk1e - 3] + vz1e[nk1e - 2] - vz1e[nk1e - 4]
    vz1e[nk1e] = vz1e[nk1e - 2] + vz1e[nk1e - 2] - vz1e[nk1e - 4]

    def extra(xa, xb, xc, xd, xe, Fa, Fb, Fc):     # extra_2_top_3d (x1..x5, Xf1..Xf3) == extra_2_bottom_3d mirrored
        A = f4(xb - xa); B = f4(xc - xb); Cc = f4(xd - xc); D = f4(xe - xd)
        ALF = Fb - Fa; BET = Fc - Fb
        if A == 0 or B == 0 or Cc == 0:
            return Fc.copy(), Fc.copy()
        F4 = Cc * (f4(2) * BET / B - ALF / A) + Fc
        F5 = F4 + F4 * D / Cc + BET * D / B - ALF * D / A - Fc * D / Cc
        return F4, F5
    xf1e[2], xf1e[1] = extra(vz1e[5], vz1e[4], vz1e[3], vz1e[2], vz1e[1], xf1e[5], xf1e[4], xf1e[3])
    xf1e[nk1e - 1], xf1e[nk1e] = extra(vz1e[nk1e - 4], vz1e[nk1e - 3], vz1e[nk1e - 2], vz1e[nk1e - 1], vz1e[nk1e],
                                      xf1e[nk1e - 4], xf1e[nk1e - 3], xf1e[nk1e - 2])
    nz = nk1e
    slope = np.zeros_like(xf1e)
    with np.errstate(all="ignore"):
        for k in range(3, nz - 1):
            m1 = (xf1e[k - 1] - xf1e[k - 2]) / f4(vz1e[k - 1] - vz1e[k - 2])
            m2 = (xf1e[k] - xf1e[k - 1]) / f4(vz1e[k] - vz1e[k - 1])
            m3 = (xf1e[k + 1] - xf1e[k]) / f4(vz1e[k + 1] - vz1e[k])
            m4 = (xf1e[k + 2] - xf1e[k + 1]) / f4(vz1e[k + 2] - vz1e[k + 1])
            gen = (np.abs(m4 - m3) * m2 + np.abs(m2 - m1) * m3) / (np.abs(m4 - m3) + np.abs(m2 - m1))
            slope[k] = np.where((m1 == m2) & (m3 == m4), f4(0.5) * (m2 + m3), gen)
    xf2 = np.zeros((nk2,) + xf1.shape[1:], f4)
    for jk2 in range(1, nk2 + 1):
        z = vz2[jk2 - 1]
        do, fnd, jk1, jp1 = True, False, 1, 0
        if z < vz1[0]:
            xf2[jk2 - 1] = xf1[0]; do = False
        while (not fnd) and do:
            if jk1 == nk1:
                break
            if z > vz1[jk1 - 1] and z <= vz1[jk1]:
                jp1 = jk1 + 2; fnd = True
            else:
                if z == vz1[jk1 - 1]:
                    xf2[jk2 - 1] = xf1[jk1 - 1]; do = False; break
                elif z == vz1[jk1]:
                    xf2[jk2 - 1] = xf1[jk1]; do = False; break
            jk1 += 1
        if jk1 == nk1 and not fnd:
            xf2[jk2 - 1] = xf1[nk1 - 1]; do = False
        if not fnd:
            xf2[jk2 - 1] = f4(rfill_val); do = False
        if do:
            jp2 = jp1 + 1
            a0 = xf1e[jp1]; a1 = slope[jp1]
            dz = f4(1.0) / f4(vz1e[jp2] - vz1e[jp1]); dz2 = f4(dz * dz)
            a2 = (f4(3) * (xf1e[jp2] - xf1e[jp1]) * dz - f4(2) * slope[jp1] - slope[jp2]) * dz
            a3 = (slope[jp1] + slope[jp2] - f4(2) * (xf1e[jp2] - xf1e[jp1]) / f4(vz1e[jp2] - vz1e[jp1])) * dz2
            dz = f4(z - vz1e[jp1]); dz2 = f4(dz * dz)
            xf2[jk2 - 1] = a0 + a1 * dz + a2 * dz2 + a3 * dz2 * dz
    jk2 = 1
    while vz2[jk2 - 1] < vz1[0] and jk2 < nk2:
        xf2[jk2 - 1] = xf1[0]; jk2 += 1
    jk2 = nk2
    while jk2 > 0 and vz2[jk2 - 1] > vz1[nk1 - 1]:
        xf2[jk2 - 1] = xf1[nk1 - 1]; jk2 -= 1
    return xf2


def lbc_lnk_py(nperio, X, cd_type, psgn, pval=None):
    """lbc_lnk_2d_r4 (through REAL(8), src/mod_nemotools.f90:154-190): X (jpj, jpi) float32; returns float32."""
    return _lbc_core(nperio, np.asarray(X, np.float64), cd_type, psgn, pval).astype(np.float32)


def _lbc_core(nperio, X, cd_type, psgn, pval=None):
    """lbc_lnk_2d_r8 + lbc_nfd_2d (src/mod_nemotools.f90:65-417) with explicit sequential Python loops; X (jpj, jpi) f64."""
    P = np.asarray(X, np.float64).T.copy()       # P[i-1, j-1]
    jpi, jpj = P.shape
    zland = 0.0 if pval is None else float(pval)

    def g(i, j):
        return P[i - 1, j - 1]

    def s(i, j, v):
        P[i - 1, j - 1] = v
    if nperio in (1, 4, 6):
        col = P[jpi - 2, :].copy(); P[0, :] = col
        col = P[1, :].copy(); P[jpi - 1, :] = col
    else:
        if cd_type in "TUVW":
            P[0, :] = zland; P[jpi - 1, :] = zland
        elif cd_type == "F":
            P[jpi - 1, :] = zland
    if nperio == 2:
        if cd_type in "TUW":
            P[:, 0] = P[:, 2].copy(); P[:, jpj - 1] = zland
        elif cd_type in "VF":
            P[:, 0] = psgn * P[:, 1]; P[:, jpj - 1] = zland
    elif nperio in (3, 4, 5, 6):
        if cd_type in "TUVWI":
            P[:, 0] = zland
        J, G = jpj, jpi
        if nperio in (3, 4):
            if cd_type in "TW":
                for ji in range(2, G + 1):
                    s(ji, J, psgn * g(G - ji + 2, J - 2))
                s(1, J, psgn * g(3, J - 2))
                for ji in range(G // 2 + 1, G + 1):
                    s(ji, J - 1, psgn * g(G - ji + 2, J - 1))
            elif cd_type == "U":
                for ji in range(1, G):
                    s(ji, J, psgn * g(G - ji + 1, J - 2))
                s(1, J, psgn * g(2, J - 2)); s(G, J, psgn * g(G - 1, J - 2)); s(1, J - 1, psgn * g(G, J - 1))
                for ji in range(G // 2, G):
                    s(ji, J - 1, psgn * g(G - ji + 1, J - 1))
            elif cd_type == "V":
                for jl in (-1, 0):
                    for ji in range(2, G + 1):
                        s(ji, J + jl, psgn * g(G - ji + 2, J - 3 - jl))
                s(1, J, psgn * g(3, J - 3))
            elif cd_type == "F":
                for jl in (-1, 0):
                    for ji in range(1, G):
                        s(ji, J + jl, psgn * g(G - ji + 1, J - 3 - jl))
                s(1, J, psgn * g(2, J - 3)); s(G, J, psgn * g(G - 1, J - 3))
                s(G, J - 1, psgn * g(G - 1, J - 2)); s(1, J - 1, psgn * g(2, J - 2))
        else:
            if cd_type in "TW":
                for ji in range(1, G + 1):
                    s(ji, J, psgn * g(G - ji + 1, J - 1))
            elif cd_type == "U":
                for ji in range(1, G):
                    s(ji, J, psgn * g(G - ji, J - 1))
                s(G, J, psgn * g(1, J - 1))
            elif cd_type == "V":
                for ji in range(1, G + 1):
                    s(ji, J, psgn * g(G - ji + 1, J - 2))
                for ji in range(G // 2 + 1, G + 1):
                    s(ji, J - 1, psgn * g(G - ji + 1, J - 1))
            elif cd_type == "F":
                for ji in range(1, G):
                    s(ji, J, psgn * g(G - ji, J - 2))
                s(G, J, psgn * g(1, J - 2))
                for ji in range(G // 2 + 1, G):
                    s(ji, J - 1, psgn * g(G - ji, J - 1))
    else:
        if cd_type in "TUVW":
            P[:, 0] = zland; P[:, jpj - 1] = zland
        elif cd_type == "F":
            P[:, jpj - 1] = zland
    return P.T.copy()


def interp3d_post_np(X, mask, ixtrpl_bot, vmin, vmax, rmiss, i_orca_trg, c_orca_trg, lmout):
    """post-interpolation block of INTERP_3D (src/mod_interp.f90:447-511), numpy; X (nk, nj, ni)"""
    f4 = np.float32
    X = np.asarray(X, f4).copy()
    nk = X.shape[0]
    if ixtrpl_bot == 1:
        first0 = np.where((mask == 0).any(axis=0), (mask == 0).argmax(axis=0) + 1, 0)    # FINDLOC(mask(ji,jj,:), 0)
        for jj, ji in zip(*np.nonzero((mask[0] == 1) & (first0 > 1))):
            kb = first0[jj, ji]
            X[kb - 1:, jj, ji] = X[kb - 2, jj, ji]
    for k in range(nk):
        Xk = X[k]
        Xk[(Xk > f4(vmax)) & (Xk != f4(rmiss))] = f4(vmax)
        Xk[(Xk < f4(vmin)) & (Xk != f4(rmiss))] = f4(vmin)
        if i_orca_trg > 0:
            X[k] = lbc_lnk_py(i_orca_trg, Xk, c_orca_trg, 1.0)
        if lmout:
            m = mask[k].astype(f4)
            X[k] = X[k] * m + (f4(1.0) - m) * f4(rmiss)
    return X


def angle2_np(nperio, lamt, phit, lamu, phiu, lamv, phiv, lamf, phif):
    """angle2 (src/mod_nemotools.f90:607-823), vectorised numpy (numpy's own libm: compare with a tolerance)."""
    f8 = np.float64
    rpi = f8(np.float32(3.141592653)); rad = rpi / f8(180.0)
    ny, nx = lamt.shape
    outs = [np.zeros((ny, nx)) for _ in range(8)]
    cut = (lamt.max() <= 180.) and (lamt.max() > 175.) and (lamt.min() >= -180.) and (lamt.min() < -175.)
    J = slice(1, ny - 1); I = slice(1, nx)
    Jm = slice(0, ny - 2); Jp = slice(2, ny); Im = slice(0, nx - 1)
    zt, zu, zv, zf = lamt[J, I].copy(), lamu[J, I].copy(), lamv[J, I].copy(), lamf[J, I].copy()
    zvjm1, zfjm1, zfim1, zujp1 = lamv[Jm, I].copy(), lamf[Jm, I].copy(), lamf[J, Im].copy(), lamu[Jp, I].copy()
    if cut:
        near = (np.abs(zt) > 170.) | (np.abs(zu) > 170.) | (np.abs(zv) > 170.) | (np.abs(zf) > 170.)
        for a in (zt, zu, zv, zf, zvjm1, zfjm1, zfim1, zujp1):
            a[near] = np.fmod(360. + a[near], 360.)
    T = lambda ph: np.tan(rpi / 4. - rad * ph / 2.)
    xn = lambda lam, ph: 0. - 2. * np.cos(rad * lam) * T(ph)
    yn = lambda lam, ph: 0. - 2. * np.sin(rad * lam) * T(ph)
    zxnpt, zynpt = xn(zt, phit[J, I]), yn(zt, phit[J, I]); znnpt = zxnpt ** 2 + zynpt ** 2
    zxnpu, zynpu = xn(zu, phiu[J, I]), yn(zu, phiu[J, I]); znnpu = zxnpu ** 2 + zynpu ** 2
    zxnpv, zynpv = xn(zv, phiv[J, I]), yn(zv, phiv[J, I]); znnpv = zxnpv ** 2 + zynpv ** 2
    zxnpf, zynpf = xn(zf, phif[J, I]), yn(zf, phif[J, I]); znnpf = zxnpf ** 2 + zynpf ** 2
    d = lambda la, pa, lb, pb, fn: 2. * fn(rad * la) * T(pa) - 2. * fn(rad * lb) * T(pb)
    eps = f8(np.float32(1.e-14))
    zxvvt, zyvvt = d(zv, phiv[J, I], zvjm1, phiv[Jm, I], np.cos), d(zv, phiv[J, I], zvjm1, phiv[Jm, I], np.sin)
    znvvt = np.maximum(np.sqrt(znnpt * (zxvvt ** 2 + zyvvt ** 2)), eps)
    zxffu, zyffu = d(zf, phif[J, I], zfjm1, phif[Jm, I], np.cos), d(zf, phif[J, I], zfjm1, phif[Jm, I], np.sin)
    znffu = np.maximum(np.sqrt(znnpu * (zxffu ** 2 + zyffu ** 2)), eps)
    zxffv, zyffv = d(zf, phif[J, I], zfim1, phif[J, Im], np.cos), d(zf, phif[J, I], zfim1, phif[J, Im], np.sin)
    znffv = np.maximum(np.sqrt(znnpv * (zxffv ** 2 + zyffv ** 2)), eps)
    zxuuf, zyuuf = d(zujp1, phiu[Jp, I], zu, phiu[J, I], np.cos), d(zujp1, phiu[Jp, I], zu, phiu[J, I], np.sin)
    znuuf = np.maximum(np.sqrt(znnpf * (zxuuf ** 2 + zyuuf ** 2)), eps)
    cost, sint, cosu, sinu, cosv, sinv, cosf, sinf = outs
    sint[J, I] = (zxnpt * zyvvt - zynpt * zxvvt) / znvvt; cost[J, I] = (zxnpt * zxvvt + zynpt * zyvvt) / znvvt
    sinu[J, I] = (zxnpu * zyffu - zynpu * zxffu) / znffu; cosu[J, I] = (zxnpu * zxffu + zynpu * zyffu) / znffu
    sinf[J, I] = (zxnpf * zyuuf - zynpf * zxuuf) / znuuf; cosf[J, I] = (zxnpf * zxuuf + zynpf * zyuuf) / znuuf
    sinv[J, I] = (zxnpv * zxffv + zynpv * zyffv) / znffv; cosv[J, I] = -(zxnpv * zyffv - zynpv * zxffv) / znffv
    if nperio > 0:
        for a, cd in zip(outs, "TTUUVVFF"):
            a[:] = lbc_lnk_r8_py(nperio, a, cd, -1.0)
    else:
        def ex1(xa, xb, xc, xd, ya, yb, yc):
            A, B, Cc = xb - xa, xc - xb, xd - xc
            with np.errstate(all="ignore"):
                r = Cc * (2 * (yc - yb) / B - (yb - ya) / A) + yc
            return np.where((A == 0) | (B == 0) | (Cc == 0), yc, r)
        PH = (phit, phit, phiu, phiu, phiv, phiv, phif, phif); LA = (lamt, lamt, lamu, lamu, lamv, lamv, lamf, lamf)
        for a, P in zip(outs, PH):
            a[ny - 1, :] = ex1(P[ny - 4], P[ny - 3], P[ny - 2], P[ny - 1], a[ny - 4], a[ny - 3], a[ny - 2])
        for a, P in zip(outs, PH):
            a[0, :] = ex1(P[3], P[2], P[1], P[0], a[3], a[2], a[1])
        for a, L in zip(outs, LA):
            a[:, 0] = ex1(L[:, 3], L[:, 2], L[:, 1], L[:, 0], a[:, 3], a[:, 2], a[:, 1])
    return tuple(outs)


def lbc_lnk_r8_py(nperio, X, cd_type, psgn, pval=None):
    """lbc_lnk_2d_r8 on a float64 array: same loops as lbc_lnk_py without the REAL(4) round trip."""
    return _lbc_core(nperio, np.asarray(X, np.float64), cd_type, psgn, pval)


def akima_1d_1d_py(vz1, vf1, vz2, rmask_val=None, l_sp=False, l_bp=False):
    """AKIMA_1D_1D, src/mod_akima_1d.f90:26-232 (+ slopes_1d :392-455, extra_2_west_r4 / extra_2_east_r4 of mod_manip.f90),
    one column, Python loops on numpy scalars: 1-based arrays with a dummy element 0.  Returns (vf2 f32, status) with the
    status convention of the oracle (0 ok, 1 all missing, -1 / -2 undefined in the reference)."""
    f4, f8 = np.float32, np.float64
    vz1 = np.concatenate([[0.0], np.asarray(vz1, f8)]); vf1 = np.concatenate([[0.0], np.asarray(vf1, f8)])
    vz2 = np.concatenate([[0.0], np.asarray(vz2, f8)])
    nk1, nk2 = vz1.size - 1, vz2.size - 1
    vf2 = np.full(nk2 + 1, f4(-6666.0), f4)
    nz1, k1_1, k1_n, nmissv = nk1, 1, nk1, 0
    if rmask_val is not None:
        imiss1 = np.zeros(nk1 + 1, np.int32)
        imiss1[1:][vf1[1:] == f8(rmask_val)] = 1
        nmissv = int(imiss1.sum())
    if not nmissv < nk1:
        return vf2[1:], 1
    if nmissv > 0:
        k1_1 = int(np.nonzero(imiss1[1:] == 0)[0][0]) + 1
        w = np.nonzero(imiss1[k1_1:] == 1)[0]
        k1_n = (int(w[0]) + 1 if w.size else 0) + k1_1 - 2
        nz1 = nk1 - nmissv
        if k1_n - k1_1 + 1 != nz1:
            return vf2[1:], -1
    if nz1 < 3:
        return vf2[1:], -2
    nk1e = nz1 + 4
    ze = np.zeros(nk1e + 1, f4); fe = np.zeros(nk1e + 1, f4); sl = np.zeros(nk1e + 1, f4)
    ze[3:nz1 + 3] = vz1[k1_1:k1_n + 1].astype(f4)
    fe[3:nz1 + 3] = vf1[k1_1:k1_n + 1].astype(f4)
    ze[2] = ze[4] - (ze[5] - ze[3])
    ze[1] = ze[3] - (ze[5] - ze[3])
    ze[nk1e - 1] = ze[nk1e - 3] + ze[nk1e - 2] - ze[nk1e - 4]
    ze[nk1e] = ze[nk1e - 2] + ze[nk1e - 2] - ze[nk1e - 4]
    two = f4(2)

    def extra(xa, xb, xc, xd, xe, ya, yb, yc):   # same body for east (1..5) and west (5..1)
        A = xb - xa; B = xc - xb; Cc = xd - xc; D = xe - xd
        ALF = yb - ya; BET = yc - yb
        if A == 0 or B == 0 or Cc == 0:
            return yc, yc
        y4 = Cc * (two * BET / B - ALF / A) + yc
        y5 = y4 + y4 * D / Cc + BET * D / B - ALF * D / A - yc * D / Cc
        return y4, y5
    with np.errstate(all="ignore"):
        fe[2], fe[1] = extra(ze[5], ze[4], ze[3], ze[2], ze[1], fe[5], fe[4], fe[3])
        fe[nk1e - 1], fe[nk1e] = extra(ze[nk1e - 4], ze[nk1e - 3], ze[nk1e - 2], ze[nk1e - 1], ze[nk1e], fe[nk1e - 4], fe[nk1e - 3], fe[nk1e - 2])
        nz = nk1e
        for k in range(3, nz - 1):
            m1 = (fe[k - 1] - fe[k - 2]) / (ze[k - 1] - ze[k - 2])
            m2 = (fe[k] - fe[k - 1]) / (ze[k] - ze[k - 1])
            m3 = (fe[k + 1] - fe[k]) / (ze[k + 1] - ze[k])
            m4 = (fe[k + 2] - fe[k + 1]) / (ze[k + 2] - ze[k + 1])
            if m1 == m2 and m3 == m4:
                sl[k] = f4(0.5) * (m2 + m3)
            else:
                sl[k] = (abs(m4 - m3) * m2 + abs(m2 - m1) * m3) / (abs(m4 - m3) + abs(m2 - m1))
        for jk2 in range(1, nk2 + 1):
            do, fnd = True, False
            jk1 = k1_1
            if vz2[jk2] < vz1[k1_1] and l_sp:
                vf2[jk2] = f4(vf1[k1_1]); do = False; fnd = True
            while (not fnd) and do:
                if jk1 == k1_n:
                    break
                if vz2[jk2] > vz1[jk1] and vz2[jk2] <= vz1[jk1 + 1]:
                    jp1 = jk1 + 2 - k1_1 + 1; jp2 = jp1 + 1; fnd = True
                elif vz2[jk2] == vz1[jk1]:
                    vf2[jk2] = f4(vf1[jk1]); do = False
                    break
                elif vz2[jk2] == vz1[jk1 + 1]:
                    vf2[jk2] = f4(vf1[jk1 + 1]); do = False
                    break
                jk1 += 1
            if not fnd:
                do = False
                if jk1 == k1_n and l_bp:
                    vf2[jk2] = f4(vf1[k1_n])
            if do:
                a0 = f8(fe[jp1]); a1 = f8(sl[jp1])
                dz = f8(f4(1.0) / (ze[jp2] - ze[jp1])); dz2 = dz * dz
                a2 = (f8(f4(3) * (fe[jp2] - fe[jp1])) * dz - f8(two * sl[jp1]) - f8(sl[jp2])) * dz
                a3 = f8(sl[jp1] + sl[jp2] - two * (fe[jp2] - fe[jp1]) / (ze[jp2] - ze[jp1])) * dz2
                dz = vz2[jk2] - f8(ze[jp1]); dz2 = dz * dz
                vf2[jk2] = f4(a0 + a1 * dz + a2 * dz2 + a3 * dz2 * dz)
    return vf2[1:], 0


def interp3d_sigma_py(depth_src, data3d_tmp, depth_trg, data3d_trg, mask_trg=None, lmout=False, src_sigma=False, trg_sigma=False):
    """generic-coordinate branch of INTERP_3D, src/mod_interp.f90:374-438 (inputs left untouched); cubes (nk, nj, ni)"""
    f4 = np.float32
    depth_src = np.asarray(depth_src, f4); data3d_tmp = np.asarray(data3d_tmp, f4); depth_trg = np.asarray(depth_trg, f4)
    out = np.array(data3d_trg, f4)
    nk_trg, nj, ni = depth_trg.shape
    for jj in range(nj):
        for ji in range(ni):
            zs = depth_src[:, jj, ji]; fs = data3d_tmp[:, jj, ji]; zt = depth_trg[:, jj, ji]
            if src_sigma:
                zs = zs[::-1]; fs = fs[::-1]
            if trg_sigma:
                zt = zt[::-1]
            zmax_src, zmax_trg = zs.max(), zt.max()
            if zmax_trg > zmax_src:
                jklast = 1
                while jklast < nk_trg:
                    if zt[jklast] > zmax_src:
                        break
                    jklast += 1
            else:
                jklast = nk_trg
            if (not lmout) or mask_trg[0, jj, ji] == 1:
                nlev = int((mask_trg[:, jj, ji] == 1).sum()) if lmout else jklast
                t = out[:, jj, ji].copy()
                v, _ = akima_1d_1d_py(zs.astype(np.float64), fs.astype(np.float64), zt[:nlev].astype(np.float64))
                t[:nlev] = v
                if 0 < jklast < nk_trg:
                    t[jklast:] = t[jklast - 1]
                if trg_sigma:
                    t = t[::-1]
                out[:, jj, ji] = t
    return out


def interp_bl_np(k_ew, iP, jP, iqd, xa, xb, Z):
    """INTERP_BL, src/mod_bilin_2d.f90:275-361, on Z (nj, ni) float32 (1-based indices translated at the reads)"""
    f4 = np.float32
    nyi, nxi = Z.shape
    im1, ip1 = iP - 1, iP + 1
    if im1 == 0 and k_ew >= 0:
        im1 = nxi - k_ew
    if ip1 == nxi + 1 and k_ew >= 0:
        ip1 = 1 + k_ew
    quad = {1: ((iP, jP), (ip1, jP), (ip1, jP + 1), (iP, jP + 1)), 2: ((iP, jP), (iP, jP - 1), (ip1, jP - 1), (ip1, jP)),
            3: ((iP, jP), (im1, jP), (im1, jP - 1), (iP, jP - 1)), 4: ((iP, jP), (iP, jP + 1), (im1, jP + 1), (im1, jP))}[iqd]
    w = [f4((1.0 - xa) * (1.0 - xb)), f4(xa * (1.0 - xb)), f4(xa * xb), f4((1.0 - xa) * xb)]
    wup = w[0] + w[1] + w[2] + w[3]
    if wup == 0:
        return f4(-9998.0)
    if any(i < 1 for i, _ in quad):
        return f4(-9997.0)
    if any(j < 1 for _, j in quad):
        return f4(-9996.0)
    if any(j > nyi for _, j in quad):
        return f4(-9995.0)
    if any(i > nxi for i, _ in quad):
        return f4(-9994.0)
    z = [Z[j - 1, i - 1] for i, j in quad]
    return (z[0] * w[0] + z[1] * w[1] + z[2] * w[2] + z[3] * w[3]) / wup


def ground_track_loop_np(fields, vt_mod, imask, vtf, metrics, ab, F_gt_f, rmissval=-9999.0):
    """the time loop of interp_to_ground_track.f90:674-787, literally: the whole field xvar is rebuilt for every track point.
    fields (Ntm, nj, ni) f32 (what GETVAR_2D would read), imask (nj, ni) i8, metrics (3, Ntf) i32, ab (2, Ntf) f64."""
    f4, f8 = np.float32, np.float64
    Ntm, nj, ni = fields.shape
    Ntf = vtf.size
    t_min_m, t_max_m = vt_mod.min(), vt_mod.max()
    mod_np = np.full(Ntf, rmissval, f8); obs = np.full(Ntf, rmissval, f8); mod = np.full(Ntf, rmissval, f8); Fmask = np.zeros(Ntf, np.int8)
    jtm_1_o = jtm_2_o = -100
    jt_s = 1
    xvar1 = xvar2 = xdum_r4 = None
    M = lambda i, j: int(imask[j - 1, i - 1])
    for jtf in range(Ntf):
        rt = f8(vtf[jtf])
        if rt >= t_min_m and rt < t_max_m:
            jt = jt_s
            while jt <= Ntm - 1:
                if rt >= vt_mod[jt - 1] and rt < vt_mod[jt]:
                    break
                jt += 1
            jtm_1, jtm_2 = jt, jt + 1
            if jtm_1 > jtm_1_o and jtm_2 > jtm_2_o:
                if jtm_1_o == -100 or jtm_1 > jtm_2_o:
                    xvar1 = fields[jtm_1 - 1].copy()
                else:
                    xvar1 = xvar2.copy()
                xvar2 = fields[jtm_2 - 1].copy()
                xdum_r4 = ((xvar2 - xvar1).astype(f8) / (vt_mod[jtm_2 - 1] - vt_mod[jtm_1 - 1])).astype(f4)
            xvar = (xvar1.astype(f8) + xdum_r4.astype(f8) * (rt - vt_mod[jtm_1 - 1])).astype(f4)
            iP, jP, iq = int(metrics[0, jtf]), int(metrics[1, jtf]), int(metrics[2, jtf])
            alpha, beta = f8(ab[0, jtf]), f8(ab[1, jtf])
            if iP > 0 and jP > 0:
                if M(iP, jP) == 1:
                    r_obs = F_gt_f[jtf]
                    if r_obs > -20.0 and r_obs < 20.0:
                        ip1, jp1, im1, jm1 = min(iP + 1, ni), min(jP + 1, nj), max(iP - 1, 1), max(jP - 1, 1)
                        idot = M(ip1, jP) + M(ip1, jp1) + M(iP, jp1) + M(im1, jp1) + M(im1, jP) + M(im1, jm1) + M(iP, jm1) + M(ip1, jm1)
                        if idot == 8:
                            mod_np[jtf] = f8(xvar[jP - 1, iP - 1])
                            mod[jtf] = f8(interp_bl_np(-1, iP, jP, iq, alpha, beta, xvar))
                            obs[jtf] = r_obs
                            Fmask[jtf] = 1
            jtm_1_o, jtm_2_o, jt_s = jtm_1, jtm_2, jtm_1
    mod[Fmask == 0] = rmissval; mod_np[Fmask == 0] = rmissval; obs[Fmask == 0] = rmissval
    mod[mod < f8(f4(-9990.0))] = rmissval
    return mod, mod_np, obs, Fmask


# ---------------------------------------------------------------------------------------------------------------
# FIND_NEAREST_POINT prelude (src/mod_manip.f90:884-947), IS_POLAR_PROJ (:1843-1879) and FIND_NEAREST_TWISTED (:1073-1440):
# a second, independent transcription of the search on curvilinear sources (written from the Fortran text with numpy
# array syntax / Python loops, glibc libm through `math`).  Arrays are Fortran-shaped [i, j], 0-based inside, indices
# returned 1-based.  Undefined behaviour of the reference is DEFINED as SURVEY Appendix A says (Q3: ji_s = jj_s = 0 before
# the first luck window; Q4: J_VLAT_S(nsplit+1,1) aliases J_VLAT_S(1,2); Q5: an out-of-range probe means "not polar").
def l_is_grid_regular(Xlon, Xlat):
    nx, ny = Xlon.shape
    for jj in range(1, ny):
        s = 0.0
        for v in np.abs(Xlon[:, jj] - Xlon[:, 0]):      # SUM() in index order
            s += float(v)
        if s > float(F32(1.0e-12)):
            return False
    for ji in range(nx):
        s = 0.0
        for v in np.abs(Xlat[ji, :] - Xlat[0, :]):
            s += float(v)
        if s > float(F32(1.0e-12)):
            return False
    return True


def _minloc_dim2(Y, mask, largest=False):
    """MINLOC / MAXLOC(Y, mask=mask, dim=2): per row index i, the 1-based first j of the extreme value, 0 if the mask is empty"""
    nx, ny = Y.shape
    out = np.zeros(nx, np.int64)
    for i in range(nx):
        best = None
        for j in range(ny):
            if not mask[i, j]:
                continue
            if best is None or (Y[i, j] > best if largest else Y[i, j] < best):
                best, out[i] = Y[i, j], j + 1
    return out


def find_nearest_prelude(Xtrg, Ytrg, Xsrc, Ysrc):
    """(l_is_reg_s, l_is_reg_t, j_strt_t, j_stop_t, jlat_icr), :884-947"""
    nx_t, ny_t = Xtrg.shape
    reg_s, reg_t = l_is_grid_regular(Xsrc, Ysrc), l_is_grid_regular(Xtrg, Ytrg)
    y_min_s, y_max_s = Ysrc.min(), Ysrc.max()
    j_strt, j_stop, icr = 1, ny_t, 1
    rmin = float(F32(-100.0))
    if nx_t > 2:
        z = np.zeros((nx_t, ny_t))
        z[:, 1:] = Ytrg[:, 1:] - Ytrg[:, :-1]
        rtmp = 0.0
        for v in z[:, ny_t // 2 - 1]:                    # SUM(ztmp_t(:,ny_t/2)), index order
            rtmp += float(v)
        z = math.copysign(1.0, rtmp) * z
        rmin = z[:, 1:].min()
    if rmin > float(F32(-1.0e-12)) or reg_t:
        i1 = _minloc_dim2(Ytrg, Ytrg >= y_min_s)
        j_strt = int(i1[i1 > 0].min())
        j_stop = int(_minloc_dim2(Ytrg, Ytrg <= y_max_s, largest=True).max())
        if j_strt > j_stop:
            icr = -1
    return reg_s, reg_t, j_strt, j_stop, icr


def is_polar_proj(XX, YY):
    nx, ny = YY.shape
    if not (YY > 88.0).any():
        return False
    d = np.abs(YY - 90.0)
    k = int(np.argmin(d.T.reshape(-1)))                  # MINLOC: first minimum in column-major (i fastest) order
    jNP, iNP = k // nx, k % nx                           # 0-based
    if jNP - 3 < 0 or jNP + 3 > ny - 1:
        return False                                     # Q5
    lipp = (YY[iNP, jNP - 3] < YY[iNP, jNP]) and (YY[iNP, jNP + 3] < YY[iNP, jNP])
    if lipp:
        a = abs(XX[iNP, jNP - 3] - XX[iNP, jNP + 3])
        lipp = (a > 100.0) and (a < 220.0)
    return bool(lipp)


def _distance_block(rlon, rlat, Xs, Ys, i1, i2, j1, j2):
    """DISTANCE_2D (:1509-1573) on the section (i1:i2, j1:j2) (1-based, inclusive): array [i2-i1+1, j2-j1+1]"""
    zconv = math.acos(-1.0) / 180.0
    zlatar, zlonar = rlat * zconv, rlon * zconv
    zux, zuy, zuz = math.cos(zlonar) * math.cos(zlatar), math.sin(zlonar) * math.cos(zlatar), math.sin(zlatar)
    out = np.empty((i2 - i1 + 1, j2 - j1 + 1))
    for jj in range(j1 - 1, j2):
        for ji in range(i1 - 1, i2):
            zlatbr, zlonbr = Ys[ji, jj] * zconv, Xs[ji, jj] * zconv
            zvx, zvy, zvz = math.cos(zlonbr) * math.cos(zlatbr), math.sin(zlonbr) * math.cos(zlatbr), math.sin(zlatbr)
            zpds = zux * zvx + zuy * zvy + zuz * zvz
            out[ji - i1 + 1, jj - j1 + 1] = ZR * math.acos(min(zpds, 1.0))
    return out


def find_nearest_twisted(Xtrg, Ytrg, Xsrc, Ysrc, mask_t, j_strt_t, j_stop_t, jlat_icr):
    """returns (JIpos, JJpos, mask_t) Fortran-shaped, or raises RuntimeError where the reference STOPs"""
    nx_s, ny_s = Xsrc.shape
    nx_t, ny_t = Xtrg.shape
    mask_t = mask_t.copy()
    y_min_s, y_max_s, y_min_t, y_max_t = Ysrc.min(), Ysrc.max(), Ytrg.min(), Ytrg.max()
    if (y_min_t >= y_max_s) or (y_max_t <= y_min_s):
        raise RuntimeError("Target and source latitudes do not overlap!")
    JI = np.full((nx_t, ny_t), -1, np.int32)
    JJ = np.full((nx_t, ny_t), -1, np.int32)
    e1 = np.full((nx_s, ny_s), float(F32(40000.0)))
    e2 = np.full((nx_s, ny_s), float(F32(40000.0)))
    th = float(F32(1000.0))
    for jj in range(ny_s):
        for ji in range(nx_s - 1):
            e1[ji, jj] = distance(Xsrc[ji, jj], Xsrc[ji + 1, jj], Ysrc[ji, jj], Ysrc[ji + 1, jj]) * th
    for jj in range(ny_s - 1):
        for ji in range(nx_s):
            e2[ji, jj] = distance(Xsrc[ji, jj], Xsrc[ji, jj + 1], Ysrc[ji, jj], Ysrc[ji, jj + 1]) * th
    if nx_s > 1:
        e1[nx_s - 1, :] = e1[nx_s - 2, :]
    if ny_s > 1:
        e2[:, ny_s - 1] = e2[:, ny_s - 2]
    lStupido = is_polar_proj(Xsrc, Ysrc) or is_polar_proj(Xtrg, Ytrg)
    nsplit = 0
    if not lStupido:
        y_max_bnd = y_max_bnd0 = min(y_max_t, y_max_s)
        y_min_bnd = y_min_bnd0 = max(y_min_t, y_min_s)
        y_max_bnd = min(float(int(y_max_bnd + 2.0)), 90.0)               # INT() truncates toward zero
        y_min_bnd = max(float(int(y_min_bnd - 2.0)), -90.0)
        nint = lambda x: int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)   # NINT: half away from zero
        y_max_bnd = min(nint(y_max_bnd / 0.5) * 0.5, 90.0)
        y_min_bnd = max(nint(y_min_bnd / 0.5) * 0.5, -90.0)
        y_max_bnd = max(y_max_bnd, y_max_bnd0)
        y_min_bnd = min(y_min_bnd, y_min_bnd0)
        nsplit = max(int(float(F32(40.0)) * (y_max_bnd - y_min_bnd) / 180.0), 1)
        dy = (y_max_bnd - y_min_bnd) / float(F32(nsplit))
        VB = [y_min_bnd + float(F32(jl)) * dy for jl in range(nsplit + 1)]
        if (y_min_bnd > y_min_bnd0) or (y_max_bnd < y_max_bnd0):
            raise RuntimeError("Bounds for latitude for VLAT_SPLIT_BOUNDS are bad!")
        JV = np.zeros((nsplit, 2), np.int64)
        for jl in range(nsplit):
            rlat_1, rlat_2 = VB[jl], VB[jl + 1]
            jmax_band = int(_minloc_dim2(Ysrc, Ysrc <= rlat_2, largest=True).max())
            i1dum = _minloc_dim2(Ysrc, Ysrc > rlat_1)
            jmin_band = int(i1dum[i1dum > 0].min())
            JV[jl, 0] = max(jmin_band - 1, 1)
            JV[jl, 1] = min(jmax_band + 1, ny_s)
        for jl in range(nsplit):
            if JV[jl, 1] <= JV[jl, 0]:
                raise RuntimeError("jj_stop > jj_start !")
        JVflat = JV.T.reshape(-1)                                         # column-major storage of J_VLAT_S(nsplit,2)
    frac_emax, sqrt2 = float(F32(0.51)), float(np.sqrt(F32(2.0)))
    rlat_old = float(F32(-9999.0))
    ji_s = jj_s = 0                                                       # Q3
    jlat = 1
    for jj_t in range(j_strt_t, j_stop_t + (1 if jlat_icr > 0 else -1), jlat_icr):
        for ji_t in range(1, nx_t + 1):
            if mask_t[ji_t - 1, jj_t - 1] != 1:
                continue
            rlon, rlat = Xtrg[ji_t - 1, jj_t - 1], Ytrg[ji_t - 1, jj_t - 1]
            if not lStupido and rlat != rlat_old:
                jlat = nsplit + 1                                         # value of the DO variable when the loop runs out
                for jl in range(1, nsplit + 1):
                    if rlat == VB[jl - 1] or (rlat > VB[jl - 1] and rlat <= VB[jl]):
                        jlat = jl
                        break
            lagain, niter = True, -1
            if rlat > 60.0:
                niter = 0
            while lagain:
                if niter == -1:
                    i1s, i2s = max(ji_s - 5, 1), min(ji_s + 5, nx_s)
                    j1s, j2s = max(jj_s - 5, 1), min(jj_s + 5, ny_s)
                else:
                    i1s, i2s = 1, nx_s
                    if lStupido:
                        j1s, j2s = 1, ny_s
                    else:
                        j1s = int(JVflat[max(jlat - niter, 1) - 1])       # J_VLAT_S(.,1); (nsplit+1,1) aliases (1,2): Q4
                        j2s = int(JVflat[nsplit + min(jlat + niter, nsplit) - 1])
                D = _distance_block(rlon, rlat, Xsrc, Ysrc, i1s, i2s, j1s, j2s)
                k = int(np.argmin(D.T.reshape(-1)))                       # MINLOC of the section, column-major first minimum
                ni_blk = i2s - i1s + 1
                ji_s, jj_s = k % ni_blk + i1s, k // ni_blk + j1s
                emax = max(e1[ji_s - 1, jj_s - 1], e2[ji_s - 1, jj_s - 1]) / th * sqrt2
                if D[ji_s - i1s, jj_s - j1s] <= frac_emax * emax:
                    lagain = False
                    JI[ji_t - 1, jj_t - 1], JJ[ji_t - 1, jj_t - 1] = ji_s, jj_s
                else:
                    if niter == 0:
                        mlon = (Xsrc > max(rlon - 0.5, 0.0)) & (Xsrc < min(rlon + 0.5, 360.0))
                        mlat = (Ysrc > max(rlat - 0.5, -90.0)) & (Ysrc < min(rlat + 0.5, 90.0))
                        if (mlon & mlat).sum() == 0:
                            lagain = False
                    if niter > nsplit // 3:
                        lagain = False
                niter += 1
            rlat_old = rlat
    mask_t[JI == -1] = -1
    mask_t[JJ == -1] = -2
    return JI, JJ, mask_t


# ---------------------------------------------------------------------------------------------------------------
# AKIMA_2D (src/mod_akima_2d.f90:59-239) with SLOPES_AKIMA (:483-664), build_pol (:249-441), pol_val (:445-480) and
# find_nearest_akima (:670-745): second transcription, whole-array numpy for the slopes and the 16 coefficients of every
# cell (the reference's own data flow -- the oracle and the GPU rebuild the coefficients per target instead), Python
# loops for the index walk.  No libm: IEEE + - * / only, so the comparison with the oracle is exact in either libm mode.
def _guard(rx):
    return np.copysign(1.0, rx) * np.maximum(np.abs(rx), REPS)


def slopes_akima(k_ew, XX, XY, XF):
    nx, ny = XF.shape
    ZX, ZY, ZF = fill_extra_bands(k_ew, XX, XY, XF)
    I = lambda o: slice(o, o + nx)       # ji + o (0-based o: ji -> 0, ip1 -> 1, ip2 -> 2, ji+3 -> 3, ji+4 -> 4)
    J = lambda o: slice(o, o + ny)
    with np.errstate(divide="ignore", invalid="ignore"):
        def dq(A, Z, ia, ja, ib, jb):    # (Z(b) - Z(a)) / guard(A(b) - A(a))
            return (Z[I(ib), J(jb)] - Z[I(ia), J(ja)]) / _guard(A[I(ib), J(jb)] - A[I(ia), J(ja)])
        m1, m2, m3, m4 = dq(ZX, ZF, 0, 2, 1, 2), dq(ZX, ZF, 1, 2, 2, 2), dq(ZX, ZF, 2, 2, 3, 2), dq(ZX, ZF, 3, 2, 4, 2)
        eqx = (m1 == m2) & (m3 == m4)
        Wx2 = np.where(eqx, 0.0, np.abs(m4 - m3)); Wx3 = np.where(eqx, 0.0, np.abs(m2 - m1))
        dFdX = np.where(eqx, 0.5 * (m2 + m3), (Wx2 * m2 + Wx3 * m3) / (Wx2 + Wx3))
        m1, m2, m3, m4 = dq(ZY, ZF, 2, 0, 2, 1), dq(ZY, ZF, 2, 1, 2, 2), dq(ZY, ZF, 2, 2, 2, 3), dq(ZY, ZF, 2, 3, 2, 4)
        eqy = (m1 == m2) & (m3 == m4)
        Wy2 = np.where(eqy, 0.0, np.abs(m4 - m3)); Wy3 = np.where(eqy, 0.0, np.abs(m2 - m1))
        dFdY = np.where(eqy, 0.5 * (m2 + m3), (Wy2 * m2 + Wy3 * m3) / (Wy2 + Wy3))
        # cross derivative: m2, m3 are now the y-direction quotients (the scalars were overwritten, :570-577)
        d22, d23 = dq(ZY, ZF, 1, 1, 1, 2), dq(ZY, ZF, 1, 2, 1, 3)
        d42, d43 = dq(ZY, ZF, 3, 1, 3, 2), dq(ZY, ZF, 3, 2, 3, 3)
        e22 = (m2 - d22) / _guard(ZX[I(2), J(1)] - ZX[I(1), J(1)])
        e23 = (m3 - d23) / _guard(ZX[I(2), J(2)] - ZX[I(1), J(2)])
        e32 = (d42 - m2) / _guard(ZX[I(3), J(2)] - ZX[I(2), J(2)])
        e33 = (d43 - m3) / _guard(ZX[I(3), J(2)] - ZX[I(2), J(2)])
        zx = (Wx2 == 0) & (Wx3 == 0)
        zy = (Wy2 == 0) & (Wy3 == 0)
        Wx2 = np.where(zx, 1.0, Wx2); Wx3 = np.where(zx, 1.0, Wx3)
        Wy2 = np.where(zy, 1.0, Wy2); Wy3 = np.where(zy, 1.0, Wy3)
        d2 = (Wx2 * (Wy2 * e22 + Wy3 * e23) + Wx3 * (Wy2 * e32 + Wy3 * e33)) / ((Wx2 + Wx3) * (Wy2 + Wy3))
    return dFdX, dFdY, d2


def build_pol(ZX, ZY, ZF, sx, sy, sxy):
    """XPLNM(nx-1, ny-1, 16) for every cell, array syntax"""
    a = (slice(0, -1), slice(0, -1)); b = (slice(1, None), slice(0, -1)); c = (slice(1, None), slice(1, None)); d = (slice(0, -1), slice(1, None))
    x = ZX[b] - ZX[a]
    y = ZY[d] - ZY[a]
    x2 = x * x; x3 = x2 * x
    y2 = y * y; y3 = y2 * y
    xy = x * y
    b1, b2, b3, b4 = ZF[a], ZF[b], ZF[c], ZF[d]
    b5, b6, b7, b8 = sx[a], sx[b], sx[c], sx[d]
    b9, b10, b11, b12 = sy[a], sy[b], sy[c], sy[d]
    b13, b14, b15, b16 = sxy[a], sxy[b], sxy[c], sxy[d]
    V = np.zeros(x.shape + (16,))
    V[..., 10], V[..., 11], V[..., 14], V[..., 15] = b13, b5, b9, b1
    b5, b6, b7, b8 = x * b5, x * b6, x * b7, x * b8
    b9, b10, b11, b12 = y * b9, y * b10, y * b11, y * b12
    b13, b14, b15, b16 = xy * b13, xy * b14, xy * b15, xy * b16
    c1, c2, c3 = b1 - b2, b3 - b4, b5 + b6
    c4, c5, c6 = b7 + b8, b9 - b10, b11 - b12
    c7, c8, c9 = b13 + b14, b15 + b16, 2 * b5 + b6
    c10, c11, c12 = b7 + 2 * b8, 2 * b13 + b14, b15 + 2 * b16
    c13, c14, c15 = b5 - b8, b1 - b4, b13 + b16
    c16, c17, c18 = 2 * b13 + b16, b9 + b12, 2 * b9 + b12
    d1, d2, d3 = c1 + c2, c3 - c4, c5 - c6
    d4, d5, d6 = c7 + c8, c9 - c10, 2 * c5 - c6
    d7, d8, d9 = 2 * c7 + c8, c11 + c12, 2 * c11 + c12
    f1, f2, f3 = 2 * d1 + d2, 2 * d3 + d4, 2 * d6 + d7
    f4, f5, f6 = 3 * d1 + d5, 3 * d3 + d8, 3 * d6 + d9
    with np.errstate(divide="ignore", invalid="ignore"):
        V[..., 0] = (2 * f1 + f2) / (x3 * y3); V[..., 1] = (-(3 * f1 + f3)) / (x3 * y2); V[..., 2] = (2 * c5 + c7) / (x3 * y)
        V[..., 3] = (2 * c1 + c3) / x3; V[..., 4] = (-(2 * f4 + f5)) / (x2 * y3); V[..., 5] = (3 * f4 + f6) / (x2 * y2)
        V[..., 6] = (-(3 * c5 + c11)) / (x2 * y); V[..., 7] = (-(3 * c1 + c9)) / x2; V[..., 8] = (2 * c13 + c15) / (x * y3)
        V[..., 9] = (-(3 * c13 + c16)) / (x * y2); V[..., 12] = (2 * c14 + c17) / y3; V[..., 13] = (-(3 * c14 + c18)) / y2
    return V


def pol_val(x, y, V):
    p1 = ((V[0] * y + V[1]) * y + V[2]) * y + V[3]
    p2 = ((V[4] * y + V[5]) * y + V[6]) * y + V[7]
    p3 = ((V[8] * y + V[9]) * y + V[10]) * y + V[11]
    p4 = ((V[12] * y + V[13]) * y + V[14]) * y + V[15]
    return ((p1 * x + p2) * x + p3) * x + p4


def find_nearest_akima(plon, plat, rng4, X2, Y2):
    nxs, nys = plon.shape
    nxt, nyt = X2.shape
    ixy = np.zeros((nxt, nyt, 2), np.int32)
    for jjt in range(nyt):
        for jit in range(nxt):
            pxt, pyt = X2[jit, jjt], Y2[jit, jjt]
            if not ((pxt >= rng4[0]) and (pxt <= rng4[1]) and (pyt >= rng4[2]) and (pyt <= rng4[3])):
                continue
            jis = jjs = 1
            lx = ly = False
            while not (lx and ly):
                lx = False
                while not lx:
                    if jis < nxs:
                        if plon[jis - 1, jjs - 1] <= pxt and plon[jis, jjs - 1] > pxt:
                            lx = True
                        else:
                            jis += 1
                    else:
                        jis -= 1
                        lx = True
                ly = False
                while not ly:
                    if jjs < nys:
                        if plat[jis - 1, jjs - 1] <= pyt and plat[jis - 1, jjs] > pyt:
                            ly = True
                        else:
                            jjs += 1
                            lx = False
                            ly = True
                    else:
                        jjs = nys - 1
                        ly = True
            ixy[jit, jjt, :] = (jis, jjs)
    return ixy


def akima_2d(k_ew_per, X1, Y1, Z1, X2, Y2):
    """one slab, 2-D Fortran-shaped arrays, i_orca_src = 0; returns (Z2 REAL(4), ixy_pos)"""
    zXs, zYs, zFs = fill_extra_bands(k_ew_per, X1, Y1, Z1.astype(np.float64))
    kewp = k_ew_per + 4 if k_ew_per > -1 else -1
    slpx, slpy, slpxy = slopes_akima(kewp, zXs, zYs, zFs)
    poly = build_pol(zXs, zYs, zFs, slpx, slpy, slpxy)
    rng4 = (zXs.min(), zXs.max(), zYs.min(), zYs.max())
    ixy = find_nearest_akima(zXs, zYs, rng4, X2, Y2)
    nx2, ny2 = X2.shape
    Z2 = np.zeros((nx2, ny2), F32)
    for jj2 in range(ny2):
        for ji2 in range(nx2):
            px2, py2 = X2[ji2, jj2], Y2[ji2, jj2]
            if (px2 >= rng4[0]) and (px2 <= rng4[1]) and (py2 >= rng4[2]) and (py2 <= rng4[3]):
                ji1, jj1 = ixy[ji2, jj2]
                px2 = px2 - zXs[ji1 - 1, jj1 - 1]
                py2 = py2 - zYs[ji1 - 1, jj1 - 1]
                with np.errstate(all="ignore"):
                    Z2[ji2, jj2] = F32(pol_val(px2, py2, poly[ji1 - 1, jj1 - 1, :]))
            else:
                Z2[ji2, jj2] = 0.0
    return Z2, ixy


def mapping_point_2d(k_ew, Xs, Ys, xP, yP, iP, jP):
    """One target of MAPPING_BL (src/mod_bilin_2d.f90:459-634) on a curvilinear source given by Fortran-shaped 2-D arrays
    Xs, Ys [i, j]; (iP, jP) 1-based nearest point.  Returns (iqdrn, alpha, beta, ipb) or None where the reference skips
    the target (:468, :487-493)."""
    nxi, nyi = Xs.shape
    if not (jP < nyi):
        return None
    iPm1, iPp1, jPm1, jPp1 = iP - 1, iP + 1, jP - 1, jP + 1
    if iPm1 == 0 and k_ew >= 0:
        iPm1 = nxi - k_ew
    if iPp1 == nxi + 1 and k_ew >= 0:
        iPp1 = 1 + k_ew
    if iPm1 < 1 or jPm1 < 1 or iPp1 > nxi:
        return None
    X = lambda i, j: math.fmod(Xs[i - 1, j - 1], 360.0)   # noqa: E731
    Y = lambda i, j: float(Ys[i - 1, j - 1])              # noqa: E731
    lonP, latP = X(iP, jP), Y(iP, jP)
    lonN, latN, lonE, latE = X(iP, jPp1), Y(iP, jPp1), X(iPp1, jP), Y(iPp1, jP)
    lonS, latS, lonW, latW = X(iP, jPm1), Y(iP, jPm1), X(iPm1, jP), Y(iPm1, jP)
    xP = math.fmod(xP, 360.0)
    hP = heading(lonP, xP, latP, yP)
    hN, hE = heading(lonP, lonN, latP, latN), heading(lonP, lonE, latP, latE)
    hS, hW = heading(lonP, lonS, latP, latS), heading(lonP, lonW, latP, latW)
    iq = 4
    hPp = degE_to_degWE(hP)
    if hN > hE:
        hN -= 360.0
    if hPp > hN and hPp <= hE:
        iq = 1
    if hP > hE and hP <= hS:
        iq = 2
    if hP > hS and hP <= hW:
        iq = 3
    if hP > hW and hPp <= hN:
        iq = 4
    icpt = 0
    while True:
        if iq == 1:
            quad = [(lonE, latE), (X(iPp1, jPp1), Y(iPp1, jPp1)), (lonN, latN)]
        elif iq == 2:
            quad = [(lonS, latS), (X(iPp1, jPm1), Y(iPp1, jPm1)), (lonE, latE)]
        elif iq == 3:
            quad = [(lonW, latW), (X(iPm1, jPm1), Y(iPm1, jPm1)), (lonS, latS)]
        elif iq == 4:
            quad = [(lonN, latN), (X(iPm1, jPp1), Y(iPm1, jPp1)), (lonW, latW)]
        loni = [xP, lonP] + [q[0] for q in quad]
        lati = [yP, latP] + [q[1] for q in quad]
        loni = [v + 360.0 if v <= 0.0 else v for v in loni]
        ok = l_inpoly(loni[1:], lati[1:], xP, yP)
        if (not ok) and yP < 88.0:
            icpt += 1
            iq = icpt
        else:
            break
        if icpt == 5:
            iq = 0
            break
    a, b, ipb = local_coord(loni, lati)
    return iq, a, b, ipb


# ---------------------------------------------------------------------------------------------------------------
# DROWN (src/mod_drown.f90:18-198), REAL(4) arithmetic spelled out; X, mask are (ni, nj) [i, j]
def drown(k_ew, X, mask, ninc_max):
    X = (X.astype(F32) * mask.astype(F32)).astype(F32)       # X = X * mask (:70)
    ni, nj = X.shape
    maskv = mask.astype(np.int64).copy()
    nsweeps = 0

    def fold(i):
        if k_ew >= 0:
            if i > ni - k_ew:
                i = i - ni + 2 * k_ew
            if i < 1 + k_ew:
                i = i + ni - 2 * k_ew
        return i
    for jinc in range(1, ninc_max + 1):
        if not (maskv == 0).any():
            break
        nsweeps += 1
        coast = np.zeros((ni, nj), np.int64)
        for jj in range(1, nj + 1):
            for ji in range(1, ni + 1):
                jjp1, jjm1 = min(jj + 1, nj), max(jj - 1, 1)
                ji0, jip1, jim1 = ji, ji + 1, ji - 1
                if k_ew >= 0:
                    ji0, jip1, jim1 = fold(ji0), fold(jip1), fold(jim1)
                else:
                    jip1, jim1 = min(jip1, ni), max(jim1, 1)
                s = maskv[jim1 - 1, jj - 1] + maskv[jip1 - 1, jj - 1] + maskv[ji0 - 1, jjm1 - 1] + maskv[ji0 - 1, jjp1 - 1]
                coast[ji - 1, jj - 1] = min(s, 1) * (1 - maskv[ji0 - 1, jj - 1])
        ns = min(jinc, 50)
        xtmp = X.copy()
        for jj in range(1, nj + 1):
            for ji in range(1, ni + 1):
                if coast[ji - 1, jj - 1] != 1:
                    continue
                summsk = datmsk = F32(0.0)
                for jjm in range(-ns, ns + 1):
                    for jim in range(-ns, ns + 1):
                        jic, jjc = fold(ji + jim), jj + jjm
                        if 1 <= jic <= ni and 1 <= jjc <= nj:
                            arg = F32(F32(-1.0) * F32(jjm ** 2 + jim ** 2)) / F32(jinc ** 2)     # REAL(4) division
                            zweight = F32(F32(math.exp(float(arg))) * F32(maskv[jic - 1, jjc - 1]))
                            summsk = F32(summsk + zweight)
                            datmsk = F32(datmsk + F32(zweight * X[jic - 1, jjc - 1]))
                with np.errstate(all="ignore"):
                    xtmp[ji - 1, jj - 1] = F32(datmsk / summsk)
        X = xtmp
        maskv = maskv + coast
    return X, nsweeps


# ---------------------------------------------------------------------------------------------------------------
# BILIN_2D after the mapping is known (src/mod_bilin_2d.f90:209-263): the apply loop with the identical-point copy, the
# identity "masking" statement, the last-row test and the ORCA north-fold patch of the last target row.  Arrays (nj, ni).
def bilin_2d_apply(k_ew, X1w, Y1w, Z1w, X2, Y2, metrics, ab, i_orca_trg, first_call=True):
    f4 = np.float32
    ny2, nx2 = X2.shape
    rmv = f4(-9999.0)
    Z2 = np.full((ny2, nx2), rmv, f4)
    IM = np.maximum(metrics[:2], 1)
    eps5 = float(f4(1.0e-5))
    for jj in range(ny2):
        for ji in range(nx2):
            iP, jP, iq = int(IM[0, jj, ji]), int(IM[1, jj, ji]), int(metrics[2, jj, ji])
            if abs(degE_to_degWE(X1w[jP - 1, iP - 1]) - degE_to_degWE(X2[jj, ji])) < eps5 and abs(Y1w[jP - 1, iP - 1] - Y2[jj, ji]) < eps5:
                Z2[jj, ji] = Z1w[jP - 1, iP - 1]
            else:
                Z2[jj, ji] = interp_bl_np(k_ew, iP, jP, iq, ab[0, jj, ji], ab[1, jj, ji], Z1w)
    Z2 = (Z2 * f4(1.0) + f4(0.0) * f4(-9995.0)).astype(f4)         # mask_ignore_trg was reset to 1 (:212): Q7
    missing = False
    if first_call and i_orca_trg >= 4:
        s = f4(0.0)
        for v in Z2[ny2 - 1, :]:
            s = f4(s + v)
        rmeanv = f4(s / f4(nx2))
        missing = bool((rmeanv < f4(rmv + f4(0.1))) and (rmeanv > f4(rmv - f4(0.1))))
    if missing:
        F = Z2.T                                                    # [i, j] view, 0-based
        if i_orca_trg == 4:
            nA, nB = nx2 // 2 - 1, nx2 // 2 + 3                     # extents of the two left-hand sections (the LHS governs)
            rhs = [F[nx2 - 1 - k, ny2 - 3] for k in range(nA)]
            for k in range(nA):
                F[1 + k, ny2 - 1] = rhs[k]
            rhs = [F[1 + k, ny2 - 3] for k in range(nB)]
            for k in range(nB):
                F[nx2 - 1 - k, ny2 - 1] = rhs[k]
        if i_orca_trg == 6:
            nA = nx2 // 2 - 1
            rhs = [F[nx2 - 2 - k, ny2 - 2] for k in range(nA)]
            for k in range(nA):
                F[1 + k, ny2 - 1] = rhs[k]
            rhs = [F[1 + k, ny2 - 2] for k in range(nA)]
            for k in range(nA):
                F[nx2 - 2 - k, ny2 - 1] = rhs[k]
    return Z2, missing
