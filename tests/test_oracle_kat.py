"""Known-answer / property tests of the CPU oracle and cross-checks against the independent numpy transliteration
(tests/pyref.py).  The reference has no tests of this path, so these are the pins."""
import math

import numpy as np
import pytest

from sosie_b200 import grids as G
from tests import pyref as P


@pytest.fixture(autouse=True)
def _glibc(orc):
    orc.set_libm(orc.GLIBC)
    yield
    orc.set_libm(orc.PMATH)


def test_literal_kinds(orc):
    # Appendix A, Q1: REAL(4) literals widened to REAL(8)
    assert orc.earth_radius() == 6367.44482421875
    assert P.ZR == 6367.44482421875
    assert float(np.float32(0.1)) == 0.10000000149011612 and float(np.float32(1e-9)) == 9.999999717180685e-10
    assert float(np.float32(0.51)) == 0.5099999904632568 and float(np.sqrt(np.float32(2.0))) == 1.4142135381698608


def test_heading_and_distance_kat(orc):
    assert orc.heading(10, 11, 20, 20) == 90.0          # due east
    assert orc.heading(10, 10, 20, 21) == 0.0           # due north
    assert orc.heading(10, 10, 20, 19) == 180.0
    assert orc.heading(10, 9, 20, 20) == 270.0
    assert orc.heading(359.5, 0.5, 0, 0) == 90.0        # across Greenwich
    assert orc.distance(10, 10, 20, 20) == 0.0          # zpds >= 1 branch
    d = orc.distance(0, 90, 0, 0)
    assert abs(d - orc.earth_radius() * math.pi / 2) < 1e-9
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.uniform(0, 360, 2); b = rng.uniform(-89, 89, 2)
        assert orc.distance(a[0], a[1], b[0], b[1]) == P.distance(a[0], a[1], b[0], b[1])
        assert orc.heading(a[0], a[1], b[0], b[1]) == P.heading(a[0], a[1], b[0], b[1])


def test_l_inpoly_nudge_and_frames(orc):
    sq = ([10, 11, 11, 10], [20, 20, 21, 21])           # integer vertices: +0.001 nudge (Q9)
    assert orc.l_inpoly(*sq, 10.5, 20.5)
    assert not orc.l_inpoly(*sq, 10.5, 20.0005)         # inside the nudged strip -> rejected
    assert not orc.l_inpoly(*sq, 12.0, 20.5)
    hs = ([10.5, 11.5, 11.5, 10.5], [20.5, 20.5, 21.5, 21.5])   # half-integer vertices: clean
    assert orc.l_inpoly(*hs, 10.6, 20.5004)
    gw = ([359.5, 0.5, 0.5, 359.5], [10.5, 10.5, 11.5, 11.5])   # spans Greenwich -> -180..180 frame
    assert orc.l_inpoly(*gw, 0.2, 11.0) and orc.l_inpoly(*gw, 359.8, 11.0) and not orc.l_inpoly(*gw, 1.0, 11.0)
    rng = np.random.default_rng(1)
    for _ in range(3000):
        x0, y0 = rng.uniform(1, 358), rng.uniform(-80, 80)
        if rng.random() < 0.3:
            x0, y0 = float(int(x0)), float(int(y0))
        dx, dy = rng.uniform(0.05, 1.5, 2)
        sk = rng.uniform(-0.3, 0.3, 4)
        vx = [x0, x0 + dx, x0 + dx + sk[0], x0 + sk[1]]
        vy = [y0, y0 + sk[2], y0 + dy, y0 + dy + sk[3]]
        px, py = x0 + rng.uniform(-0.2, 1.2) * dx, y0 + rng.uniform(-0.2, 1.2) * dy
        if min(vx) < 0 or px < 0:
            continue
        assert orc.l_inpoly(vx, vy, px, py) == P.l_inpoly(vx, vy, px, py), (vx, vy, px, py)


def test_local_coord_rectangle_and_transliteration(orc):
    # rectangular cell: alpha, beta are the exact fractions to ~1e-12 (A.1), three iterations
    lam = [10.3, 10.0, 11.0, 11.0, 10.0]; phi = [20.6, 20.0, 20.0, 21.0, 21.0]
    a, b, ipb = orc.local_coord(lam, phi)
    assert ipb == 0 and abs(a - 0.3) < 1e-12 and abs(b - 0.6) < 1e-12
    a, b, ipb = orc.local_coord([10.3, 10, 11, 11, 10], [20.0, 20, 20, 20, 20])     # all latitudes equal -> 12
    assert (a, b, ipb) == (0.5, 0.5, 12)
    rng = np.random.default_rng(2)
    for _ in range(500):
        x0, y0 = rng.uniform(0.5, 359), rng.uniform(-80, 80)
        sk = rng.uniform(-0.2, 0.2, 6)
        lam = [x0 + rng.uniform(0, 1), x0, x0 + 1 + sk[0], x0 + 1 + sk[1], x0 + sk[2]]
        phi = [y0 + rng.uniform(0, 1), y0, y0 + sk[3], y0 + 1 + sk[4], y0 + 1 + sk[5]]
        assert orc.local_coord(lam, phi) == P.local_coord(lam, phi)


def test_interp_bl_error_codes_and_weights(orc):
    Z = np.arange(30, dtype=np.float32).reshape(5, 6)          # (nj=5, ni=6)
    assert orc.interp_bl(-1, 3, 3, 1, 0.0, 0.0, Z) == Z[2, 2]
    assert orc.interp_bl(-1, 3, 3, 1, 1.0, 1.0, Z) == Z[3, 3]
    assert orc.interp_bl(-1, 1, 3, 3, 0.5, 0.5, Z) == -9997.0  # i-1 = 0, not periodic
    assert orc.interp_bl(-1, 3, 1, 2, 0.5, 0.5, Z) == -9996.0  # j-1 = 0
    assert orc.interp_bl(-1, 3, 5, 1, 0.5, 0.5, Z) == -9995.0  # j+1 > nj
    assert orc.interp_bl(-1, 6, 3, 1, 0.5, 0.5, Z) == -9994.0  # i+1 > ni
    assert orc.interp_bl(2, 1, 3, 3, 0.5, 0.5, Z) != -9997.0   # periodic wrap i-1 -> ni-k_ew
    # affine fields are reproduced exactly inside a cell (bilinear property)
    f = lambda i, j: 2.0 * i - 3.0 * j + 1.0   # noqa: E731
    Zf = np.array([[f(i, j) for i in range(1, 7)] for j in range(1, 6)], np.float32)
    assert abs(orc.interp_bl(-1, 2, 2, 1, 0.25, 0.75, Zf) - f(2.25, 2.75)) < 1e-5


def test_easy_search_and_mapping_vs_transliteration(orc):
    """EASY search + MAPPING_BL body: C++ oracle vs the numpy/Python transliteration, bit for bit."""
    lon, lat = G.regular_latlon(90, 45)
    rng = np.random.default_rng(3)
    nt = 300
    tx = rng.uniform(0.0, 359.9, nt); ty = rng.uniform(-84, 84, nt)
    tx[:40] = lon[rng.integers(0, 90, 40)]            # exact alignments / ties
    ty[40:80] = lat[rng.integers(3, 42, 40)]
    Xt, Yt = tx.reshape(3, 100), ty.reshape(3, 100)
    Xs, Ys = G.expand_2d(lon, lat)
    m = orc.mapping_bl(0, Xs, Ys, Xt, Yt)
    met, ab, ipb = m["metrics"], m["alphabeta"], m["ipb"]
    ndiff = 0
    for j in range(3):
        for i in range(100):
            iP, jP = P.find_nearest_easy_point(Xt[j, i], Yt[j, i], lon, lat)
            assert (met[0, j, i], met[1, j, i]) == (iP, jP)
            r = P.mapping_point_regular(0, lon, lat, Xt[j, i], Yt[j, i], iP, jP)
            if r is None:
                assert met[2, j, i] == 1 and ipb[j, i] == 1 and ab[0, j, i] == 0.0
                continue
            iq, a, b, ip = r
            # whole-array fix-ups (src/mod_bilin_2d.f90:654-679)
            eps = float(np.float32(1e-9))
            if a < 0 and a > -eps: a = 0.0
            if b < 0 and b > -eps: b = 0.0
            if a > -9999.0 and a < 0: a, ip = 0.5, 4       # (ZAB > rmv) .AND. (ZAB < 0.)  (:657)
            if a > 1: a, ip = 0.5, 5
            if b > -9999.0 and b < 0: b, ip = 0.5, 6
            if b > 1: b, ip = 0.5, 7
            if iq < 1: iq, ip = 1, 1
            assert (met[2, j, i], ab[0, j, i], ab[1, j, i], ipb[j, i]) == (iq, a, b, ip), (i, j)
            ndiff += 1
    assert ndiff > 250


def test_mapping_invariants_all_configs(orc):
    for name, scale in (("A", 0.2), ("B", 0.2), ("C", 0.25), ("D", 0.05), ("E", 0.01)):
        cfg = G.config(name, scale)
        Xs, Ys = G.src_2d(cfg); Z1 = G.field_slabs(Xs, Ys, 1)
        Z2, m = orc.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"],
                             l_reg_src=cfg["l_reg_src"], i_orca_trg=cfg["i_orca_trg"])
        met, ab, ipb = m["metrics"], m["alphabeta"], m["ipb"]
        found = met[0] >= 1
        assert ((met[0][found] <= cfg["nxs"]) & (met[1][found] >= 1)).all()
        assert set(np.unique(met[2])) <= {1, 2, 3, 4}
        assert (ab >= 0).all() and (ab <= 1).all()                 # after the fix-ups
        assert set(np.unique(ipb)) <= {-2, -1, 0, 1, 4, 5, 6, 7, 11, 12}
        ok = (ipb == 0)
        assert ok.mean() > 0.6
        # a well-mapped target is interpolated within the range of the source field
        assert Z2[0][ok].min() >= Z1.min() - 1e-3 and Z2[0][ok].max() <= Z1.max() + 1e-3


def test_bdrown_vs_transliteration_and_invariants(orc):
    rng = np.random.default_rng(4)
    for ni, nj, kew in ((24, 17, -1), (26, 15, 2), (20, 20, 0)):
        X = (rng.random((nj, ni)) * 30).astype(np.float32)
        mask = (rng.random((nj, ni)) > 0.45).astype(np.int32)
        mask[3:9, 5:13] = 0                       # a continent wider than nb_inc
        mask[:, 0] = rng.integers(0, 2, nj); mask[0, :] = rng.integers(0, 2, ni); mask[-1, :] = rng.integers(0, 2, ni)
        if kew > 0:
            mask[:, -kew:] = mask[:, :kew]; X[:, -kew:] = X[:, :kew]
        X[mask == 0] = -9999.0
        for ninc, nsm, pval, lo, hi in ((3, 2, -5.0, -1e3, 1e3), (50, 4, 0.0, 5.0, 25.0)):
            Xo, nsw = orc.bdrown(kew, X, mask, ninc, nsm, lo, hi, pval)
            Xp, nswp = P.bdrown(kew, P.to_f(X), P.to_f(mask), ninc, nsm, lo, hi, pval)
            assert nsw[0] == nswp
            assert np.array_equal(Xo, P.from_f(Xp)), (ni, nj, kew, ninc)
            sea = mask == 1
            if nsw[0] > 0:
                assert np.array_equal(Xo[sea], np.clip(X[sea], lo, hi))     # mask==1 untouched (mod_bdrown.f90:22-24)
    # no land at all: nothing happens, not even the clamp (the loop exits before :414)
    X = (rng.random((9, 11)) * 30).astype(np.float32)
    Xo, nsw = orc.bdrown(-1, X, np.ones((9, 11), np.int32), 10, 0, 5.0, 10.0, 0.0)
    assert nsw[0] == 0 and np.array_equal(Xo, X)


def test_smoother_vs_transliteration(orc):
    rng = np.random.default_rng(5)
    for ni, nj, kew in ((19, 14, -1), (22, 13, 2)):
        X = (rng.random((nj, ni)) * 10).astype(np.float32)
        msk = (rng.random((nj, ni)) > 0.4).astype(np.int32)
        a = orc.smoother(kew, X, 3)
        assert np.array_equal(a, P.from_f(P.smoother(kew, P.to_f(X), 3)))
        a = orc.smoother(kew, X, 3, msk=msk)
        assert np.array_equal(a, P.from_f(P.smoother(kew, P.to_f(X), 3, msk=P.to_f(msk))))
        assert np.array_equal(a[msk == 0], X[msk == 0])             # msk==0 preserved
        a = orc.smoother(kew, X, 3, msk=msk, l_exclude_mask_points=True)
        assert np.array_equal(a, P.from_f(P.smoother(kew, P.to_f(X), 3, msk=P.to_f(msk), l_emp=True)))
    c = np.full((12, 15), 7.25, np.float32)                         # constants are preserved to rounding
    assert np.allclose(orc.smoother(-1, c, 5), c, rtol=1e-6)
    assert np.array_equal(orc.smoother(-1, c, 5)[0], c[0])          # rows 1 and nj never touched


def test_drown_gaussian_invariants(orc):
    rng = np.random.default_rng(6)
    X = (rng.random((20, 30)) * 10 + 1).astype(np.float32)
    mask = (rng.random((20, 30)) > 0.3).astype(np.int8)
    Xo, nsw = orc.drown(-1, X, mask, 30)
    assert np.array_equal(Xo[mask == 1], X[mask == 1])
    assert nsw[0] >= 1 and np.isfinite(Xo).all() and Xo[mask == 0].min() >= X[mask == 1].min() - 1e-4
    assert Xo[mask == 0].max() <= X[mask == 1].max() + 1e-4       # weighted means of sea values


def test_fill_extra_bands_vs_transliteration(orc):
    rng = np.random.default_rng(7)
    lon, lat = G.regular_latlon(30, 16)
    XX, YY = G.expand_2d(lon, lat)
    F = rng.random(XX.shape)
    for kew in (0, 2, -1):
        a = orc.fill_extra_bands(kew, XX, YY, F)
        b = P.fill_extra_bands(kew, P.to_f(XX), P.to_f(YY), P.to_f(F))
        for u, v in zip(a, b):
            assert np.array_equal(u, P.from_f(v)), kew


def test_akima_reproduces_linear_and_quadratic(orc):
    lon = 10.0 + 0.5 * np.arange(40); lat = -10.0 + 0.5 * np.arange(30)
    X, Y = G.expand_2d(lon, lat)
    tl = 12.03 + 0.37 * np.arange(40); tt = -8.02 + 0.29 * np.arange(40)
    TX, TY = G.expand_2d(tl, tt)
    for f, tol in ((lambda x, y: 2 * x - 3 * y + 1, 2e-5), (lambda x, y: 0.01 * x * x - 0.02 * y * y + 0.03 * x * y, 2e-5)):
        Z2, _ = orc.akima_2d(-1, lon, lat, f(X, Y).astype(np.float32), tl, tt)
        ref = f(TX, TY)
        assert np.max(np.abs(Z2[0] - ref) / np.maximum(1.0, np.abs(ref))) < tol
    # outside the (frame-extended) source domain -> 0.0 (src/mod_akima_2d.f90:222-224)
    Z2, _ = orc.akima_2d(-1, lon, lat, np.ones(X.shape, np.float32), np.array([0.0, 200.0]), np.array([50.0, 60.0]))
    assert (Z2 == 0).all()


def test_ext_north_to_90_regg(orc):
    lon, lat = G.regular_latlon(24, 22)
    X, Y = G.expand_2d(lon, lat)
    F = np.random.default_rng(8).random(X.shape).astype(np.float32)
    XP, YP, FP = orc.ext_north_to_90_regg(X, Y, F)
    assert YP.shape == (24, 24) and (YP[22] == 90.0).all() and np.allclose(YP[23], 90.0 + (lat[-1] - lat[-2]))
    im = (np.arange(24) + 12) % 24                       # the column across the pole
    assert np.array_equal(FP[22], np.float32(0.5) * (F[21] + F[21][im]))
    assert np.array_equal(FP[23], F[20][im])


def test_rotation_roundtrip(orc):
    rng = np.random.default_rng(9)
    U = rng.standard_normal((2, 500)).astype(np.float32); V = rng.standard_normal((2, 500)).astype(np.float32)
    ang = rng.random(500) * 2 * np.pi
    uo, vo = orc.rotate_uv(U, V, np.cos(ang), np.sin(ang))
    ub, vb = orc.rotate_uv(uo, vo, np.cos(ang), np.sin(ang), inverse=True)
    assert np.allclose(ub, U, atol=2e-6) and np.allclose(vb, V, atol=2e-6)
    assert np.allclose(uo ** 2 + vo ** 2, U ** 2 + V ** 2, rtol=1e-5)


def test_faithful_fill_equals_elided(orc):
    """the per-target Xdist = 1.E12 whole-array store (src/mod_manip.f90:1038,1368) is dead: results are identical"""
    for name, scale in (("D", 0.04), ("B", 0.15)):
        cfg = G.config(name, scale)
        Xs, Ys = G.src_2d(cfg); Z1 = G.field_slabs(Xs, Ys, 1)
        a, ma = orc.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"],
                             l_reg_src=cfg["l_reg_src"], elide_fill=False)
        b, mb = orc.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"],
                             l_reg_src=cfg["l_reg_src"], elide_fill=True, nthreads=2)
        assert np.array_equal(a, b) and all(np.array_equal(ma[k], mb[k]) for k in ("metrics", "alphabeta", "ipb"))


# ---- trajectory front end: the time loop of interp_to_ground_track.f90
def track_case(seed=3, ni=40, nj=30, Ntm=5, Ntf=300):
    rng = np.random.default_rng(seed)
    fields = rng.normal(0.0, 1.0, (Ntm, nj, ni)).astype(np.float32)
    fields[:, 5, 7] = -9999.0                                   # a missing value inside the field: the < -9990 clip
    vt_mod = np.cumsum(rng.uniform(0.5, 1.5, Ntm)) + 20000.0
    imask = (rng.uniform(size=(nj, ni)) > 0.04).astype(np.int8)
    vtf = np.sort(rng.uniform(vt_mod[0] - 0.3, vt_mod[-1] + 0.3, Ntf))
    vtf[10] = vt_mod[1]                                         # exactly on a model record
    vtf.sort()                                                  # the reference scans forward from the last interval
    metrics = np.stack([rng.integers(0, ni + 1, Ntf), rng.integers(0, nj + 1, Ntf), rng.integers(1, 5, Ntf)]).astype(np.int32)
    metrics[0, :40] = 7; metrics[1, :40] = 6                    # neighbours of the missing value
    ab = rng.uniform(0.0, 1.0, (2, Ntf))
    F_gt_f = rng.normal(0.0, 9.0, Ntf)                          # some outside (-20, 20)
    return fields, vt_mod, imask, vtf, metrics, ab, F_gt_f


def track_by_intervals(run, fields, vt_mod, imask, vtf, metrics, ab, F_gt_f, rmissval=-9999.0):
    """what a tool using the library does: group the track points by model interval, one call per interval"""
    Ntf = vtf.size
    mod = np.full(Ntf, rmissval); mod_np = np.full(Ntf, rmissval); fmask = np.zeros(Ntf, np.int8)
    for jt in range(vt_mod.size - 1):
        sel = np.nonzero((vtf >= vt_mod[jt]) & (vtf < vt_mod[jt + 1]) & (vtf >= vt_mod.min()) & (vtf < vt_mod.max()))[0]
        if sel.size == 0:
            continue
        a, b, m = run(fields[jt], fields[jt + 1], vt_mod[jt], vt_mod[jt + 1], imask, vtf[sel], metrics[0, sel], metrics[1, sel], metrics[2, sel],
                      ab[0, sel], ab[1, sel], F_gt_f[sel])
        mod[sel] = a[:, 0]; mod_np[sel] = b[:, 0]; fmask[sel] = m[:, 0]
    obs = np.where(fmask == 1, F_gt_f, rmissval)
    return mod, mod_np, obs, fmask


def test_ground_track_loop_vs_transliteration(orc):
    from tests import pyref
    case = track_case()
    a = track_by_intervals(orc.track_interp, *case)
    b = pyref.ground_track_loop_np(*case)
    assert b[3].sum() > 50 and (b[3] == 0).sum() > 20
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert (b[0][b[3] == 1] == -9999.0).sum() > 0                # the clip fired on valid points


def test_track_interp_levels_no_time(orc):
    """interp_to_hydro_section.f90:567-607: nk levels of one record, per-level masks, obs window (-20, 50)"""
    from tests import pyref
    rng = np.random.default_rng(5)
    nk, nj, ni, n = 6, 20, 25, 80
    F = rng.normal(size=(nk, nj, ni)).astype(np.float32)
    imask = (rng.uniform(size=(nk, nj, ni)) > 0.03).astype(np.int8)
    iP, jP, iq = rng.integers(0, ni + 1, n), rng.integers(0, nj + 1, n), rng.integers(1, 5, n)
    xa, xb = rng.uniform(size=n), rng.uniform(size=n)
    r_obs = rng.uniform(-30, 60, n)
    f_bl, f_np, fm = orc.track_interp(F, None, 0.0, 1.0, imask, None, iP, jP, iq, xa, xb, r_obs, -20.0, 50.0, False)
    for p in range(n):
        for k in range(nk):
            M = lambda i, j: int(imask[k, j - 1, i - 1])
            ok = iP[p] > 0 and jP[p] > 0 and M(iP[p], jP[p]) == 1 and -20.0 < r_obs[p] < 50.0
            if ok:
                i, j = int(iP[p]), int(jP[p])
                ip1, jp1, im1, jm1 = min(i + 1, ni), min(j + 1, nj), max(i - 1, 1), max(j - 1, 1)
                ok = M(ip1, j) + M(ip1, jp1) + M(i, jp1) + M(im1, jp1) + M(im1, j) + M(im1, jm1) + M(i, jm1) + M(ip1, jm1) == 8
            assert fm[p, k] == int(ok)
            if ok:
                assert f_np[p, k] == np.float64(F[k, j - 1, i - 1])
                assert f_bl[p, k] == np.float64(pyref.interp_bl_np(-1, i, j, int(iq[p]), xa[p], xb[p], F[k]))
            else:
                assert f_bl[p, k] == -9999.0 and f_np[p, k] == -9999.0


def test_drown_refuses_grids_narrower_than_the_fold(orc):
    """src/mod_drown.f90:110-117: with ni < 2*k_ew + 1 the folded neighbour indices leave 1..ni (the reference would read
    maskv out of bounds): the oracle reports it instead of computing something."""
    import pytest
    X = np.zeros((6, 4), np.float32)
    m = np.ones((6, 4), np.int8)
    with pytest.raises(RuntimeError):
        orc.drown(2, X, m, 3)
    orc.drown(2, np.zeros((6, 5), np.float32), np.ones((6, 5), np.int8), 3)   # ni = 2*k_ew + 1 is the smallest legal width


@pytest.mark.parametrize("case", ["B_small", "C_small", "track", "regional_target", "masked", "polar_target"])
def test_twisted_search_vs_second_transcription(orc, case):
    """FIND_NEAREST_POINT on curvilinear sources (prelude + IS_POLAR_PROJ + FIND_NEAREST_TWISTED, src/mod_manip.f90:884-947,
    1073-1440, 1843-1879): the C++ oracle against the independent Python transcription in tests/pyref.py, both on glibc's
    libm -- JIp, JJp, the search mask and the target row range must agree exactly."""
    rng = np.random.default_rng(5)
    mask = None
    if case == "B_small":
        cfg = G.config("B", 0.12)
        Xs, Ys = G.src_2d(cfg); Xt, Yt = G.trg_2d(cfg)
    elif case == "C_small":
        cfg = G.config("C", 0.2)
        Xs, Ys = G.src_2d(cfg); Xt, Yt = G.trg_2d(cfg)
    else:
        Xs, Ys = G.orca_like(48, 36, "F", 30.0)
        if case == "track":          # (1, N) target: nx_t = 1, no row-range logic
            n = 160
            t = np.linspace(0.0, 1.0, n)
            Xt = np.mod(10.0 + 340.0 * t, 360.0).reshape(n, 1)
            Yt = (70.0 * np.sin(9.0 * t)).reshape(n, 1)
        elif case == "polar_target":      # a polar-stereographic target with the pole inside: IS_POLAR_PROJ -> the full scan (lStupido)
            xx, yy = np.meshgrid(np.linspace(-1.0, 1.0, 15), np.linspace(-1.0, 1.0, 15))
            Yt = 90.0 - 28.0 * np.hypot(xx, yy)
            Xt = np.mod(np.degrees(np.arctan2(xx, -yy)), 360.0)
        elif case == "regional_target":   # rows south of the source and a tilted grid (latitudes not monotone along j everywhere)
            u, v = np.meshgrid(np.linspace(0, 1, 23), np.linspace(0, 1, 17))
            Xt = np.mod(200.0 + 60.0 * u + 8.0 * v, 360.0)
            Yt = -85.0 + 120.0 * v + 6.0 * u
        else:
            Xt, Yt = np.meshgrid(np.linspace(3.0, 357.0, 31), np.linspace(-70.0, 80.0, 19))
            Xt = Xt + 0.3 * np.sin(Yt)    # irregular target
            mask = (rng.random(Xt.shape) > 0.25).astype(np.int32)
    orc.set_libm(orc.GLIBC)
    try:
        JIo, JJo, mo, used, info = orc.find_nearest_point(Xt, Yt, Xs, Ys, mask=mask)
        assert orc.get_error() == 0
    finally:
        orc.set_libm(orc.PMATH)
    XsF, YsF, XtF, YtF = P.to_f(Xs), P.to_f(Ys), P.to_f(Xt), P.to_f(Yt)
    reg_s, reg_t, j_strt, j_stop, icr = P.find_nearest_prelude(XtF, YtF, XsF, YsF)
    assert not reg_s
    assert (j_strt, j_stop, icr, int(reg_s), int(reg_t)) == tuple(int(v) for v in info[:5])
    m0 = np.ones(XtF.shape, np.int32) if mask is None else P.to_f(mask).astype(np.int32)
    JI, JJ, m = P.find_nearest_twisted(XtF, YtF, XsF, YsF, m0, j_strt, j_stop, icr)
    assert np.array_equal(P.from_f(JI), JIo), f"{case}: JIp differs at {(P.from_f(JI) != JIo).sum()} targets"
    assert np.array_equal(P.from_f(JJ), JJo)
    assert np.array_equal(P.from_f(m), mo)
    assert (JIo > 0).mean() > 0.3
    if case == "polar_target":
        assert P.is_polar_proj(XtF, YtF) and not P.is_polar_proj(XsF, YsF)


@pytest.mark.parametrize("kind,kew", [("cell", 0), ("regional", -1), ("overlap", 2)])
def test_akima_2d_vs_second_transcription(orc, kind, kew):
    """AKIMA_2D end to end (frame, SLOPES_AKIMA, the 16 coefficients of build_pol for EVERY cell as the reference stores them,
    find_nearest_akima's walk, pol_val, the REAL(4) cast, 0 outside the source): the C++ oracle -- which rebuilds the
    coefficients per target -- against the whole-array numpy transcription in tests/pyref.py, bit for bit.  The field has
    flat and linear patches so that the equal-slope branch (:558-564) and the zero-weight fallback (:638-647) are taken."""
    rng = np.random.default_rng(11)
    if kind == "cell":
        lon = (np.arange(36) + 0.5) * 10.0; lat = -85.0 + np.arange(18) * 10.0
    elif kind == "regional":
        lon = 20.0 + 0.7 * np.arange(30); lat = -12.0 + 0.9 * np.arange(22)
    else:
        lon = np.arange(38) * (360.0 / 36.0); lat = np.linspace(-70.0, 80.0, 21)     # 2 overlap columns: lon(37) = 360, lon(38) = 370
    X, Y = G.expand_2d(lon, lat)
    F = (10.0 * np.cos(np.deg2rad(3 * X)) * np.sin(np.deg2rad(2 * Y)) + rng.normal(0.0, 0.5, X.shape)).astype(np.float32)
    F[5:12, 8:20] = 3.25                              # flat patch: all divided differences equal
    F[13:, :10] = (0.5 * X[13:, :10] - 0.25 * Y[13:, :10]).astype(np.float32)    # linear patch
    if kind == "overlap":
        F[:, -2:] = F[:, :2]
    tl = np.sort(rng.uniform(lon.min() - 3.0, lon.max() + 3.0, 41)); tt = np.sort(rng.uniform(lat.min() - 3.0, lat.max() + 3.0, 33))
    tl[:6] = lon[rng.integers(0, lon.size, 6)]; tl.sort()      # targets exactly on source nodes (half-open cell tests)
    tt[:5] = lat[rng.integers(0, lat.size, 5)]; tt.sort()
    TX, TY = G.expand_2d(tl, tt)
    TX = TX + 0.05 * np.sin(TY)                       # 2-D target arrays
    Z2o, ixo = orc.akima_2d(kew, X, Y, F, TX, TY)
    Z2p, ixp = P.akima_2d(kew, P.to_f(X), P.to_f(Y), P.to_f(F), P.to_f(TX), P.to_f(TY))
    assert np.array_equal(ixo[0], P.from_f(ixp[:, :, 0])) and np.array_equal(ixo[1], P.from_f(ixp[:, :, 1]))
    Z2p = P.from_f(Z2p)
    assert np.array_equal(np.isnan(Z2o[0]), np.isnan(Z2p))
    ok = ~np.isnan(Z2p)
    assert np.array_equal(Z2o[0][ok], Z2p[ok]), f"{(Z2o[0][ok] != Z2p[ok]).sum()} of {ok.sum()} differ, max {np.abs(Z2o[0][ok] - Z2p[ok]).max()}"
    assert (ixo[0] > 0).mean() > 0.7
    if kind == "regional":
        assert (Z2p == 0).any()       # targets beyond the 2-cell frame: 0.0 (:222-224)
    # the slopes alone, on the frame-extended arrays
    a = orc.fill_extra_bands(kew, X, Y, F.astype(np.float64))
    so = orc.slopes_akima(kew + 4 if kew > -1 else -1, *a)
    sp = P.slopes_akima(kew + 4 if kew > -1 else -1, *(P.to_f(v) for v in a))
    for u, v in zip(so, sp):
        v = P.from_f(v)
        assert np.array_equal(np.isnan(u), np.isnan(v)) and np.array_equal(u[~np.isnan(u)], v[~np.isnan(v)])


@pytest.mark.parametrize("piv", ["F", "T"])
def test_mapping_bl_on_curvilinear_source_vs_transliteration(orc, piv):
    """MAPPING_BL on an ORCA-shaped source (TWISTED search, 2-D neighbour lookups, periodic wrap of iP-1 / iP+1, the
    quadrant retry loop, LOCAL_COORD, the whole-array fix-ups): C++ oracle against the per-target Python transliteration,
    both on glibc's libm (heading = TAN / LOG / ATAN2 decides the first-guess quadrant)."""
    Xs, Ys = G.orca_like(54, 40, piv, 25.0)
    Xt, Yt = G.expand_2d(np.linspace(1.5, 358.5, 40), np.linspace(-76.0, 87.0, 31))
    orc.set_libm(orc.GLIBC)
    try:
        m = orc.mapping_bl(2, Xs, Ys, Xt, Yt)
        assert orc.get_error() == 0
    finally:
        orc.set_libm(orc.PMATH)
    met, ab, ipb = m["metrics"], m["alphabeta"], m["ipb"]
    XsF, YsF = P.to_f(Xs), P.to_f(Ys)
    eps = float(np.float32(1e-9))
    ncmp, quads = 0, set()
    for j in range(Xt.shape[0]):
        for i in range(Xt.shape[1]):
            iP, jP = int(met[0, j, i]), int(met[1, j, i])
            if iP < 1:
                assert ipb[j, i] in (-1, -2)         # not found by the search (:681-682)
                continue
            r = P.mapping_point_2d(2, XsF, YsF, Xt[j, i], Yt[j, i], iP, jP)
            if r is None:
                assert met[2, j, i] == 1 and ipb[j, i] == 1 and ab[0, j, i] == 0.0 and ab[1, j, i] == 0.0
                continue
            iq, a, b, ip = r
            if a < 0 and a > -eps: a = 0.0
            if b < 0 and b > -eps: b = 0.0
            if a > -9999.0 and a < 0: a, ip = 0.5, 4       # (ZAB > rmv) .AND. (ZAB < 0.)  (:657)
            if a > 1: a, ip = 0.5, 5
            if b > -9999.0 and b < 0: b, ip = 0.5, 6
            if b > 1: b, ip = 0.5, 7
            if iq < 1: iq, ip = 1, 1
            assert (int(met[2, j, i]), ab[0, j, i], ab[1, j, i], int(ipb[j, i])) == (iq, a, b, ip), (i, j, iP, jP)
            ncmp += 1
            quads.add(iq)
    assert ncmp > 0.7 * Xt.size and quads == {1, 2, 3, 4}


@pytest.mark.parametrize("kew", [-1, 0, 2])
def test_drown_gaussian_vs_transliteration(orc, kew):
    """DROWN (src/mod_drown.f90): coastline by the four folded neighbours, Gaussian box mean with REAL(4) sums in loop order,
    weights EXP(-r2/jinc**2) in REAL(4): C++ oracle against Python loops, bit for bit, sweep counts included."""
    rng = np.random.default_rng(8 + kew)
    ni, nj = 15, 11
    X = (rng.random((nj, ni)) * 10 + 1).astype(np.float32)
    yy, xx = np.mgrid[0:nj, 0:ni]
    mask = ((np.sin(xx * 0.9) * np.cos(yy * 0.7) < 0.35) | (rng.random((nj, ni)) < 0.15)).astype(np.int8)
    mask[2:9, 4:11] = 0                                   # a land mass several sweeps deep
    for ninc in (1, 3, 7):
        Xo, nsw = orc.drown(kew, X, mask, ninc)
        Xp, nsp = P.drown(kew, P.to_f(X), P.to_f(mask), ninc)
        Xp = P.from_f(Xp)
        assert int(nsw[0]) == nsp
        assert np.array_equal(np.isnan(Xo), np.isnan(Xp)) and np.array_equal(Xo[~np.isnan(Xo)], Xp[~np.isnan(Xp)]), (kew, ninc)
    assert nsp >= 3


@pytest.mark.parametrize("piv,iorca,mask_last_row", [("T", 4, True), ("F", 6, True), ("T", 4, False), ("F", 0, True)])
def test_bilin_2d_apply_and_north_fold_patch_vs_transliteration(orc, piv, iorca, mask_last_row):
    """BILIN_2D after the mapping (src/mod_bilin_2d.f90:209-263): index clamp, identical-point copy, INTERP_BL, the last-row
    test (REAL(4) mean within 0.1 of rmv) and the north-fold patch of the last target row (T and F pivots; the T-pivot
    sections do not conform: the left-hand extent governs).  The last target row is masked out of the mapping so that it
    receives Z1w(1,1) = -9999 and the patch fires; with the row left in, or i_orca_trg = 0, it must not."""
    Xs, Ys = G.orca_like(44, 32, "F", 10.0)
    Xt, Yt = G.orca_like(52, 38, piv, 10.0)
    Xt[3, 5], Yt[3, 5] = Xs[4, 7], Ys[4, 7]               # a target ON a source node: the identical-point copy
    Z1 = G.field_slabs(Xs, Ys, 1, seed=12)
    Z1[0, 0, 0] = -9999.0                                 # Z1w(1,1): what clamped (not searched / not found) targets receive
    mask = np.ones(Xt.shape, np.int32)
    if mask_last_row:
        mask[-1, :] = 0
    Z2o, mo = orc.bilin_2d(2, Xs, Ys, Z1, Xt, Yt, mask_domain_trg=mask, l_reg_src=False, i_orca_trg=iorca)
    Z2p, missing = P.bilin_2d_apply(2, Xs, Ys, Z1[0], Xt, Yt, mo["metrics"], mo["alphabeta"], iorca)
    assert missing == (mask_last_row and iorca >= 4)
    assert bool(mo["l_last_y_row_missing"]) == missing
    assert np.array_equal(Z2o[0], Z2p), f"{(Z2o[0] != Z2p).sum()} points differ"
    assert Z2p[3, 5] == Z1[0, 4, 7]
    if missing:
        assert (Z2p[-1, 1:-1] != np.float32(-9999.0)).mean() > 0.5      # the patched row holds copies of row ny-2 / ny-1


def test_bilin_2d_regular_source_with_pole_rows_vs_transliteration(orc):
    """BILIN_2D on a regular, cell-centred global source: the two rows of EXT_NORTH_TO_90_REGG are appended
    (src/mod_bilin_2d.f90:119-171) and targets north of the last source latitude interpolate in them.  The oracle's
    fields against the apply transliteration fed with the extended working arrays."""
    lon, lat = G.regular_latlon(48, 24)                   # top latitude 86.25: 2*ymx - y(ny-1) >= 90
    X, Y = G.expand_2d(lon, lat)
    Z1 = G.field_slabs(X, Y, 1, seed=13)
    tl = np.linspace(2.0, 358.0, 37); tt = np.linspace(60.0, 89.5, 25)
    tl[5], tt[3] = lon[9], lat[20]                        # one target on a source node
    Xt, Yt = G.expand_2d(tl, tt)
    Z2o, mo = orc.bilin_2d(0, lon, lat, Z1, tl, tt, l_reg_src=True)
    assert mo["info"][4] == 1 or mo["metrics"][1].max() > 24        # the search saw the extended grid
    X1w, Y1w, Z1w = orc.ext_north_to_90_regg(X, Y, Z1[0])
    Z2p, _ = P.bilin_2d_apply(0, X1w, Y1w, Z1w, Xt, Yt, mo["metrics"], mo["alphabeta"], 0)
    assert np.array_equal(Z2o[0], Z2p), f"{(Z2o[0] != Z2p).sum()} points differ"
    assert (mo["metrics"][1] >= 24).any()                 # some targets have their nearest point on the last real row or the pole row
    assert Z2p[3, 5] == Z1[0, 20, 9]
