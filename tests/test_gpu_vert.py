"""GPU parity, vertical stage of INTERP_3D (csrc/vert.cu) through the C ABI against the CPU oracle: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _levels(n, zmax, stretch=2.2):
    s = np.linspace(0.0, 1.0, n)
    return (5.0 + (zmax - 5.0) * (np.sinh(stretch * s) / np.sinh(stretch))).astype(np.float32)


@pytest.mark.parametrize("nk1,nk2,ny,nx", [(31, 75, 40, 52), (3, 5, 7, 9), (46, 31, 33, 17), (75, 31, 292, 362)])
def test_akima_1d_3d_vs_oracle(gpu, orc, nk1, nk2, ny, nx):
    rng = np.random.default_rng(nk1 * 100 + nk2)
    vz1 = _levels(nk1, 5500.0)
    vz2 = _levels(nk2, 5900.0, 1.7)
    vz2[0] = 1.0
    if nk2 > 3:
        vz2[2] = vz1[1]
    xf1 = (np.linspace(25.0, 2.0, nk1, dtype=np.float32)[:, None, None] + rng.normal(0.0, 0.4, (nk1, ny, nx))).astype(np.float32)
    xf1[:, ::5, ::3] = 0.0                     # flat columns: equal slopes (m1 == m2 and m3 == m4) and 0/0-free paths
    a = gpu.akima_1d_3d(vz1, xf1, vz2, -9999.0)
    b = orc.akima_1d_3d(vz1, xf1, vz2, -9999.0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"max diff {np.abs(a - b).max()} at {(a != b).sum()} points"


def test_akima_1d_3d_identical_and_degenerate_levels(gpu, orc):
    """source == target levels (every level hits an equality or the `<=` bracket), and repeated source depths
    (A == 0 / B == 0 in the frame extrapolation, zero divisors in the slopes -> inf/NaN must match too)."""
    rng = np.random.default_rng(3)
    vz1 = _levels(10, 3000.0)
    xf1 = rng.normal(5.0, 2.0, (10, 6, 8)).astype(np.float32)
    a = gpu.akima_1d_3d(vz1, xf1, vz1.copy(), -9999.0)
    b = orc.akima_1d_3d(vz1, xf1, vz1.copy(), -9999.0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    vzd = vz1.copy(); vzd[1] = vzd[0]; vzd[-1] = vzd[-2]
    vz2 = _levels(14, 3100.0, 1.5)
    with np.errstate(all="ignore"):
        a = gpu.akima_1d_3d(vzd, xf1, vz2, -9999.0)
        b = orc.akima_1d_3d(vzd, xf1, vz2, -9999.0)
    # NaN payload / sign bits are not specified by IEEE-754 (x86 gives 0xFFC00000 for 0/0, the GPU 0x7FFFFFFF): NaNs must sit
    # at the same places, everything else must be bit-identical
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    assert np.array_equal(a[ok].view(np.uint32), b[ok].view(np.uint32))


@pytest.mark.parametrize("nperio", [0, 1, 2, 3, 4, 5, 6])
def test_lbc_lnk_vs_oracle(gpu, orc, nperio):
    rng = np.random.default_rng(5)
    for (jpj, jpi, ns) in ((149, 182, 3), (292, 362, 2), (9, 12, 4), (10, 13, 1), (3059, 4322, 1) if nperio == 6 else (11, 16, 1)):
        X = rng.normal(size=(ns, jpj, jpi)).astype(np.float32)
        for cd in "TUVF":
            for psgn, pval in ((1.0, None), (-1.0, None), (1.0, -2.5)):
                a = gpu.lbc_lnk(nperio, X, cd, psgn, pval)
                b = orc.lbc_lnk(nperio, X, cd, psgn, pval)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (nperio, jpi, jpj, cd, psgn, pval, (a != b).sum())


def test_interp3d_post_vs_oracle(gpu, orc):
    rng = np.random.default_rng(13)
    nk, nj, ni = 75, 60, 82
    X = rng.normal(10.0, 8.0, (nk, nj, ni)).astype(np.float32)
    X[5, 7, 9] = -9999.0
    bed = rng.integers(0, nk + 1, (nj, ni))
    mask = (np.arange(nk)[:, None, None] < bed[None]).astype(np.int32)
    for ixb, iorca, cd, lmout in ((1, 0, "T", True), (0, 6, "T", True), (1, 4, "U", False), (0, 0, "T", False), (1, 6, "V", True)):
        a = gpu.interp3d_post(X, mask, ixb, 0.0, 20.0, -9999.0, iorca, cd, lmout)
        b = orc.interp3d_post(X, mask, ixb, 0.0, 20.0, -9999.0, iorca, cd, lmout)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (ixb, iorca, cd, lmout, (a != b).sum())
    a = gpu.interp3d_post(X, None, 0, -5.0, 15.0)      # no mask needed when neither persistence nor lmout
    b = orc.interp3d_post(X, None, 0, -5.0, 15.0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_3d_job_on_device(gpu, orc):
    """config C end to end on the device, one record: per-level BDROWN + bilinear (TWISTED mapping) of 31 levels, the
    cube stays in HBM for AKIMA_1D_3D to 75 levels and the INTERP_3D post-ops; compared with the oracle's chain."""
    import torch
    from sosie_b200 import grids as G
    cfg = G.config("C", 0.35)
    nk1, nk2 = 31, 75
    Xs, Ys = cfg["src_lon"], cfg["src_lat"]
    X = G.field_slabs(Xs, Ys, nk1, seed=21)
    masks = np.stack([G.land_mask(Xs, Ys, thresh=0.25 - 0.4 * k / nk1) for k in range(nk1)])
    X[masks == 0] = -9999.0
    vz1, vz2 = _levels(nk1, 5000.0), _levels(nk2, 5800.0, 2.6)
    nxt, nyt = cfg["nxt"], cfg["nyt"]
    rng = np.random.default_rng(2)
    bed = rng.integers(0, nk2 + 1, (nyt, nxt))
    mask_trg = (np.arange(nk2)[:, None, None] < bed[None]).astype(np.int32)
    # oracle chain
    Xd, _ = orc.bdrown(cfg["ewper_src"], X, masks, 20, 5, -1.0e3, 1.0e3, 0.0)
    T, _ = orc.bilin_2d(cfg["ewper_src"], Xs, Ys, Xd, cfg["trg_lon"], cfg["trg_lat"], i_orca_trg=cfg["i_orca_trg"])
    V = orc.akima_1d_3d(vz1, T, vz2, -9999.0)
    R = orc.interp3d_post(V, mask_trg, 1, -2.0, 35.0, -9999.0, cfg["i_orca_trg"], "T", True)
    # device chain
    dev = torch.device("cuda", 0)
    dX, dM = torch.from_numpy(X).to(dev), torch.from_numpy(masks).to(dev)
    gpu.bdrown_dev(cfg["ewper_src"], cfg["nxs"], cfg["nys"], nk1, dX, dM, 1, 20, 5, -1.0e3, 1.0e3, 0.0)
    gpu.bilin_set_grids(cfg["ewper_src"], Xs, Ys, cfg["trg_lon"], cfg["trg_lat"], i_orca_trg=cfg["i_orca_trg"])
    gpu.bilin_map()
    dT = torch.empty((nk1, nyt, nxt), device=dev, dtype=torch.float32)
    gpu.bilin_apply_dev(nk1, dX, dT, first_call=True)
    dV = torch.empty((nk2, nyt, nxt), device=dev, dtype=torch.float32)
    gpu.akima_1d_3d_dev(nxt, nyt, vz1, dT, vz2, dV, -9999.0)
    gpu.interp3d_post_dev(nxt, nyt, nk2, dV, torch.from_numpy(mask_trg).to(dev), 1, -2.0, 35.0, -9999.0, cfg["i_orca_trg"], "T", True)
    gpu.sync()
    out = dV.cpu().numpy()
    assert np.array_equal(out.view(np.uint32), R.view(np.uint32)), (out != R).sum()


def test_extrp_hl_vs_oracle(gpu, orc):
    rng = np.random.default_rng(4)
    X = rng.normal(size=(3, 40, 25)).astype(np.float32)
    for top, btm, icr in ((33, 6, 1), (0, 4, 1), (37, 0, 1), (8, 31, -1), (0, 0, 1)):
        a = gpu.extrp_hl(X, top, btm, icr)
        b = orc.extrp_hl(X, top, btm, icr)
        assert np.array_equal(a, b), (top, btm, icr)
    a = orc.extrp_hl(X, 33, 6, 1)
    assert np.array_equal(a[:, 32:], np.broadcast_to(X[:, 31:32], a[:, 32:].shape)) and np.array_equal(a[:, :6], np.broadcast_to(X[:, 6:7], a[:, :6].shape))


def _staggered(lon, lat):
    lon = np.asarray(lon, np.float64); lat = np.asarray(lat, np.float64)
    dl = np.diff(np.unwrap(np.deg2rad(lon), axis=1), axis=1)
    dl = np.rad2deg(np.concatenate([dl, dl[:, -1:]], axis=1))
    dp = np.diff(lat, axis=0); dp = np.concatenate([dp, dp[-1:]], axis=0)
    dq = np.diff(np.rad2deg(np.unwrap(np.deg2rad(lon), axis=0)), axis=0); dq = np.concatenate([dq, dq[-1:]], axis=0)
    dr = np.diff(lat, axis=1); dr = np.concatenate([dr, dr[:, -1:]], axis=1)
    lamu, phiu = lon + 0.5 * dl, lat + 0.5 * dr
    lamv, phiv = lon + 0.5 * dq, lat + 0.5 * dp
    return lon, lat, lamu, phiu, lamv, phiv, lamu + 0.5 * dq, phiu + 0.5 * dp


def test_angle2_vs_oracle(gpu, orc):
    """ANGLE2 (src/mod_nemotools.f90:607-823): ORCA grids with both fold pivots (lbc_lnk, psgn = -1), a regional grid
    (Akima edge extrapolation) and a -180..180 grid that crosses the cut line (l_cut_180)."""
    from sosie_b200 import grids as G
    cases = []
    cfg = G.config("B", 1.0); cases.append((6, cfg["src_lon"], cfg["src_lat"]))
    cfg = G.config("C", 1.0); cases.append((4, cfg["src_lon"], cfg["src_lat"]))
    cfg = G.config("E", 0.03); Xt, Yt = G.trg_2d(cfg); cases.append((0, Xt, Yt))
    lo = np.where(Xt > 180.0, Xt - 360.0, Xt)
    cases.append((0, lo - 178.0 + (-179.9 - (lo - 178.0).min()) * ((lo - 178.0).min() < -180.0), Yt))
    for nperio, lo, la in cases:
        args = _staggered(lo, la)
        a = gpu.angle2(nperio, *args)
        b = orc.angle2(nperio, *args)
        for q in range(8):
            assert np.array_equal(np.isnan(a[q]), np.isnan(b[q])), (nperio, q)
            ok = ~np.isnan(a[q])
            assert np.array_equal(a[q][ok].view(np.uint64), b[q][ok].view(np.uint64)), (nperio, q, np.nanmax(np.abs(a[q] - b[q])))


# ---- AKIMA_1D_1D (column variant) and the generic-coordinate branch of INTERP_3D
def _columns(seed, ncol, nk1, nk2):
    rng = np.random.default_rng(seed)
    vz1 = np.cumsum(rng.uniform(3.0, 200.0, (ncol, nk1)), axis=1)
    vf1 = 20.0 - 0.004 * vz1 + rng.normal(0.0, 0.5, (ncol, nk1))
    vz2 = np.sort(rng.uniform(0.0, vz1.max() * 1.05, (ncol, nk2)), axis=1)
    vz2[:, 1] = vz1[:, 0]
    if nk1 > 3 and nk2 > 4:
        vz2[:, 3] = vz1[:, 2]
        vz2[:, 4] = vz1[:, 3].astype(np.float32).astype(np.float64) + 1e-9
    return vz1, vf1, vz2


@pytest.mark.parametrize("ncol,nk1,nk2", [(3000, 31, 75), (17, 3, 4), (500, 46, 12), (1, 8, 8)])
@pytest.mark.parametrize("opts", [dict(), dict(l_surf_persist=True, l_botm_persist=True), dict(l_botm_persist=True)])
def test_akima_1d_1d_vs_oracle(gpu, orc, ncol, nk1, nk2, opts):
    vz1, vf1, vz2 = _columns(ncol + nk1, ncol, nk1, nk2)
    vf1[::7] = 3.25                                          # flat columns: the equal-slope branch
    a, sa = gpu.akima_1d_1d(vz1, vf1, vz2, **opts)
    b, sb = orc.akima_1d_1d(vz1, vf1, vz2, **opts)
    assert np.array_equal(sa, sb) and np.all(sa == 0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{(a != b).sum()} differ, max {np.abs(a - b).max()}"
    # one source depth vector / one target vector shared by every column (interp_to_hydro_section passes vdp_m like that)
    a, sa = gpu.akima_1d_1d(vz1[0], vf1, vz2[0], **opts)
    b, sb = orc.akima_1d_1d(vz1[0], vf1, vz2[0], **opts)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_akima_1d_1d_missing_values(gpu, orc):
    vz1, vf1, vz2 = _columns(5, 400, 20, 30)
    vf1[::2, -4:] = -9999.0                                   # trailing run
    vf1[1::4, :3] = -9999.0; vf1[1::4, -2:] = -9999.0         # leading + trailing
    vf1[3::8, :] = -9999.0                                    # everything missing: DOING NOTHING
    opts = dict(rmask_val=-9999.0, l_surf_persist=True, l_botm_persist=True)
    a, sa = gpu.akima_1d_1d(vz1, vf1, vz2, **opts)
    b, sb = orc.akima_1d_1d(vz1, vf1, vz2, **opts)
    assert np.array_equal(sa, sb) and set(np.unique(sa)) == {0, 1}
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # inputs for which the reference is undefined are refused, the status says which columns
    import sosie_b200
    vf1[6, 9] = -9999.0                                       # a hole in the middle
    with pytest.raises(sosie_b200.SosieError, match="missing values"):
        gpu.akima_1d_1d(vz1, vf1, vz2, **opts)
    b, sb = orc.akima_1d_1d(vz1, vf1, vz2, **opts)
    assert np.array_equal(gpu.last_status, sb) and sb[6] == -1
    with pytest.raises(sosie_b200.SosieError, match="fewer than 3"):
        gpu.akima_1d_1d(vz1[:, :2], np.abs(vf1[:, :2]), vz2)


def _sigma_cubes(seed, nk_src, nk_trg, nj, ni, src_sigma, trg_sigma):
    rng = np.random.default_rng(seed)
    H = rng.uniform(200.0, 4000.0, (nj, ni))
    s_src = np.linspace(0.02, 1.0, nk_src)[:, None, None] ** 1.5
    s_trg = np.linspace(0.0, 1.0, nk_trg)[:, None, None] ** 1.3
    depth_src = (s_src * H * rng.uniform(0.8, 1.0, (1, nj, ni))).astype(np.float32)
    depth_trg = (s_trg * H * rng.uniform(0.9, 1.2, (1, nj, ni))).astype(np.float32)
    data = (18.0 - 0.003 * depth_src + rng.normal(0, 0.3, depth_src.shape)).astype(np.float32)
    if src_sigma:
        depth_src = depth_src[::-1].copy(); data = data[::-1].copy()
    if trg_sigma:
        depth_trg = depth_trg[::-1].copy()
    nwet = rng.integers(0, nk_trg + 1, (nj, ni))
    mask = (np.arange(nk_trg)[:, None, None] < nwet[None]).astype(np.int32)
    prev = rng.normal(size=(nk_trg, nj, ni)).astype(np.float32)
    return depth_src, data, depth_trg, prev, mask


@pytest.mark.parametrize("src_sigma,trg_sigma,lmout", [(False, False, False), (True, False, False), (False, True, True), (True, True, True),
                                                       (False, False, True), (True, True, False)])
def test_interp3d_sigma_vs_oracle(gpu, orc, src_sigma, trg_sigma, lmout):
    ds, da, dt, prev, mask = _sigma_cubes(21, 31, 40, 60, 90, src_sigma, trg_sigma)
    a = gpu.interp3d_sigma(ds, da, dt, prev, mask, lmout, src_sigma, trg_sigma)
    b, st = orc.interp3d_sigma(ds, da, dt, prev, mask, lmout, src_sigma, trg_sigma)
    assert st == 0
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{(a != b).sum()} differ"
    if lmout:
        land = mask[0] != 1
        assert np.array_equal(a[:, land], prev[:, land])
    else:
        assert not np.array_equal(a, prev)


def test_interp3d_sigma_on_device_after_horizontal(gpu, orc):
    """z -> sigma job with everything resident: bilinear per level into data3d_tmp, source depths carried to the target grid
    the same way (as INTERP_3D does, src/mod_interp.f90:300-320), then the column stage; nothing leaves HBM in between."""
    import torch
    rng = np.random.default_rng(8)
    lon_s = (np.arange(72) + 0.5) * 5.0
    lat_s = -90.0 + (np.arange(36) + 0.5) * 5.0
    lon_t = np.linspace(20.0, 120.0, 50)[None, :].repeat(30, 0) + np.linspace(0, 3.0, 30)[:, None]
    lat_t = np.linspace(-40.0, 40.0, 30)[:, None].repeat(50, 1) + np.linspace(0, 2.0, 50)[None, :]
    nk_src, nk_trg = 12, 20
    vz = np.cumsum(rng.uniform(5, 300, nk_src)).astype(np.float32)
    cube = (15.0 - 0.002 * vz[:, None, None] + rng.normal(0, 0.3, (nk_src, 36, 72))).astype(np.float32)
    depth_src_cube = np.broadcast_to(vz[:, None, None], cube.shape).astype(np.float32).copy()
    H = rng.uniform(300.0, float(vz[-1]) * 1.1, lon_t.shape)
    depth_trg = (np.linspace(0.01, 1.0, nk_trg)[:, None, None] * H).astype(np.float32)
    gpu.bilin_set_grids(0, lon_s, lat_s, lon_t, lat_t, l_reg_src=True)
    gpu.bilin_map()
    dev = torch.device("cuda:0")
    d_cube = torch.from_numpy(cube).to(dev); d_dsrc = torch.from_numpy(depth_src_cube).to(dev)
    d_tmp = torch.empty((nk_src, 30, 50), dtype=torch.float32, device=dev); d_dtmp = torch.empty_like(d_tmp)
    gpu.bilin_apply_dev(nk_src, d_cube, d_tmp)
    gpu.bilin_apply_dev(nk_src, d_dsrc, d_dtmp)
    d_dtrg = torch.from_numpy(depth_trg).to(dev)
    d_out = torch.zeros((nk_trg, 30, 50), dtype=torch.float32, device=dev)
    gpu.interp3d_sigma_dev(50, 30, nk_src, d_dtmp, d_tmp, nk_trg, d_dtrg, None, False, False, True, d_out)
    gpu.sync()
    tmp = np.stack([orc.bilin_2d(0, lon_s, lat_s, cube[k], lon_t, lat_t, l_reg_src=True)[0][0] for k in range(nk_src)])
    dtmp = np.stack([orc.bilin_2d(0, lon_s, lat_s, depth_src_cube[k], lon_t, lat_t, l_reg_src=True)[0][0] for k in range(nk_src)])
    ref, st = orc.interp3d_sigma(dtmp, tmp, depth_trg, np.zeros((nk_trg, 30, 50), np.float32), None, False, False, True)
    assert st == 0
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), ref.view(np.uint32))
