"""GPU parity at the BASELINE.json full sizes.  Small configs (A, B, C) are compared with the oracle in full; the large
ones (D, E) on a spread sample of target rows (the per-target computation is independent of the other rows on a
regular source) plus size-independent properties: index ranges, nearest-point optimality, weights in [0,1], and
exact reproduction of fields that are linear in (lon, lat)."""
import numpy as np
import pytest

from sosie_b200 import grids as G

pytestmark = pytest.mark.gpu


def _sample_rows_vs_oracle(gpu, orc, cfg, mg, Z1, Z2g, nrows):
    ny2 = cfg["nyt"]
    jj = np.unique(np.linspace(0, ny2 - 1, nrows).round().astype(int))
    Xt, Yt = G.trg_2d(cfg)
    Z2o, mo = orc.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, np.ascontiguousarray(Xt[jj]), np.ascontiguousarray(Yt[jj]),
                           l_reg_src=cfg["l_reg_src"], nthreads=orc.max_threads(), virtual_src=True)
    assert np.array_equal(mg["metrics"][:, jj], mo["metrics"]), "mapping indices differ on the sampled rows"
    assert np.array_equal(mg["ipb"][jj], mo["ipb"])
    assert np.array_equal(mg["alphabeta"][:, jj], mo["alphabeta"])
    assert np.array_equal(mg["dist_np"][jj], mo["dist_np"])
    assert np.array_equal(Z2g[:, jj], Z2o)


def _properties(cfg, mg, rng, nprobe=20000):
    met, ab, ipb = mg["metrics"], mg["alphabeta"], mg["ipb"]
    nxs, nys = cfg["nxs"], cfg["nys"] + (2 if mg["info"][7] else 0)
    assert met[0].min() >= 1 and met[0].max() <= nxs and met[1].min() >= 1 and met[1].max() <= nys
    assert set(np.unique(met[2])) <= {1, 2, 3, 4}
    assert ab.min() >= 0.0 and ab.max() <= 1.0
    assert set(np.unique(ipb)) <= {0, 1, 4, 5, 6, 7, 11, 12}
    assert (ipb == 0).mean() > 0.99
    # nearest-point optimality on a random probe: no 8-neighbour of (iP,jP) is closer (great circle, float64)
    Xt, Yt = G.trg_2d(cfg)
    lat_s = np.concatenate([cfg["src_lat"], [90.0, 90.0 + cfg["src_lat"][-1] - cfg["src_lat"][-2]]]) if mg["info"][7] else cfg["src_lat"]
    lon_s = cfg["src_lon"]
    k = rng.integers(0, Xt.size, nprobe)
    tj, ti = np.unravel_index(k, Xt.shape)
    iP, jP = met[0][tj, ti] - 1, met[1][tj, ti] - 1
    d2r = np.pi / 180.0

    def dist(i, j):
        lo, la = lon_s[np.clip(i, 0, nxs - 1)] * d2r, lat_s[np.clip(j, 0, nys - 1)] * d2r
        tlo, tla = Xt[tj, ti] * d2r, Yt[tj, ti] * d2r
        c = np.cos(tlo) * np.cos(tla) * np.cos(lo) * np.cos(la) + np.sin(tlo) * np.cos(tla) * np.sin(lo) * np.cos(la) + np.sin(tla) * np.sin(la)
        return np.arccos(np.minimum(c, 1.0))
    d0 = dist(iP, jP)
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            ok = (iP + di >= 0) & (iP + di < nxs) & (jP + dj >= 0) & (jP + dj < nys)
            assert (d0[ok] <= dist(iP + di, jP + dj)[ok] + 1e-12).all()


def _linear_field_reproduced(gpu, cfg, mg, a, b, c):
    """f = a*lon + b*lat + c on the source; well-mapped targets away from the 0/360 seam get f(lon_trg, lat_trg)."""
    Xs, Ys = np.meshgrid(cfg["src_lon"], cfg["src_lat"])
    Z1 = (a * Xs + b * Ys + c).astype(np.float32)
    Z2 = gpu.bilin_apply(Z1)[0]
    Xt, Yt = G.trg_2d(cfg)
    ref = a * Xt + b * Yt + c
    lon0, lon1 = cfg["src_lon"][1], cfg["src_lon"][-2]
    good = (mg["ipb"] == 0) & (Xt > lon0) & (Xt < lon1) & (mg["metrics"][1] < cfg["nys"] - 1) & (mg["metrics"][1] > 1)
    err = np.abs(Z2 - ref)[good] / np.maximum(1.0, np.abs(ref[good]))
    assert good.mean() > 0.5 and err.max() < 5e-5, err.max()     # REAL(4) weights + 4-term sum: a few float32 ulps of the field magnitude


def test_config_E_full(gpu, orc):
    cfg = G.config("E", 1.0)
    rng = np.random.default_rng(20260101)
    Z1 = (np.cos(4 * np.deg2rad(cfg["src_lon"]))[None, :] * np.sin(4 * np.deg2rad(cfg["src_lat"]))[:, None] * 5000 - 1000).astype(np.float32)
    gpu.l_first_call_interp_routine = True
    Z2 = gpu.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
    mg = gpu.bilin_get_map()
    assert mg["info"][5] == 0 and mg["info"][7] == 1          # EASY search, pole rows added
    _properties(cfg, mg, rng)
    _sample_rows_vs_oracle(gpu, orc, cfg, mg, Z1[None], Z2, 32)
    _linear_field_reproduced(gpu, cfg, mg, 0.37, -1.3, 11.0)
    # sharded mapping (multi-GPU path) on this one device: rows 1000..1499 only, identical to the unsharded rows
    gpu.bilin_set_shard(1000, 1499)
    try:
        gpu.bilin_map()
        ms = gpu.bilin_get_map()
        for k in ("metrics", "alphabeta"):
            assert np.array_equal(ms[k][:, 999:1499], mg[k][:, 999:1499])
        assert np.array_equal(ms["ipb"][999:1499], mg["ipb"][999:1499])
    finally:
        gpu.bilin_set_shard(0, 0)


def test_config_D_full(gpu, orc):
    cfg = G.config("D", 1.0)
    rng = np.random.default_rng(7)
    Xs, Ys = G.src_2d(cfg)
    Z1 = G.field_slabs(Xs, Ys, 9, lo=230.0, hi=310.0, seed=5)
    gpu.l_first_call_interp_routine = True
    Z2 = gpu.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
    mg = gpu.bilin_get_map()
    _properties(cfg, mg, rng)
    _sample_rows_vs_oracle(gpu, orc, cfg, mg, Z1, Z2, 12)
    _linear_field_reproduced(gpu, cfg, mg, -0.21, 0.9, 250.0)
    # batched (staged, 9 slabs) and single-slab kernels agree bit for bit
    one = np.concatenate([gpu.bilin_apply(Z1[s]) for s in range(9)])
    assert np.array_equal(one, Z2)


@pytest.mark.parametrize("name", ["B", "C"])
def test_config_BC_full_vs_oracle(gpu, orc, name):
    cfg = G.config(name, 1.0)
    Xs, Ys = G.src_2d(cfg)
    Z1 = G.field_slabs(Xs, Ys, 3, seed=2)
    gpu.l_first_call_interp_routine = True
    Z2 = gpu.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"], l_reg_src=False,
                      i_orca_trg=cfg["i_orca_trg"])
    mg = gpu.bilin_get_map()
    Z2o, mo = orc.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"], l_reg_src=False,
                           i_orca_trg=cfg["i_orca_trg"])
    assert mg["info"][5] == 1                                  # TWISTED search
    for k in ("metrics", "ipb", "alphabeta", "dist_np"):
        assert np.array_equal(mg[k], mo[k]), k
    assert np.array_equal(Z2, Z2o)


def test_config_C_full_bdrown(gpu, orc):
    cfg = G.config("C", 1.0)
    S = 31
    X = G.field_slabs(cfg["src_lon"], cfg["src_lat"], S, seed=3)
    masks = np.stack([G.land_mask(cfg["src_lon"], cfg["src_lat"], thresh=0.25 - 0.4 * k / 31) for k in range(S)])
    masks[-1] = 0
    X[masks == 0] = -9999.0
    Xo, no = orc.bdrown(cfg["ewper_src"], X, masks, cfg["nb_inc"], cfg["nb_smooth"], cfg["vmin"], cfg["vmax"], 0.0, nthreads=orc.max_threads())
    for general in (0, 1, 2):
        gpu.bdrown_force_general(general)
        try:
            Xg, ng = gpu.bdrown(cfg["ewper_src"], X, masks, cfg["nb_inc"], cfg["nb_smooth"], cfg["vmin"], cfg["vmax"], 0.0)
        finally:
            gpu.bdrown_force_general(False)
        assert np.array_equal(ng, no) and np.array_equal(Xg, Xo), general


def test_config_A_full_akima_bdrown(gpu, orc):
    cfg = G.config("A", 1.0)
    Xs, Ys = G.src_2d(cfg)
    mask = G.land_mask(Xs, Ys)
    Z1 = G.field_slabs(Xs, Ys, 2, seed=4)
    Z1[:, mask == 0] = -9999.0
    Xg, ng = gpu.bdrown(cfg["ewper_src"], Z1, mask, cfg["nb_inc"], cfg["nb_smooth"], cfg["vmin"], cfg["vmax"], 0.0)
    Xo, no = orc.bdrown(cfg["ewper_src"], Z1, mask, cfg["nb_inc"], cfg["nb_smooth"], cfg["vmin"], cfg["vmax"], 0.0)
    assert np.array_equal(ng, no) and np.array_equal(Xg, Xo)
    Zg = gpu.akima_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Xg, cfg["trg_lon"], cfg["trg_lat"])
    Zo, ixy = orc.akima_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Xo, cfg["trg_lon"], cfg["trg_lat"])
    assert np.array_equal(gpu.akima_get_ixy(), ixy) and np.array_equal(Zg, Zo)
