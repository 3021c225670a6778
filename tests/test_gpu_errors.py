"""Error behaviour through the C ABI: where the reference PRINTs and STOPs the library returns a code and the reference's
message (the Fortran shim prints it and STOPs); call-order and argument errors have their own codes."""
import numpy as np
import pytest

import sosie_b200
from sosie_b200 import grids as G

pytestmark = pytest.mark.gpu


def test_reference_stop_conditions(gpu):
    # FIND_NEAREST_TWISTED: "Target and source latitudes do not overlap!" (src/mod_manip.f90:1127-1131)
    cfg = G.config("B", 0.25)
    lon_t = np.linspace(10.0, 50.0, 30)[None, :].repeat(12, 0) + np.linspace(0, 3, 12)[:, None]
    lat_t = (-89.9 + 0.005 * np.abs(np.arange(12) - 5))[:, None].repeat(30, 1) + 0.001 * np.arange(30)[None, :]   # south of the ORCA grid, not monotone in j
    gpu.bilin_set_grids(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], lon_t, lat_t)
    with pytest.raises(sosie_b200.SosieError, match="do not overlap"):
        gpu.bilin_map()
    # SMOOTHER: l_exclude_mask_points without a mask (src/mod_bdrown.f90:499-503)
    X = np.zeros((8, 9), np.float32)
    with pytest.raises(sosie_b200.SosieError, match="need to provide"):
        gpu.smoother(0, X, 3, None, True)
    # L_InPoly: negative longitudes (src/mod_poly.f90:79)
    lon_s = np.linspace(-20.0, 20.0, 41); lat_s = np.linspace(-10.0, 10.0, 21)
    Xs, Ys = G.expand_2d(lon_s, lat_s)
    Xt, Yt = G.expand_2d(np.linspace(-5.0, 5.0, 7), np.linspace(-3.0, 3.0, 5))
    Xt = Xt + 0.01 * Yt   # irregular target, regular source -> EASY search, then L_InPoly sees the negative longitudes
    gpu.bilin_set_grids(-1, Xs, Ys, Xt, Yt)
    with pytest.raises(sosie_b200.SosieError, match="positive longitudes"):
        gpu.bilin_map()


def test_call_order_and_arguments():
    ctx = sosie_b200.Sosie(0)
    try:
        with pytest.raises(sosie_b200.SosieError, match="error 3"):     # SOSIE_ERR_STATE: no mapping yet
            ctx._shape = (10, 10, 5, 5)
            ctx.bilin_get_map()
        with pytest.raises(sosie_b200.SosieError, match="error 2"):     # SOSIE_ERR_ARG: extents
            ctx.bilin_set_grids(0, np.zeros(2), np.zeros(2), np.zeros(4), np.zeros(4))
        with pytest.raises(sosie_b200.SosieError, match="error 2"):     # AKIMA_1D_3D needs >= 3 source levels for its frame
            ctx.akima_1d_3d(np.array([1.0, 2.0], np.float32), np.zeros((2, 3, 3), np.float32), np.array([1.5], np.float32))
        # DROWN on a grid narrower than its periodic folding allows (src/mod_drown.f90:110-117 would index maskv out of bounds)
        with pytest.raises(sosie_b200.SosieError, match="2\\*k_ew"):
            ctx.drown(2, np.zeros((6, 4), np.float32), np.ones((6, 4), np.int8), 3)
    finally:
        ctx.close()
    with pytest.raises(sosie_b200.SosieError):
        sosie_b200.Sosie(99)        # no such device


def test_shard_changes_invalidate_a_partial_mapping(gpu):
    """the EASY search fills the mapping only for the rows of the active shard; moving the shard outside them must
    not let an apply pack rows that were never computed, and a shard beyond the grid is an error, not "all rows"."""
    cfg = G.config("E", 0.02)
    Xs, Ys = G.src_2d(cfg)
    Z1 = G.field_slabs(Xs, Ys, 1, seed=3)
    ny2 = cfg["nyt"]
    gpu.bilin_set_shard(0, 0)
    gpu.bilin_set_grids(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
    gpu.bilin_map()
    full = gpu.bilin_apply(Z1, first_call=True)
    try:
        with pytest.raises(sosie_b200.SosieError, match="exceeds"):
            gpu.bilin_set_shard(1, ny2 + 1)
        gpu.bilin_set_shard(10, 40)
        gpu.bilin_map()                                  # mapping now holds rows 10..40 only
        gpu.bilin_set_shard(20, 30)                      # contained: the mapping is kept
        part = gpu.bilin_apply(Z1)
        assert np.array_equal(part[0, 19:30], full[0, 19:30])
        gpu.bilin_set_shard(35, 60)                      # leaves the mapped rows: the mapping is dropped
        with pytest.raises(sosie_b200.SosieError, match="error 3"):
            gpu.bilin_apply(Z1)
        gpu.bilin_map()
        part = gpu.bilin_apply(Z1)
        assert np.array_equal(part[0, 34:60], full[0, 34:60])
    finally:
        gpu.bilin_set_shard(0, 0)
