"""The batched apply splits the target tiles between three kernels (csrc/bilin_apply.cu): the group-staged kernel (16-byte
cp.async, 8 / 4 / 2 / 1 slabs per block barrier depending on the tile's box: the default), the 4-byte cp.async kernel (any alignment, odd boxes) and the TMA-staged
kernel (cp.async.bulk.tensor through a 3-D tensor map; off by default, see profiles/r02_tma_nvproxy.md).  Every combination
must give the same bits as the oracle.  Also: the reciprocal-based correctly rounded division of these kernels against the
IEEE division."""
import os

import numpy as np
import pytest

from sosie_b200 import grids as G

pytestmark = pytest.mark.gpu

TMA_OK = os.environ.get("SOSIE_TEST_TMA") == "1"     # cp.async.bulk.tensor traps under the nvproxy driver of this pool


def _variants(gpu, Z1):
    out = {}
    modes = {"grp": 0, "old": 4}
    if TMA_OK:
        modes.update({"tma+grp": 1, "tma+old": 5})
    try:
        for name, m in modes.items():
            gpu.apply_tma(m)
            out[name] = gpu.bilin_apply(Z1)
    finally:
        gpu.apply_tma(0)
    return out, gpu.apply_tile_classes()


@pytest.mark.parametrize("name,scale,nslab", [("D", 0.25, 4), ("D", 0.25, 7), ("D", 0.125, 33), ("D", 0.5, 9), ("E", 0.02, 5), ("A", 0.5, 6),
                                               ("D", 0.25, 2), ("D", 1.0, 5), ("D", 0.25, 19), ("D", 0.5, 16), ("E", 0.02, 24)])
def test_batched_apply_kernels_agree_with_oracle(gpu, orc, name, scale, nslab):
    cfg = G.config(name, scale)
    assert cfg["nxs"] % 4 == 0               # 16-byte row pitch: the group-staged kernel is eligible
    Xs, Ys = G.src_2d(cfg)
    Z1 = G.field_slabs(Xs, Ys, nslab, lo=-50.0, hi=310.0, seed=21)
    Z1[:, ::7, ::5] = 0.0                    # zeros (both signs) and tiny / huge values: the division's guarded ranges
    Z1[:, 1::9, 2::11] = -0.0
    Z1[:, 3::13, ::17] = 1.0e-38
    Z1[:, 5::13, 1::17] = -3.0e37
    gpu.bilin_set_shard(0, 0)
    gpu.l_first_call_interp_routine = True
    Z2 = gpu.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"], l_reg_src=cfg["l_reg_src"],
                      i_orca_trg=cfg["i_orca_trg"])
    if scale < 1.0:
        Z2o, _ = orc.bilin_2d(cfg["ewper_src"], cfg["src_lon"], cfg["src_lat"], Z1, cfg["trg_lon"], cfg["trg_lat"], l_reg_src=cfg["l_reg_src"],
                              i_orca_trg=cfg["i_orca_trg"])
    else:                                    # full size: the single-slab kernel (IEEE division, direct gathers) is the reference
        Z2o = np.concatenate([gpu.bilin_apply(Z1[s]) for s in range(nslab)])
    res, cls = _variants(gpu, Z1)
    assert cls[1] > 0, f"no tile fits the group-staged kernel: {cls}"
    for k, v in res.items():
        assert np.array_equal(v.view(np.uint32), Z2o.view(np.uint32)), f"kernel set '{k}' differs from the reference at {(v != Z2o).sum()} points"
    assert np.array_equal(Z2.view(np.uint32), Z2o.view(np.uint32))


def test_unaligned_sources_fall_back(gpu, orc):
    """source width not a multiple of 4 (no 16-byte row pitch): only the 4-byte cp.async kernel runs, results unchanged"""
    cfg = G.config("D", 0.25)
    lon, lat = G.regular_latlon(cfg["nxs"] + 1, cfg["nys"])
    Xs, Ys = G.expand_2d(lon, lat)
    for nslab in (2, 3, 6):
        Z1 = G.field_slabs(Xs, Ys, nslab, seed=5)
        gpu.l_first_call_interp_routine = True
        Z2 = gpu.bilin_2d(0, lon, lat, Z1, cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
        Z2o, _ = orc.bilin_2d(0, lon, lat, Z1, cfg["trg_lon"], cfg["trg_lat"], l_reg_src=True)
        assert np.array_equal(Z2, Z2o)


def test_division_by_reciprocal_is_ieee(gpu):
    """div_by_rcp (csrc/bilin_apply.cu) == __fdiv_rn bit for bit on 2^28 operand pairs: x any REAL(4) bit pattern (3/8 of them
    packed around the guard thresholds and around 1), wup any REAL(4) in [0.5, 2]"""
    assert gpu.selftest_div(1 << 28, seed=1) == 0
    assert gpu.selftest_div(1 << 26, seed=0xC0FFEE) == 0
