"""Randomised GPU-vs-oracle parity (tools/fuzz_parity.py): random grids (regular axes cell- or node-centred, with overlap
columns, regional; ORCA-shaped curvilinear), random targets (1-D axes, rotated/sheared, polar caps), masks, periodicity,
sequential and pipelined first calls, Akima, BDROWN (all three implementations) / DROWN / SMOOTHER with random masks and
options, the vertical stage (AKIMA_1D_3D / _1D with missing values, the sigma branch, lbc_lnk, post-ops, extrp_hl), rotation.  Results must be bit-identical; where the oracle says the reference would STOP the library must raise as well."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("seed,ncases", [(11, 160), (12, 160)])
def test_fuzz_parity(gpu, orc, seed, ncases):
    """the first `ncases` cases of `python tools/fuzz_parity.py N seed` (same sub-seeds): 500 cases take ~3 s on the box"""
    import fuzz_parity as F
    rng = np.random.default_rng(seed)
    fails, done, stops = [], 0, 0
    for k in range(ncases):
        sub = np.random.default_rng(rng.integers(1 << 62))
        fn = F.CASES[k % len(F.CASES)]
        try:
            tag = fn(gpu, sub, k)
            done += tag is not None
            stops += bool(tag) and tag.endswith("(both STOP)")
        except AssertionError as e:
            fails.append(str(e))
    gpu.set_pipeline_min_targets(1 << 20)
    gpu.bdrown_force_general(0)
    gpu.akima_untiled_slopes(False)
    gpu.l_first_call_interp_routine = True
    assert not fails, fails[:5]
    assert done > 0.85 * ncases and stops < done // 4, (done, stops)
