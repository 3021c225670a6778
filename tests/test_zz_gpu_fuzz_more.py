"""More of the randomised GPU-vs-oracle parity run (tools/fuzz_parity.py): the first 600 cases of three seeds whose 2500-case
runs were verified on the box.  Collected last on purpose: the targeted parity tests report first."""
import pytest

from tests.test_gpu_fuzz import test_fuzz_parity as _fuzz

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [41, 42, 43])
def test_fuzz_parity_more(gpu, orc, seed):
    _fuzz(gpu, orc, seed, 600)
