import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure), built on demand."""
    from oracle import oracle as O
    O.build()
    O.set_libm(O.PMATH)
    return O


@pytest.fixture(scope="session")
def gpu():
    """A sosie_b200 context on cuda:0.  The library must be there: no fallback."""
    import sosie_b200
    if not os.path.exists(sosie_b200.lib_path()):
        from sosie_b200 import build
        build.build()
    ctx = sosie_b200.Sosie(0)
    yield ctx
    ctx.close()
