import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def gpu_ctx(pkg):
    """One context on cuda:0 for the whole GPU session.  Fails loudly (no CPU fallback)."""
    ctx = pkg.VgaContext(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import c_oracle
    c_oracle.lib()
