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
def _build_everything_once():
    """The oracle's C transcription and (if it did not travel with the snapshot) the CUDA library:
    nvcc cross-compiles sm_100a without a GPU, so the CPU suite can check the ABI."""
    from oracle import c_oracle
    c_oracle.lib()
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ge.PKG_DIR, "lib", "libebpmf_b200.so")):
        ge.build()
