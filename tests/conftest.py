import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        import micropp_b200
        return micropp_b200.LIB_PATH.exists() and micropp_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # gpu tests are only deselected by -m; if someone runs them without a device, fail loudly, never skip silently
    pass


@pytest.fixture(autouse=True)
def _finalize_after_each_test():
    """Objects that own device memory (Micropp3, SlabRVE, the compiled reference) are finalized when THEIR test ends,
    not at some later allocation of another test: a crash in a destructor is then reported against its own test."""
    yield
    import gc
    gc.collect()


@pytest.fixture(scope="session")
def refpy():
    """The compiled reference (oracle/_ref) -- the checker."""
    from oracle import refpy as R
    if not R.available():
        pytest.skip("oracle/_ref/libmicropp_ref.so not built (needs /root/reference: make -C oracle ref)")
    return R


@pytest.fixture(scope="session")
def mpp():
    import micropp_b200
    micropp_b200.load()
    return micropp_b200
