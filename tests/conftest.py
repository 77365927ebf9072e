import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _has_gpu():
    import tg_b200
    try:
        return tg_b200.lib().tgb200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """GPU tests FAIL (not skip) when libtgb200.so is missing or sees no device: no silent fallback."""
    import tg_b200
    tg_b200.lib()
    assert _has_gpu(), "no CUDA device visible to libtgb200.so"
    return True


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O
