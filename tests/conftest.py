import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libccv_ref.so (the compiled reference)")


def _has_gpu():
    try:
        from ccv_b200 import nnc
        return nnc.lib().ccv_nnc_device_count(nnc.CCV_STREAM_CONTEXT_GPU) > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """The GPU tests must run the native library: fail (not skip) if it cannot be loaded on a GPU box."""
    from ccv_b200 import nnc
    nnc.init()
    if not _has_gpu():
        pytest.skip("no CUDA device")
    return nnc


@pytest.fixture(scope="session")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libccv_ref.so not built (make -C oracle)")
    r.ref()
    return r
