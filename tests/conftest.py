import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def ref_backend():
    """The reference's own operator extension compiled unmodified for sm_100a (oracle/build_ref.py);
    present on the GPU box as oracle/_ref/_pvcnn_backend.so, None when it was not built."""
    from oracle import build_ref
    try:
        return build_ref.load()
    except Exception:
        return None
