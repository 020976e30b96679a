import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build(ref=False)
    return oracle.orc


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.ref.available():
        if os.path.isdir(oracle.REFERENCE_TREE):
            oracle.build(ref=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return oracle.ref
