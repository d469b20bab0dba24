import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a B200 skips the gpu-marked tests instead of failing in their fixtures
    (the product path has no CPU fallback); `-m gpu` on the GPU box runs them."""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); the product path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ksw_ctx():
    from pansvr_b200 import ksw
    ctx = ksw.KswContext(0)
    yield ctx
    ctx.close()
