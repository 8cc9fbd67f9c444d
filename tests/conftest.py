import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_library():
    """Build (if stale) and return the path of the in-tree shared library."""
    from jaxfluids_b200 import build
    return build.build()


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) on a machine without a CUDA device, so that a plain
    `python -m pytest tests` is green on CPU-only hosts; on a GPU box nothing is skipped here and a missing
    library still fails loudly inside the tests."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
