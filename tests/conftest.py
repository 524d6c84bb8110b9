import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device; run with `-m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed():
    # the reference seeds everything with 12345 (tests/conftest.py:43-45, configs/config.yaml:3)
    torch.manual_seed(12345)
    yield


@pytest.fixture(scope="session")
def lib():
    from myria3d_b200 import _lib
    from myria3d_b200.build import build_library

    build_library()
    return _lib.load()
