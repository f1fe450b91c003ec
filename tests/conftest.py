import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle32():
    from oracle.oracle import Oracle
    o = Oracle("f32")
    o.set_threads(1)  # deterministic accumulation order
    return o


@pytest.fixture(scope="session")
def oracle64():
    from oracle.oracle import Oracle
    o = Oracle("f64")
    o.set_threads(1)
    return o
