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


@pytest.fixture(scope="session")
def b200():
    """The product operator module (loads libg4s_rasterizer.so; no fallback)."""
    import g4splat_b200.diff_surfel_rasterization as op
    return op


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference extension built into oracle/_ref (GPU side-by-side runs)."""
    from oracle import build_ref
    if not build_ref.up_to_date():
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt files)")
    return build_ref.import_reference()
