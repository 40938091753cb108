import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device AND the built library: skip them (instead of failing by the hundred) elsewhere, so
    that a plain `pytest tests` is a usable signal on a CPU-only machine."""
    import torch
    lib = os.path.join(ROOT, "geometry_rl_b200", "libgrl_b200.so")
    if torch.cuda.is_available() and os.path.exists(lib):
        return
    why = "no CUDA device" if not torch.cuda.is_available() else "libgrl_b200.so is not built"
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
