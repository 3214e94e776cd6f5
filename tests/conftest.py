import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pi-consistency-activity-detection_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def pytest_sessionstart(session):
    """The suite exercises the in-tree C-ABI library: build it (nvcc, sm_100a, ~1 min, no GPU needed) when a fresh
    checkout does not have it yet.  A failure here is reported by tests/test_abi.py, not swallowed."""
    try:
        from b200caps import _abi, build
        if not os.path.isfile(_abi.LIB_PATH):
            build.build()
    except Exception as e:  # noqa: BLE001 - reported through the ABI tests
        sys.stderr.write(f"[conftest] could not build libb200caps.so: {e}\n")
