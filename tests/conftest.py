import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def cpu_backend(monkeypatch):
    """Host-logic tests: route the C-ABI dispatcher to oracle/abi_reference.py (CPU)."""
    from gnnepcsaft_b200 import _lib
    from gnnepcsaft_b200.train import models
    from oracle import abi_reference

    abi_reference.install(monkeypatch, _lib, models)
    try:
        from gnnepcsaft_b200.train import engine
        monkeypatch.setattr(engine, "_require_cuda", lambda t, what: None, raising=False)
    except ImportError:
        pass
    return abi_reference
