"""Shared fixtures.  GPU tests are marked ``gpu`` and call the CUDA path through the C ABI;
everything else runs on CPU (oracle vs golden vectors, host logic, symbol checks)."""

from __future__ import annotations

import os

# see spectralclustersupertree_b200/_lib.py: load every kernel before ranks start waiting for each other on the device
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def engine():
    """One GPU context for the whole session; building the library if it is stale."""
    from spectralclustersupertree_b200 import build
    from spectralclustersupertree_b200.engine import Engine

    build.build()
    eng = Engine(0)
    yield eng
    eng.close()
