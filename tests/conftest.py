import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ensure_native_library():
    """A fresh checkout has no built artefacts (they are git-ignored): build the product library once, before collection —
    test modules ask it for the device count in their skip conditions.  nvcc cross-compiles without a GPU (~1 min)."""
    from odgi_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        from odgi_b200.build import build_native
        build_native()


_ensure_native_library()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def golden_graphs():
    """name -> dict of arrays (flattened graph + XP sampling tables dumped by the reference)."""
    from odgi_b200.arrays import read_arrays
    out = {}
    for name in ("note5", "t", "overlap", "k", "DRB1-3123", "chr6.C4", "LPA"):
        out[name] = read_arrays(os.path.join(GOLDEN, f"{name}.graph.arr.gz"))
    return out
