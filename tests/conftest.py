import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "covid-spings-variant-caller_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_testfile():
    with open(os.path.join(GOLD, "testfile_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_synth():
    with open(os.path.join(GOLD, "synthetic_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_utils():
    with open(os.path.join(GOLD, "utils_known_answers.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (build it if the .so is missing: nvcc cross-compiles without a GPU)."""
    from lvc_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return capi.load_library()


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
