import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_MODELS = "/root/reference/models"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def have_ref():
    return os.path.isdir(REF_MODELS)


needs_ref = pytest.mark.skipif(not have_ref(), reason="/root/reference not present (GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def golden_model_path(name):
    return os.path.join(GOLDEN, name + ".pbdm")


_flat_cache = {}


def load_flat(name):
    """FlatModel of a committed golden model (.pbdm), read with the product loader."""
    if name not in _flat_cache:
        from partsbaseddetector_b200 import Model
        _flat_cache[name] = Model.load_bin(golden_model_path(name)).to_flat()
    return _flat_cache[name]
