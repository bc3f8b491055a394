"""The C-ABI shared library loads and exports every symbol include/pbd_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from partsbaseddetector_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pbd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbd_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = ctypes.CDLL(_lib.SO_PATH)
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.EXPORTS) == names


def test_version_and_error_strings():
    L = _lib.lib()
    assert b"sm_100a" in L.pbd_version()
    assert L.pbd_model_nparts(None, 0) < 0
    assert L.pbd_last_error() != b""


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from conftest import load_flat
    from partsbaseddetector_b200 import Model, PartsBasedDetector, PbdError
    det = PartsBasedDetector()
    with pytest.raises(PbdError) as e:
        det.distributeModel(Model.from_flat(load_flat("Willowcoffee_5parts")))
    assert e.value.code == -4          # PBD_E_CUDA: no CPU fallback exists
