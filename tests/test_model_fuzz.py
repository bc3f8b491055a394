"""Robustness of the model readers (MAT-v5, opencv_storage XML / YAML, .pbdm): mutated and truncated files must either load or fail
with a PbdError -- never crash, hang or allocate without bound.  (The interpreter would die with the test on a crash.)"""
import os

import numpy as np
import pytest
import scipy.io

from conftest import GOLDEN, load_flat
from partsbaseddetector_b200 import FileStorageModel, MatlabIOModel, Model, PbdError
from test_matlab_model import matlab_model_dict


def mutate(raw, rng, t):
    b = bytearray(raw)
    kind = t % 4
    if kind == 0:                                     # a few random bytes
        for _ in range(int(rng.integers(1, 8))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
    elif kind == 1:                                   # truncation
        b = b[: int(rng.integers(0, len(b)))]
    elif kind == 2:                                   # a random 32-bit word (sizes, counts, tags in the binary formats)
        i = int(rng.integers(0, len(b) - 8))
        b[i:i + 4] = int(rng.integers(0, 2 ** 32)).to_bytes(4, "little")
    else:                                             # a digit replaced by another digit (text formats stay well formed)
        idx = [i for i in rng.integers(0, len(b), 64) if 48 <= b[i] <= 57]
        for i in idx[:3]:
            b[i] = 48 + int(rng.integers(0, 10))
    return bytes(b)


def run(raw, load, path, n, seed):
    rng = np.random.default_rng(seed)
    ok = 0
    for t in range(n):
        with open(path, "wb") as f:
            f.write(mutate(raw, rng, t))
        try:
            load(path)
            ok += 1
        except PbdError as e:
            assert e.code in (-2, -3, -1), e           # I/O, format or argument (validation) errors only
    return ok


@pytest.mark.parametrize("compress", [False, True])
def test_mutated_mat_files(tmp_path, compress):
    fm = load_flat("Willowcoffee_5parts")
    src = str(tmp_path / "m.mat")
    scipy.io.savemat(src, {"model": matlab_model_dict(fm), "name": "x"}, do_compression=compress)
    ok = run(open(src, "rb").read(), lambda p: MatlabIOModel().deserialize(p), str(tmp_path / "f.mat"), 200, 1)
    assert compress or ok > 0                          # some mutations of an uncompressed file only touch weights


@pytest.mark.parametrize("ext", [".xml", ".yml"])
def test_mutated_filestorage_files(tmp_path, ext):
    m = FileStorageModel.load_bin(os.path.join(GOLDEN, "Willowcoffee_5parts.pbdm"))
    src = str(tmp_path / ("m" + ext))
    assert m.serialize(src)
    ok = run(open(src, "rb").read(), lambda p: FileStorageModel().deserialize(p), str(tmp_path / ("f" + ext)), 200, 2)
    assert ok > 0


def test_mutated_pbdm_files(tmp_path):
    raw = open(os.path.join(GOLDEN, "Willowcoffee_5parts.pbdm"), "rb").read()
    run(raw, lambda p: Model.load_bin(p), str(tmp_path / "f.pbdm"), 200, 3)
