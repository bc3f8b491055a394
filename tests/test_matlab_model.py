"""MatlabIOModel (native MAT-v5 reader, partsbaseddetector_b200/csrc/matfile.cpp) against the reference loader's semantics
(src/MatlabIOModel.cpp:71-188): a .mat written from a known model -- by scipy.io.savemat (compressed and not) and by a small
independent writer below (big endian, integer storage types for doubles, small data elements, as MATLAB itself writes them) --
must load into exactly the fields the XML loader gives."""
import os
import struct
import zlib

import numpy as np
import pytest
import scipy.io

from conftest import load_flat
from partsbaseddetector_b200 import FileStorageModel, MatlabIOModel, PbdError


def matlab_model_dict(fm):
    """The Matlab training code's `model` struct (matlab/learning, matlab/modelTransfer.m) for a FlatModel: 1-based ids, filters
    H x W x C, anchors [x y level], biasid as a (child mixtures x parent mixtures) matrix, the root without deformation."""
    nf = len(fm.filters)
    filters = np.zeros((1, nf), dtype=[("w", "O"), ("i", "O")])
    for i, f in enumerate(fm.filters):
        kh = f.shape[0]
        kw = f.shape[1] // fm.flen
        filters[0, i]["w"] = np.ascontiguousarray(f.reshape(kh, kw, fm.flen).astype(np.float64))
        filters[0, i]["i"] = float(i * 100 + 1)
    comps = np.empty((1, len(fm.comps)), dtype=object)
    for c, parts in enumerate(fm.comps):
        sa = np.zeros((1, len(parts)), dtype=[("defid", "O"), ("filterid", "O"), ("parent", "O"), ("biasid", "O")])
        for p, part in enumerate(parts):
            root = part.parentid < 0
            sa[0, p]["defid"] = np.zeros((0, 0)) if root else (np.asarray(part.defid, np.float64) + 1).reshape(1, -1)
            sa[0, p]["filterid"] = (np.asarray(part.filterid, np.float64) + 1).reshape(1, -1)
            sa[0, p]["parent"] = float(part.parentid + 1)
            rows = 1 if root else len(part.filterid)
            sa[0, p]["biasid"] = (np.asarray(part.biasid, np.float64) + 1).reshape(rows, -1)
        comps[0, c] = sa
    nd = len(fm.defs)
    defs = np.zeros((1, nd), dtype=[("w", "O"), ("i", "O"), ("anchor", "O")])
    for d in range(nd):
        defs[0, d]["w"] = np.asarray(fm.defs[d], np.float64).reshape(1, 4)
        defs[0, d]["i"] = float(d + 1)
        defs[0, d]["anchor"] = np.array([[fm.anchors[d][0] + 1.0, fm.anchors[d][1] + 1.0, 0.0]])
    bias = np.zeros((1, len(fm.biasw)), dtype=[("w", "O"), ("i", "O")])
    for b, w in enumerate(fm.biasw):
        bias[0, b]["w"] = float(w)
        bias[0, b]["i"] = float(b + 1)
    return {"interval": float(fm.interval), "thresh": float(fm.thresh), "sbin": float(fm.sbin), "filters": filters,
            "components": comps, "defs": defs, "bias": bias, "maxsize": np.array([[5.0, 5.0]]), "len": 12345.0}


def assert_same_model(got, fm, name):
    assert got.name == name
    assert (got.interval, got.sbin, got.flen, got.norient) == (fm.interval, fm.sbin, fm.flen, 18)
    assert np.float32(got.thresh) == np.float32(fm.thresh)
    assert len(got.filters) == len(fm.filters)
    for a, b in zip(got.filters, fm.filters):
        assert a.shape == b.shape and np.array_equal(a, b)
    assert np.array_equal(np.asarray(got.biasw, np.float32), np.asarray(fm.biasw, np.float32))
    assert np.array_equal(np.asarray(got.anchors), np.asarray(fm.anchors))
    assert np.array_equal(np.asarray(got.defs, np.float32), np.asarray(fm.defs, np.float32))
    assert len(got.comps) == len(fm.comps)
    for ca, cb in zip(got.comps, fm.comps):
        assert len(ca) == len(cb)
        for pa, pb in zip(ca, cb):
            assert (pa.parentid, list(pa.filterid), list(pa.biasid), list(pa.defid)) == (pb.parentid, list(pb.filterid), list(pb.biasid), list(pb.defid))


@pytest.mark.parametrize("name", ["Person_26parts", "Person_8parts", "Face_99filters", "Willowcoffee_5parts"])
@pytest.mark.parametrize("compress", [False, True])
def test_savemat_round_trip(tmp_path, name, compress):
    fm = load_flat(name)
    path = str(tmp_path / (name + ".mat"))
    scipy.io.savemat(path, {"model": matlab_model_dict(fm), "name": name, "pa": np.arange(5.0)}, do_compression=compress)
    m = MatlabIOModel()
    assert m.deserialize(path) is True
    assert_same_model(m.to_flat(), fm, name)
    assert m.serialize(str(tmp_path / "x.mat")) is False          # a stub in the reference as well


def test_name_defaults_to_file_stem_and_errors(tmp_path):
    fm = load_flat("Willowcoffee_5parts")
    path = str(tmp_path / "coffee_cup.mat")
    scipy.io.savemat(path, {"model": matlab_model_dict(fm)})
    m = MatlabIOModel()
    assert m.deserialize(path)
    assert m.name() == "coffee_cup"
    assert MatlabIOModel().deserialize(str(tmp_path / "missing.mat")) is False       # cannot open -> false, as the reference
    bad = tmp_path / "bad.mat"
    bad.write_bytes(b"not a mat file" * 20)
    with pytest.raises(PbdError):
        MatlabIOModel().deserialize(str(bad))
    scipy.io.savemat(str(tmp_path / "nomodel.mat"), {"x": np.eye(3)})
    with pytest.raises(PbdError):
        MatlabIOModel().deserialize(str(tmp_path / "nomodel.mat"))
    d = matlab_model_dict(fm)
    del d["bias"]
    scipy.io.savemat(str(tmp_path / "nobias.mat"), {"model": d})
    with pytest.raises(PbdError):
        MatlabIOModel().deserialize(str(tmp_path / "nobias.mat"))
    raw = open(path, "rb").read()
    # a cell array claiming 2^27 elements in a 100-byte element must be rejected, not allocated
    w = MatWriter(False)
    huge = w.element(6, struct.pack("<II", 1, 0)) + w.element(5, struct.pack("<2i", 1, 1 << 27)) + w.element(1, b"model")
    (tmp_path / "huge.mat").write_bytes(w.file({})[:128] + w.element(14, huge))
    with pytest.raises(PbdError):
        MatlabIOModel().deserialize(str(tmp_path / "huge.mat"))
    (tmp_path / "trunc.mat").write_bytes(raw[: len(raw) // 2])
    with pytest.raises(PbdError):
        MatlabIOModel().deserialize(str(tmp_path / "trunc.mat"))


# ------------------------------------------------------------------ an independent MAT-v5 writer (what MATLAB itself emits)
class MatWriter:
    """Minimal Level-5 writer: either byte order; doubles holding small integers are stored in the smallest integer type
    (miUINT8 / miINT16 / ...), elements of <= 4 bytes use the small data element format -- both as MATLAB does."""
    MI = {"i1": 1, "u1": 2, "i2": 3, "u2": 4, "i4": 5, "u4": 6, "f8": 9}

    def __init__(self, big_endian):
        self.e = ">" if big_endian else "<"

    def element(self, mi, payload):
        n = len(payload)
        if 0 < n <= 4:
            if self.e == "<":
                return struct.pack("<HH", mi, n) + payload.ljust(4, b"\0")
            return struct.pack(">HH", n, mi) + payload.ljust(4, b"\0")
        return struct.pack(self.e + "II", mi, n) + payload + b"\0" * (-n % 8)

    def numeric_payload(self, a):
        a = np.asarray(a, np.float64).ravel(order="F")
        if a.size and np.all(a == np.round(a)):
            for code, lo, hi in (("u1", 0, 255), ("i2", -32768, 32767), ("u2", 0, 65535), ("i4", -2 ** 31, 2 ** 31 - 1)):
                if a.min() >= lo and a.max() <= hi:
                    return self.MI[code], a.astype(self.e + code).tobytes()
        return self.MI["f8"], a.astype(self.e + "f8").tobytes()

    def matrix(self, name, value):
        if value is None:
            return self.element(14, b"")
        if isinstance(value, str):
            cls, dims = 4, (1, len(value))
            body = self.element(4, np.frombuffer(value.encode(), np.uint8).astype(self.e + "u2").tobytes())
        elif isinstance(value, list):                 # cell row vector
            cls, dims = 1, (1, len(value))
            body = b"".join(self.matrix("", v) for v in value)
        elif isinstance(value, tuple):                # struct array: (field names, [dict per element])
            fields, elems = value
            cls, dims = 2, (1, len(elems))
            flen = 32
            body = self.element(5, struct.pack(self.e + "i", flen))
            body += self.element(1, b"".join(f.encode().ljust(flen, b"\0") for f in fields))
            body += b"".join(self.matrix("", el[f]) for el in elems for f in fields)
        else:
            a = np.asarray(value, np.float64)
            if a.ndim < 2:
                a = a.reshape(1, -1)
            cls, dims = 6, a.shape
            mi, payload = self.numeric_payload(a)
            body = self.element(mi, payload) if a.size else b""
        head = self.element(6, struct.pack(self.e + "II", cls, 0))
        head += self.element(5, struct.pack(self.e + "%di" % len(dims), *dims))
        head += self.element(1, name.encode())
        return self.element(14, head + body)

    def file(self, variables, compress_names=()):
        hdr = b"MATLAB 5.0 MAT-file, written by tests/test_matlab_model.py".ljust(116) + b"\0" * 8
        hdr += struct.pack(self.e + "H", 0x0100) + (b"MI" if self.e == ">" else b"IM")
        out = hdr
        for name, v in variables.items():
            el = self.matrix(name, v)
            if name in compress_names:
                z = zlib.compress(el)
                el = struct.pack(self.e + "II", 15, len(z)) + z
            out += el
        return out


def to_writer_value(fm):
    d = matlab_model_dict(fm)

    def conv(sa, fields):
        return (fields, [{f: (None if (isinstance(sa[0, i][f], np.ndarray) and sa[0, i][f].size == 0) else sa[0, i][f]) for f in fields}
                         for i in range(sa.shape[1])])
    comps = [conv(d["components"][0, c], ["defid", "filterid", "parent", "biasid"]) for c in range(d["components"].shape[1])]
    model = {"interval": d["interval"], "thresh": d["thresh"], "sbin": d["sbin"], "filters": conv(d["filters"], ["w", "i"]),
             "components": comps, "defs": conv(d["defs"], ["w", "i", "anchor"]), "bias": conv(d["bias"], ["w", "i"])}
    return (list(model.keys()), [model])


@pytest.mark.parametrize("big_endian", [False, True])
@pytest.mark.parametrize("compressed", [False, True])
def test_matlab_style_storage_small_elements_and_byte_order(tmp_path, big_endian, compressed):
    fm = load_flat("Person_8parts")
    w = MatWriter(big_endian)
    raw = w.file({"name": "Person_8parts", "model": to_writer_value(fm)}, compress_names=("model",) if compressed else ())
    path = tmp_path / "m.mat"
    path.write_bytes(raw)
    if not big_endian and not compressed:             # the independent writer agrees with scipy's reader
        back = scipy.io.loadmat(str(path), squeeze_me=False)
        assert back["model"]["sbin"][0, 0][0, 0] == fm.sbin and str(back["name"][0]) == "Person_8parts"
    m = MatlabIOModel()
    assert m.deserialize(str(path))
    assert_same_model(m.to_flat(), fm, "Person_8parts")


def test_mat_model_equals_xml_model_through_the_reference_formats(tmp_path):
    # XML (FileStorageModel) and MAT (MatlabIOModel) of the same model give the same detector input
    fm = load_flat("Willowcoffee_5parts")
    scipy.io.savemat(str(tmp_path / "Willowcoffee_5parts.mat"), {"model": matlab_model_dict(fm)})
    a = MatlabIOModel()
    assert a.deserialize(str(tmp_path / "Willowcoffee_5parts.mat"))
    src = FileStorageModel.load_bin(os.path.join(os.path.dirname(__file__), "golden", "Willowcoffee_5parts.pbdm"))
    assert src.serialize(str(tmp_path / "w.xml"))
    x = FileStorageModel()
    assert x.deserialize(str(tmp_path / "w.xml"))
    assert_same_model(a.to_flat(), x.to_flat(), "Willowcoffee_5parts")
