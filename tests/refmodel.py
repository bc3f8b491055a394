"""Reference-format model reader built on cv2.FileStorage (the authority for the
opencv_storage XML/YAML written by reference src/FileStorageModel.cpp:42-94).
TEST INFRASTRUCTURE: used to feed the oracle and to check the product's own C++ loader."""
import os

import cv2
import numpy as np

from partsbaseddetector_b200.flatmodel import FlatModel, FlatPart

REF_MODELS = "/root/reference/models"


def _seq_ints(node):
    if node.isNone() or node.empty():
        return []
    if node.isInt() or node.isReal():
        return [int(node.real())]
    return [int(node.at(i).real()) for i in range(node.size())]


def load_xml_cv2(path):
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    assert fs.isOpened(), path
    m = FlatModel()
    m.name = fs.getNode("name").string()
    m.interval = int(fs.getNode("interval").real())
    m.thresh = float(np.float32(fs.getNode("thresh").real()))
    m.sbin = int(fs.getNode("sbin").real())
    m.norient = int(fs.getNode("norient").real())
    m.flen = int(fs.getNode("flen").real())
    fw = fs.getNode("filtersw")
    m.filters = [np.ascontiguousarray(fw.at(i).mat(), np.float64) for i in range(fw.size())]
    bw = fs.getNode("biasw")
    m.biasw = np.array([bw.at(i).real() for i in range(bw.size())], np.float64).astype(np.float32)
    an = fs.getNode("anchors")
    m.anchors = np.array([int(an.at(i).real()) for i in range(an.size())], np.int32).reshape(-1, 2)
    df = fs.getNode("defs")
    m.defs = np.array([[df.at(i).at(j).real() for j in range(4)] for i in range(df.size())], np.float64).astype(np.float32).reshape(-1, 4)
    comps = fs.getNode("indexers")
    for c in range(comps.size()):
        cn = comps.getNode("component-%d" % c)
        parts = []
        for p in range(cn.size()):
            pn = cn.getNode("part-%d" % p)
            defid = _seq_ints(pn.getNode("defid")) or [0]       # T2: full sequence; root/empty -> [0]
            parts.append(FlatPart(int(pn.getNode("parentid").real()), _seq_ints(pn.getNode("filterid")),
                                  _seq_ints(pn.getNode("biasid")), defid))
        m.comps.append(parts)
    fs.release()
    return m


def ref_model_path(name):
    return os.path.join(REF_MODELS, name)
