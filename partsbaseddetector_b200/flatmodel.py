"""Flat (SoA) view of a parts model -- the layout the C-ABI exchanges.

Mirrors the fields of the reference `Model` (include/Model.hpp:49-122) as they are
(de)serialised by FileStorageModel (src/FileStorageModel.cpp:42-159):
name, interval (stored in Model::nscales_), thresh, sbin, norient, flen, filtersw,
biasw, anchors, defs and the per-component/part indexers.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class FlatPart:
    parentid: int
    filterid: List[int]
    biasid: List[int]
    defid: List[int]


@dataclass
class FlatModel:
    name: str = ""
    interval: int = 0
    thresh: float = 0.0
    sbin: int = 0
    norient: int = 18
    flen: int = 32
    filters: List[np.ndarray] = field(default_factory=list)   # each (kh, kw*flen) float64, HWC
    biasw: np.ndarray = None                                  # float32
    anchors: np.ndarray = None                                # (ndefs, 2) int32 (x, y)
    defs: np.ndarray = None                                   # (ndefs, 4) float32
    comps: List[List[FlatPart]] = field(default_factory=list)

    def to_arrays(self):
        nf = len(self.filters)
        fdims = np.zeros((nf, 2), np.int32)
        for i, f in enumerate(self.filters):
            fdims[i] = (f.shape[0], f.shape[1] // self.flen)
        filt = np.concatenate([np.ascontiguousarray(f, np.float64).ravel() for f in self.filters]) if nf else np.zeros(0)
        idx = []
        for comp in self.comps:
            idx.append(len(comp))
            for p in comp:
                idx += [p.parentid, len(p.filterid), len(p.biasid), len(p.defid)]
                idx += list(p.filterid) + list(p.biasid) + list(p.defid)
        hdr = np.array([self.interval, self.sbin, self.norient, self.flen, nf, len(self.biasw),
                        len(self.defs), len(self.comps)], np.int32)
        return dict(hdr=hdr, fdims=np.ascontiguousarray(fdims.ravel()),
                    filters=np.ascontiguousarray(filt, np.float64),
                    biasw=np.ascontiguousarray(self.biasw, np.float32),
                    anchors=np.ascontiguousarray(np.asarray(self.anchors, np.int32).ravel()),
                    defs=np.ascontiguousarray(np.asarray(self.defs, np.float32).ravel()),
                    indexers=np.array(idx, np.int32))

    def nfilters(self):
        return len(self.filters)

    def nmix(self, c, p):
        return len(self.comps[c][p].filterid)
