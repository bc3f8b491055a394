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

    @classmethod
    def from_arrays(cls, a, name="", thresh=0.0):
        """Inverse of to_arrays() (hdr / fdims / filters / biasw / anchors / defs / indexers)."""
        hdr = [int(v) for v in a["hdr"]]
        m = cls(name=str(name), interval=hdr[0], thresh=float(thresh), sbin=hdr[1], norient=hdr[2], flen=hdr[3])
        fd = np.asarray(a["fdims"], np.int32).reshape(-1, 2)
        flt = np.asarray(a["filters"], np.float64)
        off = 0
        for kh, kw in fd:
            n = int(kh) * int(kw) * m.flen
            m.filters.append(flt[off:off + n].reshape(int(kh), int(kw) * m.flen).copy())
            off += n
        m.biasw = np.asarray(a["biasw"], np.float32).copy()
        m.anchors = np.asarray(a["anchors"], np.int32).reshape(-1, 2).copy()
        m.defs = np.asarray(a["defs"], np.float32).reshape(-1, 4).copy()
        ix = [int(v) for v in a["indexers"]]
        i = 0
        for _ in range(hdr[7]):
            nparts = ix[i]; i += 1
            parts = []
            for _ in range(nparts):
                par, nf, nb, nd = ix[i:i + 4]; i += 4
                fid = ix[i:i + nf]; i += nf
                bid = ix[i:i + nb]; i += nb
                did = ix[i:i + nd]; i += nd
                parts.append(FlatPart(par, fid, bid, did))
            m.comps.append(parts)
        return m

    def nfilters(self):
        return len(self.filters)

    def nmix(self, c, p):
        return len(self.comps[c][p].filterid)
