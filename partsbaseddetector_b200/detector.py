"""Host-side mirror of the reference's interface for the detect() path, over the C-ABI.

Names follow the reference: `FileStorageModel.deserialize/serialize` (src/FileStorageModel.cpp),
`PartsBasedDetector.distributeModel/detect` (src/PartsBasedDetector.cpp:69-127), `Candidate`
(include/Candidate.hpp) and the stage interfaces `pyramid` (IFeatures), `pdf` (IConvolutionEngine),
`min` / `argmin` (DynamicProgram).  Everything numeric happens in libpbd_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from .flatmodel import FlatModel, FlatPart


class Model:
    """reference `Model` (include/Model.hpp:49-122): owns a pbd_model handle."""

    def __init__(self, handle=None):
        self._h = handle

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().pbd_model_free(self._h)
                self._h = None
        except Exception:          # interpreter shutdown: module globals may already be gone
            pass

    @property
    def handle(self):
        if not self._h:
            raise _lib.PbdError(-5, "model is empty; call deserialize() first")
        return self._h

    # --- getters (include/Model.hpp:98-118) ---
    def _hdr(self):
        hdr = np.zeros(8, np.int32)
        th = C.c_float()
        _lib.check(_lib.lib().pbd_model_header(self.handle, hdr, C.byref(th)))
        return hdr, th.value

    def name(self):
        return _lib.lib().pbd_model_name(self.handle).decode()

    def nscales(self):
        return int(self._hdr()[0][0])          # the reference keeps `interval` in nscales_

    def binsize(self):
        return int(self._hdr()[0][1])

    def norient(self):
        return int(self._hdr()[0][2])

    def flen(self):
        return int(self._hdr()[0][3])

    def thresh(self):
        return self._hdr()[1]

    def ncomponents(self):
        return int(self._hdr()[0][7])

    def to_flat(self):
        L = _lib.lib()
        hdr, th = self._hdr()
        m = FlatModel(name=self.name(), interval=int(hdr[0]), thresh=float(np.float32(th)), sbin=int(hdr[1]),
                      norient=int(hdr[2]), flen=int(hdr[3]))
        for i in range(hdr[4]):
            r, k = C.c_int(), C.c_int()
            p = C.POINTER(C.c_double)()
            _lib.check(L.pbd_model_filter(self.handle, i, C.byref(r), C.byref(k), C.byref(p)))
            m.filters.append(np.ctypeslib.as_array(p, (r.value, k.value * m.flen)).copy())
        n = C.c_int()
        pf = C.POINTER(C.c_float)()
        _lib.check(L.pbd_model_bias(self.handle, C.byref(pf), C.byref(n)))
        m.biasw = np.ctypeslib.as_array(pf, (n.value,)).copy()
        pi = C.POINTER(C.c_int)()
        _lib.check(L.pbd_model_anchors(self.handle, C.byref(pi), C.byref(n)))
        m.anchors = np.ctypeslib.as_array(pi, (n.value, 2)).copy() if n.value else np.zeros((0, 2), np.int32)
        _lib.check(L.pbd_model_defs(self.handle, C.byref(pf), C.byref(n)))
        m.defs = np.ctypeslib.as_array(pf, (n.value, 4)).copy() if n.value else np.zeros((0, 4), np.float32)
        for c in range(hdr[7]):
            parts = []
            for p in range(L.pbd_model_nparts(self.handle, c)):
                lists = []
                par = C.c_int()
                for which in range(3):
                    buf = np.zeros(256, np.int32)
                    cnt = C.c_int()
                    _lib.check(L.pbd_model_part(self.handle, c, p, C.byref(par), which, buf, 256, C.byref(cnt)))
                    lists.append([int(v) for v in buf[:cnt.value]])
                parts.append(FlatPart(par.value, *lists))
            m.comps.append(parts)
        return m

    @classmethod
    def from_flat(cls, fm):
        a = fm.to_arrays()
        h = C.c_void_p()
        _lib.check(_lib.lib().pbd_model_create(fm.name.encode(), a["hdr"], float(fm.thresh), a["fdims"], a["filters"], a["biasw"],
                                               a["anchors"] if len(a["anchors"]) else np.zeros(2, np.int32),
                                               a["defs"] if len(a["defs"]) else np.zeros(4, np.float32), a["indexers"], C.byref(h)))
        return cls(h)

    def save_bin(self, path):
        _lib.check(_lib.lib().pbd_model_save_bin(self.handle, path.encode()))

    @classmethod
    def load_bin(cls, path):
        h = C.c_void_p()
        _lib.check(_lib.lib().pbd_model_load_bin(path.encode(), C.byref(h)))
        return cls(h)


class FileStorageModel(Model):
    """reference FileStorageModel (src/FileStorageModel.cpp): XML / YAML (de)serialisation as cv::FileStorage (format from the
    content on reading, from the extension on writing); bool returns."""

    def deserialize(self, filename):
        h = C.c_void_p()
        rc = _lib.lib().pbd_model_load_storage(filename.encode(), C.byref(h))
        if rc == -2:            # cannot open: the reference returns false (src/FileStorageModel.cpp:100-101)
            return False
        _lib.check(rc)
        if self._h:
            _lib.lib().pbd_model_free(self._h)
        self._h = h
        return True

    def serialize(self, filename):
        _lib.check(_lib.lib().pbd_model_save_storage(self.handle, filename.encode()))
        return True


class MatlabIOModel(Model):
    """reference MatlabIOModel (src/MatlabIOModel.cpp:71-188): reads the Matlab training code's `model` struct from a version-5
    .mat file (native reader inside libpbd_b200.so; the reference needs cvmatio).  serialize() is a stub in the reference too."""

    def deserialize(self, filename):
        h = C.c_void_p()
        rc = _lib.lib().pbd_model_load_mat(filename.encode(), C.byref(h))
        if rc == -2:            # cannot open: `if (!ok) return false;` (src/MatlabIOModel.cpp:76-77)
            return False
        _lib.check(rc)
        if self._h:
            _lib.lib().pbd_model_free(self._h)
        self._h = h
        return True

    def serialize(self, filename):
        return False            # src/MatlabIOModel.cpp:191-195: "TODO: implement", returns false


class Candidate:
    """reference Candidate (include/Candidate.hpp:56-99) + frame/level/part locations."""

    __slots__ = ("frame", "level", "component_", "confidence_", "parts_", "x", "y", "m")

    def parts(self):
        return self.parts_             # (nparts, 4) int32: cv::Rect x, y, width, height

    def confidence(self):
        return self.confidence_

    def score(self):
        return self.confidence_[0]

    def component(self):
        return self.component_

    @staticmethod
    def sort(cands):
        """In place on a list; a CandidateList is returned as a new sorted list."""
        if isinstance(cands, list):
            cands.sort(key=lambda c: -float(c.score()))      # stable, descending (Candidate.hpp:97-99)
            return cands
        return sorted(cands, key=lambda c: -float(c.score()))

    def boundingBox(self):
        """cv::Rect hull of the part boxes (include/Candidate.hpp:104-110) as (x, y, width, height)."""
        p = self.parts_
        x0, y0 = int(p[:, 0].min()), int(p[:, 1].min())
        return (x0, y0, int((p[:, 0] + p[:, 2]).max()) - x0, int((p[:, 1] + p[:, 3]).max()) - y0)

    @staticmethod
    def nonMaximaSuppression(im, candidates, overlap=0.0):
        """Candidate::nonMaximaSuppression (include/Candidate.hpp:277-304).  `im` is an image (or its (h, w) shape);
        `candidates` a list of Candidate or a CandidateList in the order to be processed (sort first).  A list is
        filtered in place (and returned), a CandidateList yields a new CandidateList."""
        shape = im if isinstance(im, tuple) else np.asarray(im).shape
        h, w = int(shape[0]), int(shape[1])
        n = len(candidates)
        if isinstance(candidates, CandidateList):
            meta, scores, parts = candidates.meta, candidates.scores, candidates.parts
        else:
            mp = max([len(c.x) for c in candidates] + [1])
            meta = np.zeros((n, 4), np.int32)
            scores = np.zeros(n, np.float32)
            parts = np.zeros((n, mp, 7), np.int32)
            for i, c in enumerate(candidates):
                k = len(c.x)
                meta[i] = (c.frame, c.level, c.component_, k)
                scores[i] = c.score()
                parts[i, :k, 0], parts[i, :k, 1], parts[i, :k, 2] = c.x, c.y, c.m
                parts[i, :k, 3:7] = c.parts_
        hnd = C.c_void_p()
        L = _lib.lib()
        mp = parts.shape[1]
        _lib.check(L.pbd_candidates_create(n, mp, np.ascontiguousarray(meta).reshape(-1), np.ascontiguousarray(scores),
                                           np.ascontiguousarray(parts).reshape(-1), C.byref(hnd)))
        _lib.check(L.pbd_candidates_nms(hnd, h, w, float(overlap)))
        out = _unpack_candidates(hnd)
        if isinstance(candidates, CandidateList):
            return out
        candidates[:] = list(out)
        return candidates


def filterCandidatesByDepth(model, candidates, depth, zfactor=0.03):
    """SearchSpacePruning<T>::filterCandidatesByDepth (src/SearchSpacePruning.cpp:73-95): `candidates` (a CandidateList or a list of
    Candidate, all of ONE frame) filtered by the frame's depth image (h x w float32; 0 = no reading).  Returns a CandidateList."""
    depth = np.ascontiguousarray(depth, np.float32)
    if isinstance(candidates, CandidateList):
        meta, scores, parts = candidates.meta, candidates.scores, candidates.parts
    else:
        n = len(candidates)
        mp = max([len(c.x) for c in candidates] + [1])
        meta, scores, parts = np.zeros((n, 4), np.int32), np.zeros(n, np.float32), np.zeros((n, mp, 7), np.int32)
        for i, c in enumerate(candidates):
            k = len(c.x)
            meta[i] = (c.frame, c.level, c.component_, k)
            scores[i] = c.score()
            parts[i, :k, 0], parts[i, :k, 1], parts[i, :k, 2] = c.x, c.y, c.m
            parts[i, :k, 3:7] = c.parts_
    hnd = C.c_void_p()
    L = _lib.lib()
    _lib.check(L.pbd_candidates_create(len(scores), parts.shape[1], np.ascontiguousarray(meta).reshape(-1), np.ascontiguousarray(scores),
                                       np.ascontiguousarray(parts).reshape(-1), C.byref(hnd)))
    try:
        _lib.check(L.pbd_candidates_filter_by_depth(hnd, model.handle, depth.reshape(-1), depth.shape[0], depth.shape[1], 0, float(zfactor)))
    except Exception:
        L.pbd_candidates_free(hnd)
        raise
    return _unpack_candidates(hnd)


class CandidateList:
    """Read-only sequence of Candidate backed by the arrays of one bulk export; Candidate objects are built on access.
    `meta` = (n,4) frame/level/component/nparts, `scores` = (n,), `parts` = (n, max_parts, 7) x,y,mixture,rect."""

    def __init__(self, meta, scores, parts):
        self.meta, self.scores, self.parts = meta, scores, parts

    def __len__(self):
        return len(self.scores)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        c = Candidate()
        npart = int(self.meta[i, 3])
        c.frame, c.level, c.component_ = int(self.meta[i, 0]), int(self.meta[i, 1]), int(self.meta[i, 2])
        conf = np.zeros(npart, np.float32)
        conf[0] = self.scores[i]                                 # root = rootv, others 0.0 (DynamicProgram.cpp:241-244)
        c.confidence_ = conf
        p = self.parts[i, :npart]
        c.parts_ = p[:, 3:7]
        c.x, c.y, c.m = p[:, 0], p[:, 1], p[:, 2]
        return c

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _unpack_candidates(handle, free=True):
    """pbd_candidates -> CandidateList (one bulk export call)."""
    L = _lib.lib()
    n = L.pbd_candidates_count(handle)
    meta = np.empty((n, 4), np.int32)
    scores = np.empty(n, np.float32)
    parts = np.empty((n, 1, 7), np.int32)
    if n:
        mp = max(L.pbd_candidates_nparts(handle, 0), 1)
        while True:
            parts = np.empty((n, mp, 7), np.int32)
            rc = L.pbd_candidates_export(handle, meta.reshape(-1), scores, parts.reshape(-1), mp)
            if rc == 0:
                break
            mp *= 2                                              # components with more parts than the first candidate
            if mp > 4096:
                _lib.check(rc)
    if free:
        L.pbd_candidates_free(handle)
    return CandidateList(meta, scores, parts)


class PartsBasedDetector:
    """reference PartsBasedDetector<float> (include/PartsBasedDetector.hpp:152-175)."""

    def __init__(self, device=0, stream=0):
        self._d = None
        self._device = device
        self._stream = stream
        self._name = ""
        self._model = None

    def __del__(self):
        try:
            self.close()
        except Exception:          # interpreter shutdown: module globals may already be gone
            pass

    def close(self):
        if getattr(self, "_d", None):
            _lib.lib().pbd_destroy(self._d)
            self._d = None

    @property
    def handle(self):
        if not self._d:
            raise _lib.PbdError(-5, "distributeModel() has not been called")
        return self._d

    def distributeModel(self, model):
        self.close()
        d = C.c_void_p()
        _lib.check(_lib.lib().pbd_create(model.handle, self._device, C.c_void_p(self._stream), C.byref(d)))
        self._d = d
        self._name = model.name()
        self._model = model

    def name(self):
        return self._name

    def set_option(self, key, value):
        _lib.check(_lib.lib().pbd_set_option(self.handle, key.encode(), float(value)))

    def get_option(self, key):
        v = C.c_double()
        _lib.check(_lib.lib().pbd_get_option(self.handle, key.encode(), C.byref(v)))
        return v.value

    # ---- whole path ----
    @staticmethod
    def _frames(im):
        a = np.asarray(im)
        if a.dtype != np.uint8:
            raise _lib.PbdError(-6, "Unsupported image type (only 8-bit frames)")     # reference CV_Error
        if a.ndim == 2:
            a = a[None, :, :, None]
        elif a.ndim == 3:
            a = a[None] if a.shape[2] in (1, 3) else a[..., None]
        return np.ascontiguousarray(a)

    def detect(self, im, depth=None, candidates=None):
        """detect(im[, depth], candidates): appends to `candidates` like the reference (never clears it).
        `im` is one HxWx3 (BGR) / HxW frame or an NxHxWxC batch; `depth` is accepted and ignored (T7)."""
        a = self._frames(im)
        n, h, w, c = a.shape
        out = C.c_void_p()
        _lib.check(_lib.lib().pbd_detect_batch_u8(self.handle, a.ctypes.data, n, h, w, c, 0, 0, C.byref(out)))
        res = _unpack_candidates(out)
        if candidates is None:
            return res                                   # CandidateList (lazy)
        candidates.extend(res)                           # reference semantics: append to the caller's vector
        return candidates

    # ---- pipelined (streaming) API: submit batch i+1 before collecting batch i ----
    def submit(self, im):
        """Enqueue H2D + all stages for a batch of frames without waiting; returns a ticket.  `im` must stay untouched (and
        should be pinned memory) until collect_ticket(ticket)."""
        a = self._frames(im)
        n, h, w, c = a.shape
        t = C.c_int()
        _lib.check(_lib.lib().pbd_submit_batch_u8(self.handle, a.ctypes.data, n, h, w, c, C.byref(t)))
        return t.value, a            # keep a reference to the frames alive with the ticket

    def collect_ticket(self, ticket):
        t = ticket[0] if isinstance(ticket, tuple) else ticket
        out = C.c_void_p()
        _lib.check(_lib.lib().pbd_collect_ticket(self.handle, t, C.byref(out)))
        return _unpack_candidates(out)

    def detect_stream(self, batches):
        """Generator over an iterable of frame batches: yields each batch's CandidateList, keeping two batches in flight."""
        prev = None
        for b in batches:
            cur = self.submit(b)
            if prev is not None:
                yield self.collect_ticket(prev)
            prev = cur
        if prev is not None:
            yield self.collect_ticket(prev)

    def detect_device(self, dptr, n, h, w, c):
        out = C.c_void_p()
        _lib.check(_lib.lib().pbd_detect_batch_u8_device(self.handle, C.c_void_p(dptr), n, h, w, c, C.byref(out)))
        return _unpack_candidates(out)

    def enqueue_device(self, dptr, n, h, w, c):
        _lib.check(_lib.lib().pbd_enqueue_batch_u8_device(self.handle, C.c_void_p(dptr), n, h, w, c))

    def collect(self):
        out = C.c_void_p()
        _lib.check(_lib.lib().pbd_collect_candidates(self.handle, C.byref(out)))
        return _unpack_candidates(out)

    # ---- stage level (IFeatures / IConvolutionEngine / DynamicProgram) ----
    def pyramid(self, im):
        a = self._frames(im)
        n, h, w, c = a.shape
        _lib.check(_lib.lib().pbd_stage_pyramid(self.handle, a.ctypes.data, n, h, w, c, 0, 0))

    def pdf(self):
        _lib.check(_lib.lib().pbd_stage_pdf(self.handle))

    def min(self):
        _lib.check(_lib.lib().pbd_stage_dp_min(self.handle))

    def argmin(self):
        out = C.c_void_p()
        _lib.check(_lib.lib().pbd_stage_dp_argmin(self.handle, C.byref(out)))
        return _unpack_candidates(out)

    def nscales(self):
        return _lib.lib().pbd_num_levels(self.handle)

    def level_info(self, l):
        v = [C.c_int() for _ in range(4)]
        s = C.c_float()
        _lib.check(_lib.lib().pbd_level_info(self.handle, l, *[C.byref(x) for x in v], C.byref(s)))
        return dict(img_h=v[0].value, img_w=v[1].value, oh=v[2].value, ow=v[3].value, scale=np.float32(s.value))

    def scales(self):
        return [self.level_info(l)["scale"] for l in range(self.nscales())]

    def pyramid_image(self, frame, l, channels=3):
        li = self.level_info(l)
        out = np.empty(li["img_h"] * li["img_w"] * channels, np.uint8)
        _lib.check(_lib.lib().pbd_get_pyramid_image(self.handle, frame, l, out))
        return out.reshape(li["img_h"], li["img_w"], channels)

    def features(self, frame, l):
        li = self.level_info(l)
        out = np.empty(li["oh"] * li["ow"] * 32, np.float32)
        _lib.check(_lib.lib().pbd_get_features(self.handle, frame, l, out))
        return out.reshape(li["oh"], li["ow"], 32)

    def response(self, frame, l, f):
        li = self.level_info(l)
        out = np.empty(li["oh"] * li["ow"], np.float32)
        _lib.check(_lib.lib().pbd_get_response(self.handle, frame, l, f, out))
        return out.reshape(li["oh"], li["ow"])

    def rootv(self, frame, l, c=0):
        li = self.level_info(l)
        out = np.empty(li["oh"] * li["ow"], np.float32)
        _lib.check(_lib.lib().pbd_get_rootv(self.handle, frame, l, c, out))
        return out.reshape(li["oh"], li["ow"])

    def rooti(self, frame, l, c=0):
        li = self.level_info(l)
        out = np.empty(li["oh"] * li["ow"], np.int32)
        _lib.check(_lib.lib().pbd_get_rooti(self.handle, frame, l, c, out))
        return out.reshape(li["oh"], li["ow"])

    def backptr(self, frame, l, c, p, m):
        li = self.level_info(l)
        n = li["oh"] * li["ow"]
        ix, iy, ik = (np.empty(n, np.int32) for _ in range(3))
        _lib.check(_lib.lib().pbd_get_backptr(self.handle, frame, l, c, p, m, ix, iy, ik))
        s = (li["oh"], li["ow"])
        return ix.reshape(s), iy.reshape(s), ik.reshape(s)

    def set_levels(self, n_frames, ohow, scales):
        ohow = np.ascontiguousarray(ohow, np.int32).reshape(-1)
        scales = np.ascontiguousarray(scales, np.float32)
        _lib.check(_lib.lib().pbd_set_levels(self.handle, n_frames, len(scales), ohow, scales))

    def set_features(self, frame, l, arr):
        _lib.check(_lib.lib().pbd_set_features(self.handle, frame, l, np.ascontiguousarray(arr, np.float32).reshape(-1)))

    def set_response(self, frame, l, f, arr):
        _lib.check(_lib.lib().pbd_set_response(self.handle, frame, l, f, np.ascontiguousarray(arr, np.float32).reshape(-1)))

    def launch_count(self):
        return int(_lib.lib().pbd_launch_count(self.handle))

    def stage_times_ms(self):
        ms = np.zeros(6, np.float32)
        _lib.check(_lib.lib().pbd_stage_times_ms(self.handle, ms))
        return dict(zip(["h2d", "pyramid", "hog", "pdf", "dp_min", "argmin"], [float(x) for x in ms]))

    def kernel_times_ms(self):
        """Per-kernel device time of the last batch run with set_option("timing", 2)."""
        ms = np.zeros(6, np.float32)
        _lib.check(_lib.lib().pbd_kernel_times_ms(self.handle, ms))
        return dict(zip(["feat_split", "part_response", "dt_rows", "dt_cols", "mix_max", "root_select"], [float(x) for x in ms]))

    def device_bytes(self):
        return int(_lib.lib().pbd_device_bytes(self.handle))


class Dt2dPlan:
    """Standalone generalised 2-D distance transform (DistanceTransform<float>::compute) of n maps of h x w with all tables and
    scratch buffers pre-allocated: run() only enqueues kernels (device pointers, e.g. torch tensors' data_ptr()).
    impl: 0 default (1), 1 streaming envelope (any line length <= 4096), 2 parallel-in-q (lines <= 1024), 3 streaming with lagged-scan
    emission, 4 windowed certified evaluation with replay (the detector's default transform).  Identical results."""

    def __init__(self, n, h, w, defw4, anchor_xy, impl=0):
        self.n, self.h, self.w = n, h, w
        defw4 = np.ascontiguousarray(np.broadcast_to(np.asarray(defw4, np.float32).reshape(-1, 4), (n, 4)))
        anchor = np.ascontiguousarray(np.broadcast_to(np.asarray(anchor_xy, np.int32).reshape(-1, 2), (n, 2)))
        self._p = C.c_void_p()
        _lib.check(_lib.lib().pbd_dt2d_plan_create(n, h, w, defw4.reshape(-1), anchor.reshape(-1), impl, C.byref(self._p)))

    def impl(self):
        return _lib.lib().pbd_dt2d_plan_impl(self._p)

    def set_segment(self, steps):
        """impl 4: lines cut into segments of `steps` walk steps (multiple of 16; 0 = off), as the detector does for small launches"""
        _lib.check(_lib.lib().pbd_dt2d_plan_set_segment(self._p, int(steps)))

    def replayed(self):
        """impl 4: lines handed to the stack algorithm since the last call (synchronises)"""
        return int(_lib.lib().pbd_dt2d_plan_replayed(self._p))

    def run(self, d_in, d_out, d_ix, d_iy, backptr_mode=0, stream=0):
        _lib.check(_lib.lib().pbd_dt2d_plan_run(self._p, C.c_void_p(stream), C.c_void_p(d_in), C.c_void_p(d_out), C.c_void_p(d_ix), C.c_void_p(d_iy),
                                                backptr_mode))

    def close(self):
        if getattr(self, "_p", None):
            _lib.lib().pbd_dt2d_plan_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dt2d(score, defw4, anchor_xy, backptr_mode=0):
    """Generalised 2-D distance transform of one or more maps on the GPU (DistanceTransform<float>::compute)."""
    a = np.ascontiguousarray(score, np.float32)
    if a.ndim == 2:
        a = a[None]
    n, h, w = a.shape
    defw4 = np.ascontiguousarray(np.broadcast_to(np.asarray(defw4, np.float32).reshape(-1, 4), (n, 4)))
    anchor = np.ascontiguousarray(np.broadcast_to(np.asarray(anchor_xy, np.int32).reshape(-1, 2), (n, 2)))
    out = np.empty(n * h * w, np.float32)
    ix = np.empty(n * h * w, np.int32)
    iy = np.empty(n * h * w, np.int32)
    _lib.check(_lib.lib().pbd_dt2d_f32(a.reshape(-1), n, h, w, defw4.reshape(-1), anchor.reshape(-1), out, ix, iy, backptr_mode))
    return out.reshape(n, h, w), ix.reshape(n, h, w), iy.reshape(n, h, w)
