"""ctypes binding of oracle/_ref/libpbd_ref.so -- the reference's OWN sources (include/DistanceTransform.hpp, include/Math.hpp,
src/HOGFeatures.cpp, src/DynamicProgram.cpp, include/Candidate.hpp) compiled unmodified against the minimal cv:: stand-in of
oracle/ref_shim.  TEST INFRASTRUCTURE: pins the restated oracle; never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_ref", "libpbd_ref.so")
_lib = None

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")


def available():
    """The library is built where /root/reference exists (make -C oracle ref) and travels as a prebuilt file otherwise."""
    if not os.path.exists(_SO) and os.path.isdir("/root/reference/src"):
        subprocess.call(["make", "-C", os.path.join(_ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        assert available(), "oracle/_ref/libpbd_ref.so missing"
        L = C.CDLL(_SO)
        L.ref_dt2d_f32.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, _f32p, _i32p, _i32p]
        L.ref_dt2d_f64.argtypes = [_f64p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, _f64p, _i32p, _i32p]
        L.ref_reduce_max_pick_f32.argtypes = [_f32p, _i32p, C.c_int, C.c_int, C.c_int, _f32p, _i32p, _i32p]
        L.ref_hog_pyramid.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_hog_level_dims.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.ref_hog_level.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.ref_dp_create.argtypes = [_i32p, _i32p, _f64p, _f32p, _i32p, _f32p, _i32p, C.c_int]
        L.ref_dp_create.restype = C.c_void_p
        L.ref_dp_destroy.argtypes = [C.c_void_p]
        L.ref_dp_set_levels.argtypes = [C.c_void_p, C.c_int, _i32p, _f32p]
        L.ref_dp_set_response.argtypes = [C.c_void_p, C.c_int, C.c_int, _f64p]
        L.ref_dp_run.argtypes = [C.c_void_p, C.c_double]
        L.ref_dp_get_root.argtypes = [C.c_void_p, C.c_int, C.c_int, _f64p, _i32p]
        L.ref_dp_get_backptr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p]
        L.ref_dp_candidate_nparts.argtypes = [C.c_void_p, C.c_int]
        L.ref_dp_get_candidate.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), _i32p, _f32p]
        L.ref_dp_sort_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, _i32p, _f32p]
        L.ref_rootmap_nms.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p]
        L.ref_dp_filter_by_depth.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_float, _i32p]
        L.ref_dp_filter_by_depth.restype = C.c_int
        _lib = L
    return _lib


def hog_pyramid(img, sbin, interval, flen=32, norient=18, precision=32):
    """HOGFeatures<T>::pyramid of the reference: list of (features (oh, ow, flen), scale)."""
    L = lib()
    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim == 2:
        img = img[:, :, None]
    h, w, c = img.shape
    n = L.ref_hog_pyramid(img.reshape(-1), h, w, c, sbin, interval, flen, norient, precision)
    out = []
    dt = np.float64 if precision == 64 else np.float32
    for l in range(n):
        r, cc, s = C.c_int(), C.c_int(), C.c_float()
        L.ref_hog_level_dims(l, precision, C.byref(r), C.byref(cc), C.byref(s))
        a = np.empty((r.value, cc.value), dt)
        L.ref_hog_level(l, precision, a.ctypes.data)
        out.append((a.reshape(r.value, cc.value // flen, flen), np.float32(s.value)))
    return out


class RefDP:
    """DynamicProgram<T>::min / argmin of the reference over caller-supplied response maps."""

    def __init__(self, flat_model, precision=32):
        self.L = lib()
        a = flat_model.to_arrays()
        self.model = flat_model
        self.precision = precision
        self.h = self.L.ref_dp_create(a["hdr"], a["fdims"], a["filters"], a["biasw"], a["anchors"], a["defs"], a["indexers"], precision)
        self.ohow = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_dp_destroy(self.h)
            self.h = None

    def set_levels(self, ohow, scales):
        self.ohow = np.ascontiguousarray(ohow, np.int32).reshape(-1, 2)
        self.L.ref_dp_set_levels(self.h, len(self.ohow), self.ohow.reshape(-1), np.ascontiguousarray(scales, np.float32))

    def set_response(self, level, f, arr):
        self.L.ref_dp_set_response(self.h, level, f, np.ascontiguousarray(arr, np.float64).ravel())

    def run(self, thresh):
        self.ncand = self.L.ref_dp_run(self.h, float(thresh))
        return self.ncand

    def root(self, level, comp=0):
        oh, ow = self.ohow[level]
        v, i = np.empty(oh * ow, np.float64), np.empty(oh * ow, np.int32)
        self.L.ref_dp_get_root(self.h, level, comp, v, i)
        return v.reshape(oh, ow), i.reshape(oh, ow)

    def backptr(self, level, comp, part, pm):
        oh, ow = self.ohow[level]
        ix, iy, ik = (np.empty(oh * ow, np.int32) for _ in range(3))
        self.L.ref_dp_get_backptr(self.h, level, comp, part, pm, ix, iy, ik)
        return ix.reshape(oh, ow), iy.reshape(oh, ow), ik.reshape(oh, ow)

    def candidates(self):
        out = []
        for i in range(self.ncand):
            n = self.L.ref_dp_candidate_nparts(self.h, i)
            comp = C.c_int()
            rects, conf = np.empty(n * 4, np.int32), np.empty(n, np.float32)
            self.L.ref_dp_get_candidate(self.h, i, C.byref(comp), rects, conf)
            out.append((comp.value, rects.reshape(n, 4), conf))
        return out

    def filter_by_depth(self, depth, zfactor):
        """SearchSpacePruning<float>::filterCandidatesByDepth on the last run's candidates: 0/1 per candidate.  Math::median prints every
        box it looks at to std::cerr (include/Math.hpp:70), so fd 2 is parked on /dev/null for the call."""
        import os
        depth = np.ascontiguousarray(depth, np.float32)
        keep = np.zeros(max(self.ncand, 1), np.int32)
        saved, null = os.dup(2), os.open(os.devnull, os.O_WRONLY)
        os.dup2(null, 2)
        try:
            self.L.ref_dp_filter_by_depth(self.h, depth.reshape(-1), depth.shape[0], depth.shape[1], float(zfactor), keep)
        finally:
            os.dup2(saved, 2)
            os.close(saved)
            os.close(null)
        return keep[:self.ncand]

    def sort_nms(self, im_h, im_w, overlap):
        n = max(self.L.ref_dp_candidate_nparts(self.h, 0), 1) if self.ncand else 1
        rects, scores = np.zeros(max(self.ncand, 1) * n * 4, np.int32), np.zeros(max(self.ncand, 1), np.float32)
        k = self.L.ref_dp_sort_nms(self.h, im_h, im_w, float(overlap), self.ncand, rects, scores)
        return rects.reshape(-1, n, 4)[:k], scores[:k]
