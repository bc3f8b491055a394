"""ctypes binding of the CPU oracle (oracle/libpbd_oracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libpbd_oracle.so")
_lib = None

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")


def build():
    src = os.path.join(_ROOT, "oracle", "pbd_oracle.cpp")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    L.orc_resize_u8.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
    L.orc_pyrdown_u8.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
    L.orc_pyramid_geometry.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _f32p]
    L.orc_pyramid_geometry.restype = C.c_int
    L.orc_hog_dims.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_hog_f32.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
    L.orc_hog_f64.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p]
    L.orc_convolve_f32.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, _f32p]
    L.orc_convolve_f64.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, _f64p]
    L.orc_dt1d_f32.argtypes = [_f32p, C.c_int, C.c_double, C.c_double, C.c_int, _f32p, _i32p]
    L.orc_dt1d_f64.argtypes = [_f64p, C.c_int, C.c_double, C.c_double, C.c_int, _f64p, _i32p]
    L.orc_dt2d_f32.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _i32p, _i32p]
    L.orc_dt2d_f64.argtypes = [_f64p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _f64p, _i32p, _i32p]
    L.orc_nms.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _i32p]
    L.orc_nms.restype = C.c_int
    L.orc_rootmap_nms.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p]
    L.orc_filter_by_depth.argtypes = [_i32p, C.c_int, C.c_int, _i32p, _i32p, _f32p, C.c_int, C.c_int, C.c_float, _i32p]
    L.orc_filter_by_depth.restype = C.c_int
    L.orc_create.argtypes = [_i32p, C.c_float, _i32p, _f64p, _f32p, _i32p, _f32p, _i32p, C.c_int]
    L.orc_create.restype = C.c_void_p
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_thresh.argtypes = [C.c_void_p, C.c_double]
    L.orc_set_backptr_mode.argtypes = [C.c_void_p, C.c_int]
    L.orc_set_max_levels.argtypes = [C.c_void_p, C.c_int]
    L.orc_num_threads.restype = C.c_int
    L.orc_set_num_threads.argtypes = [C.c_int]
    L.orc_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_set_levels.argtypes = [C.c_void_p, C.c_int, _i32p, _f32p]
    L.orc_set_features.argtypes = [C.c_void_p, C.c_int, _f64p]
    L.orc_set_response.argtypes = [C.c_void_p, C.c_int, C.c_int, _f64p]
    L.orc_nlevels.argtypes = [C.c_void_p]
    L.orc_level_info.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_float)]
    L.orc_get_image.argtypes = [C.c_void_p, C.c_int, _u8p]
    L.orc_get_features.argtypes = [C.c_void_p, C.c_int, _f64p]
    L.orc_get_response.argtypes = [C.c_void_p, C.c_int, C.c_int, _f64p]
    L.orc_get_rootv.argtypes = [C.c_void_p, C.c_int, C.c_int, _f64p]
    L.orc_get_rooti.argtypes = [C.c_void_p, C.c_int, C.c_int, _i32p]
    L.orc_get_backptr.argtypes = [C.c_void_p] + [C.c_int] * 4 + [_i32p] * 3
    L.orc_ncandidates.argtypes = [C.c_void_p]
    L.orc_candidate_nparts.argtypes = [C.c_void_p, C.c_int]
    L.orc_get_candidate.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float),
                                    _i32p, _i32p, _i32p, _i32p]
    L.orc_get_timings.argtypes = [C.c_void_p, _f64p]
    _lib = L
    return L


def load_flat_npz(path):
    """FlatModel from a committed *_flat.npz fixture (tests/golden/make_golden.py): no product library involved."""
    from partsbaseddetector_b200.flatmodel import FlatModel
    with np.load(path) as z:
        return FlatModel.from_arrays({k: z[k] for k in ("hdr", "fdims", "filters", "biasw", "anchors", "defs", "indexers")},
                                     name=str(z["name"]), thresh=float(z["thresh"]))


def physical_cores():
    """Number of physical cores this process may run on (SMT siblings counted once)."""
    try:
        allowed = os.sched_getaffinity(0)
        cores = set()
        cpu = phys = core = None
        for line in open("/proc/cpuinfo"):
            if line.startswith("processor"):
                cpu = int(line.split(":")[1])
            elif line.startswith("physical id"):
                phys = int(line.split(":")[1])
            elif line.startswith("core id"):
                core = int(line.split(":")[1])
                if cpu in allowed:
                    cores.add((phys, core))
        return max(len(cores), 1) if cores else max(len(allowed), 1)
    except (OSError, ValueError, AttributeError):
        return os.cpu_count() or 1


def use_all_cores():
    """The reference parallelises with OpenMP over all host threads; launchers such as torchrun export OMP_NUM_THREADS=1,
    so the thread count is set explicitly (one thread per physical core)."""
    n = physical_cores()
    lib().orc_set_num_threads(n)
    return n


class OracleDetector:
    """Restated PartsBasedDetector<T> (reference src/PartsBasedDetector.cpp:69-127) on the CPU."""

    def __init__(self, flat_model, precision=32):
        self.L = lib()
        self.model = flat_model
        self.precision = precision
        a = flat_model.to_arrays()
        self.h = self.L.orc_create(a["hdr"], float(flat_model.thresh), a["fdims"], a["filters"], a["biasw"],
                                   a["anchors"], a["defs"], a["indexers"], precision)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_thresh(self, t):
        self.L.orc_set_thresh(self.h, float(t))

    def set_backptr_mode(self, m):
        self.L.orc_set_backptr_mode(self.h, int(m))

    def set_max_levels(self, m):
        self.L.orc_set_max_levels(self.h, int(m))

    def run(self, img, first=1, last=4):
        if img is not None:
            img = np.ascontiguousarray(img, np.uint8)
            if img.ndim == 2:
                img = img[:, :, None]
            h, w, c = img.shape
            self._cn = c
            self.L.orc_run(self.h, img.ctypes.data, h, w, c, first, last)
        else:
            self.L.orc_run(self.h, None, 0, 0, 0, first, last)

    def set_levels(self, ohow, scales):
        ohow = np.ascontiguousarray(ohow, np.int32)
        scales = np.ascontiguousarray(scales, np.float32)
        self.L.orc_set_levels(self.h, len(scales), ohow, scales)

    def set_features(self, level, arr):
        self.L.orc_set_features(self.h, level, np.ascontiguousarray(arr, np.float64).ravel())

    def set_response(self, level, f, arr):
        self.L.orc_set_response(self.h, level, f, np.ascontiguousarray(arr, np.float64).ravel())

    def nlevels(self):
        return self.L.orc_nlevels(self.h)

    def level_info(self, l):
        v = [C.c_int() for _ in range(4)]
        s = C.c_float()
        self.L.orc_level_info(self.h, l, *[C.byref(x) for x in v], C.byref(s))
        return dict(img_h=v[0].value, img_w=v[1].value, oh=v[2].value, ow=v[3].value, scale=np.float32(s.value))

    def image(self, l):
        li = self.level_info(l)
        out = np.empty((li["img_h"], li["img_w"], getattr(self, "_cn", 3)), np.uint8)
        self.L.orc_get_image(self.h, l, out.reshape(-1))
        return out

    def _dt(self):
        return np.float32 if self.precision == 32 else np.float64

    def features(self, l):
        li = self.level_info(l)
        out = np.empty(li["oh"] * li["ow"] * self.model.flen, np.float64)
        self.L.orc_get_features(self.h, l, out)
        return out.reshape(li["oh"], li["ow"], self.model.flen).astype(self._dt())

    def response(self, l, f):
        li = self.level_info(l)
        out = np.empty(li["oh"] * li["ow"], np.float64)
        self.L.orc_get_response(self.h, l, f, out)
        return out.reshape(li["oh"], li["ow"]).astype(self._dt())

    def rootv(self, l, c=0):
        li = self.level_info(l)
        out = np.empty(li["oh"] * li["ow"], np.float64)
        self.L.orc_get_rootv(self.h, l, c, out)
        return out.reshape(li["oh"], li["ow"]).astype(self._dt())

    def rooti(self, l, c=0):
        li = self.level_info(l)
        out = np.empty(li["oh"] * li["ow"], np.int32)
        self.L.orc_get_rooti(self.h, l, c, out)
        return out.reshape(li["oh"], li["ow"])

    def backptr(self, l, c, p, m):
        li = self.level_info(l)
        n = li["oh"] * li["ow"]
        ix, iy, ik = (np.empty(n, np.int32) for _ in range(3))
        self.L.orc_get_backptr(self.h, l, c, p, m, ix, iy, ik)
        s = (li["oh"], li["ow"])
        return ix.reshape(s), iy.reshape(s), ik.reshape(s)

    def candidates(self):
        out = []
        for i in range(self.L.orc_ncandidates(self.h)):
            npart = self.L.orc_candidate_nparts(self.h, i)
            lv, cp, sc = C.c_int(), C.c_int(), C.c_float()
            xs, ys, ms = (np.empty(npart, np.int32) for _ in range(3))
            rc = np.empty(npart * 4, np.int32)
            self.L.orc_get_candidate(self.h, i, C.byref(lv), C.byref(cp), C.byref(sc), xs, ys, ms, rc)
            out.append(dict(level=lv.value, component=cp.value, score=np.float32(sc.value), x=xs, y=ys, m=ms,
                            rects=rc.reshape(npart, 4)))
        return out

    def timings(self):
        t = np.zeros(5)
        self.L.orc_get_timings(self.h, t)
        return dict(pyramid=t[0], hog=t[1], pdf=t[2], dp_min=t[3], argmin=t[4])
