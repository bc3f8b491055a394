"""Ingest in front of detect(): what the reference's callers do before handing a cv::Mat to the detector -- cv::imread
(src/demo.cpp:88-99) and cv_bridge::toCvCopy of sensor_msgs/Image (ros/Node.cpp:165-176) -- over the C-ABI (csrc/ingest.cpp)."""
import ctypes as C

import numpy as np

from . import _lib


def imdecode(buf):
    """cv::imdecode(buf, IMREAD_COLOR): PNG / binary PNM bytes -> (h, w, 3) uint8 BGR."""
    b = np.frombuffer(bytes(buf), np.uint8)
    h, w, c, bits = (C.c_int() for _ in range(4))
    _lib.check(_lib.lib().pbd_image_info(b, b.size, C.byref(h), C.byref(w), C.byref(c), C.byref(bits)))
    out = np.empty((h.value, w.value, 3), np.uint8)
    _lib.check(_lib.lib().pbd_image_decode_bgr8(b, b.size, out.reshape(-1), out.size, None, None))
    return out


def imread(path):
    """cv::imread(path): (h, w, 3) uint8 BGR."""
    with open(path, "rb") as f:
        return imdecode(f.read())


def imdecode_depth(buf, scale=1.0 / 1000.0):
    """cv::imdecode(buf, IMREAD_ANYDEPTH) * scale (the demo converts millimetres to metres): (h, w) float32."""
    b = np.frombuffer(bytes(buf), np.uint8)
    h, w = C.c_int(), C.c_int()
    _lib.check(_lib.lib().pbd_image_info(b, b.size, C.byref(h), C.byref(w), None, None))
    out = np.empty((h.value, w.value), np.float32)
    _lib.check(_lib.lib().pbd_image_decode_depth_f32(b, b.size, float(scale), out.reshape(-1), out.size, None, None))
    return out


def from_ros_image(encoding, height, width, data, step=0, is_bigendian=0):
    """sensor_msgs/Image fields -> (h, w, 3) uint8 BGR, as cv_bridge::toCvCopy(msg, "bgr8")."""
    src = np.frombuffer(bytes(data), np.uint8)
    out = np.empty((height, width, 3), np.uint8)
    _lib.check(_lib.lib().pbd_ros_image_to_bgr8(encoding.encode(), height, width, step, int(is_bigendian), src, out.reshape(-1)))
    return out


def depth_from_ros_image(encoding, height, width, data, step=0, is_bigendian=0):
    """sensor_msgs/Image depth fields (32FC1 / 16UC1) -> (h, w) float32, as cv_bridge::toCvCopy(msg, "32FC1")."""
    src = np.frombuffer(bytes(data), np.uint8)
    out = np.empty((height, width), np.float32)
    _lib.check(_lib.lib().pbd_ros_depth_to_f32(encoding.encode(), height, width, step, int(is_bigendian), src, out.reshape(-1)))
    return out


class PinnedFrames:
    """A pinned (page-locked) host buffer of n frames for PartsBasedDetector.submit(): `.array` is an (n, h, w, c) uint8 view."""

    def __init__(self, n, h, w, c=3):
        self._p = C.c_void_p()
        size = n * h * w * c
        _lib.check(_lib.lib().pbd_host_alloc_pinned(size, C.byref(self._p)))
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint8)), (size,)).reshape(n, h, w, c)

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            _lib.lib().pbd_host_free_pinned(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
