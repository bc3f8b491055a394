"""ctypes binding of the C-ABI (include/pbd_b200.h) exported by the in-tree libpbd_b200.so.

The library is the product: there is no Python or CPU fallback.  If the shared library is missing it is
built from csrc/ with nvcc (sm_100a); if that fails the import raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpbd_b200.so")
_lib = None

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")

# every symbol include/pbd_b200.h declares (checked by tests/test_abi_symbols.py)
EXPORTS = [
    "pbd_last_error", "pbd_version", "pbd_model_load_xml", "pbd_model_save_xml", "pbd_model_load_storage", "pbd_model_save_storage", "pbd_model_load_mat", "pbd_model_load_bin",
    "pbd_model_save_bin", "pbd_model_create", "pbd_model_free", "pbd_model_name", "pbd_model_header",
    "pbd_model_filter", "pbd_model_bias", "pbd_model_anchors", "pbd_model_defs", "pbd_model_nparts",
    "pbd_model_part", "pbd_create", "pbd_destroy", "pbd_set_option", "pbd_get_option", "pbd_detect_batch_u8",
    "pbd_detect_batch_u8_device", "pbd_enqueue_batch_u8_device", "pbd_collect_candidates", "pbd_submit_batch_u8", "pbd_collect_ticket",
    "pbd_candidates_count", "pbd_candidates_nparts", "pbd_candidates_get", "pbd_candidates_export", "pbd_candidates_free",
    "pbd_candidates_sort", "pbd_candidates_nms", "pbd_candidates_filter_by_depth", "pbd_candidates_create", "pbd_stage_pyramid", "pbd_stage_pdf", "pbd_stage_dp_min", "pbd_stage_dp_argmin",
    "pbd_pyramid_geometry", "pbd_num_frames", "pbd_num_levels", "pbd_level_info", "pbd_get_pyramid_image", "pbd_get_features",
    "pbd_get_response", "pbd_get_rootv", "pbd_get_rooti", "pbd_get_backptr", "pbd_set_levels",
    "pbd_set_features", "pbd_set_response", "pbd_dt2d_f32_device", "pbd_dt2d_f32", "pbd_dt2d_plan_create", "pbd_dt2d_plan_impl", "pbd_dt2d_plan_set_segment", "pbd_dt2d_plan_replayed",
    "pbd_dt2d_plan_run", "pbd_dt2d_plan_destroy", "pbd_image_info", "pbd_image_decode_bgr8", "pbd_image_decode_depth_f32", "pbd_imread_bgr8",
    "pbd_ros_image_to_bgr8", "pbd_ros_depth_to_f32", "pbd_host_alloc_pinned", "pbd_host_free_pinned", "pbd_launch_count",
    "pbd_stage_times_ms", "pbd_kernel_times_ms", "pbd_device_bytes",
]


class PbdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pbd_b200 error %d: %s" % (code, msg))
        self.code = code


def build(force=False):
    """Compile csrc/ into libpbd_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", csrc, "-j8"], stdout=subprocess.DEVNULL)
    return SO_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PBD_B200_LIB") or SO_PATH      # developer knob: A/B a differently compiled build of the same library
    if path == SO_PATH and not os.path.exists(SO_PATH):
        build()
    L = C.CDLL(path)
    vp, ci, cf, cd = C.c_void_p, C.c_int, C.c_float, C.c_double
    P = C.POINTER
    L.pbd_last_error.restype = C.c_char_p
    L.pbd_version.restype = C.c_char_p
    L.pbd_model_load_xml.argtypes = [C.c_char_p, P(vp)]
    L.pbd_model_load_mat.argtypes = [C.c_char_p, P(vp)]
    L.pbd_model_load_storage.argtypes = [C.c_char_p, P(vp)]
    L.pbd_model_save_storage.argtypes = [vp, C.c_char_p]
    L.pbd_model_save_xml.argtypes = [vp, C.c_char_p]
    L.pbd_model_load_bin.argtypes = [C.c_char_p, P(vp)]
    L.pbd_model_save_bin.argtypes = [vp, C.c_char_p]
    L.pbd_model_create.argtypes = [C.c_char_p, _i32p, cf, _i32p, _f64p, _f32p, _i32p, _f32p, _i32p, P(vp)]
    L.pbd_model_free.argtypes = [vp]
    L.pbd_model_name.argtypes = [vp]
    L.pbd_model_name.restype = C.c_char_p
    L.pbd_model_header.argtypes = [vp, _i32p, P(cf)]
    L.pbd_model_filter.argtypes = [vp, ci, P(ci), P(ci), P(P(cd))]
    L.pbd_model_bias.argtypes = [vp, P(P(cf)), P(ci)]
    L.pbd_model_anchors.argtypes = [vp, P(P(ci)), P(ci)]
    L.pbd_model_defs.argtypes = [vp, P(P(cf)), P(ci)]
    L.pbd_model_nparts.argtypes = [vp, ci]
    L.pbd_model_part.argtypes = [vp, ci, ci, P(ci), ci, _i32p, ci, P(ci)]
    L.pbd_create.argtypes = [vp, ci, vp, P(vp)]
    L.pbd_destroy.argtypes = [vp]
    L.pbd_set_option.argtypes = [vp, C.c_char_p, cd]
    L.pbd_get_option.argtypes = [vp, C.c_char_p, P(cd)]
    L.pbd_detect_batch_u8.argtypes = [vp, vp, ci, ci, ci, ci, C.c_size_t, C.c_size_t, P(vp)]
    L.pbd_detect_batch_u8_device.argtypes = [vp, vp, ci, ci, ci, ci, P(vp)]
    L.pbd_enqueue_batch_u8_device.argtypes = [vp, vp, ci, ci, ci, ci]
    L.pbd_collect_candidates.argtypes = [vp, P(vp)]
    L.pbd_submit_batch_u8.argtypes = [vp, vp, ci, ci, ci, ci, P(ci)]
    L.pbd_collect_ticket.argtypes = [vp, ci, P(vp)]
    L.pbd_candidates_count.argtypes = [vp]
    L.pbd_candidates_nparts.argtypes = [vp, ci]
    L.pbd_candidates_get.argtypes = [vp, ci, P(ci), P(ci), P(ci), P(cf), _i32p, _i32p, _i32p, _i32p]
    L.pbd_candidates_export.argtypes = [vp, _i32p, _f32p, _i32p, ci]
    L.pbd_candidates_free.argtypes = [vp]
    L.pbd_candidates_sort.argtypes = [vp]
    L.pbd_candidates_nms.argtypes = [vp, ci, ci, cf]
    L.pbd_candidates_filter_by_depth.argtypes = [vp, vp, _f32p, ci, ci, C.c_size_t, cf]
    L.pbd_candidates_create.argtypes = [ci, ci, _i32p, _f32p, _i32p, P(vp)]
    L.pbd_stage_pyramid.argtypes = [vp, vp, ci, ci, ci, ci, C.c_size_t, C.c_size_t]
    L.pbd_stage_pdf.argtypes = [vp]
    L.pbd_stage_dp_min.argtypes = [vp]
    L.pbd_stage_dp_argmin.argtypes = [vp, P(vp)]
    L.pbd_pyramid_geometry.argtypes = [ci, ci, ci, ci, ci, ci, _i32p, _f32p]
    L.pbd_num_frames.argtypes = [vp]
    L.pbd_num_levels.argtypes = [vp]
    L.pbd_level_info.argtypes = [vp, ci, P(ci), P(ci), P(ci), P(ci), P(cf)]
    L.pbd_get_pyramid_image.argtypes = [vp, ci, ci, _u8p]
    L.pbd_get_features.argtypes = [vp, ci, ci, _f32p]
    L.pbd_get_response.argtypes = [vp, ci, ci, ci, _f32p]
    L.pbd_get_rootv.argtypes = [vp, ci, ci, ci, _f32p]
    L.pbd_get_rooti.argtypes = [vp, ci, ci, ci, _i32p]
    L.pbd_get_backptr.argtypes = [vp, ci, ci, ci, ci, ci, _i32p, _i32p, _i32p]
    L.pbd_set_levels.argtypes = [vp, ci, ci, _i32p, _f32p]
    L.pbd_set_features.argtypes = [vp, ci, ci, _f32p]
    L.pbd_set_response.argtypes = [vp, ci, ci, ci, _f32p]
    L.pbd_dt2d_f32_device.argtypes = [vp, vp, ci, ci, ci, _f32p, _i32p, vp, vp, vp, ci]
    L.pbd_dt2d_f32.argtypes = [_f32p, ci, ci, ci, _f32p, _i32p, _f32p, _i32p, _i32p, ci]
    L.pbd_dt2d_plan_create.argtypes = [ci, ci, ci, _f32p, _i32p, ci, P(vp)]
    L.pbd_dt2d_plan_impl.argtypes = [vp]
    L.pbd_dt2d_plan_set_segment.argtypes = [vp, ci]
    L.pbd_dt2d_plan_replayed.argtypes = [vp]
    L.pbd_dt2d_plan_replayed.restype = C.c_longlong
    L.pbd_dt2d_plan_run.argtypes = [vp, vp, vp, vp, vp, vp, ci]
    L.pbd_dt2d_plan_destroy.argtypes = [vp]
    L.pbd_dt2d_plan_destroy.restype = None
    L.pbd_image_info.argtypes = [_u8p, C.c_size_t, P(ci), P(ci), P(ci), P(ci)]
    L.pbd_image_decode_bgr8.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, P(ci), P(ci)]
    L.pbd_image_decode_depth_f32.argtypes = [_u8p, C.c_size_t, cf, _f32p, C.c_size_t, P(ci), P(ci)]
    L.pbd_imread_bgr8.argtypes = [C.c_char_p, vp, C.c_size_t, P(ci), P(ci)]
    L.pbd_ros_image_to_bgr8.argtypes = [C.c_char_p, ci, ci, C.c_size_t, ci, _u8p, _u8p]
    L.pbd_ros_depth_to_f32.argtypes = [C.c_char_p, ci, ci, C.c_size_t, ci, _u8p, _f32p]
    L.pbd_host_alloc_pinned.argtypes = [C.c_size_t, P(vp)]
    L.pbd_host_free_pinned.argtypes = [vp]
    L.pbd_host_free_pinned.restype = None
    L.pbd_launch_count.argtypes = [vp]
    L.pbd_launch_count.restype = C.c_longlong
    L.pbd_stage_times_ms.argtypes = [vp, _f32p]
    L.pbd_kernel_times_ms.argtypes = [vp, _f32p]
    L.pbd_device_bytes.argtypes = [vp]
    L.pbd_device_bytes.restype = C.c_size_t
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise PbdError(rc, lib().pbd_last_error().decode("utf-8", "replace"))
