"""ctypes binding of libimagestitch_b200.so (include/imagestitch.h).

The library is the product; this file only declares its C ABI for Python callers (tests, bench).  It
never falls back to a CPU implementation: if the shared library is missing it raises, and every compute
call fails with IS_ERR_CUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libimagestitch_b200.so")

IS_OK = 0
IS_ERR_NO_MEM, IS_ERR_BAD_ARG, IS_ERR_UNSUPPORTED, IS_ERR_ASSERT, IS_ERR_CUDA, IS_ERR_INTERNAL = -4, -5, -213, -215, -1000, -1001
IS_8U, IS_16S, IS_32S, IS_32F = 0, 3, 4, 5
PROJ_CYLINDRICAL, PROJ_SPHERICAL, PROJ_PLANE, PROJ_FISHEYE, PROJ_STEREOGRAPHIC = 0, 1, 2, 3, 4
INTER_NEAREST, INTER_LINEAR = 0, 1
BORDER_CONSTANT, BORDER_REFLECT = 0, 2
COST_COLOR, COST_COLOR_GRAD = 0, 1
WEIGHT_32F, WEIGHT_16S = 5, 3
FEED_COPY, FEED_BORROW, FEED_DEFER_WEIGHTS = 0, 1, 2
SEAM_NONE, SEAM_DP = 0, 1
EXPOSURE_NONE, EXPOSURE_GAIN = 0, 1
BLEND_MULTI_BAND, BLEND_FEATHER = 0, 1


class KeyPoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float), ("response", C.c_float), ("octave", C.c_int),
                ("class_id", C.c_int)]


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("grid_width", C.c_int), ("grid_height", C.c_int)]


class Mat(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("channels", C.c_int), ("depth", C.c_int),
                ("step", C.c_size_t), ("device", C.c_int)]


class Point(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int)]


class Size(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int)]


class Rect(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("width", C.c_int), ("height", C.c_int)]


class Camera(C.Structure):
    _fields_ = [("K", C.c_float * 9), ("R", C.c_float * 9)]


DETECT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(Mat))
MATCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int)
ESTIMATE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(Camera), C.POINTER(C.c_float))


class RegistrationHooks(C.Structure):
    _fields_ = [("user", C.c_void_p), ("detect", DETECT_FN), ("match", MATCH_FN), ("estimate", ESTIMATE_FN)]


class PipelineConfig(C.Structure):
    _fields_ = [("projection", C.c_int), ("seam", C.c_int), ("seam_cost", C.c_int), ("num_bands", C.c_int),
                ("weight_type", C.c_int), ("scale", C.c_float), ("exposure", C.c_int), ("blender", C.c_int), ("sharpness", C.c_float),
                ("seam_dilate", C.c_int)]


# every symbol include/imagestitch.h declares: name -> (restype, argtypes)
_P = C.POINTER
_F9 = _P(C.c_float)
SYMBOLS = {
    "is_version": (C.c_char_p, []),
    "is_status_string": (C.c_char_p, [C.c_int]),
    "is_ctx_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "is_ctx_destroy": (C.c_int, [C.c_void_p]),
    "is_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "is_ctx_last_error": (C.c_char_p, [C.c_void_p]),
    "is_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "is_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "is_ctx_reset_stream": (C.c_int, [C.c_void_p]),
    "is_ctx_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "is_ctx_device": (C.c_int, [C.c_void_p]),
    "is_ctx_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "is_ctx_kernel_timing_report": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "is_warp_roi": (C.c_int, [C.c_void_p, C.c_int, Size, _F9, _F9, C.c_float, _P(Point), _P(Size)]),
    "is_build_maps": (C.c_int, [C.c_void_p, C.c_int, Size, _F9, _F9, C.c_float, _P(Mat), _P(Mat), _P(Rect)]),
    "is_warp": (C.c_int, [C.c_void_p, C.c_int, _P(Mat), _F9, _F9, C.c_float, C.c_int, C.c_int, _P(Mat), _P(Point)]),
    "is_bmp_info": (C.c_int, [C.c_void_p, C.c_char_p, _P(Size), _P(C.c_int)]),
    "is_imread_bmp": (C.c_int, [C.c_void_p, C.c_char_p, _P(Mat)]),
    "is_imwrite_bmp": (C.c_int, [C.c_void_p, C.c_char_p, _P(Mat)]),
    "is_orb_find": (C.c_int, [C.c_void_p, _P(Mat), _P(OrbParams), _P(KeyPoint), C.c_void_p, C.c_int, _P(C.c_int)]),
    "is_remap": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), _P(Mat), C.c_int, C.c_int, _P(Mat)]),
    "is_warp_with_mask": (C.c_int, [C.c_void_p, C.c_int, _P(Mat), _F9, _F9, C.c_float, _P(Mat), _P(Mat), _P(Point)]),
    "is_seam_dp_find": (C.c_int, [C.c_void_p, C.c_int, _P(Mat), _P(Point), _P(Mat), C.c_int]),
    "is_seam_dp_find_trace": (C.c_int, [C.c_void_p, C.c_int, _P(Mat), _P(Point), _P(Mat), C.c_int, _P(C.c_int32), C.c_size_t, _P(C.c_size_t)]),
    "is_ctx_seam_speculation": (C.c_int, [C.c_void_p]),
    "is_ctx_seam_path": (C.c_int, [C.c_void_p]),
    "is_ctx_seam_waves": (C.c_int, [C.c_void_p]),
    "is_ctx_clear_plan_cache": (C.c_int, [C.c_void_p]),
    "is_debug_seam_pair_finish": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                            _P(C.c_int32), C.c_size_t]),
    "is_debug_dp_bench": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, _P(C.c_int32), _P(C.c_float)]),
    "is_debug_seam_pair_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                          _P(C.c_int32), C.c_size_t, _P(C.c_size_t)]),
    "is_seam_pair_run": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), Point, Point, _P(Mat), _P(Mat), _P(Mat), _P(Mat), _P(C.c_void_p)]),
    "is_seam_pair_check": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), Point, Point, _P(Mat), _P(Mat), C.c_void_p, _P(C.c_int)]),
    "is_seam_pair_destroy": (C.c_int, [C.c_void_p]),
    "is_seam_pair_same_structure": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), _P(Mat), Point, Point, _P(C.c_int)]),
    "is_mask_and": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat)]),
    "is_blender_strip_needs": (C.c_int, [C.c_void_p, Size, Point, C.c_int, C.c_int, _P(C.c_int)]),
    "is_blender_blend_strip": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(Mat), _P(Mat)]),
    "is_seam_cost_maps": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), Point, Point, _P(Mat), Point, C.c_int, Rect, _P(Mat), _P(Mat)]),
    "is_blender_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(C.c_void_p)]),
    "is_blender_destroy": (C.c_int, [C.c_void_p]),
    "is_blender_prepare": (C.c_int, [C.c_void_p, C.c_int, _P(Point), _P(Size)]),
    "is_blender_prepare_roi": (C.c_int, [C.c_void_p, Rect]),
    "is_blender_num_bands": (C.c_int, [C.c_void_p]),
    "is_blender_dst_size": (C.c_int, [C.c_void_p, _P(Size)]),
    "is_blender_feed": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), Point, C.c_int]),
    "is_blender_feed_ex": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), Point, C.c_int, C.c_longlong]),
    "is_blender_blend": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat)]),
    "is_linear_blend_size": (C.c_int, [Size, Size, Point, Point, _P(Size)]),
    "is_linear_blend_pair": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), Point, Point, _P(Mat), _P(C.c_int)]),
    "is_mask_dilate_and": (C.c_int, [C.c_void_p, _P(Mat), C.c_int, C.c_int, _P(Mat)]),
    "is_feather_weight_map": (C.c_int, [C.c_void_p, _P(Mat), C.c_float, _P(Mat)]),
    "is_feather_create": (C.c_int, [C.c_void_p, C.c_float, _P(C.c_void_p)]),
    "is_feather_destroy": (C.c_int, [C.c_void_p]),
    "is_feather_prepare": (C.c_int, [C.c_void_p, C.c_int, _P(Point), _P(Size)]),
    "is_feather_prepare_roi": (C.c_int, [C.c_void_p, Rect]),
    "is_feather_dst_size": (C.c_int, [C.c_void_p, _P(Size)]),
    "is_feather_feed": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), Point]),
    "is_feather_blend": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat)]),
    "is_gain_feed": (C.c_int, [C.c_void_p, C.c_int, _P(Point), _P(Mat), _P(Mat), _P(C.c_double)]),
    "is_gain_apply": (C.c_int, [C.c_void_p, _P(Mat), C.c_double]),
    "is_debug_seam_wave_schedule": (C.c_int, [C.c_int, _P(Point), _P(Size), _P(C.c_int32), C.c_size_t, _P(C.c_size_t)]),
    "is_pipeline_estimate": (C.c_int, [C.c_void_p, C.c_int, _P(Mat), _P(RegistrationHooks), _P(Camera), _P(C.c_float)]),
    "is_pipeline_plan": (C.c_int, [C.c_void_p, C.c_int, _P(Size), _P(Camera), _P(PipelineConfig), _P(Point), _P(Size), _P(Rect)]),
    "is_pipeline_run": (C.c_int, [C.c_void_p, C.c_int, _P(Mat), _P(Camera), _P(RegistrationHooks), _P(PipelineConfig), _P(Mat), _P(Mat), _P(Mat)]),
    "is_pipeline_last_gains": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_double)]),
    "is_pipeline_last_timings": (C.c_int, [C.c_void_p, _P(C.c_float)]),
}

_lib = None


def load():
    """Loads the shared library and binds every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m imagestitch_b200.build` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"imagestitch_b200 error {status}: {message}")
        self.status = status
