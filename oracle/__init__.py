"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

ctypes/numpy front end of ``oracle/_build/liboracle.so`` (sources: oracle/*.cpp, header oracle/oracle.h).
The oracle is the parity checker for the CUDA path and the timed CPU baseline of bench.py.  It must
never be imported from ``imagestitch_b200`` (the product); only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it.

Parity pin: see oracle/oracle.h (pinned against OpenCV 4.13 fixtures in tests/golden/; the
hand-written linear blend against the reference's own block compiled into oracle/_ref, `build_ref()`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_build", "liboracle.so")

PROJ_CYLINDRICAL, PROJ_SPHERICAL, PROJ_PLANE, PROJ_FISHEYE, PROJ_STEREOGRAPHIC = 0, 1, 2, 3, 4
INTER_NEAREST, INTER_LINEAR = 0, 1
BORDER_CONSTANT, BORDER_REFLECT = 0, 2
COST_COLOR, COST_COLOR_GRAD = 0, 1
WEIGHT_32F, WEIGHT_16S = 5, 3


def build(force: bool = False) -> str:
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".cpp", ".h"))]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _DIR] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_get_max_threads.restype = C.c_int
        L.orc_dp_seam_find.restype = C.c_int
        L.orc_mb_create.restype = C.c_void_p
        L.orc_mb_num_bands.restype = C.c_int
        L.orc_lin_blend.restype = C.c_int
        L.orc_pipeline_plan.restype = C.c_int
        L.orc_pipeline_run.restype = C.c_int
        L.orc_pipeline_run_ex.restype = C.c_int
        L.orc_pipeline_run_ex2.restype = C.c_int
        L.orc_gain_feed.restype = C.c_int
        L.orc_fb_create.restype = C.c_void_p
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f9(m):
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(9))


def set_threads(n: int):
    lib().orc_set_threads(C.c_int(int(n)))


def max_threads() -> int:
    return int(lib().orc_get_max_threads())


# ---------------------------------------------------------------- warp
def camera_params(K, R):
    k_rinv = np.zeros(9, np.float32)
    r_kinv = np.zeros(9, np.float32)
    lib().orc_camera_params(_p(_f9(K)), _p(_f9(R)), _p(k_rinv), _p(r_kinv))
    return k_rinv.reshape(3, 3), r_kinv.reshape(3, 3)


def detect_roi(proj, src_size_wh, K, R, scale, full_scan=True):
    """-> (tl_x, tl_y, br_x, br_y)   [WARP]:64-88"""
    roi = np.zeros(4, np.int32)
    lib().orc_detect_roi(C.c_int(proj), C.c_int(src_size_wh[0]), C.c_int(src_size_wh[1]), _p(_f9(K)), _p(_f9(R)),
                         C.c_float(scale), C.c_int(1 if full_scan else 0), _p(roi))
    return tuple(int(v) for v in roi)


def build_maps(proj, src_size_wh, K, R, scale, full_scan=True):
    """-> (roi, xmap, ymap)   [WARP]:122-144.  roi = (tl_x, tl_y, br_x, br_y)."""
    roi = detect_roi(proj, src_size_wh, K, R, scale, full_scan)
    h, w = roi[3] - roi[1] + 1, roi[2] - roi[0] + 1
    xmap = np.empty((h, w), np.float32)
    ymap = np.empty((h, w), np.float32)
    r = np.asarray(roi, np.int32)
    lib().orc_build_maps(C.c_int(proj), _p(_f9(K)), _p(_f9(R)), C.c_float(scale), _p(r), _p(xmap), _p(ymap))
    return roi, xmap, ymap


def remap(src, xmap, ymap, interp, border):
    src = np.ascontiguousarray(src)
    assert src.dtype == np.uint8
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = xmap.shape
    dst = np.empty((h, w) if src.ndim == 2 else (h, w, ch), np.uint8)
    xmap = np.ascontiguousarray(xmap, np.float32)
    ymap = np.ascontiguousarray(ymap, np.float32)
    lib().orc_remap_u8(_p(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]), C.c_int(ch), C.c_size_t(src.strides[0]),
                       _p(xmap), _p(ymap), C.c_int(h), C.c_int(w), C.c_int(interp), C.c_int(border), _p(dst))
    return dst


def warp(proj, src, K, R, scale, interp, border, full_scan=True):
    """-> ((tl_x, tl_y), dst)   [WARP]:145-161"""
    roi, xmap, ymap = build_maps(proj, (src.shape[1], src.shape[0]), K, R, scale, full_scan)
    return (roi[0], roi[1]), remap(src, xmap, ymap, interp, border)


# ---------------------------------------------------------------- seam
def dp_seam_find(images, corners, masks, cost_fn=COST_COLOR, want_trace=False):
    """[SEAM]:87-124.  images: list of HxWx3 (or HxWx4: the fourth channel is ignored) float32 or uint8; masks: list of HxW uint8.
    Returns new masks (inputs are not modified) and, optionally, the seam trace
    [(i, j, comp, is_horizontal, points Nx2 in pano coords), ...]."""
    n = len(images)
    is_u8 = images[0].dtype == np.uint8
    imgs = [np.ascontiguousarray(im, np.uint8 if is_u8 else np.float32) for im in images]
    out = [np.ascontiguousarray(m, np.uint8).copy() for m in masks]
    ip = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
    mp = (C.c_void_p * n)(*[m.ctypes.data for m in out])
    rows = np.asarray([im.shape[0] for im in imgs], np.int32)
    cols = np.asarray([im.shape[1] for im in imgs], np.int32)
    cxy = np.asarray(corners, np.int32).reshape(-1).copy()
    cap = 0
    trace = None
    if want_trace:
        cap = int(sum(5 + 2 * (im.shape[0] + im.shape[1]) for im in imgs) * max(1, n) * 2)
        trace = np.zeros(cap, np.int32)
    tlen = C.c_size_t(0)
    flags = (1 if is_u8 else 0) | (2 if imgs[0].ndim == 3 and imgs[0].shape[2] == 4 else 0)      # bit 1: CV_8UC4 / CV_32FC4 ([SEAM]:745-748)
    rc = lib().orc_dp_seam_find(C.c_int(n), ip, C.c_int(flags), _p(rows), _p(cols), _p(cxy), mp,
                                C.c_int(cost_fn), _p(trace) if want_trace else None, C.c_size_t(cap), C.byref(tlen))
    if rc:
        raise RuntimeError(f"orc_dp_seam_find failed: {rc}")
    if not want_trace:
        return out
    return out, parse_trace(trace[: min(cap, tlen.value)])


def parse_trace(t):
    res = []
    k = 0
    while k + 5 <= len(t):
        i, j, comp, horiz, npts = (int(v) for v in t[k:k + 5])
        pts = np.asarray(t[k + 5:k + 5 + 2 * npts]).reshape(-1, 2).copy()
        res.append((i, j, comp, bool(horiz), pts))
        k += 5 + 2 * npts
    return res


def seam_gradients(img):
    """[SEAM]:549-572: Sobel x / y (CV_32F, ksize 3) of cvtColor(BGR2GRAY) of the image taken as CV_32F"""
    is_u8 = img.dtype == np.uint8
    a = np.ascontiguousarray(img, np.uint8 if is_u8 else np.float32)
    gx = np.empty(a.shape[:2], np.float32)
    gy = np.empty(a.shape[:2], np.float32)
    lib().orc_seam_gradients(_p(a), C.c_int(1 if is_u8 else 0), C.c_int(a.shape[0]), C.c_int(a.shape[1]), _p(gx), _p(gy))
    return gx, gy


def seam_costs(img1, img2, tl1, tl2, labels, union_tl, l, roi_xywh, cost_fn=COST_COLOR):
    is_u8 = img1.dtype == np.uint8
    a = np.ascontiguousarray(img1)
    b = np.ascontiguousarray(img2)
    labels = np.ascontiguousarray(labels, np.int32)
    x, y, w, h = roi_xywh
    costV = np.empty((h, w + 1), np.float32)
    costH = np.empty((h + 1, w), np.float32)
    roi = np.asarray(roi_xywh, np.int32)
    lib().orc_seam_costs_ex(_p(a), _p(b), C.c_int(1 if is_u8 else 0), C.c_int(a.shape[0]), C.c_int(a.shape[1]),
                         C.c_int(b.shape[0]), C.c_int(b.shape[1]), C.c_int(tl1[0]), C.c_int(tl1[1]),
                         C.c_int(tl2[0]), C.c_int(tl2[1]), _p(labels), C.c_int(labels.shape[0]), C.c_int(labels.shape[1]),
                         C.c_int(union_tl[0]), C.c_int(union_tl[1]), C.c_int(l), _p(roi), C.c_int(cost_fn), _p(costV), _p(costH))
    return costV, costH


# ---------------------------------------------------------------- pyramids / blend
def pyr_down_s16(src):
    src = np.ascontiguousarray(src, np.int16)
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    dst = np.empty(((h + 1) // 2, (w + 1) // 2) + (() if src.ndim == 2 else (ch,)), np.int16)
    lib().orc_pyr_down_s16(_p(src), C.c_int(h), C.c_int(w), C.c_int(ch), _p(dst))
    return dst


def pyr_up_s16(src, dsize_hw):
    src = np.ascontiguousarray(src, np.int16)
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    dst = np.empty(tuple(dsize_hw) + (() if src.ndim == 2 else (ch,)), np.int16)
    lib().orc_pyr_up_s16(_p(src), C.c_int(h), C.c_int(w), C.c_int(ch), C.c_int(dsize_hw[0]), C.c_int(dsize_hw[1]), _p(dst))
    return dst


def pyr_down_f32(src):
    src = np.ascontiguousarray(src, np.float32)
    h, w = src.shape
    dst = np.empty(((h + 1) // 2, (w + 1) // 2), np.float32)
    lib().orc_pyr_down_f32(_p(src), C.c_int(h), C.c_int(w), _p(dst))
    return dst


class MultiBandBlender:
    """Call shape of cv::detail::MultiBandBlender as used at [SEAM]:1244-1252,1271,1280."""

    def __init__(self, num_bands=5, weight_type=WEIGHT_32F):
        self._h = C.c_void_p(lib().orc_mb_create(C.c_int(num_bands), C.c_int(weight_type)))
        self._roi = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_mb_destroy(self._h)
            self._h = None

    def prepare(self, dst_roi_xywh):
        self._roi = tuple(int(v) for v in dst_roi_xywh)
        lib().orc_mb_prepare(self._h, _p(np.asarray(self._roi, np.int32)))

    def prepare_corners(self, corners, sizes_wh):
        self.prepare(result_roi(corners, sizes_wh))

    def num_bands(self):
        return int(lib().orc_mb_num_bands(self._h))

    def feed(self, img, mask, tl):
        img = np.ascontiguousarray(img, np.int16)
        mask = np.ascontiguousarray(mask, np.uint8)
        lib().orc_mb_feed(self._h, _p(img), _p(mask), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_int(tl[0]), C.c_int(tl[1]))

    def blend(self):
        x, y, w, h = self._roi
        dst = np.empty((h, w, 3), np.int16)
        dmask = np.empty((h, w), np.uint8)
        lib().orc_mb_blend(self._h, _p(dst), _p(dmask))
        return dst, dmask


def result_roi(corners, sizes_wh):
    """cv::detail::resultRoi(corners, sizes) -> (x, y, w, h)"""
    tlx = min(c[0] for c in corners)
    tly = min(c[1] for c in corners)
    brx = max(c[0] + s[0] for c, s in zip(corners, sizes_wh))
    bry = max(c[1] + s[1] for c, s in zip(corners, sizes_wh))
    return (tlx, tly, brx - tlx, bry - tly)


# ---------------------------------------------------------------- linear blend ([BLEND])
def lin_blend(img1, img2, tl1, tl2, want_cost=False):
    a = np.ascontiguousarray(img1, np.float32)
    b = np.ascontiguousarray(img2, np.float32)
    he = C.c_int(0)
    br = C.c_int(0)
    lib().orc_lin_geometry(C.c_int(a.shape[0]), C.c_int(a.shape[1]), C.c_int(b.shape[0]), C.c_int(b.shape[1]),
                           C.c_int(tl1[0]), C.c_int(tl1[1]), C.c_int(tl2[0]), C.c_int(tl2[1]), C.byref(he), C.byref(br))
    pano = np.zeros((he.value, br.value, 3), np.float32)
    seam = np.zeros(he.value, np.int32)
    ib = a.shape[1] - (tl2[0] - tl1[0])
    cost = np.zeros((he.value, ib + 2), np.float32) if want_cost else None
    rc = lib().orc_lin_blend(_p(a), C.c_int(a.shape[0]), C.c_int(a.shape[1]), _p(b), C.c_int(b.shape[0]), C.c_int(b.shape[1]),
                             C.c_int(tl1[0]), C.c_int(tl1[1]), C.c_int(tl2[0]), C.c_int(tl2[1]), _p(pano), _p(seam),
                             _p(cost) if want_cost else None)
    if rc == 1:
        return None
    return (pano, seam, cost) if want_cost else (pano, seam)


# ---------------------------------------------------------------- oracle/_ref: the reference's own code
_REF_SO = os.path.join(_DIR, "_ref", "libref_linblend.so")
_REF_WARP_SO = os.path.join(_DIR, "_ref", "libref_warp.so")
_REF_SEAM_SO = os.path.join(_DIR, "_ref", "libref_seam.so")
_ref = None
_ref_warp = None
_ref_seam = None


def build_ref() -> str | None:
    """Compiles the reference's own hand-written pair blend ([BLEND]:141-717) from /root/reference against
    oracle/ref_shim/cvshim.h (recipe: `make -C oracle ref`).  Returns the .so path, or None where neither the
    reference sources nor a prebuilt oracle/_ref exist."""
    if os.path.isdir("/root/reference"):
        srcs = [os.path.join(_DIR, "ref_shim", f) for f in os.listdir(os.path.join(_DIR, "ref_shim"))] + [os.path.join(_DIR, "Makefile")]
        outs = (_REF_SO, _REF_WARP_SO, _REF_SEAM_SO)
        if not all(os.path.exists(o) for o in outs) or any(os.path.getmtime(x) > min(os.path.getmtime(o) for o in outs) for x in srcs):
            subprocess.check_call(["make", "-s", "-C", _DIR, "ref"], stdout=subprocess.DEVNULL)
    return _REF_SO if all(os.path.exists(o) for o in (_REF_SO, _REF_WARP_SO, _REF_SEAM_SO)) else None


def ref_dp_seam_find(images, corners, masks, cost_fn=COST_COLOR):
    """The reference's own find() ([SEAM]:87-1093, compiled into oracle/_ref) -> new masks.  COLOR_GRAD on 8-bit images goes
    through an 8-bit gray image there (cv::cvtColor on CV_8U), unlike the CV_32F images the mains pass."""
    global _ref_seam
    if _ref_seam is None:
        if build_ref() is None:
            raise RuntimeError("oracle/_ref/libref_seam.so is not available (no /root/reference here)")
        _ref_seam = C.CDLL(_REF_SEAM_SO)
        _ref_seam.ref_dp_seam_find.restype = C.c_int
    n = len(images)
    if n == 0:
        return []
    is_u8 = images[0].dtype == np.uint8
    imgs = [np.ascontiguousarray(im, np.uint8 if is_u8 else np.float32) for im in images]
    out = [np.ascontiguousarray(m, np.uint8).copy() for m in masks]
    ip = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
    mp = (C.c_void_p * n)(*[m.ctypes.data for m in out])
    rows = np.asarray([im.shape[0] for im in imgs], np.int32)
    cols = np.asarray([im.shape[1] for im in imgs], np.int32)
    cxy = np.asarray(corners, np.int32).reshape(-1).copy()
    flags = (1 if is_u8 else 0) | (2 if imgs[0].ndim == 3 and imgs[0].shape[2] == 4 else 0)
    rc = _ref_seam.ref_dp_seam_find(C.c_int(n), ip, C.c_int(flags), _p(rows), _p(cols), _p(cxy), mp, C.c_int(cost_fn))
    if rc:
        raise RuntimeError(f"the reference's find() raised cv::Error {rc}")
    return out


def ref_last_seams():
    """Seams the reference's estimateSeam() produced during the last ref_dp_seam_find call, in order:
    [(comp, is_horizontal, points Nx2 in pano coords), ...] -- compare with dp_seam_find(..., want_trace=True)[1]."""
    _ref_seam.ref_last_seam_trace.restype = C.c_size_t
    n = int(_ref_seam.ref_last_seam_trace(None, C.c_size_t(0)))
    t = np.zeros(max(n, 1), np.int32)
    _ref_seam.ref_last_seam_trace(_p(t), C.c_size_t(n))
    res, k = [], 0
    while k + 3 <= n:
        comp, horiz, npts = (int(v) for v in t[k:k + 3])
        res.append((comp, bool(horiz), t[k + 3:k + 3 + 2 * npts].reshape(-1, 2).copy()))
        k += 3 + 2 * npts
    return res


def ref_cylindrical_maps(src_size_wh, K, R, scale):
    """The reference's own detectResultRoi + mapBackward ([WARP]:47-88) -> (roi (tlx, tly, brx, bry), xmap, ymap).
    k_rinv / r_kinv are the oracle's (setCameraParams forms them with OpenCV matrix operators, pinned against cv2)."""
    global _ref_warp
    if _ref_warp is None:
        if build_ref() is None:
            raise RuntimeError("oracle/_ref/libref_warp.so is not available (no /root/reference here)")
        _ref_warp = C.CDLL(_REF_WARP_SO)
    k_rinv, r_kinv = camera_params(K, R)
    _ref_warp.ref_warp_set(_p(np.ascontiguousarray(k_rinv.reshape(9))), _p(np.ascontiguousarray(r_kinv.reshape(9))), C.c_float(scale))
    roi = np.zeros(4, np.int32)
    _ref_warp.ref_detect_roi(C.c_int(src_size_wh[0]), C.c_int(src_size_wh[1]), _p(roi))
    h, w = int(roi[3] - roi[1] + 1), int(roi[2] - roi[0] + 1)
    xmap = np.empty((h, w), np.float32)
    ymap = np.empty((h, w), np.float32)
    _ref_warp.ref_build_maps(C.c_int(int(roi[0])), C.c_int(int(roi[1])), C.c_int(int(roi[2])), C.c_int(int(roi[3])), _p(xmap), _p(ymap))
    return tuple(int(v) for v in roi), xmap, ymap


def ref_lin_blend(img1, img2, tl1, tl2):
    """The reference's compiled block on two CV_32FC3 images -> (pano, seam_x, costV), None for its early return
    ([BLEND]:182-183).  Raises if oracle/_ref is not available."""
    global _ref
    if _ref is None:
        so = build_ref()
        if so is None:
            raise RuntimeError("oracle/_ref/libref_linblend.so is not available (no /root/reference here)")
        _ref = C.CDLL(so)
        _ref.ref_lin_blend.restype = C.c_int
    a = np.ascontiguousarray(img1, np.float32)
    b = np.ascontiguousarray(img2, np.float32)
    he, br = C.c_int(0), C.c_int(0)
    lib().orc_lin_geometry(C.c_int(a.shape[0]), C.c_int(a.shape[1]), C.c_int(b.shape[0]), C.c_int(b.shape[1]),
                           C.c_int(tl1[0]), C.c_int(tl1[1]), C.c_int(tl2[0]), C.c_int(tl2[1]), C.byref(he), C.byref(br))
    ib = a.shape[1] - (tl2[0] - tl1[0])
    pano = np.zeros((max(he.value, 0), max(br.value, 0), 3), np.float32)
    seam = np.zeros(max(he.value, 0), np.int32)
    cost = np.zeros((max(he.value, 0), max(ib + 2, 0)), np.float32)
    rc = _ref.ref_lin_blend(_p(a), C.c_int(a.shape[0]), C.c_int(a.shape[1]), _p(b), C.c_int(b.shape[0]), C.c_int(b.shape[1]),
                            C.c_int(int(tl1[0])), C.c_int(int(tl1[1])), C.c_int(int(tl2[0])), C.c_int(int(tl2[1])),
                            _p(pano), C.c_int(pano.shape[0]), C.c_int(pano.shape[1]), _p(seam), _p(cost), C.c_int(cost.shape[1]))
    if rc == 1:
        return None
    if rc != 0:
        raise RuntimeError(f"ref_lin_blend: geometry mismatch between the reference block and orc_lin_geometry ({rc})")
    return pano, seam, cost


def ref_seam_costs(img1, img2, tl1, tl2, labels, union_tl, l, roi_xywh, cost_fn=COST_COLOR):
    """The reference's own computeCosts ([SEAM]:733-803) -> (costV, costH); same arguments as seam_costs()."""
    global _ref_seam
    if _ref_seam is None:
        ref_dp_seam_find([], [], [])
    is_u8 = img1.dtype == np.uint8
    a = np.ascontiguousarray(img1)
    b = np.ascontiguousarray(img2)
    labels = np.ascontiguousarray(labels, np.int32)
    x, y, w, h = roi_xywh
    costV = np.empty((h, w + 1), np.float32)
    costH = np.empty((h + 1, w), np.float32)
    roi = np.asarray(roi_xywh, np.int32)
    rc = _ref_seam.ref_seam_costs(_p(a), _p(b), C.c_int(1 if is_u8 else 0), C.c_int(a.shape[0]), C.c_int(a.shape[1]),
                                  C.c_int(b.shape[0]), C.c_int(b.shape[1]), C.c_int(tl1[0]), C.c_int(tl1[1]), C.c_int(tl2[0]), C.c_int(tl2[1]),
                                  _p(labels), C.c_int(labels.shape[0]), C.c_int(labels.shape[1]), C.c_int(union_tl[0]), C.c_int(union_tl[1]),
                                  C.c_int(l), _p(roi), C.c_int(cost_fn), _p(costV), _p(costH))
    if rc:
        raise RuntimeError(f"the reference's computeCosts raised cv::Error {rc}")
    return costV, costH


# ---------------------------------------------------------------- whole path
def pipeline_plan(proj, src_sizes_hw, Ks, Rs, scale):
    n = len(src_sizes_hw)
    rows = np.asarray([s[0] for s in src_sizes_hw], np.int32)
    cols = np.asarray([s[1] for s in src_sizes_hw], np.int32)
    K = np.ascontiguousarray(np.asarray(Ks, np.float32).reshape(n, 9))
    R = np.ascontiguousarray(np.asarray(Rs, np.float32).reshape(n, 9))
    corners = np.zeros(2 * n, np.int32)
    sizes = np.zeros(2 * n, np.int32)
    roi = np.zeros(4, np.int32)
    lib().orc_pipeline_plan(C.c_int(n), C.c_int(proj), _p(rows), _p(cols), _p(K), _p(R), C.c_float(scale), _p(corners), _p(sizes), _p(roi))
    return corners.reshape(n, 2), sizes.reshape(n, 2), tuple(int(v) for v in roi)


def dilate_rect(mask, ksize_wh=(20, 20)):
    """cv::dilate with getStructuringElement(MORPH_RECT, ksize) ([SEAM]:1258,1264)"""
    m = np.ascontiguousarray(mask, np.uint8)
    out = np.empty_like(m)
    lib().orc_dilate_rect(_p(m), C.c_int(m.shape[0]), C.c_int(m.shape[1]), C.c_int(ksize_wh[0]), C.c_int(ksize_wh[1]), _p(out))
    return out


def distance_l1(mask):
    m = np.ascontiguousarray(mask, np.uint8)
    out = np.empty(m.shape, np.float32)
    lib().orc_distance_l1(_p(m), C.c_int(m.shape[0]), C.c_int(m.shape[1]), _p(out))
    return out


def feather_weight(mask, sharpness):
    m = np.ascontiguousarray(mask, np.uint8)
    out = np.empty(m.shape, np.float32)
    lib().orc_feather_weight(_p(m), C.c_int(m.shape[0]), C.c_int(m.shape[1]), C.c_float(sharpness), _p(out))
    return out


class FeatherBlender:
    """cv::detail::FeatherBlender ([SEAM]:1249-1252,1271,1280)"""

    def __init__(self, sharpness=0.02):
        self.h = C.c_void_p(lib().orc_fb_create(C.c_float(sharpness)))
        self.roi = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fb_destroy(self.h)
            self.h = None

    def prepare(self, dst_roi_xywh):
        self.roi = tuple(int(v) for v in dst_roi_xywh)
        r = np.asarray(self.roi, np.int32)
        lib().orc_fb_prepare(self.h, _p(r))

    def feed(self, img, mask, tl):
        img = np.ascontiguousarray(img, np.int16)
        mask = np.ascontiguousarray(mask, np.uint8)
        lib().orc_fb_feed(self.h, _p(img), _p(mask), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_int(int(tl[0])), C.c_int(int(tl[1])))

    def blend(self):
        pano = np.empty((self.roi[3], self.roi[2], 3), np.int16)
        pmask = np.empty((self.roi[3], self.roi[2]), np.uint8)
        lib().orc_fb_blend(self.h, _p(pano), _p(pmask))
        return pano, pmask


def gain_feed(corners, images, masks):
    """cv::detail::GainCompensator::feed -> gains (float64[n]); images u8 BGR, masks u8"""
    n = len(images)
    images = [np.ascontiguousarray(a, np.uint8) for a in images]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    rows = np.asarray([a.shape[0] for a in images], np.int32)
    cols = np.asarray([a.shape[1] for a in images], np.int32)
    c = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(-1))
    gains = np.zeros(n, np.float64)
    ip = (C.c_void_p * n)(*[a.ctypes.data for a in images])
    mp = (C.c_void_p * n)(*[a.ctypes.data for a in masks])
    rc = lib().orc_gain_feed(C.c_int(n), ip, mp, _p(rows), _p(cols), _p(c), _p(gains))
    if rc:
        raise RuntimeError("orc_gain_feed: singular system")
    return gains


def gain_apply(image, gain):
    """cv::detail::GainCompensator::apply = multiply(image, gain) on CV_8UC3 -> new array"""
    out = np.ascontiguousarray(image, np.uint8).copy()
    lib().orc_gain_apply(_p(out), C.c_size_t(out.size), C.c_double(float(gain)))
    return out


def pipeline_run(proj, srcs, Ks, Rs, scale, seam=True, num_bands=5, weight_type=WEIGHT_32F, want_intermediates=False, exposure_gain=False,
                 blender="multiband", sharpness=0.02, seam_dilate=0, seam_cost=COST_COLOR):
    """warp -> [gain exposure] -> [DP seam] -> multi-band blend.  Returns dict(pano, pano_mask, corners, sizes, roi, seconds, gains[, warped, masks])."""
    n = len(srcs)
    srcs = [np.ascontiguousarray(s, np.uint8) for s in srcs]
    corners, sizes, roi = pipeline_plan(proj, [s.shape[:2] for s in srcs], Ks, Rs, scale)
    rows = np.asarray([s.shape[0] for s in srcs], np.int32)
    cols = np.asarray([s.shape[1] for s in srcs], np.int32)
    K = np.ascontiguousarray(np.asarray(Ks, np.float32).reshape(n, 9))
    R = np.ascontiguousarray(np.asarray(Rs, np.float32).reshape(n, 9))
    pano = np.empty((roi[3], roi[2], 3), np.int16)
    pmask = np.empty((roi[3], roi[2]), np.uint8)
    secs = np.zeros(4, np.float64)
    sp = (C.c_void_p * n)(*[s.ctypes.data for s in srcs])
    warped = masks = None
    wp = mp = None
    if want_intermediates:
        warped = [np.empty((int(sz[1]), int(sz[0]), 3), np.uint8) for sz in sizes]
        masks = [np.empty((int(sz[1]), int(sz[0])), np.uint8) for sz in sizes]
        wp = (C.c_void_p * n)(*[a.ctypes.data for a in warped])
        mp = (C.c_void_p * n)(*[a.ctypes.data for a in masks])
    c = np.ascontiguousarray(corners.reshape(-1))
    s = np.ascontiguousarray(sizes.reshape(-1))
    r = np.asarray(roi, np.int32)
    gains = np.ones(n, np.float64)
    rc = lib().orc_pipeline_run_ex2(C.c_int(n), C.c_int(proj), sp, _p(rows), _p(cols), _p(K), _p(R), C.c_float(scale),
                                    C.c_int((2 if seam_cost == COST_COLOR_GRAD else 1) if seam else 0), C.c_int(num_bands), C.c_int(weight_type), C.c_int(1 if exposure_gain else 0),
                                    C.c_int(1 if blender == "feather" else 0), C.c_float(sharpness), C.c_int(int(seam_dilate)),
                                    _p(c), _p(s), _p(r), wp, mp, _p(pano), _p(pmask), _p(secs), _p(gains))
    if rc:
        raise RuntimeError(f"orc_pipeline_run failed: {rc}")
    out = dict(pano=pano, pano_mask=pmask, corners=corners, sizes=sizes, roi=roi, seconds=secs, gains=gains)
    if want_intermediates:
        out["warped"] = warped
        out["masks"] = masks
    return out


# ---------------------------------------------------------------- ORB features finder ([FEAT])
def bgr2gray(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(img.shape[:2], np.uint8)
    lib().orc_bgr2gray(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_int(img.shape[2]), C.c_size_t(img.strides[0]), _p(out))
    return out


def resize_linear_exact(gray, dsize_wh):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.empty((dsize_wh[1], dsize_wh[0]), np.uint8)
    lib().orc_resize_linear_exact_u8(_p(gray), C.c_int(gray.shape[0]), C.c_int(gray.shape[1]), _p(out), C.c_int(out.shape[0]), C.c_int(out.shape[1]))
    return out


def fast(gray, threshold=20):
    """cv::FAST(gray, threshold, nonmaxSuppression=true) -> int array n x 3 (x, y, score), raster order"""
    gray = np.ascontiguousarray(gray, np.uint8)
    cap = gray.size // 4 + 16
    out = np.zeros((cap, 3), np.int32)
    lib().orc_fast.restype = C.c_int
    n = lib().orc_fast(_p(gray), C.c_int(gray.shape[0]), C.c_int(gray.shape[1]), C.c_int(threshold), _p(out), C.c_int(cap))
    return out[:n].copy()


def gaussian7(gray):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.empty_like(gray)
    lib().orc_gaussian7_u8(_p(gray), C.c_int(gray.shape[0]), C.c_int(gray.shape[1]), _p(out))
    return out


def fast_atan2(y, x):
    lib().orc_fast_atan2.restype = C.c_float
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def orb_find(img, grid_wh=(3, 1), nfeatures=510, scale_factor=1.3, nlevels=5):
    """find() [FEAT]:948 -> (keypoints n x 6 float32: x, y, size, angle, response, octave; descriptors n x 32 uint8)"""
    img = np.ascontiguousarray(img, np.uint8)
    ch = 1 if img.ndim == 2 else img.shape[2]
    cap = (nfeatures * 2 + 64) * grid_wh[0] * grid_wh[1]
    kps = np.zeros((cap, 6), np.float32)
    desc = np.zeros((cap, 32), np.uint8)
    lib().orc_orb_find.restype = C.c_int
    n = lib().orc_orb_find(_p(img), C.c_int(img.shape[0]), C.c_int(img.shape[1]), C.c_int(ch), C.c_size_t(img.strides[0]), C.c_int(grid_wh[0]),
                           C.c_int(grid_wh[1]), C.c_int(nfeatures), C.c_float(scale_factor), C.c_int(nlevels), _p(kps), _p(desc), C.c_int(cap))
    assert n <= cap
    return kps[:n].copy(), desc[:n].copy()
