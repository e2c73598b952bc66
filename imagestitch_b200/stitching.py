"""Python mirror of the reference's hot-path interface over the C ABI.

Same names and argument meaning as the cv::detail interfaces the reference re-implements
(SURVEY.md 8b), so a parity test reads like one of the reference's own mains:

    warper = RotationWarper(ctx, "cylindrical", scale)            # WarperCreator::create(scale)     [BLEND]:99
    corner, warped = warper.warp(img, K, R, INTER_LINEAR, BORDER_REFLECT)                          # [WARP]:145
    masks = DpSeamFinder(ctx, "COLOR").find(images_f, corners, masks)                              # [SEAM]:87
    blender = MultiBandBlender(ctx, num_bands=5); blender.prepare(corners, sizes)                  # [SEAM]:1244-1252
    blender.feed(img_s, mask, corner); pano, pano_mask = blender.blend()                           # [SEAM]:1271,1280

Arrays may be numpy arrays (host buffers: staged inside the call, results returned as numpy) or torch
CUDA tensors (device buffers: used in place, results returned as torch tensors on the same device).
All computation happens in libimagestitch_b200.so; nothing here computes pixels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .capi import (BORDER_CONSTANT, BORDER_REFLECT, COST_COLOR, COST_COLOR_GRAD, EXPOSURE_GAIN, EXPOSURE_NONE, FEED_BORROW, FEED_COPY,
                   INTER_LINEAR, INTER_NEAREST, PROJ_CYLINDRICAL, PROJ_FISHEYE, PROJ_PLANE, PROJ_SPHERICAL, PROJ_STEREOGRAPHIC, SEAM_DP, SEAM_NONE, WEIGHT_16S, WEIGHT_32F)

_NP_DEPTH = {np.dtype(np.uint8): capi.IS_8U, np.dtype(np.int16): capi.IS_16S, np.dtype(np.int32): capi.IS_32S,
             np.dtype(np.float32): capi.IS_32F}
_PROJ = {"cylindrical": PROJ_CYLINDRICAL, "spherical": PROJ_SPHERICAL, "plane": PROJ_PLANE, "fisheye": PROJ_FISHEYE,
         "stereographic": PROJ_STEREOGRAPHIC}              # the warper creators of [BLEND]:91-95
_PROJ.update({v: v for v in list(_PROJ.values())})


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _torch_depth(t):
    import torch
    return {torch.uint8: capi.IS_8U, torch.int16: capi.IS_16S, torch.int32: capi.IS_32S, torch.float32: capi.IS_32F}[t.dtype]


def as_mat(a):
    """numpy array (H,W[,C]) or torch CUDA tensor -> (capi.Mat, keepalive)."""
    if _is_torch(a):
        assert a.is_cuda, "torch tensors must live on a CUDA device (use numpy for host buffers)"
        assert a.stride(-1) == 1 and (a.dim() == 2 or a.stride(1) == a.shape[2]), "rows must be dense"
        ch = 1 if a.dim() == 2 else a.shape[2]
        m = capi.Mat(a.data_ptr(), a.shape[0], a.shape[1], ch, _torch_depth(a), a.stride(0) * a.element_size(), a.device.index or 0)
        return m, a
    a = np.asarray(a)
    if not (a.strides[-1] == a.itemsize and (a.ndim == 2 or a.strides[1] == a.itemsize * a.shape[2])):
        a = np.ascontiguousarray(a)
    ch = 1 if a.ndim == 2 else a.shape[2]
    m = capi.Mat(a.ctypes.data, a.shape[0], a.shape[1], ch, _NP_DEPTH[a.dtype], a.strides[0], -1)
    return m, a


def _f9(m):
    a = np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(9))
    return a.ctypes.data_as(C.POINTER(C.c_float)), a


def _alloc_like(ref, shape, dtype):
    """Output buffer on the same side (host/device) as `ref`."""
    if _is_torch(ref):
        import torch
        tdt = {np.uint8: torch.uint8, np.int16: torch.int16, np.int32: torch.int32, np.float32: torch.float32}[dtype]
        return torch.empty(shape, dtype=tdt, device=ref.device)
    return np.empty(shape, dtype)


class Context:
    """One is_ctx: a CUDA stream, the workspace pool and the kernel-launch counter."""

    def __init__(self, device: int = 0, use_torch_stream: bool = False):
        self.lib = capi.load()
        h = C.c_void_p()
        st = self.lib.is_ctx_create(int(device), C.byref(h))
        if st != capi.IS_OK:
            raise capi.Error(st, "is_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.device = int(device)
        if use_torch_stream:
            import torch
            self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.is_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, status):
        if status < 0:
            raise capi.Error(status, self.lib.is_ctx_last_error(self.h).decode(errors="replace"))
        return status

    def synchronize(self):
        self.check(self.lib.is_ctx_synchronize(self.h))

    def set_stream(self, cuda_stream_ptr):
        self.check(self.lib.is_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def kernel_timing(self, enable: bool):
        self.check(self.lib.is_ctx_kernel_timing(self.h, 1 if enable else 0))

    def kernel_timing_report(self):
        """[{name, launches, ms, bytes}] per kernel since the last report (CUDA events on the context's stream)."""
        import json
        buf = C.create_string_buffer(1 << 20)      # one call: the report clears the records
        self.lib.is_ctx_kernel_timing_report(self.h, buf, len(buf))
        return json.loads(buf.value.decode() or "[]")

    @property
    def seam_speculation(self) -> int:
        """1: concurrent pair execution accepted, 0: fell back to the sequential loop, -1: sequential by construction."""
        return int(self.lib.is_ctx_seam_speculation(self.h))

    @property
    def seam_path(self) -> int:
        """2 = batched pair loop, 1 = one host thread + stream per pair, 0 = the reference's sequential loop (last call)."""
        return int(self.lib.is_ctx_seam_path(self.h))

    @property
    def seam_waves(self) -> int:
        return int(self.lib.is_ctx_seam_waves(self.h))

    def clear_plan_cache(self):
        """forget the memoised warp plans: the next plan / run scans the image borders again (detectResultRoi)"""
        self.check(self.lib.is_ctx_clear_plan_cache(self.h))

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.is_ctx_kernel_launches(self.h))


class RotationWarper:
    """cv::detail::RotationWarper as the reference restates it in [WARP] (buildMaps :122, warp :145)."""

    def __init__(self, ctx: Context, projection="cylindrical", scale: float = 1.0):
        self.ctx, self.proj, self.scale = ctx, _PROJ[projection], float(scale)

    def warp_roi(self, src_size_wh, K, R):
        """-> ((tl_x, tl_y), (width, height)) of the destination the reference allocates."""
        kp, _k = _f9(K)
        rp, _r = _f9(R)
        tl, sz = capi.Point(), capi.Size()
        self.ctx.check(self.ctx.lib.is_warp_roi(self.ctx.h, self.proj, capi.Size(int(src_size_wh[0]), int(src_size_wh[1])), kp, rp,
                                                self.scale, C.byref(tl), C.byref(sz)))
        return (tl.x, tl.y), (sz.width, sz.height)

    def buildMaps(self, src_size_wh, K, R, like=None):
        """-> (roi (x, y, w, h) as cv::Rect(dst_tl, dst_br), xmap, ymap)"""
        (_, _), (w, h) = self.warp_roi(src_size_wh, K, R)
        xmap = _alloc_like(like, (h, w), np.float32) if like is not None else np.empty((h, w), np.float32)
        ymap = _alloc_like(like, (h, w), np.float32) if like is not None else np.empty((h, w), np.float32)
        mx, _a = as_mat(xmap)
        my, _b = as_mat(ymap)
        kp, _k = _f9(K)
        rp, _r = _f9(R)
        roi = capi.Rect()
        self.ctx.check(self.ctx.lib.is_build_maps(self.ctx.h, self.proj, capi.Size(int(src_size_wh[0]), int(src_size_wh[1])), kp, rp,
                                                  self.scale, C.byref(mx), C.byref(my), C.byref(roi)))
        return (roi.x, roi.y, roi.width, roi.height), xmap, ymap

    def warp(self, src, K, R, interp_mode=INTER_LINEAR, border_mode=BORDER_REFLECT):
        """-> ((tl_x, tl_y), dst)"""
        ms, src_k = as_mat(src)
        (_, _), (w, h) = self.warp_roi((ms.cols, ms.rows), K, R)
        dst = _alloc_like(src_k, (h, w) if ms.channels == 1 else (h, w, ms.channels), np.uint8)
        md, _d = as_mat(dst)
        kp, _k = _f9(K)
        rp, _r = _f9(R)
        tl = capi.Point()
        self.ctx.check(self.ctx.lib.is_warp(self.ctx.h, self.proj, C.byref(ms), kp, rp, self.scale, int(interp_mode), int(border_mode),
                                            C.byref(md), C.byref(tl)))
        return (tl.x, tl.y), dst

    def warp_with_mask(self, src, K, R):
        """The two warp calls of the reference's warp loop ([BLEND]:105,109) fused -> ((tl_x, tl_y), dst, mask)."""
        ms, src_k = as_mat(src)
        (_, _), (w, h) = self.warp_roi((ms.cols, ms.rows), K, R)
        dst = _alloc_like(src_k, (h, w, 3), np.uint8)
        mask = _alloc_like(src_k, (h, w), np.uint8)
        md, _d = as_mat(dst)
        mm, _m = as_mat(mask)
        kp, _k = _f9(K)
        rp, _r = _f9(R)
        tl = capi.Point()
        self.ctx.check(self.ctx.lib.is_warp_with_mask(self.ctx.h, self.proj, C.byref(ms), kp, rp, self.scale, C.byref(md), C.byref(mm), C.byref(tl)))
        return (tl.x, tl.y), dst, mask


def remap(ctx: Context, src, xmap, ymap, interp_mode=INTER_LINEAR, border_mode=BORDER_REFLECT):
    """cv::remap ([WARP]:157) for 8-bit images with 1 or 3 channels through CV_32F maps -> dst of the maps' size"""
    ms, src_k = as_mat(src)
    mx, _x = as_mat(xmap)
    my, _y = as_mat(ymap)
    dst = _alloc_like(src_k, (mx.rows, mx.cols) if ms.channels == 1 else (mx.rows, mx.cols, ms.channels), np.uint8)
    md, _d = as_mat(dst)
    ctx.check(ctx.lib.is_remap(ctx.h, C.byref(ms), C.byref(mx), C.byref(my), int(interp_mode), int(border_mode), C.byref(md)))
    return dst


def imread(ctx: Context, path, like=None):
    """cv::imread(path) for bitmaps ([BLEND]:31-34) -> HxWx3 uint8 (numpy, or a CUDA tensor when `like` is one)"""
    sz, bpp = capi.Size(), C.c_int()
    ctx.check(ctx.lib.is_bmp_info(ctx.h, os.fsencode(path), C.byref(sz), C.byref(bpp)))
    dst = _alloc_like(like, (sz.height, sz.width, 3), np.uint8) if like is not None else np.empty((sz.height, sz.width, 3), np.uint8)
    md, _d = as_mat(dst)
    ctx.check(ctx.lib.is_imread_bmp(ctx.h, os.fsencode(path), C.byref(md)))
    return dst


def imwrite(ctx: Context, path, img):
    """cv::imwrite(path, img) for bitmaps ([BLEND]:717, [SEAM]:1195-1206): uint8 / int16 / float32 with 1 or 3 channels"""
    ms, _k = as_mat(img)
    ctx.check(ctx.lib.is_imwrite_bmp(ctx.h, os.fsencode(path), C.byref(ms)))


def orb_find(ctx: Context, image, grid_wh=(3, 1), nfeatures=510, scale_factor=1.3, nlevels=5):
    """find(image, features) [FEAT]:948 -> (key points n x 6 float32: x, y, size, angle, response, octave; descriptors n x 32 uint8)"""
    ms, _k = as_mat(image)
    prm = capi.OrbParams(int(nfeatures), float(scale_factor), int(nlevels), int(grid_wh[0]), int(grid_wh[1]))
    cap = (2 * int(nfeatures) + 64) * int(grid_wh[0]) * int(grid_wh[1])
    kps = (capi.KeyPoint * cap)()
    desc = np.zeros((cap, 32), np.uint8)
    n = C.c_int()
    ctx.check(ctx.lib.is_orb_find(ctx.h, C.byref(ms), C.byref(prm), kps, desc.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
    assert n.value <= cap
    out = np.array([(k.x, k.y, k.size, k.angle, k.response, k.octave) for k in kps[:n.value]], np.float32).reshape(-1, 6)
    return out, desc[:n.value].copy()


def _mat_array(arrs):
    mats, keep = [], []
    for a in arrs:
        m, k = as_mat(a)
        mats.append(m)
        keep.append(k)
    return (capi.Mat * len(mats))(*mats), keep


class GainCompensator:
    """cv::detail::GainCompensator (ExposureCompensator::createDefault(GAIN)), the compensator of every main:
    feed(corners, images_warped, masks_warped) [BLEND]:117-123, apply(index, corner, image, mask) [SEAM]:1165-1171."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self._gains = None

    def feed(self, corners, images, masks):
        n = len(images)
        im, _k1 = _mat_array(images)
        mk, _k2 = _mat_array(masks)
        pts = (capi.Point * n)(*[capi.Point(int(c[0]), int(c[1])) for c in corners])
        g = (C.c_double * n)()
        self.ctx.check(self.ctx.lib.is_gain_feed(self.ctx.h, n, pts, im, mk, g))
        self._gains = np.array(list(g), np.float64)
        return self._gains

    def gains(self):
        """getMatGains()"""
        return self._gains

    def apply(self, index, corner, image, mask=None):
        """in place, like the reference; also returns the image"""
        if not _is_torch(image) and not image.flags["C_CONTIGUOUS"]:
            raise ValueError("apply() works in place: pass a contiguous array")
        m, _k = as_mat(image)
        self.ctx.check(self.ctx.lib.is_gain_apply(self.ctx.h, C.byref(m), float(self._gains[index])))
        return image


class DpSeamFinder:
    """cv::detail::DpSeamFinder == the free function find() of [SEAM]:87."""

    def __init__(self, ctx: Context, cost_func="COLOR"):
        self.ctx = ctx
        self.cost = {"COLOR": COST_COLOR, "COLOR_GRAD": COST_COLOR_GRAD}.get(cost_func, cost_func)

    def find(self, src, corners, masks, want_trace=False):
        """src: CV_32FC3 or CV_8UC3 images; masks CV_8U.  The masks are modified IN PLACE (as in the
        reference) and also returned.  want_trace -> (masks, [(i, j, comp, horizontal, points Nx2)])."""
        n = len(src)
        if n == 0:
            return (masks, []) if want_trace else masks
        im, _k1 = _mat_array(src)
        for m in masks:
            if not _is_torch(m) and not (isinstance(m, np.ndarray) and m.flags.c_contiguous and m.dtype == np.uint8):
                raise ValueError("find() works in place: pass contiguous uint8 masks")
        mk, _k2 = _mat_array(masks)
        pts = (capi.Point * n)(*[capi.Point(int(c[0]), int(c[1])) for c in corners])
        if not want_trace:
            self.ctx.check(self.ctx.lib.is_seam_dp_find(self.ctx.h, n, im, pts, mk, int(self.cost)))
            return masks
        cap = int(sum(5 + 2 * (m.rows + m.cols) for m in im) * max(1, n) * 2)
        trace = np.zeros(cap, np.int32)
        tlen = C.c_size_t(0)
        self.ctx.check(self.ctx.lib.is_seam_dp_find_trace(self.ctx.h, n, im, pts, mk, int(self.cost),
                                                          trace.ctypes.data_as(C.POINTER(C.c_int32)), cap, C.byref(tlen)))
        res, k = [], 0
        t = trace[: min(cap, tlen.value)]
        while k + 5 <= len(t):
            i, j, comp, horiz, npts = (int(v) for v in t[k:k + 5])
            res.append((i, j, comp, bool(horiz), t[k + 5:k + 5 + 2 * npts].reshape(-1, 2).copy()))
            k += 5 + 2 * npts
        return masks, res

    def pair_run(self, image_i, image_j, tl_i, tl_j, mask_i, mask_j):
        """One pair of the loop on the given input masks -> (out_i, out_j, handle).  The inputs are not modified."""
        mi, _a = as_mat(image_i)
        mj, _b = as_mat(image_j)
        ki, _c = as_mat(mask_i)
        kj, _d = as_mat(mask_j)
        out_i = _alloc_like(mask_i, (ki.rows, ki.cols), np.uint8)
        out_j = _alloc_like(mask_j, (kj.rows, kj.cols), np.uint8)
        oi, _e = as_mat(out_i)
        oj, _f = as_mat(out_j)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.is_seam_pair_run(self.ctx.h, C.byref(mi), C.byref(mj), capi.Point(int(tl_i[0]), int(tl_i[1])),
                                                     capi.Point(int(tl_j[0]), int(tl_j[1])), C.byref(ki), C.byref(kj), C.byref(oi), C.byref(oj),
                                                     C.byref(h)))
        return out_i, out_j, h

    def pair_check(self, image_i, image_j, tl_i, tl_j, mask_i, mask_j, handle) -> bool:
        """True when the pair run that produced `handle` is proven equal to a run on these (true) input masks."""
        mi, _a = as_mat(image_i)
        mj, _b = as_mat(image_j)
        ki, _c = as_mat(mask_i)
        kj, _d = as_mat(mask_j)
        same = C.c_int(0)
        self.ctx.check(self.ctx.lib.is_seam_pair_check(self.ctx.h, C.byref(mi), C.byref(mj), capi.Point(int(tl_i[0]), int(tl_i[1])),
                                                       capi.Point(int(tl_j[0]), int(tl_j[1])), C.byref(ki), C.byref(kj), handle, C.byref(same)))
        return bool(same.value)

    def pair_free(self, handle):
        self.ctx.lib.is_seam_pair_destroy(handle)

    def pair_same_structure(self, mask_i, mask_j_a, mask_j_b, tl_i, tl_j) -> bool:
        """True when pair (i, j) would take the same decisions with mask_j_b in place of mask_j_a (is_seam_pair_same_structure)."""
        ki, _a = as_mat(mask_i)
        ka, _b = as_mat(mask_j_a)
        kb, _c = as_mat(mask_j_b)
        same = C.c_int(0)
        self.ctx.check(self.ctx.lib.is_seam_pair_same_structure(self.ctx.h, C.byref(ki), C.byref(ka), C.byref(kb), capi.Point(int(tl_i[0]), int(tl_i[1])),
                                                                capi.Point(int(tl_j[0]), int(tl_j[1])), C.byref(same)))
        return bool(same.value)

    def mask_and(self, dst, src):
        """dst = 0 where src == 0 (intersection of clear sets), in place."""
        md, _a = as_mat(dst)
        ms, _b = as_mat(src)
        self.ctx.check(self.ctx.lib.is_mask_and(self.ctx.h, C.byref(md), C.byref(ms)))
        return dst

    def cost_maps(self, image1, image2, tl1, tl2, labels, union_tl, label, roi_xywh):
        """computeCosts [SEAM]:733-803 -> (costV h x (w+1), costH (h+1) x w)"""
        x, y, w, h = (int(v) for v in roi_xywh)
        costV = _alloc_like(image1, (h, w + 1), np.float32)
        costH = _alloc_like(image1, (h + 1, w), np.float32)
        m1, _a = as_mat(image1)
        m2, _b = as_mat(image2)
        ml, _c = as_mat(labels)
        mv, _d = as_mat(costV)
        mh, _e = as_mat(costH)
        self.ctx.check(self.ctx.lib.is_seam_cost_maps(self.ctx.h, C.byref(m1), C.byref(m2), capi.Point(*tl1), capi.Point(*tl2), C.byref(ml),
                                                      capi.Point(*union_tl), int(label), capi.Rect(x, y, w, h), C.byref(mv), C.byref(mh)))
        return costV, costH


class MultiBandBlender:
    """cv::detail::MultiBandBlender as the reference's mains call it ([SEAM]:1244-1252,1271,1280)."""

    def __init__(self, ctx: Context, try_gpu=0, num_bands=5, weight_type=WEIGHT_32F):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx.lib.is_blender_create(ctx.h, int(num_bands), int(weight_type), C.byref(h)))
        self.h = h
        self._like = None
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, "h", None) and getattr(self.ctx, "h", None):
                self.ctx.lib.is_blender_destroy(self.h)
            self.h = None
        except Exception:
            pass

    def prepare(self, corners_or_roi, sizes=None):
        if sizes is None:
            x, y, w, h = (int(v) for v in corners_or_roi)
            self.ctx.check(self.ctx.lib.is_blender_prepare_roi(self.h, capi.Rect(x, y, w, h)))
        else:
            n = len(sizes)
            pts = (capi.Point * n)(*[capi.Point(int(c[0]), int(c[1])) for c in corners_or_roi])
            szs = (capi.Size * n)(*[capi.Size(int(s[0]), int(s[1])) for s in sizes])
            self.ctx.check(self.ctx.lib.is_blender_prepare(self.h, n, pts, szs))

    def numBands(self):
        return int(self.ctx.lib.is_blender_num_bands(self.h))

    def feed(self, img, mask, tl, borrow=False, defer=False, key=None):
        """key: position in the feed order (1-based, ascending).  defer=True (device buffers): only the image pyramid is built
        now, on a side stream; the mask is read when blend() / blend_strip() is called and may change until then."""
        mi, ki = as_mat(img)
        mm, km = as_mat(mask)
        self._like = ki
        if borrow or defer:
            self._keep += [ki, km]
        if key is None and not defer:
            self.ctx.check(self.ctx.lib.is_blender_feed(self.h, C.byref(mi), C.byref(mm), capi.Point(int(tl[0]), int(tl[1])),
                                                        FEED_BORROW if borrow else FEED_COPY))
            return
        self._auto_key = getattr(self, "_auto_key", 0) + 1
        flags = (FEED_BORROW if (borrow or defer) else FEED_COPY) | (capi.FEED_DEFER_WEIGHTS if defer else 0)
        self.ctx.check(self.ctx.lib.is_blender_feed_ex(self.h, C.byref(mi), C.byref(mm), capi.Point(int(tl[0]), int(tl[1])), flags,
                                                       int(key) if key is not None else self._auto_key))

    def strip_needs(self, size_wh, tl, x0, x1) -> bool:
        """Does an image of this size / corner contribute to the destination columns [x0, x1)?"""
        needed = C.c_int(0)
        self.ctx.check(self.ctx.lib.is_blender_strip_needs(self.h, capi.Size(int(size_wh[0]), int(size_wh[1])), capi.Point(int(tl[0]), int(tl[1])),
                                                           int(x0), int(x1), C.byref(needed)))
        return bool(needed.value)

    def blend_strip(self, x0, x1, out=None):
        """Columns [x0, x1) of the destination ROI -> (dst H x (x1-x0) x 3 int16, mask)."""
        sz = capi.Size()
        self.ctx.check(self.ctx.lib.is_blender_dst_size(self.h, C.byref(sz)))
        like = self._like if self._like is not None else np.empty(0)
        if out is None:
            dst = _alloc_like(like, (sz.height, x1 - x0, 3), np.int16)
            dmask = _alloc_like(like, (sz.height, x1 - x0), np.uint8)
        else:
            dst, dmask = out
        md, _a = as_mat(dst)
        mm, _b = as_mat(dmask)
        self.ctx.check(self.ctx.lib.is_blender_blend_strip(self.h, int(x0), int(x1), C.byref(md), C.byref(mm)))
        self._keep = []
        return dst, dmask

    def blend(self):
        sz = capi.Size()
        self.ctx.check(self.ctx.lib.is_blender_dst_size(self.h, C.byref(sz)))
        like = self._like if self._like is not None else np.empty(0)
        dst = _alloc_like(like, (sz.height, sz.width, 3), np.int16)
        dmask = _alloc_like(like, (sz.height, sz.width), np.uint8)
        md, _a = as_mat(dst)
        mm, _b = as_mat(dmask)
        self.ctx.check(self.ctx.lib.is_blender_blend(self.h, C.byref(md), C.byref(mm)))
        self._keep = []
        return dst, dmask


def dilate_and(ctx: Context, mask, ksize_wh=(20, 20), and_mask=None):
    """dilate(mask, mask, getStructuringElement(MORPH_RECT, ksize)); mask &= and_mask   ([SEAM]:1258-1269).  In place."""
    if not _is_torch(mask) and not mask.flags["C_CONTIGUOUS"]:
        raise ValueError("dilate_and() works in place: pass a contiguous array")
    m, _k = as_mat(mask)
    if and_mask is None:
        ctx.check(ctx.lib.is_mask_dilate_and(ctx.h, C.byref(m), int(ksize_wh[0]), int(ksize_wh[1]), None))
    else:
        a, _k2 = as_mat(and_mask)
        ctx.check(ctx.lib.is_mask_dilate_and(ctx.h, C.byref(m), int(ksize_wh[0]), int(ksize_wh[1]), C.byref(a)))
    return mask


def feather_weight_map(ctx: Context, mask, sharpness):
    """createWeightMap(mask, sharpness, weight) of the feather blender -> float32 map"""
    w = _alloc_like(mask, tuple(mask.shape[:2]), np.float32)
    m, _k = as_mat(mask)
    mw, _k2 = as_mat(w)
    ctx.check(ctx.lib.is_feather_weight_map(ctx.h, C.byref(m), float(sharpness), C.byref(mw)))
    return w


class FeatherBlender:
    """cv::detail::FeatherBlender as the reference's mains call it ([SEAM]:1249-1252,1271,1280)."""

    def __init__(self, ctx: Context, sharpness=0.02):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx.lib.is_feather_create(ctx.h, float(sharpness), C.byref(h)))
        self.h = h
        self._like = None

    def __del__(self):
        try:
            if getattr(self, "h", None) and getattr(self.ctx, "h", None):
                self.ctx.lib.is_feather_destroy(self.h)
            self.h = None
        except Exception:
            pass

    def prepare(self, corners_or_roi, sizes=None):
        if sizes is None:
            x, y, w, h = (int(v) for v in corners_or_roi)
            self.ctx.check(self.ctx.lib.is_feather_prepare_roi(self.h, capi.Rect(x, y, w, h)))
        else:
            n = len(sizes)
            pts = (capi.Point * n)(*[capi.Point(int(c[0]), int(c[1])) for c in corners_or_roi])
            szs = (capi.Size * n)(*[capi.Size(int(s[0]), int(s[1])) for s in sizes])
            self.ctx.check(self.ctx.lib.is_feather_prepare(self.h, n, pts, szs))

    def feed(self, img, mask, tl):
        mi, ki = as_mat(img)
        mm, _km = as_mat(mask)
        self._like = ki
        self.ctx.check(self.ctx.lib.is_feather_feed(self.h, C.byref(mi), C.byref(mm), capi.Point(int(tl[0]), int(tl[1]))))

    def blend(self):
        sz = capi.Size()
        self.ctx.check(self.ctx.lib.is_feather_dst_size(self.h, C.byref(sz)))
        like = self._like if self._like is not None else np.empty(0)
        dst = _alloc_like(like, (sz.height, sz.width, 3), np.int16)
        dmask = _alloc_like(like, (sz.height, sz.width), np.uint8)
        md, _a = as_mat(dst)
        mm, _b = as_mat(dmask)
        self.ctx.check(self.ctx.lib.is_feather_blend(self.h, C.byref(md), C.byref(mm)))
        return dst, dmask


def linear_blend_pair(ctx: Context, img1, img2, tl1, tl2):
    """The reference's hand-written pair blend [BLEND]:141-717 -> (pano float32 HxWx3, seam_x) or None."""
    m1, k1 = as_mat(img1)
    m2, k2 = as_mat(img2)
    sz = capi.Size()
    ctx.lib.is_linear_blend_size(capi.Size(m1.cols, m1.rows), capi.Size(m2.cols, m2.rows), capi.Point(*tl1), capi.Point(*tl2), C.byref(sz))
    pano = _alloc_like(k1, (sz.height, sz.width, 3), np.float32)
    mp, _p = as_mat(pano)
    seam = np.zeros(sz.height, np.int32)
    st = ctx.check(ctx.lib.is_linear_blend_pair(ctx.h, C.byref(m1), C.byref(m2), capi.Point(*tl1), capi.Point(*tl2), C.byref(mp),
                                                seam.ctypes.data_as(C.POINTER(C.c_int))))
    if st == 1:
        return None
    return pano, seam


class Stitcher:
    """The composite call sequence detect -> match -> homography -> warp -> seam -> blend (is_pipeline_run).
    Registration stages are host control flow: pass cameras, or Python callables as hooks."""

    def __init__(self, ctx: Context, projection="cylindrical", seam="dp", num_bands=5, weight_type=WEIGHT_32F, exposure=None,
                 blender="multiband", sharpness=0.02, seam_dilate=0, seam_cost="COLOR"):
        """blender="feather", sharpness=0.1, seam_dilate=20, exposure="gain" is the configuration the reference's mains run."""
        self.ctx = ctx
        self.cfg = capi.PipelineConfig(_PROJ[projection], SEAM_DP if seam in ("dp", SEAM_DP, True) else SEAM_NONE,
                                       COST_COLOR_GRAD if seam_cost in ("COLOR_GRAD", COST_COLOR_GRAD) else COST_COLOR,
                                       int(num_bands), int(weight_type), 1.0, EXPOSURE_GAIN if exposure in ("gain", EXPOSURE_GAIN, True) else EXPOSURE_NONE,
                                       capi.BLEND_FEATHER if blender in ("feather", capi.BLEND_FEATHER) else capi.BLEND_MULTI_BAND, float(sharpness),
                                       int(seam_dilate))
        self.timings_ms = None

    @staticmethod
    def _cameras(Ks, Rs):
        n = len(Ks)
        cams = (capi.Camera * n)()
        for i in range(n):
            cams[i].K[:] = [float(v) for v in np.asarray(Ks[i], np.float32).reshape(9)]
            cams[i].R[:] = [float(v) for v in np.asarray(Rs[i], np.float32).reshape(9)]
        return cams

    def plan(self, src_sizes_wh, Ks, Rs, scale):
        n = len(src_sizes_wh)
        self.cfg.scale = float(scale)
        cams = self._cameras(Ks, Rs)
        szs = (capi.Size * n)(*[capi.Size(int(s[0]), int(s[1])) for s in src_sizes_wh])
        corners = (capi.Point * n)()
        sizes = (capi.Size * n)()
        roi = capi.Rect()
        self.ctx.check(self.ctx.lib.is_pipeline_plan(self.ctx.h, n, szs, cams, C.byref(self.cfg), corners, sizes, C.byref(roi)))
        return ([(c.x, c.y) for c in corners], [(s.width, s.height) for s in sizes], (roi.x, roi.y, roi.width, roi.height))

    def stitch(self, images, Ks, Rs, scale, want_seam_masks=False, out=None, hooks=None):
        """-> dict(pano, pano_mask, corners, sizes, roi[, seam_masks]).  `out` = (pano, pano_mask) buffers to reuse."""
        n = len(images)
        im, keep = _mat_array(images)
        corners, sizes, roi = self.plan([(m.cols, m.rows) for m in im], Ks, Rs, scale)
        cams = self._cameras(Ks, Rs)
        like = keep[0]
        if out is None:
            pano = _alloc_like(like, (roi[3], roi[2], 3), np.int16)
            pmask = _alloc_like(like, (roi[3], roi[2]), np.uint8)
        else:
            pano, pmask = out
        mp, _a = as_mat(pano)
        mm, _b = as_mat(pmask)
        sm = None
        seam_masks = None
        if want_seam_masks:
            seam_masks = [_alloc_like(like, (s[1], s[0]), np.uint8) for s in sizes]
            sm, _c = _mat_array(seam_masks)
        self.ctx.check(self.ctx.lib.is_pipeline_run(self.ctx.h, n, im, cams, hooks, C.byref(self.cfg), C.byref(mp), C.byref(mm), sm))
        t = (C.c_float * 4)()
        self.ctx.lib.is_pipeline_last_timings(self.ctx.h, t)
        self.timings_ms = dict(warp=t[0], seam=t[1], blend=t[2], total=t[3])
        res = dict(pano=pano, pano_mask=pmask, corners=corners, sizes=sizes, roi=roi)
        if self.cfg.exposure == EXPOSURE_GAIN:
            g = (C.c_double * n)()
            self.ctx.check(self.ctx.lib.is_pipeline_last_gains(self.ctx.h, n, g))
            res["gains"] = np.array(list(g), np.float64)
        if want_seam_masks:
            res["seam_masks"] = seam_masks
        return res
