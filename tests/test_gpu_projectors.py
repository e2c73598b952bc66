"""CUDA-vs-oracle parity of the projectors the reference keeps as commented-out alternatives ([BLEND]:91-95: cv::PlaneWarper,
FisheyeWarper, StereographicWarper) and of cv::remap on its own ([WARP]:157), through the C ABI.  The oracle is pinned to
cv2.PyRotationWarper / cv2.remap for all of them (tests/test_oracle_cv2.py), incl. NaN / out-of-range map values."""
import numpy as np
import pytest

from helpers import random_camera
from imagestitch_b200 import stitching as S, synth

pytestmark = pytest.mark.gpu


def _eq(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} of {a.size} differ, first at {bad[0].tolist()}: got {a[tuple(bad[0])]} want {b[tuple(bad[0])]}")


@pytest.mark.parametrize("proj", [2, 3, 4])
def test_other_projectors_roi_maps_image_mask(ctx, oracle, proj):
    O = oracle
    rng = np.random.default_rng(30 + proj)
    for t in range(3):
        w, h = int(rng.integers(200, 600)), int(rng.integers(150, 450))
        K, R, scale = random_camera(rng, w, h)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        wp = S.RotationWarper(ctx, proj, scale)
        roi, oxm, oym = O.build_maps(proj, (w, h), K, R, scale)
        (tlx, tly), (dw, dh) = wp.warp_roi((w, h), K, R)
        assert (tlx, tly, tlx + dw - 1, tly + dh - 1) == roi
        groi, xm, ym = wp.buildMaps((w, h), K, R)
        assert groi == (roi[0], roi[1], roi[2] - roi[0], roi[3] - roi[1])
        _eq(xm.view(np.uint32), oxm.view(np.uint32), "xmap bits")
        _eq(ym.view(np.uint32), oym.view(np.uint32), "ymap bits")
        for interp in (O.INTER_LINEAR, O.INTER_NEAREST):
            for border in (O.BORDER_REFLECT, O.BORDER_CONSTANT):
                tl, dst = wp.warp(img, K, R, interp, border)
                assert tl == (roi[0], roi[1])
                _eq(dst, O.remap(img, oxm, oym, interp, border), f"warp proj={proj} interp={interp} border={border}")
        tl, dimg, dmask = wp.warp_with_mask(img, K, R)
        _eq(dimg, O.remap(img, oxm, oym, O.INTER_LINEAR, O.BORDER_REFLECT), "warp_with_mask image")
        _eq(dmask, O.remap(np.full((h, w), 255, np.uint8), oxm, oym, O.INTER_NEAREST, O.BORDER_CONSTANT), "warp_with_mask mask")


def test_remap_incl_extreme_maps(ctx, oracle):
    import torch
    O = oracle
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)
    vals = np.array([np.nan, np.inf, -np.inf, 3e9, -3e9, 1e8, -1e8, 7e7, -7e7, 2 ** 31 / 32, 2 ** 31 / 32 - 4, -2 ** 31 / 32, 1e12, -1e12, 1e20, -1e20,
                     6.7e7, 40.3, -0.5, 0.5, 1.5, 79.5, 78.999, -1.0, 32767.4, 32768.6, -32768.5], np.float32)
    xm = np.ascontiguousarray(np.concatenate([np.tile(vals, (len(vals), 1)), rng.uniform(-200, 300, (len(vals), len(vals))).astype(np.float32)]))
    ym = np.ascontiguousarray(np.concatenate([np.tile(vals, (len(vals), 1)).T, rng.uniform(-150, 250, (len(vals), len(vals))).astype(np.float32)]))
    for src in (img, np.ascontiguousarray(img[:, :, 1])):
        for interp in (O.INTER_LINEAR, O.INTER_NEAREST):
            for border in (O.BORDER_REFLECT, O.BORDER_CONSTANT):
                _eq(S.remap(ctx, src, xm, ym, interp, border), O.remap(src, xm, ym, interp, border), f"remap ch={src.ndim} interp={interp} border={border}")
    big = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    bx = rng.uniform(-50, 450, (500, 700)).astype(np.float32)
    by = rng.uniform(-50, 350, (500, 700)).astype(np.float32)
    got = S.remap(ctx, torch.from_numpy(big).cuda(), torch.from_numpy(bx).cuda(), torch.from_numpy(by).cuda())
    torch.cuda.synchronize()
    _eq(got.cpu().numpy(), O.remap(big, bx, by, O.INTER_LINEAR, O.BORDER_REFLECT), "remap (device buffers)")


@pytest.mark.parametrize("cfg", [("plane", 2, 3, 2.0), ("fisheye", 3, 3, 1.2), ("stereographic", 4, 2, 1.2)])
def test_pipeline_other_projectors(ctx, oracle, cfg):
    """warp -> DP seam -> multi-band blend through the plane projector (fused warp path) and the two per-pixel projectors (maps)"""
    O = oracle
    name, proj, n, fw = cfg
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(n, 320, 240, fw, 0.3)
    want = O.pipeline_run(proj, imgs, Ks, Rs, scale, seam=True, num_bands=4, weight_type=O.WEIGHT_32F, want_intermediates=True)
    got = S.Stitcher(ctx, name, "dp", 4, S.WEIGHT_32F).stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    assert got["roi"] == want["roi"]
    for k in range(n):
        _eq(got["seam_masks"][k], want["masks"][k], f"{name}: seam mask {k}")
    _eq(got["pano_mask"], want["pano_mask"], f"{name}: pano mask")
    _eq(got["pano"], want["pano"], f"{name}: pano")
