"""Live cross-check of the oracle against OpenCV (python cv2) on fresh random cases; skipped when cv2 is
not importable.  Complements tests/test_oracle_golden.py, which needs neither cv2 nor a GPU."""
import numpy as np
import pytest

from helpers import blob_masks, random_camera, warped_set
from imagestitch_b200 import synth

cv2 = pytest.importorskip("cv2")


def test_warp_vs_cv2(oracle):
    O = oracle
    rng = np.random.default_rng(1234)
    for proj, name in ((0, "cylindrical"), (1, "spherical"), (2, "plane"), (3, "fisheye"), (4, "stereographic")):
        for _ in range(4):
            w, h = int(rng.integers(120, 360)), int(rng.integers(100, 300))
            K, R, scale = random_camera(rng, w, h)
            wp = cv2.PyRotationWarper(name, scale)
            roi, xm, ym = wp.buildMaps((w, h), K, R)
            oroi, oxm, oym = O.build_maps(proj, (w, h), K, R, scale, full_scan=True)
            assert (roi[0], roi[1], roi[0] + roi[2], roi[1] + roi[3]) == oroi
            assert O.detect_roi(proj, (w, h), K, R, scale, full_scan=False) == oroi
            if proj >= 2:                                     # the mask warp of the mains ([BLEND]:109) through these projectors too
                tlm, wmk = wp.warp(np.full((h, w), 255, np.uint8), K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
                otlm, owmk = O.warp(proj, np.full((h, w), 255, np.uint8), K, R, scale, O.INTER_NEAREST, O.BORDER_CONSTANT)
                assert tuple(tlm) == otlm and np.array_equal(wmk, owmk)
            assert np.array_equal(xm.view(np.uint32), oxm.view(np.uint32)) and np.array_equal(ym.view(np.uint32), oym.view(np.uint32))
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            tl, wi = wp.warp(img, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
            otl, owi = O.warp(proj, img, K, R, scale, O.INTER_LINEAR, O.BORDER_REFLECT)
            assert tuple(tl) == otl and np.array_equal(wi, owi)


def test_remap_vs_cv2(oracle):
    O = oracle
    rng = np.random.default_rng(4321)
    img = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)
    xm = rng.uniform(-200, 300, (100, 120)).astype(np.float32)
    ym = rng.uniform(-150, 250, (100, 120)).astype(np.float32)
    for interp, ci in ((O.INTER_LINEAR, cv2.INTER_LINEAR), (O.INTER_NEAREST, cv2.INTER_NEAREST)):
        for border, cb in ((O.BORDER_REFLECT, cv2.BORDER_REFLECT), (O.BORDER_CONSTANT, cv2.BORDER_CONSTANT)):
            assert np.array_equal(cv2.remap(img, xm, ym, ci, borderMode=cb), O.remap(img, xm, ym, interp, border))


def test_remap_extreme_maps_vs_cv2(oracle):
    """NaN, infinities and coordinates beyond the int range: cv::remap rounds them with cvtps2dq (INT_MIN); the oracle follows"""
    O = oracle
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)
    vals = np.array([np.nan, np.inf, -np.inf, 3e9, -3e9, 1e8, -1e8, 7e7, -7e7, 2 ** 31 / 32, 2 ** 31 / 32 - 4, -2 ** 31 / 32, 1e12, -1e12, 1e20, -1e20,
                     6.7e7, 40.3, -0.5, 0.5, 1.5, 79.5, 78.999, -1.0, 32767.4, 32768.6, -32768.5], np.float32)
    xm = np.tile(vals, (len(vals), 1))
    ym = xm.T.copy()
    for interp, ci in ((O.INTER_LINEAR, cv2.INTER_LINEAR), (O.INTER_NEAREST, cv2.INTER_NEAREST)):
        for border, cb in ((O.BORDER_REFLECT, cv2.BORDER_REFLECT), (O.BORDER_CONSTANT, cv2.BORDER_CONSTANT)):
            assert np.array_equal(cv2.remap(img, xm, ym, ci, borderMode=cb), O.remap(img, xm, ym, interp, border))


def _cv_pairwise(wi, corners, masks, cost="COLOR"):
    n = len(wi)
    masks = [m.copy() for m in masks]
    for (i, j) in [(i, j) for i in range(n) for j in range(i + 1, n)][::-1]:      # [SEAM]:100-111
        res = cv2.detail_DpSeamFinder(cost).find([cv2.UMat(wi[i].astype(np.float32)), cv2.UMat(wi[j].astype(np.float32))],
                                                    [corners[i], corners[j]], [cv2.UMat(masks[i]), cv2.UMat(masks[j])])
        masks[i], masks[j] = res[0].get(), res[1].get()
    return masks


@pytest.mark.parametrize("case", [(2, 260, 200, 0.25, 1, False), (3, 200, 150, 0.6, 1, False), (4, 160, 120, 0.3, 2, False), (3, 180, 130, 0.4, 1, True)])
def test_seam_and_blend_vs_cv2(oracle, case):
    O = oracle
    n, w, h, ov, rows, irregular = case
    corners, wi, wm = warped_set(O, n, w, h, overlap=ov, grid_rows=rows)
    if irregular:
        holes = blob_masks(np.random.default_rng(8), [m.shape for m in wm], holes=4)
        wm = [np.where(hm > 0, m, 0).astype(np.uint8) for m, hm in zip(wm, holes)]
    want = _cv_pairwise(wi, corners, wm)
    got = O.dp_seam_find(wi, corners, wm)
    for i in range(n):
        assert np.array_equal(got[i], want[i])
    sizes = [(a.shape[1], a.shape[0]) for a in wi]
    roi = O.result_roi(corners, sizes)
    for wt, cwt in ((O.WEIGHT_16S, cv2.CV_16S), (O.WEIGHT_32F, cv2.CV_32F)):
        mb = cv2.detail_MultiBandBlender(0, 5, cwt)
        mb.prepare(roi)
        ob = O.MultiBandBlender(5, wt)
        ob.prepare(roi)
        for i in range(n):
            mb.feed(wi[i].astype(np.int16), got[i], corners[i])
            ob.feed(wi[i].astype(np.int16), got[i], corners[i])
        cd, cm = mb.blend(None, None)
        od, om = ob.blend()
        assert np.array_equal(cm, om)
        if wt == O.WEIGHT_16S:
            assert np.array_equal(cd, od)
        else:
            d = np.abs(cd.astype(np.int32) - od.astype(np.int32))
            assert d.max() <= 2 and (d == 0).mean() >= 0.99


@pytest.mark.parametrize("case", [(2, 260, 200, 0.25, 1, False), (3, 200, 150, 0.6, 1, False), (4, 160, 120, 0.3, 2, False),
                                  (5, 220, 160, 0.35, 1, True)])
def test_color_grad_seam_vs_cv2(oracle, case):
    O = oracle
    n, w, h, ov, rows, irregular = case
    corners, wi, wm = warped_set(O, n, w, h, overlap=ov, grid_rows=rows)
    if irregular:
        holes = blob_masks(np.random.default_rng(8), [m.shape for m in wm], holes=4)
        wm = [np.where(hm > 0, m, 0).astype(np.uint8) for m, hm in zip(wm, holes)]
    want = _cv_pairwise(wi, corners, wm, "COLOR_GRAD")
    for imgs in (wi, [a.astype(np.float32) for a in wi]):
        got = O.dp_seam_find(imgs, corners, wm, cost_fn=O.COST_COLOR_GRAD)
        for i in range(n):
            assert np.array_equal(got[i], want[i])
    rng = np.random.default_rng(n)
    a = (rng.random((h, w, 3)) * 255).astype(np.float32)
    g = cv2.cvtColor(a, cv2.COLOR_BGR2GRAY)
    gx, gy = O.seam_gradients(a)
    assert np.abs(gx - cv2.Sobel(g, cv2.CV_32F, 1, 0)).max() <= 2.5e-4 and np.abs(gy - cv2.Sobel(g, cv2.CV_32F, 0, 1)).max() <= 2.5e-4


def test_gain_compensator_vs_cv2(oracle):
    """cv::detail::GainCompensator (feed / apply): gains to 1e-12 relative (the sums are doubles in raster order on both sides;
    OpenCV 4.13 agrees to the last bit or two), apply() bit for bit."""
    O = oracle
    rng = np.random.default_rng(99)
    for trial in range(6):
        n = int(rng.integers(2, 6))
        imgs, masks, corners, x = [], [], [], 0
        for i in range(n):
            h, w = int(rng.integers(40, 90)), int(rng.integers(60, 120))
            imgs.append(np.clip(rng.integers(0, 256, (h, w, 3)).astype(np.float32) * rng.uniform(0.6, 1.2), 0, 255).astype(np.uint8))
            m = np.full((h, w), 255, np.uint8)
            m[int(rng.integers(0, h // 2)):, :int(rng.integers(1, w // 3))] = 0
            if trial % 2:
                m[::7, ::5] = 128                      # only mask == 255 counts
            masks.append(m)
            corners.append((x, int(rng.integers(-8, 8))))
            x += int(w * rng.uniform(0.5, 0.9))
        c = cv2.detail.ExposureCompensator_createDefault(cv2.detail.ExposureCompensator_GAIN)
        c.feed(corners, imgs, masks)
        g_cv = np.array([float(g[0, 0]) for g in c.getMatGains()])
        g_or = O.gain_feed(corners, imgs, masks)
        assert np.max(np.abs(g_cv - g_or) / np.abs(g_cv)) < 1e-12
        for i in range(n):
            assert np.array_equal(c.apply(i, corners[i], imgs[i].copy(), masks[i]), O.gain_apply(imgs[i], g_cv[i]))
    # apply(): every 8-bit value under many gains, including products that land on .5 in float but not in double
    vals = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, 2)
    for g in list(rng.uniform(0.3, 3.0, 1500)) + [1.0, 0.5, 2.0, 1.5, 1.2208333386655252]:
        c = cv2.detail.ExposureCompensator_createDefault(cv2.detail.ExposureCompensator_GAIN)
        c.setMatGains([np.array([[g]], np.float64)])
        assert np.array_equal(c.apply(0, (0, 0), vals.copy(), np.full((1, 256), 255, np.uint8)), O.gain_apply(vals, g)), g


def test_dilate_distance_feather_vs_cv2(oracle):
    """The mains' live blend path ([SEAM]:1249-1280): dilate 20x20, distanceTransform(L1, 3), FeatherBlender -- bit for bit."""
    O = oracle
    rng = np.random.default_rng(5)
    for t in range(6):
        h, w = int(rng.integers(30, 90)), int(rng.integers(30, 120))
        m = ((rng.random((h, w)) > 0.7).astype(np.uint8) * 255) if t % 2 else blob_masks(rng, [(h, w)])[0]
        for k in ((20, 20), (3, 3), (5, 8), (1, 1)):
            assert np.array_equal(cv2.dilate(m, cv2.getStructuringElement(cv2.MORPH_RECT, k)), O.dilate_rect(m, k)), (t, k)
        assert np.array_equal(cv2.distanceTransform(m, cv2.DIST_L1, 3), O.distance_l1(m)), t
    full = np.full((20, 30), 255, np.uint8)
    assert np.array_equal(cv2.distanceTransform(full, cv2.DIST_L1, 3), O.distance_l1(full))      # FLT_MAX everywhere
    corners, wi, wm = warped_set(O, 3, 320, 240, overlap=0.3)
    sm = O.dp_seam_find(wi, corners, wm)
    sizes = [(a.shape[1], a.shape[0]) for a in wi]
    roi = O.result_roi(corners, sizes)
    for sharp in (0.02, 0.1, 5.0):
        for masks in ([O.dilate_rect(s) & m for s, m in zip(sm, wm)], wm):
            fb = cv2.detail_FeatherBlender(sharp)
            fb.prepare(roi)
            ob = O.FeatherBlender(sharp)
            ob.prepare(roi)
            for i in range(3):
                fb.feed(wi[i].astype(np.int16), masks[i], corners[i])
                ob.feed(wi[i].astype(np.int16), masks[i], corners[i])
            a, am = fb.blend(None, None)
            b, bm = ob.blend()
            assert np.array_equal(am, bm) and np.array_equal(a, b), sharp


def test_mains_sequence_gain_before_seam_vs_cv2(oracle):
    """The whole composite sequence of the reference's main() ([SEAM]:1150-1285) driven through cv2: warp -> GAIN feed ->
    apply IN PLACE -> convertTo(CV_32F) -> DpSeamFinder -> dilate(20x20) & warped mask -> FeatherBlender(0.1).  The seam
    finder must see the COMPENSATED images ([SEAM]:1165-1171 precede :1188-1192); with gains this far from 1 the seam masks
    differ from those of the uncompensated images, which the last assertion checks."""
    O = oracle
    n = 4
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(n, 384, 288, 1.2, 0.25)
    imgs = [np.clip(a.astype(np.float32) * g, 0, 255).astype(np.uint8) for a, g in zip(imgs, (0.7, 1.0, 1.25, 0.85))]
    warper = cv2.PyRotationWarper("cylindrical", scale)
    corners, wi, wm = [], [], []
    for i in range(n):
        tl, a = warper.warp(imgs[i], Ks[i], Rs[i], cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, m = warper.warp(np.full(imgs[i].shape[:2], 255, np.uint8), Ks[i], Rs[i], cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        corners.append(tuple(int(v) for v in tl)); wi.append(a); wm.append(m)
    comp = cv2.detail.ExposureCompensator_createDefault(cv2.detail.ExposureCompensator_GAIN)
    comp.feed(corners, wi, wm)
    raw = [a.copy() for a in wi]
    wi = [comp.apply(i, corners[i], wi[i], wm[i]) for i in range(n)]
    seam = _cv_pairwise(wi, corners, [m.copy() for m in wm], "COLOR")
    seam_raw = _cv_pairwise(raw, corners, [m.copy() for m in wm], "COLOR")
    el = cv2.getStructuringElement(cv2.MORPH_RECT, (20, 20))
    masks = [cv2.dilate(s, el) & m for s, m in zip(seam, wm)]
    sizes = [(a.shape[1], a.shape[0]) for a in wi]
    fb = cv2.detail_FeatherBlender(0.1)
    fb.prepare(O.result_roi(corners, sizes))
    for i in range(n):
        fb.feed(wi[i].astype(np.int16), masks[i], corners[i])
    pano, pmask = fb.blend(None, None)
    got = O.pipeline_run(O.PROJ_CYLINDRICAL, imgs, Ks, Rs, scale, seam=True, want_intermediates=True, exposure_gain=True, blender="feather",
                         sharpness=0.1, seam_dilate=20)
    for i in range(n):
        assert np.array_equal(got["warped"][i], wi[i]), f"compensated image {i}"
        assert np.array_equal(got["masks"][i], masks[i]), f"seam mask {i} (dilated & warped)"
    assert np.array_equal(got["pano_mask"], pmask) and np.array_equal(got["pano"], pano)
    assert any(not np.array_equal(a, b) for a, b in zip(seam, seam_raw)), "the case must distinguish compensated from raw seam inputs"
