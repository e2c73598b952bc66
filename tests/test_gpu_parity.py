"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on the same inputs.

Bars: bit-exact for ROI/corners, maps, warped images and masks, cost maps, seam point
lists, seam masks, linear-blend seam indices and the CV_16S-weight multi-band blend; the CV_32F-weight
blend is bit-exact against the oracle as well (same association order on both sides -- the tolerance
max|d| <= 2 int16 units / >= 99 % exact applies to oracle-vs-OpenCV, tests/test_oracle_cv2.py).
The linear blend's float panorama: |d| <= 1e-3 * 255.
"""
import numpy as np
import pytest

from helpers import blob_masks, random_camera, warped_set
from imagestitch_b200 import stitching as S, synth

pytestmark = pytest.mark.gpu


def _eq(a, b, what):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} of {a.size} differ, first at {bad[0].tolist()}: got {a[tuple(bad[0])]} want {b[tuple(bad[0])]}")


@pytest.mark.parametrize("proj", [0, 1])
def test_warp_roi_maps_image_mask(ctx, oracle, proj):
    O = oracle
    rng = np.random.default_rng(10 + proj)
    for t in range(4):
        w, h = int(rng.integers(200, 700)), int(rng.integers(150, 500))
        K, R, scale = random_camera(rng, w, h)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        wp = S.RotationWarper(ctx, proj, scale)
        roi, oxm, oym = O.build_maps(proj, (w, h), K, R, scale, full_scan=True)   # the reference's full scan
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
                _eq(dst, O.remap(img, oxm, oym, interp, border), f"warp interp={interp} border={border}")
        g = img[:, :, 1].copy()
        _, dst1 = wp.warp(g, K, R, O.INTER_LINEAR, O.BORDER_REFLECT)
        _eq(dst1, O.remap(g, oxm, oym, O.INTER_LINEAR, O.BORDER_REFLECT), "warp 8UC1")
        tl, dimg, dmask = wp.warp_with_mask(img, K, R)
        _eq(dimg, O.remap(img, oxm, oym, O.INTER_LINEAR, O.BORDER_REFLECT), "warp_with_mask image")
        _eq(dmask, O.remap(np.full((h, w), 255, np.uint8), oxm, oym, O.INTER_NEAREST, O.BORDER_CONSTANT), "warp_with_mask mask")


def test_warp_device_buffers(ctx, oracle):
    import torch
    O = oracle
    rng = np.random.default_rng(5)
    w, h = 640, 480
    K, R, scale = random_camera(rng, w, h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    wp = S.RotationWarper(ctx, 0, scale)
    timg = torch.from_numpy(img).cuda()
    torch.cuda.synchronize()
    tl, dimg, dmask = wp.warp_with_mask(timg, K, R)
    ctx.synchronize()
    otl, oimg = O.warp(0, img, K, R, scale, O.INTER_LINEAR, O.BORDER_REFLECT)
    assert tl == otl
    _eq(dimg.cpu().numpy(), oimg, "device warp")


SEAM_CASES = [(2, 400, 300, 0.25, 1), (3, 320, 240, 0.3, 1), (4, 256, 200, 0.25, 1), (2, 300, 400, 0.5, 1), (3, 300, 200, 0.6, 1),
              (4, 240, 180, 0.3, 2)]


@pytest.mark.parametrize("case", SEAM_CASES)
@pytest.mark.parametrize("as_float", [False, True])
def test_dp_seam_masks_and_seams(ctx, oracle, case, as_float):
    O = oracle
    n, w, h, ov, grid_rows = case
    corners, wi, wm = warped_set(O, n, w, h, overlap=ov, grid_rows=grid_rows)
    imgs = [a.astype(np.float32) for a in wi] if as_float else wi
    want, wtrace = O.dp_seam_find(imgs, corners, wm, want_trace=True)
    got, gtrace = S.DpSeamFinder(ctx, "COLOR").find(imgs, corners, [m.copy() for m in wm], want_trace=True)
    assert len(gtrace) == len(wtrace), f"number of seams {len(gtrace)} vs {len(wtrace)}"
    for (gi, gj, gc, gh, gp), (oi, oj, oc, oh, op) in zip(gtrace, wtrace):
        assert (gi, gj, gc, gh) == (oi, oj, oc, oh)
        _eq(gp, op, f"seam points of pair ({gi},{gj})")
    for k in range(n):
        _eq(got[k], want[k], f"seam mask {k}")


def test_dp_seam_irregular_masks(ctx, oracle):
    """Holes and notches: many components, single-neighbour relabelling, unreachable seams."""
    O = oracle
    rng = np.random.default_rng(99)
    for t in range(4):
        corners, wi, wm = warped_set(O, 3, 220, 160, overlap=0.4)
        holes = blob_masks(rng, [m.shape for m in wm], holes=4)
        wm = [np.where(hm > 0, m, 0).astype(np.uint8) for m, hm in zip(wm, holes)]
        want = O.dp_seam_find(wi, corners, wm)
        got = S.DpSeamFinder(ctx, "COLOR").find(wi, corners, [m.copy() for m in wm])
        for k in range(3):
            _eq(got[k], want[k], f"irregular case {t} mask {k}")


def test_dp_seam_edge_cases(ctx, oracle):
    O = oracle
    f = S.DpSeamFinder(ctx, "COLOR")
    assert f.find([], [], []) == []                                    # [SEAM]:94-95
    a = np.zeros((40, 50, 3), np.uint8)
    m = [np.full((40, 50), 255, np.uint8), np.full((40, 50), 255, np.uint8)]
    got = f.find([a, a], [(0, 0), (100, 0)], [x.copy() for x in m])    # disjoint ROIs: [SEAM]:142-143
    _eq(got[0], m[0], "disjoint 0")
    _eq(got[1], m[1], "disjoint 1")
    # identical placement (full overlap), flat images: all costs tie -> exercises the (cost, step) tie-break
    want = O.dp_seam_find([a, a], [(0, 0), (20, 7)], m)
    got = f.find([a, a], [(0, 0), (20, 7)], [x.copy() for x in m])
    _eq(got[0], want[0], "flat 0")
    _eq(got[1], want[1], "flat 1")
    want = O.dp_seam_find([a, a], [(0, 0), (20, 7)], m, cost_fn=O.COST_COLOR_GRAD)
    got = S.DpSeamFinder(ctx, "COLOR_GRAD").find([a, a], [(0, 0), (20, 7)], [x.copy() for x in m])
    _eq(got[0], want[0], "flat 0 (COLOR_GRAD)")
    _eq(got[1], want[1], "flat 1 (COLOR_GRAD)")


def test_seam_cost_maps(ctx, oracle):
    O = oracle
    rng = np.random.default_rng(3)
    h1, w1, h2, w2 = 60, 80, 70, 64
    tl1, tl2 = (5, -3), (40, 4)
    utl = (min(tl1[0], tl2[0]), min(tl1[1], tl2[1]))
    ubr = (max(tl1[0] + w1, tl2[0] + w2), max(tl1[1] + h1, tl2[1] + h2))
    W, H = ubr[0] - utl[0], ubr[1] - utl[1]
    labels = np.zeros((H, W), np.int32)
    ix0, iy0 = tl2[0] - utl[0], tl2[1] - utl[1]
    ix1, iy1 = tl1[0] + w1 - utl[0], tl1[1] + h1 - utl[1]
    labels[iy0:iy1, ix0:ix1] = 2
    labels[iy0 + 5:iy0 + 9, ix0 + 3:ix0 + 10] = 1          # a hole of another label inside the component
    roi = (ix0, iy0, ix1 - ix0, iy1 - iy0)
    for dt in (np.uint8, np.float32):
        a = rng.integers(0, 256, (h1, w1, 3)).astype(dt)
        b = rng.integers(0, 256, (h2, w2, 3)).astype(dt)
        if dt == np.float32:
            a += rng.uniform(-0.5, 0.5, a.shape).astype(np.float32)
            b += rng.uniform(-0.5, 0.5, b.shape).astype(np.float32)
        wv, wh = O.seam_costs(a, b, tl1, tl2, labels, utl, 2, roi)
        gv, gh = S.DpSeamFinder(ctx, "COLOR").cost_maps(a, b, tl1, tl2, labels, utl, 2, roi)
        _eq(gv.view(np.uint32), wv.view(np.uint32), f"costV {dt.__name__}")
        _eq(gh.view(np.uint32), wh.view(np.uint32), f"costH {dt.__name__}")


BLEND_CASES = [(2, 400, 300, 0.25, 5), (3, 320, 240, 0.3, 5), (4, 256, 200, 0.25, 3), (2, 300, 400, 0.5, 6), (2, 200, 150, 0.3, 0)]


@pytest.mark.parametrize("case", BLEND_CASES)
@pytest.mark.parametrize("wt", [S.WEIGHT_32F, S.WEIGHT_16S])
@pytest.mark.parametrize("u8", [False, True])
def test_multiband_blend(ctx, oracle, case, wt, u8):
    O = oracle
    n, w, h, ov, nb = case
    corners, wi, wm = warped_set(O, n, w, h, overlap=ov)
    sm = O.dp_seam_find(wi, corners, wm)
    sizes = [(a.shape[1], a.shape[0]) for a in wi]
    for masks in (sm, wm):
        ob = O.MultiBandBlender(nb, wt)
        ob.prepare(O.result_roi(corners, sizes))
        gb = S.MultiBandBlender(ctx, 0, nb, wt)
        gb.prepare(corners, sizes)
        for i in range(n):
            ob.feed(wi[i].astype(np.int16), masks[i], corners[i])
            gb.feed(wi[i] if u8 else wi[i].astype(np.int16), masks[i], corners[i])
        assert gb.numBands() == ob.num_bands()
        want, wmask = ob.blend()
        got, gmask = gb.blend()
        _eq(gmask, wmask, "blend mask")
        _eq(got, want, f"blend nb={nb} wt={wt}")


@pytest.mark.parametrize("wt", [S.WEIGHT_32F, S.WEIGHT_16S])
def test_multiband_blend_gray_masks(ctx, oracle, wt):
    """Masks with arbitrary byte values (fractional level-0 weights), overlapping everywhere: the general
    multiply / divide path of the tiled level-0 kernel next to its mask == 255 fast path."""
    O = oracle
    rng = np.random.default_rng(5)
    corners, wi, wm = warped_set(O, 3, 320, 240, overlap=0.3)
    sizes = [(a.shape[1], a.shape[0]) for a in wi]
    vals = np.array([0, 1, 77, 128, 254, 255, 255, 255], np.uint8)
    masks = []
    for m in wm:
        blocks = vals[rng.integers(0, len(vals), (m.shape[0] // 16 + 1, m.shape[1] // 16 + 1))]
        g = np.kron(blocks, np.ones((16, 16), np.uint8))[:m.shape[0], :m.shape[1]]
        masks.append(np.where(m > 0, g, 0).astype(np.uint8))
    ob = O.MultiBandBlender(5, wt)
    ob.prepare(O.result_roi(corners, sizes))
    gb = S.MultiBandBlender(ctx, 0, 5, wt)
    gb.prepare(corners, sizes)
    for i in range(3):
        ob.feed(wi[i].astype(np.int16), masks[i], corners[i])
        gb.feed(wi[i], masks[i], corners[i])
    want, wmask = ob.blend()
    got, gmask = gb.blend()
    _eq(gmask, wmask, "blend mask (gray masks)")
    _eq(got, want, f"blend gray masks wt={wt}")


def test_gain_compensator(ctx, oracle):
    """GainCompensator feed / apply through the C ABI vs the oracle: gains to 1e-12 relative (tree sum on the device, raster
    sum on the CPU), applied images bit for bit; host and device buffers."""
    import torch
    O = oracle
    corners, wi, wm = warped_set(O, 4, 320, 240, overlap=0.3)
    rng = np.random.default_rng(3)
    wi = [np.clip(a.astype(np.float32) * g, 0, 255).astype(np.uint8) for a, g in zip(wi, (0.8, 1.0, 1.15, 0.95))]
    wm = [m.copy() for m in wm]
    wm[1][10:40, 5:60] = 0
    wm[2][::9, ::4] = 77                                   # only mask == 255 counts
    want = O.gain_feed(corners, wi, wm)
    gc = S.GainCompensator(ctx)
    got = gc.feed(corners, wi, wm)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-12, (got, want)
    dev = S.GainCompensator(ctx).feed(corners, [torch.from_numpy(a).cuda() for a in wi], [torch.from_numpy(m).cuda() for m in wm])
    assert np.array_equal(dev, got), "device buffers give different gains than host buffers"
    for i in range(4):
        a = wi[i].copy()
        gc.apply(i, corners[i], a, wm[i])
        _eq(a, O.gain_apply(wi[i], got[i]), f"apply {i}")
    vals = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, 2)
    for g in list(rng.uniform(0.3, 3.0, 200)) + [1.2208333386655252, 0.5, 1.5]:
        gc._gains = np.array([g])
        t = torch.from_numpy(vals.copy()).cuda()
        gc.apply(0, (0, 0), t)
        torch.cuda.synchronize()
        _eq(t.cpu().numpy(), O.gain_apply(vals, g), f"apply gain {g!r}")


def test_pipeline_with_gain_exposure(ctx, oracle):
    O = oracle
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(4, 384, 288, 1.2, 0.25)
    imgs = [np.clip(a.astype(np.float32) * g, 0, 255).astype(np.uint8) for a, g in zip(imgs, (0.85, 1.0, 1.1, 0.9))]
    want = O.pipeline_run(O.PROJ_CYLINDRICAL, imgs, Ks, Rs, scale, seam=True, num_bands=5, weight_type=O.WEIGHT_32F, want_intermediates=True,
                          exposure_gain=True)
    got = S.Stitcher(ctx, "cylindrical", "dp", 5, S.WEIGHT_32F, exposure="gain").stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    assert np.max(np.abs(got["gains"] - want["gains"]) / want["gains"]) < 1e-12
    assert np.ptp(want["gains"]) > 0.05, "the test images should need visibly different gains"
    for k in range(4):
        _eq(got["seam_masks"][k], want["masks"][k], f"seam mask {k}")
    _eq(got["pano_mask"], want["pano_mask"], "panorama mask")
    _eq(got["pano"], want["pano"], "panorama (gain compensated)")


def test_dilate_and_feather_weights(ctx, oracle):
    """Mask preparation of the mains ([SEAM]:1257-1270) and createWeightMap: dilate / & / L1 distance weights, bit for bit."""
    import torch
    O = oracle
    rng = np.random.default_rng(8)
    for t in range(5):
        h, w = int(rng.integers(30, 300)), int(rng.integers(30, 400))
        m = ((rng.random((h, w)) > 0.6).astype(np.uint8) * 255) if t % 2 else blob_masks(rng, [(h, w)])[0]
        other = blob_masks(rng, [(h, w)])[0]
        for k in ((20, 20), (3, 5), (1, 1), (33, 7)):
            a = m.copy()
            S.dilate_and(ctx, a, k, other)
            _eq(a, O.dilate_rect(m, k) & other, f"dilate {k} & mask")
            b = torch.from_numpy(m.copy()).cuda()
            S.dilate_and(ctx, b, k)
            torch.cuda.synchronize()
            _eq(b.cpu().numpy(), O.dilate_rect(m, k), f"dilate {k} (device buffer)")
        for sharp in (0.02, 0.1, 5.0):
            _eq(S.feather_weight_map(ctx, m, sharp).view(np.uint32), O.feather_weight(m, sharp).view(np.uint32), f"weight map sharpness {sharp}")
    full = np.full((40, 50), 255, np.uint8)          # no zero pixel anywhere: distance FLT_MAX, weight 1
    _eq(S.feather_weight_map(ctx, full, 0.1), O.feather_weight(full, 0.1), "weight map of a mask without zeros")


@pytest.mark.parametrize("sharp", [0.02, 0.1])
@pytest.mark.parametrize("u8", [False, True])
def test_feather_blend(ctx, oracle, sharp, u8):
    """The mains' live blend path: dilated seam masks & warped masks, FeatherBlender(sharpness) prepare / feed / blend."""
    O = oracle
    corners, wi, wm = warped_set(O, 4, 320, 240, overlap=0.3)
    sm = O.dp_seam_find(wi, corners, wm)
    sizes = [(a.shape[1], a.shape[0]) for a in wi]
    for masks in ([O.dilate_rect(s) & m for s, m in zip(sm, wm)], wm):
        ob = O.FeatherBlender(sharp)
        ob.prepare(O.result_roi(corners, sizes))
        gb = S.FeatherBlender(ctx, sharp)
        gb.prepare(corners, sizes)
        for i in range(4):
            ob.feed(wi[i].astype(np.int16), masks[i], corners[i])
            gb.feed(wi[i] if u8 else wi[i].astype(np.int16), masks[i], corners[i])
        want, wmask = ob.blend()
        got, gmask = gb.blend()
        _eq(gmask, wmask, "feather mask")
        _eq(got, want, f"feather blend sharpness={sharp}")


@pytest.mark.parametrize("cfg", [dict(blender="feather", sharpness=0.1, seam_dilate=20, exposure_gain=True),      # what the reference's mains run
                                 dict(blender="feather", sharpness=0.02, seam_dilate=0, exposure_gain=False),
                                 dict(blender="multiband", sharpness=0.02, seam_dilate=20, exposure_gain=True)])
def test_pipeline_reference_configuration(ctx, oracle, cfg):
    """warp -> gain exposure -> DP seam -> dilate(20x20) & warped mask -> feather blend(0.1): the sequence of [SEAM]:1094-1285."""
    O = oracle
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(4, 384, 288, 1.2, 0.25)
    imgs = [np.clip(a.astype(np.float32) * g, 0, 255).astype(np.uint8) for a, g in zip(imgs, (0.85, 1.0, 1.1, 0.9))]
    want = O.pipeline_run(O.PROJ_CYLINDRICAL, imgs, Ks, Rs, scale, seam=True, num_bands=5, weight_type=O.WEIGHT_32F, want_intermediates=True, **cfg)
    st = S.Stitcher(ctx, "cylindrical", "dp", 5, S.WEIGHT_32F, exposure="gain" if cfg["exposure_gain"] else None, blender=cfg["blender"],
                    sharpness=cfg["sharpness"], seam_dilate=cfg["seam_dilate"])
    got = st.stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    for k in range(4):
        _eq(got["seam_masks"][k], want["masks"][k], f"seam mask {k}")
    _eq(got["pano_mask"], want["pano_mask"], "panorama mask")
    _eq(got["pano"], want["pano"], f"panorama {cfg}")
    import torch
    dev = st.stitch([torch.from_numpy(a).cuda() for a in imgs], Ks, Rs, scale)
    torch.cuda.synchronize()
    _eq(dev["pano"].cpu().numpy(), want["pano"], "panorama (device buffers)")


def test_linear_blend_pair(ctx, oracle):
    O = oracle
    for (w, h, ov) in ((400, 300, 0.25), (320, 260, 0.4)):
        corners, wi, _ = warped_set(O, 2, w, h, overlap=ov)
        a, b = wi[0].astype(np.float32), wi[1].astype(np.float32)
        for tl2 in (corners[1], (corners[1][0], corners[0][1]), (corners[1][0], corners[0][1] + 6), (corners[1][0], corners[0][1] - 5)):
            want = O.lin_blend(a, b, corners[0], tl2)
            got = S.linear_blend_pair(ctx, a, b, corners[0], tl2)
            assert (want is None) == (got is None)
            if want is None:
                continue
            _eq(got[1], want[1], f"greedy seam indices tl2={tl2}")
            assert got[0].shape == want[0].shape
            d = np.abs(got[0] - want[0])
            assert np.nanmax(d) <= 1e-3 * 255, f"linear blend pano max diff {np.nanmax(d)}"
            assert np.array_equal(np.isnan(got[0]), np.isnan(want[0]))
    assert S.linear_blend_pair(ctx, a, b, (0, 0), (5000, 0)) is None     # [BLEND]:182-183


@pytest.mark.parametrize("cfg", [(3, 384, 288, 1.2, 5, S.WEIGHT_32F, 0), (2, 512, 384, 1.2, 5, S.WEIGHT_16S, 0), (4, 300, 220, 1.5, 4, S.WEIGHT_32F, 1)])
def test_pipeline_end_to_end(ctx, oracle, cfg):
    O = oracle
    n, w, h, fw, nb, wt, proj = cfg
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(n, w, h, fw, 0.25)
    want = O.pipeline_run(proj, imgs, Ks, Rs, scale, seam=True, num_bands=nb, weight_type=wt, want_intermediates=True)
    st = S.Stitcher(ctx, proj, "dp", nb, wt)
    got = st.stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    assert got["roi"] == want["roi"]
    assert [tuple(c) for c in got["corners"]] == [tuple(int(v) for v in c) for c in want["corners"]]
    for k in range(n):
        _eq(got["seam_masks"][k], want["masks"][k], f"pipeline seam mask {k}")
    _eq(got["pano_mask"], want["pano_mask"], "pano mask")
    _eq(got["pano"], want["pano"], "pano")
    assert ctx.kernel_launches > 0


def test_pipeline_device_resident(ctx, oracle):
    import torch
    O = oracle
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(3, 384, 288, 1.2, 0.25)
    want = O.pipeline_run(0, imgs, Ks, Rs, scale, seam=True, num_bands=5, weight_type=O.WEIGHT_32F)
    timgs = [torch.from_numpy(a).cuda() for a in imgs]
    torch.cuda.synchronize()
    got = S.Stitcher(ctx, 0, "dp", 5, S.WEIGHT_32F).stitch(timgs, Ks, Rs, scale)
    ctx.synchronize()
    _eq(got["pano"].cpu().numpy(), want["pano"], "device-resident pano")
    _eq(got["pano_mask"].cpu().numpy(), want["pano_mask"], "device-resident pano mask")


def test_seam_concurrent_pairs_are_proven_or_redone(ctx, oracle, monkeypatch):
    """The pair loop runs concurrently and is accepted only when proven equal to the sequential loop
    (is_ctx_seam_speculation); mutually overlapping images (rho = 0.6) exercise the dependent case."""
    O = oracle
    f = S.DpSeamFinder(ctx, "COLOR")
    corners, wi, wm = warped_set(O, 4, 256, 200, overlap=0.25)          # strip: independent pairs
    want = O.dp_seam_find(wi, corners, wm)
    got = f.find(wi, corners, [m.copy() for m in wm])
    assert ctx.seam_speculation == 1
    for k in range(4):
        _eq(got[k], want[k], f"concurrent mask {k}")
    monkeypatch.setenv("IS_SEAM_SEQUENTIAL", "1")
    got = f.find(wi, corners, [m.copy() for m in wm])
    assert ctx.seam_speculation == -1
    for k in range(4):
        _eq(got[k], want[k], f"sequential mask {k}")
    monkeypatch.delenv("IS_SEAM_SEQUENTIAL")
    corners, wi, wm = warped_set(O, 3, 300, 200, overlap=0.6)           # images 0 and 2 overlap as well
    want = O.dp_seam_find(wi, corners, wm)
    got = f.find(wi, corners, [m.copy() for m in wm])
    assert ctx.seam_speculation in (0, 1)
    for k in range(3):
        _eq(got[k], want[k], f"dependent mask {k}")


def test_blend_strips_equal_full_blend(ctx, oracle):
    """Column-strip blending (one strip per GPU): strips fed only with the images is_blender_strip_needs asks for
    must reproduce the same columns of the full blend, for odd / unaligned cuts too."""
    O = oracle
    corners, wi, wm = warped_set(O, 4, 256, 200, overlap=0.25)
    sm = O.dp_seam_find(wi, corners, wm)
    sizes = [(a.shape[1], a.shape[0]) for a in wi]
    roi = O.result_roi(corners, sizes)
    for wt in (S.WEIGHT_32F, S.WEIGHT_16S):
        ob = O.MultiBandBlender(5, wt)
        ob.prepare(roi)
        for i in range(4):
            ob.feed(wi[i].astype(np.int16), sm[i], corners[i])
        want, wmask = ob.blend()
        for cuts in ([0, 256, roi[2]], [0, 97, 301, 555, roi[2]], [0, 1, roi[2] - 1, roi[2]]):
            for x0, x1 in zip(cuts, cuts[1:]):
                gb = S.MultiBandBlender(ctx, 0, 5, wt)
                gb.prepare(roi)
                fed = [i for i in range(4) if gb.strip_needs(sizes[i], corners[i], x0, x1)]
                assert fed, "a strip inside the ROI needs at least one image"
                for i in fed:
                    gb.feed(wi[i], sm[i], corners[i])
                got, gmask = gb.blend_strip(x0, x1)
                _eq(gmask, wmask[:, x0:x1], f"strip mask [{x0},{x1})")
                _eq(got, want[:, x0:x1], f"strip [{x0},{x1}) fed {fed}")


def test_seam_pair_primitives(ctx, oracle):
    O = oracle
    corners, wi, wm = warped_set(O, 3, 320, 240, overlap=0.3)
    want = O.dp_seam_find(wi, corners, wm)
    f = S.DpSeamFinder(ctx, "COLOR")
    # reference order: (1,2) then (0,1); run both on the entry masks, then prove (0,1) on the mask it would have seen
    o1, o2, h12 = f.pair_run(wi[1], wi[2], corners[1], corners[2], wm[1], wm[2])
    p0, p1, h01 = f.pair_run(wi[0], wi[1], corners[0], corners[1], wm[0], wm[1])
    assert f.pair_check(wi[0], wi[1], corners[0], corners[1], wm[0], o1, h01)
    assert not f.pair_check(wi[0], wi[1], corners[0], corners[1], wm[0], np.zeros_like(wm[1]), h01)   # a different structure is detected
    m1 = wm[1].copy()
    f.mask_and(m1, o1)
    f.mask_and(m1, p1)
    _eq(p0, want[0], "mask 0")
    _eq(m1, want[1], "mask 1")
    _eq(o2, want[2], "mask 2")
    f.pair_free(h12)
    f.pair_free(h01)


def test_sharded_stitcher_single_rank(oracle):
    import torch

    from imagestitch_b200 import sharded
    O = oracle
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(4, 300, 220, 1.2, 0.25)
    want = O.pipeline_run(0, imgs, Ks, Rs, scale, seam=True, num_bands=4, weight_type=O.WEIGHT_32F, want_intermediates=True)
    be = sharded.GpuBackend(0)
    plan = sharded.ShardPlan.build(want["corners"], want["sizes"], want["roi"], 1, 4)
    st = sharded.ShardedStitcher(be, sharded.Comm(None), 4)
    res = st.stitch([torch.from_numpy(a).cuda() for a in imgs], Ks, Rs, scale, plan)
    assert st.info["seam_speculation"] == 1
    _eq(res["pano"].cpu().numpy(), want["pano"], "sharded (1 rank) pano")
    _eq(res["pano_mask"].cpu().numpy(), want["pano_mask"], "sharded (1 rank) pano mask")
    for i in range(4):
        _eq(res["seam_masks"][i].cpu().numpy(), want["masks"][i], f"sharded (1 rank) seam mask {i}")


def test_dp_seam_noisy_masks_take_the_dense_labelling_path(ctx, oracle, monkeypatch):
    """Salt-and-pepper masks have far too many runs for the run-based labelling: the dense union-find fallback must
    give the same masks, and forcing it (IS_SEAM_DENSE=1) on clean masks must change nothing either."""
    O = oracle
    rng = np.random.default_rng(4)
    corners, wi, wm = warped_set(O, 2, 160, 120, overlap=0.4)
    noisy = [np.where(rng.random(m.shape) < 0.35, 0, m).astype(np.uint8) for m in wm]
    want = O.dp_seam_find(wi, corners, noisy)
    got = S.DpSeamFinder(ctx, "COLOR").find(wi, corners, [m.copy() for m in noisy])
    for k in range(2):
        _eq(got[k], want[k], f"noisy mask {k}")
    corners, wi, wm = warped_set(O, 3, 256, 200, overlap=0.3)
    want = O.dp_seam_find(wi, corners, wm)
    monkeypatch.setenv("IS_SEAM_DENSE", "1")
    got = S.DpSeamFinder(ctx, "COLOR").find(wi, corners, [m.copy() for m in wm])
    for k in range(3):
        _eq(got[k], want[k], f"dense-path mask {k}")


def test_registration_hooks_two_phase(ctx, oracle):
    """detect -> match -> estimate as host hooks ([FEAT]:948, [MATCH]:123, [CAM]:118): is_pipeline_estimate hands back the cameras and
    the scale, is_pipeline_plan sizes the panorama from them, is_pipeline_run with those cameras equals the run with the hooks
    passed directly -- and the oracle.  An invalid image is rejected before any hook sees it."""
    import ctypes as C

    from imagestitch_b200 import capi
    O = oracle
    n = 3
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(n, 320, 240, 1.2, 0.3)
    calls = []

    def detect(user, i, mat):
        calls.append(("detect", i, mat.contents.rows, mat.contents.cols))
        return 0

    def match(user, k):
        calls.append(("match", k))
        return 0

    def estimate(user, k, cams, sc):
        calls.append(("estimate", k))
        for i in range(k):
            for q in range(9):
                cams[i].K[q] = float(np.asarray(Ks[i], np.float32).reshape(9)[q])
                cams[i].R[q] = float(np.asarray(Rs[i], np.float32).reshape(9)[q])
        sc[0] = float(scale)
        return 0

    hooks = capi.RegistrationHooks(None, capi.DETECT_FN(detect), capi.MATCH_FN(match), capi.ESTIMATE_FN(estimate))
    mats = (capi.Mat * n)()
    keep = []
    for i, a in enumerate(imgs):
        a = np.ascontiguousarray(a)
        keep.append(a)
        mats[i] = capi.Mat(a.ctypes.data, a.shape[0], a.shape[1], 3, 0, a.strides[0], -1)
    cams = (capi.Camera * n)()
    sc = C.c_float(0)
    ctx.check(ctx.lib.is_pipeline_estimate(ctx.h, n, mats, C.byref(hooks), cams, C.byref(sc)))
    assert calls == [("detect", i, 240, 320) for i in range(n)] + [("match", n), ("estimate", n)]
    assert sc.value == np.float32(scale)
    Ke = [np.asarray(list(cams[i].K), np.float32).reshape(3, 3) for i in range(n)]
    Re = [np.asarray(list(cams[i].R), np.float32).reshape(3, 3) for i in range(n)]
    st = S.Stitcher(ctx, "cylindrical", "dp", 3, S.WEIGHT_32F)
    got = st.stitch(imgs, Ke, Re, sc.value)
    want = O.pipeline_run(O.PROJ_CYLINDRICAL, imgs, Ks, Rs, scale, seam=True, num_bands=3, weight_type=O.WEIGHT_32F)
    _eq(got["pano"], want["pano"], "panorama from the estimated cameras")
    calls.clear()
    got2 = st.stitch(imgs, Ks, Rs, scale, hooks=C.byref(hooks))        # hooks passed to the run itself: same geometry, same panorama
    assert [c[0] for c in calls] == ["detect"] * n + ["match", "estimate"]
    _eq(got2["pano"], want["pano"], "panorama with the hooks in is_pipeline_run")
    calls.clear()
    bad = capi.Mat(keep[0].ctypes.data, 240, 320, 1, 0, keep[0].strides[0], -1)     # one channel: rejected before the hooks run
    mats[1] = bad
    assert ctx.lib.is_pipeline_estimate(ctx.h, n, mats, C.byref(hooks), cams, C.byref(sc)) < 0
    assert calls == []
