"""The oracle against the reference's OWN code: the hand-written pair blend ([BLEND]:141-717), the cylindrical
projector (detectResultRoi + mapBackward, [WARP]:47-88) and the refactored DP seam finder ([SEAM]:87-1093).

tests/golden/linblend_ref_cases.npz holds outputs of the reference's block compiled from /root/reference
(`make -C oracle ref`, generator tests/golden/make_reference_golden.py); the oracle's restatement must reproduce them bit
for bit -- panorama including its NaNs (0/0 weights where a row's left == seam + 1), greedy seam, cost map.  Where the
reference build is available (this container, or a prebuilt oracle/_ref on the GPU box) the same is checked live on
further cases."""
import os

import numpy as np
import pytest

from helpers import blob_masks, random_camera, seam_edge_cases, warped_set

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linblend_ref_cases.npz")


def _flat_trace(trace):
    """oracle trace [(i, j, comp, horizontal, points)] -> the flat layout of tests/golden/make_reference_golden._flat"""
    out = []
    for _i, _j, comp, horiz, pts in trace:
        out += [comp, int(horiz), len(pts)] + [int(v) for v in np.asarray(pts).reshape(-1)]
    return np.asarray(out, np.int32)


def _same(got, want, what):
    pano, seam, cost = got
    assert np.array_equal(seam, want[1]), f"{what}: greedy seam differs"
    assert np.array_equal(cost.view(np.uint32), want[2].view(np.uint32)), f"{what}: cost map differs"
    assert np.array_equal(np.isnan(pano), np.isnan(want[0])), f"{what}: NaN pattern differs"
    assert np.array_equal(np.nan_to_num(pano).view(np.uint32), np.nan_to_num(want[0]).view(np.uint32)), f"{what}: panorama differs"


def test_linear_blend_matches_reference_golden(oracle):
    O = oracle
    z = np.load(GOLD)
    assert int(z["n"]) >= 6
    for k in range(int(z["n"])):
        tl1, tl2 = (tuple(int(v) for v in t) for t in z[f"c{k}_tl"])
        got = O.lin_blend(z[f"c{k}_img1"].astype(np.float32), z[f"c{k}_img2"].astype(np.float32), tl1, tl2, want_cost=True)
        _same(got, (z[f"c{k}_pano_ref"], z[f"c{k}_seam_ref"], z[f"c{k}_cost_ref"]), f"case {k}")


def test_linear_blend_matches_reference_build_live(oracle):
    O = oracle
    if O.build_ref() is None:
        pytest.skip("oracle/_ref (the compiled reference block) is not available on this machine")
    for (w, h, ov) in ((400, 300, 0.25), (320, 260, 0.4), (257, 193, 0.33)):
        corners, wi, _ = warped_set(O, 2, w, h, overlap=ov)
        a, b = wi[0].astype(np.float32), wi[1].astype(np.float32)
        for tl2 in (corners[1], (corners[1][0], corners[0][1]), (corners[1][0], corners[0][1] + 6), (corners[1][0], corners[0][1] - 5)):
            want = O.ref_lin_blend(a, b, corners[0], tl2)
            got = O.lin_blend(a, b, corners[0], tl2, want_cost=True)
            assert (want is None) == (got is None)
            if want is not None:
                _same(got, want, f"{w}x{h} tl2={tuple(int(v) for v in tl2)}")
    assert O.ref_lin_blend(a, b, (0, 0), (5000, 0)) is None and O.lin_blend(a, b, (0, 0), (5000, 0)) is None     # [BLEND]:182-183


def test_cylindrical_maps_match_reference_golden(oracle):
    """ROI (the reference's full forward scan and the oracle's border scan) and backward maps, bit for bit"""
    O = oracle
    z = np.load(os.path.join(os.path.dirname(GOLD), "warp_ref_cases.npz"))
    for k in range(int(z["n"])):
        w, h = (int(v) for v in z[f"w{k}_size"])
        K, R, scale = z[f"w{k}_K"], z[f"w{k}_R"], float(z[f"w{k}_scale"])
        roi, xm, ym = O.build_maps(O.PROJ_CYLINDRICAL, (w, h), K, R, scale, full_scan=True)
        assert roi == tuple(int(v) for v in z[f"w{k}_roi_ref"])
        assert O.detect_roi(O.PROJ_CYLINDRICAL, (w, h), K, R, scale, full_scan=False) == roi
        assert np.array_equal(xm.view(np.uint32), z[f"w{k}_xmap_ref"].view(np.uint32))
        assert np.array_equal(ym.view(np.uint32), z[f"w{k}_ymap_ref"].view(np.uint32))


def test_cylindrical_maps_match_reference_build_live(oracle):
    O = oracle
    if O.build_ref() is None:
        pytest.skip("oracle/_ref (the compiled reference block) is not available on this machine")
    rng = np.random.default_rng(77)
    for _ in range(8):
        w, h = int(rng.integers(100, 500)), int(rng.integers(80, 400))
        K, R, scale = random_camera(rng, w, h)
        roi, xm, ym = O.ref_cylindrical_maps((w, h), K, R, scale)
        oroi, oxm, oym = O.build_maps(O.PROJ_CYLINDRICAL, (w, h), K, R, scale, full_scan=True)
        assert roi == oroi and O.detect_roi(O.PROJ_CYLINDRICAL, (w, h), K, R, scale, full_scan=False) == roi
        assert np.array_equal(xm.view(np.uint32), oxm.view(np.uint32)) and np.array_equal(ym.view(np.uint32), oym.view(np.uint32))


def test_dp_seam_matches_reference_golden(oracle):
    """seam masks written by the reference's own compiled find() for the inputs of seam_blend_cases.npz: COLOR (8-bit and
    float images give the same masks there) and COLOR_GRAD (float images, what the mains pass)"""
    O = oracle
    zin = np.load(os.path.join(os.path.dirname(GOLD), "seam_blend_cases.npz"))
    z = np.load(os.path.join(os.path.dirname(GOLD), "seam_ref_cases.npz"))
    for k in range(int(zin["n_cases"])):
        p = f"s{k}_"
        n = int(zin[p + "n"])
        corners = [tuple(int(v) for v in c) for c in zin[p + "corners"]]
        wi = [zin[p + f"img{i}"] for i in range(n)]
        wm = [zin[p + f"mask{i}"] for i in range(n)]
        for imgs in (wi, [a.astype(np.float32) for a in wi]):
            got, trace = O.dp_seam_find(imgs, corners, wm, want_trace=True)
            for i in range(n):
                assert np.array_equal(got[i], z[p + f"seam_mask{i}_ref"]), f"case {k} mask {i}"
            # every seam estimateSeam() produced in the reference: component, orientation, each point, in order
            assert np.array_equal(_flat_trace(trace), z[p + "seams_ref"]), f"case {k}: seam point lists differ"
        got, trace = O.dp_seam_find([a.astype(np.float32) for a in wi], corners, wm, cost_fn=O.COST_COLOR_GRAD, want_trace=True)
        assert np.array_equal(_flat_trace(trace), z[p + "seams_grad_ref"]), f"case {k}: seam point lists differ (COLOR_GRAD)"
        for i in range(n):
            assert np.array_equal(got[i], z[p + f"seam_mask{i}_grad_ref"]), f"case {k} mask {i} (COLOR_GRAD)"
            # the reference's copy and OpenCV's class agree with each other as well
            assert np.array_equal(z[p + f"seam_mask{i}_ref"], zin[p + f"seam_mask{i}_cv"])


@pytest.mark.parametrize("case", [(2, 260, 200, 0.25, 1, False), (3, 200, 150, 0.6, 1, False), (4, 160, 120, 0.3, 2, False),
                                  (3, 180, 130, 0.4, 1, True), (5, 220, 160, 0.35, 1, True), (6, 128, 96, 0.3, 2, True),
                                  (2, 1500, 1000, 0.3, 1, False)])
def test_dp_seam_matches_reference_build_live(oracle, case):
    O = oracle
    if O.build_ref() is None:
        pytest.skip("oracle/_ref (the compiled reference block) is not available on this machine")
    n, w, h, ov, rows, irregular = case
    corners, wi, wm = warped_set(O, n, w, h, overlap=ov, grid_rows=rows)
    if irregular:
        holes = blob_masks(np.random.default_rng(8), [m.shape for m in wm], holes=4)
        wm = [np.where(hm > 0, m, 0).astype(np.uint8) for m, hm in zip(wm, holes)]
    wf = [a.astype(np.float32) for a in wi]
    for imgs, cost in ((wi, O.COST_COLOR), (wf, O.COST_COLOR), (wf, O.COST_COLOR_GRAD)):
        want = O.ref_dp_seam_find(imgs, corners, wm, cost)
        want_seams = O.ref_last_seams()
        got, trace = O.dp_seam_find(imgs, corners, wm, cost_fn=cost, want_trace=True)
        for i in range(n):
            assert np.array_equal(got[i], want[i]), f"mask {i}, cost {cost}, {imgs[0].dtype}"
        assert len(trace) == len(want_seams) and len(trace) >= 1
        for a, b in zip(want_seams, trace):
            assert a[0] == b[2] and a[1] == b[3] and np.array_equal(a[2], b[4]), f"seam point list differs, cost {cost}, {imgs[0].dtype}"


def test_reference_find_on_its_own_artefacts(oracle):
    """images_warped_f[*].bmp -> mask_seam[*].bmp (tests/test_oracle_reference_artefacts.py): on those inputs the
    reference's compiled find() and the oracle give the same masks, so what separates both from the checked-in masks
    (854 of 1100 rows exact) is the reconstruction of the inputs, not the algorithm."""
    O = oracle
    cv2 = pytest.importorskip("cv2")
    d = "/root/reference/动态规划法寻找最佳缝合线/动态规划法寻找最佳缝合线/"
    if O.build_ref() is None or not os.path.isdir(d):
        pytest.skip("reference not present on this machine")
    from test_oracle_reference_artefacts import _footprint, _read
    imgs = [_read(f"images_warped_f[{i}].bmp", cv2.IMREAD_COLOR).astype(np.float32) for i in range(2)]
    masks = [_footprint(_read(f"mask_seam[{i}].bmp", cv2.IMREAD_GRAYSCALE)) for i in range(2)]
    want = O.ref_dp_seam_find(imgs, [(0, 5), (799, 0)], masks)
    got = O.dp_seam_find(imgs, [(0, 5), (799, 0)], masks)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_seam_cost_maps_match_reference_build_live(oracle):
    """computeCosts ([SEAM]:733-803): costV / costH of the oracle against the reference's own function, bit for bit, with
    fractional float images (sums that do round), a hole of another label in the component, COLOR and COLOR_GRAD"""
    O = oracle
    if O.build_ref() is None:
        pytest.skip("oracle/_ref (the compiled reference block) is not available on this machine")
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
    labels[iy0 + 5:iy0 + 9, ix0 + 3:ix0 + 10] = 1
    roi = (ix0, iy0, ix1 - ix0, iy1 - iy0)
    for dt in (np.uint8, np.float32):
        a = rng.integers(0, 256, (h1, w1, 3)).astype(dt)
        b = rng.integers(0, 256, (h2, w2, 3)).astype(dt)
        if dt == np.float32:
            a += rng.uniform(-0.5, 0.5, a.shape).astype(np.float32)
            b += rng.uniform(-0.5, 0.5, b.shape).astype(np.float32)
        for cost in ((O.COST_COLOR, O.COST_COLOR_GRAD) if dt == np.float32 else (O.COST_COLOR,)):
            wv, wh = O.ref_seam_costs(a, b, tl1, tl2, labels, utl, 2, roi, cost)
            gv, gh = O.seam_costs(a, b, tl1, tl2, labels, utl, 2, roi, cost)
            assert np.array_equal(gv.view(np.uint32), wv.view(np.uint32)), f"costV {dt.__name__} cost {cost}"
            assert np.array_equal(gh.view(np.uint32), wh.view(np.uint32)), f"costH {dt.__name__} cost {cost}"


def test_dp_seam_edge_cases_match_reference(oracle):
    """ties everywhere, containment, one-pixel overlaps, empty / gray / checkerboard masks, noise with irregular masks:
    against the masks the reference's own find() wrote for the same inputs (golden), and live where oracle/_ref exists"""
    O = oracle
    z = np.load(os.path.join(os.path.dirname(GOLD), "seam_ref_edge_cases.npz"))
    live = O.build_ref() is not None
    for k, (name, imgs, cs, ms, cost) in enumerate(seam_edge_cases()):
        got, trace = O.dp_seam_find(imgs, cs, ms, cost_fn=cost, want_trace=True)
        assert np.array_equal(_flat_trace(trace), z[f"e{k}_seams_ref"]), f"{name}: seam point lists differ from the reference's (golden)"
        want = O.ref_dp_seam_find(imgs, cs, ms, cost) if live else None
        for i in range(len(imgs)):
            assert np.array_equal(got[i], z[f"e{k}_mask{i}_ref"]), f"{name}: mask {i} differs from the reference's (golden)"
            if live:
                assert np.array_equal(got[i], want[i]), f"{name}: mask {i} differs from the reference's (live)"
        if cost == 0:                                          # 8-bit images: identical costs, identical masks ([SEAM]:742-743)
            got8 = O.dp_seam_find([a.astype(np.uint8) for a in imgs], cs, ms)
            assert all(np.array_equal(a, b) for a, b in zip(got8, got)), name


def _with_alpha(a, rng):
    """HxWx3 -> HxWx4 with a random fourth channel (it must not influence anything)"""
    return np.concatenate([a, rng.integers(0, 256, a.shape[:2] + (1,)).astype(a.dtype)], axis=2)


@pytest.mark.parametrize("case", [(2, 260, 200, 0.25, 1, False), (4, 160, 120, 0.3, 2, False), (3, 180, 130, 0.4, 1, True)])
def test_dp_seam_four_channels_match_reference_build(oracle, case):
    """CV_8UC4 / CV_32FC4 images ([SEAM]:722-730 diffL2Square4, 745-748): the oracle equals the reference's own find() on them --
    masks and seam point lists -- and the three-channel result (the fourth channel is skipped)."""
    O = oracle
    if O.build_ref() is None:
        pytest.skip("oracle/_ref (the compiled reference block) is not available on this machine")
    n, w, h, ov, rows, irregular = case
    corners, wi, wm = warped_set(O, n, w, h, overlap=ov, grid_rows=rows)
    if irregular:
        holes = blob_masks(np.random.default_rng(8), [m.shape for m in wm], holes=4)
        wm = [np.where(hm > 0, m, 0).astype(np.uint8) for m, hm in zip(wm, holes)]
    rng = np.random.default_rng(5)
    w4 = [_with_alpha(a, rng) for a in wi]
    f4 = [a.astype(np.float32) for a in w4]
    for imgs, cost in ((w4, O.COST_COLOR), (f4, O.COST_COLOR), (f4, O.COST_COLOR_GRAD)):
        want = O.ref_dp_seam_find(imgs, corners, wm, cost)
        want_seams = O.ref_last_seams()
        got, trace = O.dp_seam_find(imgs, corners, wm, cost_fn=cost, want_trace=True)
        three = O.dp_seam_find([np.ascontiguousarray(a[:, :, :3]) for a in imgs], corners, wm, cost_fn=cost)
        for i in range(n):
            assert np.array_equal(got[i], want[i]), f"mask {i}, cost {cost}, {imgs[0].dtype}"
            assert np.array_equal(got[i], three[i]), f"four channels != three channels, mask {i}"
        assert len(trace) == len(want_seams) and len(trace) >= 1
        for a, b in zip(want_seams, trace):
            assert a[0] == b[2] and a[1] == b[3] and np.array_equal(a[2], b[4])
