"""The oracle against the reference's OWN code: the hand-written pair blend ([BLEND]:141-717) and the cylindrical
projector (detectResultRoi + mapBackward, [WARP]:47-88).

tests/golden/linblend_ref_cases.npz holds outputs of the reference's block compiled from /root/reference
(`make -C oracle ref`, generator tests/golden/make_reference_golden.py); the oracle's restatement must reproduce them bit
for bit -- panorama including its NaNs (0/0 weights where a row's left == seam + 1), greedy seam, cost map.  Where the
reference build is available (this container, or a prebuilt oracle/_ref on the GPU box) the same is checked live on
further cases."""
import os

import numpy as np
import pytest

from helpers import random_camera, warped_set

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linblend_ref_cases.npz")


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
