"""Golden vectors produced by the REFERENCE'S OWN code, compiled from /root/reference against oracle/ref_shim/cvshim.h
(`make -C oracle ref`): the hand-written pair blend ([BLEND]:141-717) -> linblend_ref_cases.npz, the cylindrical
projector (detectResultRoi + mapBackward, [WARP]:47-88) -> warp_ref_cases.npz, and the refactored DP seam finder
(find() ... updateLabelsUsingSeam, [SEAM]:87-1093) on the inputs of seam_blend_cases.npz -> seam_ref_cases.npz.  Runs only where /root/reference exists; the
resulting files travel, the reference does not.
    python tests/golden/make_reference_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle as O  # noqa: E402
from helpers import random_camera, seam_edge_cases, warped_set  # noqa: E402


def _flat(seams):
    """[(comp, horizontal, points Nx2)] -> int32 vector comp, horizontal, npts, x0, y0, ... (the layout of the oracle's trace
    without the pair indices)"""
    out = []
    for comp, horiz, pts in seams:
        out += [comp, int(horiz), len(pts)] + [int(v) for v in pts.reshape(-1)]
    return np.asarray(out, np.int32)


def cases():
    """(img1 u8, img2 u8, tl1, tl2): warped synthetic pairs with their black borders (dy < 0, == 0, > 0) and noise
    pairs with dark patches around the classification thresholds 10 / 20 ([BLEND]:331-460)."""
    out = []
    corners, wi, _ = warped_set(O, 2, 120, 90, overlap=0.3)
    for dy in (None, 0, 4, -3):
        tl2 = corners[1] if dy is None else (corners[1][0], corners[0][1] + dy)
        out.append((wi[0], wi[1], tuple(int(v) for v in corners[0]), tuple(int(v) for v in tl2)))
    # (image heights are chosen so that the block never reads image rows past the end of an image: for panoHe - dy2 > rows
    #  [BLEND]:216-219 does, and what it finds there is whatever follows the buffer)
    rng = np.random.default_rng(41)
    for k, (h1, w1, h2, w2, tl2) in enumerate([(70, 100, 70, 96, (60, 0)), (66, 90, 64, 90, (41, 2)), (72, 88, 70, 92, (50, -2))]):
        a = rng.integers(0, 256, (h1, w1, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (h2, w2, 3), dtype=np.uint8)
        for im in (a, b):                                   # dark patches: values 0..40 straddle both thresholds
            for _ in range(6):
                y, x = int(rng.integers(0, im.shape[0] - 12)), int(rng.integers(0, im.shape[1] - 12))
                im[y:y + 12, x:x + 12] = rng.integers(0, 41, (12, 12, 3), dtype=np.uint8)
        a[:, :3] = 0
        b[:, -3:] = 0
        out.append((a, b, (0, 0), tl2))
    return out


if __name__ == "__main__":
    if O.build_ref() is None:
        sys.exit("oracle/_ref is not available here")
    z = {}
    cs = cases()
    for k, (a, b, tl1, tl2) in enumerate(cs):
        pano, seam, cost = O.ref_lin_blend(a.astype(np.float32), b.astype(np.float32), tl1, tl2)
        z[f"c{k}_img1"], z[f"c{k}_img2"] = a, b
        z[f"c{k}_tl"] = np.asarray([tl1, tl2], np.int32)
        z[f"c{k}_pano_ref"], z[f"c{k}_seam_ref"], z[f"c{k}_cost_ref"] = pano, seam, cost
    z["n"] = np.int32(len(cs))
    path = os.path.join(HERE, "linblend_ref_cases.npz")
    np.savez_compressed(path, **z)
    print(path, os.path.getsize(path), "bytes,", len(cs), "cases")
    rng = np.random.default_rng(2024)
    z = {}
    for k in range(6):
        w, h = int(rng.integers(60, 140)), int(rng.integers(50, 120))
        K, R, scale = random_camera(rng, w, h)
        roi, xm, ym = O.ref_cylindrical_maps((w, h), K, R, scale)
        z[f"w{k}_size"] = np.asarray([w, h], np.int32)
        z[f"w{k}_K"], z[f"w{k}_R"], z[f"w{k}_scale"] = np.asarray(K, np.float32), np.asarray(R, np.float32), np.float32(scale)
        z[f"w{k}_roi_ref"], z[f"w{k}_xmap_ref"], z[f"w{k}_ymap_ref"] = np.asarray(roi, np.int32), xm, ym
    z["n"] = np.int32(6)
    path = os.path.join(HERE, "warp_ref_cases.npz")
    np.savez_compressed(path, **z)
    print(path, os.path.getsize(path), "bytes")
    zin = np.load(os.path.join(HERE, "seam_blend_cases.npz"))
    z = {}
    for k in range(int(zin["n_cases"])):
        p = f"s{k}_"
        n = int(zin[p + "n"])
        corners = [tuple(int(v) for v in c) for c in zin[p + "corners"]]
        wi = [zin[p + f"img{i}"] for i in range(n)]
        wm = [zin[p + f"mask{i}"] for i in range(n)]
        color = O.ref_dp_seam_find([a.astype(np.float32) for a in wi], corners, wm, O.COST_COLOR)
        color_u8 = O.ref_dp_seam_find(wi, corners, wm, O.COST_COLOR)
        assert all(np.array_equal(a, b) for a, b in zip(color, color_u8))
        color = O.ref_dp_seam_find([a.astype(np.float32) for a in wi], corners, wm, O.COST_COLOR)
        z[p + "seams_ref"] = _flat(O.ref_last_seams())
        grad = O.ref_dp_seam_find([a.astype(np.float32) for a in wi], corners, wm, O.COST_COLOR_GRAD)
        z[p + "seams_grad_ref"] = _flat(O.ref_last_seams())
        for i in range(n):
            z[p + f"seam_mask{i}_ref"] = color[i]
            z[p + f"seam_mask{i}_grad_ref"] = grad[i]
    path = os.path.join(HERE, "seam_ref_cases.npz")
    np.savez_compressed(path, **z)
    print(path, os.path.getsize(path), "bytes")
    z = {}
    for k, (name, imgs, cs, ms, cost) in enumerate(seam_edge_cases()):
        ref = O.ref_dp_seam_find(imgs, cs, ms, cost)
        z[f"e{k}_seams_ref"] = _flat(O.ref_last_seams())
        for i, m in enumerate(ref):
            z[f"e{k}_mask{i}_ref"] = m
    path = os.path.join(HERE, "seam_ref_edge_cases.npz")       # inputs are regenerated by tests/helpers.seam_edge_cases()
    np.savez_compressed(path, **z)
    print(path, os.path.getsize(path), "bytes")

