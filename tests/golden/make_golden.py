"""Generates the golden fixtures under tests/golden/ from OpenCV (python cv2 4.13.0), the library whose
arithmetic the reference delegates to on this path (remap, DpSeamFinder, MultiBandBlender, pyrDown/pyrUp;
reference pin: OpenCV 3.4.2, un-vendored).  The reference itself ships no golden vectors.

    python tests/golden/make_golden.py

Inputs are generated with numpy / imagestitch_b200.synth only; every *_cv entry is an OpenCV output.
The oracle is NOT used here.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import blob_masks, random_camera  # noqa: E402
from imagestitch_b200 import synth  # noqa: E402

cv2.setNumThreads(1)
cv2.ocl.setUseOpenCL(False)


def warp_cases():
    rng = np.random.default_rng(2026)
    out = {}
    k = 0
    for name in ("cylindrical", "spherical"):
        for _ in range(3):
            w, h = int(rng.integers(60, 120)), int(rng.integers(50, 100))
            K, R, scale = random_camera(rng, w, h)
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            wp = cv2.PyRotationWarper(name, scale)
            roi, xm, ym = wp.buildMaps((w, h), K, R)
            tl, wi = wp.warp(img, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
            _, wm = wp.warp(np.full((h, w), 255, np.uint8), K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
            p = f"c{k}_"
            out.update({p + "proj": np.int32(0 if name == "cylindrical" else 1), p + "K": K, p + "R": R, p + "scale": np.float32(scale),
                        p + "img": img, p + "roi_cv": np.asarray(roi, np.int32), p + "tl_cv": np.asarray(tl, np.int32),
                        p + "xmap_cv": xm, p + "ymap_cv": ym, p + "warped_cv": wi, p + "mask_cv": wm})
            k += 1
    out["n"] = np.int32(k)
    np.savez_compressed(os.path.join(HERE, "warp_cases.npz"), **out)


def remap_cases():
    rng = np.random.default_rng(7)
    out = {}
    src = rng.integers(0, 256, (40, 50, 3), dtype=np.uint8)
    xm = rng.uniform(-120, 170, (48, 64)).astype(np.float32)
    ym = rng.uniform(-90, 130, (48, 64)).astype(np.float32)
    xm[0, :8] = -1
    ym[0, :8] = -1
    xm[1, :16] = np.arange(16, dtype=np.float32)          # exact integer coordinates: the (0,0) weight entry
    ym[1, :16] = np.arange(16, dtype=np.float32) + 2
    out.update(src=src, xmap=xm, ymap=ym)
    for iname, ci in (("linear", cv2.INTER_LINEAR), ("nearest", cv2.INTER_NEAREST)):
        for bname, cb in (("reflect", cv2.BORDER_REFLECT), ("constant", cv2.BORDER_CONSTANT)):
            out[f"{iname}_{bname}_c3_cv"] = cv2.remap(src, xm, ym, ci, borderMode=cb)
            out[f"{iname}_{bname}_c1_cv"] = cv2.remap(src[:, :, 0].copy(), xm, ym, ci, borderMode=cb)
    np.savez_compressed(os.path.join(HERE, "remap_cases.npz"), **out)


def _cv_warped_set(n, w, h, ov, grid_rows=1):
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(n, w, h, 1.2, ov, grid_rows=grid_rows)
    wp = cv2.PyRotationWarper("cylindrical", scale)
    corners, wi, wm = [], [], []
    for i in range(n):
        tl, a = wp.warp(imgs[i], Ks[i], Rs[i], cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, m = wp.warp(np.full(imgs[i].shape[:2], 255, np.uint8), Ks[i], Rs[i], cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        corners.append(tuple(tl))
        wi.append(a)
        wm.append(m)
    return corners, wi, wm


def _cv_seam_pairwise(wi, corners, masks):
    """cv2.detail_DpSeamFinder driven one pair at a time in the reference's order ([SEAM]:100-111)."""
    n = len(wi)
    masks = [m.copy() for m in masks]
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)][::-1]
    for (i, j) in pairs:
        sf = cv2.detail_DpSeamFinder("COLOR")
        res = sf.find([cv2.UMat(wi[i].astype(np.float32)), cv2.UMat(wi[j].astype(np.float32))], [corners[i], corners[j]],
                      [cv2.UMat(masks[i]), cv2.UMat(masks[j])])
        masks[i], masks[j] = res[0].get(), res[1].get()
    return masks


def seam_blend_cases():
    rng = np.random.default_rng(11)
    out = {}
    cases = [(2, 160, 120, 0.25, 1, False), (3, 128, 96, 0.35, 1, False), (3, 128, 96, 0.4, 1, True), (4, 112, 84, 0.3, 2, False)]
    for k, (n, w, h, ov, rows, irregular) in enumerate(cases):
        corners, wi, wm = _cv_warped_set(n, w, h, ov, rows)
        if irregular:
            holes = blob_masks(rng, [m.shape for m in wm], holes=4)
            wm = [np.where(hm > 0, m, 0).astype(np.uint8) for m, hm in zip(wm, holes)]
        sm = _cv_seam_pairwise(wi, corners, wm)
        p = f"s{k}_"
        out[p + "n"] = np.int32(n)
        out[p + "corners"] = np.asarray(corners, np.int32)
        for i in range(n):
            out[p + f"img{i}"] = wi[i]
            out[p + f"mask{i}"] = wm[i]
            out[p + f"seam_mask{i}_cv"] = sm[i]
        sizes = [(a.shape[1], a.shape[0]) for a in wi]
        tlx = min(c[0] for c in corners); tly = min(c[1] for c in corners)
        brx = max(c[0] + s[0] for c, s in zip(corners, sizes)); bry = max(c[1] + s[1] for c, s in zip(corners, sizes))
        roi = (tlx, tly, brx - tlx, bry - tly)
        for nb in (3, 5):
            for wname, cwt in (("f32", cv2.CV_32F), ("s16", cv2.CV_16S)):
                mb = cv2.detail_MultiBandBlender(0, nb, cwt)
                mb.prepare(roi)
                for i in range(n):
                    mb.feed(wi[i].astype(np.int16), sm[i], corners[i])
                d, dm = mb.blend(None, None)
                out[p + f"blend_nb{nb}_{wname}_cv"] = d
                out[p + f"blend_nb{nb}_{wname}_mask_cv"] = dm
                out[p + f"blend_nb{nb}_numbands_cv"] = np.int32(mb.numBands())
    out["n_cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(HERE, "seam_blend_cases.npz"), **out)


def pyr_cases():
    rng = np.random.default_rng(5)
    out = {}
    for k, (h, w) in enumerate(((17, 23), (32, 48), (5, 9), (41, 6))):
        a = rng.integers(-3000, 3000, (h, w, 3)).astype(np.int16)
        out[f"p{k}_src"] = a
        out[f"p{k}_down_cv"] = cv2.pyrDown(a)
        out[f"p{k}_up_cv"] = cv2.pyrUp(a, dstsize=(2 * w, 2 * h))
        out[f"p{k}_up_odd_cv"] = cv2.pyrUp(a, dstsize=(2 * w - 1, 2 * h - 1))
        f = rng.uniform(0, 1, (h, w)).astype(np.float32)
        out[f"p{k}_f32"] = f
        out[f"p{k}_f32_down_cv"] = cv2.pyrDown(f)
    out["n"] = np.int32(4)
    np.savez_compressed(os.path.join(HERE, "pyr_cases.npz"), **out)


def exposure_feather_cases():
    """GainCompensator feed / apply, dilate 20x20, distanceTransform(L1, 3), FeatherBlender (the mains' live blend path)."""
    rng = np.random.default_rng(77)
    out = {}
    for k in range(3):
        n = 3
        imgs, masks, corners, x = [], [], [], 0
        for i in range(n):
            h, w = int(rng.integers(50, 80)), int(rng.integers(70, 110))
            imgs.append(np.clip(rng.integers(0, 256, (h, w, 3)).astype(np.float32) * rng.uniform(0.6, 1.2), 0, 255).astype(np.uint8))
            m = blob_masks(rng, [(h, w)])[0]
            masks.append(m)
            corners.append((x, int(rng.integers(-6, 6))))
            x += int(w * rng.uniform(0.5, 0.8))
        c = cv2.detail.ExposureCompensator_createDefault(cv2.detail.ExposureCompensator_GAIN)
        c.feed(corners, imgs, masks)
        gains = np.array([float(g[0, 0]) for g in c.getMatGains()])
        sizes = [(a.shape[1], a.shape[0]) for a in imgs]
        x0 = min(cc[0] for cc in corners); y0 = min(cc[1] for cc in corners)
        x1 = max(cc[0] + sz[0] for cc, sz in zip(corners, sizes)); y1 = max(cc[1] + sz[1] for cc, sz in zip(corners, sizes))
        roi = (x0, y0, x1 - x0, y1 - y0)
        fb = cv2.detail_FeatherBlender(0.1)
        fb.prepare(roi)
        e = cv2.getStructuringElement(cv2.MORPH_RECT, (20, 20))
        out[f"e{k}_corners"] = np.asarray(corners, np.int32)
        out[f"e{k}_gains_cv"] = gains
        out[f"e{k}_roi"] = np.asarray(roi, np.int32)
        for i in range(n):
            out[f"e{k}_img{i}"] = imgs[i]
            out[f"e{k}_mask{i}"] = masks[i]
            out[f"e{k}_applied{i}_cv"] = c.apply(i, corners[i], imgs[i].copy(), masks[i])
            out[f"e{k}_dilated{i}_cv"] = cv2.dilate(masks[i], e)
            out[f"e{k}_dist{i}_cv"] = cv2.distanceTransform(masks[i], cv2.DIST_L1, 3)
            fb.feed(imgs[i].astype(np.int16), masks[i], corners[i])
        pano, pmask = fb.blend(None, None)
        out[f"e{k}_feather_cv"] = pano
        out[f"e{k}_feather_mask_cv"] = pmask
    out["n"] = np.int32(3)
    np.savez_compressed(os.path.join(HERE, "exposure_feather_cases.npz"), **out)


def color_grad_cases():
    """DpSeamFinder("COLOR_GRAD") masks for the inputs already stored in seam_blend_cases.npz, plus cvtColor + Sobel gradients."""
    z = np.load(os.path.join(HERE, "seam_blend_cases.npz"))
    out = {}
    for k in range(int(z["n_cases"])):
        p = f"s{k}_"
        n = int(z[p + "n"])
        corners = [tuple(int(v) for v in c) for c in z[p + "corners"]]
        wi = [z[p + f"img{i}"] for i in range(n)]
        masks = [z[p + f"mask{i}"].copy() for i in range(n)]
        for (i, j) in [(i, j) for i in range(n) for j in range(i + 1, n)][::-1]:
            res = cv2.detail_DpSeamFinder("COLOR_GRAD").find([cv2.UMat(wi[i].astype(np.float32)), cv2.UMat(wi[j].astype(np.float32))],
                                                         [corners[i], corners[j]], [cv2.UMat(masks[i]), cv2.UMat(masks[j])])
            masks[i], masks[j] = res[0].get(), res[1].get()
        for i in range(n):
            out[p + f"seam_mask{i}_grad_cv"] = masks[i]
    rng = np.random.default_rng(5)
    for k, shape in enumerate([(48, 64), (37, 53)]):
        a = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        g = cv2.cvtColor(a.astype(np.float32), cv2.COLOR_BGR2GRAY)
        out[f"g{k}_img"] = a
        out[f"g{k}_gradx_cv"] = cv2.Sobel(g, cv2.CV_32F, 1, 0)
        out[f"g{k}_grady_cv"] = cv2.Sobel(g, cv2.CV_32F, 0, 1)
    np.savez_compressed(os.path.join(HERE, "color_grad_cases.npz"), **out)


def mains_sequence_case():
    """The composite sequence of the reference's main() ([SEAM]:1150-1285) through OpenCV: warp -> GAIN feed -> apply in place
    -> DpSeamFinder(COLOR) on the compensated images -> dilate(20x20) & warped mask -> FeatherBlender(0.1).  The inputs are
    regenerated by the test from imagestitch_b200.synth (same call), only OpenCV's outputs are stored."""
    n = 3
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(n, 256, 192, 1.2, 0.25)
    gains_in = (0.7, 1.0, 1.25)
    imgs = [np.clip(a.astype(np.float32) * g, 0, 255).astype(np.uint8) for a, g in zip(imgs, gains_in)]
    warper = cv2.PyRotationWarper("cylindrical", scale)
    corners, wi, wm = [], [], []
    for i in range(n):
        tl, a = warper.warp(imgs[i], Ks[i], Rs[i], cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, m = warper.warp(np.full(imgs[i].shape[:2], 255, np.uint8), Ks[i], Rs[i], cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        corners.append(tuple(int(v) for v in tl)); wi.append(a); wm.append(m)
    comp = cv2.detail.ExposureCompensator_createDefault(cv2.detail.ExposureCompensator_GAIN)
    comp.feed(corners, wi, wm)
    wi = [comp.apply(i, corners[i], wi[i], wm[i]) for i in range(n)]
    masks = [m.copy() for m in wm]
    for (i, j) in [(i, j) for i in range(n) for j in range(i + 1, n)][::-1]:
        res = cv2.detail_DpSeamFinder("COLOR").find([cv2.UMat(wi[i].astype(np.float32)), cv2.UMat(wi[j].astype(np.float32))],
                                                    [corners[i], corners[j]], [cv2.UMat(masks[i]), cv2.UMat(masks[j])])
        masks[i], masks[j] = res[0].get(), res[1].get()
    el = cv2.getStructuringElement(cv2.MORPH_RECT, (20, 20))
    masks = [cv2.dilate(s, el) & m for s, m in zip(masks, wm)]
    sizes = [(a.shape[1], a.shape[0]) for a in wi]
    tl = (min(c[0] for c in corners), min(c[1] for c in corners))
    br = (max(c[0] + s[0] for c, s in zip(corners, sizes)), max(c[1] + s[1] for c, s in zip(corners, sizes)))
    fb = cv2.detail_FeatherBlender(0.1)
    fb.prepare((tl[0], tl[1], br[0] - tl[0], br[1] - tl[1]))
    for i in range(n):
        fb.feed(wi[i].astype(np.int16), masks[i], corners[i])
    pano, pmask = fb.blend(None, None)
    out = {"gains_in": np.asarray(gains_in), "pano_cv": pano, "pano_mask_cv": pmask}
    for i in range(n):
        out[f"seam_mask{i}_cv"] = masks[i]
    np.savez_compressed(os.path.join(HERE, "mains_sequence_case.npz"), **out)


def widened_cases():
    """the rows added last in round 2: plane / fisheye / stereographic projectors, cv::remap on extreme maps, the ORB
    features finder (cv2.ORB per grid cell, as find() [FEAT]:948 drives it), bitmap files (cv2.imwrite bytes, cv2.imread pixels)"""
    rng = np.random.default_rng(4242)
    out = {}
    k = 0
    for proj, name in ((2, "plane"), (3, "fisheye"), (4, "stereographic")):
        for _ in range(2):
            w, h = int(rng.integers(60, 110)), int(rng.integers(50, 90))
            K, R, scale = random_camera(rng, w, h)
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            wp = cv2.PyRotationWarper(name, scale)
            roi, xm, ym = wp.buildMaps((w, h), K, R)
            tl, wi = wp.warp(img, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
            _, wm = wp.warp(np.full((h, w), 255, np.uint8), K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
            p = f"w{k}_"
            out.update({p + "proj": np.int32(proj), p + "K": K, p + "R": R, p + "scale": np.float32(scale), p + "img": img,
                        p + "roi_cv": np.asarray(roi, np.int32), p + "xmap_cv": xm, p + "ymap_cv": ym, p + "warped_cv": wi, p + "mask_cv": wm})
            k += 1
    out["n_warp"] = np.int32(k)
    src = rng.integers(0, 256, (30, 40, 3), dtype=np.uint8)
    vals = np.array([np.nan, np.inf, -np.inf, 3e9, -3e9, 1e8, -1e8, 2 ** 31 / 32, -2 ** 31 / 32, 1e20, -1e20, 20.3, -0.5, 0.5, 39.5, 32767.4, 32768.6, -32768.5], np.float32)
    xm = np.tile(vals, (len(vals), 1))
    ym = xm.T.copy()
    out.update(x_src=src, x_xmap=xm, x_ymap=ym)
    for iname, ci in (("linear", cv2.INTER_LINEAR), ("nearest", cv2.INTER_NEAREST)):
        for bname, cb in (("reflect", cv2.BORDER_REFLECT), ("constant", cv2.BORDER_CONSTANT)):
            out[f"x_{iname}_{bname}_cv"] = cv2.remap(src, xm, ym, ci, borderMode=cb)
    orb = cv2.ORB_create(510, 1.3, 5)
    smooth = synth.make_panorama_inputs(2, 420, 260, 1.2, 0.25)[0][1]
    noise = rng.integers(0, 256, (150, 330), dtype=np.uint8)
    for j, (img, grid) in enumerate(((smooth, (3, 1)), (noise, (2, 1)))):
        gray = img if img.ndim == 2 else cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        h, w = gray.shape
        kps, ds = [], []
        for c in range(grid[0]):
            xl, xr = c * w // grid[0], (c + 1) * w // grid[0]
            kk, dd = orb.detectAndCompute(np.ascontiguousarray(gray[:, xl:xr]), None)
            kps += [(p.pt[0] + xl, p.pt[1], p.size, p.angle, p.response, p.octave) for p in kk]
            ds.append(dd)
        out.update({f"o{j}_img": img, f"o{j}_grid": np.asarray(grid, np.int32), f"o{j}_kps_cv": np.array(kps, np.float32).reshape(-1, 6),
                    f"o{j}_desc_cv": np.concatenate(ds)})
    out["n_orb"] = np.int32(2)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        f32 = rng.uniform(-40, 300, (7, 9, 3)).astype(np.float32)
        f32[0, 0] = (np.nan, np.inf, -3e9)
        for j, img in enumerate((rng.integers(0, 256, (6, 7, 3), dtype=np.uint8), rng.integers(0, 256, (5, 6), dtype=np.uint8),
                                 rng.integers(-300, 600, (4, 5, 3)).astype(np.int16), f32)):
            path = os.path.join(d, f"b{j}.bmp")
            assert cv2.imwrite(path, img)
            out.update({f"b{j}_img": img, f"b{j}_file_cv": np.frombuffer(open(path, "rb").read(), np.uint8), f"b{j}_read_cv": cv2.imread(path)})
    out["n_bmp"] = np.int32(4)
    np.savez_compressed(os.path.join(HERE, "widened_cases.npz"), **out)


if __name__ == "__main__":
    if "--only-widened" in sys.argv:
        widened_cases()
        sys.exit(0)
    if "--only-mains" in sys.argv:
        mains_sequence_case()
        sys.exit(0)
    mains_sequence_case()
    if "--only-new" not in sys.argv:
        warp_cases()
        remap_cases()
        seam_blend_cases()
        pyr_cases()
    if "--only-color-grad" not in sys.argv:
        exposure_feather_cases()
    color_grad_cases()
    widened_cases()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
    print("cv2", cv2.__version__)
