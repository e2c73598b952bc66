"""Shared input builders for the tests (inputs only; no reference arithmetic here)."""
import numpy as np

from imagestitch_b200 import synth


def rot(yaw, pitch, roll=0.0):
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    Rz = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
    return (Ry @ Rx @ Rz).astype(np.float32)


def random_camera(rng, w, h):
    f = float(rng.uniform(0.8, 2.0) * w)
    K = np.array([[f, 0, w / 2 + rng.uniform(-5, 5)], [0, f * rng.uniform(0.95, 1.05), h / 2 + rng.uniform(-5, 5)], [0, 0, 1]], np.float32)
    R = rot(rng.uniform(-1, 1), rng.uniform(-0.1, 0.1), rng.uniform(-0.05, 0.05))
    scale = float(f * rng.uniform(0.9, 1.1))
    return K, R, scale


def warped_set(O, n, w, h, f_over_w=1.2, overlap=0.25, proj=0, grid_rows=1):
    """Synthetic strip -> oracle-warped images, masks, corners (inputs of the seam / blend stages)."""
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(n, w, h, f_over_w, overlap, grid_rows=grid_rows)
    corners, wi, wm = [], [], []
    for i in range(n):
        tl, a = O.warp(proj, imgs[i], Ks[i], Rs[i], scale, O.INTER_LINEAR, O.BORDER_REFLECT, full_scan=False)
        _, m = O.warp(proj, np.full(imgs[i].shape[:2], 255, np.uint8), Ks[i], Rs[i], scale, O.INTER_NEAREST, O.BORDER_CONSTANT, full_scan=False)
        corners.append(tl)
        wi.append(a)
        wm.append(m)
    return corners, wi, wm


def blob_masks(rng, shapes, holes=3):
    """Irregular masks: full rectangles with a few rectangular holes / notches (multi-component cases)."""
    out = []
    for (h, w) in shapes:
        m = np.full((h, w), 255, np.uint8)
        for _ in range(holes):
            y0, x0 = int(rng.integers(0, h - 8)), int(rng.integers(0, w - 8))
            m[y0:y0 + int(rng.integers(3, max(4, h // 4))), x0:x0 + int(rng.integers(3, max(4, w // 4)))] = 0
        out.append(m)
    return out


def seam_edge_cases():
    """(name, images f32, corners, masks, cost_fn) -- degenerate and adversarial inputs of the DP seam finder: ties everywhere
    (flat images), containment, one-pixel overlaps, empty masks, touching / disjoint rectangles, many components, gray
    masks, noise images with irregular masks."""
    rng = np.random.default_rng(99)
    full = lambda h, w: np.full((h, w), 255, np.uint8)      # noqa: E731
    flat = np.zeros((40, 50, 3), np.float32)
    b = rng.integers(0, 256, (40, 50, 3)).astype(np.float32)
    c = rng.integers(0, 256, (20, 25, 3)).astype(np.float32)
    cb = ((np.indices((40, 50)).sum(0) // 4) % 2 * 255).astype(np.uint8)
    out = [("flat full overlap", [flat, flat], [(0, 0), (20, 7)], [full(40, 50), full(40, 50)], 0),
           ("flat same place", [flat, flat], [(3, 3), (3, 3)], [full(40, 50), full(40, 50)], 0),
           ("inside", [b, c], [(0, 0), (10, 8)], [full(40, 50), full(20, 25)], 0),
           ("inside reversed", [c, b], [(10, 8), (0, 0)], [full(20, 25), full(40, 50)], 0),
           ("1-col overlap", [b, b], [(0, 0), (49, 0)], [full(40, 50), full(40, 50)], 0),
           ("1-row overlap", [b, b], [(0, 0), (0, 39)], [full(40, 50), full(40, 50)], 0),
           ("2-col overlap", [b, b], [(0, 0), (48, 3)], [full(40, 50), full(40, 50)], 0),
           ("zero masks", [b, b], [(0, 0), (20, 5)], [np.zeros((40, 50), np.uint8), np.zeros((40, 50), np.uint8)], 0),
           ("one zero mask", [b, b], [(0, 0), (20, 5)], [full(40, 50), np.zeros((40, 50), np.uint8)], 0),
           ("disjoint", [b, b], [(0, 0), (100, 0)], [full(40, 50), full(40, 50)], 0),
           ("touching", [b, b], [(0, 0), (50, 0)], [full(40, 50), full(40, 50)], 0),
           ("checker masks", [b, b], [(0, 0), (20, 5)], [cb, cb.copy()], 0),
           ("gray masks", [b, b], [(0, 0), (20, 5)], [np.full((40, 50), 7, np.uint8), np.full((40, 50), 127, np.uint8)], 0)]
    for seed in range(6):
        r = np.random.default_rng(seed)
        n = int(r.integers(2, 5))
        imgs, ms, cs, x = [], [], [], 0
        for _ in range(n):
            h, w = int(r.integers(30, 70)), int(r.integers(40, 90))
            imgs.append(r.integers(0, 256, (h, w, 3)).astype(np.float32))
            ms.append(blob_masks(r, [(h, w)], holes=int(r.integers(0, 5)))[0])
            cs.append((x, int(r.integers(-8, 8))))
            x += int(w * r.uniform(0.3, 0.9))
        out.append((f"noise {seed} n={n}", imgs, cs, ms, 0))
        out.append((f"noise {seed} n={n} COLOR_GRAD", imgs, cs, ms, 1))
    return out
