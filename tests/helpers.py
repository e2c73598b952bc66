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
