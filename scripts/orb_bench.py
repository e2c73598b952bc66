"""Times is_orb_find (the reference's find(), [FEAT]:948: 3 x 1 grid, 510 features per cell, 5 levels at 1.3) on one source image of
the C2 size beside cv2.ORB on the box's host cores, checks that both return the same key points and descriptors, and prints the
per-kernel CUDA-event times.  One JSON line; no oracle involved."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cv2
    import torch
    from imagestitch_b200 import stitching as S
    h, w = 4000, 6000
    rng = np.random.default_rng(3)
    img = np.zeros((h, w, 3), np.uint8)
    for s in (32, 8, 2):                                          # texture at several scales: corners on every pyramid level
        b = rng.integers(0, 86, (h // s + 1, w // s + 1, 3), dtype=np.uint8)
        img += np.kron(b, np.ones((s, s, 1), np.uint8))[:h, :w]
    ctx = S.Context(0)
    dev = torch.from_numpy(img).cuda()
    for _ in range(2):
        k, d = S.orb_find(ctx, dev)
    torch.cuda.synchronize()
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); S.orb_find(ctx, dev); t.append(time.perf_counter() - t0)
    th = []
    for _ in range(3):
        t0 = time.perf_counter(); S.orb_find(ctx, img); th.append(time.perf_counter() - t0)
    os.environ["IS_ORB_LAPS"] = "1"
    S.orb_find(ctx, dev)
    del os.environ["IS_ORB_LAPS"]
    ctx.kernel_timing(True)
    S.orb_find(ctx, dev)
    rep = ctx.kernel_timing_report()
    ctx.kernel_timing(False)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    orb = cv2.ORB_create(510, 1.3, 5)
    tc = []
    for _ in range(2):
        t0 = time.perf_counter()
        ck, cd = [], []
        g2 = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        for c in range(3):
            xl, xr = c * w // 3, (c + 1) * w // 3
            kk, dd = orb.detectAndCompute(np.ascontiguousarray(g2[:, xl:xr]), None)
            ck += [(p.pt[0] + xl, p.pt[1], p.size, p.angle, p.response, p.octave) for p in kk]
            cd.append(dd)
        tc.append(time.perf_counter() - t0)
    ck = np.array(ck, np.float32).reshape(-1, 6)
    cd = np.concatenate(cd)
    same = len(ck) == len(k) and np.array_equal(ck.view(np.uint32), k.view(np.uint32)) and np.array_equal(cd, d)
    print(json.dumps({"image": [h, w, 3], "keypoints": int(len(k)), "equals_cv2_orb": bool(same), "device_resident_ms": round(1e3 * min(t), 3),
                      "host_image_ms": round(1e3 * min(th), 3), "cv2_orb_ms": round(1e3 * min(tc), 2), "cv2_threads": cv2.getNumThreads(),
                      "kernels": [{"name": r["name"], "ms": round(r["ms"], 4), "launches": r["launches"]} for r in rep]}))


if __name__ == "__main__":
    main()
