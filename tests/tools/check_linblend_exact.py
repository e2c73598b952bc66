"""Is the device pair blend (is_linear_blend_pair) bit-exact against the oracle -- which is bit-exact against the reference's
own compiled block?  tests/test_gpu_parity.py only asserts |d| <= 1e-3 * 255 for the panorama; this reports the exact
figure so that the test can be tightened.   python tests/tools/check_linblend_exact.py   (needs a GPU)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
from helpers import warped_set  # noqa: E402
from imagestitch_b200 import build as B, stitching as S  # noqa: E402

O.build()
B.build()
ctx = S.Context(0)
bad = 0
for (w, h, ov) in ((400, 300, 0.25), (320, 260, 0.4), (1024, 768, 0.25)):
    corners, wi, _ = warped_set(O, 2, w, h, overlap=ov)
    a, b = wi[0].astype(np.float32), wi[1].astype(np.float32)
    for tl2 in (corners[1], (corners[1][0], corners[0][1]), (corners[1][0], corners[0][1] + 6), (corners[1][0], corners[0][1] - 5)):
        want = O.lin_blend(a, b, corners[0], tl2)
        got = S.linear_blend_pair(ctx, a, b, corners[0], tl2)
        if want is None or got is None:
            continue
        same = np.array_equal(np.nan_to_num(got[0]).view(np.uint32), np.nan_to_num(want[0]).view(np.uint32)) and \
            np.array_equal(np.isnan(got[0]), np.isnan(want[0]))
        d = np.nanmax(np.abs(got[0] - want[0]))
        print(f"{w}x{h} tl2={tuple(int(v) for v in tl2)}: seam {'ok' if np.array_equal(got[1], want[1]) else 'DIFFERS'}, panorama "
              f"{'bit-exact' if same else f'max |d| = {d:g}, {int((got[0] != want[0]).sum())} values differ'}", flush=True)
        bad += not same
ctx.close()
print("pair blend bit-exact:", "YES" if bad == 0 else f"NO ({bad} cases)")
