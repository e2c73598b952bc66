"""Degenerate / adversarial DP-seam inputs (tests/helpers.seam_edge_cases: ties everywhere, containment, one-pixel overlaps,
empty / gray / checkerboard masks, noise with irregular masks) on the device against the oracle -- which equals the
reference's own find() on every one of them (tests/test_oracle_reference_build.py).  COLOR cases go through the shipped
path; COLOR_GRAD cases through the experimental switch.   python tests/tools/check_seam_edge_cases.py   (needs a GPU)"""
import os
import sys

os.environ["IS_EXPERIMENTAL_COLOR_GRAD"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
from helpers import seam_edge_cases  # noqa: E402
from imagestitch_b200 import build as B, stitching as S  # noqa: E402

O.build()
B.build()
ctx = S.Context(0)
bad_color = bad_grad = 0
for name, imgs, cs, ms, cost in seam_edge_cases():
    want = O.dp_seam_find(imgs, cs, ms, cost_fn=cost)
    variants = [("f32", imgs)] + ([("u8", [a.astype(np.uint8) for a in imgs])] if cost == 0 else [])
    for kind, im in variants:
        try:
            got = S.DpSeamFinder(ctx, "COLOR_GRAD" if cost else "COLOR").find(im, cs, [m.copy() for m in ms])
            ok = all(np.array_equal(a, b) for a, b in zip(got, want))
            note = "" if ok else " " + str([int((a != b).sum()) for a, b in zip(got, want)])
        except Exception as e:                                   # noqa: BLE001
            ok, note = False, f" raised {type(e).__name__}: {e}"
        print(f"{name} [{kind}]: {'ok' if ok else 'DIFFERS' + note}", flush=True)
        if not ok:
            if cost:
                bad_grad += 1
            else:
                bad_color += 1
ctx.close()
print(f"seam edge cases: COLOR {'PASS' if bad_color == 0 else f'FAIL ({bad_color})'}, COLOR_GRAD (experimental) {'PASS' if bad_grad == 0 else f'FAIL ({bad_grad})'}")
sys.exit(1 if bad_color else 0)
