"""BASELINE.json configs[0]: 2x(1024x768) RGB pair, cylindrical warp + the reference's hand-written linear blend, CPU, one core.
The blend is timed twice: the reference's OWN block ([BLEND]:141-717 compiled into oracle/_ref, when available) and the
oracle's restatement of it; the warp is the oracle's (the reference calls OpenCV's warper for it).  Prints one JSON line.
    python tests/tools/c1_reference_timing.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
from imagestitch_b200 import synth  # noqa: E402

O.build()
O.set_threads(1)
rows, cols = 768, 1024
imgs, Ks, Rs, scale = synth.make_panorama_inputs(2, cols, rows, 1.2, 0.25)
t0 = time.perf_counter()
warped, corners = [], []
for i in range(2):
    tl, a = O.warp(O.PROJ_CYLINDRICAL, imgs[i], Ks[i], Rs[i], scale, O.INTER_LINEAR, O.BORDER_REFLECT, full_scan=True)
    warped.append(a.astype(np.float32))
    corners.append(tl)
t_warp = time.perf_counter() - t0


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t)
    return min(ts), r


t_port, port = best(lambda: O.lin_blend(warped[0], warped[1], corners[0], corners[1]))
out = {"config": "2x(768x1024) RGB pair, cylindrical warp + reference linear blend, CPU single core", "warp_seconds_full_roi_scan": t_warp,
       "linear_blend_port_seconds": t_port, "input_mp": 2 * rows * cols / 1e6}
if O.build_ref() is not None:
    t_ref, ref = best(lambda: O.ref_lin_blend(warped[0], warped[1], corners[0], corners[1]))
    out["linear_blend_reference_seconds"] = t_ref
    out["reference_equals_port"] = bool(np.array_equal(np.nan_to_num(ref[0]), np.nan_to_num(port[0])) and np.array_equal(ref[1], port[1]))
    out["mp_per_s_warp_plus_reference_blend"] = out["input_mp"] / (t_warp + t_ref)
print(json.dumps(out))
