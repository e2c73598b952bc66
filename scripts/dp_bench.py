"""DP forward / back-track micro-benchmark on the device (is_debug_dp_bench): both formulations on synthetic cost tables of the
shapes the seam stage meets, seams compared for equality.   python scripts/dp_bench.py   (needs a GPU)"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagestitch_b200 import build as B, capi, stitching as S  # noqa: E402

B.build()
lib = capi.load()
ctx = S.Context(0)


def run(lanes, steps, njobs, variant, iters=5, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = str(v)
    seam = np.zeros((njobs, steps), np.int32)
    ms = (C.c_float * 1)()
    rc = lib.is_debug_dp_bench(ctx.h, lanes, steps, njobs, variant, 1234, iters, seam.ctypes.data_as(C.POINTER(C.c_int32)), ms)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    if rc != 0:
        return None, None
    return float(ms[0]), seam


out = []
for (lanes, steps, njobs) in ((1500, 4029, 1), (1500, 4029, 5), (1500, 4029, 11), (2000, 6029, 5), (3000, 8029, 8), (4029, 1500, 5), (6029, 2000, 4), (300, 12000, 5), (700, 2000, 3), (100, 700, 2)):
    base_ms, base = run(lanes, steps, njobs, 0)
    row = {"lanes": lanes, "steps": steps, "njobs": njobs, "v0_ms": base_ms}
    variants = [("v1", {}), ("v1_barrier", {"IS_DP_CL_BARRIER": 1}), ("v1_cl1", {"IS_DP_CLUSTER": 1}), ("v1_cl1_r4", {"IS_DP_CLUSTER": 1, "IS_DP_CL_RING": 4})]
    for cl in (2, 4, 8):
        variants.append((f"v1_cl{cl}", {"IS_DP_CLUSTER": cl}))
        variants.append((f"v1_cl{cl}_barrier", {"IS_DP_CLUSTER": cl, "IS_DP_CL_BARRIER": 1}))
    for name, env in variants:
        ms, seam = run(lanes, steps, njobs, 1, env=env)
        row[name + "_ms"] = ms
        row[name + "_equal"] = None if seam is None or base is None else bool(np.array_equal(seam, base))
    row["reached"] = None if base is None else [bool(r[0] >= 0) for r in base]
    print(json.dumps(row), flush=True)
    out.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dp_bench.json"), "w"), indent=1)
ctx.close()
