"""One configuration of the DP micro-benchmark (for ncu):  python scripts/dp_one.py LANES STEPS NJOBS [ITERS]   (environment knobs as in dp_bench.py)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagestitch_b200 import build as B, capi, stitching as S  # noqa: E402

B.build()
lib = capi.load()
ctx = S.Context(0)
lanes, steps, njobs = (int(v) for v in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
seam = np.zeros((njobs, steps), np.int32)
ms = (C.c_float * 1)()
rc = lib.is_debug_dp_bench(ctx.h, lanes, steps, njobs, 1, 1234, iters, seam.ctypes.data_as(C.POINTER(C.c_int32)), ms)
print("rc", rc, "ms", float(ms[0]))
ctx.close()
