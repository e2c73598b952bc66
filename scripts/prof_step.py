"""A few device-resident steps of the C2 pipeline (6 x (4000 x 6000), cylindrical warp, DP seam, 5-band blend) for ncu:
    ncu ... python scripts/prof_step.py [steps] [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from imagestitch_b200 import build as B, stitching as S, synth

B.build()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n, rows, cols, fw, ov, grid_rows, _ = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "c2"]
Ks, Rs, scale = synth.strip_cameras(n, cols, rows, fw, ov, grid_rows=grid_rows)
imgs = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device="cuda:0") for i in range(n)]
ctx = S.Context(0, use_torch_stream=True)
st = S.Stitcher(ctx, "cylindrical", "dp", bench.NUM_BANDS, S.WEIGHT_32F)
corners, sizes, roi = st.plan([(cols, rows)] * n, Ks, Rs, scale)
pano = torch.empty((roi[3], roi[2], 3), dtype=torch.int16, device="cuda:0")
pmask = torch.empty((roi[3], roi[2]), dtype=torch.uint8, device="cuda:0")
for _ in range(steps):
    ctx.clear_plan_cache()
    st.stitch(imgs, Ks, Rs, scale, out=(pano, pmask))
torch.cuda.synchronize()
print("done", ctx.kernel_launches)
