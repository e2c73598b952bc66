"""One 4 x 3 mosaic through the pipeline with IS_SEAM_DEBUG=1: which pairs leave the batched seam path and why.   python scripts/mosaic_debug.py [rows cols]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IS_SEAM_DEBUG"] = "1"
import torch

from imagestitch_b200 import build as B, stitching as S, synth

B.build()
rows, cols = (int(v) for v in (sys.argv[1:3] + ["2000", "3000"])[:2])
n, grid = 12, 4
Ks, Rs, scale = synth.strip_cameras(n, cols, rows, 1.5, 0.25, grid_rows=grid)
imgs = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device="cuda:0") for i in range(n)]
ctx = S.Context(0, use_torch_stream=True)
st = S.Stitcher(ctx, "cylindrical", "dp", 5, S.WEIGHT_32F)
st.stitch(imgs, Ks, Rs, scale)
torch.cuda.synchronize()
print("path", ctx.seam_path, "waves", ctx.seam_waves)
