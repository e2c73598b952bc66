"""Profiling driver: one 2-image pair of the bench workload through the pipeline (used under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagestitch_b200 import stitching as S, synth

n, rows, cols = int(os.environ.get("N", 2)), int(os.environ.get("ROWS", 4000)), int(os.environ.get("COLS", 6000))
Ks, Rs, scale = synth.strip_cameras(n, cols, rows, 1.2, 0.25)
imgs = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device="cuda:0") for i in range(n)]
torch.cuda.synchronize()
ctx = S.Context(0, use_torch_stream=True)
st = S.Stitcher(ctx, "cylindrical", "dp", 5, S.WEIGHT_32F)
for it in range(int(os.environ.get("ITERS", 2))):
    r = st.stitch(imgs, Ks, Rs, scale)
    torch.cuda.synchronize()
    print(it, st.timings_ms)
if os.environ.get("KT"):      # per-kernel CUDA-event times of a few more steps
    ctx.kernel_timing(True); ctx.kernel_timing_report()
    k = int(os.environ.get("KT"))
    for it in range(k):
        st.stitch(imgs, Ks, Rs, scale)
    rep = sorted(ctx.kernel_timing_report(), key=lambda r: -r["ms"])
    for r in rep[:int(os.environ.get("KT_TOP", 12))]:
        print(f'{r["name"]:50s} {r["launches"] / k:6.1f} launches/step {r["ms"] / k:8.4f} ms/step')
