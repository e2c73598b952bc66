"""Single-GPU pipeline vs the CPU oracle at larger sizes; prints the seam speculation verdict."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle as O
from imagestitch_b200 import stitching as S, synth
O.set_threads(os.cpu_count())
ctx = S.Context(0)
for (n, rows, cols, fw) in [(6, 2000, 3000, 1.2), (12, 1000, 1500, 1.525), (12, 2000, 3000, 1.525)]:
    Ks, Rs, scale = synth.strip_cameras(n, cols, rows, fw, 0.25)
    imgs = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device="cuda:0").cpu().numpy() for i in range(n)]
    t = time.time()
    want = O.pipeline_run(0, imgs, Ks, Rs, scale, seam=True, num_bands=5, weight_type=O.WEIGHT_32F, want_intermediates=True)
    to = time.time() - t
    for mode in ("", "1"):
        os.environ["IS_SEAM_SEQUENTIAL"] = mode
        got = S.Stitcher(ctx, 0, "dp", 5, S.WEIGHT_32F).stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
        bad = [k for k in range(n) if not np.array_equal(got["seam_masks"][k], want["masks"][k])]
        print(f"n={n} {rows}x{cols} fw={fw} sequential_env={mode!r} speculation={ctx.seam_speculation} bad_masks={bad} "
              f"pano_equal={np.array_equal(got['pano'], want['pano'])} oracle={to:.1f}s", flush=True)
