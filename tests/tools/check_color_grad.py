"""COLOR_GRAD seam cost on the device against the CPU oracle (bit-exact seam masks and seam point lists).

The device path is switched on by IS_EXPERIMENTAL_COLOR_GRAD=1 (set here); tests/test_gpu_zz_reports.py runs this
script in a process of its own so that a fault cannot take the test session's CUDA context with it.
usage: python tests/tools/check_color_grad.py          exit code 0 = parity on every case
"""
import os
import sys

os.environ["IS_EXPERIMENTAL_COLOR_GRAD"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    import oracle as O
    from helpers import blob_masks, warped_set
    from imagestitch_b200 import build as B, stitching as S
    O.build()
    B.build()
    ctx = S.Context(0)
    bad = 0
    cases = [(2, 260, 200, 0.25, 1, False), (3, 200, 150, 0.6, 1, False), (4, 160, 120, 0.3, 2, False), (3, 180, 130, 0.4, 1, True),
             (5, 220, 160, 0.35, 1, True), (2, 1500, 1000, 0.3, 1, False)]
    for case in cases:
        n, w, h, ov, rows, irregular = case
        corners, wi, wm = warped_set(O, n, w, h, overlap=ov, grid_rows=rows)
        if irregular:
            holes = blob_masks(np.random.default_rng(8), [m.shape for m in wm], holes=4)
            wm = [np.where(hm > 0, m, 0).astype(np.uint8) for m, hm in zip(wm, holes)]
        for kind in ("u8", "f32"):
            imgs = wi if kind == "u8" else [a.astype(np.float32) for a in wi]
            want, wtrace = O.dp_seam_find(imgs, corners, wm, cost_fn=O.COST_COLOR_GRAD, want_trace=True)
            got, gtrace = S.DpSeamFinder(ctx, "COLOR_GRAD").find(imgs, corners, [m.copy() for m in wm], want_trace=True)
            ok = all(np.array_equal(a, b) for a, b in zip(got, want))
            ok_trace = len(wtrace) == len(gtrace) and all(a[:4] == b[:4] and np.array_equal(a[4], b[4]) for a, b in zip(wtrace, gtrace))
            color = O.dp_seam_find(imgs, corners, wm)
            print(f"case {case} {kind}: masks {'ok' if ok else 'DIFFER'}, seam points {'ok' if ok_trace else 'DIFFER'}, "
                  f"differs from COLOR in {sum(int((a != b).sum()) for a, b in zip(want, color))} px", flush=True)
            bad += (not ok) + (not ok_trace)
    # the whole sequence with the COLOR_GRAD cost (is_pipeline_run -> seam_find_device)
    from imagestitch_b200 import synth
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(3, 384, 288, 1.2, 0.25)
    want = O.pipeline_run(0, imgs, Ks, Rs, scale, seam=True, num_bands=5, weight_type=O.WEIGHT_32F, want_intermediates=True, seam_cost=O.COST_COLOR_GRAD)
    got = S.Stitcher(ctx, "cylindrical", "dp", 5, S.WEIGHT_32F, seam_cost="COLOR_GRAD").stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    okp = all(np.array_equal(a, b) for a, b in zip(got["seam_masks"], want["masks"])) and np.array_equal(got["pano"], want["pano"])
    print(f"pipeline with COLOR_GRAD: {'ok' if okp else 'DIFFERS'}", flush=True)
    bad += not okp
    ctx.close()
    print("COLOR_GRAD parity:", "PASS" if bad == 0 else f"FAIL ({bad})")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
