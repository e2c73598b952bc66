#!/usr/bin/env bash
# Round 2, GPU session 31: staged plan (components cut by two seams stay on the batched path): full suite, host/device validation cross-check,
# mosaic timing on one GPU.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s31_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s31_pytest_gpu.log 2>&1
echo "pytest gpu: exit $?" | tee gpurun_out/s31_status.txt
tail -4 gpurun_out/s31_pytest_gpu.log
IS_SEAM_CHECK_BOTH=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_seam_more.py -m gpu -q -x > gpurun_out/s31_pytest_both.log 2>&1
echo "pytest parity + seam, host and device validation cross-checked: exit $?" | tee -a gpurun_out/s31_status.txt
tail -3 gpurun_out/s31_pytest_both.log
python - <<'PY' 2>&1 | tee gpurun_out/s31_mosaic_timing.log
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from imagestitch_b200 import stitching as S, synth
n, rows, cols, grid = 12, 2000, 3000, 4
Ks, Rs, scale = synth.strip_cameras(n, cols, rows, 1.5, 0.25, grid_rows=grid)
imgs = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device="cuda:0") for i in range(n)]
for knob in ("1", ""):
    if knob: os.environ["IS_SEAM_NO_RESUME"] = knob
    else: os.environ.pop("IS_SEAM_NO_RESUME", None)
    ctx = S.Context(0, use_torch_stream=True)
    st = S.Stitcher(ctx, "cylindrical", "dp", 5, S.WEIGHT_32F)
    res = None
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = st.stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"4x3 mosaic of 12 x ({rows} x {cols}), staged plan {'off' if knob else 'on'}: {dt:.1f} ms per panorama, seam stage {st.timings_ms.get('seam'):.1f} ms, path {ctx.seam_path}, waves {ctx.seam_waves}")
    if knob: ref = [m.clone() for m in r["seam_masks"]], r["pano"].clone()
    else: print("same masks and panorama:", all(bool((a == b).all()) for a, b in zip(ref[0], r["seam_masks"])) and bool((ref[1] == r["pano"]).all()))
    ctx.close()
PY
