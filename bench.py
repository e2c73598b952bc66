#!/usr/bin/env python
"""bench.py -- stitched megapixels/s of the composite path (warp -> DP seam -> 5-band blend).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|tiny]

One "step" = one pass of the whole hot path over one synthetic panorama (BASELINE.json configs[1] at N=1:
6 x (4000 x 6000) RGB strip, cylindrical warp, DP seam masks, multi-band blend with 5 bands).

  value     input megapixels / s with the sources already resident in HBM when the timed region starts and
            the panorama left in HBM (CUDA events on the stream the kernels run on, max over ranks)
  e2e       the same metric through the C ABI with HOST buffers (pinned): H2D of the sources and D2H of the
            panorama + mask inside the timed region
  roofline  the dominant kernel of the step: algorithmic bytes per launch / CUDA-event duration per launch
            (per-launch events recorded by the library in a second timed region of the same K steps)
  cpu_baseline   the oracle's CPU restatement of the same path on the host cores, bounded sample

--impl reference times the CPU port (oracle/, all host threads) on the same workload definition; the
reference's own sources cannot be compiled here (OpenCV C++ headers/libs absent), see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_images per GPU, rows, cols, f_over_w, overlap, grid_rows, description)
    "c2": (6, 4000, 6000, 1.2, 0.25, 1, "6x(4000x6000) RGB strip, cylindrical warp + DP seam masks + multi-band blend (5 bands)"),
    "c3": (12, 4000, 6000, 1.5, 0.25, 1, "12x(4000x6000) RGB strip, cylindrical warp + DP seam + multi-band blend (5 bands)"),
    "c2_8k": (6, 6000, 8000, 1.2, 0.25, 1, "6x(6000x8000) RGB strip (8K-wide images), cylindrical warp + DP seam masks + multi-band blend (5 bands)"),
    "c1": (2, 768, 1024, 1.2, 0.25, 1, "2x(768x1024) RGB pair, cylindrical warp + hand-written linear blend ([BLEND]:141-717)"),
    # BASELINE.json configs[3], configs[4]: per-GPU shares of the 8-GPU panoramas (run with --gpus 8)
    "c4": (3, 6000, 8000, 3.0, 0.25, 1, "24x(6000x8000) RGB 346-degree strip over 8 GPUs (3 images per GPU), column-strip sharded, NCCL halo exchange"),
    "c5": (6, 8000, 12000, 1.5, 0.25, 4, "48x(8000x12000) RGB 4x12 mosaic over 8 GPUs (6 images per GPU), column-strip sharded, NCCL halo exchange"),
    "tiny": (3, 384, 512, 1.2, 0.25, 1, "3x(384x512) debug strip"),
}
NUM_BANDS = 5
L2_BYTES = 126 * 1024 * 1024


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()            # the exact process we started
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass

    def summary(self, windows):
        sm, smax, reasons = [], [], set()
        for t, line in self.rows:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(n, rows, cols, sizes, roi):
    """compulsory traffic of the path, implementation independent."""
    src = n * 3 * rows * cols
    area = sum(w * h for (w, h) in sizes)
    pano = roi[2] * roi[3]
    return {"warp": src + 4 * area,                    # u8 source read, u8x3 warped + u8 mask written
            "warp_blend_fused": src + area + 7 * pano,  # B_warp+blend: source + seam mask read, s16x3 pano + u8 mask written
            "blend_level0": 4 * area + 7 * pano,       # warped u8x3 + mask read once, pano + mask written once
            "pyrdown_level0": 4 * area + (6 + 4) * area // 4}


def cpu_port_run(workload, threads, sample_rows=None, n_sample=2, steps=1, warmup=0):
    """The oracle's CPU restatement of the path on a bounded sample: n_sample neighbouring images (full
    width, sample_rows rows).  Returns (MP/s, description, seconds per step, per-stage seconds)."""
    import numpy as np

    import oracle as O
    from imagestitch_b200 import synth
    n, rows, cols, fw, ov, grid_rows, _ = WORKLOADS[workload]
    rows_s = min(rows, sample_rows or rows)
    n_s = min(n, n_sample)
    Ks, Rs, scale = synth.strip_cameras(n, cols, rows_s, fw * 1.0, ov, grid_rows=grid_rows)
    # keep the focal length of the full workload: f = f_over_w * cols
    imgs = [synth.make_image(i, cols, rows_s, Ks[i], Rs[i], device="cpu").numpy() for i in range(n_s)]
    O.set_threads(threads)
    best = None
    stages = None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = O.pipeline_run(O.PROJ_CYLINDRICAL, imgs, Ks[:n_s], Rs[:n_s], scale, seam=True, num_bands=NUM_BANDS, weight_type=O.WEIGHT_32F)
        dt = time.perf_counter() - t0
        if it >= warmup and (best is None or dt < best):
            best, stages = dt, [float(v) for v in r["seconds"]]
    mp = n_s * rows_s * cols / 1e6
    desc = f"{n_s} neighbouring images of the workload, {rows_s}x{cols} each (warp + DP seam + {NUM_BANDS}-band blend), {threads} OpenMP threads, best of {steps}"
    return mp / best, desc, best, stages


def _informational_leg(flag, workload, sample_rows, timeout=600):
    """An informational CPU leg in a process of its own: nothing it does (cv2's threading runtime, the reference's code reading
    past a buffer) can disturb or end the bench.  Returns its dict, or {"unavailable": why}."""
    import subprocess
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), flag, str(int(sample_rows)), "--workload", workload],
                           capture_output=True, text=True, timeout=timeout)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": f"no output (exit {r.returncode}): {r.stderr.strip()[-160:]}"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def opencv_run(workload, sample_rows):
    return _informational_leg("--opencv-sample", workload, sample_rows)


def reference_find_check(workload, sample_rows):
    return _informational_leg("--reference-find-sample", workload, sample_rows)


def weights16s_leg_inline(workload, steps):
    """The device-resident step of `workload` on one GPU with MultiBandBlender's CV_16S weights ("report CV_16S too")."""
    import torch

    from imagestitch_b200 import stitching as S, synth
    n, rows, cols, fw, ov, grid_rows, _desc = WORKLOADS[workload]
    Ks, Rs, scale = synth.strip_cameras(n, cols, rows, fw, ov, grid_rows=grid_rows)
    dev = "cuda:0"
    torch.cuda.set_device(0)
    ctx = S.Context(0, use_torch_stream=True)
    st = S.Stitcher(ctx, "cylindrical", "dp", NUM_BANDS, S.WEIGHT_16S)
    imgs = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device=dev) for i in range(n)]
    _corners, _sizes, roi = st.plan([(cols, rows)] * n, Ks, Rs, scale)
    pano = torch.empty((roi[3], roi[2], 3), dtype=torch.int16, device=dev)
    pmask = torch.empty((roi[3], roi[2]), dtype=torch.uint8, device=dev)

    def step():
        ctx.clear_plan_cache()
        st.stitch(imgs, Ks, Rs, scale, out=(pano, pmask))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "value": n * rows * cols / 1e6 / (ms * 1e-3), "unit": "MP/s", "steps": steps, "weight_type": "CV_16S"}


def opencv_run_inline(workload, sample_rows, n_sample=2):
    """Informational: the reference-equivalent OpenCV path through python cv2 on the host cores, same bounded
    sample as cpu_port_run: PyRotationWarper.warp x2 per image -> float32 -> detail_DpSeamFinder -> int16 ->
    MultiBandBlender(5 bands).  Returns None when cv2 (or its stitching module) is not importable."""
    try:
        import cv2
        import numpy as np

        from imagestitch_b200 import synth
        n, rows, cols, fw, ov, grid_rows, _ = WORKLOADS[workload]
        rows_s = min(rows, sample_rows or rows)
        n_s = min(n, n_sample)
        Ks, Rs, scale = synth.strip_cameras(n, cols, rows_s, fw * 1.0, ov, grid_rows=grid_rows)
        imgs = [synth.make_image(i, cols, rows_s, Ks[i], Rs[i], device="cpu").numpy() for i in range(n_s)]
        t0 = time.perf_counter()
        wp = cv2.PyRotationWarper("cylindrical", float(scale))
        corners, wi, wm = [], [], []
        for i in range(n_s):
            K, R = np.asarray(Ks[i], np.float32), np.asarray(Rs[i], np.float32)
            tl, a = wp.warp(imgs[i], K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
            _, m = wp.warp(np.full(imgs[i].shape[:2], 255, np.uint8), K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
            corners.append(tuple(int(v) for v in tl)); wi.append(a); wm.append(m)
        t1 = time.perf_counter()
        masks = cv2.detail_DpSeamFinder("COLOR").find([cv2.UMat(a.astype(np.float32)) for a in wi], corners, [cv2.UMat(m) for m in wm])
        masks = [m.get() for m in masks]
        t2 = time.perf_counter()
        x0 = min(c[0] for c in corners); y0 = min(c[1] for c in corners)
        x1 = max(c[0] + a.shape[1] for c, a in zip(corners, wi)); y1 = max(c[1] + a.shape[0] for c, a in zip(corners, wi))
        mb = cv2.detail_MultiBandBlender(0, NUM_BANDS, cv2.CV_32F)
        mb.prepare((x0, y0, x1 - x0, y1 - y0))
        for i in range(n_s):
            mb.feed(wi[i].astype(np.int16), masks[i], corners[i])
        mb.blend(None, None)
        t3 = time.perf_counter()
        return {"value": n_s * rows_s * cols / 1e6 / (t3 - t0), "unit": "MP/s", "threads": int(cv2.getNumThreads()), "version": cv2.__version__,
                "stage_seconds": {"warp": t1 - t0, "seam": t2 - t1, "blend": t3 - t2, "total": t3 - t0}}
    except Exception as e:                      # informational only: never fail the bench over it
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def reference_find_check_inline(workload, sample_rows, n_sample=2):
    """Informational: the reference's OWN find() ([SEAM]:87-1093 compiled into oracle/_ref, single-threaded as in the reference)
    beside the port's seam finder on the CPU baseline's sample, and whether their masks agree.  None when oracle/_ref is absent."""
    try:
        import numpy as np

        import oracle as O
        from imagestitch_b200 import synth
        if O.build_ref() is None:
            return {"unavailable": "oracle/_ref not present"}
        n, rows, cols, fw, ov, grid_rows, _ = WORKLOADS[workload]
        rows_s = min(rows, sample_rows or rows)
        n_s = min(n, n_sample)
        Ks, Rs, scale = synth.strip_cameras(n, cols, rows_s, fw * 1.0, ov, grid_rows=grid_rows)
        wi, wm, cs = [], [], []
        for i in range(n_s):
            img = synth.make_image(i, cols, rows_s, Ks[i], Rs[i], device="cpu").numpy()
            tl, a = O.warp(O.PROJ_CYLINDRICAL, img, Ks[i], Rs[i], scale, O.INTER_LINEAR, O.BORDER_REFLECT, full_scan=False)
            _, m = O.warp(O.PROJ_CYLINDRICAL, np.full(img.shape[:2], 255, np.uint8), Ks[i], Rs[i], scale, O.INTER_NEAREST, O.BORDER_CONSTANT, full_scan=False)
            wi.append(a.astype(np.float32)); wm.append(m); cs.append(tl)
        t0 = time.perf_counter()
        ref = O.ref_dp_seam_find(wi, cs, wm)
        t1 = time.perf_counter()
        port = O.dp_seam_find(wi, cs, wm)
        t2 = time.perf_counter()
        return {"find_seconds_reference_1_thread": t1 - t0, "find_seconds_port": t2 - t1,
                "masks_equal": bool(all(np.array_equal(a, b) for a, b in zip(ref, port)))}
    except Exception as e:                      # informational only
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args):
    """--impl reference: the CPU port of the reference path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle as O
    O.build()
    threads = os.cpu_count() or 1
    n, rows, cols, fw, ov, grid_rows, desc = WORKLOADS[args.workload]
    # size the sample so that (steps + warmup) steps stay within a few minutes: probe at 1/8 of the rows
    probe_rows = max(64, rows // 8)
    v, _, dt, _ = cpu_port_run(args.workload, threads, sample_rows=probe_rows)
    budget = 150.0 / max(1, args.steps + args.warmup)
    rows_s = int(min(rows, max(probe_rows, probe_rows * budget / max(dt, 1e-3))))
    rows_s -= rows_s % 32
    rows_s = max(rows_s, 64)
    import numpy as np
    from imagestitch_b200 import synth
    n_s = min(n, 2)
    Ks, Rs, scale = synth.strip_cameras(n, cols, rows_s, fw, ov, grid_rows=grid_rows)
    imgs = [synth.make_image(i, cols, rows_s, Ks[i], Rs[i], device="cpu").numpy() for i in range(n_s)]
    O.set_threads(threads)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.pipeline_run(O.PROJ_CYLINDRICAL, imgs, Ks[:n_s], Rs[:n_s], scale, seam=True, num_bands=NUM_BANDS, weight_type=O.WEIGHT_32F)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    mp = n_s * rows_s * cols / 1e6
    value = mp / (ms / 1e3)
    sample = f"{n_s} neighbouring images of the workload at {rows_s}x{cols} per step (bounded sample), {threads} OpenMP threads"
    line = {"impl": "reference", "metric": "stitched_megapixels_per_sec", "value": value, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s16",
            "data": "synthetic", "config": {"workload": desc, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


PATH_KERNEL_WORDS = ("k_warp", "k_pyrdown", "k_blend", "k_mask_summary", "k_gain", "k_dilate", "k_feather", "k_dt_")   # warp / pyramid / blend side of the path
SEAM_KERNEL_WORDS = ("k_seam", "k_cost", "k_row_toggles", "k_special", "k_label", "k_bt_", "k_apply_clear", "k_ccl", "k_uls", "k_contour", "k_mask_row",
                     "k_mask_update", "k_mask_and", "k_collect", "k_scatter", "k_relabel", "k_classify", "k_sobel", "k_scan")


def _kernel_group(name):
    n = name.strip("()")
    if any(n.startswith(w) for w in PATH_KERNEL_WORDS):
        return "warp_blend"
    if any(n.startswith(w) for w in SEAM_KERNEL_WORDS):
        return "seam"
    return "other"


def run_c1(args, dev, local_rank):
    """BASELINE.json configs[0]: 2 x (768 x 1024) pair, cylindrical warp + the reference's hand-written linear blend
    ([BLEND]:141-717, is_linear_blend_pair) -- the configuration the reference's own CPU path (single core) is quoted on."""
    import numpy as np
    import torch

    import oracle as O
    from imagestitch_b200 import stitching as S, synth
    n, rows, cols, fw, ov, grid_rows, desc = WORKLOADS["c1"]
    Ks, Rs, scale = synth.strip_cameras(n, cols, rows, fw, ov)
    imgs = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device="cpu").numpy() for i in range(n)]
    ctx = S.Context(local_rank, use_torch_stream=True)
    wp = S.RotationWarper(ctx, "cylindrical", scale)
    dimgs = [torch.from_numpy(a).to(dev) for a in imgs]

    def step_device():
        ctx.clear_plan_cache()
        ws, tls = [], []
        for i in range(n):
            tl, w, _m = wp.warp_with_mask(dimgs[i], Ks[i], Rs[i])
            ws.append(w.float()); tls.append(tl)               # images_warped[i].convertTo(images_warped_f[i], CV_32F)  [BLEND]:140
        return S.linear_blend_pair(ctx, ws[0], ws[1], tls[0], tls[1])

    def step_host():
        ctx.clear_plan_cache()
        ws, tls = [], []
        for i in range(n):
            tl, w, _m = wp.warp_with_mask(imgs[i], Ks[i], Rs[i])
            ws.append(w.astype(np.float32)); tls.append(tl)
        return S.linear_blend_pair(ctx, ws[0], ws[1], tls[0], tls[1])

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, (ctx.kernel_launches - l0) // steps

    for _ in range(max(args.warmup, 3)):
        step_device()
    ms_dev, launches = timed(step_device, args.steps)
    step_host()
    ms_e2e, _ = timed(step_host, args.steps)
    got = step_host()
    # parity + CPU baseline: the oracle's restatement of the same two stages, ONE thread (the reference is single threaded)
    O.build()
    O.set_threads(1)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        ws, tls = [], []
        for i in range(n):
            tl, w = O.warp(O.PROJ_CYLINDRICAL, imgs[i], Ks[i], Rs[i], scale, O.INTER_LINEAR, O.BORDER_REFLECT, full_scan=False)
            O.warp(O.PROJ_CYLINDRICAL, np.full(imgs[i].shape[:2], 255, np.uint8), Ks[i], Rs[i], scale, O.INTER_NEAREST, O.BORDER_CONSTANT, full_scan=False)
            ws.append(w.astype(np.float32)); tls.append(tl)
        want = O.lin_blend(ws[0], ws[1], tls[0], tls[1])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    gp = np.asarray(got[0])
    seam_equal = bool(np.array_equal(np.asarray(got[1]), np.asarray(want[1])))
    both = np.isfinite(gp) & np.isfinite(want[0])
    pano_err = float(np.abs(gp[both] - want[0][both]).max()) if both.any() else 0.0
    in_mp = n * rows * cols / 1e6
    line = {"metric": "stitched_megapixels_per_sec", "value": in_mp / (ms_dev * 1e-3), "unit": "MP/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "l2_policy": "inputs (4.7 MB) fit the L2: the plan memo is cleared every step, nothing else is cached between steps"},
            "e2e": {"value": in_mp / (ms_e2e * 1e-3), "unit": "MP/s", "h2d_bytes_per_step": n * rows * cols * 3, "d2h_bytes_per_step": int(gp.size * 4), "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "parity": {"seam_indices_equal_oracle": seam_equal, "pano_max_abs_err": pano_err, "pano_tolerance": 1e-3 * 255},
            "roofline": None,
            "cpu_baseline": {"value": in_mp / best, "unit": "MP/s", "cores": 1, "kind": "port", "seconds": best,
                             "sample": "the whole workload (2 warps with masks + [BLEND]:141-717 pair blend), oracle port, 1 thread, best of 3"}}
    print(json.dumps(line), flush=True)
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-lanes", type=int, default=4, help="N = 1: panoramas in flight in the pipelined end-to-end measurement")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the sharded-vs-single-GPU comparison on rank 0")
    ap.add_argument("--kernel-report", default=None, help="write the per-kernel timing table (JSON) to this file")
    ap.add_argument("--opencv-sample", type=int, default=None, help="internal: time python cv2's path on a sample of this many rows, print JSON")
    ap.add_argument("--reference-find-sample", type=int, default=None, help="internal: the reference's own find() vs the port on such a sample, print JSON")
    ap.add_argument("--weights16s-leg", type=int, default=None, help="internal: time this many device-resident steps with CV_16S blend weights, print JSON")
    args = ap.parse_args()
    if args.opencv_sample is not None:
        print(json.dumps(opencv_run_inline(args.workload, args.opencv_sample)), flush=True)
        return 0
    if args.reference_find_sample is not None:
        print(json.dumps(reference_find_check_inline(args.workload, args.reference_find_sample)), flush=True)
        return 0
    if args.weights16s_leg is not None:
        print(json.dumps(weights16s_leg_inline(args.workload, args.weights16s_leg)), flush=True)
        return 0
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    from imagestitch_b200 import build as B, stitching as S, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    if rank == 0:
        B.build()
    if dist:
        dist.barrier()
    dev = f"cuda:{local_rank}"
    if args.workload == "c1":
        if rank == 0:
            run_c1(args, dev, local_rank)
        if dist:
            dist.destroy_process_group()
        return 0

    n, rows, cols, fw, ov, grid_rows, desc = WORKLOADS[args.workload]
    # Weak scaling: every GPU brings n images; for N > 1 they form ONE panorama of n*N images (a strip, or a grid_rows-row
    # mosaic) that is sharded by column strip (imagestitch_b200.sharded: NCCL halo exchange at the strip boundaries).  A
    # cylinder cannot hold more than 360 degrees, so the focal length grows with N to keep a strip below ~340 degrees.
    import math
    n_all = n * world
    per_row = n_all // grid_rows
    fw = max(fw, 1.0 / (2.0 * math.tan(5.9 / (2.0 * (1.0 - ov) * per_row)))) if world > 1 else fw
    Ks_all, Rs_all, scale = synth.strip_cameras(n_all, cols, rows, fw, ov, grid_rows=grid_rows)
    in_mp = n * rows * cols / 1e6
    ctx = S.Context(local_rank, use_torch_stream=True)     # kernels run on torch's current stream -> torch events see them
    st = S.Stitcher(ctx, "cylindrical", "dp", NUM_BANDS, S.WEIGHT_32F)
    sh = plan = None
    if world == 1:
        idx = list(range(n))
        Ks, Rs = Ks_all, Rs_all
        imgs_dev = [synth.make_image(i, cols, rows, Ks_all[i], Rs_all[i], device=dev) for i in idx]
        corners, sizes, roi = st.plan([(cols, rows)] * n, Ks, Rs, scale)
        out_shape = (roi[3], roi[2])
        pano_dev = torch.empty(out_shape + (3,), dtype=torch.int16, device=dev)
        pmask_dev = torch.empty(out_shape, dtype=torch.uint8, device=dev)
        contexts = [ctx]

        def step_device():
            # a stream of different panoramas never finds its cameras in the plan memo: every step scans its image borders
            # (detectResultRoi, [WARP]:64-88) once -- plan() does it, is_pipeline_run() then reuses it for the same panorama
            ctx.clear_plan_cache()
            st.stitch(imgs_dev, Ks, Rs, scale, out=(pano_dev, pmask_dev))
            return pano_dev, pmask_dev
    else:
        from imagestitch_b200 import sharded
        be = sharded.GpuBackend(local_rank, ctx=ctx)
        contexts = [be.ctx] + be.workers
        corners_all, sizes_all, roi = st.plan([(cols, rows)] * n_all, Ks_all, Rs_all, scale)
        plan = sharded.ShardPlan.build(corners_all, sizes_all, roi, world, NUM_BANDS)
        idx = [i for i in range(n_all) if plan.owner[i] == rank]
        sizes = [sizes_all[i] for i in idx]
        imgs_dev = [synth.make_image(i, cols, rows, Ks_all[i], Rs_all[i], device=dev) for i in idx]
        sh = sharded.ShardedStitcher(be, sharded.Comm(dist), NUM_BANDS)
        out_shape = (roi[3], plan.cuts[rank + 1] - plan.cuts[rank])

        def step_device():
            ctx.clear_plan_cache()
            r = sh.stitch(imgs_dev, Ks_all, Rs_all, scale, plan)
            return r["pano"], r["pano_mask"]
    torch.cuda.synchronize()

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def total_launches():
        return sum(c.kernel_launches for c in contexts)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    windows = []

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        l0 = total_launches()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        windows.append((t0, time.perf_counter()))
        ms = e0.elapsed_time(e1) / steps
        launches = (total_launches() - l0) // steps
        if dist:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    for _ in range(max(args.warmup, 3)):
        step_device()
    ms_dev, launches = timed(step_device, args.steps)
    stage_ms = dict(st.timings_ms) if world == 1 else None
    seam_info = {"path": ctx.seam_path, "waves": ctx.seam_waves} if world == 1 else dict(sh.info.get("seam", {}))
    dev_pano, dev_pmask = step_device()
    dev_pano, dev_pmask = dev_pano.clone(), dev_pmask.clone()

    # second timed region with per-launch events: the roofline figures
    for c in contexts:
        c.kernel_timing(True)
        c.kernel_timing_report()
    ms_dev_ev, _ = timed(step_device, args.steps)
    ktable = {}
    for c in contexts:
        for r in c.kernel_timing_report():
            a = ktable.setdefault(r["name"], {"name": r["name"], "launches": 0, "ms": 0.0, "bytes": 0.0})
            a["launches"] += r["launches"]; a["ms"] += r["ms"]; a["bytes"] += r["bytes"]
        c.kernel_timing(False)
    ktable = sorted(ktable.values(), key=lambda r: -r["ms"])

    # end to end with HOST buffers (pinned): H2D of the sources and D2H of the panorama (strip) inside the timed region
    imgs_pin = [torch.empty((rows, cols, 3), dtype=torch.uint8, pin_memory=True) for _ in range(n)]
    for p, d in zip(imgs_pin, imgs_dev):
        p.copy_(d)
    pano_pin = torch.empty(out_shape + (3,), dtype=torch.int16, pin_memory=True)
    pmask_pin = torch.empty(out_shape, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    e2e_pipe = None
    if world == 1:
        imgs_host = [p.numpy() for p in imgs_pin]
        out_host = (pano_pin.numpy(), pmask_pin.numpy())

        def step_host():                                  # host buffers straight through the C ABI (is_pipeline_run)
            ctx.clear_plan_cache()
            st.stitch(imgs_host, Ks, Rs, scale, out=out_host)
    else:
        def step_host():                                  # per rank: its sources up, its strip of the panorama down
            ctx.clear_plan_cache()
            up = [p.to(dev, non_blocking=True) for p in imgs_pin]
            r = sh.stitch(up, Ks_all, Rs_all, scale, plan)
            pano_pin.copy_(r["pano"], non_blocking=True)
            pmask_pin.copy_(r["pano_mask"], non_blocking=True)
            torch.cuda.synchronize()

    step_host()
    e2e_steps = max(2, min(args.steps, 6))
    ms_e2e_serial, _ = timed(step_host, e2e_steps)
    same = bool(torch.equal(dev_pano.cpu(), pano_pin)) and bool(torch.equal(dev_pmask.cpu(), pmask_pin))
    ms_e2e = ms_e2e_serial
    if world == 1:
        # Throughput of a STREAM of panoramas through the same synchronous call: a few contexts on as many host threads (one context
        # per thread is the library's threading model), each with its own pinned output buffers, so that the D2H of one panorama
        # runs under the H2D + compute of the next ones (the host link is full duplex).  Every step still uploads its sources and
        # downloads its panorama inside the timed region.
        n_lanes = max(2, args.e2e_lanes)
        extra = [S.Context(local_rank) for _ in range(n_lanes - 1)]
        lanes = [(ctx, st, out_host)]
        pins = []
        for c in extra:
            pp = torch.empty(out_shape + (3,), dtype=torch.int16, pin_memory=True)
            pm = torch.empty(out_shape, dtype=torch.uint8, pin_memory=True)
            pins.append((pp, pm))
            lanes.append((c, S.Stitcher(c, "cylindrical", "dp", NUM_BANDS, S.WEIGHT_32F), (pp.numpy(), pm.numpy())))

        def lane_steps(k, count):
            c, s_, out = lanes[k]
            if k:                                         # panorama k arrives a fraction of a step later: its upload + compute run
                time.sleep(ms_e2e_serial * 1e-3 * k / n_lanes)   # under the others' downloads instead of competing with their uploads
            for _ in range(count):
                c.clear_plan_cache()
                s_.stitch(imgs_host, Ks, Rs, scale, out=out)

        def run_lanes(count_each):
            th = [threading.Thread(target=lane_steps, args=(k, count_each)) for k in range(n_lanes)]
            for x in th:
                x.start()
            for x in th:
                x.join()

        run_lanes(2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        per_lane = max(3, e2e_steps // 2 + 1)
        run_lanes(per_lane)
        e1.record()
        torch.cuda.synchronize()
        windows.append((t0, time.perf_counter()))
        ms_pipe = e0.elapsed_time(e1) / (n_lanes * per_lane)
        same2 = all(bool(torch.equal(dev_pano.cpu(), pp)) and bool(torch.equal(dev_pmask.cpu(), pm)) for pp, pm in pins)
        e2e_pipe = {"ms_per_step": ms_pipe, "panoramas_in_flight": n_lanes, "steps": n_lanes * per_lane, "matches_device_path": same2}
        ms_e2e = min(ms_e2e_serial, ms_pipe)
        for c in extra:
            c.close()
        del pins
    # what the host link of this box delivers for the same buffers (plain pinned copies): the floor of the e2e figure
    pcie = {}
    try:
        tmp = torch.empty_like(pano_pin, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for name, fn, nbytes in (("h2d_gbs", lambda: tmp.copy_(pano_pin, non_blocking=True), pano_pin.numel() * 2),
                                 ("d2h_gbs", lambda: pano_pin.copy_(tmp, non_blocking=True), pano_pin.numel() * 2)):
            fn(); torch.cuda.synchronize()
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            pcie[name] = nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
        del tmp
    except Exception:
        pass
    if dist:
        t = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        same = bool(t.item())

    # N > 1: the sharded result against the single-GPU pipeline on a bounded subset (rank 0: the images of ranks 0 and 1)
    sharded_matches = None
    if world > 1 and not args.no_parity_check:
        sharded_matches = sharded_parity_check(rank, world, plan, sh, st, ctx, imgs_dev, dev_pano, dev_pmask, Ks_all, Rs_all, scale, rows, cols, dev, dist)
    if sampler:
        sampler.stop()
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    alg = algorithmic_bytes(n, rows, cols, sizes, (roi[0], roi[1], out_shape[1], out_shape[0]))   # this rank's share of the panorama
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    groups = {"warp_blend": 0.0, "seam": 0.0, "other": 0.0}
    for r in ktable:
        groups[_kernel_group(r["name"])] += r["ms"] / args.steps
    # The roofline figure of the PATH: compulsory bytes of warp + blend (sources and seam masks read once, int16
    # panorama + mask written once; every intermediate -- warped images, pyramids -- counts against it) over the summed CUDA-event
    # time of every warp / pyramid / blend kernel of a step.  The DP seam stage is latency bound and reported beside it.
    wb_ms = groups["warp_blend"]
    achieved = alg["warp_blend_fused"] / (wb_ms * 1e-3) / 1e9 if wb_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": "warp + pyramid + blend kernels of a step (summed CUDA-event time)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_step": alg["warp_blend_fused"],
                "algorithmic_bytes_formula": "sum 3hw (u8 BGR sources) + sum A (u8 seam masks) + 7P (s16 BGR panorama + u8 mask); no intermediates",
                "kernel_ms_per_step": {k: round(v, 4) for k, v in groups.items()},
                "whole_step": {"ms": ms_dev, "achieved_gbs": alg["warp_blend_fused"] / (ms_dev * 1e-3) / 1e9, "frac": alg["warp_blend_fused"] / (ms_dev * 1e-3) / 1e9 / peak},
                "kernels": [{"name": r["name"], "group": _kernel_group(r["name"]), "ms_per_step": round(r["ms"] / args.steps, 5), "launches_per_step": r["launches"] / args.steps,
                             "declared_gbs": round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] > 0 and r["bytes"] > 0 else None} for r in ktable]}
    # DRAM traffic of the same kernels from one `ncu --set full` capture of a whole step of THIS workload on one GPU (scripts/
    # ncu_traffic.py -> profiles/r2_ncu_traffic_c2.json): a constant of the profile, not a measurement of this run -- only
    # attached where it applies (N = 1, C2)
    if world == 1 and args.workload == "c2":
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic_c2.json")))
            roofline["traffic"] = sum(v["dram_bytes"] for k, v in tr["kernels"].items() if _kernel_group(k) == "warp_blend") / max(1, tr.get("steps", 1))
            roofline["traffic_source"] = "profiles/r2_ncu_traffic_c2.json: dram__bytes_read.sum + dram__bytes_write.sum of the warp / pyramid / blend kernels of one step (ncu --set full)"
        except Exception:
            pass
    if args.kernel_report:
        os.makedirs(os.path.dirname(os.path.abspath(args.kernel_report)), exist_ok=True)
        with open(args.kernel_report, "w") as f:
            json.dump({"steps": args.steps, "ms_per_step": ms_dev, "ms_per_step_with_events": ms_dev_ev, "kernels": ktable}, f, indent=1)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        import oracle as O
        O.build()
        threads = os.cpu_count() or 1
        sample_rows = rows if rows * cols <= 8e6 else rows // 4
        v, sdesc, dt, stages = cpu_port_run(args.workload, threads, sample_rows=sample_rows, steps=3)
        cpu = {"value": v, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sdesc, "seconds": dt,
               "stage_seconds": dict(zip(("warp", "seam", "blend", "total"), stages)),
               "opencv_cv2_same_sample": opencv_run(args.workload, sample_rows),
               "reference_find_same_sample": reference_find_check(args.workload, sample_rows)}

    h2d = world * n * rows * cols * 3
    d2h = roi[2] * roi[3] * 7
    if world == 1:
        wl = desc
    elif grid_rows > 1:
        wl = f"{n_all}x({rows}x{cols}) RGB {grid_rows}x{per_row} mosaic ({n} images per GPU), cylindrical warp + DP seam masks + multi-band blend (5 bands)"
    else:
        wl = f"{n_all}x({rows}x{cols}) RGB strip ({n} images per GPU), cylindrical warp + DP seam masks + multi-band blend (5 bands)"
    line = {
        "metric": "stitched_megapixels_per_sec", "value": world * in_mp / (ms_dev * 1e-3), "unit": "MP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "s16", "data": "synthetic",
        "config": {"workload": wl, "images_per_gpu": n, "image_rows": rows, "image_cols": cols, "num_bands": NUM_BANDS, "weight_type": "CV_32F",
                   "seam": "dp_color", "projection": "cylindrical", "f_over_w": round(fw, 4), "overlap": ov, "pano_roi": list(roi),
                   "l2_policy": f"inputs ({(h2d // world) >> 20} MiB per GPU) and panorama exceed the {L2_BYTES >> 20} MiB L2; no flush needed",
                   "plan_memo": "cleared at the start of every step: each step scans its image borders (detectResultRoi) once",
                   "parallelism": "single GPU" if world == 1 else
                   f"one {n_all}-image panorama sharded by column strip over {world} GPUs, NCCL P2P halo exchange at strip boundaries",
                   "seam_pairs": seam_info},
        "clocks": sampler.summary(windows) if sampler else None,
        "e2e": {"value": world * in_mp / (ms_e2e * 1e-3), "unit": "MP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "one_panorama_at_a_time": {"ms_per_step": ms_e2e_serial, "value": world * in_mp / (ms_e2e_serial * 1e-3)},
                "panoramas_in_flight": e2e_pipe, "matches_device_path": same, "host_link_pinned_copy": pcie},
        "gpu_launches": int(launches),
        "stage_ms": stage_ms, "ms_per_step_with_kernel_events": ms_dev_ev,
        "roofline": roofline, "cpu_baseline": cpu,
    }
    if sharded_matches is not None:
        line["sharded_matches"] = sharded_matches
    if world == 1:                                        # "also output-pano MP/s", and the CV_16S-weight step
        out_mp = out_shape[0] * out_shape[1] / 1e6
        line["output_panorama"] = {"megapixels": out_mp, "mp_per_s": out_mp / (ms_dev * 1e-3), "mp_per_s_e2e": out_mp / (ms_e2e * 1e-3)}
        # "default weight type CV_32F (report CV_16S too)": the same device-resident step with CV_16S blend weights, timed in a
        # process of its own so that nothing it does can disturb or end this line (informational)
        line["weights_cv_16s"] = _informational_leg("--weights16s-leg", args.workload, max(3, min(args.steps, 10)), timeout=120) if not args.no_cpu_baseline else None
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()
    return 0


def sharded_parity_check(rank, world, plan, sh, st, ctx, imgs_dev, dev_pano, dev_pmask, Ks_all, Rs_all, scale, rows, cols, dev, dist):
    """Rank 0 stitches, on its GPU alone, the images of ranks 0 and 1 (every image its strip can see plus their pair partners)
    with the single-GPU code -- seam finder, then the blender on the GLOBAL panorama ROI restricted to rank 0's columns -- and
    compares bit for bit: the seam masks of its own images and its strip of the panorama.  The other ranks only send images."""
    import torch

    from imagestitch_b200 import stitching as S, synth
    n_all = len(plan.corners)
    subset = [i for i in range(n_all) if plan.owner[i] in (0, 1)]
    ok = True
    if rank == 0:
        r = sh.stitch(imgs_dev, Ks_all, Rs_all, scale, plan)       # one more step, keeping the seam masks
        mine = [i for i in range(n_all) if plan.owner[i] == 0]
        have = dict(zip(mine, imgs_dev))
        imgs = [have[i] if i in have else synth.make_image(i, cols, rows, Ks_all[i], Rs_all[i], device=dev) for i in subset]
        wp = S.RotationWarper(ctx, "cylindrical", scale)
        ws, ms, cs = [], [], []
        for k, i in enumerate(subset):
            tl, w, m = wp.warp_with_mask(imgs[k], Ks_all[i], Rs_all[i])
            ws.append(w); ms.append(m); cs.append(tl)
        sm = S.DpSeamFinder(ctx, "COLOR").find(ws, cs, ms)
        for k, i in enumerate(subset):
            if i in mine:
                ok = ok and bool(torch.equal(sm[k], r["seam_masks"][i]))
        b = S.MultiBandBlender(ctx, 0, NUM_BANDS, S.WEIGHT_32F)
        b.prepare(plan.roi)
        for k, i in enumerate(subset):
            if b.strip_needs(plan.sizes[i], plan.corners[i], plan.cuts[0], plan.cuts[1]):
                b.feed(ws[k], sm[k], cs[k], borrow=True)
        pano, pmask = b.blend_strip(plan.cuts[0], plan.cuts[1])
        ok = ok and bool(torch.equal(pano, r["pano"])) and bool(torch.equal(pmask, r["pano_mask"]))
    else:
        sh.stitch(imgs_dev, Ks_all, Rs_all, scale, plan)           # the step is collective
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


if __name__ == "__main__":
    sys.exit(main())
