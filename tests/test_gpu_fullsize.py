"""Parity at BASELINE.json's full sizes: bench workloads C2 (6 x (4000 x 6000)) and C3 (12 images) against the oracle, plus size-independent properties.

C2 is small enough for the CPU oracle to finish in seconds with all host threads, so the whole panorama is compared bit
for bit; on top of that the properties that do not need an oracle: the concurrent (speculative) seam stage equals the
reference's sequential pair loop, the host-buffer path equals the device-resident path, and the seam masks partition
every overlap (no pixel is claimed by two images, none of the covered pixels is lost).
"""
import os

import numpy as np
import pytest

from imagestitch_b200 import stitching as S, synth

pytestmark = pytest.mark.gpu


def _inputs(n, rows, cols, fw, device):
    Ks, Rs, scale = synth.strip_cameras(n, cols, rows, fw, 0.25)
    imgs = [synth.make_image(i, cols, rows, Ks[i], Rs[i], device=device) for i in range(n)]
    import torch
    torch.cuda.synchronize()      # the library runs on its own stream
    return imgs, Ks, Rs, scale


def _coverage(res, masks):
    """per panorama pixel: how many final masks are set"""
    x0, y0, w, h = res["roi"]
    cnt = np.zeros((h, w), np.uint8)
    for m, c in zip(masks, res["corners"]):
        mm = m.cpu().numpy() if hasattr(m, "cpu") else m
        cnt[c[1] - y0:c[1] - y0 + mm.shape[0], c[0] - x0:c[0] - x0 + mm.shape[1]] += (mm != 0)
    return cnt


def test_c2_full_size_equals_oracle(ctx, oracle, monkeypatch):
    O = oracle
    O.set_threads(os.cpu_count() or 1)
    imgs, Ks, Rs, scale = _inputs(6, 4000, 6000, 1.2, "cuda:0")
    st = S.Stitcher(ctx, "cylindrical", "dp", 5, S.WEIGHT_32F)
    got = st.stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    assert ctx.seam_speculation == 1, "the strip's pairs are expected to validate concurrently"
    host = [t.cpu().numpy() for t in imgs]
    want = O.pipeline_run(O.PROJ_CYLINDRICAL, host, Ks, Rs, scale, seam=True, num_bands=5, weight_type=O.WEIGHT_32F, want_intermediates=True)
    assert got["roi"] == want["roi"]
    for k in range(6):
        assert np.array_equal(got["seam_masks"][k].cpu().numpy(), want["masks"][k]), f"seam mask {k}"
    assert np.array_equal(got["pano_mask"].cpu().numpy(), want["pano_mask"])
    assert np.array_equal(got["pano"].cpu().numpy(), want["pano"])
    # the reference's sequential pair loop gives the same masks and panorama
    monkeypatch.setenv("IS_SEAM_SEQUENTIAL", "1")
    seq = st.stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    monkeypatch.delenv("IS_SEAM_SEQUENTIAL")
    for k in range(6):
        assert bool((seq["seam_masks"][k] == got["seam_masks"][k]).all())
    assert bool((seq["pano"] == got["pano"]).all())
    # host buffers through the same entry point
    h = st.stitch(host, Ks, Rs, scale)
    assert np.array_equal(h["pano"], want["pano"]) and np.array_equal(h["pano_mask"], want["pano_mask"])


def test_c3_properties(ctx, monkeypatch):
    imgs, Ks, Rs, scale = _inputs(12, 4000, 6000, 1.5, "cuda:0")
    st = S.Stitcher(ctx, "cylindrical", "dp", 5, S.WEIGHT_32F)
    a = st.stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    cnt = _coverage(a, a["seam_masks"])
    assert cnt.max() == 1, "a pixel is claimed by two images after the seam stage"
    pm = a["pano_mask"].cpu().numpy()
    assert np.array_equal(pm != 0, cnt == 1), "panorama mask != union of the seam masks"
    monkeypatch.setenv("IS_SEAM_SEQUENTIAL", "1")
    b = st.stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    monkeypatch.delenv("IS_SEAM_SEQUENTIAL")
    for k in range(12):
        assert bool((a["seam_masks"][k] == b["seam_masks"][k]).all()), f"concurrent != sequential, mask {k}"
    assert bool((a["pano"] == b["pano"]).all())
    # idempotence: a second run on the same inputs gives the same bits
    c = st.stitch(imgs, Ks, Rs, scale)
    assert bool((c["pano"] == a["pano"]).all()) and bool((c["pano_mask"] == a["pano_mask"]).all())


def test_c3_full_size_equals_oracle(ctx, oracle):
    """BASELINE.json configs[2] (12 x (4000 x 6000), DP seam + 5-band blend) against the oracle at full size: seam masks, panorama
    mask and panorama bit for bit (about half a minute of CPU with all host threads)."""
    O = oracle
    O.set_threads(os.cpu_count() or 1)
    imgs, Ks, Rs, scale = _inputs(12, 4000, 6000, 1.5, "cuda:0")
    st = S.Stitcher(ctx, "cylindrical", "dp", 5, S.WEIGHT_32F)
    got = st.stitch(imgs, Ks, Rs, scale, want_seam_masks=True)
    host = [t.cpu().numpy() for t in imgs]
    want = O.pipeline_run(O.PROJ_CYLINDRICAL, host, Ks, Rs, scale, seam=True, num_bands=5, weight_type=O.WEIGHT_32F, want_intermediates=True)
    assert got["roi"] == want["roi"]
    for k in range(12):
        assert np.array_equal(got["seam_masks"][k].cpu().numpy(), want["masks"][k]), f"seam mask {k}"
    assert np.array_equal(got["pano_mask"].cpu().numpy(), want["pano_mask"])
    assert np.array_equal(got["pano"].cpu().numpy(), want["pano"])
