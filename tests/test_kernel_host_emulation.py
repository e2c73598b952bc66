"""Host emulation of device code that could not be run on hardware when it was written (the COLOR_GRAD additions to the
seam cost kernels): the regions of imagestitch_b200/csrc/seam.cu marked @emu-begin / @emu-end are compiled for the host
(tests/emu/cuda_host_emul.h: qualifiers vanish, *_rn intrinsics are the IEEE operations, threadIdx/blockIdx are stepped by
a loop) and their results compared with the oracle bit for bit.  This checks the per-thread arithmetic and every index
computation of k_sobel_window and k_cost_pq<T, GRAD>; it does not check the launch plumbing around them."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
OUT = os.path.join(EMU, "_build")


@pytest.fixture(scope="module")
def emu():
    src = open(os.path.join(ROOT, "imagestitch_b200", "csrc", "seam.cu")).read()
    regions = re.findall(r"// @emu-begin[^\n]*\n(.*?)// @emu-end", src, flags=re.S)
    assert len(regions) == 3, "expected three marked regions in seam.cu"
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "seam_regions.inc"), "w") as f:
        f.write("\n".join(regions))
    so = os.path.join(OUT, "libseam_cost_emul.so")
    subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-I", OUT, "-shared", "-o", so,
                           os.path.join(EMU, "seam_cost_emul.cpp")])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_sobel_window_kernel_matches_oracle(emu, oracle, dtype):
    O = oracle
    rng = np.random.default_rng(5)
    h, w = 57, 83
    img = rng.integers(0, 256, (h, w, 3)).astype(dtype)
    if dtype == np.float32:
        img += rng.uniform(-0.5, 0.5, img.shape).astype(np.float32)
    want_gx, want_gy = O.seam_gradients(img)
    # windows: the whole image (touches all four borders -> reflect-101) and an interior window with an offset frame
    for (dx, dy, ox, oy, ww, wh) in ((0, 0, 0, 0, w, h), (-7, 3, 7 + 10, 5 - 3, 40, 31), (5, -2, -5, 2, 1, h), (0, 0, w - 1, h - 1, 1, 1)):
        pitch = (ww + 31) & ~31
        gx = np.full((wh, pitch), np.nan, np.float32)
        gy = np.full((wh, pitch), np.nan, np.float32)
        emu.emu_sobel_window(_p(img), 1 if dtype == np.uint8 else 0, h, w, dx, dy, ox, oy, ww, wh, _p(gx), _p(gy), pitch)
        ix0, iy0 = ox + dx, oy + dy                      # image coordinates of the window's first pixel
        assert np.array_equal(gx[:, :ww].view(np.uint32), want_gx[iy0:iy0 + wh, ix0:ix0 + ww].view(np.uint32))
        assert np.array_equal(gy[:, :ww].view(np.uint32), want_gy[iy0:iy0 + wh, ix0:ix0 + ww].view(np.uint32))
        assert np.isnan(gx[:, ww:]).all()                # nothing written outside the window


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("grad", [False, True])
@pytest.mark.parametrize("horizontal", [0, 1])
def test_cost_pq_kernel_matches_oracle(emu, oracle, dtype, grad, horizontal):
    """P / Q in DP layout == costV / costH of the oracle (== the reference's computeCosts, test_oracle_reference_build.py)"""
    O = oracle
    rng = np.random.default_rng(3)
    h1, w1, h2, w2 = 60, 80, 70, 64
    tl1, tl2 = (5, -3), (40, 4)
    utl = (min(tl1[0], tl2[0]), min(tl1[1], tl2[1]))
    ubr = (max(tl1[0] + w1, tl2[0] + w2), max(tl1[1] + h1, tl2[1] + h2))
    W, H = ubr[0] - utl[0], ubr[1] - utl[1]
    labels = np.zeros((H, W), np.int32)
    ix0, iy0 = tl2[0] - utl[0], tl2[1] - utl[1]                         # intersection rectangle in the union frame
    ix1, iy1 = tl1[0] + w1 - utl[0], tl1[1] + h1 - utl[1]
    labels[iy0:iy1, ix0:ix1] = 2
    labels[iy0 + 5:iy0 + 9, ix0 + 3:ix0 + 10] = 1                      # a hole of another label inside the component
    rx, ry, rw, rh = ix0, iy0, ix1 - ix0, iy1 - iy0
    a = rng.integers(0, 256, (h1, w1, 3)).astype(dtype)
    b = rng.integers(0, 256, (h2, w2, 3)).astype(dtype)
    if dtype == np.float32:
        a += rng.uniform(-0.5, 0.5, a.shape).astype(np.float32)
        b += rng.uniform(-0.5, 0.5, b.shape).astype(np.float32)
    cost = O.COST_COLOR_GRAD if grad else O.COST_COLOR
    want_v, want_h = O.seam_costs(a, b, tl1, tl2, labels, utl, 2, (rx, ry, rw, rh), cost)
    dx1, dy1, dx2, dy2 = utl[0] - tl1[0], utl[1] - tl1[1], utl[0] - tl2[0], utl[1] - tl2[1]
    g = None
    gpitch = (rw + 31) & ~31
    if grad:                                                             # what PairSeam::compute_gradients launches
        g = np.zeros((4, rh, gpitch), np.float32)
        emu.emu_sobel_window(_p(a), 1 if dtype == np.uint8 else 0, h1, w1, dx1, dy1, rx, ry, rw, rh, _p(g[0]), _p(g[1]), gpitch)
        emu.emu_sobel_window(_p(b), 1 if dtype == np.uint8 else 0, h2, w2, dx2, dy2, rx, ry, rw, rh, _p(g[2]), _p(g[3]), gpitch)
    lanes, steps = (rh, rw) if horizontal else (rw, rh)
    pitch = ((lanes + 127) // 128) * 128
    P = np.full((steps, pitch), np.nan, np.float32)
    Q = np.full((steps, pitch), np.nan, np.float32)
    emu.emu_cost_pq(_p(a), _p(b), 1 if dtype == np.uint8 else 0, h1, w1, h2, w2, dx1, dy1, dx2, dy2, _p(labels), H, W, 2, rx, ry, rw, rh, horizontal,
                    _p(g) if grad else None, gpitch, rx, ry, rh, _p(P), _p(Q), pitch)
    inside = labels[ry:ry + rh, rx:rx + rw] == 2
    cv_, ch_ = want_v[:, :rw], want_h[:rh, :]                            # costV is h x (w+1), costH (h+1) x w
    if horizontal:                                                       # step = x, lane = y: P = costH, Q = costV
        gotP, gotQ, wantP, wantQ, ins = P[:, :lanes].T, Q[:, :lanes].T, ch_, cv_, inside
    else:
        gotP, gotQ, wantP, wantQ, ins = P[:, :lanes], Q[:, :lanes], cv_, ch_, inside
    assert np.array_equal(gotQ.view(np.uint32), wantQ.view(np.uint32))
    assert np.array_equal(gotP[ins].view(np.uint32), wantP[ins].view(np.uint32))
    assert np.isinf(gotP[~ins]).all()                                    # cells outside the component can never be on a path
    assert np.isinf(P[:, lanes:]).all() and (Q[:, lanes:] == 0).all()    # padding lanes
