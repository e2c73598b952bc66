"""Host emulation of device code: the warp path (camera products, ROI scan, trig tables, k_warp, k_build_maps -- the product's
source checked against the oracle without a GPU) and the code that could not be run on hardware when it was written (the
COLOR_GRAD additions to the seam cost kernels).  The regions of imagestitch_b200/csrc/the .cu files marked @emu-begin / @emu-end are compiled for the host
(tests/emu/cuda_host_emul.h: qualifiers vanish, *_rn intrinsics are the IEEE operations, threadIdx/blockIdx are stepped by
a loop) and their results compared with the oracle bit for bit.  This checks the per-thread arithmetic and every index
computation of k_sobel_window and k_cost_pq<T, GRAD>; it does not check the launch plumbing around them."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.timeout(600)      # the multi-threaded emulator spins on emulated mbarriers: never let a bug hang the suite

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
OUT = os.path.join(EMU, "_build")


def _build(name, sources, n_regions, inc_name):
    text = "\n".join(open(os.path.join(ROOT, "imagestitch_b200", "csrc", f)).read() for f in sources)
    regions = re.findall(r"// @emu-begin[^\n]*\n(.*?)// @emu-end", text, flags=re.S)
    assert len(regions) == n_regions, f"expected {n_regions} marked regions in {sources}, found {len(regions)}"
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, inc_name), "w") as f:
        f.write("\n".join(regions))
    so = os.path.join(OUT, f"lib{name}.so")
    subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-I", OUT, "-shared", "-o", so,
                           os.path.join(EMU, f"{name}.cpp")])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def emu():
    return _build("seam_cost_emul", ["seam.cu"], 3, "seam_regions.inc")


@pytest.fixture(scope="module")
def emu_warp():
    lib = _build("warp_emul", ["internal.cuh", "warp.cu"], 2, "warp_regions.inc")
    lib.emu_warp.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu_warp_g1():
    """the fused warp + mask + Gaussian level 1 kernel on the multi-threaded block emulator"""
    regions = []
    for f in ("internal.cuh", "warp.cu"):
        regions += re.findall(r"// @emu-begin[^\n]*\n(.*?)// @emu-end", open(os.path.join(ROOT, "imagestitch_b200", "csrc", f)).read(), flags=re.S)
    g1 = re.findall(r"// @emu-g1-begin[^\n]*\n(.*?)// @emu-g1-end", open(os.path.join(ROOT, "imagestitch_b200", "csrc", "warp.cu")).read(), flags=re.S)
    assert len(regions) == 2 and len(g1) == 1
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "warp_g1_regions.inc"), "w") as f:
        # `__shared__ __align__(n) T x` -> `alignas(n) static T x` (standard attributes must come first)
        f.write(re.sub(r"__shared__\s+__align__\((\d+)\)", r"alignas(\1) static", "\n".join(regions + g1)))
    so = os.path.join(OUT, "libwarp_g1_emul.so")
    subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-I", OUT, "-shared",
                           "-o", so, os.path.join(EMU, "warp_g1_emul.cpp")])
    lib = C.CDLL(so)
    lib.emu_warp_g1.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu_orb():
    """the ORB kernels and their driver (orb_find_core) on the host"""
    text = open(os.path.join(ROOT, "imagestitch_b200", "csrc", "orb.cu")).read()
    regions = re.findall(r"// @emu-begin[^\n]*\n(.*?)// @emu-end", text, flags=re.S)
    assert len(regions) == 1
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "orb_region.inc"), "w") as f:
        f.write(regions[0])
    so = os.path.join(OUT, "liborb_emul.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-I", OUT, "-I",
                           os.path.join(ROOT, "imagestitch_b200", "csrc"), "-shared", "-o", so, os.path.join(EMU, "orb_emul.cpp")])
    lib = C.CDLL(so)
    lib.emu_orb_find.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu_dp():
    """the DP kernel on the multi-threaded block emulator (tests/emu/cuda_host_emul_mt.h, tests/emu/tma.cuh)"""
    src = open(os.path.join(ROOT, "imagestitch_b200", "csrc", "seam.cu")).read()
    regions = re.findall(r"// @emu-dp-begin[^\n]*\n(.*?)// @emu-dp-end", src, flags=re.S)
    assert len(regions) == 1
    # dynamic shared memory: `extern __shared__ ... unsigned char name[];` becomes a pointer to the emulator's buffer
    text, n = re.subn(r"extern\s+__shared__[^;]*?unsigned char (\w+)\[\];", r"unsigned char* \1 = emu_dynamic_smem;", regions[0])
    assert n == 1
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "seam_dp_region.inc"), "w") as f:
        f.write(text)
    so = os.path.join(OUT, "libseam_dp_emul.so")
    subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-I", OUT, "-shared",
                           "-o", so, os.path.join(EMU, "seam_dp_emul.cpp")])
    lib = C.CDLL(so)
    lib.emu_seam_dp.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu_linblend():
    text = open(os.path.join(ROOT, "imagestitch_b200", "csrc", "linblend.cu")).read()
    regions = re.findall(r"// @emu-begin[^\n]*\n(.*?)// @emu-end", text, flags=re.S)
    assert len(regions) == 1
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "linblend_region.inc"), "w") as f:
        f.write(regions[0])
    so = os.path.join(OUT, "liblinblend_emul.so")
    subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-I", OUT, "-shared",
                           "-o", so, os.path.join(EMU, "linblend_emul.cpp")])
    lib = C.CDLL(so)
    lib.emu_linear_blend_pair.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu_feather():
    text = open(os.path.join(ROOT, "imagestitch_b200", "csrc", "feather.cu")).read()
    regions = re.findall(r"// @emu-begin[^\n]*\n(.*?)// @emu-end", text, flags=re.S)
    assert len(regions) == 2
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "feather_regions.inc"), "w") as f:
        f.write("\n".join(regions))
    so = os.path.join(OUT, "libfeather_emul.so")
    subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-I", OUT, "-shared",
                           "-o", so, os.path.join(EMU, "feather_emul.cpp")])
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


def _f9(m):
    return np.ascontiguousarray(np.asarray(m, np.float32).reshape(9))


@pytest.mark.parametrize("proj", [0, 1, 2, 3, 4])
def test_warp_kernels_match_oracle(emu_warp, oracle, proj):
    """ROI, backward maps, warped image and warped mask of the product's warp source == the oracle (== cv2 / the reference)"""
    from helpers import random_camera
    O = oracle
    rng = np.random.default_rng(10 + proj)
    for t in range(3):
        w, h = int(rng.integers(90, 260)), int(rng.integers(70, 200))
        K, R, scale = random_camera(rng, w, h)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        roi = np.zeros(4, np.int32)
        emu_warp.emu_warp_roi(proj, w, h, _p(_f9(K)), _p(_f9(R)), C.c_float(scale), _p(roi))
        oroi, oxm, oym = O.build_maps(proj, (w, h), K, R, scale, full_scan=True)
        assert tuple(int(v) for v in roi) == oroi
        dh, dw = oroi[3] - oroi[1] + 1, oroi[2] - oroi[0] + 1
        xm = np.empty((dh, dw), np.float32)
        ym = np.empty((dh, dw), np.float32)
        emu_warp.emu_build_maps(proj, w, h, _p(_f9(K)), _p(_f9(R)), C.c_float(scale), _p(xm), _p(ym))
        assert np.array_equal(xm.view(np.uint32), oxm.view(np.uint32)) and np.array_equal(ym.view(np.uint32), oym.view(np.uint32))
        dst = np.zeros((dh, dw, 3), np.uint8)
        msk = np.zeros((dh, dw), np.uint8)
        assert emu_warp.emu_warp(proj, _p(img), h, w, 3, C.c_size_t(w * 3), _p(_f9(K)), _p(_f9(R)), C.c_float(scale), O.INTER_LINEAR, O.BORDER_REFLECT,
                                 _p(dst), C.c_size_t(dw * 3), _p(msk), C.c_size_t(dw)) == 0
        _, want = O.warp(proj, img, K, R, scale, O.INTER_LINEAR, O.BORDER_REFLECT)
        _, wmask = O.warp(proj, np.full((h, w), 255, np.uint8), K, R, scale, O.INTER_NEAREST, O.BORDER_CONSTANT)
        assert np.array_equal(dst, want) and np.array_equal(msk, wmask)
        # the stand-alone mask warp (1 channel, nearest, constant border) through the same kernel
        m1 = np.zeros((dh, dw), np.uint8)
        assert emu_warp.emu_warp(proj, _p(np.full((h, w), 255, np.uint8)), h, w, 1, C.c_size_t(w), _p(_f9(K)), _p(_f9(R)), C.c_float(scale),
                                 O.INTER_NEAREST, O.BORDER_CONSTANT, _p(m1), C.c_size_t(dw), None, C.c_size_t(0)) == 0
        assert np.array_equal(m1, wmask)


@pytest.mark.parametrize("proj", [0, 1, 2])
def test_fused_warp_g1_kernel_matches_oracle(emu_warp_g1, oracle, proj):
    """k_warp_g1 (the pipeline's warp: image + all-255 mask + pyrDown of the BORDER_REFLECT-padded frame, shared-memory tile, dp4a
    sampling, three-row software pipeline) thread by thread on the host == oracle warp, mask and pyrDown(copyMakeBorder(warp))"""
    from helpers import random_camera
    O = oracle
    rng = np.random.default_rng(40 + proj)
    ran = 0
    for t in range(2):
        w, h = int(rng.integers(100, 200)), int(rng.integers(70, 130))
        K, R, scale = random_camera(rng, w, h)
        if proj == 2 and t == 1:
            R = np.ascontiguousarray(__import__("helpers").rot(1.15, 0.05, 0.02))    # far off axis: the plane's unguarded quotients grow large
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if t == 1:
            img = np.ascontiguousarray(np.concatenate([img, img[:, :1]], axis=1)[:, :w + 1])[:, :w]     # a pitch that is not a multiple of 4
        _, want = O.warp(proj, img, K, R, scale, O.INTER_LINEAR, O.BORDER_REFLECT)
        _, wmask = O.warp(proj, np.full((h, w), 255, np.uint8), K, R, scale, O.INTER_NEAREST, O.BORDER_CONSTANT)
        dh, dw = want.shape[:2]
        if dh * dw > 400 * 400:
            continue
        top, left = (5, 9) if t == 0 else (0, 16)
        height, width = -(-(dh + top + 3) // 16) * 16, -(-(dw + left + 1) // 16) * 16
        dst = np.zeros((dh, dw, 3), np.uint8)
        msk = np.zeros((dh, dw), np.uint8)
        g1 = np.zeros(((height + 1) // 2, (width + 1) // 2, 3), np.int16)
        assert emu_warp_g1.emu_warp_g1(proj, _p(img), h, w, C.c_size_t(img.strides[0]), _p(_f9(K)), _p(_f9(R)), C.c_float(scale), _p(dst), _p(msk),
                                       top, left, height, width, _p(g1)) == 0
        assert np.array_equal(dst, want) and np.array_equal(msk, wmask), (proj, t)
        frame = np.pad(want, ((top, height - top - dh), (left, width - left - dw), (0, 0)), mode="symmetric").astype(np.int16)   # copyMakeBorder(BORDER_REFLECT)
        assert np.array_equal(g1, O.pyr_down_s16(frame)), (proj, t)
        ran += 1
    assert ran >= 1


def test_remap_kernel_matches_oracle_incl_extreme_maps(emu_warp, oracle):
    """k_remap (is_remap, and the fisheye / stereographic warps) on caller maps with NaN, infinities and values beyond the int range:
    cv::remap samples those at cvRound's INT_MIN on x86 (the oracle is pinned to cv2 on exactly these values)."""
    O = oracle
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)
    vals = np.array([np.nan, np.inf, -np.inf, 3e9, -3e9, 1e8, -1e8, 7e7, -7e7, 2 ** 31 / 32, 2 ** 31 / 32 - 4, -2 ** 31 / 32, 1e12, -1e12, 1e20, -1e20,
                     6.7e7, 40.3, -0.5, 0.5, 1.5, 79.5, 78.999, -1.0, 32767.4, 32768.6, -32768.5], np.float32)
    xm = np.concatenate([np.tile(vals, (len(vals), 1)), rng.uniform(-200, 300, (len(vals), len(vals))).astype(np.float32)])
    ym = np.concatenate([np.tile(vals, (len(vals), 1)).T, rng.uniform(-150, 250, (len(vals), len(vals))).astype(np.float32)])
    xm, ym = np.ascontiguousarray(xm), np.ascontiguousarray(ym)
    dh, dw = xm.shape
    for ch in (3, 1):
        src = img if ch == 3 else np.ascontiguousarray(img[:, :, 1])
        for interp in (O.INTER_LINEAR, O.INTER_NEAREST):
            for border in (O.BORDER_REFLECT, O.BORDER_CONSTANT):
                dst = np.zeros((dh, dw, 3) if ch == 3 else (dh, dw), np.uint8)
                assert emu_warp.emu_remap(_p(src), 60, 80, ch, C.c_size_t(80 * ch), _p(xm), _p(ym), dh, dw, interp, border, _p(dst), C.c_size_t(dw * ch),
                                          None, C.c_size_t(0)) == 0
                assert np.array_equal(dst, O.remap(src, xm, ym, interp, border)), (ch, interp, border)


def test_warp_kernel_other_sampling_modes(emu_warp, oracle):
    from helpers import random_camera
    O = oracle
    rng = np.random.default_rng(77)
    w, h = 150, 110
    K, R, scale = random_camera(rng, w, h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    oroi = O.detect_roi(0, (w, h), K, R, scale)
    dh, dw = oroi[3] - oroi[1] + 1, oroi[2] - oroi[0] + 1
    for ch, interp, border in ((3, O.INTER_LINEAR, O.BORDER_CONSTANT), (3, O.INTER_NEAREST, O.BORDER_REFLECT), (1, O.INTER_LINEAR, O.BORDER_REFLECT)):
        src = img if ch == 3 else np.ascontiguousarray(img[:, :, 0])
        dst = np.zeros((dh, dw, 3) if ch == 3 else (dh, dw), np.uint8)
        assert emu_warp.emu_warp(0, _p(src), h, w, ch, C.c_size_t(w * ch), _p(_f9(K)), _p(_f9(R)), C.c_float(scale), interp, border, _p(dst),
                                 C.c_size_t(dw * ch), None, C.c_size_t(0)) == 0
        _, want = O.warp(0, src, K, R, scale, interp, border)
        assert np.array_equal(dst, want), (ch, interp, border)


@pytest.mark.parametrize("case", [(200, 150, 0.3, 0, 4, 8, 2), (180, 140, 0.45, 3, 4, 3, 2), (150, 110, 0.35, -4, 8, 16, 2), (260, 120, 0.3, 2, 4, 5, 3)])
def test_dp_kernel_matches_oracle(emu, emu_dp, oracle, case):
    """k_seam_dp (forward pass over the TMA ring, reachability, back-track) fed by k_cost_pq, both from the product's source and
    both emulated on the host, against the seam the oracle traces for the same pair (== the reference's estimateSeam)."""
    from helpers import warped_set
    O = oracle
    w, h, ov, dy, lpt, G, D = case
    corners, wi, wm = warped_set(O, 2, w, h, overlap=ov)
    corners = [tuple(int(v) for v in corners[0]), (int(corners[1][0]), int(corners[0][1]) + dy)]
    got_masks, trace = O.dp_seam_find(wi, corners, wm, want_trace=True)
    assert len(trace) == 1
    _i, _j, _comp, horizontal, pts = trace[0]
    assert not horizontal
    (h1, w1), (h2, w2) = wi[0].shape[:2], wi[1].shape[:2]
    utl = (min(corners[0][0], corners[1][0]), min(corners[0][1], corners[1][1]))
    ubr = (max(corners[0][0] + w1, corners[1][0] + w2), max(corners[0][1] + h1, corners[1][1] + h2))
    W, H = ubr[0] - utl[0], ubr[1] - utl[1]
    m1 = np.zeros((H, W), np.uint8)
    m2 = np.zeros((H, W), np.uint8)
    m1[corners[0][1] - utl[1]:corners[0][1] - utl[1] + h1, corners[0][0] - utl[0]:corners[0][0] - utl[0] + w1] = wm[0]
    m2[corners[1][1] - utl[1]:corners[1][1] - utl[1] + h2, corners[1][0] - utl[0]:corners[1][0] - utl[0] + w2] = wm[1]
    inters = (m1 > 0) & (m2 > 0)                                           # the INTERS component of a two-image strip
    ys, xs = np.nonzero(inters)
    rx, ry, rw, rh = int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)
    labels = np.where(inters, 2, 0).astype(np.int32)
    dx1, dy1, dx2, dy2 = utl[0] - corners[0][0], utl[1] - corners[0][1], utl[0] - corners[1][0], utl[1] - corners[1][1]
    nt = ((rw + lpt - 1) // lpt + 31) // 32 * 32
    pitch = nt * lpt
    P = np.zeros((rh, pitch), np.float32)
    Q = np.zeros((rh, pitch), np.float32)
    emu.emu_cost_pq(_p(wi[0]), _p(wi[1]), 1, h1, w1, h2, w2, dx1, dy1, dx2, dy2, _p(labels), H, W, 2, rx, ry, rw, rh, 0, None, 0, 0, 0, 0, _p(P), _p(Q), pitch)
    seam = np.asarray(pts) - np.asarray(utl) - np.asarray([rx, ry])           # bbox coordinates (x = lane, y = step)
    order = np.argsort(seam[:, 1])
    seam = seam[order]
    s0, lane0, s1, lane1 = int(seam[0, 1]), int(seam[0, 0]), int(seam[-1, 1]), int(seam[-1, 0])
    assert np.array_equal(seam[:, 1], np.arange(s0, s1 + 1))                # a vertical seam has one point per step
    control = np.zeros((rh, pitch), np.uint8)
    seam_lane = np.full(rh + 1, -1, np.int32)
    reached = np.zeros(1, np.int32)
    assert emu_dp.emu_seam_dp(_p(P), _p(Q), rw, pitch, rh, s0, lane0, s1, lane1, lpt, G, D, _p(control), _p(seam_lane), _p(reached)) == 0
    assert reached[0] == 1
    assert np.array_equal(seam_lane[:s1 - s0 + 1], seam[:, 0]), "DP kernel seam differs from the oracle's"
    # an unreachable destination: a wall of +inf across the component
    P2 = P.copy()
    P2[(s0 + s1) // 2, :] = np.inf
    assert emu_dp.emu_seam_dp(_p(P2), _p(Q), rw, pitch, rh, s0, lane0, s1, lane1, lpt, G, D, _p(control), _p(seam_lane), _p(reached)) == 0
    assert reached[0] == 0


def test_pair_blend_kernels_match_oracle(emu_linblend, oracle):
    """the six kernels of is_linear_blend_pair, emulated, against the oracle -- which equals the reference's own compiled block
    bit for bit (tests/test_oracle_reference_build.py): greedy seam, panorama and its NaN pattern"""
    from helpers import warped_set
    O = oracle
    exact = total = 0
    for (w, h, ov) in ((240, 180, 0.25), (200, 160, 0.4)):
        corners, wi, _ = warped_set(O, 2, w, h, overlap=ov)
        a, b = wi[0].astype(np.float32), wi[1].astype(np.float32)
        for tl2 in (corners[1], (corners[1][0], corners[0][1]), (corners[1][0], corners[0][1] + 6), (corners[1][0], corners[0][1] - 5)):
            want = O.lin_blend(a, b, corners[0], tl2)
            pano = np.zeros_like(want[0])
            seam = np.zeros(want[0].shape[0], np.int32)
            rc = emu_linblend.emu_linear_blend_pair(_p(a), a.shape[0], a.shape[1], _p(b), b.shape[0], b.shape[1], int(corners[0][0]), int(corners[0][1]),
                                                    int(tl2[0]), int(tl2[1]), _p(pano), _p(seam))
            assert rc == 0
            assert np.array_equal(seam, want[1]), f"greedy seam differs, tl2={tl2}"
            assert np.array_equal(np.isnan(pano), np.isnan(want[0]))
            d = np.nanmax(np.abs(pano - want[0]))
            assert d <= 1e-3 * 255, d                                         # the bar of tests/test_gpu_parity.py
            total += 1
            exact += int(np.array_equal(np.nan_to_num(pano).view(np.uint32), np.nan_to_num(want[0]).view(np.uint32)))
    print(f"pair blend kernels: {exact} of {total} panoramas bit-exact against the oracle")
    assert exact == total, f"only {exact} of {total} panoramas are bit-exact"
    z = np.zeros((40, 50, 3), np.float32)
    assert emu_linblend.emu_linear_blend_pair(_p(z), 40, 50, _p(z), 40, 50, 0, 0, 5000, 0, _p(np.zeros(3, np.float32)), _p(np.zeros(1, np.int32))) == 1


def test_feather_path_kernels_match_oracle(emu_feather, oracle):
    """the blend path the reference's mains execute ([SEAM]:1249-1282): dilate 20x20 & mask, distance-transform weight map,
    feather blend -- the product's kernels, emulated, against the oracle (== cv2, tests/test_oracle_cv2.py)"""
    from helpers import blob_masks
    O = oracle
    rng = np.random.default_rng(21)
    n = 3
    imgs, masks, corners, x = [], [], [], 0
    for i in range(n):
        h, w = int(rng.integers(50, 80)), int(rng.integers(70, 110))
        imgs.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        masks.append(blob_masks(rng, [(h, w)])[0])
        corners.append((x, int(rng.integers(0, 9))))
        x += int(w * rng.uniform(0.5, 0.8))
    for k, m in enumerate(masks + [np.zeros((9, 40), np.uint8), np.full((30, 33), 255, np.uint8)]):
        for (kw, kh) in (((20, 20), (3, 7)) if k < 2 else ((20, 20),)):  # dilation, with and without the `&`
            stripes = np.where(np.arange(m.shape[1])[None, :] % 5 > 0, 255, 0).astype(np.uint8).repeat(m.shape[0], 0)
            for andm in ((None, stripes) if kw == 20 else (None,)):
                got = m.copy()
                emu_feather.emu_mask_dilate_and(_p(got), m.shape[0], m.shape[1], kw, kh, _p(andm) if andm is not None else None)
                want = O.dilate_rect(m, (kw, kh))
                if andm is not None:
                    want = want & andm
                assert np.array_equal(got, want), (m.shape, kw, kh, andm is not None)
        for sharp in (0.02, 0.1, 5.0):                                   # weight map; no zero pixel -> FLT_MAX * sharpness, clamped to 1
            got = np.zeros(m.shape, np.float32)
            emu_feather.emu_feather_weight(_p(m), m.shape[0], m.shape[1], C.c_float(sharp), _p(got))
            assert np.array_equal(got.view(np.uint32), O.feather_weight(m, sharp).view(np.uint32)), (m.shape, sharp)
    sizes = [(a.shape[1], a.shape[0]) for a in imgs]
    roi = O.result_roi(corners, sizes)
    for dtype in (np.uint8, np.int16):
        fb = O.FeatherBlender(0.1)
        fb.prepare(roi)
        src = [a.astype(dtype) for a in imgs]
        for i in range(n):
            fb.feed(src[i].astype(np.int16), masks[i], corners[i])
        want, wmask = fb.blend()
        dst = np.zeros((roi[3], roi[2], 3), np.int16)
        dmask = np.zeros((roi[3], roi[2]), np.uint8)
        ip = (C.c_void_p * n)(*[a.ctypes.data for a in src])
        mp = (C.c_void_p * n)(*[m.ctypes.data for m in masks])
        rows = np.asarray([a.shape[0] for a in src], np.int32)
        cols = np.asarray([a.shape[1] for a in src], np.int32)
        x0 = np.asarray([c[0] - roi[0] for c in corners], np.int32)
        y0 = np.asarray([c[1] - roi[1] for c in corners], np.int32)
        emu_feather.emu_feather_blend(n, ip, 1 if dtype == np.uint8 else 0, mp, _p(rows), _p(cols), _p(x0), _p(y0), C.c_float(0.1), roi[2], roi[3],
                                      _p(dst), _p(dmask))
        assert np.array_equal(dmask, wmask) and np.array_equal(dst, want), dtype


@pytest.mark.parametrize("case", [("synth", (1, 1), 3), ("synth", (3, 1), 3), ("noise", (2, 2), 1), ("gray", (1, 1), 1), ("bgra", (3, 1), 4)])
def test_orb_kernels_and_driver_match_oracle(emu_orb, oracle, case):
    """is_orb_find's whole computation -- gray conversion, pyramid, FAST + non-maximum suppression, Harris, orientation, blur,
    descriptors and the host selections between them -- from the product's source on the host == the oracle (== cv2.ORB): every field
    of every key point, their order, every descriptor byte"""
    from imagestitch_b200 import synth
    O = oracle
    kind, grid, ch = case
    rng = np.random.default_rng(21)
    if kind == "noise":
        img = rng.integers(0, 256, (260, 330), dtype=np.uint8)
    else:
        img = synth.make_panorama_inputs(2, 480, 300, 1.2, 0.25)[0][1]
        if kind == "gray":
            img = np.ascontiguousarray(img[:, :, 1])
        if kind == "bgra":
            img = np.ascontiguousarray(np.concatenate([img, rng.integers(0, 256, img.shape[:2] + (1,), dtype=np.uint8)], axis=2))
    want_k, want_d = O.orb_find(img, grid)
    cap = 1200 * grid[0] * grid[1]
    kps = np.zeros((cap, 6), np.float32)
    desc = np.zeros((cap, 32), np.uint8)
    n = emu_orb.emu_orb_find(_p(img), img.shape[0], img.shape[1], 1 if img.ndim == 2 else img.shape[2], C.c_size_t(img.strides[0]), grid[0], grid[1], 510,
                             C.c_float(1.3), 5, _p(kps), _p(desc), cap)
    assert n == len(want_k) and n > 50
    assert np.array_equal(kps[:n].view(np.uint32), want_k.view(np.uint32))
    assert np.array_equal(desc[:n], want_d)


@pytest.mark.parametrize("case", [(100, 150, (3, 1)), (70, 210, (3, 1)), (200, 90, (1, 2)), (333, 517, (2, 3)), (129, 1000, (3, 1))])
def test_orb_edge_sizes_match_oracle(emu_orb, oracle, case):
    """the product's ORB source on cells / levels around the 2 x 31 pixel border, odd sizes, grids with rows, tie-heavy images"""
    O = oracle
    h, w, grid = case
    rng = np.random.default_rng(9)
    noise = rng.integers(0, 256, (h, w), dtype=np.uint8)
    blocks = np.ascontiguousarray(np.kron(rng.integers(0, 256, (h // 3 + 1, w // 3 + 1), dtype=np.uint8), np.ones((3, 3), np.uint8))[:h, :w])
    for img in (noise, blocks):
        want_k, want_d = O.orb_find(img, grid)
        cap = 1200 * grid[0] * grid[1]
        kps = np.zeros((cap, 6), np.float32)
        desc = np.zeros((cap, 32), np.uint8)
        n = emu_orb.emu_orb_find(_p(img), h, w, 1, C.c_size_t(img.strides[0]), grid[0], grid[1], 510, C.c_float(1.3), 5, _p(kps), _p(desc), cap)
        assert n == len(want_k)
        assert np.array_equal(kps[:n].view(np.uint32), want_k.view(np.uint32)) and np.array_equal(desc[:n], want_d)
