"""The oracle's ORB features finder (oracle/orb.cpp, restating [FEAT]:56-418, 727-1021) pinned to OpenCV: every OpenCV call it
restates on its own -- cvtColor, resize(INTER_LINEAR_EXACT), cv::FAST, fastAtan2, the float separable filter behind
GaussianBlur on a sub-matrix -- and the whole of find() against cv2.ORB_create(510, 1.3, 5).detectAndCompute per grid cell:
key-point fields bit for bit, in the same order, and every descriptor byte.

One soft edge, stated: the blur's float sums are fused multiply-adds in OpenCV's AVX2 / FMA3 dispatch (any current x86 host) and
separate operations in its SSE baseline; the two differ in about one rounded pixel in 10^5.  The oracle follows the fused form;
on a host without FMA3 the blur comparison below would have to run with the other one."""
import numpy as np
import pytest

from imagestitch_b200 import synth

cv2 = pytest.importorskip("cv2")


def _cv_find(img, grid):
    """find() [FEAT]:948-1021 with cv2's ORB per cell"""
    gray = img if img.ndim == 2 else cv2.cvtColor(img, cv2.COLOR_BGR2GRAY if img.shape[2] == 3 else cv2.COLOR_BGRA2GRAY)
    orb = cv2.ORB_create(510, 1.3, 5)
    h, w = gray.shape
    kps, ds = [], []
    for r in range(grid[1]):
        for c in range(grid[0]):
            xl, yl, xr, yr = c * w // grid[0], r * h // grid[1], (c + 1) * w // grid[0], (r + 1) * h // grid[1]
            k, d = orb.detectAndCompute(np.ascontiguousarray(gray[yl:yr, xl:xr]), None)
            kps += [(p.pt[0] + xl, p.pt[1] + yl, p.size, p.angle, p.response, p.octave) for p in k]
            if d is not None:
                ds.append(d)
    return np.array(kps, np.float32).reshape(-1, 6), (np.concatenate(ds) if ds else np.zeros((0, 32), np.uint8))


def test_orb_building_blocks_vs_cv2(oracle):
    O = oracle
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(gray, O.bgr2gray(img))
    for (w, h) in ((231, 154), (178, 118), (299, 199), (150, 100), (77, 51), (300, 200), (301, 203)):
        assert np.array_equal(cv2.resize(gray, (w, h), interpolation=cv2.INTER_LINEAR_EXACT), O.resize_linear_exact(gray, (w, h))), (w, h)
    smooth = cv2.cvtColor(synth.make_panorama_inputs(2, 640, 480, 1.2, 0.25)[0][0], cv2.COLOR_BGR2GRAY)
    for g in (gray, smooth):
        k = cv2.FastFeatureDetector_create(20, True).detect(g, None)
        want = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)
        assert np.array_equal(want, O.fast(g, 20))
    assert all(cv2.fastAtan2(float(y), float(x)) == O.fast_atan2(float(y), float(x)) for y, x in rng.integers(-100000, 100000, (3000, 2)))
    kx = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()
    for g in (gray, smooth):
        want = cv2.sepFilter2D(g, cv2.CV_8U, kx, kx, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(want[3:-3, 3:-3], O.gaussian7(g)[3:-3, 3:-3])
    big = rng.integers(0, 256, (700, 1100), dtype=np.uint8)          # ~10^6 pixels: enough to tell fused from separate steps
    assert np.array_equal(cv2.sepFilter2D(big, cv2.CV_8U, kx, kx, borderType=cv2.BORDER_REFLECT_101)[3:-3, 3:-3], O.gaussian7(big)[3:-3, 3:-3])


@pytest.mark.parametrize("case", [("synth", (1, 1)), ("synth", (3, 1)), ("noise", (1, 1)), ("noise", (3, 1)), ("noise", (2, 2)), ("gray", (3, 1)), ("bgra", (1, 1))])
def test_orb_find_vs_cv2(oracle, case):
    O = oracle
    kind, grid = case
    rng = np.random.default_rng(4)
    if kind == "noise":
        img = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    else:
        img = synth.make_panorama_inputs(2, 900, 600, 1.2, 0.25)[0][0]
        if kind == "gray":
            img = np.ascontiguousarray(img[:, :, 2])
        if kind == "bgra":
            img = np.ascontiguousarray(np.concatenate([img, rng.integers(0, 256, img.shape[:2] + (1,), dtype=np.uint8)], axis=2))
    want_k, want_d = _cv_find(img, grid)
    got_k, got_d = O.orb_find(img, grid)
    assert len(want_k) == len(got_k) and len(got_k) > 100
    assert np.array_equal(want_k.view(np.uint32), got_k.view(np.uint32)), "key points (x, y, size, angle, response, octave), in OpenCV's order"
    assert np.array_equal(want_d, got_d), f"{int(np.unpackbits(want_d ^ got_d).sum())} descriptor bits differ"


@pytest.mark.parametrize("case", [(100, 150, (3, 1)), (70, 210, (3, 1)), (63, 64, (1, 1)), (200, 90, (1, 2)), (333, 517, (2, 3)), (64, 400, (4, 1)), (129, 1000, (3, 1))])
def test_orb_edge_sizes_vs_cv2(oracle, case):
    """cells and pyramid levels around the 2 x 31 pixel border (none / some levels too small for a key point), odd sizes, grids with
    rows, and blocky images whose responses tie by the hundred (retainBest keeps ties; the order is nth_element's)"""
    O = oracle
    h, w, grid = case
    rng = np.random.default_rng(9)
    noise = rng.integers(0, 256, (h, w), dtype=np.uint8)
    blocks = np.ascontiguousarray(np.kron(rng.integers(0, 256, (h // 3 + 1, w // 3 + 1), dtype=np.uint8), np.ones((3, 3), np.uint8))[:h, :w])
    for img in (noise, blocks):
        want_k, want_d = _cv_find(img, grid)
        got_k, got_d = O.orb_find(img, grid)
        assert len(want_k) == len(got_k)
        assert np.array_equal(want_k.view(np.uint32), got_k.view(np.uint32)) and np.array_equal(want_d, got_d)
