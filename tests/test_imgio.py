"""Bitmap files either side of the path (imagestitch_b200/csrc/imgio.cu; cv::imread [BLEND]:31-34, cv::imwrite [BLEND]:717,
[SEAM]:1195-1206) against cv2.imread / cv2.imwrite, byte for byte: without a GPU the marked region of the .cu file (header parse
and write, the unpack / pack kernels) runs on the host emulator; with one, the same cases go through the C ABI."""
import ctypes as C
import glob
import os
import re
import struct
import subprocess

import numpy as np
import pytest

try:
    import cv2
except ImportError:                     # the golden test below needs neither cv2 nor a GPU
    cv2 = None
needs_cv2 = pytest.mark.skipif(cv2 is None, reason="cv2 not importable")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
OUT = os.path.join(EMU, "_build")
IS_8U, IS_16S, IS_32F = 0, 3, 5


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def emu_io():
    text = open(os.path.join(ROOT, "imagestitch_b200", "csrc", "imgio.cu")).read()
    regions = re.findall(r"// @emu-begin[^\n]*\n(.*?)// @emu-end", text, flags=re.S)
    assert len(regions) == 1
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "imgio_region.inc"), "w") as f:
        f.write(regions[0])
    so = os.path.join(OUT, "libimgio_emul.so")
    subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-I", OUT, "-shared", "-o", so,
                           os.path.join(EMU, "imgio_emul.cpp")])
    lib = C.CDLL(so)
    lib.emu_bmp_write.restype = C.c_size_t
    return lib


def _palette_bmp(idx, bpp, palette, top_down=False, clr_used=0, core=False):
    """an uncompressed palettized bitmap built by hand (cv2 writes none with 1 / 4 bits or a colour palette)"""
    h, w = idx.shape
    step = ((w * bpp + 7) // 8 + 3) & ~3
    rows = []
    for y in (range(h) if top_down else range(h - 1, -1, -1)):
        r = bytearray(step)
        for x in range(w):
            v = int(idx[y, x])
            if bpp == 8:
                r[x] = v
            elif bpp == 4:
                r[x >> 1] |= v << (0 if x & 1 else 4)
            else:
                r[x >> 3] |= v << (7 - (x & 7))
        rows.append(bytes(r))
    n = clr_used or (1 << bpp)
    if core:
        pal = b"".join(bytes(palette[i][:3]) for i in range(n))
        head = struct.pack("<IHHHH", 12, w, h, 1, bpp)
    else:
        pal = b"".join(bytes(palette[i][:3]) + b"\0" for i in range(n))
        head = struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, bpp, 0, 0, 0, 0, clr_used, 0)
    off = 14 + len(head) + len(pal)
    body = b"".join(rows)
    return b"BM" + struct.pack("<IHHI", off + len(body), 0, 0, off) + head + pal + body


def _rgb24_bmp(img, top_down=False):
    h, w = img.shape[:2]
    pad = b"\0" * ((-3 * w) % 4)
    body = b"".join(img[y].tobytes() + pad for y in (range(h) if top_down else range(h - 1, -1, -1)))
    head = struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, 24, 0, 0, 0, 0, 0, 0)
    return b"BM" + struct.pack("<IHHI", 54 + len(body), 0, 0, 54) + head + body


def _rgb32_bmp(img, bitfields=False, top_down=False):
    h, w = img.shape[:2]
    body = b"".join(np.concatenate([img[y], np.full((w, 1), 77, np.uint8)], axis=1).tobytes() for y in (range(h) if top_down else range(h - 1, -1, -1)))
    head = struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, 32, 3 if bitfields else 0, 0, 0, 0, 0, 0)
    masks = struct.pack("<III", 0x00ff0000, 0x0000ff00, 0x000000ff) if bitfields else b""
    off = 14 + len(head) + len(masks)
    return b"BM" + struct.pack("<IHHI", off + len(body), 0, 0, off) + head + masks + body


def _files(tmp_path, with_reference=True):
    """name -> path of bitmaps of every supported layout"""
    rng = np.random.default_rng(3)
    out = {}
    for name, shape in (("c3_w7", (5, 7, 3)), ("c3_w8", (9, 8, 3)), ("c3_big", (130, 301, 3)), ("c1_w5", (6, 5)), ("c1_w256", (3, 256))):
        p = str(tmp_path / f"{name}.bmp")
        assert cv2.imwrite(p, rng.integers(0, 256, shape, dtype=np.uint8))
        out[name] = p
    pal = rng.integers(0, 256, (256, 3), dtype=np.uint8).tolist()
    for name, data in (("p8", _palette_bmp(rng.integers(0, 256, (7, 13)), 8, pal)),
                       ("p8_topdown", _palette_bmp(rng.integers(0, 256, (7, 13)), 8, pal, top_down=True)),
                       ("p8_clrused", _palette_bmp(rng.integers(0, 40, (6, 9)), 8, pal, clr_used=20)),
                       ("p4", _palette_bmp(rng.integers(0, 16, (5, 11)), 4, pal)),
                       ("p1", _palette_bmp(rng.integers(0, 2, (4, 19)), 1, pal)),
                       ("p8_core", _palette_bmp(rng.integers(0, 256, (5, 6)), 8, pal, core=True)),
                       ("p4_odd", _palette_bmp(rng.integers(0, 16, (3, 7)), 4, pal, top_down=True)),
                       ("p1_w33", _palette_bmp(rng.integers(0, 2, (5, 33)), 1, pal)),
                       ("rgb24_topdown", _rgb24_bmp(rng.integers(0, 256, (7, 6, 3), dtype=np.uint8), top_down=True)),
                       ("rgb24_w1", _rgb24_bmp(rng.integers(0, 256, (3, 1, 3), dtype=np.uint8))),
                       ("rgb32_topdown", _rgb32_bmp(rng.integers(0, 256, (4, 3, 3), dtype=np.uint8), top_down=True)),
                       ("rgb32", _rgb32_bmp(rng.integers(0, 256, (6, 5, 3), dtype=np.uint8))),
                       ("rgb32_bitfields", _rgb32_bmp(rng.integers(0, 256, (6, 5, 3), dtype=np.uint8), bitfields=True))):
        p = str(tmp_path / f"{name}.bmp")
        open(p, "wb").write(data)
        out[name] = p
    for p in (sorted(glob.glob("/root/reference/*/*/*.bmp")) if with_reference else []):          # the reference's own artefacts, where they are present
        out["ref:" + os.path.basename(p)] = p
    return out


def _write_cases():
    rng = np.random.default_rng(9)
    f = rng.uniform(-40, 300, (11, 13, 3)).astype(np.float32)
    f[0, 0] = (np.nan, np.inf, -np.inf)
    f[0, 1] = (0.5, 1.5, 2.5)
    f[0, 2] = (254.5, 255.5, 3e9)
    return [("u8c3", rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)), ("u8c1", rng.integers(0, 256, (5, 6), dtype=np.uint8)),
            ("u8c1_w4", rng.integers(0, 256, (3, 4), dtype=np.uint8)), ("s16c3", rng.integers(-300, 600, (8, 10, 3)).astype(np.int16)),
            ("s16c1", rng.integers(-300, 600, (8, 3)).astype(np.int16)), ("f32c3", f), ("f32c1", np.ascontiguousarray(f[:, :, 0]))]


@needs_cv2
def test_bmp_decode_matches_cv2(emu_io, tmp_path):
    for name, path in _files(tmp_path).items():
        want = cv2.imread(path)
        assert want is not None, name
        data = np.frombuffer(open(path, "rb").read(), np.uint8)
        rows, cols, bpp = C.c_int(), C.c_int(), C.c_int()
        assert emu_io.emu_bmp_info(_p(data), C.c_size_t(data.size), C.byref(rows), C.byref(cols), C.byref(bpp)) == 0, name
        assert (rows.value, cols.value) == want.shape[:2], name
        got = np.zeros((rows.value, cols.value, 3), np.uint8)
        assert emu_io.emu_bmp_read(_p(data), C.c_size_t(data.size), _p(got), C.c_size_t(got.strides[0])) == 0
        assert np.array_equal(got, want), name


def test_bmp_decode_rejects_what_it_does_not_read(emu_io, tmp_path):
    rows, cols, bpp = C.c_int(), C.c_int(), C.c_int()
    good = np.frombuffer(_palette_bmp(np.zeros((4, 4), np.int64), 8, [[0, 0, 0]] * 256), np.uint8)
    for bad in (good[:30], np.concatenate([np.frombuffer(b"XX", np.uint8), good[2:]]), good[:-5]):
        bad = np.ascontiguousarray(bad)
        assert emu_io.emu_bmp_info(_p(bad), C.c_size_t(bad.size), C.byref(rows), C.byref(cols), C.byref(bpp)) < 0
    rle = good.copy()
    rle[30] = 1                                                       # BI_RLE8
    assert emu_io.emu_bmp_info(_p(rle), C.c_size_t(rle.size), C.byref(rows), C.byref(cols), C.byref(bpp)) == -213


@needs_cv2
def test_bmp_encode_matches_cv2(emu_io, tmp_path):
    depth = {np.dtype(np.uint8): IS_8U, np.dtype(np.int16): IS_16S, np.dtype(np.float32): IS_32F}
    for name, img in _write_cases():
        p = str(tmp_path / f"{name}.bmp")
        assert cv2.imwrite(p, img)
        want = open(p, "rb").read()
        ch = 1 if img.ndim == 2 else img.shape[2]
        out = np.zeros(54 + 1024 + ((img.shape[1] * ch + 3) & ~3) * img.shape[0], np.uint8)
        n = emu_io.emu_bmp_write(_p(img), depth[img.dtype], img.shape[0], img.shape[1], ch, C.c_size_t(img.strides[0]), _p(out))
        assert out[:n].tobytes() == want, name


def test_bmp_golden_fixtures(emu_io):
    """committed cv2.imwrite bytes / cv2.imread pixels (tests/golden/widened_cases.npz, made by tests/golden/make_golden.py)"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "widened_cases.npz"))
    depth = {np.dtype(np.uint8): IS_8U, np.dtype(np.int16): IS_16S, np.dtype(np.float32): IS_32F}
    for j in range(int(z["n_bmp"])):
        img, want = np.ascontiguousarray(z[f"b{j}_img"]), z[f"b{j}_file_cv"]
        ch = 1 if img.ndim == 2 else img.shape[2]
        out = np.zeros(54 + 1024 + ((img.shape[1] * ch + 3) & ~3) * img.shape[0], np.uint8)
        n = emu_io.emu_bmp_write(_p(img), depth[img.dtype], img.shape[0], img.shape[1], ch, C.c_size_t(img.strides[0]), _p(out))
        assert np.array_equal(out[:n], want), j
        data = np.ascontiguousarray(want)
        got = np.zeros(z[f"b{j}_read_cv"].shape, np.uint8)
        assert emu_io.emu_bmp_read(_p(data), C.c_size_t(data.size), _p(got), C.c_size_t(got.strides[0])) == 0
        assert np.array_equal(got, z[f"b{j}_read_cv"]), j


@needs_cv2
@pytest.mark.gpu
def test_bmp_files_through_the_c_abi(ctx, tmp_path):
    import torch
    from imagestitch_b200 import stitching as S
    files = _files(tmp_path, with_reference=False)
    for name, path in files.items():
        want = cv2.imread(path)
        assert np.array_equal(S.imread(ctx, path), want), name
    big = files["c3_big"]
    dev = S.imread(ctx, big, like=torch.empty(1, device="cuda"))
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), cv2.imread(big))
    for name, img in _write_cases():
        p, q = str(tmp_path / f"cv_{name}.bmp"), str(tmp_path / f"is_{name}.bmp")
        assert cv2.imwrite(p, img)
        S.imwrite(ctx, q, img)
        assert open(q, "rb").read() == open(p, "rb").read(), name
    t = torch.from_numpy(_write_cases()[3][1]).cuda()
    S.imwrite(ctx, str(tmp_path / "dev.bmp"), t)
    assert open(str(tmp_path / "dev.bmp"), "rb").read() == open(str(tmp_path / "cv_s16c3.bmp"), "rb").read()
    with pytest.raises(Exception):
        S.imread(ctx, str(tmp_path / "missing.bmp"))
