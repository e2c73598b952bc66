"""include/imagestitch.hpp (the C++ mirror of the reference's RotationWarper / SeamFinder / Blender classes) compiled into a
reference-style main() (tests/cpp/hpp_stitch.cpp): the build + link against the C-ABI library on the CPU, the run against the
oracle on the GPU."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from imagestitch_b200 import build as B, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "hpp_stitch.cpp")


def _build(tmp_path, src=SRC, name="hpp_stitch"):
    lib = B.build()
    exe = str(tmp_path / name)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    libdir = os.path.dirname(lib)
    cmd = [cxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", libdir, "-l:" + os.path.basename(lib), "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_hpp_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)        # no arguments: usage error before any CUDA call
    assert r.returncode == 2


def test_hpp_features_and_io_compile_and_link(tmp_path):
    """OrbFeaturesFinder / ImageFeatures, imread / imwrite and the stand-alone remap of the header, instantiated"""
    exe = _build(tmp_path, os.path.join(ROOT, "tests", "cpp", "hpp_features_io.cpp"), "hpp_features_io")
    assert subprocess.run([exe], capture_output=True, text=True).returncode == 2


@pytest.mark.gpu
def test_hpp_main_equals_oracle(tmp_path, oracle):
    O = oracle
    n, h, w = 3, 240, 320
    imgs, Ks, Rs, scale = synth.make_panorama_inputs(n, w, h, 1.2, 0.3)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<iiif", n, h, w, float(scale)))
        for i in range(n):
            f.write(np.asarray(Ks[i], np.float32).tobytes())
            f.write(np.asarray(Rs[i], np.float32).tobytes())
        for i in range(n):
            f.write(np.ascontiguousarray(imgs[i], np.uint8).tobytes())
    exe = _build(tmp_path)
    r = subprocess.run([exe, str(inp), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = O.pipeline_run(O.PROJ_CYLINDRICAL, imgs, Ks, Rs, scale, seam=True, num_bands=3, weight_type=O.WEIGHT_32F, want_intermediates=True)
    raw = open(out, "rb").read()
    roi = struct.unpack_from("<4i", raw, 0)
    pos = 16
    assert tuple(roi) == tuple(int(v) for v in want["roi"])
    for i in range(n):
        x, y, sw, sh = struct.unpack_from("<4i", raw, pos)
        pos += 16
        assert (x, y) == tuple(int(v) for v in want["corners"][i]) and (sw, sh) == tuple(int(v) for v in want["sizes"][i])
        m = np.frombuffer(raw, np.uint8, sw * sh, pos).reshape(sh, sw)
        pos += sw * sh
        assert np.array_equal(m, want["masks"][i]), f"seam mask {i}"
    H, W = roi[3], roi[2]
    pano = np.frombuffer(raw, np.int16, H * W * 3, pos).reshape(H, W, 3)
    pos += H * W * 6
    pmask = np.frombuffer(raw, np.uint8, H * W, pos).reshape(H, W)
    assert np.array_equal(pmask, want["pano_mask"])
    assert np.array_equal(pano, want["pano"])
