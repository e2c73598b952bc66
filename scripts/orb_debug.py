"""Diagnosis helper: the intermediate buffers of is_orb_find (IS_ORB_DUMP) on the device against those of the host emulation of the
same source, stage by stage.  `python scripts/orb_debug.py emu` writes the emulation's files (no GPU needed); `... gpu` compares."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EMU_DIR = os.path.join(ROOT, "scripts", "probe", "orb_emu_dump")


def image():
    from imagestitch_b200 import synth
    return synth.make_panorama_inputs(2, 1200, 800, 1.2, 0.25)[0][1]


def main():
    img = image()
    if sys.argv[1] == "emu":
        emu, out = os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", "_build")
        text = open(os.path.join(ROOT, "imagestitch_b200", "csrc", "orb.cu")).read()
        os.makedirs(out, exist_ok=True)
        open(os.path.join(out, "orb_region.inc"), "w").write(re.findall(r"// @emu-begin[^\n]*\n(.*?)// @emu-end", text, flags=re.S)[0])
        so = os.path.join(out, "liborb_emul.so")
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-pthread", "-fno-fast-math", "-w", "-I", emu, "-I", out, "-I",
                               os.path.join(ROOT, "imagestitch_b200", "csrc"), "-shared", "-o", so, os.path.join(emu, "orb_emul.cpp")])
        os.environ["IS_ORB_DUMP"] = EMU_DIR
        os.makedirs(EMU_DIR, exist_ok=True)
        lib = C.CDLL(so)
        kps, desc = np.zeros((4000, 6), np.float32), np.zeros((4000, 32), np.uint8)
        n = lib.emu_orb_find(img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1], 3, C.c_size_t(img.strides[0]), 3, 1, 510, C.c_float(1.3), 5,
                             kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), 4000)
        np.save(os.path.join(EMU_DIR, "img.npy"), img)
        print("emu key points:", n)
        return
    from imagestitch_b200 import stitching as S
    assert np.array_equal(img, np.load(os.path.join(EMU_DIR, "img.npy"))), "input differs"
    d = "/tmp/orb_gpu_dump"
    os.makedirs(d, exist_ok=True)
    os.environ["IS_ORB_DUMP"] = d
    ctx = S.Context(0)
    k, _ = S.orb_find(ctx, img, (3, 1))
    print("gpu key points:", len(k))
    for name, dt in (("pyr", np.uint8), ("score", np.uint8), ("hbuf", np.uint32), ("blur", np.uint8), ("kp2n", np.int32), ("resp", np.uint32), ("kpn", np.int32), ("ang", np.uint32)):
        a = np.fromfile(os.path.join(d, name + ".bin"), dt)
        b = np.fromfile(os.path.join(EMU_DIR, name + ".bin"), dt)
        if a.shape != b.shape:
            print(name, "sizes differ", a.shape, b.shape)
            continue
        bad = np.flatnonzero(a != b)
        print(name, "equal" if len(bad) == 0 else f"{len(bad)} of {a.size} differ, first at {bad[:8].tolist()}: gpu {a[bad[:8]].tolist()} emu {b[bad[:8]].tolist()}")


if __name__ == "__main__":
    main()
