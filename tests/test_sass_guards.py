"""Guards on the generated SASS, read with cuobjdump (skipped where the CUDA toolkit is not installed).

k_orb_fast: ptxas 12.9 for sm_100a miscompiled the corner score written as max over arcs of max(min d, -(max d)) -- the chain of
maxima became VIMNMX3 and the negation survived for the first arc only (seen on a B200).  The kernel now takes
minima over d and over p - v, so no integer negation may feed its min / max network; this test fails if one comes back."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sass(obj, kernel):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", obj], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", out)
    body = [p for p in parts if p.startswith(kernel) or kernel in p.split("\n", 1)[0]]
    assert len(body) == 1, f"{kernel}: {len(body)} SASS functions"
    return body[0]


def test_fast_score_has_no_negation_in_its_min_max_network():
    from imagestitch_b200 import build as B
    B.build()
    sass = _sass(os.path.join(ROOT, "imagestitch_b200", "_build", "orb.o"), "_ZN2is10k_orb_fastE")
    assert sass.count("VIMNMX") > 50, "the score is a min / max network"
    neg = re.findall(r"IMAD\.MOV R\d+, RZ, RZ, -R\d+|IADD3 R\d+, PT, PT, -R\d+, RZ, RZ", sass)
    assert not neg, f"integer negations next to VIMNMX3 in k_orb_fast: {neg}"
