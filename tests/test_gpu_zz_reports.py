"""COLOR_GRAD seam cost on the device.  The kernels were written after the round's GPU time was
spent, so the path sits behind IS_EXPERIMENTAL_COLOR_GRAD=1 and this test reports instead of gating: it runs
tests/tools/check_color_grad.py in a process of its own (a fault there cannot poison this session's CUDA context) and turns
a failure into an expected-failure record with the script's output.  Once it has passed on hardware the switch goes and
the check moves into tests/test_gpu_parity.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_color_grad_device_matches_oracle():
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "check_color_grad.py")], capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        pytest.xfail("tests/tools/check_color_grad.py timed out (experimental path)")
    tail = (r.stdout + r.stderr)[-1500:]
    print(tail)
    if r.returncode != 0:
        pytest.xfail("experimental COLOR_GRAD device path does not match the oracle yet:\n" + tail)


def test_seam_edge_cases_device_report():
    """tests/helpers.seam_edge_cases (pinned against the reference's own find() on the CPU side) on the device.  Written
    after the round's GPU time was spent: reports, in a process of its own, and records a mismatch as an expected failure
    with the script's output; becomes a hard parity test once it has been seen to pass on hardware."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "check_seam_edge_cases.py")], capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        pytest.xfail("tests/tools/check_seam_edge_cases.py timed out")
    tail = (r.stdout + r.stderr)[-2500:]
    print(tail)
    if r.returncode != 0:
        pytest.xfail("device seam finder differs from the oracle on an edge case (COLOR path):\n" + tail)
