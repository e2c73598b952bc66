"""Multi-GPU parity (skipped below 2 devices): the column-strip sharded stitcher over NCCL on 2 GPUs -- a strip and a 2-row mosaic --
assembled on rank 0 and compared bit for bit with the single-GPU pipeline (which tests/test_gpu_parity.py compares with the oracle):
seam masks of every image, panorama, panorama mask.  The check itself is scripts/sharded_check.py, launched through torchrun."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,per_rank,grid_rows", [(600, 900, 3, 1), (480, 640, 4, 2)])
def test_sharded_equals_single_gpu(rows, cols, per_rank, grid_rows):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "scripts", "sharded_check.py"), str(rows), str(cols), str(per_rank), str(grid_rows)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "SHARDED == SINGLE GPU: True" in out, out[-3000:]
    assert "differs" not in out, out[-3000:]
