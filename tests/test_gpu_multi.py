"""Multi-GPU tests (-m gpu): skipped on boxes with fewer than 2 GPUs.  The checks themselves run under torchrun
(tests/dist_check_native.py): native NCCL setup, NVLink peer-memory exchange, distributed cyclic reduction, graph replay,
against the single-GPU solve of the same scene."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_gpu_distributed_solve_matches_single_gpu():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_check_native.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DIST_CHECK_NATIVE PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
