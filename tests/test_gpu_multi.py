"""Slab-decomposed multi-GPU parity (P = 2, 4, 8 as far as the box has GPUs): P-rank results equal
the oracle / the single-GPU result.  Spawns tests/mgpu_worker.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("P", [2, 4, 8])
def test_multi_gpu_matches_oracle(lib_built, P):
    if _ngpu() < P:
        pytest.skip(f"needs {P} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={P}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + P), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "MGPU_OK" in out, out[-4000:]
