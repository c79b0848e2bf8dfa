"""N > 1 path on CPU: world_size 2 and 4 over gloo (SURVEY.md section 8e protocol)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("P", [2, 4])
def test_slab_protocol_matches_single_process_oracle(P):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={P}", "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + P), os.path.join(ROOT, "tests", "slab_gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "SLAB_GLOO_OK" in out, out[-4000:]
