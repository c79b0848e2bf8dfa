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


def test_host_ngpus_key_drives_a_slab_decomposed_run_world2():
    """conf key NGPUs = 2 through evolvePMCosmology (per-rank upload, redistribute, reduced stats, gathers at outputs,
    rank-0 files) with the device replaced by the CPU slab protocol: equals the single-process oracle loop"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "host_ngpus_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "HOST_NGPUS_OK" in out, out[-4000:]
