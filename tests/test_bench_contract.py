"""bench.py's reference arm runs on the host cores (no GPU): check the JSON line it prints against the contract the
driver reads -- same metric / unit / direction as the native arm, the sample it actually ran stated in `config`,
cpu_baseline and e2e describing this very run -- and that the native arm's line carries the keys the contract names."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_line():
    import bench
    line = json.loads(_run("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "3", "--cpu-sample", "32", "--cpu-threads", "2"))
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 3 and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["dtype"] == "f64" and line["data"] == "synthetic"
    # the line describes what was RUN: the bounded sample, with the problem it stands in for named beside it
    cfg = line["config"]
    assert cfg["particles"] == 32 ** 3 and cfg["mesh_cells"] == 32 ** 3 and "32^3" in cfg["workload"]
    assert cfg["sample_of"] == {"particles": 1024 ** 3, "mesh_cells": 1024 ** 3}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] == line["value"] and "32^3" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    assert _run("--impl", "reference", "--gpus", "2", "--cpu-sample", "16", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""


def test_native_arm_line_carries_the_contract_keys():
    """static: the dict literal rank 0 prints names every key of the contract (the GPU run itself is the driver's)"""
    src = open(os.path.join(ROOT, "bench.py"), encoding="utf-8").read()
    m = re.search(r"line = \{\"metric\": METRIC.*?\n\s+print\(json\.dumps\(line\)\)", src, flags=re.S)
    assert m, "native arm's JSON line not found"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "parity", "cpu_baseline"):
        assert f'"{key}"' in m.group(0), key
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "per_kernel"):
        assert f'"{key}"' in src[src.index("roofline = {"):src.index("t_hbm =")], key
    # the oracle is imported only inside the checker / cpu-baseline functions, never at module level or in the timed loop
    top = src[:src.index("def peaks()")]
    assert "oracle" not in re.sub(r'""".*?"""', "", top, flags=re.S)
