"""Dry run of bench.py's NATIVE arm on CPU: torch.cuda and the device context are replaced by stand-ins (the context by
an adapter over the C restatement, so the in-run parity block really compares something), and main() runs end to end --
argument handling, warm-up / timed loop, roofline arithmetic, the parity block, the e2e loop, the CPU baseline and the
JSON line.  It guards the one script the driver's numbers come from against integration errors that only a GPU run would
otherwise show; it measures nothing."""
import json
import sys
import types

import numpy as np
import pytest

from oracle import pm_ref as O


class _Stats:
    def __init__(self, mabs, mv, n):
        self.max_abs_v, self.max_v, self.max_pa = mabs, mv, 0.0
        self.n_halo_violations, self.n_untiled, self.n_migrated, self.npart_local, self.kernel_launches = 0, 0, 0, n, 9

    def as_dict(self):
        from noname_b200._lib import MS_NAMES
        d = {k: getattr(self, k) for k in ("max_abs_v", "max_v", "max_pa", "n_halo_violations", "npart_local", "n_migrated",
                                           "n_untiled", "kernel_launches")}
        d["ms"] = {n: (1.0 if n in ("deposit", "solve", "push", "fft_z", "fft_xy", "fft_x", "fft_y") else 0.0) for n in MS_NAMES}
        return d


class _Dev:
    """the part of DevicePM bench.py uses, over oracle/pm_ref.c"""

    def __init__(self, dims, npart_total, rank=3, **kw):
        from oracle.pm_ref_c import PMRefC
        self.dims, self.rank, self.nranks, self.rank_id = tuple(dims), rank, 1, 0
        self.f = O.Fields(self.dims, npart_total, rank)
        self.c = PMRefC(threads=1)
        self.options = {}

    def set_option(self, k, v):
        self.options[k] = v

    def init_grf(self, pdims, seed=12345, ps_index=-1.0, disp_rms=0.0, vfac=0.0, m=1.0, peaks=None):
        self.x, self.v = O.synthetic_grf(tuple(pdims), self.rank, seed, ps_index, disp_rms, vfac, peaks=peaks)
        self.m = m

    def upload(self, x, v, m, first_id=0):
        self.x, self.v, self.m = np.array(x), np.array(v), m

    def download(self, x=None, v=None, ids=False):
        if x is None:
            x, v = np.empty_like(self.x), np.empty_like(self.v)
        x[...], v[...] = self.x, self.v
        return (x, v, np.arange(len(self.x))) if ids else (x, v)

    @property
    def local_count(self):
        return len(self.x)

    def step_comoving(self, dt, a1, ak, adk, a2):
        mabs, mv = self.c.step_comoving(self.f, self.x, self.v, self.m, dt, a1, ak, adk, a2)
        return _Stats(mabs, mv, len(self.x))

    def deposit(self):
        self.c.deposit(self.f.rho, self.x, self.m, "cic")

    def get_field(self, which):
        return {"rho": self.f.rho, "phi": self.f.phi}[which].copy()

    def power_spectrum(self):
        return O.powerSpectrum(self.x.copy(), self.m, self.dims, "cic")

    def close(self):
        pass


class _Event:
    clock = [0.0]

    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def test_native_arm_runs_end_to_end_on_stand_ins(monkeypatch, capsys, tmp_path):
    import torch
    import bench
    import noname_b200
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{**k, "device": "cpu"}))     # device="cuda" scratch tensors
    monkeypatch.setattr(noname_b200, "DevicePM", _Dev)
    monkeypatch.setattr(noname_b200, "pinned_empty", lambda shape, dtype=np.float64: np.empty(shape, dtype=dtype))
    monkeypatch.setattr(noname_b200, "pinned_free", lambda a: None)
    monkeypatch.setattr(bench, "PK_REF", str(tmp_path / "pk_ref.json"))
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)        # no nvidia-smi here
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)

    def run(*extra):
        monkeypatch.setattr(sys, "argv", ["bench.py", "--n", "16", "--mesh", "16", "--steps", "4", "--warmup", "3", "--cpu-sample", "16",
                                          "--e2e-steps", "1", *extra])
        assert bench.main() == 0
        return json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    first = run("--pk-ref-write")                                           # stores this run's P(k) as the reference ...
    line = run()                                                            # ... which the second run is compared with
    assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["n_gpus"] == 1 and line["steps"] == 4
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["dtype"] == "f64" and line["scaling"] == "strong"
    assert line["gpu_launches"] == 4 * 9 and line["vs_baseline"] is None
    cfg = line["config"]
    assert cfg["particles"] == 16 ** 3 and cfg["mesh_cells"] == 16 ** 3 and "16^3" in cfg["workload"]
    assert cfg["cosmology"]["hubble_time_gyr"] == 9.77793 and cfg["cosmology"]["zinit"] == 50.0
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["achieved"] > 0 and roof["frac"] == pytest.approx(roof["achieved"] / roof["peak"])
    assert roof["algorithmic_bytes_per_launch"] == 96.0 * 16 ** 3 and set(roof["per_kernel"]) >= {"deposit", "push", "fft_z"}
    assert roof["step_roofline_ms"] == pytest.approx((16 ** 3 * (144.0 + 96.0)) / (roof["peak"] * 1e9) * 1e3)
    par = line["parity"]
    assert par["small_step"]["ok"] and par["mass_rel"] < 1e-12 and par["n_halo_violations"] == 0
    assert par["pk_vs_P1_maxrel"] is not None and par["pk_vs_P1_maxrel"] < 1e-12 and par["ok"] is True
    assert first["parity"]["ok"] is True
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 2 * 16 ** 3 * 24 and e2e["d2h_bytes_per_step"] == 2 * 16 ** 3 * 24 + 32
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] > 0
    assert line["clocks"]["reasons"] == ["nvidia-smi unavailable"]
