#!/usr/bin/env python
"""Markdown table of the bench lines kept under profiles/ (bench_r02_*.json): python profiles/bench_table.py"""
import glob
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
rows = []
for f in sorted(glob.glob(os.path.join(HERE, "bench_r02_*.json"))):
    try:
        p = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    r = p["roofline"]
    ph = r.get("phase_ms", {})
    par = p.get("parity") or {}
    rows.append((os.path.basename(f)[10:-5], p["n_gpus"], p["steps"], p["ms_per_step"], p["value"], r["step_frac"],
                 r.get("step_frac_at_measured_nvlink_770"), ph, par, (p.get("clocks") or {}).get("sm_mhz")))
print("| line | GPUs | steps | ms/step | pu/s | of roofline (900 GB/s NVLink) | (770) | deposit | solve (comm) | push | sort | exchange | SM MHz | parity |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, n, k, ms, val, fr, fr2, ph, par, mhz in rows:
    pk = par.get("pk_vs_P1_maxrel")
    ptxt = "—" if not par else ("ok" if par.get("ok") else "FAIL") + (f", P(k) {pk:.1e}" if pk is not None else "")
    print(f"| {name} | {n} | {k} | {ms:.2f} | {val:.3g} | {fr:.3f} | {'' if fr2 is None else f'{fr2:.3f}'} | {ph.get('deposit', 0):.2f} | "
          f"{ph.get('solve', 0):.2f} ({ph.get('comm', 0):.2f}) | {ph.get('push', 0):.2f} | {ph.get('sort', 0):.2f} | {ph.get('exchange', 0):.2f} | {mhz} | {ptxt} |")
