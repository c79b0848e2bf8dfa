#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): one line per captured launch with the metrics the
roofline discussion uses.   python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "ms", 1), ("dram__bytes_read.sum", "GB_rd", 1), ("dram__bytes_write.sum", "GB_wr", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%", 1),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu%", 1),
        ("lts__t_sector_hit_rate.pct", "l2hit%", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1),
        ("smsp__inst_executed.sum", "winst", 1)]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0][:48]
        parts = []
        for k, lab, _ in KEYS:
            if k in d and d[k] != "":
                v = d[k]
                try:
                    v = float(v.replace(",", ""))
                    if lab == "ms":
                        v = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1, "ms": 1, "second": 1e3, "nsecond": 1e-6}.get(u[k], 1) * v
                    if lab.startswith("GB"):
                        v = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1, "Tbyte": 1e3}.get(u[k], 1) * v
                    v = f"{v:.4g}"
                except ValueError:
                    pass
                parts.append(f"{lab}={v}")
        print(f"{name:48s} " + " ".join(parts))
        if "--stalls" in sys.argv:
            st = []
            for k, v in d.items():
                if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
                    try:
                        st.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            print("      stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))


if __name__ == "__main__":
    main()
