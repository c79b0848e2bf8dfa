#!/usr/bin/env python
"""Where a kernel's warp-stall samples fall, from the SASS source page of an .ncu-rep (read on the CPU box):
    python profiles/ncu_hotspots.py rep.ncu-rep <kernel regex> [top N]
prints the N instructions with the most samples (with their dominant stall reason) and a coarse profile
along the instruction stream (samples per 5 % of the SASS)."""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # the page repeats per captured launch: take the first kernel block
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        body.append(dict(zip(hdr, r)))
    stall_cols = [c for c in hdr if c.startswith("stall_") and "(" not in c]
    tot = sum(int(b["# Samples"] or 0) for b in body)
    print(f"{len(body)} SASS instructions, {tot} samples")
    ranked = sorted(enumerate(body), key=lambda t: -int(t[1]["# Samples"] or 0))[:top]
    for i, b in ranked:
        s = int(b["# Samples"] or 0)
        dom = max(stall_cols, key=lambda c: int(b[c] or 0))
        print(f"  #{i:5d} {100.0 * s / tot:5.1f}%  {dom[6:]:14s} {b['Source'].strip()[:90]}")
    nb = 20
    print("profile along the instruction stream (% of samples per 5 % of the SASS):")
    line = []
    for k in range(nb):
        lo, hi = k * len(body) // nb, (k + 1) * len(body) // nb
        line.append(f"{100.0 * sum(int(b['# Samples'] or 0) for b in body[lo:hi]) / tot:.0f}")
    print("  " + " ".join(line))
    agg = {c: sum(int(b[c] or 0) for b in body) for c in stall_cols}
    print("stall totals: " + ", ".join(f"{c[6:]}={100.0 * v / tot:.0f}%" for c, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))


if __name__ == "__main__":
    main()
