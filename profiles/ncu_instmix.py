#!/usr/bin/env python
"""Instruction mix / hottest SASS lines of one kernel of an .ncu-rep (needs --import-source on).
   python profiles/ncu_instmix.py rep.ncu-rep <kernel-substring> [--lines N]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{pat}", "-c", "1"] if False else
                         ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # the dump concatenates kernels: "Kernel Name",<name> line, then header, then instructions
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        if pat not in b["name"]:
            continue
        h = b["hdr"]
        iS, iE, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        mix, stall, tot = collections.Counter(), collections.Counter(), 0
        per = []
        for r in b["rows"]:
            if len(r) <= iE:
                continue
            op = [o for o in r[iS].split() if not o.startswith("@")]
            if not op:
                continue
            nm = op[0].split(".")[0]
            n = int(r[iE] or 0)
            w = int(r[iW] or 0)
            mix[nm] += n
            stall[nm] += w
            tot += n
            per.append((w, n, r[iS].strip()))
        print(b["name"][:100])
        print("  total warp instructions", tot)
        for k, v in mix.most_common(22):
            print(f"  {k:10s} {v:12d} {100 * v / tot:5.1f}%   stall samples {stall[k]}")
        if nlines:
            print("  hottest lines (stall samples, executed, sass):")
            for w, n, sline in sorted(per, reverse=True)[:nlines]:
                print(f"   {w:7d} {n:11d}  {sline}")
        break


if __name__ == "__main__":
    main()
