#!/usr/bin/env python
"""Per-kernel share of a step from an ncu launch list (--metrics gpu__time_duration.sum --csv):
    python profiles/launch_shares.py gpurun_out/launches.csv [first_kernel_of_a_step]
Steps are delimited by the step's first launch (default k_reset_red); steps that contain a sort are
left out; the mean over the remaining steps is printed (ms per step and share)."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    first = sys.argv[2] if len(sys.argv) > 2 else "k_reset_red"
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    steps, cur = [], None
    for r in rows[1:]:
        name = r[kn].split("(")[0]
        if name.startswith("void "):
            name = name[5:]
        if name.startswith(first):
            cur = []
            steps.append(cur)
        if cur is not None:
            cur.append((name, float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)))
    steps = [s for s in steps[:-1] if not any(n.startswith("k_sort") or n.startswith("k_permute") for n, _ in s)]
    tot = collections.defaultdict(float)
    for s in steps:
        for n, t in s:
            tot[n[:70]] += t / len(steps)
    print(f"# mean over {len(steps)} sort-free steps; ms per step and share (cold-cache, serialised launches: compare SHARES)")
    total = sum(tot.values())
    for n, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{n:72s} {t:8.3f} ms {100 * t / total:5.1f}%")
    print(f"{'sum':72s} {total:8.3f} ms")


if __name__ == "__main__":
    main()
