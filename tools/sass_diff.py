#!/usr/bin/env python
"""Per-kernel SASS comparison of two builds of libnnpm.so (no GPU needed).

    python tools/sass_diff.py OLD.so [NEW.so]        # NEW defaults to noname_b200/libnnpm.so
    python tools/sass_diff.py --rev HEAD~3           # builds that revision's csrc/ into a temp dir first

Prints the kernels present in only one library and the kernels whose instruction stream differs (addresses and
encodings stripped).  Used to prove that a change leaves every kernel the GPU tests and the bench were measured
with byte-identical when no GPU time is left to re-run them."""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    table, name, buf = {}, None, []

    def close():
        if name:
            table[name] = (hashlib.md5("\n".join(buf).encode()).hexdigest(), len(buf))

    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            close()
            name, buf = m.group(1), []
        elif name:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
            if m:
                buf.append(re.sub(r"\s+", " ", m.group(1)))
    close()
    return table


def build_rev(rev):
    d = tempfile.mkdtemp(prefix="nnpm_rev_")
    tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "noname_b200/csrc", "include"], capture_output=True, check=True).stdout
    subprocess.run(["tar", "-x", "-C", d], input=tar, check=True)
    out = os.path.join(d, "libnnpm.so")
    subprocess.run(["make", "-s", "-C", os.path.join(d, "noname_b200", "csrc"), f"OUT={out}"], check=True)
    return out


def main(argv):
    if len(argv) >= 2 and argv[0] == "--rev":
        old, new = build_rev(argv[1]), (argv[2] if len(argv) > 2 else os.path.join(ROOT, "noname_b200", "libnnpm.so"))
    elif argv:
        old, new = argv[0], (argv[1] if len(argv) > 1 else os.path.join(ROOT, "noname_b200", "libnnpm.so"))
    else:
        print(__doc__)
        return 2
    a, b = kernels(old), kernels(new)
    only_a, only_b = sorted(set(a) - set(b)), sorted(set(b) - set(a))
    changed = sorted(k for k in a if k in b and a[k] != b[k])
    print(f"{old}: {len(a)} kernels; {new}: {len(b)} kernels; identical: {len(a) - len(only_a) - len(changed)}")
    for k in only_a:
        print(f"  only in old: {k} ({a[k][1]} instructions)")
    for k in only_b:
        print(f"  only in new: {k} ({b[k][1]} instructions)")
    for k in changed:
        print(f"  changed:     {k} ({a[k][1]} -> {b[k][1]} instructions)")
    return 0 if not (changed or only_a or only_b) else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
