# perf probe over kernel variants on ONE device context (scratch; not part of the product):
#   python tools/probe_variants.py 1024 base "zfused=2" "rowfft=0" "rowfft=0,zfused=2"
# Each argument after the size is one variant (comma-separated option=value pairs applied on top of
# the defaults; "base" = defaults).  Prints mean deposit / solve / push / wall ms over the last 4 of 5 steps.
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noname_b200 import DevicePM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variants = sys.argv[2:] or ["base"]
ic = os.environ.get("PROBE_IC", "grf")
dev = DevicePM((n, n, n), n ** 3, rank=3, timing=True, part_capacity=n ** 3, sort_every=0)
if ic == "clustered":
    dev.init_synthetic((n, n, n), kind="clustered", seed=777, nmodes=100, kmax=1000, disp_rms=1.5 / n, vfac=2e-5, m=0.3)
elif ic == "modes":
    dev.init_synthetic((n, n, n), seed=1, nmodes=32, kmax=8, disp_rms=0.3 / n, vfac=0.05, m=0.3)
else:
    dev.init_grf((n, n, n), seed=12345, ps_index=-1.0, disp_rms=0.3 / n, vfac=0.05, m=0.3)
dev.sort()
print("device GB", dev.device_bytes / 1e9, "ic", ic, flush=True)
DEFAULTS = {"rowfft": 1, "colfft": 1, "zfused": 3, "fft3d": 0, "tile_deposit": 1, "tile_push": 1, "bulk_flush": 1,
            "l2pf_push": 1, "l2pf_deposit": 1, "cfrun_y": 1, "zsplit": 0, "sort_cells": 1, "fft2d": 0, "fft2d_block": 4,
            "keep_rho": 0}
known = {}
for var in variants:
    opts = {} if var == "base" else dict(kv.split("=") for kv in var.split(","))
    for k in known:                       # back to the defaults recorded so far
        dev.set_option(k, known[k])
    for k, v in opts.items():
        known.setdefault(k, float(DEFAULTS[k]))
        dev.set_option(k, float(v))
    acc = {}
    walls = []
    nsteps = int(os.environ.get('PROBE_STEPS', 5))
    for it in range(nsteps):
        t = time.time()
        st = dev.step_comoving(0.2 / n, 1.0, 1.0, 0.8, 1.0)
        walls.append((time.time() - t) * 1e3)
        if it:
            for k, v in st.as_dict()["ms"].items():
                acc[k] = acc.get(k, 0.0) + v / (nsteps - 1)
    d = st.as_dict()
    print("%-32s dep %.2f solve %.2f (x %.2f y %.2f z %.2f) push %.2f wall %.2f | untiled %d maxv %.9g maxpa %.9g" % (
        var, acc["deposit"], acc["solve"], acc["fft_x"], acc["fft_y"], acc["fft_z"], acc["push"], sum(walls[1:]) / (nsteps - 1),
        d["n_untiled"], d["max_abs_v"], d["max_pa"]), flush=True)
