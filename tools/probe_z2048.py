# z-pass timing on a fine mesh along z (scratch): python tools/probe_z2048.py [N1 N2 N3]
# fused radix-2-wrapped persistent kernel (zfused=2) vs cuFFT z forward + Green kernel + cuFFT z inverse (zfused=0)
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noname_b200 import DevicePM

dims = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 512, 2048)
rng = np.random.default_rng(1)
x = rng.random((64, 3))
v = np.zeros((64, 3))
for zf in (2, 0, 2, 0):
    with DevicePM(dims, 64, rank=3, timing=True) as dev:
        dev.set_option("zfused", zf)
        dev.upload(x, v, 1.0)
        ms = []
        for it in range(4):
            st = dev.step_static(1e-6)
            ms.append(st.as_dict()["ms"])
        print(dims, "zfused", zf, "solve %.2f fft_z %.2f fft_xy %.2f" % (ms[-1]["solve"], ms[-1]["fft_z"], ms[-1]["fft_xy"]), flush=True)
