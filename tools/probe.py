# quick perf probe (scratch; not part of the product)
import sys, time
import numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from noname_b200 import DevicePM
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
opts = dict(kv.split('=') for kv in sys.argv[2:])
dev = DevicePM((n, n, n), n ** 3, rank=3, timing=True, part_capacity=int(opts.pop('cap_extra', 0)) + n ** 3, sort_every=int(opts.pop('sort_every', 0)),
               deposit_mode=opts.pop('deposit_mode', 'atomic'))
for k, v in opts.items():
    dev.set_option(k, float(v))
dev.init_synthetic((n, n, n), seed=1, nmodes=32, kmax=8, disp_rms=0.3 / n, vfac=0.05, m=0.3)
print('device GB', dev.device_bytes / 1e9)
for it in range(6):
    t = time.time()
    st = dev.step_comoving(0.2 / n, 1.0, 1.0, 0.8, 1.0)
    el = (time.time() - t) * 1e3
    d = st.as_dict()
    print(it, 'wall ms %.2f' % el, {k: round(v, 3) for k, v in d['ms'].items()}, 'maxv %.4g' % d['max_abs_v'], 'launches', d['kernel_launches'])
print('pu/s %.3e' % (n ** 3 / (el * 1e-3)))
