import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200.workloads import config_B, build_kernel
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
n = 16384
cfg = config_B(n)
for algo in (0, 1):
    k = build_kernel(cfg); k.esp_algorithm = algo
    spec = k._make_spec()
    Xd = _device.to_device(cfg["X"])
    px = _device.Points(spec, Xd); px2 = _device.Points(spec, Xd)
    out = torch.empty((n, n), dtype=torch.float64, device="cuda")
    t = timeit(lambda: _device.gram(spec, px, px2, out=out))
    ts = timeit(lambda: _device.gram(spec, px, out=out))
    print(f"lib={os.environ.get('OAK_B200_LIB','default')} variant={os.environ.get('OAK_GRAM_VARIANT')} algo={algo}: cross {t:.3f} ms ({n*n/t*1e-6:.2f} G/s)  sym {ts:.3f} ms ({n*(n+1)/2/ts*1e-6:.2f} G/s)")
    spec.close()
