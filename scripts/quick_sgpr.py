import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200.workloads import config_C, build_kernel
def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
n = 1_000_000
cfg = config_C(n)
k = build_kernel(cfg); spec = k._make_spec()
Xd, Zd, yd = _device.to_device(cfg["X"]), _device.to_device(cfg["Z"]), _device.to_device(cfg["y"])
pz, px = _device.Points(spec, Zd), _device.Points(spec, Xd)
peak = 18.37e12
ref = None
for chunk in (262144,):
    st = _device.sgpr_stats(spec, pz, px, yd, chunk=chunk)
    t = timeit(lambda: _device.sgpr_stats(spec, pz, px, yd, chunk=chunk))
    if ref is None: ref = st.clone()
    err = float((st - ref).abs().max() / ref.abs().max())
    print(f"mode={os.environ.get('OAK_SYRK_MODE')} chunk={chunk}: {t:.2f} ms per {n} -> {t*1e6/n:.1f} ms per 1M  frac(865.5)={1024*n*865.5/(t*1e-3)/peak:.3f} relerr_vs_first={err:.1e}")
