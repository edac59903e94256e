"""Quick device-resident timing of the fused Gram kernel (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200.workloads import config_B, config_C, build_kernel

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

peak = _device.measure_fp64_peak(1.0)
print(f"fp64 peak slots/s {peak:.4e}  ({2*peak/1e12:.2f} TFLOP/s)")
for algo in (0, 1):
    for n in (8192, 16384, 32768):
        cfg = config_B(n)
        k = build_kernel(cfg); k.esp_algorithm = algo
        spec = k._make_spec()
        Xd = _device.to_device(cfg["X"])
        px = _device.Points(spec, Xd)
        out = torch.empty((n, n), dtype=torch.float64, device="cuda")
        tmin, tmed = timeit(lambda: _device.gram(spec, px, out=out))
        uniq = n * (n + 1) / 2
        print(f"algo {algo} B sym   n={n}: {tmed:.3f} ms  unique entries/s {uniq/tmed*1e3:.3e}  frac(306 slots) {uniq*306/(tmed*1e-3)/peak:.3f}")
        px2 = _device.Points(spec, Xd)
        tmin, tmed = timeit(lambda: _device.gram(spec, px, px2, out=out))
        print(f"algo {algo} B cross n={n}: {tmed:.3f} ms  entries/s {n*n/tmed*1e3:.3e}  frac {n*n*306/(tmed*1e-3)/peak:.3f}")
        spec.close()
