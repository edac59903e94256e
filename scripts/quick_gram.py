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
print(f"variant {os.environ.get('OAK_GRAM_VARIANT')} fp64 peak slots/s {peak:.4e}  ({2*peak/1e12:.2f} TFLOP/s)")
for algo in (0, 1):
    for n in (16384, 32768):
        cfg = config_B(n)
        k = build_kernel(cfg); k.esp_algorithm = algo
        spec = k._make_spec()
        Xd = _device.to_device(cfg["X"])
        px = _device.Points(spec, Xd)
        out = torch.empty((n, n), dtype=torch.float64, device="cuda")
        tmin, tmed = timeit(lambda: _device.gram(spec, px, out=out))
        uniq = n * (n + 1) / 2
        print(f"algo {algo} B sym   n={n}: {tmed:.3f} ms  unique entries/s {uniq/tmed*1e3:.3e}  frac(306 slots) {uniq*306/(tmed*1e-3)/peak:.3f}")
        tmin, tmed = timeit(lambda: _device.gram_lower(spec, px, 0, n, out=out))
        print(f"algo {algo} B lower n={n}: {tmed:.3f} ms  unique entries/s {uniq/tmed*1e3:.3e}  frac(306 slots) {uniq*306/(tmed*1e-3)/peak:.3f}")
        px2 = _device.Points(spec, Xd)
        tmin, tmed = timeit(lambda: _device.gram(spec, px, px2, out=out))
        print(f"algo {algo} B cross n={n}: {tmed:.3f} ms  entries/s {n*n/tmed*1e3:.3e}  frac {n*n*306/(tmed*1e-3)/peak:.3f}")
        spec.close()
# SGPR stats breakdown (config C at 200k)
cfg = config_C(200_000)
k = build_kernel(cfg); spec = k._make_spec()
Xd, Zd, yd = _device.to_device(cfg["X"]), _device.to_device(cfg["Z"]), _device.to_device(cfg["y"])
pz, px = _device.Points(spec, Zd), _device.Points(spec, Xd)
for chunk in (2048, 4096, 8192, 16384):
    tmin, tmed = timeit(lambda: _device.sgpr_stats(spec, pz, px, yd, chunk=chunk), reps=3, warm=1)
    print(f"sgpr_stats N=200k M=1024 chunk={chunk}: {tmed:.3f} ms  frac(865.5) {1024*200000*865.5/(tmed*1e-3)/peak:.3f}")
out = torch.empty((1024, 200_000), dtype=torch.float64, device="cuda")
tmin, tmed = timeit(lambda: _device.gram(spec, pz, px, out=out), reps=3, warm=1)
print(f"Kuf generation alone: {tmed:.3f} ms frac(352) {1024*200000*352/(tmed*1e-3)/peak:.3f}")
A = out
tmin, tmed = timeit(lambda: torch.mm(A, A.T), reps=3, warm=1)
print(f"torch DGEMM Kuf Kuf^T (full, not syrk): {tmed:.3f} ms -> {2*1024*1024*200000/(tmed*1e-3)/1e12:.2f} TFLOP/s")
