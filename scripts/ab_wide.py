"""Timing of the depth-2 shapes (development aid, not the bench): config D Gram (symmetric, mixed inputs), its
cross-covariance, config E's Kuf tiles and a depth-2 ELBO.  Written for the A/B of the 64 x 128 tile geometry
(profiles/r02bb_ab_gram_wide_tile_negative.txt; that build read OAK_GRAM_WIDE, the shipped one ignores it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200.models import SGPR
from oak_b200.workloads import config_D, config_E, build_kernel

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

peak = _device.measure_fp64_peak(0.5)
line = [f"[OAK_GRAM_WIDE={os.environ.get('OAK_GRAM_WIDE', '1')}] peak {2*peak/1e12:.2f} TF"]
n = 32768
cfg = config_D(n)
k = build_kernel(cfg); spec = k._make_spec()
Xd = _device.to_device(cfg["X"])
px = _device.Points(spec, Xd)
out = torch.empty((n, n), dtype=torch.float64, device="cuda")
t = timeit(lambda: _device.gram(spec, px, out=out))
line.append(f"D sym n={n}: {t:.3f} ms frac(135) {n*(n+1)/2*135/(t*1e-3)/peak:.3f}")
ref = out[:512, :640].clone()
px2 = _device.Points(spec, Xd)
t = timeit(lambda: _device.gram(spec, px, px2, out=out))
line.append(f"D cross: {t:.3f} ms frac(135) {n*n*135/(t*1e-3)/peak:.3f}")
line.append(f"sym-vs-cross corner max|diff| {float((out[:512, :640] - ref).abs().max()):.1e} symmetry {float((ref[:512, :512] - ref[:512, :512].T).abs().max()):.1e}")
spec.close()
cfg = config_E(200_000)
k = build_kernel(cfg); spec = k._make_spec()
Xd, Zd = _device.to_device(cfg["X"][:131072]), _device.to_device(cfg["Z"])
pz, px = _device.Points(spec, Zd), _device.Points(spec, Xd)
out = torch.empty((512, 131072), dtype=torch.float64, device="cuda")
t = timeit(lambda: _device.gram(spec, pz, px, out=out))
line.append(f"E Kuf 512x131072 (D=50, P=2): {t:.3f} ms frac(757) {512*131072*757/(t*1e-3)/peak:.3f}")
spec.close()
model = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=262144)
model.likelihood.variance.assign(cfg["noise"])
val = model.elbo()
t = timeit(lambda: model.elbo(), reps=5)
line.append(f"E ELBO N=200000 M=512: {t:.3f} ms value {float(val):.12e}")
print(" | ".join(line), flush=True)
