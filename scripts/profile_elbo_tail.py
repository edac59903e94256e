"""Where the non-statistics time of one ELBO evaluation goes (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200._gpflow_shim import DEFAULT_JITTER
from oak_b200.models import SGPR
from oak_b200.workloads import build_kernel, config_C
n = int(os.environ.get("AB_N", 125_000))
cfg = config_C(n, 20, 1024, 3)
model = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=65536)
model.likelihood.variance.assign(cfg["noise"])
Xd, Yd = model._device_data()
for _ in range(3): model.elbo()
def T(fn, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, r
kern = model.kernel
t_total, _ = T(model.elbo)
t_spec, _ = T(lambda: kern._make_spec().close())
spec = kern._make_spec()
Zs = model._Z_device()
t_zdev, _ = T(model._Z_device)
t_pz, pz = T(lambda: _device.Points(spec, Zs))
t_px, px = T(lambda: _device.Points(spec, Xd))
t_stats, stats = T(lambda: _device.sgpr_stats(spec, pz, px, Yd, chunk=65536))
t_kuu, Kuu = T(lambda: _device.gram(spec, pz))
def fin():
    return _device.sgpr_finish(Kuu.clone(), stats.clone(), n, 0.01, DEFAULT_JITTER, want_alpha=False)
t_fin, out = T(fin)
t_clone, _ = T(lambda: (Kuu.clone(), stats.clone()))
t_item, _ = T(lambda: float(out[0][0].item()))
print(f"N={n}: elbo total {t_total:.3f} ms | spec create+destroy {t_spec:.3f} | Z to device {t_zdev:.3f} | Points(Z) {t_pz:.3f} | "
      f"Points(X) {t_px:.3f} | stats {t_stats:.3f} | Kuu gram {t_kuu:.3f} | finish {t_fin - t_clone:.3f} (+clone {t_clone:.3f}) | item {t_item:.3f}")
