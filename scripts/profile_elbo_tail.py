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
model = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=262144)
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
t_chk, _ = T(lambda: (kern._check_discrete(Xd, spec._keep), kern._check_discrete(Zs, spec._keep)))
t_pz, pz = T(lambda: _device.Points(spec, Zs))
t_px, px = T(lambda: _device.Points(spec, Xd))
t_fac, fac = T(lambda: _device.sgpr_factor(spec, pz, DEFAULT_JITTER))
t_stats, stats = T(lambda: _device.sgpr_stats2(spec, pz, px, Yd, fac, chunk=262144))
t_fin, tail = T(lambda: _device.sgpr_finish2(fac, stats, n, 0.01, want_alpha=False))
t_item, _ = T(lambda: tail.host())
print(f"N={n}: elbo total {t_total:.3f} ms | spec create+destroy {t_spec:.3f} | Z to device {t_zdev:.3f} | check_discrete {t_chk:.3f} | "
      f"Points(Z) {t_pz:.3f} | Points(X) {t_px:.3f} | factor front {t_fac:.3f} | stats {t_stats:.3f} | finish {t_fin:.3f} | "
      f"readback {t_item:.3f} | route {model.last_route} cond {model.last_cond_estimate:.3e}")
for route in (0, 1):
    f2 = _device.sgpr_factor(spec, pz, DEFAULT_JITTER, route=route)
    t_s, _ = T(lambda: _device.sgpr_stats2(spec, pz, px, Yd, f2, chunk=262144), reps=5)
    print(f"  stats phase, forced route {route}: {t_s:.3f} ms")
# the fused call: factorisation on a side stream next to the first chunk's Kuf tiles
for ov in [int(v) for v in os.environ.get("AB_OVERLAPS", "0,4,6,8,12,16").split(",")]:
    t_fs, _ = T(lambda: _device.sgpr_factor_stats(spec, pz, px, Yd, DEFAULT_JITTER, chunk=262144, overlap_ctas=ov), reps=10)
    model.overlap_ctas = ov
    t_e, _ = T(model.elbo, reps=10)
    print(f"  overlap_ctas {ov:3d}: factor + stats {t_fs:.3f} ms, whole ELBO evaluation {t_e:.3f} ms")
