"""Timing of ELBO + gradient (backward tiles) at config C scale (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200.models import SGPR
from oak_b200.training import freeze_unsupported, sgpr_elbo_and_grad
from oak_b200.workloads import build_kernel, config_C
n = int(os.environ.get("AB_N", 1_000_000))
cfg = config_C(n, 20, 1024, 3)
model = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=262144)
model.likelihood.variance.assign(cfg["noise"])
freeze_unsupported(model)
model._device_data()
for _ in range(2):
    out = sgpr_elbo_and_grad(model)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    out = sgpr_elbo_and_grad(model)
torch.cuda.synchronize()
t = (time.perf_counter() - t0) / reps
e0 = time.perf_counter(); v = model.elbo(); torch.cuda.synchronize(); te = time.perf_counter() - e0
print(f"N={n}: elbo+grad {t*1e3:.1f} ms ({1/t:.2f} evals/s), elbo only {te*1e3:.1f} ms; elbo {out[0]:.6f} vs {v:.6f}; |g_ls| {np.abs(out[1]).max():.3e}")
# kernel-level: backward tile kernel alone on one chunk
k = build_kernel(cfg); spec = k._make_spec()
Xd, Zd = _device.to_device(cfg["X"][:65536]), _device.to_device(cfg["Z"])
px, pz = _device.Points(spec, Xd), _device.Points(spec, Zd)
W = torch.randn(65536, 1024, dtype=torch.float64, device="cuda")
for _ in range(2): _device.gram_backward(spec, px, W, px2=pz)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): _device.gram_backward(spec, px, W, px2=pz)
e1.record(); torch.cuda.synchronize()
tb = e0.elapsed_time(e1) / 5
out_k = torch.empty(65536, 1024, dtype=torch.float64, device="cuda")
e0.record()
for _ in range(5): _device.gram(spec, px, pz, out=out_k)
e1.record(); torch.cuda.synchronize()
tf = e0.elapsed_time(e1) / 5
print(f"one 65536 x 1024 chunk: backward tiles {tb:.3f} ms, forward tiles {tf:.3f} ms (ratio {tb/tf:.2f})")
