"""A/B timing of the backward tiles at config C's chunk shape (development aid): plain backward, backward with
row-point gradients, and the full SGPR training step with fixed / trainable inducing points.
usage: [OAK_B200_LIB=scripts/ubench/liboak_<variant>.so] python scripts/ab_backward.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200.models import SGPR
from oak_b200.training import sgpr_elbo_and_grad
from oak_b200.workloads import build_kernel, config_C

tag = os.environ.get("OAK_B200_LIB", "default")
have_rows = "bwprev" not in tag
n, nc = int(os.environ.get("AB_N", 1_000_000)), 262144
cfg = config_C(n, 20, 1024, 3)
k = build_kernel(cfg)
spec = k._make_spec()
px = _device.Points(spec, _device.to_device(cfg["X"][:nc]))
pz = _device.Points(spec, _device.to_device(cfg["Z"]))
W = torch.randn(1024, nc, dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = [f"[{tag}]"]
out.append(f"backward 1024 x {nc}: {timed(lambda: _device.gram_backward(spec, pz, W, px2=px)):.3f} ms")
Wt = W.T.contiguous()
out.append(f"backward {nc} x 1024: {timed(lambda: _device.gram_backward(spec, px, Wt, px2=pz)):.3f} ms")
if have_rows:
    out.append(f"backward + rows 1024 x {nc}: {timed(lambda: _device.gram_backward_rows(spec, pz, W, px2=px)):.3f} ms")
Kf = torch.empty(1024, nc, dtype=torch.float64, device="cuda")
out.append(f"forward 1024 x {nc}: {timed(lambda: _device.gram(spec, pz, px, out=Kf)):.3f} ms")
del W, Wt, Kf
spec.close()
model = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=262144)
model.likelihood.variance.assign(cfg["noise"])
model._device_data()
for ztrain in ((False, True) if have_rows else (False,)):
    model.inducing_variable.Z.trainable = ztrain
    out.append(f"training step (Z trainable={ztrain}): {timed(lambda: sgpr_elbo_and_grad(model), reps=3):.1f} ms")
print(" | ".join(out))
