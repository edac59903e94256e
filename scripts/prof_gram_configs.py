"""Symmetric Gram of the named configuration (B, D, A-shaped) for ncu captures / quick timings (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oak_b200 import _device
from oak_b200.workloads import build_kernel, config_A, config_B, config_D
which = os.environ.get("AB_CFG", "D")
cfg = {"B": lambda: config_B(int(os.environ.get("AB_N", 32768))), "D": lambda: config_D(),
       "A": lambda: config_A(int(os.environ.get("AB_N", 32768)))}[which]()
slots = {"B": 306.0, "D": 135.0, "A": 244.0}[which]
k = build_kernel(cfg); spec = k._make_spec()
Xd = _device.to_device(cfg["X"]); n = Xd.shape[0]
px = _device.Points(spec, Xd)
out = torch.empty((n, n), dtype=torch.float64, device="cuda")
for _ in range(2): _device.gram(spec, px, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): _device.gram(spec, px, out=out)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
peak = _device.measure_fp64_peak(0.5)
print(f"config {which}: n={n} {ms:.3f} ms, {n*(n+1)/2/ms/1e6:.2f} G unique entries/s, frac {n*(n+1)/2*slots/(ms*1e-3)/peak:.3f} of the {slots:.0f}-slot model")
