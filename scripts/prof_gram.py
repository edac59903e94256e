"""Launches for an ncu capture: config B, symmetric then cross Gram (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oak_b200 import _device
from oak_b200.workloads import config_B, build_kernel
n = int(os.environ.get("AB_N", 16384))
cfg = config_B(n)
k = build_kernel(cfg); spec = k._make_spec()
Xd = _device.to_device(cfg["X"])
px = _device.Points(spec, Xd); px2 = _device.Points(spec, Xd)
out = torch.empty((n, n), dtype=torch.float64, device="cuda")
for _ in range(2):
    _device.gram(spec, px, out=out)
    _device.gram(spec, px, px2, out=out)
torch.cuda.synchronize()
