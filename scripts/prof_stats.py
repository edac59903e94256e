"""One statistics pass at config C's chunk shape (for ncu captures of the contraction kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oak_b200 import _device
from oak_b200.workloads import build_kernel, config_C
n = int(os.environ.get("AB_N", 262144))
route = int(os.environ.get("AB_ROUTE", 0))
cfg = config_C(n, 20, 1024, 3)
k = build_kernel(cfg)
spec = k._make_spec()
pz = _device.Points(spec, _device.to_device(cfg["Z"]))
px = _device.Points(spec, _device.to_device(cfg["X"]))
y = _device.to_device(cfg["y"])
fac = _device.sgpr_factor(spec, pz, 1e-6, route=route)
for _ in range(int(os.environ.get("AB_REPS", 3))):
    st = _device.sgpr_stats2(spec, pz, px, y, fac, chunk=262144)
torch.cuda.synchronize()
print("ok", float(st[0]))
