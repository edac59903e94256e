"""Launches for an ncu capture of the backward tiles at config C's chunk shape (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oak_b200 import _device
from oak_b200.workloads import build_kernel, config_C
nc = int(os.environ.get("AB_NC", 262144))
cfg = config_C(nc, 20, 1024, 3)
spec = build_kernel(cfg)._make_spec()
px = _device.Points(spec, _device.to_device(cfg["X"][:nc]))
pz = _device.Points(spec, _device.to_device(cfg["Z"]))
W = torch.randn(1024, nc, dtype=torch.float64, device="cuda")
for _ in range(3):
    _device.gram_backward(spec, pz, W, px2=px)
    _device.gram_backward_rows(spec, pz, W, px2=px)
torch.cuda.synchronize()
