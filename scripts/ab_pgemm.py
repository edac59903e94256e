"""Cotangent product W = G2 Kuf + g_b y^T (1024 x 1024 x 262144): torch.addmm (cuBLAS) vs the in-house DMMA panel product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oak_b200 import _device
M, n = 1024, 262144
g = torch.Generator(device="cuda").manual_seed(0)
G2 = torch.randn((M, M), generator=g, dtype=torch.float64, device="cuda")
K = torch.randn((M, n), generator=g, dtype=torch.float64, device="cuda")
u = torch.randn(M, generator=g, dtype=torch.float64, device="cuda")
v = torch.randn(n, generator=g, dtype=torch.float64, device="cuda")
out = torch.empty((M, n), dtype=torch.float64, device="cuda")
def T(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r
t_lib, ref = T(lambda: torch.addmm(u[:, None] @ v[None, :], G2, K))
t_own, got = T(lambda: _device.panel_gemm(G2, K, u=u, v=v, out=out))
err = float((got - ref).abs().max() / ref.abs().max())
t_tri, _ = T(lambda: _device.panel_gemm(torch.tril(G2), K, lower=True, out=out))
fl = 2.0 * M * M * n
print(f"cuBLAS addmm + outer {t_lib:.2f} ms ({fl / t_lib / 1e9:.1f} TFLOP/s) | panel_gemm {t_own:.2f} ms ({fl / t_own / 1e9:.1f} TFLOP/s), "
      f"max rel diff {err:.1e} | lower-triangular {t_tri:.2f} ms ({fl * 0.53 / t_tri / 1e9:.1f} TFLOP/s of useful work)")
