"""Where the time of compute_sobol_oak goes on configs A and E (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200.models import GPR, SGPR
from oak_b200.oak_kernel import get_list_representation
from oak_b200.utils import compute_sobol_oak
from oak_b200.workloads import build_kernel, config_A, config_E
def T(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, r
for name, cfg, mk in (("A", config_A(), lambda c: GPR((c["X"], c["y"]), kernel=build_kernel(c))),
                      ("E", config_E(), lambda c: SGPR((c["X"], c["y"]), kernel=build_kernel(c), inducing_variable=c["Z"], chunk=65536))):
    m = mk(cfg); m.likelihood.variance.assign(cfg["noise"])
    t_all, (idx, sob) = T(lambda: compute_sobol_oak(m, 1.0, 0.0))
    t_alpha, alpha = T(m.sufficient_statistics)
    t_list, (sel, kl) = T(lambda: get_list_representation(m.kernel, num_dims=cfg["X"].shape[1]))
    kern = m.kernel; spec = kern._make_spec()
    Xc = m._slice_for_kernel(_device.to_device(cfg["Z"] if name == "E" else cfg["X"]))
    mm, D = int(Xc.shape[0]), len(kern.kernels)
    Ls = torch.empty((D, mm, mm), dtype=torch.float64, device="cuda")
    def build_L():
        for d in range(D): _device.sobol_L(spec, d, Xc, 1.0, 0.0, out=Ls[d])
    t_L, _ = T(build_L)
    subsets = [sorted(int(i) for i in c.iComponent_list) for c in kl[1:]]
    t_q, _ = T(lambda: _device.sobol_quadforms(Ls, subsets, [1.0] * len(subsets), alpha))
    spec.close()
    print(f"config {name}: compute_sobol_oak {t_all:.2f} ms | alpha {t_alpha:.2f} | get_list_representation {t_list:.2f} | "
          f"{D} L matrices ({mm} x {mm}) {t_L:.2f} | {len(subsets)} quadratic forms {t_q:.2f} | rest (host loops) "
          f"{t_all - t_alpha - t_list - t_L - t_q:.2f}")
