"""Host-side profile of one ELBO evaluation (cProfile over 200 evaluations at the 8-GPU per-rank size)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200.models import SGPR
from oak_b200.workloads import build_kernel, config_C
n = int(os.environ.get("AB_N", 125_000))
cfg = config_C(n, 20, 1024, 3)
model = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=262144)
model.likelihood.variance.assign(cfg["noise"])
for _ in range(5): model.elbo()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): model.elbo()
torch.cuda.synchronize(); print(f"{(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per evaluation")
# host time until everything is enqueued (no read-back): how far ahead of the GPU the host runs
t0 = time.perf_counter()
for _ in range(50):
    tail = model._statistics(False)[0]
t_enq = (time.perf_counter() - t0) / 50 * 1e3
torch.cuda.synchronize(); print(f"{t_enq:.3f} ms of host time to enqueue one evaluation (GPU busy ~7.6 ms)")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): model.elbo()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
