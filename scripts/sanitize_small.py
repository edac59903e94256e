"""Small invocations of every hand-written kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
Gram tiles (TMA mirror path and direct-store path, general mode with the folded K y), lower trapezoid + mirror
strips, the DMMA contractions, the panel product, the bordered Cholesky, the SGPR factor / statistics / tail on both
routes, the backward tiles.  Sizes are tiny: the tools slow kernels down 10-100x."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200.models import SGPR
from oak_b200.workloads import build_kernel, config_C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import mixed_config

which = set((os.environ.get("SAN_WHICH") or "gram,strips,sgpr,chol,pgemm,backward").split(","))
cfg = config_C(700, 6, 130, 3)
k = build_kernel(cfg)
spec = k._make_spec()
Xd = _device.to_device(cfg["X"])
px = _device.Points(spec, Xd)
pz = _device.Points(spec, _device.to_device(cfg["Z"]))
if "gram" in which:
    K = _device.gram(spec, px)                                   # symmetric: TMA mirrored stores
    Kodd = torch.empty((700, 701), dtype=torch.float64, device="cuda")[:, :700]
    _device.gram(spec, px, out=Kodd)                             # odd pitch: direct mirrored stores
    Kc = _device.gram(spec, pz, px)                              # general mode
    print("gram", float(K[3, 5]), float(Kodd[3, 5]), float(Kc[1, 2]))
if "strips" in which:
    o, ot = _device.gram_lower_mirror(spec, px, 256, 640)
    o2 = _device.gram_lower(spec, px, 0, 256)
    print("strips", float(o[0, 0]), float(ot[0, 0]), float(o2[0, 0]))
if "chol" in which:
    n = 200
    A = np.random.default_rng(0).standard_normal((n, n)); A = A @ A.T / n + np.eye(n)
    buf = np.zeros((n, 2 * n + 8)); buf[:, :n] = A; buf[:, n:2 * n] = np.eye(n)
    d = torch.as_tensor(buf).cuda()
    info, ld = _device.chol(d, n, 2 * n, border_identity=True)
    print("chol", int(info.item()), float(ld.item()))
if "pgemm" in which:
    T = torch.tril(torch.randn(130, 130, dtype=torch.float64, device="cuda"))
    B = torch.randn(130, 300, dtype=torch.float64, device="cuda")
    print("pgemm", float(_device.panel_gemm(T, B, lower=True)[5, 7]))
if "sgpr" in which:
    for chunk in (64, 256):
        for whiten in (False, True):
            m = SGPR((cfg["X"], cfg["y"]), kernel=k, inducing_variable=cfg["Z"], chunk=chunk, whiten_stats=whiten)
            m.likelihood.variance.assign(cfg["noise"])
            print("sgpr", chunk, whiten, m.elbo())
    mc = mixed_config(n=300, seed=2, depth=2)
    mm = SGPR((mc["X"], mc["y"]), kernel=build_kernel(mc), inducing_variable=mc["Z"], chunk=64)
    mm.likelihood.variance.assign(mc["noise"])
    print("sgpr mixed", mm.elbo())
if "backward" in which:
    W = torch.randn(130, 700, dtype=torch.float64, device="cuda")
    g = _device.gram_backward(spec, pz, W, px2=px)
    print("backward", float(g[0]))
if "overlap" in which:
    # the fused factor + statistics call with the factorisation on the side stream (forced: the sizes are tiny)
    for whiten in (False, True):
        m = SGPR((cfg["X"], cfg["y"]), kernel=k, inducing_variable=cfg["Z"], chunk=256, whiten_stats=whiten)
        m.likelihood.variance.assign(cfg["noise"])
        m.overlap_ctas = 4
        print("sgpr overlapped", whiten, m.elbo(), m.elbo())
if "kmeans" in which:
    from oak_b200.kmeans import KMeans
    rng = np.random.default_rng(1)
    for n_, d_, k_ in ((900, 5, 12), (5000, 20, 70), (600, 70, 7)):
        Xk = rng.standard_normal((n_, d_)) + 3.0 * rng.standard_normal((8, d_))[rng.integers(0, 8, n_)]
        km = KMeans(n_clusters=k_, random_state=0).fit(Xk)
        print("kmeans", n_, d_, k_, km.n_iter_, float(km.cluster_centers_[0, 0]))
if "unique" in which:
    col = np.round(np.random.default_rng(2).standard_normal(5000) * 4) / 4
    Xu = _device.to_device(np.column_stack([col, col[::-1]]))
    v, c = _device.column_unique(Xu, 0)
    print("unique", len(v), int(c.sum()), _device.column_mean(Xu, 1))
spec.close()
torch.cuda.synchronize()
print("sanitize_small done")
