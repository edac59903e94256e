"""Timing of the device k-means at the inducing-point scale of config C (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200.kmeans import KMeans
from oak_b200 import _device
n, d, k = int(os.environ.get("KM_N", 1_000_000)), 20, int(os.environ.get("KM_K", 1024))
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)) + 2.0 * rng.standard_normal((64, d))[rng.integers(0, 64, n)]
Xd = _device.to_device(X)
KMeans(n_clusters=8, random_state=0, max_iter=2).fit(Xd[:10000])   # warm-up
for max_iter in (1, 300):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    km = KMeans(n_clusters=k, random_state=0, max_iter=max_iter).fit(Xd)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"N={n} d={d} k={k} max_iter={max_iter}: {t:.3f} s, {km.n_iter_} Lloyd iterations")
if os.environ.get("KM_SKLEARN"):
    from sklearn.cluster import KMeans as Sk
    ns = int(os.environ["KM_SKLEARN"])
    t0 = time.perf_counter(); sk = Sk(n_clusters=k, random_state=0).fit(X[:ns]); t = time.perf_counter() - t0
    print(f"scikit-learn on the first {ns} points: {t:.2f} s, {sk.n_iter_} iterations")
    t0 = time.perf_counter(); km = KMeans(n_clusters=k, random_state=0).fit(Xd[:ns]); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"device on the same {ns} points: {t:.3f} s, {km.n_iter_} iterations, max rel diff of the centres "
          f"{np.abs(km.cluster_centers_ - sk.cluster_centers_).max() / np.abs(sk.cluster_centers_).max():.2e}")
