"""k-means on the device for the inducing-point initialisation.

The reference calls scikit-learn: ``KMeans(n_clusters=K, random_state=0).fit(X).cluster_centers_``
(oak/model_utils.py:31-41, 376-383; oak/utils.py:549-552, 570-573).  ``KMeans`` below keeps that interface and
follows scikit-learn's algorithm step by step (sklearn/cluster/_kmeans.py of the 1.x line: ``n_init="auto"`` = one
k-means++ initialisation, Lloyd, ``tol`` relative to the mean column variance) so that the centres agree with the
library's to rounding on data without exact distance ties:

* the random numbers are drawn here, on the host, from ``numpy.random.RandomState`` with exactly the calls
  ``_kmeans_plusplus`` makes (``choice(n, p=...)`` for the first centre, ``uniform(size=2 + int(log k))`` per further
  centre), so the stream is consumed identically;
* everything that touches all N points -- centring, distances to the candidates, the running minimum, potentials,
  the cumulative sum and its ``searchsorted``, label assignment, per-cluster sums -- runs on the device
  (``csrc/oak_kmeans.cu``), with fixed-order reductions: the same centres on every run and on every rank;
* empty clusters are rare and are relocated with scikit-learn's own rule on the host (``_relocate_empty``).

No CPU fallback: without the CUDA library ``fit`` raises (``_cabi.OakNativeError``).  Up to 64 columns run the
register-tiled assignment kernel, wider inputs (up to 256 columns) a plain one.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi, _device
from ._cabi import check
from ._device import _p, stream_ptr


def _random_state(seed):
    """sklearn.utils.check_random_state."""
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, (int, np.integer)):
        return np.random.RandomState(int(seed))
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError(f"{seed!r} cannot be used to seed a numpy.random.RandomState instance")


class KMeans:
    """``sklearn.cluster.KMeans(n_clusters, random_state=..., n_init="auto", init="k-means++", algorithm="lloyd")``."""

    def __init__(self, n_clusters: int = 8, *, random_state=None, max_iter: int = 300, tol: float = 1e-4):
        self.n_clusters, self.random_state, self.max_iter, self.tol = int(n_clusters), random_state, int(max_iter), float(tol)

    # ---- pieces -----------------------------------------------------------------------------------------------
    def _seed(self, lib, Xc, n, d, k, rs, work):
        """k-means++ (_kmeans.py:_kmeans_plusplus): returns the indices of the chosen points."""
        torch = _device._torch()
        dev = Xc.device
        trials = 2 + int(np.log(k))
        cand = torch.zeros(trials, dtype=torch.int64, device=dev)
        dist = torch.empty((trials, n), dtype=torch.float64, device=dev)
        closest = torch.empty(n, dtype=torch.float64, device=dev)
        cum = torch.empty(n, dtype=torch.float64, device=dev)
        pots = torch.empty(trials, dtype=torch.float64, device=dev)
        thr = torch.empty(trials, dtype=torch.float64, device=dev)
        indices = np.full(k, -1, dtype=np.int64)
        sw = np.ones(n)
        indices[0] = rs.choice(n, p=sw / sw.sum())
        cand[0] = int(indices[0])
        check(lib.oak_kmeanspp_round_f64(_p(Xc), n, d, _p(None), _p(None), 1, _p(cand), _p(None), _p(dist), _p(pots),
                                         _p(work), C.c_void_p(stream_ptr())), "oak_kmeanspp_round_f64")
        closest.copy_(dist[0])
        current_pot = float(pots[0].item())
        for c in range(1, k):
            thr.copy_(torch.as_tensor(rs.uniform(size=trials) * current_pot))
            check(lib.oak_kmeanspp_round_f64(_p(Xc), n, d, _p(closest), _p(thr), trials, _p(cand), _p(cum), _p(dist),
                                             _p(pots), _p(work), C.c_void_p(stream_ptr())), "oak_kmeanspp_round_f64")
            pots_h = pots.cpu().numpy()
            best = int(np.argmin(pots_h))
            current_pot = float(pots_h[best])
            closest.copy_(dist[best])
            indices[c] = int(cand[best].item())
        return indices

    @staticmethod
    def _relocate_empty(Xc, labels, centers_old, sums, counts):
        """_k_means_common.pyx:_relocate_empty_clusters_dense + _average_centers + _center_shift on the host (rare)."""
        X = Xc.cpu().numpy()
        lab = labels.cpu().numpy()
        old = centers_old.cpu().numpy()
        new = sums.cpu().numpy().copy()
        w = counts.cpu().numpy().copy()
        empty = np.where(w == 0)[0]
        distances = ((X - old[lab]) ** 2).sum(axis=1)
        if np.max(distances) != 0:
            far = np.argpartition(distances, -len(empty))[: -len(empty) - 1: -1]
            for idx, new_id in enumerate(empty):
                far_idx = far[idx]
                old_id = lab[far_idx]
                new[old_id] -= X[far_idx]
                new[new_id] = X[far_idx]
                w[new_id] = 1.0
                w[old_id] -= 1.0
        argmax_w = int(np.argmax(w))
        for j in range(len(w)):  # in order: an empty cluster copies the row of the biggest one as it is at that moment
            if w[j] > 0:
                new[j] *= 1.0 / w[j]
            else:
                new[j] = new[argmax_w]
        shift = np.sqrt(((new - old) ** 2).sum(axis=1))
        return new, float((shift ** 2).sum())

    # ---- sklearn API ----------------------------------------------------------------------------------------------
    def fit(self, X, y=None):
        torch = _device._torch()
        lib = _cabi.load()
        Xd = _device.to_device(X).contiguous()
        if Xd.dim() != 2:
            raise ValueError("Expected 2D array")
        n, d = int(Xd.shape[0]), int(Xd.shape[1])
        k = self.n_clusters
        if n < k:
            raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")
        if d > 256:
            raise NotImplementedError("the device k-means supports at most 256 columns")
        rs = _random_state(self.random_state)
        dev = Xd.device
        trials = 2 + int(np.log(k))
        work = torch.empty(max(int(lib.oak_kmeans_work_bytes(n, d, k, trials)) // 8, 1), dtype=torch.float64, device=dev)
        Xc = torch.empty((n, d), dtype=torch.float64, device=dev)
        stats = torch.empty(d + 1, dtype=torch.float64, device=dev)
        check(lib.oak_kmeans_center_f64(_p(Xd), n, d, int(Xd.stride(0)), _p(Xc), _p(stats), _p(work),
                                        C.c_void_p(stream_ptr())), "oak_kmeans_center_f64")
        tol_abs = float(stats[d].item()) * self.tol
        indices = self._seed(lib, Xc, n, d, k, rs, work)
        centers = Xc[torch.as_tensor(indices, device=dev)].contiguous()
        centers_new = torch.empty_like(centers)
        sums = torch.empty_like(centers)
        counts = torch.empty(k, dtype=torch.float64, device=dev)
        labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
        out = torch.empty(4, dtype=torch.float64, device=dev)

        def lloyd(update):
            check(lib.oak_kmeans_lloyd_f64(_p(Xc), n, d, _p(centers), k, _p(labels), _p(centers_new), _p(sums), _p(counts),
                                           _p(out), int(update), _p(work), C.c_void_p(stream_ptr())), "oak_kmeans_lloyd_f64")
            o = out.cpu()
            return float(o[0]), int(o.view(torch.int64)[2]), int(o.view(torch.int64)[3])

        strict = False
        n_iter = 0
        for i in range(self.max_iter):
            n_iter = i + 1
            shift_tot, n_empty, changed = lloyd(True)
            if n_empty:
                new, shift_tot = self._relocate_empty(Xc, labels, centers, sums, counts)
                centers_new.copy_(torch.as_tensor(new, device=dev))
            centers, centers_new = centers_new, centers
            if changed == 0:      # np.array_equal(labels, labels_old)
                strict = True
                break
            if shift_tot <= tol_abs:
                break
        if not strict:
            lloyd(False)          # labels consistent with the final centres (centres unchanged)
        self.cluster_centers_ = _device.from_device(centers + stats[:d], True)
        self.labels_ = labels.cpu().numpy()
        self.n_iter_ = n_iter
        self.seed_indices_ = indices
        return self


def kmeans_class():
    """The k-means the model-building helpers use: the device implementation on a machine with a CUDA device (a missing
    library is an error there, ``_device.cuda_available``); scikit-learn's -- the reference's own call -- only where
    there is no device at all, i.e. in the CPU-only unit tests of the host logic.  One-off preprocessing, not the hot
    path; ``KMeans`` itself never falls back."""
    if _device.cuda_available():
        return KMeans
    from sklearn.cluster import KMeans as SklearnKMeans

    return SklearnKMeans
