"""One-dimensional spherical Gaussian mixture on the device: the fit behind the MOG input measure.

The reference calls ``GaussianMixture(n_components=K, random_state=0, covariance_type="spherical").fit(x[:, None])``
(oak/model_utils.py:753-770).  ``GaussianMixture1D`` follows scikit-learn's algorithm (sklearn/mixture/_base.py,
_gaussian_mixture.py of the 1.x line): responsibilities initialised from ``KMeans(n_clusters=K, n_init=1,
random_state=<the same RandomState>)`` labels, EM with ``tol = 1e-3`` on the change of the mean log-likelihood,
``max_iter = 100``, ``reg_covar = 1e-6``.  The O(N K) part of every iteration -- log-densities, ``logsumexp``,
responsibilities and the three weighted sums of the M step -- is one kernel with fixed-order reductions
(``oak_gmm1d_estep_f64``); the K-element M step is NumPy, written as the library writes it.  The k-means labels come
from the device k-means (``kmeans.KMeans``)."""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _cabi, _device
from ._cabi import check
from ._device import _p, stream_ptr
from .kmeans import KMeans, _random_state


class GaussianMixture1D:
    def __init__(self, n_components: int = 1, *, random_state=None, tol: float = 1e-3, max_iter: int = 100,
                 reg_covar: float = 1e-6):
        self.n_components, self.random_state = int(n_components), random_state
        self.tol, self.max_iter, self.reg_covar = float(tol), int(max_iter), float(reg_covar)

    def _m_step(self, sums, n):
        """_estimate_gaussian_parameters (spherical, one feature) + GaussianMixture._m_step."""
        K = self.n_components
        nk = sums[:K] + 10 * np.finfo(np.float64).eps
        means = sums[K: 2 * K] / nk
        cov = sums[2 * K: 3 * K] / nk - means ** 2 + self.reg_covar
        return nk, means, cov

    def fit(self, x):
        torch = _device._torch()
        lib = _cabi.load()
        xd = _device.to_device(np.asarray(x, dtype=np.float64).reshape(-1) if _device.is_host(x) else x, ndim=1)
        xd = xd.reshape(-1).contiguous()
        n, K = int(xd.numel()), self.n_components
        if K < 1 or K > 16:
            raise NotImplementedError("1 to 16 mixture components")
        if n < K:
            raise ValueError(f"Expected n_samples >= n_components but got n_components = {K}, n_samples = {n}")
        rs = _random_state(self.random_state)
        dev = xd.device
        work = torch.empty(max(int(lib.oak_gmm1d_work_bytes(n, K)) // 8, 1), dtype=torch.float64, device=dev)
        out = torch.empty(3 * K + 1, dtype=torch.float64, device=dev)
        par = torch.zeros(4 * K, dtype=torch.float64, device=dev)

        def sums(labels=None):
            check(lib.oak_gmm1d_estep_f64(_p(xd), n, K, _p(par), _p(labels), _p(out), _p(work),
                                          C.c_void_p(stream_ptr())), "oak_gmm1d_estep_f64")
            return out.cpu().numpy()

        def set_par(weights, means, cov):
            pc = 1.0 / np.sqrt(cov)                      # _compute_precision_cholesky (spherical)
            par.copy_(torch.as_tensor(np.concatenate([means, pc ** 2, np.log(pc), np.log(weights)])))

        # initialisation: hard responsibilities from k-means labels (mixture/_base.py:119-128, 158)
        km = KMeans(n_clusters=K, random_state=rs).fit(xd.reshape(-1, 1))
        labels = torch.as_tensor(km.labels_, dtype=torch.int32, device=dev)
        nk, means, cov = self._m_step(sums(labels), n)
        weights = nk / n
        set_par(weights, means, cov)
        lower_bound, converged, n_iter = -np.inf, False, 0
        for n_iter in range(1, self.max_iter + 1):
            prev = lower_bound
            s = sums()                                   # E step with the current parameters
            nk, means, cov = self._m_step(s, n)
            weights = nk / n
            weights = weights / weights.sum()
            set_par(weights, means, cov)
            lower_bound = s[3 * K] / n                   # mean of log p(x) under the parameters of the E step
            if abs(lower_bound - prev) < self.tol:
                converged = True
                break
        if not converged:
            warnings.warn("GaussianMixture1D did not converge; try different init parameters, or increase max_iter, "
                          "tol, or check for degenerate data.")
        self.weights_, self.means_, self.covariances_ = weights, means.reshape(-1, 1), cov
        self.converged_, self.n_iter_, self.lower_bound_ = converged, n_iter, float(lower_bound)
        return self


def on_device() -> bool:
    """Device implementation on a machine with a CUDA device (a missing library is an error there, see
    ``_device.cuda_available``); scikit-learn's class -- the reference's own call -- only where there is no device at
    all (CPU-only unit tests of the host logic)."""
    return _device.cuda_available()
