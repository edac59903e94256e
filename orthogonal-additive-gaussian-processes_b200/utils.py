"""Analysis utilities on top of the fused tiles: sufficient statistics, Sobol indices and
per-component predictions.

Drop-ins for the reference's ``oak/utils.py`` entry points on the hot path:
``get_model_sufficient_statistics`` (:168-218), ``compute_L`` / ``compute_L_binary_kernel`` /
``compute_L_categorical_kernel`` / ``compute_L_empirical_measure`` (:221-335),
``compute_sobol_oak`` (:338-435) and ``get_prediction_component`` (:491-530).  Same names,
argument meaning and error behaviour; the arithmetic runs in ``csrc/oak_sobol.cu``,
``csrc/oak_sgpr.cu`` and ``csrc/oak_component.cu``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from . import _cabi, _device
from ._gpflow_shim import DEFAULT_JITTER, scalar_of, value_of
from .input_measures import EmpiricalMeasure, MOGMeasure
from .models import GPR, SGPR, SVGP
from .oak_kernel import KernelComponenent, OAKKernel, get_list_representation
from .ortho_binary_kernel import OrthogonalBinary
from .ortho_categorical_kernel import OrthogonalCategorical
from .ortho_rbf_kernel import RBF, OrthogonalRBFKernel


def get_model_sufficient_statistics(m, get_L: bool = True):
    """``alpha`` (and the effective ``L``) used for prediction and Sobol indices.

    SGPR: ``alpha = L^-T LB^-T c`` (utils.py:180-198); GPR: ``alpha = (K + noise I)^-1 y``
    (:206-211).  Returned as NumPy (column vector) like ``tensor.numpy()`` consumers expect.
    """
    import torch

    if isinstance(m, SVGP):
        # posterior().alpha = L^-T q_mu; L = chol(inv(Qinv)), Qinv = L^-T (I - diag(q_sqrt^2)) L^-1 (utils.py:174-179,
        # gpflow 2.2.1 posteriors.py, whitened q); chol raises when some q_sqrt >= 1, as in the reference
        alpha = m.sufficient_statistics()
        alpha_h = alpha.reshape(-1, 1).cpu().numpy()
        if not get_L:
            return alpha_h
        Zs = m._Z_device()
        spec = m.kernel._make_spec()
        try:
            Kuu = _device.gram(spec, _device.Points(spec, Zs))
        finally:
            spec.close()
        Kuu.diagonal().add_(DEFAULT_JITTER)
        Lm = torch.linalg.cholesky(Kuu)
        s2 = _device.to_device(m.q_sqrt.numpy(), ndim=1).reshape(-1) ** 2
        Linv = torch.linalg.solve_triangular(Lm, torch.eye(Lm.shape[0], dtype=Lm.dtype, device=Lm.device), upper=False)
        Qinv = Linv.T @ ((1.0 - s2)[:, None] * Linv)
        return alpha_h, torch.linalg.cholesky(torch.linalg.inv(Qinv)).cpu().numpy()
    if isinstance(m, SGPR):
        tail, fac, _ = m._statistics(True)
        tail.host()  # raises when a factorisation failed
        alpha_h = tail.alpha.reshape(-1, 1).cpu().numpy()
        if not get_L:
            return alpha_h
        # effective L (utils.py:200-204): inv(L^-1 - LB^-1 L^-1); L^-1 comes with the factor
        LAi = fac.Linv()
        LBiLAi = torch.linalg.solve_triangular(tail.LB(), LAi, upper=False)
        return alpha_h, torch.linalg.inv(LAi - LBiLAi).cpu().numpy()
    if isinstance(m, GPR):
        Kfac, _, alpha = m._factorise()
        alpha_h = alpha.reshape(-1, 1).cpu().numpy()
        if not get_L:
            return alpha_h
        return alpha_h, torch.triu(Kfac).T.cpu().numpy()
    raise NotImplementedError


# ---- closed forms of the Gaussian-measure L (utils.py:116-165): eq. (44)-(47) of the paper -------------
def _sobol_term(k: int, x, y, sigma, lengthscales, delta, mu):
    """Elementwise over broadcast (x, y); evaluated by the device functions that also build ``compute_L``."""
    host = _device.is_host(x) and _device.is_host(y)
    xb, yb = np.broadcast_arrays(np.asarray(value_of(x), dtype=np.float64), np.asarray(value_of(y), dtype=np.float64)) \
        if host else (x, y)
    xd = _device.to_device(np.ascontiguousarray(xb).reshape(-1) if host else xb.reshape(-1), ndim=1)
    yd = _device.to_device(np.ascontiguousarray(yb).reshape(-1) if host else yb.reshape(-1), ndim=1)
    out = _device.sobol_gaussian_terms(xd, yd, float(sigma), float(lengthscales), float(delta), float(mu))[k]
    return out.cpu().numpy().reshape(np.shape(xb)) if host else out.reshape(xb.shape)


def f1(x, y, sigma, lengthscales, delta, mu):
    return _sobol_term(0, x, y, sigma, lengthscales, delta, mu)


def f2(x, y, sigma, lengthscales, delta, mu):
    return _sobol_term(1, x, y, sigma, lengthscales, delta, mu)


def f3(x, y, sigma, lengthscales, delta, mu):
    return _sobol_term(2, x, y, sigma, lengthscales, delta, mu)


def f4(x, y, sigma, lengthscales, delta, mu):
    return _sobol_term(3, x, y, sigma, lengthscales, delta, mu)


# ---- k-means inducing points with discrete columns (utils.py:533-574) ---------------------------------------------
def initialize_kmeans_with_binary(X, binary_index: list, continuous_index: Optional[list] = None,
                                  n_clusters: Optional[int] = 200) -> np.ndarray:
    """One k-means per binary column (centres truncated to integers) and one over the continuous block.

    The continuous block -- the O(N k d) part -- runs on the device (``kmeans.KMeans``: scikit-learn's algorithm and
    random stream).  The one-column k-means of a DISCRETE feature asks for more clusters than the column has distinct
    values; what comes out of scikit-learn there is decided by its handling of duplicate points and empty clusters on
    rounding-level distances, so that call stays the library's (N x 1, negligible)."""
    from sklearn.cluster import KMeans as SklearnKMeans

    from .kmeans import kmeans_class

    KMeans = kmeans_class()
    X = np.asarray(X, dtype=np.float64)
    Z = np.zeros([n_clusters, X.shape[1]])
    for index in binary_index:
        km = SklearnKMeans(n_clusters=n_clusters, random_state=0).fit(X[:, index][:, None])
        Z[:, index] = km.cluster_centers_.astype(int)[:, 0]
    if continuous_index is not None:
        km = KMeans(n_clusters=n_clusters, random_state=0).fit(X[:, continuous_index])
        Z[:, continuous_index] = km.cluster_centers_
    return Z


def initialize_kmeans_with_categorical(X, binary_index: list, categorical_index: list, continuous_index: list,
                                       n_clusters: Optional[int] = 200) -> np.ndarray:
    """The same with categorical columns treated like the binary ones (:555-574)."""
    return initialize_kmeans_with_binary(X, list(binary_index) + list(categorical_index), continuous_index, n_clusters)


# ---- single L matrices (same call signatures as the reference) -----------------------------
def _L_single(dim_spec, Xcol, delta, mu):
    spec = _cabi.Spec([dim_spec], 1, [0.0, 1.0], True, stream=_device.stream_ptr())
    try:
        Xd = _device.to_device(np.asarray(Xcol, dtype=np.float64).reshape(-1, 1))
        return _device.sobol_L(spec, 0, Xd, delta, mu)
    finally:
        spec.close()


def compute_L(X, lengthscale: float, variance: float, dim: int, delta: float, mu: float) -> np.ndarray:
    """Gaussian-measure ``L`` (utils.py:221-240): ``variance**2 * (f1 - f2 - f3 + f4)``."""
    X = np.asarray(value_of(X))
    ds = _cabi.DimSpec(_cabi.DIM_RBF, 0, measure=_cabi.MEASURE_GAUSSIAN, lengthscale=float(lengthscale),
                       variance=1.0, m0=float(mu), m1=float(delta) ** 2)
    L = _L_single(ds, X[:, dim], delta, mu)
    return (float(variance) ** 2 * L).cpu().numpy()


def compute_L_binary_kernel(X, p0: float, variance: float, dim: int) -> np.ndarray:
    """Binary-kernel ``L`` (utils.py:243-272); scaled by ``variance**1`` exactly like the reference."""
    assert 0 <= p0 <= 1
    X = np.asarray(value_of(X))
    ds = _cabi.DimSpec(_cabi.DIM_BINARY, 0, variance=1.0, m0=float(p0))
    L = _L_single(ds, X[:, dim], 1.0, 0.0)
    return (float(variance) * L).cpu().numpy()


def compute_L_categorical_kernel(X, W, kappa, p, variance: float, dim: int) -> np.ndarray:
    """Categorical-kernel ``L`` (utils.py:275-309)."""
    p = np.asarray(value_of(p), dtype=np.float64)
    assert np.abs(p.sum() - 1) < 1e-6
    X = np.asarray(value_of(X))
    Wv = value_of(W)
    ds = _cabi.DimSpec(_cabi.DIM_CATEGORICAL, 0, variance=1.0, v0=Wv, v1=value_of(kappa), v2=p.reshape(-1),
                       rank=Wv.shape[1])
    L = _L_single(ds, X[:, dim], 1.0, 0.0)
    return (float(variance) ** 2 * L).cpu().numpy()


def compute_L_empirical_measure(x, w, kernel: OrthogonalRBFKernel, z) -> np.ndarray:
    """Empirical-measure ``L = (w o k~(x, z))^T k~(x, z)`` (utils.py:312-335)."""
    ds = kernel.base_kernel._dim_spec(0, EmpiricalMeasure(np.asarray(x).reshape(-1, 1), np.asarray(w).reshape(-1, 1)))
    return _L_single(ds, np.asarray(value_of(z)).reshape(-1), 1.0, 0.0).cpu().numpy()


def compute_sobol_oak(model, delta: float, mu: float,
                      share_var_across_orders: Optional[bool] = True) -> Tuple[List[List[int]], List[float]]:
    """Sobol index of every additive component of an OAK model (utils.py:338-435).

    One ``L_d`` per input dimension is built on the device (the reference rebuilds it for every
    subset containing d), then all ``alpha^T (prod_d L_d) alpha`` quadratic forms run in one launch.
    """
    import torch

    assert isinstance(model.kernel, OAKKernel), "only work for OAK kernel"
    kern: OAKKernel = model.kernel
    num_dims = np.shape(model.data[0])[1]
    # the subsets of get_list_representation (oak_kernel.py:338-364) without its KernelComponenent objects: at D = 50,
    # depth 2 building 1276 of them cost 5 ms of a 17 ms call, and only their index lists were read here
    import itertools

    selected_dims_oak = []
    for order in range(1, int(kern.max_interaction_depth) + 1):
        selected_dims_oak += [list(t) for t in itertools.combinations(np.arange(num_dims), order)]
    if isinstance(model, (SGPR, SVGP)):
        Xc = model._slice_for_kernel(_device.to_device(value_of(model.inducing_variable.Z)))
    else:
        Xc = model._slice_for_kernel(model._device_data()[0])
    alpha = model.sufficient_statistics()

    # per-component scale: order variance enters through the first dim only (utils.py:376-380).  Per sub-kernel: the
    # exponent of v (sigma^4 (:119) and v**2 (:404) for RBF; variance**1 (:266) for the binary kernel -- reference
    # quirk, replicated; B * variance on both factors (:299, :307) for the categorical one)
    power = []
    for k in kern.kernels:
        if isinstance(k, OrthogonalRBFKernel):
            power.append(None if isinstance(k.measure, MOGMeasure) else 2)  # MOG: NotImplementedError (:413-414)
        elif isinstance(k, OrthogonalBinary):
            power.append(1)
        elif isinstance(k, OrthogonalCategorical):
            power.append(2)
        else:
            raise NotImplementedError
    order_var = [scalar_of(v) for v in kern.variances]
    subsets, scales = [], []
    for comp in selected_dims_oak:
        S = sorted(int(i) for i in comp)
        scale = 1.0
        for j, d in enumerate(S):
            if share_var_across_orders:
                v = order_var[len(S)] if j < 1 else 1.0
            else:
                v = scalar_of(kern.kernels[d].base_kernel.variance)  # AttributeError for discrete kernels, as in :382
            if power[d] is None:
                raise NotImplementedError
            scale *= v ** power[d]
        subsets.append(S)
        scales.append(scale)

    m = int(Xc.shape[0])
    D = len(kern.kernels)
    spec = kern._make_spec()
    try:
        Lstack = torch.empty((D, m, m), dtype=torch.float64, device=Xc.device)
        for d in range(D):
            _device.sobol_L(spec, d, Xc, float(delta), float(mu), out=Lstack[d])
        sob = _device.sobol_quadforms(Lstack, subsets, scales, alpha)
    finally:
        spec.close()
    sobol = [float(v) for v in sob.cpu().numpy()]
    assert len(selected_dims_oak) == len(sobol)
    return selected_dims_oak, sobol


def get_prediction_component(m, alpha, X=None, share_var_across_orders: Optional[bool] = True) -> list:
    """Predictive mean of every additive component (utils.py:491-530), fused with the alpha
    contraction so that no N* x M matrix per component is materialised."""
    if X is None:
        X = m.data[0]
    kern: OAKKernel = m.kernel
    selected_dims, _ = get_list_representation(kern, num_dims=np.shape(X)[1])
    subsets = [sorted(int(i) for i in s) for s in selected_dims[1:]]
    if isinstance(m, GPR):
        Xc = m._slice_for_kernel(m._device_data()[0])
    elif isinstance(m, (SGPR, SVGP)):
        Xc = m._slice_for_kernel(_device.to_device(value_of(m.inducing_variable.Z)))
    else:
        raise NotImplementedError
    Xd = m._slice_for_kernel(_device.to_device(X))
    a = _device.to_device(value_of(alpha))
    # order variances only when sharing (utils.py:525-526)
    var = kern._order_variances() if share_var_across_orders else [1.0] * (kern._depth() + 1)
    spec = _cabi.Spec(kern._dim_specs(), kern._depth(), var, True, stream=_device.stream_ptr())
    try:
        out = _device.component_predict(spec, subsets, _device.Points(spec, Xd), _device.Points(spec, Xc), a)
    finally:
        spec.close()
    host = out.cpu().numpy()
    return [host[c] for c in range(host.shape[0])]
