"""Synthetic workloads of BASELINE.json's configs and a builder that turns
a plain configuration dict into an ``OAKKernel``.  Shared by ``bench.py`` and the parity tests so
that both exercise exactly the same inputs.  Data generation only -- no kernel arithmetic.

A configuration is a dict with keys
  ``dims``       list of per-dimension dicts:
                   {"type": "rbf", "lengthscale", "variance", "measure": None | ("gaussian", mu, var) |
                    ("uniform", a, b) | ("empirical", loc, w) | ("mog", means, variances, weights)}
                   {"type": "binary", "p0", "variance"}
                   {"type": "categorical", "p", "W", "kappa", "variance"}
  ``depth``      max_interaction_depth
  ``variances``  order variances sigma^2_0..depth (or [sigma^2_0] when not sharing)
  ``share_var``  share_var_across_orders
and optionally ``X``, ``Z``, ``y``, ``noise``.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np


def build_kernel(cfg: Dict):
    """Product-side ``OAKKernel`` with the configuration's hyper-parameters assigned."""
    from .input_measures import EmpiricalMeasure, GaussianMeasure, MOGMeasure, UniformMeasure
    from .oak_kernel import OAKKernel
    from .ortho_binary_kernel import OrthogonalBinary
    from .ortho_categorical_kernel import OrthogonalCategorical
    from .ortho_rbf_kernel import RBF, OrthogonalRBFKernel

    dims = cfg["dims"]
    D = len(dims)
    share = cfg.get("share_var", True)
    k = OAKKernel([RBF] * D, num_dims=D, max_interaction_depth=cfg["depth"], constrain_orthogonal=True,
                  share_var_across_orders=share)
    subs = []
    for d, dc in enumerate(dims):
        if dc["type"] == "rbf":
            base = RBF(variance=dc.get("variance", 1.0), lengthscales=dc["lengthscale"])
            m = dc.get("measure", ("gaussian", 0.0, 1.0))
            if m is None:
                base.active_dims = [d]
                subs.append(base)
                continue
            kind = m[0]
            if kind == "gaussian":
                meas = GaussianMeasure(m[1], m[2])
            elif kind == "uniform":
                meas = UniformMeasure(m[1], m[2])
            elif kind == "empirical":
                meas = EmpiricalMeasure(np.asarray(m[1]).reshape(-1, 1), np.asarray(m[2]).reshape(-1, 1))
            elif kind == "mog":
                meas = MOGMeasure(np.asarray(m[1]), np.asarray(m[2]), np.asarray(m[3]))
            else:
                raise ValueError(kind)
            subs.append(OrthogonalRBFKernel(base, meas, active_dims=[d]))
        elif dc["type"] == "binary":
            kb = OrthogonalBinary(p0=dc["p0"], active_dims=[d])
            kb.variance.assign(dc.get("variance", 1.0))
            subs.append(kb)
        elif dc["type"] == "categorical":
            kc = OrthogonalCategorical(p=np.asarray(dc["p"]).reshape(-1, 1), rank=np.asarray(dc["W"]).shape[1],
                                       active_dims=[d])
            kc.W.assign(dc["W"])
            kc.kappa.assign(dc["kappa"])
            kc.variance.assign(dc.get("variance", 1.0))
            subs.append(kc)
        else:
            raise ValueError(dc["type"])
    k.kernels = subs
    for prm, v in zip(k.variances, cfg["variances"]):
        prm.assign(v)
    return k


def _gauss_dims(ls) -> List[Dict]:
    return [{"type": "rbf", "lengthscale": float(l), "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)} for l in ls]


def config_A(n: int = 1030) -> Dict:
    """GPR on concrete-shaped data: N=1030, D=8, Gaussian measure, full depth 8."""
    rng = np.random.default_rng(1030)
    X = rng.standard_normal((n, 8))
    y = X[:, 0] ** 2 + 2 * X[:, 1] + X[:, 0] * X[:, 1] + np.sin(X[:, 2]) + 0.1 * rng.standard_normal(n)
    y = ((y - y.mean()) / y.std()).reshape(-1, 1)
    ls = rng.uniform(0.5, 3.0, 8)
    return dict(name="A", X=X, y=y, dims=_gauss_dims(ls), depth=8, variances=[2.0 ** (-i) for i in range(9)],
                share_var=True, noise=0.01)


def config_B(n: int = 65536, D: int = 16, depth: int = 4) -> Dict:
    """Gram sweep: K(X, X), D=16, depth 4."""
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, D))
    ls = 0.5 + 2.5 * np.arange(D) / max(D - 1, 1)
    return dict(name="B", X=X, dims=_gauss_dims(ls), depth=depth, variances=[1.0, 1.0, 0.5, 0.25, 0.125][: depth + 1],
                share_var=True)


def config_C(n: int = 1_000_000, D: int = 20, m: int = 1024, depth: int = 3) -> Dict:
    """SGPR ELBO: N=1M, D=20, M=1024 inducing points (Z = X[:M]), depth 3."""
    rng = np.random.default_rng(20)
    X = rng.standard_normal((n, D))
    y = np.sin(X).sum(1) / np.sqrt(D) + X[:, 0] * X[:, 1] + 0.1 * rng.standard_normal(n)
    ls = rng.uniform(1.0, 4.0, D)
    return dict(name="C", X=X, y=y.reshape(-1, 1), Z=X[:m].copy(), dims=_gauss_dims(ls), depth=depth,
                variances=[1.0, 1.0, 0.5, 0.25][: depth + 1], share_var=True, noise=0.01)


def config_D(n: int = 50_000, m: int = 512) -> Dict:
    """Mixed inputs: 6 Gaussian-measure + 2 empirical-measure continuous, 2 binary, 2 categorical; depth 2."""
    rng = np.random.default_rng(12)
    X = np.zeros((n, 12))
    X[:, :6] = rng.standard_normal((n, 6))
    dims = _gauss_dims(rng.uniform(0.5, 3.0, 6))
    for j in (6, 7):
        col = np.round(8 * rng.standard_normal(n)) / 8
        col = (col - col.mean()) / col.std()
        X[:, j] = col
        loc, cnt = np.unique(col, return_counts=True)
        dims.append({"type": "rbf", "lengthscale": float(rng.uniform(0.5, 3.0)), "variance": 1.0,
                     "measure": ("empirical", loc, cnt / cnt.sum())})
    for j, pr in zip((8, 9), (0.3, 0.5)):
        X[:, j] = (rng.random(n) < pr).astype(np.float64)
        dims.append({"type": "binary", "p0": float(1 - X[:, j].mean()), "variance": 1.0})
    for j, C in zip((10, 11), (4, 6)):
        X[:, j] = rng.integers(0, C, n).astype(np.float64)
        _, cnt = np.unique(X[:, j], return_counts=True)
        dims.append({"type": "categorical", "p": cnt / n, "W": rng.uniform(0, 1, (C, 2)), "kappa": np.ones(C),
                     "variance": 1.0})
    y = np.sin(X[:, 0]) + X[:, 8] * X[:, 1] + 0.3 * X[:, 10] + 0.1 * rng.standard_normal(n)
    y = ((y - y.mean()) / y.std()).reshape(-1, 1)
    return dict(name="D", X=X, y=y, Z=X[:m].copy(), dims=dims, depth=2, variances=[1.0, 1.0, 0.5], share_var=True,
                noise=0.01)


def config_E(n: int = 200_000, D: int = 50, m: int = 512) -> Dict:
    """Sobol indices on a high-dimensional model: D=50, depth 2, N=200k, Z = X[:512]."""
    rng = np.random.default_rng(50)
    X = rng.standard_normal((n, D))
    y = X[:, 0] ** 2 + 2 * X[:, 1] + X[:, 0] * X[:, 1] + np.sin(X[:, 2:6]).sum(1) + 0.1 * rng.standard_normal(n)
    y = ((y - y.mean()) / y.std()).reshape(-1, 1)
    ls = rng.uniform(1.0, 4.0, D)
    return dict(name="E", X=X, y=y, Z=X[:m].copy(), dims=_gauss_dims(ls), depth=2, variances=[1.0, 1.0, 0.5],
                share_var=True, noise=0.01)
