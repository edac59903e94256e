"""Constrained coregional kernel for categorical inputs; drop-in for ``OrthogonalCategorical``
(``oak/ortho_categorical_kernel.py:14-74``).  ``B = sigma^2 (A - Ap (Ap)^T / p^T A p)`` with
``A = W W^T + diag(kappa)`` is built in ``csrc/oak_spec.cu`` and gathered inside the Gram tile."""
from __future__ import annotations

from typing import List

import numpy as np

from . import _cabi
from ._cabi import DimSpec
from ._gpflow_shim import Parameter, positive, scalar_of, value_of
from ._native_kernel import NativeKernel


class OrthogonalCategorical(NativeKernel):
    """
    :param p: probability measure of the categories, p_i = Prob(X = i), summing to 1
    :param rank: number of degrees of correlation between the outputs (columns of W)
    :param active_dims: active dimension of input to apply this kernel to
    """

    def __init__(self, p: List, rank: int = 2, active_dims=None):
        super().__init__(active_dims=active_dims)
        num_cat = len(p)
        self.num_cat = num_cat
        self.p = p
        self.variance = Parameter(1.0, transform=positive())
        # reference: W ~ tf.random.uniform (float32 values), kappa = ones  (:28-32)
        self.W = Parameter(np.random.uniform(size=(num_cat, rank)).astype(np.float32).astype(np.float64))
        self.kappa = Parameter(np.ones(num_cat), transform=positive())

    def _p_vector(self):
        return np.asarray(self.p, dtype=np.float64).reshape(-1)

    def output_covariance(self):
        """Parameter view of B (oak/ortho_categorical_kernel.py:34-42); not used on the hot path."""
        W, kappa, p = value_of(self.W), value_of(self.kappa), self._p_vector().reshape(-1, 1)
        A = W @ W.T + np.diag(kappa)
        Ap = A @ p
        return (A - (Ap @ Ap.T) / (p.T @ Ap)[0]) * scalar_of(self.variance)

    def output_variance(self):
        W, kappa, p = value_of(self.W), value_of(self.kappa), self._p_vector().reshape(-1, 1)
        A = W @ W.T + np.diag(kappa)
        Ap = A @ p
        A_diag = np.sum(np.square(W), 1) + kappa
        return (A_diag - np.sum(np.square(Ap), 1) / (p.T @ Ap)[0]) * scalar_of(self.variance)

    def _dim_spec(self, column: int) -> DimSpec:
        W = value_of(self.W)
        return DimSpec(_cabi.DIM_CATEGORICAL, column, variance=scalar_of(self.variance), v0=W,
                       v1=value_of(self.kappa), v2=self._p_vector(), rank=W.shape[1])

    def _dim_specs(self):
        return [self._dim_spec(0)]

    def K(self, X, X2=None):
        for a in (X, X2):
            if a is not None and (len(np.shape(a)) != 2 or np.shape(a)[1] != 1):
                raise ValueError(f"expected an (N, 1) input, got shape {tuple(np.shape(a))}")
        return super().K(X, X2)

    def K_diag(self, X):
        if len(np.shape(X)) != 2 or np.shape(X)[1] != 1:
            raise ValueError(f"expected an (N, 1) input, got shape {tuple(np.shape(X))}")
        return super().K_diag(X)
