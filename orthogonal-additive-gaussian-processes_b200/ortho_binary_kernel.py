"""Constrained binary kernel; drop-in for ``OrthogonalBinary`` (``oak/ortho_binary_kernel.py:13-59``).
The 2x2 table is built in ``csrc/oak_spec.cu`` and gathered inside the fused Gram tile."""
from __future__ import annotations

import numpy as np

from . import _cabi
from ._cabi import DimSpec
from ._gpflow_shim import Parameter, positive, scalar_of
from ._native_kernel import NativeKernel


class OrthogonalBinary(NativeKernel):
    """
    :param p0: probability of binary measure, P(x = 0)
    :param active_dims: active dimension along which the kernel is to be applied
    """

    def __init__(self, p0: float = 0.5, active_dims=None):
        super().__init__(active_dims=active_dims)
        self.variance = Parameter(1.0, transform=positive())
        self.p0 = p0

    def output_covariance(self):
        """sigma^2 [[p1^2, -p0 p1], [-p0 p1, p0^2]] (oak/ortho_binary_kernel.py:29-33) -- parameter view."""
        p0 = float(self.p0)
        p1 = 1.0 - p0
        return np.array([[p1 * p1, -p0 * p1], [-p0 * p1, p0 * p0]]) * scalar_of(self.variance)

    def output_variance(self):
        p0 = float(self.p0)
        p1 = 1.0 - p0
        return np.array([p1 * p1, p0 * p0]) * scalar_of(self.variance)

    def _dim_spec(self, column: int) -> DimSpec:
        return DimSpec(_cabi.DIM_BINARY, column, variance=scalar_of(self.variance), m0=float(self.p0))

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
