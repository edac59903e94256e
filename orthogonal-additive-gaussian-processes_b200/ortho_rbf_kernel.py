"""Constrained squared-exponential kernel for continuous inputs.

Drop-in for the reference's ``OrthogonalRBFKernel`` (``oak/ortho_rbf_kernel.py:20-177``):
``k~(x, y) = k(x, y) - cov_X_s(x) cov_X_s(y) / var_s()`` with the closed-form correction for the
Gaussian / uniform / empirical / mixture-of-Gaussians input measure.  All arithmetic runs in
``liboak_b200.so`` (``csrc/oak_prepare.cu``: per-point correction; ``csrc/oak_gram.cu``: tiles).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi, _device
from ._cabi import DimSpec
from ._gpflow_shim import Parameter, positive, scalar_of
from ._native_kernel import NativeKernel
from .input_measures import EmpiricalMeasure, GaussianMeasure, Measure, MOGMeasure, UniformMeasure


class RBF(NativeKernel):
    """Stand-in for ``gpflow.kernels.RBF`` on ONE input column: ``variance * exp(-(x-y)^2 / (2 l^2))``
    (call sites ``oak/ortho_rbf_kernel.py:107,116,169,176``; plain use ``oak/oak_kernel.py:203``)."""

    def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None, name=None):
        super().__init__(active_dims=active_dims, name=name)
        self.variance = Parameter(variance, transform=positive())
        self.lengthscales = Parameter(lengthscales, transform=positive())

    def _dim_spec(self, column: int, measure: Measure = None) -> DimSpec:
        l, s2 = scalar_of(self.lengthscales), scalar_of(self.variance)
        if measure is None:
            return DimSpec(_cabi.DIM_RBF, column, measure=_cabi.MEASURE_NONE, lengthscale=l, variance=s2)
        if isinstance(measure, GaussianMeasure):
            return DimSpec(_cabi.DIM_RBF, column, measure=_cabi.MEASURE_GAUSSIAN, lengthscale=l, variance=s2,
                           m0=float(measure.mu), m1=float(measure.var))
        if isinstance(measure, UniformMeasure):
            return DimSpec(_cabi.DIM_RBF, column, measure=_cabi.MEASURE_UNIFORM, lengthscale=l, variance=s2,
                           m0=float(measure.a), m1=float(measure.b))
        if isinstance(measure, EmpiricalMeasure):
            loc, w = np.asarray(measure.location, dtype=np.float64), np.asarray(measure.weights, dtype=np.float64)
            if loc.ndim != 2 or loc.shape[1] != 1 or w.shape != loc.shape:
                raise ValueError("empirical location and weights must both have shape (M, 1)")
            return DimSpec(_cabi.DIM_RBF, column, measure=_cabi.MEASURE_EMPIRICAL, lengthscale=l, variance=s2,
                           v0=loc, v1=w)
        if isinstance(measure, MOGMeasure):
            return DimSpec(_cabi.DIM_RBF, column, measure=_cabi.MEASURE_MOG, lengthscale=l, variance=s2,
                           v0=measure.means, v1=measure.variances, v2=measure.weights)
        raise NotImplementedError

    def _dim_specs(self):
        return [self._dim_spec(0)]

    def K(self, X, X2=None):
        if tuple(np.shape(X))[-1] != 1:
            raise ValueError("RBF stand-in acts on exactly one input column (SURVEY.md 2.2 quirk 7)")
        return super().K(X, X2)


SquaredExponential = RBF


class OrthogonalRBFKernel(NativeKernel):
    """
    :param base_kernel: base RBF kernel before applying orthogonality constraint
    :param measure: input measure
    :param active_dims: active dimension
    :return: constrained RBF kernel
    """

    def __init__(self, base_kernel: RBF, measure: Measure, active_dims=None):
        super().__init__(active_dims=active_dims)
        self.base_kernel, self.measure = base_kernel, measure
        if not isinstance(base_kernel, RBF):
            raise NotImplementedError
        if not isinstance(measure, (UniformMeasure, GaussianMeasure, EmpiricalMeasure, MOGMeasure)):
            raise NotImplementedError

    def _dim_spec(self, column: int) -> DimSpec:
        return self.base_kernel._dim_spec(column, self.measure)

    def _dim_specs(self):
        return [self._dim_spec(0)]

    @staticmethod
    def _one_column(X):
        if len(np.shape(X)) != 2 or np.shape(X)[1] != 1:
            raise ValueError(f"expected an (N, 1) input, got shape {tuple(np.shape(X))}")

    def K(self, X, X2=None):
        self._one_column(X)
        if X2 is not None:
            self._one_column(X2)
        return super().K(X, X2)

    def K_diag(self, X):
        self._one_column(X)
        return super().K_diag(X)

    # --- the correction terms, exposed like the reference's closures ------------------
    def var_s(self):
        """``var_s()`` (oak/ortho_rbf_kernel.py:65-78, 94-97, 109-120, 138-152)."""
        spec = self._make_spec()
        try:
            out = C.c_double(0.0)
            _cabi.check(_cabi.load().oak_spec_var_s_f64(spec.handle, 0, C.byref(out),
                                                        C.c_void_p(_device.stream_ptr())), "oak_spec_var_s_f64")
            return float(out.value)
        finally:
            spec.close()

    def cov_X_s(self, X):
        """``cov_X_s(X)`` -> (N, 1) (oak/ortho_rbf_kernel.py:49-63, 82-92, 101-107, 124-136)."""
        self._one_column(X)
        host = _device.is_host(X)
        Xd = _device.to_device(X)
        spec = self._make_spec()
        try:
            px = _device.Points(spec, Xd)
            n_pad = px.buf.numel() // 2
            chat = px.buf.view(n_pad, 2)[: px.n, 1]
            out = C.c_double(0.0)
            _cabi.check(_cabi.load().oak_spec_var_s_f64(spec.handle, 0, C.byref(out),
                                                        C.c_void_p(_device.stream_ptr())), "oak_spec_var_s_f64")
            c = (chat * float(np.sqrt(out.value))).reshape(-1, 1).clone()
            return _device.from_device(c, host)
        finally:
            spec.close()
