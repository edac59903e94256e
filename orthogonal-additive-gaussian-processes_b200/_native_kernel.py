"""Shared host logic of every kernel class: turn the current parameter values into an
``oak_spec`` and run prepare + gram / gram_diag through the C ABI."""
from __future__ import annotations

from typing import List, Sequence


from . import _cabi, _device
from ._cabi import DimSpec
from ._gpflow_shim import Kernel

# Process-wide default for how e_n is formed.  The reference accumulates power sums and applies Newton-Girard
# (oak_kernel.py:236-249, kept as ESP_NEWTON_GIRARD and exercised by the same parity tests); the direct recurrence
# e_n += k_d e_{n-1} evaluates the same polynomial with P instead of S_pow(P) FP64 instructions per entry and
# dimension and without the cancellation of the Newton-Girard identities.  Measured on B200
# (profiles/r02au_ab_gram_esp_default.txt): depth 4 (config B) +3 %, depth 8 (config A's shape) +22 %, depth 2 equal.
_DEFAULT_ALGORITHM = _cabi.ESP_DIRECT


class NativeKernel(Kernel):
    """A kernel whose ``K`` / ``K_diag`` run as fused CUDA tiles.

    Sub-classes provide ``_dim_specs(columns)`` (one ``DimSpec`` per sub-kernel, acting on the
    given columns of the *sliced* input), ``_depth()`` and ``_order_variances()``.
    """

    esp_algorithm = None  # per-instance override of the elementary-symmetric-polynomial scheme

    # -- to be provided ---------------------------------------------------------------
    def _dim_specs(self) -> List[DimSpec]:
        raise NotImplementedError

    def _depth(self) -> int:
        return 1

    def _order_variances(self) -> Sequence[float]:
        return [0.0, 1.0]  # K = e_1 = the single sub-kernel itself

    def _share_var(self) -> bool:
        return True

    # -- native path ------------------------------------------------------------------
    def _make_spec(self) -> _cabi.Spec:
        algo = self.esp_algorithm if self.esp_algorithm is not None else _DEFAULT_ALGORITHM
        return _cabi.Spec(self._dim_specs(), self._depth(), self._order_variances(), self._share_var(), algo,
                          stream=_device.stream_ptr())

    def _check_discrete(self, Xd, dims: List[DimSpec]):
        cols = [d.column for d in dims if d.type != _cabi.DIM_RBF]
        cnts = [d.count for d in dims if d.type != _cabi.DIM_RBF]
        if cols:
            _device.validate_discrete(Xd, cols, cnts)

    def _K_device(self, Xd, X2d=None, spec=None):
        own = spec is None
        spec = self._make_spec() if own else spec
        try:
            self._check_discrete(Xd, spec._keep)
            px = _device.Points(spec, Xd)
            px2 = None
            if X2d is not None:
                self._check_discrete(X2d, spec._keep)
                px2 = _device.Points(spec, X2d)
            return _device.gram(spec, px, px2)
        finally:
            if own:
                spec.close()

    def _K_diag_device(self, Xd, spec=None):
        own = spec is None
        spec = self._make_spec() if own else spec
        try:
            self._check_discrete(Xd, spec._keep)
            return _device.gram_diag(spec, _device.Points(spec, Xd))
        finally:
            if own:
                spec.close()

    @staticmethod
    def _check_2d(X, name="X"):
        shp = tuple(X.shape)
        if len(shp) != 2:
            raise ValueError(f"{name} must be a matrix (N, D); got shape {shp}")

    def K(self, X, X2=None):
        host = _device.is_host(X)
        Xd = _device.to_device(X)
        self._check_2d(Xd)
        X2d = None
        if X2 is not None:
            X2d = _device.to_device(X2)
            self._check_2d(X2d, "X2")
            if X2d.shape[1] != Xd.shape[1]:
                raise ValueError("X and X2 must have the same number of columns")
        return _device.from_device(self._K_device(Xd, X2d), host)

    def K_diag(self, X):
        host = _device.is_host(X)
        Xd = _device.to_device(X)
        self._check_2d(Xd)
        return _device.from_device(self._K_diag_device(Xd), host)
