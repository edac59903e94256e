"""The orthogonal additive kernel.

Drop-in for the reference's ``oak/oak_kernel.py`` (``bounded_param`` :24-33, ``OAKKernel`` :36-278,
``KernelComponenent`` :281-335, ``get_list_representation`` :338-364): same names, constructor
arguments, attribute paths and error behaviour.  ``K`` / ``K_diag`` do not build one matrix per
input dimension; they pack the current hyper-parameters into an ``oak_spec`` and call the fused
CUDA tiles (``csrc/oak_gram.cu``), which evaluate every per-dimension constrained kernel, the
Newton-Girard recurrence and the variance-weighted sum in registers.
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Tuple, Type

import numpy as np

from . import _cabi, _device
from ._gpflow_shim import Kernel, Parameter, Sigmoid, positive, scalar_of
from ._native_kernel import NativeKernel
from .input_measures import EmpiricalMeasure, GaussianMeasure, MOGMeasure
from .ortho_binary_kernel import OrthogonalBinary
from .ortho_categorical_kernel import OrthogonalCategorical
from .ortho_rbf_kernel import RBF, OrthogonalRBFKernel


def bounded_param(low: float, high: float, param: float) -> Parameter:
    """Parameter constrained to (low, high) through a sigmoid (oak/oak_kernel.py:24-33)."""
    return Parameter(param, transform=Sigmoid(low, high), dtype=np.float64)


class OAKKernel(NativeKernel):
    """
    Compute OAK kernel
    :param base_kernels: list of base kernel classes for non-binary inputs (``RBF`` or None)
    :param num_dims: dimensionality of input data
    :param max_interaction_depth: maximum order of interactions
    :param active_dims: list of single-column lists, one per sub-kernel
    :param constrain_orthogonal: whether to use the orthogonal (constrained) kernels
    :param p0: per-dim probability P(x=0) for binary inputs, None elsewhere
    :param p: per-dim category probabilities for categorical inputs, None elsewhere
    :param lengthscale_bounds: [low, high] common sigmoid bounds for every lengthscale
    :param empirical_locations / empirical_weights: per-dim empirical measure, None elsewhere
    :param gmm_measures: per-dim ``MOGMeasure`` or None
    :param share_var_across_orders: one variance per interaction order (OAK) if True,
           else the constrained kernel prod_i (1 + k_i) with one variance for the constant
    """

    def __init__(
        self,
        base_kernels: List[Optional[Type[Kernel]]],
        num_dims: int,
        max_interaction_depth: int,
        active_dims: Optional[List[List[int]]] = None,
        constrain_orthogonal: bool = False,
        p0: Optional[List[float]] = None,
        p: Optional[List[float]] = None,
        lengthscale_bounds: Optional[List[float]] = None,
        empirical_locations: Optional[List[float]] = None,
        empirical_weights: Optional[List[float]] = None,
        gmm_measures: Optional[List[MOGMeasure]] = None,
        share_var_across_orders: Optional[bool] = True,
    ):
        super().__init__(active_dims=range(num_dims))
        if active_dims is None:
            active_dims = [[dim] for dim in range(num_dims)]
        flat_dims = [dim for sublist in active_dims for dim in sublist]
        # same checks as oak_kernel.py:79-82 (the first one is off by one there; kept)
        assert max(flat_dims) <= num_dims, "Active dims exceeding num dims."
        assert len(flat_dims) == len(np.unique(flat_dims)), "Active dims contains duplicates."
        if any(len(a) != 1 for a in active_dims):
            raise ValueError("every sub-kernel acts on exactly one input column (SURVEY.md 2.2 quirk 7)")
        if max_interaction_depth > _cabi.OAK_MAX_DEPTH:
            raise ValueError(f"max_interaction_depth > {_cabi.OAK_MAX_DEPTH} is not supported by the CUDA tiles")

        delta2 = 1  # prior measure variance hard-coded to 1 (oak_kernel.py:84)
        self.base_kernels, self.max_interaction_depth = base_kernels, max_interaction_depth
        self.share_var_across_orders = share_var_across_orders
        n_k = len(active_dims)
        if p0 is None:
            p0 = [None] * n_k
        if p is None:
            p = [None] * n_k

        self.kernels = []
        if constrain_orthogonal:
            if empirical_locations is None:
                assert empirical_weights is None, "Cannot have weights without locations"
                empirical_locations = [None] * n_k
                empirical_weights = [None] * n_k
            elif empirical_weights is not None:
                loc_shapes = [None if l is None else len(l) for l in empirical_locations[:n_k]]
                w_shapes = [None if w is None else len(w) for w in empirical_weights[:n_k]]
                assert loc_shapes == w_shapes, (
                    f"Shape of empirical measure locations {loc_shapes} do not match weights {w_shapes}"
                )
            else:
                empirical_weights = [None] * n_k
            if gmm_measures is None:
                gmm_measures = [None] * n_k

            for dim in range(n_k):
                if empirical_locations[dim] is not None and gmm_measures[dim] is not None:
                    raise ValueError(f"Both empirical and GMM measure defined for input {dim}")
                if (p0[dim] is None) and (p[dim] is None):
                    if empirical_locations[dim] is not None:
                        k = OrthogonalRBFKernel(
                            base_kernels[dim](),
                            EmpiricalMeasure(empirical_locations[dim], empirical_weights[dim]),
                            active_dims=active_dims[dim],
                        )
                    elif gmm_measures[dim] is not None:
                        k = OrthogonalRBFKernel(base_kernels[dim](), measure=gmm_measures[dim],
                                                active_dims=active_dims[dim])
                    else:
                        k = OrthogonalRBFKernel(base_kernels[dim](), GaussianMeasure(0, delta2),
                                                active_dims=active_dims[dim])
                        if share_var_across_orders:
                            # constant, non-trainable unit variance (oak_kernel.py:163-166)
                            k.base_kernel.variance = np.ones(1, dtype=np.float64)
                    if lengthscale_bounds is not None:
                        k.base_kernel.lengthscales = bounded_param(lengthscale_bounds[0], lengthscale_bounds[1], 1)
                elif p[dim] is not None:
                    assert base_kernels[dim] is None
                    k = OrthogonalCategorical(p=p[dim], active_dims=active_dims[dim])
                    if share_var_across_orders:
                        k.variance = np.ones(1, dtype=np.float64)
                else:
                    assert base_kernels[dim] is None
                    k = OrthogonalBinary(p0=p0[dim], active_dims=active_dims[dim])
                    if share_var_across_orders:
                        k.variance = np.ones(1, dtype=np.float64)
                self.kernels.append(k)
        else:
            # unconstrained kernels with the additive structure (oak_kernel.py:191-210)
            assert empirical_locations is None, "Cannot have empirical locations without orthogonal constraint"
            assert empirical_weights is None, "Cannot have empirical weights without orthogonal constraint"
            for dim in range(n_k):
                if p0[dim] is None:
                    k = base_kernels[dim](active_dims=active_dims[dim])
                else:
                    assert base_kernels[dim] is None
                    k = OrthogonalBinary(p0=p0[dim], active_dims=active_dims[dim])
                if share_var_across_orders:
                    k.variance = np.ones(1, dtype=np.float64)
                self.kernels.append(k)

        # variances of the interaction orders (+1 for the constant term) (oak_kernel.py:212-221)
        n_var = max_interaction_depth + 1 if self.share_var_across_orders else 1
        self.variances = [Parameter(1.0, transform=positive()) for _ in range(n_var)]

    # ---- spec packing ----------------------------------------------------------------
    def _column_of(self, k) -> int:
        dims = k.active_dims
        if isinstance(dims, slice):
            raise ValueError("sub-kernels of OAKKernel need explicit active_dims")
        return int(np.asarray(dims).reshape(-1)[0])

    def _dim_specs(self):
        return [k._dim_spec(self._column_of(k)) for k in self.kernels]

    def _depth(self) -> int:
        return int(self.max_interaction_depth)

    def _order_variances(self):
        return [scalar_of(v) for v in self.variances]

    def _share_var(self) -> bool:
        return bool(self.share_var_across_orders)

    # ---- reference API ---------------------------------------------------------------
    def compute_additive_terms(self, kernel_matrices):
        """Elementary symmetric polynomials e_0..e_P of a list of matrices via Newton-Girard
        (oak/oak_kernel.py:223-249).  On the hot path this recurrence runs inside the CUDA tile
        epilogue; this host method exists for API parity with callers that pass explicit
        matrices (tests/test_kernel_properties.py:70-86) and accepts NumPy or torch inputs."""
        import torch

        mats = list(kernel_matrices)
        host = _device.is_host(mats[0])
        stack = torch.stack([_device.to_device(m, ndim=0) for m in mats])
        e = _device.additive_terms(stack, int(self.max_interaction_depth))
        return [_device.from_device(t, host) for t in e]


class KernelComponenent(NativeKernel):
    """One additive component sigma^2_|S| prod_{d in S} k_d (oak/oak_kernel.py:281-335)."""

    def __init__(self, oak_kernel: OAKKernel, iComponent_list: List[int],
                 share_var_across_orders: Optional[bool] = True):
        super().__init__(active_dims=oak_kernel.active_dims)
        self.oak_kernel = oak_kernel
        self.iComponent_list = iComponent_list
        self.share_var_across_orders = share_var_across_orders
        self.kernels = [k for i, k in enumerate(self.oak_kernel.kernels) if i in self.iComponent_list]

    def _make_spec(self):
        """Spec of the parent kernel whose order variances implement (:313-318, :329-334):
        sigma^2_n when sharing, 1 otherwise; the constant component always uses variances[0]."""
        ok = self.oak_kernel
        n = len(self.iComponent_list)
        depth = max(n, 1)
        var = [0.0] * (depth + 1)
        if n == 0:
            var[0] = scalar_of(ok.variances[0])
        else:
            var[n] = scalar_of(ok.variances[n]) if self.share_var_across_orders else 1.0
        return _cabi.Spec(ok._dim_specs(), depth, var, True, _cabi.ESP_NEWTON_GIRARD, stream=_device.stream_ptr())

    def _subset(self):
        return sorted(int(i) for i in self.iComponent_list)

    def _K_device(self, Xd, X2d=None, spec=None):
        spec = self._make_spec()
        try:
            self._check_discrete(Xd, spec._keep)
            px = _device.Points(spec, Xd)
            px2 = None
            if X2d is not None:
                self._check_discrete(X2d, spec._keep)
                px2 = _device.Points(spec, X2d)
            return _device.component_gram(spec, self._subset(), px, px2)
        finally:
            spec.close()

    def _K_diag_device(self, Xd, spec=None):
        # diagonal of the component: computed as the diagonal of the component Gram on the
        # device (O(N) kernel launch over the prepared points)
        spec = self._make_spec()
        try:
            self._check_discrete(Xd, spec._keep)
            px = _device.Points(spec, Xd)
            return _device.component_diag(spec, self._subset(), px)
        finally:
            spec.close()


def get_list_representation(
    kernel: OAKKernel, num_dims: int, share_var_across_orders: Optional[bool] = True
) -> Tuple[List[List[int]], List[KernelComponenent]]:
    """List representation of the OAK kernel: all subsets of dims up to the interaction depth,
    ordered by order then lexicographically (oak/oak_kernel.py:338-364)."""
    assert isinstance(kernel, OAKKernel)
    selected_dims: List[List[int]] = [[]]
    kernel_list = [KernelComponenent(kernel, [], share_var_across_orders=share_var_across_orders)]
    if kernel.max_interaction_depth > 0:
        for order in range(1, kernel.max_interaction_depth + 1):
            combos = [list(t) for t in itertools.combinations(np.arange(num_dims), order)]
            selected_dims = selected_dims + combos
            for c in combos:
                # NB the reference drops share_var_across_orders here (:362); kept
                kernel_list.append(KernelComponenent(kernel, c))
    return selected_dims, kernel_list
