"""Input measures: parameter carriers for the constrained kernels.

Mirrors the reference's ``oak/input_measures.py:16-78`` (same class names, constructor
arguments, attributes and validation); the arithmetic that consumes them lives in
``csrc/oak_spec.cu`` / ``csrc/oak_prepare.cu``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def _weights_sum_to_one(weights: np.ndarray, what: str) -> np.ndarray:
    weights = np.asarray(weights)
    total = weights.sum()
    if not np.isclose(total, 1.0, atol=1e-6):
        raise AssertionError(f"{what} {total}")  # the reference asserts
    return weights


class Measure:
    """Base class; the kernels dispatch on the concrete type."""


class GaussianMeasure(Measure):
    """N(mu, var) input density."""

    def __init__(self, mu: float, var: float):
        self.mu = mu
        self.var = var


class UniformMeasure(Measure):
    """Uniform(a, b) input density."""

    def __init__(self, a: float, b: float):
        self.a = a
        self.b = b


class MOGMeasure(Measure):
    """Mixture of K one-dimensional Gaussians: vectors ``means``, ``variances``, ``weights`` of length K."""

    def __init__(self, means: np.ndarray, variances: np.ndarray, weights: np.ndarray):
        parts = [np.asarray(v) for v in (means, variances, weights)]
        if any(v.ndim != 1 for v in parts) or len({len(v) for v in parts}) != 1:
            raise ValueError("means, variances and weights must be vectors of one common length K")
        self.weights = _weights_sum_to_one(parts[2], "Weights not close to 1")
        self.means = parts[0].astype(float)
        self.variances = parts[1].astype(float)


class EmpiricalMeasure(Measure):
    """Weighted Dirac measure on ``location`` (M, 1); weights default to 1/M and must sum to 1."""

    def __init__(self, location: np.ndarray, weights: Optional[np.ndarray] = None):
        self.location = np.asarray(location)
        m = self.location.shape[0]
        self.weights = _weights_sum_to_one(np.full((m, 1), 1.0 / m) if weights is None else weights,
                                           "not close to 1")
