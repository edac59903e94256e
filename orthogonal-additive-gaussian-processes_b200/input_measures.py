"""Input measures: parameter carriers for the constrained kernels.

Mirrors the reference's ``oak/input_measures.py:16-78`` (same class names, constructor
arguments, attributes and validation); the arithmetic that consumes them lives in
``csrc/oak_spec.cu`` / ``csrc/oak_prepare.cu``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


class Measure:
    pass


class UniformMeasure(Measure):
    """Uniform(a, b) input density."""

    def __init__(self, a: float, b: float):
        self.a, self.b = a, b


class GaussianMeasure(Measure):
    """N(mu, var) input density."""

    def __init__(self, mu: float, var: float):
        self.mu, self.var = mu, var


class EmpiricalMeasure(Measure):
    """Weighted Dirac measure on ``location`` (M,1); weights default to 1/M and must sum to 1."""

    def __init__(self, location: np.ndarray, weights: Optional[np.ndarray] = None):
        location = np.asarray(location)
        self.location = location
        if weights is None:
            weights = np.full((location.shape[0], 1), 1.0 / len(location))
        weights = np.asarray(weights)
        if not np.isclose(weights.sum(), 1.0, atol=1e-6):
            raise AssertionError(f"not close to 1 {weights.sum()}")
        self.weights = weights


class MOGMeasure(Measure):
    """Mixture of K one-dimensional Gaussians."""

    def __init__(self, means: np.ndarray, variances: np.ndarray, weights: np.ndarray):
        means, variances, weights = np.asarray(means), np.asarray(variances), np.asarray(weights)
        if not (means.ndim == variances.ndim == weights.ndim == 1 and len(means) == len(variances) == len(weights)):
            raise ValueError("means, variances and weights must be vectors of one common length K")
        if not np.isclose(weights.sum(), 1.0, atol=1e-6):
            raise AssertionError(f"Weights not close to 1 {weights.sum()}")
        self.means, self.variances, self.weights = means.astype(float), variances.astype(float), weights
