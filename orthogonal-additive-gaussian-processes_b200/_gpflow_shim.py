"""Minimal stand-in for the parts of the gpflow 2.2.1 object model the OAK hot path touches
(``Parameter`` with positive / sigmoid transforms, ``Kernel.__call__`` / ``slice`` /
``active_dims``, parameter discovery).  gpflow / TensorFlow are not installable next to this
package; when a maintainer wires the kernels into real gpflow, only this file is replaced
(see INTEGRATION.md).  Host-side bookkeeping only -- no kernel arithmetic lives here.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

DEFAULT_JITTER = 1e-6  # gpflow.config.default_jitter()
DEFAULT_POSITIVE_MINIMUM = 0.0  # gpflow.config.default_positive_minimum()
VARIANCE_LOWER_BOUND = 1e-6  # gpflow.likelihoods DEFAULT_VARIANCE_LOWER_BOUND


def default_float():
    return np.float64


def default_jitter() -> float:
    return DEFAULT_JITTER


# ---- transforms -------------------------------------------------------------------------
class Identity:
    def forward(self, u):
        return u

    def inverse(self, v):
        return v


class Softplus:
    """``gpflow.utilities.positive()``: softplus shifted by the positive minimum."""

    def __init__(self, lower: float = DEFAULT_POSITIVE_MINIMUM):
        self.lower = lower

    def forward(self, u):
        return np.logaddexp(0.0, u) + self.lower

    def inverse(self, v):
        y = np.asarray(v, dtype=np.float64) - self.lower
        return y + np.log(-np.expm1(-y))


class Sigmoid:
    """``tfp.bijectors.Sigmoid(low, high)`` as used by ``bounded_param`` (oak/oak_kernel.py:24-33)."""

    def __init__(self, low: float, high: float):
        self.low, self.high = float(low), float(high)

    def forward(self, u):
        return self.low + (self.high - self.low) / (1.0 + np.exp(-np.asarray(u, dtype=np.float64)))

    def inverse(self, v):
        y = (np.asarray(v, dtype=np.float64) - self.low) / (self.high - self.low)
        return np.log(y) - np.log1p(-y)


def positive(lower: Optional[float] = None):
    return Softplus(DEFAULT_POSITIVE_MINIMUM if lower is None else lower)


class Parameter:
    """Constrained parameter holding its unconstrained representation (gpflow.Parameter)."""

    def __init__(self, value, transform=None, trainable: bool = True, prior=None, dtype=np.float64, name=None):
        if isinstance(value, Parameter):
            value = value.numpy()
        self.transform = transform if transform is not None else Identity()
        self.trainable = trainable
        self.prior = prior
        self.name = name
        self._set_u(self.transform.inverse(np.asarray(value, dtype=np.float64)))

    def _set_u(self, u):
        self._u = np.asarray(u, dtype=np.float64)
        self._v = None       # constrained value, computed on demand (a spec is packed on every evaluation)
        self._scalar = None

    # gpflow API subset
    def numpy(self):
        if self._v is None:
            self._v = np.array(self.transform.forward(self._u), dtype=np.float64)
        return self._v.copy()

    def scalar(self) -> float:
        """float(squeeze(value)), cached until the next assignment."""
        if self._scalar is None:
            if self._v is None:
                self._v = np.array(self.transform.forward(self._u), dtype=np.float64)
            self._scalar = float(np.squeeze(self._v))
        return self._scalar

    def assign(self, value):
        self._set_u(self.transform.inverse(np.asarray(value, dtype=np.float64)))
        return self

    @property
    def unconstrained_variable(self):
        return self._u

    @unconstrained_variable.setter
    def unconstrained_variable(self, u):
        self._set_u(u)

    @property
    def shape(self):
        return self.numpy().shape

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(np.squeeze(self.numpy()))

    def __repr__(self):
        return f"Parameter({self.numpy()!r}, trainable={self.trainable})"

    # arithmetic delegates to the constrained value
    def __mul__(self, o):
        return self.numpy() * np.asarray(o)

    __rmul__ = __mul__

    def __add__(self, o):
        return self.numpy() + np.asarray(o)

    __radd__ = __add__

    def __sub__(self, o):
        return self.numpy() - np.asarray(o)

    def __rsub__(self, o):
        return np.asarray(o) - self.numpy()

    def __truediv__(self, o):
        return self.numpy() / np.asarray(o)

    def __rtruediv__(self, o):
        return np.asarray(o) / self.numpy()

    def __pow__(self, o):
        return self.numpy() ** o

    def __neg__(self):
        return -self.numpy()


def value_of(p) -> np.ndarray:
    """Constrained value of a Parameter / constant tensor / python scalar as float64 ndarray."""
    if isinstance(p, Parameter):
        return p.numpy()
    if hasattr(p, "detach"):
        return p.detach().cpu().numpy().astype(np.float64)
    return np.asarray(p, dtype=np.float64)


def scalar_of(p) -> float:
    if isinstance(p, Parameter):
        return p.scalar()
    return float(np.squeeze(value_of(p)))


def set_trainable(obj, flag: bool):
    for p in collect_parameters(obj):
        p.trainable = flag


def collect_parameters(obj, _seen=None) -> List[Parameter]:
    """Parameter discovery in ``tf.Module._flatten`` order, which is the order of gpflow's ``parameters`` /
    ``trainable_parameters`` and therefore of the ``hyperparams`` array in the reference's checkpoints
    (oak/model_utils.py:44-87): attributes in sorted-name order, lists / tuples / dicts flattened in place,
    a module's own Parameters first ("walk direct properties first then recurse"), then its sub-modules in the
    order they were met, each visited once."""
    if _seen is None:
        _seen = set()
    if isinstance(obj, Parameter):
        if id(obj) in _seen:
            return []
        _seen.add(id(obj))
        return [obj]
    out: List[Parameter] = []
    submodules: list = []

    def leaves(value):
        if isinstance(value, (list, tuple)):
            for v in value:
                yield from leaves(v)
        elif isinstance(value, dict):
            for k in sorted(value):
                yield from leaves(value[k])
        else:
            yield value

    roots = vars(obj) if isinstance(obj, Module) else {"": obj} if isinstance(obj, (list, tuple, dict)) else {}
    if isinstance(obj, Module):
        _seen.add(id(obj))
    for key in sorted(roots):
        if key.startswith("_"):
            continue
        for leaf in leaves(roots[key]):
            if id(leaf) in _seen:
                continue
            if isinstance(leaf, Parameter):
                _seen.add(id(leaf))
                out.append(leaf)
            elif isinstance(leaf, Module):
                _seen.add(id(leaf))
                submodules.append(leaf)
    for sub in submodules:
        _seen.discard(id(sub))  # visited below; the mark above only de-duplicates the queue
        out += collect_parameters(sub, _seen)
    return out


class Module:
    @property
    def parameters(self):
        return tuple(collect_parameters(self))

    @property
    def trainable_parameters(self):
        return tuple(p for p in collect_parameters(self) if p.trainable)


class Kernel(Module):
    """``gpflow.kernels.Kernel`` protocol: ``active_dims``, ``slice``, ``__call__``."""

    def __init__(self, active_dims=None, name=None):
        self._active_dims = self._normalize_active_dims(active_dims)
        self.name = name

    @staticmethod
    def _normalize_active_dims(value):
        if value is None:
            value = slice(None, None, None)
        if isinstance(value, range):
            value = list(value)
        if not isinstance(value, slice):
            value = np.array(value, dtype=int).reshape(-1)
        return value

    @property
    def active_dims(self):
        return self._active_dims

    @active_dims.setter
    def active_dims(self, value):
        self._active_dims = self._normalize_active_dims(value)

    def slice(self, X, X2=None):
        dims = self.active_dims
        if isinstance(dims, slice):
            X = X[..., dims]
            if X2 is not None:
                X2 = X2[..., dims]
        else:
            idx = dims.tolist() if isinstance(dims, np.ndarray) else dims
            X = X[..., idx]
            if X2 is not None:
                X2 = X2[..., idx]
        return X, X2

    def K(self, X, X2=None):  # pragma: no cover - abstract
        raise NotImplementedError

    def K_diag(self, X):  # pragma: no cover - abstract
        raise NotImplementedError

    def __call__(self, X, X2=None, *, full_cov: bool = True, presliced: bool = False):
        if (not full_cov) and (X2 is not None):
            raise ValueError("Ambiguous inputs: `not full_cov` and `X2` are not compatible.")
        if not presliced:
            X, X2 = self.slice(X, X2)
        if not full_cov:
            return self.K_diag(X)
        return self.K(X, X2)


class InducingPoints(Module):
    def __init__(self, Z):
        self.Z = Parameter(np.asarray(Z, dtype=np.float64))

    @property
    def num_inducing(self):
        return self.Z.numpy().shape[0]

    def __len__(self):
        return self.num_inducing


class Gaussian(Module):
    """Gaussian likelihood: only the noise variance is needed on this path."""

    def __init__(self, variance: float = 1.0):
        self.variance = Parameter(variance, transform=positive(lower=VARIANCE_LOWER_BOUND))


class InvLink:
    """Inverse link of the Bernoulli likelihood, evaluated on the device by ``oak_bernoulli_quadrature_f64``.
    ``kind`` 0: ``sigmoid(x) (1 - 2 jitter) + jitter`` -- the ``inv_logit`` the reference defines in
    examples/uci/uci_classification_train.py:43-45; 1: gpflow's ``inv_probit``."""

    def __init__(self, kind: int, jitter: float = 1e-3):
        self.kind, self.jitter = int(kind), float(jitter)

    def __call__(self, x):
        from math import sqrt

        x = np.asarray(x, dtype=np.float64)
        if self.kind == 0:
            e = np.exp(-np.abs(x))
            base = np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))
        else:
            from scipy.special import erf

            base = 0.5 * (1.0 + erf(x / sqrt(2.0)))
        return base * (1.0 - 2.0 * self.jitter) + self.jitter


inv_logit = InvLink(0)
inv_probit = InvLink(1)


class Bernoulli(Module):
    """``gpflow.likelihoods.Bernoulli(invlink=...)``: 20-point Gauss-Hermite quadrature for the variational
    expectations and the predictive density (gpflow 2.2.1 ScalarLikelihood defaults)."""

    def __init__(self, invlink: InvLink = inv_probit, num_gauss_hermite_points: int = 20):
        if not isinstance(invlink, InvLink):
            raise NotImplementedError("invlink must be oak_b200 inv_logit / inv_probit (an InvLink): arbitrary "
                                      "Python callables cannot run inside the CUDA quadrature kernel")
        self.invlink = invlink
        self.num_gauss_hermite_points = int(num_gauss_hermite_points)


class Gamma:
    """Prior placeholder for ``tfd.Gamma(concentration, rate)`` (oak/model_utils.py:163-165)."""

    def __init__(self, concentration: float, rate: float):
        self.concentration, self.rate = float(concentration), float(rate)

    def log_prob(self, x):
        from math import lgamma

        x = np.asarray(x, dtype=np.float64)
        a, b = self.concentration, self.rate
        return a * np.log(b) - lgamma(a) + (a - 1.0) * np.log(x) - b * x
