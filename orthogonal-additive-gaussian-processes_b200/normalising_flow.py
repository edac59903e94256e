"""One-dimensional normalising flow of the input pipeline (oak/normalising_flow.py:16-85).

Each continuous input column is pushed through

    x -> x - offset -> log -> + shift -> * scale -> SinhArcsinh(skewness, tailweight)

(``tfb.Chain([SinhArcsinh, Scale, Shift, Log, Shift(-offset)])``, :46-52; without the two log steps
when ``log=False``, :54-56) whose four parameters minimise ``KL_objective`` (:76-81) under L-BFGS-B
(``gpflow.optimizers.Scipy().minimize`` default, model_utils.py:313-317).  ``SinhArcsinh`` follows
tensorflow_probability 0.11 (the reference's pin, setup.py:33): ``sinh((arcsinh(x) + skewness) *
tailweight)`` -- later TFP releases add a tail-weight dependent multiplier.  The O(N) work runs on the
device (csrc/oak_flow.cu): every L-BFGS-B iteration is one ``oak_flow_objective_f64`` pass over the resident
column (objective and its four derivatives, deterministic reduction), the forward transform is
``oak_flow_forward_f64``.  ``inverse`` and ``forward_log_det_jacobian`` (plotting-side utilities on small
arrays in the reference) are NumPy.  The optimiser path is not TensorFlow's, so the fitted parameters
agree with the reference's only to optimiser tolerance; parity of the hot path is pinned on the post-flow
inputs.
"""
from __future__ import annotations

import numpy as np

from . import _device
from ._gpflow_shim import Module, Parameter


class _Exp:
    """``tfb.Exp()`` as a parameter transform."""

    def forward(self, u):
        return np.exp(u)

    def inverse(self, v):
        return np.log(np.asarray(v, dtype=np.float64))


class _Bijector:
    """The chain of :46-56 with ``forward`` / ``inverse`` / ``forward_log_det_jacobian`` and a
    TensorFlow-like ``__call__``."""

    def __init__(self, owner):
        self._o = owner

    def _pre(self, x):
        o = self._o
        x = np.asarray(x, dtype=np.float64)
        u = np.log(x - o.offset) if o.log else x
        return u, (u + float(o.shift.numpy())) * float(o.scale.numpy())

    def forward(self, x):
        """Device transform of a column: NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out."""
        o = self._o
        host = _device.is_host(x)
        xd = _device.to_device(np.asarray(x, dtype=np.float64).reshape(-1) if host else x.reshape(-1), ndim=1)
        y = _device.flow_forward(xd, o.offset, o.log, float(o.shift.numpy()), float(o.scale.numpy()),
                                 float(o.skewness.numpy()), float(o.tailweight.numpy()))
        return y.cpu().numpy().reshape(np.shape(x)) if host else y.reshape(x.shape)

    __call__ = forward

    def inverse(self, y):
        o = self._o
        y = np.asarray(y, dtype=np.float64)
        z = np.sinh(np.arcsinh(y) / float(o.tailweight.numpy()) - float(o.skewness.numpy()))
        u = z / float(o.scale.numpy()) - float(o.shift.numpy())
        return np.exp(u) + o.offset if o.log else u

    def forward_log_det_jacobian(self, x, event_ndims=0):
        o = self._o
        u, z = self._pre(x)
        tau = float(o.tailweight.numpy())
        w = (np.arcsinh(z) + float(o.skewness.numpy())) * tau
        ldj = np.log(np.cosh(w)) + np.log(tau) - 0.5 * np.log1p(z * z) + np.log(float(o.scale.numpy()))
        return ldj - u if o.log else ldj


class Normalizer(Module):
    """
    :param x: input to transform
    :param log: whether to log x first before applying flows of transformations
    :return: flows of transformations to match x to standard Gaussian
    """

    def __init__(self, x, log=True, **kwargs):
        self.x = np.asarray(x, dtype=np.float64).reshape(-1)
        self._xd = None
        self.log = bool(log)
        self.offset = float(np.min(self.x) - 1.0) if self.log else 0.0
        base = np.log(self.x - self.offset) if self.log else self.x
        # make_sinharcsinh (:16-20) and make_standardizer (:23-27)
        self.skewness = Parameter(0.0)
        self.tailweight = Parameter(1.0, transform=_Exp())
        self.scale = Parameter(1.0 / np.std(base), transform=_Exp())
        self.shift = Parameter(-np.mean(base))
        self.bijector = _Bijector(self)

    # ---- objective (:76-81) and its gradient in the unconstrained variables ----------------------
    def _theta(self):
        return np.array([float(self.scale.unconstrained_variable), float(self.shift.unconstrained_variable),
                         float(self.skewness.unconstrained_variable), float(self.tailweight.unconstrained_variable)])

    def _device_column(self):
        if self._xd is None:
            self._xd = _device.to_device(self.x, ndim=1)
        return self._xd

    def _objective_and_grad(self, theta):
        """theta = (log scale, shift, skewness, log tailweight): one device pass over the column."""
        out = _device.flow_objective(self._device_column(), self.offset, self.log, theta)
        return float(out[0]), out[1:].copy()

    def KL_objective(self) -> float:
        return self._objective_and_grad(self._theta())[0]

    def fit(self, maxiter: int = 1000):
        """``gpflow.optimizers.Scipy().minimize(n.KL_objective, n.trainable_variables)`` (L-BFGS-B)."""
        from scipy.optimize import minimize

        res = minimize(self._objective_and_grad, self._theta(), jac=True, method="L-BFGS-B", options=dict(maxiter=maxiter))
        self._xd = None  # the device copy of the column is only needed while fitting
        self.scale.unconstrained_variable = np.asarray(res.x[0])
        self.shift.unconstrained_variable = np.asarray(res.x[1])
        self.skewness.unconstrained_variable = np.asarray(res.x[2])
        self.tailweight.unconstrained_variable = np.asarray(res.x[3])
        return res

    def kstest(self):
        """Kolmogorov-Smirnov test for normality of the transformed data (:83-85)."""
        from scipy import stats

        s, pvalue = stats.kstest(self.bijector(self.x), "norm")
        print("KS test statistic is %.3f, p-value is %.8f" % (s, pvalue))
        return s, pvalue
