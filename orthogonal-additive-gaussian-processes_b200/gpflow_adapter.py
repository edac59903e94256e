"""Guarded adapter for REAL gpflow (INTEGRATION.md section 3 as code).

When ``gpflow`` and ``tensorflow`` import, ``as_gpflow_kernel(native)`` wraps one of this package's kernels
(``OAKKernel``, ``OrthogonalRBFKernel``, ...) in a ``gpflow.kernels.Kernel`` subclass, so that gpflow's own models
(``gpflow.models.GPR / SGPR / SVGP``), ``print_summary`` and ``gpflow.optimizers.Scipy`` see an ordinary kernel:

* ``K`` / ``K_diag`` run the fused CUDA tiles (``oak_gram_f64`` / ``oak_gram_diag_f64``) through ``tf.numpy_function``;
* the gradient TensorFlow's autodiff would trace through ``oak/oak_kernel.py:223-278`` is supplied by
  ``tf.custom_gradient`` from the backward tiles (``oak_gram_backward_f64`` / ``oak_gram_backward_rows_f64`` /
  ``oak_gram_diag_backward_f64``): lengthscales, base variances and order variances, plus the FIRST argument's points
  (``inducing_variable.Z`` as gpflow passes it: ``Kuf = kernel(Z, X)``, ``Kuu = kernel(Z)``);
* the trainable parameters are mirrored as ``gpflow.Parameter`` objects with the reference's transforms
  (``tfp.bijectors.Sigmoid(low, high)`` for bounded lengthscales, ``gpflow.utilities.positive()`` otherwise).

Not provided (raise / documented): gradients of the categorical ``W`` / ``kappa`` (``training.py`` chains the table
cotangent for them on this package's own models) and the gradient with respect to the SECOND argument's points.

TensorFlow / gpflow are not installable in the build image (no wheels for Python 3.12, no network), so only the
forward half is exercised by the tests -- through ``oracle/tf_shim`` standing in for the two packages
(``tests/test_gpu_gpflow_adapter.py``); the ``tf.custom_gradient`` half is UNTESTED IN THIS IMAGE and follows the
gradient layout documented in ``include/oak_b200.h`` (``oak_backward_grad_count``).  Importing this module never
fails: ``available()`` reports whether the adapter can be used.
"""
from __future__ import annotations

import numpy as np

from . import _device
from ._gpflow_shim import Sigmoid as _NativeSigmoid


def _imports():
    import gpflow  # noqa: F401
    import tensorflow as tf  # noqa: F401

    return gpflow, tf


def available() -> bool:
    try:
        _imports()
        return True
    except Exception:
        return False


def _native_parameters(native):
    """[(kind, index, native Parameter)] in the order of the backward tiles' gradient vector
    (include/oak_b200.h: lengthscale_i at i, sigma2_n at D + n, base variance of sub-kernel i at count - D + i)."""
    subs = list(getattr(native, "kernels", [native]))
    out = []
    for i, k in enumerate(subs):
        base = getattr(k, "base_kernel", None)
        if base is not None:
            out.append(("lengthscale", i, base.lengthscales))
            out.append(("base_variance", i, base.variance))
    for n, v in enumerate(getattr(native, "variances", [])):
        out.append(("order_variance", n, v))
    return out


def as_gpflow_kernel(native):
    """``native`` wrapped as a ``gpflow.kernels.Kernel``; raises ImportError when gpflow / TensorFlow are absent."""
    gpflow, tf = _imports()

    def _transform(p):
        t = getattr(p, "transform", None)
        if isinstance(t, _NativeSigmoid):
            import tensorflow_probability as tfp

            f64 = lambda v: tf.cast(v, tf.float64)  # noqa: E731
            return tfp.bijectors.Sigmoid(f64(t.low), f64(t.high))
        return gpflow.utilities.positive()

    class OAKGpflowKernel(gpflow.kernels.Kernel):
        """gpflow-facing shell of a native kernel: parameters are gpflow's, arithmetic is the CUDA tiles'."""

        def __init__(self, inner):
            super().__init__()  # no slicing here: the native kernel applies its own active_dims
            self.native = inner
            for k in getattr(inner, "kernels", [inner]):
                for name in ("W", "kappa"):
                    q = getattr(k, name, None)
                    if q is not None and getattr(q, "trainable", False):
                        raise NotImplementedError(
                            f"gradient of the categorical {name} is not provided through the adapter: freeze it "
                            "(set_trainable) or train with oak_b200.training")
            self._slots = _native_parameters(inner)
            # gpflow Parameters, discovered by tf.Module like any other kernel's
            self.oak_parameters = [
                gpflow.Parameter(np.asarray(p.numpy(), dtype=np.float64), transform=_transform(p),
                                 trainable=bool(p.trainable)) for _, _, p in self._slots]

        # ---- host callbacks (NumPy in / out; the tiles run on the current CUDA device) -----------------
        def _push(self, theta):
            for (_, _, p), v in zip(self._slots, theta):
                p.assign(np.asarray(v, dtype=np.float64).reshape(np.shape(p.numpy())))

        def _np_K(self, X, X2, same, *theta):
            self._push(theta)
            return np.asarray(self.native.K(np.asarray(X), None if same else np.asarray(X2)), dtype=np.float64)

        def _np_K_diag(self, X, *theta):
            self._push(theta)
            return np.asarray(self.native.K_diag(np.asarray(X)), dtype=np.float64)

        def _split(self, grad, D):
            g = np.asarray(grad.cpu().numpy() if hasattr(grad, "cpu") else grad, dtype=np.float64)
            out = []
            for (kind, idx, p) in self._slots:
                v = {"lengthscale": g[idx], "order_variance": g[D + idx], "base_variance": g[len(g) - D + idx]}[kind]
                out.append(np.full(np.shape(p.numpy()), v, dtype=np.float64))
            return out

        def _np_K_grad(self, X, X2, same, dy, *theta):
            self._push(theta)
            nat = self.native
            spec = nat._make_spec()
            try:
                Xs = nat.slice(_device.to_device(np.asarray(X)), None)[0].contiguous()
                px = _device.Points(spec, Xs)
                px2 = None if same else _device.Points(
                    spec, nat.slice(_device.to_device(np.asarray(X2)), None)[0].contiguous())
                W = _device.to_device(np.ascontiguousarray(dy))
                if same:
                    W = W + W.T  # both arguments are the same points
                grad, rows = _device.gram_backward_rows(spec, px, W.contiguous() if same else W, px2)
                if same:
                    grad = grad * 0.5  # the parameter gradient must see W once
                D = spec.num_dims
                # rows[:, i] belongs to the column sub-kernel i reads from the SLICED input; scatter to X's columns
                outer = np.arange(np.shape(X)[1])[nat.active_dims]
                dX = np.zeros(np.shape(X), dtype=np.float64)
                rows_h = rows.cpu().numpy()
                for i, k in enumerate(getattr(nat, "kernels", [nat])):
                    inner = np.arange(len(outer))[k.active_dims]
                    dX[:, outer[int(np.atleast_1d(inner)[0])]] += rows_h[:, i]
                return [dX] + self._split(grad, D)
            finally:
                spec.close()

        def _np_K_diag_grad(self, X, dy, *theta):
            self._push(theta)
            spec = self.native._make_spec()
            try:
                px = _device.Points(spec, self.native.slice(_device.to_device(np.asarray(X)), None)[0].contiguous())
                grad = _device.gram_diag_backward(spec, px, 1.0, w=_device.to_device(np.asarray(dy), ndim=1))
                return self._split(grad, spec.num_dims)
            finally:
                spec.close()

        # ---- gpflow.kernels.Kernel ---------------------------------------------------------------------
        def K(self, X, X2=None):
            same = X2 is None
            theta = [tf.convert_to_tensor(p) for p in self.oak_parameters]
            X2t = X if same else X2

            @tf.custom_gradient
            def fused(Xa, Xb, *th):
                Kv = tf.numpy_function(lambda a, b, *t: self._np_K(a, b, same, *t), [Xa, Xb, *th], tf.float64)

                def grad(dy):
                    outs = tf.numpy_function(lambda a, b, g, *t: self._np_K_grad(a, b, same, g, *t),
                                             [Xa, Xb, dy, *th], [tf.float64] * (1 + len(th)))
                    dXa = outs[0]  # per sub-kernel column of the SLICED first argument
                    return (dXa, None, *outs[1:])

                return Kv, grad

            return fused(X, X2t, *theta)

        def K_diag(self, X):
            theta = [tf.convert_to_tensor(p) for p in self.oak_parameters]

            @tf.custom_gradient
            def fused(Xa, *th):
                Kv = tf.numpy_function(lambda a, *t: self._np_K_diag(a, *t), [Xa, *th], tf.float64)

                def grad(dy):
                    outs = tf.numpy_function(lambda a, g, *t: self._np_K_diag_grad(a, g, *t), [Xa, dy, *th],
                                             [tf.float64] * len(th))
                    return (None, *outs)

                return Kv, grad

            return fused(X, *theta)

    return OAKGpflowKernel(native)
