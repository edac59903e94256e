"""NumPy restatement of the reference's one-dimensional normalising flow (test infrastructure only -- never
imported by the product).

oak/normalising_flow.py:46-56 chains tfb.Shift(-offset) -> Log -> Shift -> Scale -> SinhArcsinh (without the
first two when log=False); SinhArcsinh follows tensorflow_probability 0.11 (setup.py:33):
``sinh((arcsinh(x) + skewness) * tailweight)``.  KL_objective (:76-81) is
``mean(y^2 / 2) - mean(forward_log_det_jacobian(x))``.  Pinned (tests/golden/g8_normalising_flow.npz) against
the reference's own ``Normalizer`` executed over oracle/tf_shim: chain order, offset, standardiser
initialisation and KL_objective are the reference's code; the individual TFP 0.11 bijectors (Shift, Log, Scale,
SinhArcsinh and their log-det-Jacobians) are the shim's restatement of the published formulas, since
TensorFlow / TFP are not installable here -- that part is pinned only on identities in tests/ (analytic gradient
vs central differences, log-det-Jacobian vs a numerical derivative, forward/inverse round trip)."""
import numpy as np


def forward(x, offset, log, shift, scale, skewness, tailweight):
    x = np.asarray(x, dtype=np.float64)
    u = np.log(x - offset) if log else x
    z = (u + shift) * scale
    return np.sinh((np.arcsinh(z) + skewness) * tailweight)


def forward_log_det_jacobian(x, offset, log, shift, scale, skewness, tailweight):
    x = np.asarray(x, dtype=np.float64)
    u = np.log(x - offset) if log else x
    z = (u + shift) * scale
    w = (np.arcsinh(z) + skewness) * tailweight
    ldj = np.log(np.cosh(w)) + np.log(tailweight) - 0.5 * np.log1p(z * z) + np.log(scale)
    return ldj - u if log else ldj


def kl_objective_and_grad(x, offset, log, theta):
    """(J, dJ/d theta) for theta = (log scale, shift, skewness, log tailweight): the variables gpflow's Scipy
    optimiser sees (scale and tailweight carry tfb.Exp transforms, normalising_flow.py:16-27)."""
    x = np.asarray(x, dtype=np.float64)
    ta, b, eps, tt = (float(t) for t in theta)
    a, tau = np.exp(ta), np.exp(tt)
    u = np.log(x - offset) if log else x
    z = (u + b) * a
    w = (np.arcsinh(z) + eps) * tau
    y = np.sinh(w)
    n = x.shape[0]
    ldj = np.log(np.cosh(w)) + tt - 0.5 * np.log1p(z * z) + ta - (u if log else 0.0)
    J = 0.5 * np.mean(y * y) - np.mean(ldj)
    dJdw = (y * np.cosh(w) - np.tanh(w)) / n
    dJdz = dJdw * tau / np.sqrt(1.0 + z * z) + (z / (1.0 + z * z)) / n
    g = np.array([np.sum(dJdz * z) - 1.0, np.sum(dJdz) * a, np.sum(dJdw) * tau, np.sum(dJdw * w) - 1.0])
    return float(J), g
