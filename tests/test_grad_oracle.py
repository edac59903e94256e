"""CPU checks of the gradient oracle (oracle/oak_grad_oracle.py): its forward values are the NumPy
oracle's, and the statistics-space gradient formulas documented in the product's training.py agree
with autograd through the gpflow operation order."""
import numpy as np
import torch

from helpers import build_oracle
from oracle import oak_grad_oracle as go
from oracle import oak_oracle as oo


def _setup(n=160, D=4, P=3, M=24, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D))
    y = (np.sin(X[:, 0]) + X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
    ls = rng.uniform(0.5, 2.5, D)
    var = rng.uniform(0.3, 1.2, P + 1)
    cfg = dict(dims=[{"type": "rbf", "lengthscale": float(l), "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)}
                     for l in ls], depth=P, variances=list(var), share_var=True)
    return X, y, X[:M].copy(), ls, var, 0.05, build_oracle(cfg)


def test_forward_values_match_the_numpy_oracle():
    X, y, Z, ls, var, noise, ref = _setup()
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    assert np.allclose(go.oak_K(t(Z), t(X), t(ls), t(var)).numpy(), ref.K(Z, X), rtol=1e-12, atol=1e-13)
    assert np.allclose(go.oak_K_diag(t(X), t(ls), t(var)).numpy(), ref.K_diag(X), rtol=1e-12, atol=1e-13)
    v, *_ = go.value_and_grad(go.sgpr_elbo, X, y, Z, ls, var, noise)
    assert abs(v - oo.sgpr_elbo(ref, X, y, Z, noise)) < 1e-10 * abs(v)
    v, *_ = go.value_and_grad(go.gpr_lml, X, y, None, ls, var, noise)
    assert abs(v - oo.gpr_log_marginal_likelihood(ref, X, y, noise)) < 1e-10 * abs(v)


def test_statistics_space_gradients_match_autograd():
    """d ELBO / d (Phi, b, s, Q, noise) as written in training.py, chained through dK/d theta by autograd."""
    X, y, Z, ls, var, noise, _ = _setup(seed=1)
    n, M = X.shape[0], Z.shape[0]
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    _, a_ls, a_var, a_noise = go.value_and_grad(go.sgpr_elbo, X, y, Z, ls, var, noise)
    Kuf = go.oak_K(t(Z), t(X), t(ls), t(var)).numpy()
    Kuu = go.oak_K(t(Z), t(Z), t(ls), t(var)).numpy()
    kd = go.oak_K_diag(t(X), t(ls), t(var)).numpy()
    Phi, b, s = Kuf @ Kuf.T, Kuf @ y, kd.sum()
    Q = Kuu + go.JITTER * np.eye(M)
    S = Q + Phi / noise
    Si, Qi = np.linalg.inv(S), np.linalg.inv(Q)
    Sib = Si @ b
    G_phi = -0.5 * Si / noise - (Sib @ Sib.T) / (2 * noise ** 3) + Qi / (2 * noise)
    g_b = Sib / noise ** 2
    G_Q = -0.5 * Si + 0.5 * Qi - (Sib @ Sib.T) / (2 * noise ** 2) - Qi @ Phi @ Qi / (2 * noise)
    g_noise = (-n / (2 * noise) + (y.T @ y).item() / (2 * noise ** 2) - (b.T @ Sib).item() / noise ** 3
               + s / (2 * noise ** 2) - np.trace(Qi @ Phi) / (2 * noise ** 2) + 0.5 * np.trace(Si @ Phi) / noise ** 2
               + (Sib.T @ Phi @ Sib).item() / (2 * noise ** 4))
    assert abs(g_noise - a_noise) < 1e-8 * abs(a_noise)
    lsT, vT = t(ls).clone().requires_grad_(True), t(var).clone().requires_grad_(True)
    W = 2 * G_phi @ Kuf + g_b @ y.T
    tot = ((t(W) * go.oak_K(t(Z), t(X), lsT, vT)).sum() + (t(G_Q) * go.oak_K(t(Z), t(Z), lsT, vT)).sum()
           - 0.5 / noise * go.oak_K_diag(t(X), lsT, vT).sum())
    tot.backward()
    assert np.allclose(lsT.grad.numpy(), a_ls, rtol=1e-8)
    assert np.allclose(vT.grad.numpy(), a_var, rtol=1e-8)
