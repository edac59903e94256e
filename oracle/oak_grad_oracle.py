"""Gradient oracle (test infrastructure only -- never imported by the product).

The reference trains OAK by automatic differentiation of gpflow's objectives through TensorFlow
(oak/model_utils.py:168-175: ``gpflow.optimizers.Scipy().minimize(model.training_loss_closure(),
model.trainable_variables)``).  This module restates the forward computations of
oracle/oak_oracle.py in torch float64 (CPU) for Gaussian-measure RBF dimensions and lets autograd
produce d objective / d (lengthscales, order variances, likelihood variance) -- the quantities the
CUDA backward tiles compute analytically.  parity unpinned by the reference's own tests (it has
no gradient fixtures); the forward values are checked against oak_oracle.py in tests/.
"""
import math

import torch

JITTER = 1e-6


def _k_tilde(x, y, l, mu, var):
    """Constrained RBF for a Gaussian measure N(mu, var), s^2 = 1 (ortho_rbf_kernel.py:82-97,157-172)."""
    d2 = (x[:, None] - y[None, :]) ** 2
    k = torch.exp(-0.5 * d2 / l ** 2)
    cx = l / torch.sqrt(l ** 2 + var) * torch.exp(-0.5 * (x - mu) ** 2 / (l ** 2 + var))
    cy = l / torch.sqrt(l ** 2 + var) * torch.exp(-0.5 * (y - mu) ** 2 / (l ** 2 + var))
    v = l / torch.sqrt(l ** 2 + 2 * var)
    return k - cx[:, None] * cy[None, :] / v


def oak_K(X, X2, ls, variances, mu=0.0, var=1.0):
    """OAKKernel.K with Newton-Girard (oak_kernel.py:223-265)."""
    D = X.shape[1]
    P = len(variances) - 1
    ks = [_k_tilde(X[:, d], X2[:, d], ls[d], mu, var) for d in range(D)]
    s = [sum(k ** p for k in ks) for p in range(1, P + 1)]
    e = [torch.ones_like(ks[0])]
    for n in range(1, P + 1):
        acc = 0
        for q in range(1, n + 1):
            acc = acc + (-1) ** (q - 1) * e[n - q] * s[q - 1]
        e.append(acc / n)
    return sum(variances[n] * e[n] for n in range(P + 1))


def oak_K_diag(X, ls, variances, mu=0.0, var=1.0):
    D = X.shape[1]
    P = len(variances) - 1
    ks = []
    for d in range(D):
        l = ls[d]
        cx = l / torch.sqrt(l ** 2 + var) * torch.exp(-0.5 * (X[:, d] - mu) ** 2 / (l ** 2 + var))
        ks.append(1.0 - cx ** 2 / (l / torch.sqrt(l ** 2 + 2 * var)))
    s = [sum(k ** p for k in ks) for p in range(1, P + 1)]
    e = [torch.ones_like(ks[0])]
    for n in range(1, P + 1):
        acc = 0
        for q in range(1, n + 1):
            acc = acc + (-1) ** (q - 1) * e[n - q] * s[q - 1]
        e.append(acc / n)
    return sum(variances[n] * e[n] for n in range(P + 1))


def sgpr_elbo(X, Y, Z, ls, variances, noise):
    """gpflow 2.2.1 SGPR.elbo, R = 1."""
    N = X.shape[0]
    Kuf = oak_K(Z, X, ls, variances)
    Kuu = oak_K(Z, Z, ls, variances) + JITTER * torch.eye(Z.shape[0], dtype=X.dtype)
    kd = oak_K_diag(X, ls, variances)
    L = torch.linalg.cholesky(Kuu)
    sigma = torch.sqrt(noise)
    A = torch.linalg.solve_triangular(L, Kuf, upper=False) / sigma
    AAT = A @ A.T
    B = AAT + torch.eye(Z.shape[0], dtype=X.dtype)
    LB = torch.linalg.cholesky(B)
    Aerr = A @ Y
    c = torch.linalg.solve_triangular(LB, Aerr, upper=False) / sigma
    bound = -0.5 * N * math.log(2 * math.pi)
    bound = bound - torch.log(torch.diagonal(LB)).sum()
    bound = bound - 0.5 * N * torch.log(noise)
    bound = bound - 0.5 * (Y ** 2).sum() / noise
    bound = bound + 0.5 * (c ** 2).sum()
    bound = bound - 0.5 * kd.sum() / noise
    bound = bound + 0.5 * torch.diagonal(AAT).sum()
    return bound


def gpr_lml(X, Y, ls, variances, noise):
    """gpflow GPR.log_marginal_likelihood."""
    N = X.shape[0]
    K = oak_K(X, X, ls, variances) + noise * torch.eye(N, dtype=X.dtype)
    L = torch.linalg.cholesky(K)
    a = torch.linalg.solve_triangular(L, Y, upper=False)
    return -0.5 * (a ** 2).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * N * math.log(2 * math.pi)


def value_and_grad(fn, X, Y, Z, ls, variances, noise):
    """Returns (value, d/d ls, d/d variances, d/d noise) as floats / numpy arrays."""
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    lsT = t(ls).clone().requires_grad_(True)
    vT = t(variances).clone().requires_grad_(True)
    nT = t(noise).clone().requires_grad_(True)
    if Z is None:
        val = fn(t(X), t(Y), lsT, vT, nT)
    else:
        val = fn(t(X), t(Y), t(Z), lsT, vT, nT)
    val.backward()
    return float(val.detach()), lsT.grad.numpy(), vT.grad.numpy(), float(nT.grad)
