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


def _k_tilde_emp(x, y, l, loc, w):
    """Constrained RBF for an empirical measure sum_q w_q delta(s_q) (ortho_rbf_kernel.py:101-120,157-172)."""
    k = torch.exp(-0.5 * (x[:, None] - y[None, :]) ** 2 / l ** 2)
    cx = (torch.exp(-0.5 * (x[:, None] - loc[None, :]) ** 2 / l ** 2) * w[None, :]).sum(1)
    cy = (torch.exp(-0.5 * (y[:, None] - loc[None, :]) ** 2 / l ** 2) * w[None, :]).sum(1)
    v = (w[:, None] * torch.exp(-0.5 * (loc[:, None] - loc[None, :]) ** 2 / l ** 2) * w[None, :]).sum()
    return k - cx[:, None] * cy[None, :] / v


def _dim_values(X, X2, ls, measures):
    """Per-dimension constrained kernel matrices.  measures[d]: None -> Gaussian N(0, 1);
    ("gaussian", mu, var); ("empirical", loc, w); ("none",) -> plain RBF; ("table", B) -> discrete
    kernel B[x, x'] with a constant table (its parameters are not differentiated)."""
    D = X.shape[1]
    out = []
    for d in range(D):
        m = None if measures is None else measures[d]
        if m is None:
            out.append(_k_tilde(X[:, d], X2[:, d], ls[d], 0.0, 1.0))
        elif m[0] == "gaussian":
            out.append(_k_tilde(X[:, d], X2[:, d], ls[d], m[1], m[2]))
        elif m[0] == "empirical":
            out.append(_k_tilde_emp(X[:, d], X2[:, d], ls[d], torch.as_tensor(m[1], dtype=X.dtype).reshape(-1),
                                    torch.as_tensor(m[2], dtype=X.dtype).reshape(-1)))
        elif m[0] == "none":
            out.append(torch.exp(-0.5 * (X[:, d][:, None] - X2[:, d][None, :]) ** 2 / ls[d] ** 2))
        elif m[0] == "table":
            B = m[1] if isinstance(m[1], torch.Tensor) else torch.as_tensor(m[1], dtype=X.dtype)
            out.append(B[X[:, d].long()][:, X2[:, d].long()])
        else:
            raise ValueError(m[0])
    return out


def _esp_sum(ks, variances):
    P = len(variances) - 1
    s = [sum(k ** p for k in ks) for p in range(1, P + 1)]
    e = [torch.ones_like(ks[0])]
    for n in range(1, P + 1):
        acc = 0
        for q in range(1, n + 1):
            acc = acc + (-1) ** (q - 1) * e[n - q] * s[q - 1]
        e.append(acc / n)
    return sum(variances[n] * e[n] for n in range(P + 1))


def oak_K(X, X2, ls, variances, measures=None):
    """OAKKernel.K with Newton-Girard (oak_kernel.py:223-265)."""
    return _esp_sum(_dim_values(X, X2, ls, measures), variances)


def oak_K_diag(X, ls, variances, measures=None):
    """OAKKernel.K_diag (oak_kernel.py:267-278).  Test sizes only: each dimension's diagonal is read
    off its full matrix."""
    ks = [torch.diagonal(k) for k in _dim_values(X, X, ls, measures)]
    return _esp_sum(ks, variances)


def sgpr_elbo(X, Y, Z, ls, variances, noise, measures=None):
    """gpflow 2.2.1 SGPR.elbo, R = 1."""
    N = X.shape[0]
    Kuf = oak_K(Z, X, ls, variances, measures)
    Kuu = oak_K(Z, Z, ls, variances, measures) + JITTER * torch.eye(Z.shape[0], dtype=X.dtype)
    kd = oak_K_diag(X, ls, variances, measures)
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


def gpr_lml(X, Y, ls, variances, noise, measures=None):
    """gpflow GPR.log_marginal_likelihood."""
    N = X.shape[0]
    K = oak_K(X, X, ls, variances, measures) + noise * torch.eye(N, dtype=X.dtype)
    L = torch.linalg.cholesky(K)
    a = torch.linalg.solve_triangular(L, Y, upper=False)
    return -0.5 * (a ** 2).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * N * math.log(2 * math.pi)


def value_and_grad(fn, X, Y, Z, ls, variances, noise, measures=None):
    """Returns (value, d/d ls, d/d variances, d/d noise) as floats / numpy arrays."""
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    lsT = t(ls).clone().requires_grad_(True)
    vT = t(variances).clone().requires_grad_(True)
    nT = t(noise).clone().requires_grad_(True)
    if Z is None:
        val = fn(t(X), t(Y), lsT, vT, nT, measures)
    else:
        val = fn(t(X), t(Y), t(Z), lsT, vT, nT, measures)
    val.backward()
    return float(val.detach()), lsT.grad.numpy(), vT.grad.numpy(), float(nT.grad)
