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


def _k_tilde_general(x, y, l, s2, cx, cy, v):
    return s2 * torch.exp(-0.5 * (x[:, None] - y[None, :]) ** 2 / l ** 2) - cx[:, None] * cy[None, :] / v


def _uniform_parts(x, l, s2, a, b):
    """cov_X_s / var_s for a uniform measure on [a, b] (ortho_rbf_kernel.py:49-78)."""
    c = s2 * l * math.sqrt(math.pi / 2) / (b - a) * (torch.erf((b - x) / (math.sqrt(2) * l)) - torch.erf((a - x) / (math.sqrt(2) * l)))
    y = (b - a) / math.sqrt(2) / l
    v = 2 / (b - a) ** 2 * s2 * l ** 2 * (math.sqrt(math.pi) * y * torch.erf(y) + torch.exp(-y ** 2) - 1)
    return c, v


def _mog_parts(x, l, s2, means, variances, weights):
    """cov_X_s / var_s for a mixture of Gaussians (ortho_rbf_kernel.py:124-152)."""
    S = l ** 2 + variances
    c = s2 * l * (weights[None, :] * torch.exp(-0.5 * (x[:, None] - means[None, :]) ** 2 / S[None, :]) / torch.sqrt(S)[None, :]).sum(1)
    T = l ** 2 + variances[:, None] + variances[None, :]
    v = s2 * l * (weights[:, None] * weights[None, :] * torch.exp(-0.5 * (means[:, None] - means[None, :]) ** 2 / T) / torch.sqrt(T)).sum()
    return c, v


def _dim_values(X, X2, ls, measures, s2=None):
    """Per-dimension constrained kernel matrices.  measures[d]: None -> Gaussian N(0, 1);
    ("gaussian", mu, var); ("uniform", a, b); ("empirical", loc, w); ("mog", means, variances, weights);
    ("none",) -> plain RBF; ("table", B) -> discrete kernel B[x, x'] (B may be a differentiable tensor).
    s2[d]: base variance of the RBF dims (default 1)."""
    D = X.shape[1]
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(a, dtype=X.dtype)
    out = []
    for d in range(D):
        m = None if measures is None else measures[d]
        x, y = X[:, d], X2[:, d]
        if m is not None and m[0] == "table":
            B = t(m[1])
            out.append(B[x.long()][:, y.long()])
            continue
        l = ls[d]
        v2 = 1.0 if s2 is None else s2[d]
        if m is None or m[0] == "gaussian":
            mu, var = (0.0, 1.0) if m is None else (m[1], m[2])
            pre = v2 * l / torch.sqrt(l ** 2 + var)
            cx = pre * torch.exp(-0.5 * (x - mu) ** 2 / (l ** 2 + var))
            cy = pre * torch.exp(-0.5 * (y - mu) ** 2 / (l ** 2 + var))
            v = v2 * l / torch.sqrt(l ** 2 + 2 * var)
        elif m[0] == "uniform":
            cx, v = _uniform_parts(x, l, v2, m[1], m[2])
            cy, _ = _uniform_parts(y, l, v2, m[1], m[2])
        elif m[0] == "empirical":
            loc, w = t(m[1]).reshape(-1), t(m[2]).reshape(-1)
            cx = v2 * (torch.exp(-0.5 * (x[:, None] - loc[None, :]) ** 2 / l ** 2) * w[None, :]).sum(1)
            cy = v2 * (torch.exp(-0.5 * (y[:, None] - loc[None, :]) ** 2 / l ** 2) * w[None, :]).sum(1)
            v = v2 * (w[:, None] * torch.exp(-0.5 * (loc[:, None] - loc[None, :]) ** 2 / l ** 2) * w[None, :]).sum()
        elif m[0] == "mog":
            cx, v = _mog_parts(x, l, v2, t(m[1]), t(m[2]), t(m[3]))
            cy, _ = _mog_parts(y, l, v2, t(m[1]), t(m[2]), t(m[3]))
        elif m[0] == "none":
            out.append(v2 * torch.exp(-0.5 * (x[:, None] - y[None, :]) ** 2 / l ** 2))
            continue
        else:
            raise ValueError(m[0])
        out.append(_k_tilde_general(x, y, l, v2, cx, cy, v))
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


# ---- whitened SVGP, diagonal q(u), Bernoulli likelihood (gpflow 2.2.1; not vendored under /root/reference) ----
# Call sites: examples/uci/uci_classification_train.py:108-135.  Restated from the published gpflow 2.2.1 sources:
# models/svgp.py (elbo, prior_kl), kullback_leiblers.py (gauss_kl, K=None, diagonal q_sqrt),
# conditionals/util.py (base_conditional, white=True), likelihoods (Bernoulli, NDiagGHQuadrature with 20 points,
# nodes from np.polynomial.hermite.hermgauss scaled by sqrt 2 and 1 / sqrt pi), posteriors.py (alpha = L^-T q_mu).
# Parity unpinned against a gpflow run (gpflow / TensorFlow are not installed in this image); pinned on the
# mathematical identities in tests/ (q(u) = prior gives KL = 0 and the prior predictive; quadrature against
# scipy.integrate.quad).
def inv_logit(f, jitter=1e-3):
    return torch.sigmoid(f) * (1 - 2 * jitter) + jitter


def inv_probit(f, jitter=1e-3):
    return 0.5 * (1.0 + torch.erf(f / math.sqrt(2.0))) * (1 - 2 * jitter) + jitter


def gh_points_and_weights(n_gh=20):
    import numpy as np

    x, w = np.polynomial.hermite.hermgauss(n_gh)
    return torch.as_tensor(x * np.sqrt(2.0)), torch.as_tensor(w / np.sqrt(np.pi))


def svgp_conditional(Xnew, Z, ls, variances, q_mu, q_sqrt, measures=None):
    """base_conditional(Kmn, Kmm + jitter, Knn_diag, f=q_mu, q_sqrt=diag, white=True) -> (mean [N], var [N])."""
    Kmm = oak_K(Z, Z, ls, variances, measures) + JITTER * torch.eye(Z.shape[0], dtype=Z.dtype)
    Kmn = oak_K(Z, Xnew, ls, variances, measures)
    Knn = oak_K_diag(Xnew, ls, variances, measures)
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
    fvar = Knn - (A ** 2).sum(0)
    fmean = (A.T @ q_mu.reshape(-1, 1))[:, 0]
    LTA = A * q_sqrt.reshape(-1, 1)
    fvar = fvar + (LTA ** 2).sum(0)
    return fmean, fvar


def bernoulli_log_prob(F, Y, invlink=inv_logit):
    p = invlink(F)
    return torch.log(torch.where(Y == 1, p, 1 - p))


def bernoulli_variational_expectations(Fmu, Fvar, Y, invlink=inv_logit, n_gh=20):
    x, w = gh_points_and_weights(n_gh)
    F = Fmu[:, None] + torch.sqrt(Fvar)[:, None] * x[None, :]
    return (bernoulli_log_prob(F, Y[:, None], invlink) * w[None, :]).sum(1)


def bernoulli_predict_log_density(Fmu, Fvar, Y, invlink=inv_logit, n_gh=20):
    x, w = gh_points_and_weights(n_gh)
    F = Fmu[:, None] + torch.sqrt(Fvar)[:, None] * x[None, :]
    return torch.logsumexp(bernoulli_log_prob(F, Y[:, None], invlink) + torch.log(w)[None, :], dim=1)


def svgp_elbo(X, Y, Z, ls, variances, q_mu, q_sqrt, measures=None, invlink=inv_logit, num_data=None):
    fmean, fvar = svgp_conditional(X, Z, ls, variances, q_mu, q_sqrt, measures)
    var_exp = bernoulli_variational_expectations(fmean, fvar, Y.reshape(-1), invlink)
    M = Z.shape[0]
    two_kl = (q_mu ** 2).sum() - M - torch.log(q_sqrt ** 2).sum() + (q_sqrt ** 2).sum()
    scale = 1.0 if num_data is None else num_data / X.shape[0]
    return var_exp.sum() * scale - 0.5 * two_kl


def svgp_alpha(Z, ls, variances, q_mu, measures=None):
    Kmm = oak_K(Z, Z, ls, variances, measures) + JITTER * torch.eye(Z.shape[0], dtype=Z.dtype)
    return torch.linalg.solve_triangular(torch.linalg.cholesky(Kmm).T, q_mu.reshape(-1, 1), upper=True)


def value_and_grad(fn, X, Y, Z, ls, variances, noise, measures=None, wrt_Z=False):
    """Returns (value, d/d ls, d/d variances, d/d noise) as floats / numpy arrays; with ``wrt_Z`` also
    d/d Z (the inducing points, gpflow's ``inducing_variable.Z`` when zfixed=False)."""
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    lsT = t(ls).clone().requires_grad_(True)
    vT = t(variances).clone().requires_grad_(True)
    nT = t(noise).clone().requires_grad_(True)
    if Z is None:
        val = fn(t(X), t(Y), lsT, vT, nT, measures)
    else:
        ZT = t(Z).clone().requires_grad_(bool(wrt_Z))
        val = fn(t(X), t(Y), ZT, lsT, vT, nT, measures)
    val.backward()
    out = (float(val.detach()), lsT.grad.numpy(), vT.grad.numpy(), float(nT.grad))
    return out + (ZT.grad.numpy(),) if wrt_Z else out
