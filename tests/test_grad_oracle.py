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


def test_svgp_oracle_identities():
    """The SVGP / Bernoulli restatement of gpflow 2.2.1 in oracle/oak_grad_oracle.py is unpinned against a gpflow
    run; these identities pin its pieces: the 20-point Gauss-Hermite expectations against adaptive quadrature of
    the integrals they stand for, KL = 0 and the prior predictive when q(u) is the prior, and the whitened
    conditional against the textbook un-whitened formulas (mean Kfu Kuu^-1 m, var Kff - Kfu Kuu^-1 (Kuu - S) Kuu^-1 Kuf
    with m = L q_mu, S = L diag(q_sqrt^2) L^T)."""
    import torch
    from scipy import integrate, stats

    from oracle import oak_grad_oracle as go

    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
    for link, inv in (("logit", go.inv_logit), ("probit", go.inv_probit)):
        for mu, var, y in ((0.7, 1.9, 1.0), (-1.2, 0.4, 0.0), (0.3, 1.0, 0.0)):  # smooth enough for 20 nodes
            p = lambda f: float(inv(t(f)))
            lik = lambda f: p(f) if y == 1.0 else 1.0 - p(f)
            dens = lambda f: stats.norm.pdf(f, mu, np.sqrt(var))
            ve = float(go.bernoulli_variational_expectations(t([mu]), t([var]), t([y]), inv))
            ld = float(go.bernoulli_predict_log_density(t([mu]), t([var]), t([y]), inv))
            assert abs(ve - integrate.quad(lambda f: dens(f) * np.log(lik(f)), mu - 12, mu + 12)[0]) < 1e-4  # the accuracy of 20 nodes on the jitter-clipped log
            assert abs(ld - np.log(integrate.quad(lambda f: dens(f) * lik(f), mu - 12, mu + 12)[0])) < 1e-4
    rng = np.random.default_rng(0)
    X, Z = rng.standard_normal((40, 3)), rng.standard_normal((7, 3))
    y = (rng.random(40) < 0.5).astype(float)
    ls, var = t([0.8, 1.3, 2.0]), t([0.4, 1.0, 0.6])
    # q(u) = prior
    fm, fv = go.svgp_conditional(t(X), t(Z), ls, var, t(np.zeros(7)), t(np.ones(7)))
    assert float(fm.abs().max()) == 0.0
    assert float((fv - go.oak_K_diag(t(X), ls, var)).abs().max()) < 1e-12
    elbo = go.svgp_elbo(t(X), t(y), t(Z), ls, var, t(np.zeros(7)), t(np.ones(7)))
    assert abs(float(elbo) - float(go.bernoulli_variational_expectations(fm, fv, t(y)).sum())) < 1e-12
    # whitened vs un-whitened parametrisation of the same q(u)
    q_mu, q_sqrt = t(rng.standard_normal(7)), t(rng.uniform(0.3, 1.2, 7))
    fm, fv = go.svgp_conditional(t(X), t(Z), ls, var, q_mu, q_sqrt)
    Kuu = go.oak_K(t(Z), t(Z), ls, var) + go.JITTER * torch.eye(7, dtype=torch.float64)
    Kuf = go.oak_K(t(Z), t(X), ls, var)
    L = torch.linalg.cholesky(Kuu)
    m, S = L @ q_mu, L @ torch.diag(q_sqrt ** 2) @ L.T
    Ki = torch.linalg.inv(Kuu)
    assert float((fm - Kuf.T @ Ki @ m).abs().max()) < 1e-8
    want = go.oak_K_diag(t(X), ls, var) - torch.einsum("ij,jk,ki->i", Kuf.T, Ki @ (Kuu - S) @ Ki, Kuf)
    assert float((fv - want).abs().max()) < 1e-7
    # alpha of the posterior object: Kfu alpha is the predictive mean
    alpha = go.svgp_alpha(t(Z), ls, var, q_mu)
    assert float((Kuf.T @ alpha[:, 0] - fm).abs().max()) < 1e-8


def test_inducing_point_gradient_formulas():
    """The row-point gradient the CUDA tiles implement, on the CPU: d k~(z, x)/dz = -ex (z - x)/l^2 - c^'(z) c^(x)
    chained through dK/dk~ by the removal recurrence gives d/dZ of sum W * K(Z, X); here the closed form is
    assembled in NumPy for the Gaussian measure and compared with autograd, and autograd's d ELBO / dZ is
    compared with central differences (the reference differentiates Z the same way when zfixed=False)."""
    X, y, Z, ls, var, noise, _ = _setup(n=60, D=3, P=2, M=9, seed=3)
    Z = Z + 0.1 * np.random.default_rng(1).standard_normal(Z.shape)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    rng = np.random.default_rng(2)
    W = rng.standard_normal((Z.shape[0], X.shape[0]))
    ZT = t(Z).clone().requires_grad_(True)
    (t(W) * go.oak_K(ZT, t(X), t(ls), t(var))).sum().backward()
    # closed form, Gaussian measure N(0, 1): c^(x) = pre exp(-x^2 / (2 (l^2 + 1))) / sqrt(v)
    D = Z.shape[1]
    kt, ex, chz, chx = [], [], [], []
    for d in range(D):
        l = ls[d]
        pre = l / np.sqrt(l ** 2 + 1.0)
        v = l / np.sqrt(l ** 2 + 2.0)
        cz = pre * np.exp(-0.5 * Z[:, d] ** 2 / (l ** 2 + 1.0)) / np.sqrt(v)
        cx = pre * np.exp(-0.5 * X[:, d] ** 2 / (l ** 2 + 1.0)) / np.sqrt(v)
        e = np.exp(-0.5 * (Z[:, d][:, None] - X[:, d][None, :]) ** 2 / l ** 2)
        kt.append(e - cz[:, None] * cx[None, :])
        ex.append(e)
        chz.append(cz)
        chx.append(cx)
    # elementary symmetric polynomials with dimension d removed: e_0^{(-d)} = 1, e_1^{(-d)} = e_1 - k_d  (P = 2)
    e1 = sum(kt)
    gZ = np.zeros_like(Z)
    for d in range(D):
        dKdk = var[1] + var[2] * (e1 - kt[d])
        wk = W * dKdk
        l = ls[d]
        A = (wk * ex[d] * (Z[:, d][:, None] - X[:, d][None, :])).sum(1)
        B = (wk * chx[d][None, :]).sum(1)
        dchz = -chz[d] * Z[:, d] / (l ** 2 + 1.0)
        gZ[:, d] = -A / l ** 2 - dchz * B
    assert np.allclose(gZ, ZT.grad.numpy(), rtol=1e-10, atol=1e-12)
    # autograd of the ELBO with respect to Z against central differences
    _, _, _, _, a_Z = go.value_and_grad(go.sgpr_elbo, X, y, Z, ls, var, noise, wrt_Z=True)
    h = 1e-6
    for (i, d) in ((0, 0), (4, 1), (8, 2)):
        Zp, Zm = Z.copy(), Z.copy()
        Zp[i, d] += h
        Zm[i, d] -= h
        fd = (go.value_and_grad(go.sgpr_elbo, X, y, Zp, ls, var, noise)[0]
              - go.value_and_grad(go.sgpr_elbo, X, y, Zm, ls, var, noise)[0]) / (2 * h)
        assert abs(fd - a_Z[i, d]) < 1e-5 * max(1.0, abs(fd))
