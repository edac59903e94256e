"""Whitened SVGP with diagonal q(u) and a Bernoulli likelihood around the Kuf tiles (SURVEY.md section 8(f) #4;
the model of examples/uci/uci_classification_train.py:108-135) against the torch-CPU restatement of gpflow
2.2.1 in oracle/oak_grad_oracle.py: values to 1e-9, gradients against autograd."""
import numpy as np
import pytest

from helpers import build_oracle, max_rel_err
from oracle import oak_grad_oracle as go

pytestmark = pytest.mark.gpu


def _t(a):
    import torch

    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def _cfg(n, D, P, m, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D))
    logit = 2.0 * np.sin(X[:, 0]) + X[:, 1] * X[:, min(2, D - 1)]
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-logit))).astype(np.float64).reshape(-1, 1)
    ls = rng.uniform(0.6, 2.0, D)
    var = rng.uniform(0.3, 1.2, P + 1)
    dims = [{"type": "rbf", "lengthscale": float(l), "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)} for l in ls]
    cfg = dict(X=X, y=y, Z=X[:m] + 0.1 * rng.standard_normal((m, D)), dims=dims, depth=P, variances=list(var),
               share_var=True)
    q_mu = 0.5 * rng.standard_normal((m, 1))
    q_sqrt = rng.uniform(0.3, 1.4, (m, 1))
    return cfg, ls, var, q_mu, q_sqrt


def _model(cfg, q_mu, q_sqrt, link="logit", **kw):
    from oak_b200._gpflow_shim import Bernoulli, inv_logit, inv_probit
    from oak_b200.models import SVGP
    from oak_b200.workloads import build_kernel

    lk = Bernoulli(invlink=inv_logit if link == "logit" else inv_probit)
    return SVGP(kernel=build_kernel(cfg), likelihood=lk, inducing_variable=cfg["Z"], whiten=True, q_diag=True,
                q_mu=q_mu, q_sqrt=q_sqrt, **kw)


@pytest.mark.parametrize("link", ["logit", "probit"])
def test_bernoulli_quadrature_kernel_matches_oracle_and_autograd(link):
    import torch

    from oak_b200 import _device

    rng = np.random.default_rng(0)
    n = 1000
    mean = np.concatenate([rng.standard_normal(n - 6) * 3.0, [0.0, 40.0, -40.0, 8.0, -8.0, 1e-3]])
    var = np.concatenate([rng.uniform(1e-3, 9.0, n - 6), [1e-12, 4.0, 4.0, 1e-6, 25.0, 1.0]])
    y = (rng.random(n) < 0.5).astype(np.float64)
    inv = go.inv_logit if link == "logit" else go.inv_probit
    out = _device.bernoulli_quadrature(_device.to_device(mean, ndim=1), _device.to_device(var, ndim=1),
                                       _device.to_device(y, ndim=1), 0 if link == "logit" else 1, 1e-3, 20,
                                       want=("varexp", "gmean", "gvar", "logdensity"))
    mT, vT = _t(mean).requires_grad_(True), _t(var).requires_grad_(True)
    ve = go.bernoulli_variational_expectations(mT, vT, _t(y), inv)
    ve.sum().backward()
    assert max_rel_err(out["varexp"].cpu().numpy(), ve.detach().numpy()) < 1e-12
    assert max_rel_err(out["gmean"].cpu().numpy()[:-6], mT.grad.numpy()[:-6]) < 1e-11
    assert max_rel_err(out["gvar"].cpu().numpy()[:-6], vT.grad.numpy()[:-6]) < 1e-9
    ld = go.bernoulli_predict_log_density(_t(mean), _t(var), _t(y), inv)
    assert max_rel_err(out["logdensity"].cpu().numpy(), ld.numpy()) < 1e-12
    assert torch.isfinite(out["gmean"]).all() and torch.isfinite(out["gvar"]).all()


def test_moments_kernels_match_torch():
    import torch

    from oak_b200 import _device

    g = torch.Generator(device="cuda").manual_seed(1)
    m, n = 37, 1111
    A = torch.randn(m, n + 5, dtype=torch.float64, device="cuda", generator=g)[:, :n]  # pitched view
    q_mu = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    q_sqrt = torch.rand(m, dtype=torch.float64, device="cuda", generator=g) + 0.2
    kd = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 50.0
    mean, var = _device.svgp_moments(A, q_mu, q_sqrt, kd)
    assert max_rel_err(mean.cpu().numpy(), (A.T @ q_mu).cpu().numpy()) < 1e-13
    want = kd - (A * A).sum(0) + ((A * q_sqrt[:, None]) ** 2).sum(0)
    assert max_rel_err(var.cpu().numpy(), want.cpu().numpy()) < 1e-13
    gm = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    gv = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    acc = torch.ones(m, dtype=torch.float64, device="cuda")
    Abar = _device.svgp_moments_backward(A, q_mu, q_sqrt, gm, gv, acc)
    Ar, mr, sr = (x.detach().clone().requires_grad_(True) for x in (A.contiguous(), q_mu, q_sqrt))
    obj = (gm * (Ar.T @ mr)).sum() + (gv * (-(Ar * Ar).sum(0) + ((Ar * sr[:, None]) ** 2).sum(0))).sum()
    obj.backward()
    assert max_rel_err(Abar.cpu().numpy(), Ar.grad.cpu().numpy()) < 1e-13
    assert max_rel_err((acc - 1.0).cpu().numpy(), sr.grad.cpu().numpy()) < 1e-12


@pytest.mark.parametrize("link,chunk,num_data", [("logit", 65536, None), ("logit", 64, None), ("probit", 200, 5000)])
def test_svgp_elbo_and_every_gradient_match_autograd(link, chunk, num_data):
    from oak_b200.training import svgp_elbo_and_grad

    cfg, ls, var, q_mu, q_sqrt = _cfg(333, 5, 3, 29, seed=11)
    m = _model(cfg, q_mu, q_sqrt, link, chunk=chunk, num_data=num_data)
    data = (cfg["X"], cfg["y"])
    elbo, g_ls, g_var, g_noise = svgp_elbo_and_grad(m, data)
    assert g_noise is None
    T = {k: _t(v).clone().requires_grad_(True) for k, v in dict(ls=ls, var=var, q_mu=q_mu, q_sqrt=q_sqrt, Z=cfg["Z"]).items()}
    inv = go.inv_logit if link == "logit" else go.inv_probit
    val = go.svgp_elbo(_t(cfg["X"]), _t(cfg["y"]), T["Z"], T["ls"], T["var"], T["q_mu"], T["q_sqrt"], invlink=inv,
                       num_data=num_data)
    val.backward()
    assert abs(elbo - float(val)) < 1e-9 * abs(float(val))
    assert abs(m.elbo(data) - float(val)) < 1e-9 * abs(float(val))
    assert max_rel_err(g_ls, T["ls"].grad.numpy()) < 1e-7
    assert max_rel_err(g_var, T["var"].grad.numpy()) < 1e-7
    assert max_rel_err(m._variational_grads[id(m.q_mu)], T["q_mu"].grad.numpy()) < 1e-8
    assert max_rel_err(m._variational_grads[id(m.q_sqrt)], T["q_sqrt"].grad.numpy()) < 1e-8
    assert max_rel_err(m._inducing_grad, T["Z"].grad.numpy()) < 1e-7


def test_svgp_prior_q_gives_zero_kl_and_the_prior_predictive():
    """q(u) = p(u) (q_mu = 0, q_sqrt = 1, gpflow's initial state): KL = 0, predict_f = (0, K_diag)."""
    cfg, ls, var, _, _ = _cfg(150, 4, 2, 20, seed=3)
    m = _model(cfg, None, None)
    assert np.all(m.q_mu.numpy() == 0.0) and np.allclose(m.q_sqrt.numpy(), 1.0)
    mean, v = m.predict_f(cfg["X"])
    kd = build_oracle(cfg).K_diag(cfg["X"])
    assert np.abs(mean).max() == 0.0
    assert max_rel_err(v[:, 0], kd) < 1e-12
    ve = go.bernoulli_variational_expectations(_t(np.zeros(150)), _t(kd), _t(cfg["y"][:, 0]))
    assert abs(m.elbo((cfg["X"], cfg["y"])) - float(ve.sum())) < 1e-9 * abs(float(ve.sum()))


def test_svgp_predictions_and_sufficient_statistics_match_oracle():
    from oak_b200.utils import get_model_sufficient_statistics

    cfg, ls, var, q_mu, q_sqrt = _cfg(260, 5, 3, 31, seed=5)
    q_sqrt = np.minimum(q_sqrt, 0.9)  # I - diag(q_sqrt^2) positive definite: the reference's chol(inv(Qinv)) exists
    m = _model(cfg, q_mu, q_sqrt, chunk=100)
    rng = np.random.default_rng(9)
    Xn = rng.standard_normal((123, 5))
    yn = (rng.random((123, 1)) < 0.5).astype(np.float64)
    fm, fv = go.svgp_conditional(_t(Xn), _t(cfg["Z"]), _t(ls), _t(var), _t(q_mu), _t(q_sqrt))
    mean, v = m.predict_f(Xn)
    assert mean.shape == (123, 1) and v.shape == (123, 1)
    assert max_rel_err(mean[:, 0], fm.numpy()) < 1e-9
    assert max_rel_err(v[:, 0], fv.numpy()) < 1e-9
    ld = go.bernoulli_predict_log_density(fm, fv, _t(yn[:, 0]))
    assert max_rel_err(m.predict_log_density((Xn, yn)), ld.numpy()) < 1e-9
    alpha, L = get_model_sufficient_statistics(m, get_L=True)
    a_ref = go.svgp_alpha(_t(cfg["Z"]), _t(ls), _t(var), _t(q_mu)).numpy()
    assert max_rel_err(alpha, a_ref) < 1e-9
    # predictive variance = K_diag - Kfu Qinv Kuf with Qinv = (L L^T)^-1
    Kfu = build_oracle(cfg).K(Xn, cfg["Z"])
    kd = build_oracle(cfg).K_diag(Xn)
    Qinv = np.linalg.inv(L @ L.T)
    assert max_rel_err(kd - np.einsum("ij,jk,ik->i", Kfu, Qinv, Kfu), fv.numpy()) < 1e-6
    assert max_rel_err((Kfu @ alpha)[:, 0], fm.numpy()) < 1e-8


def test_svgp_classification_trains_with_bfgs_and_decomposes():
    """The flow of uci_classification_train.py:95-160 on synthetic labels: SVGP on an OAK kernel, Z fixed, BFGS
    on the closure; then accuracy, NLL, Sobol indices and the additive decomposition of the latent mean."""
    from oak_b200._gpflow_shim import Bernoulli, inv_logit, set_trainable
    from oak_b200.model_utils import create_model_oak
    from oak_b200.models import SVGP
    from oak_b200.training import optimise, trainable_parameters
    from oak_b200.utils import compute_sobol_oak, get_model_sufficient_statistics, get_prediction_component

    rng = np.random.default_rng(4)
    X = rng.standard_normal((600, 3))
    logit = 3.0 * np.sin(2.0 * X[:, 0]) + 2.0 * X[:, 1]
    y = (rng.random(600) < 1.0 / (1.0 + np.exp(-logit))).astype(np.float64).reshape(-1, 1)
    Xtr, ytr, Xte, yte = X[:450], y[:450], X[450:], y[450:]
    base = create_model_oak((Xtr, ytr), max_interaction_depth=2, optimise=False)
    m = SVGP(kernel=base.kernel, likelihood=Bernoulli(invlink=inv_logit), inducing_variable=Xtr[:40].copy(),
             whiten=True, q_diag=True)
    set_trainable(m.inducing_variable, False)
    data = (Xtr, ytr)
    assert not any(p is m.inducing_variable.Z for p in trainable_parameters(m))
    loss0 = m.training_loss(data)
    res = optimise(m.training_loss_closure(data), method="BFGS", maxiter=40)
    assert res.fun < loss0 - 20.0
    assert abs(m.training_loss(data) - res.fun) < 1e-8 * abs(res.fun)
    mu, var = m.predict_f(Xte)
    prob = inv_logit(mu)
    err = np.mean(np.abs((prob > 0.5).astype(int)[:, 0] - yte[:, 0]))
    assert err < 0.3
    nll = -np.mean(m.predict_log_density((Xte, yte)))
    assert nll < 0.65
    # Sobol indices and the component decomposition (utils.py:338-435, 491-530 with an SVGP)
    m.data = data
    idx, sob = compute_sobol_oak(m, delta=1.0, mu=0.0)
    sob = np.array(sob) / np.sum(sob)
    assert idx[:3] == [[0], [1], [2]]
    assert sob[0] + sob[1] > 0.5 and sob[2] < min(sob[0], sob[1])
    alpha = get_model_sufficient_statistics(m, get_L=False)
    comps = get_prediction_component(m, alpha, Xte)
    const = alpha.sum() * float(m.kernel.variances[0].numpy())
    assert max_rel_err(const + np.sum(comps, axis=0), mu[:, 0]) < 1e-8


def test_svgp_gradient_with_discrete_and_empirical_dims():
    """Mixed kernel: categorical W / kappa gradients through the table cotangent, empirical-measure lengthscales."""
    import torch

    from helpers import mixed_config
    from oak_b200._gpflow_shim import Bernoulli, inv_logit
    from oak_b200.models import SVGP
    from oak_b200.training import _assign_unconstrained, trainable_parameters, training_loss_and_grad
    from oak_b200.workloads import build_kernel

    cfg = mixed_config(n=240, seed=8, depth=2)
    rng = np.random.default_rng(2)
    y = (rng.random((240, 1)) < 0.4).astype(np.float64)
    mz = cfg["Z"].shape[0]
    m = SVGP(kernel=build_kernel(cfg), likelihood=Bernoulli(invlink=inv_logit), inducing_variable=cfg["Z"],
             whiten=True, q_diag=True, q_mu=0.3 * rng.standard_normal((mz, 1)), q_sqrt=rng.uniform(0.5, 1.2, (mz, 1)))
    m.inducing_variable.Z.trainable = False
    data = (cfg["X"], y)
    params = trainable_parameters(m)
    u0 = np.concatenate([np.asarray(p.unconstrained_variable, dtype=np.float64).reshape(-1) for p in params])
    loss0, g0 = training_loss_and_grad(m, data)
    assert g0.shape == u0.shape and np.all(np.isfinite(g0))
    for seed in range(3):  # central differences along random directions
        d = np.random.default_rng(seed).standard_normal(u0.shape)
        d /= np.linalg.norm(d)
        h = 1e-5
        _assign_unconstrained(params, u0 + h * d)
        lp = m.training_loss(data)
        _assign_unconstrained(params, u0 - h * d)
        lm = m.training_loss(data)
        _assign_unconstrained(params, u0)
        fd = (lp - lm) / (2 * h)
        assert abs(fd - g0 @ d) < 2e-6 * max(1.0, abs(fd))
