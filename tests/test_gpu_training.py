"""Backward tiles and objective gradients against torch autograd of the
gpflow operation order (oracle/oak_grad_oracle.py, CPU float64).  The reference differentiates by
TensorFlow autodiff and holds no gradient fixtures, so the oracle here is autograd of the restated
forward pass, whose VALUES are pinned against oracle/oak_oracle.py in the same tests."""
import numpy as np
import pytest

from helpers import build_oracle, max_rel_err
from oracle import oak_grad_oracle as go
from oracle import oak_oracle as oo

pytestmark = pytest.mark.gpu


def _cfg(n, D, P, m, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D))
    y = (np.sin(X[:, 0]) + X[:, 1] * X[:, min(2, D - 1)] + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
    ls = rng.uniform(0.5, 2.5, D)
    var = list(rng.uniform(0.2, 1.2, P + 1))
    dims = [{"type": "rbf", "lengthscale": float(l), "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)} for l in ls]
    return dict(X=X, y=y, Z=X[:m].copy(), dims=dims, depth=P, variances=var, share_var=True, noise=0.05), ls, np.array(var)


@pytest.mark.parametrize("D,P", [(3, 1), (5, 2), (6, 3), (5, 4), (20, 3), (7, 6), (8, 8)])
def test_backward_tiles_match_autograd(D, P):
    """sum W * K(X, X2) differentiated by the tiles vs autograd, ragged sizes, cross and self."""
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg, ls, var = _cfg(210, D, P, 70, seed=10 * D + P)
    k = build_kernel(cfg)
    spec = k._make_spec()
    rng = np.random.default_rng(3)
    Xd, Zd = _device.to_device(cfg["X"]), _device.to_device(cfg["Z"])
    px, pz = _device.Points(spec, Xd), _device.Points(spec, Zd)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    for (pa, A), (pb, B) in (((px, cfg["X"]), (pz, cfg["Z"])), ((pz, cfg["Z"]), (pz, cfg["Z"]))):
        W = rng.standard_normal((A.shape[0], B.shape[0]))
        g = _device.gram_backward(spec, pa, _device.to_device(W), px2=pb).cpu().numpy()
        lsT, vT = t(ls).clone().requires_grad_(True), t(var).clone().requires_grad_(True)
        (t(W) * go.oak_K(t(A), t(B), lsT, vT)).sum().backward()
        assert max_rel_err(g[:D], lsT.grad.numpy()) < 1e-10
        assert max_rel_err(g[D:D + P + 1], vT.grad.numpy()) < 1e-10
    # K_diag
    w = rng.standard_normal(210)
    g = _device.gram_diag_backward(spec, px, wscale=0.7, w=_device.to_device(w, ndim=1)).cpu().numpy()
    lsT, vT = t(ls).clone().requires_grad_(True), t(var).clone().requires_grad_(True)
    (0.7 * t(w) * go.oak_K_diag(t(cfg["X"]), lsT, vT)).sum().backward()
    assert max_rel_err(g[:D], lsT.grad.numpy()) < 1e-10
    assert max_rel_err(g[D:D + P + 1], vT.grad.numpy()) < 1e-10
    spec.close()


def test_backward_row_ranges_add_up():
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg, ls, var = _cfg(500, 6, 3, 90, seed=5)
    k = build_kernel(cfg)
    spec = k._make_spec()
    Xd, Zd = _device.to_device(cfg["X"]), _device.to_device(cfg["Z"])
    px, pz = _device.Points(spec, Xd), _device.Points(spec, Zd)
    W = torch.randn(500, 90, dtype=torch.float64, device="cuda")
    full = _device.gram_backward(spec, px, W, px2=pz)
    acc = torch.zeros_like(full)
    for b, e in ((0, 128), (128, 320), (320, 500)):
        _device.gram_backward(spec, px, W[b:e].contiguous(), px2=pz, row_begin=b, row_end=e, grad=acc)
    assert max_rel_err(acc.cpu().numpy(), full.cpu().numpy()) < 1e-12
    spec.close()


@pytest.mark.parametrize("chunk", [128, 8192])
def test_sgpr_elbo_gradient_matches_autograd(chunk):
    from oak_b200.models import SGPR
    from oak_b200.training import freeze_unsupported, sgpr_elbo_and_grad
    from oak_b200.workloads import build_kernel

    cfg, ls, var = _cfg(600, 5, 3, 48, seed=2)
    m = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=chunk)
    m.likelihood.variance.assign(cfg["noise"])
    m.inducing_variable.Z.trainable = False  # zfixed=True, the reference's default (model_utils.py:100-101)
    freeze_unsupported(m)
    elbo, g_ls, g_var, g_noise = sgpr_elbo_and_grad(m)
    v, a_ls, a_var, a_noise = go.value_and_grad(go.sgpr_elbo, cfg["X"], cfg["y"], cfg["Z"], ls, var, cfg["noise"])
    ref = oo.sgpr_elbo(build_oracle(cfg), cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    assert abs(v - ref) < 1e-9 * abs(ref)          # the autograd oracle's value is the NumPy oracle's
    assert abs(elbo - ref) < 1e-9 * abs(ref)
    assert abs(elbo - m.elbo()) < 1e-9 * abs(ref)  # and the gpflow-order value of the product path
    assert max_rel_err(g_ls, a_ls) < 1e-7
    assert max_rel_err(g_var, a_var) < 1e-7
    assert abs(g_noise - a_noise) < 1e-7 * abs(a_noise)


def test_gpr_lml_gradient_matches_autograd():
    from oak_b200.models import GPR
    from oak_b200.training import freeze_unsupported, gpr_lml_and_grad
    from oak_b200.workloads import build_kernel

    cfg, ls, var = _cfg(300, 4, 4, 10, seed=4)
    m = GPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg))
    m.likelihood.variance.assign(cfg["noise"])
    freeze_unsupported(m)
    lml, g_ls, g_var, g_noise = gpr_lml_and_grad(m)
    v, a_ls, a_var, a_noise = go.value_and_grad(go.gpr_lml, cfg["X"], cfg["y"], None, ls, var, cfg["noise"])
    assert abs(lml - v) < 1e-9 * abs(v)
    assert max_rel_err(g_ls, a_ls) < 1e-7
    assert max_rel_err(g_var, a_var) < 1e-7
    assert abs(g_noise - a_noise) < 1e-7 * abs(a_noise)


def test_training_loss_gradient_and_bfgs_improve_the_bound():
    """d training_loss / d unconstrained variables (softplus transforms, Gamma(1, 0.2) prior on the order
    variances, model_utils.py:163-167) against autograd + the same chain, then a few BFGS steps (the
    reference's own training test only asserts that the objective improves: test_optimisation.py:45,70)."""
    from oak_b200.model_utils import create_model_oak
    from oak_b200.training import optimise, trainable_parameters, training_loss_and_grad

    rng = np.random.default_rng(7)
    X = rng.standard_normal((400, 3))
    y = (X[:, 0] ** 2 + 2 * X[:, 1] + X[:, 0] * X[:, 1] + 0.1 * rng.standard_normal(400)).reshape(-1, 1)
    y = (y - y.mean()) / y.std()
    Z = X[:40].copy()
    model = create_model_oak((X, y), max_interaction_depth=2, inducing_pts=Z, optimise=False)
    params = trainable_parameters(model)
    assert len(params) == 3 + 3 + 1  # lengthscales, order variances, likelihood variance (Z is fixed)
    loss0, g = training_loss_and_grad(model)
    assert abs(loss0 - model.training_loss()) < 1e-9 * abs(loss0)
    # expected: autograd of the ELBO, prior gradient, softplus chain (initial values: l = 1, sigma2 = 1, noise = 0.01)
    v, a_ls, a_var, a_noise = go.value_and_grad(go.sgpr_elbo, X, y, Z, np.ones(3), np.ones(3), 0.01)
    sp = lambda x, lower=0.0: 1.0 - np.exp(-(x - lower))  # d softplus / du at softplus(u) + lower = x
    # gpflow's (tf.Module) parameter order: the kernel's own order variances, then the sub-kernels, then the likelihood
    assert params[0] is model.kernel.variances[0] and params[-1] is model.likelihood.variance
    want = np.concatenate([-(a_var - 0.2) * sp(1.0), -a_ls * sp(1.0), [-a_noise * sp(0.01, 1e-6)]])
    assert max_rel_err(g, want) < 1e-7
    res = optimise(model, method="BFGS", maxiter=15)
    assert model.training_loss() < loss0 - 1.0
    assert np.isfinite(res.fun)


@pytest.mark.parametrize("sparse", [False, True])
def test_oak_model_fit_with_optimisation_end_to_end(sparse):
    """oak_model.fit(optimise=True) (model_utils.py:194-427 flow without the TFP normalising flow):
    BFGS training on y = x0^2 + 2 x1 + x0 x1 (the reference's Sobol toy, test_sobol_oak_kernel.py:41-75):
    the fit must explain the data and attribute it to {x0}, {x1}, {x0, x1} and not to x2."""
    from oak_b200.model_utils import oak_model

    rng = np.random.default_rng(3)
    N = 300
    X = rng.standard_normal((N, 3))
    y = (X[:, 0] ** 2 + 2 * X[:, 1] + X[:, 0] * X[:, 1] + 0.05 * rng.standard_normal(N)).reshape(-1, 1)
    oak = oak_model(max_interaction_depth=2, use_normalising_flow=False, sparse=sparse, num_inducing=60)
    oak.fit(X, y, optimise=False, initialise_inducing_points=False)
    loss0 = oak.m.training_loss()
    rmse0 = float(np.sqrt(np.mean((oak.predict(X) - y[:, 0]) ** 2)))
    oak.optimise()
    assert oak.m.training_loss() < loss0 - 10.0
    Xt = rng.standard_normal((200, 3))
    yt = Xt[:, 0] ** 2 + 2 * Xt[:, 1] + Xt[:, 0] * Xt[:, 1]
    rmse = float(np.sqrt(np.mean((oak.predict(Xt) - yt) ** 2)))
    assert rmse < 0.5 * max(rmse0, 0.5) and rmse < 0.6
    sob = oak.get_sobol()
    # components: [0], [1], [2], [0,1], [0,2], [1,2]
    assert sob[0] > 0.1 and sob[1] > 0.3 and sob[3] > 0.03
    assert sob[2] < 0.02 and sob[4] < 0.02 and sob[5] < 0.02


def _mixed_cfg(n, m, seed, depth=2):
    """Gaussian-measure, empirical-measure and unconstrained RBF dims + a binary and a categorical dim."""
    rng = np.random.default_rng(seed)
    X = np.zeros((n, 6))
    X[:, 0] = rng.standard_normal(n)
    X[:, 1] = np.round(4 * rng.standard_normal(n)) / 4
    X[:, 2] = rng.standard_normal(n)
    X[:, 3] = np.round(3 * rng.standard_normal(n)) / 3
    X[:, 4] = (rng.random(n) < 0.3).astype(float)
    X[:, 5] = rng.integers(0, 4, n).astype(float)
    loc1, cnt1 = np.unique(X[:, 1], return_counts=True)
    loc3, cnt3 = np.unique(X[:, 3], return_counts=True)
    ls = rng.uniform(0.6, 2.0, 6)
    dims = [
        {"type": "rbf", "lengthscale": float(ls[0]), "variance": 1.0, "measure": ("gaussian", 0.3, 1.7)},
        {"type": "rbf", "lengthscale": float(ls[1]), "variance": 1.0, "measure": ("empirical", loc1, cnt1 / cnt1.sum())},
        {"type": "rbf", "lengthscale": float(ls[2]), "variance": 1.0, "measure": None},
        {"type": "rbf", "lengthscale": float(ls[3]), "variance": 1.0, "measure": ("empirical", loc3, cnt3 / cnt3.sum())},
        {"type": "binary", "p0": 0.7, "variance": 1.0},
        {"type": "categorical", "p": np.array([0.2, 0.3, 0.1, 0.4]), "W": rng.uniform(0, 1, (4, 2)),
         "kappa": rng.uniform(0.5, 1.5, 4), "variance": 1.0},
    ]
    var = list(rng.uniform(0.3, 1.2, depth + 1))
    y = (np.sin(X[:, 0]) + X[:, 1] * X[:, 4] + 0.3 * X[:, 5] + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
    cfg = dict(X=X, y=y, Z=X[:m].copy(), dims=dims, depth=depth, variances=var, share_var=True, noise=0.05)
    ref = build_oracle(cfg)
    # oracle-side description: tables of the discrete dims come from the NumPy oracle (constants)
    measures = [("gaussian", 0.3, 1.7), ("empirical", loc1, cnt1 / cnt1.sum()), ("none",),
                ("empirical", loc3, cnt3 / cnt3.sum()), ("table", ref.dims[4].table()), ("table", ref.dims[5].table())]
    return cfg, ls, np.array(var), measures, ref


def test_backward_tiles_mixed_measures_match_autograd():
    """Lengthscale gradients under the empirical measure (per-point d c^/dl block), a non-standard
    Gaussian measure and no measure, with discrete dims in the product; order variances as well."""
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg, ls, var, measures, ref = _mixed_cfg(230, 60, seed=11)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    # the autograd oracle's forward value is the NumPy oracle's
    assert max_rel_err(go.oak_K(t(cfg["Z"]), t(cfg["X"]), t(ls), t(var), measures).numpy(), ref.K(cfg["Z"], cfg["X"])) < 1e-12
    k = build_kernel(cfg)
    spec = k._make_spec()
    px, pz = _device.Points(spec, _device.to_device(cfg["X"])), _device.Points(spec, _device.to_device(cfg["Z"]))
    rng = np.random.default_rng(1)
    W = rng.standard_normal((230, 60))
    g = _device.gram_backward(spec, px, _device.to_device(W), px2=pz).cpu().numpy()
    lsT, vT = t(ls).clone().requires_grad_(True), t(var).clone().requires_grad_(True)
    (t(W) * go.oak_K(t(cfg["X"]), t(cfg["Z"]), lsT, vT, measures)).sum().backward()
    assert max_rel_err(g[:4], lsT.grad.numpy()[:4]) < 1e-10
    assert np.all(g[4:6] == 0.0)  # discrete sub-kernels: no lengthscale, entries untouched
    assert max_rel_err(g[6:9], vT.grad.numpy()) < 1e-10
    assert g.shape[0] == 6 + 3 + (4 + 2) + (16 + 4) + 6 and np.any(g[9:35] != 0.0)  # + B tables, base variances
    w = rng.standard_normal(230)
    g = _device.gram_diag_backward(spec, px, wscale=-0.3, w=_device.to_device(w, ndim=1)).cpu().numpy()
    lsT, vT = t(ls).clone().requires_grad_(True), t(var).clone().requires_grad_(True)
    (-0.3 * t(w) * go.oak_K_diag(t(cfg["X"]), lsT, vT, measures)).sum().backward()
    assert max_rel_err(g[:4], lsT.grad.numpy()[:4]) < 1e-10
    assert max_rel_err(g[6:9], vT.grad.numpy()) < 1e-10
    spec.close()


def test_sgpr_gradient_mixed_model_with_frozen_discrete_parameters():
    from oak_b200.models import SGPR
    from oak_b200.training import freeze_unsupported, sgpr_elbo_and_grad
    from oak_b200.workloads import build_kernel

    cfg, ls, var, measures, ref = _mixed_cfg(500, 40, seed=12)
    m = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=128)
    m.likelihood.variance.assign(cfg["noise"])
    assert freeze_unsupported(m) == []  # every parameter of the mixed model has a gradient, Z included
    elbo, g_ls, g_var, g_noise = sgpr_elbo_and_grad(m)
    v, a_ls, a_var, a_noise, a_Z = go.value_and_grad(go.sgpr_elbo, cfg["X"], cfg["y"], cfg["Z"], ls, var, cfg["noise"],
                                                     measures, wrt_Z=True)
    cont = [i for i, d in enumerate(cfg["dims"]) if d["type"] == "rbf"]
    assert max_rel_err(m._inducing_grad[:, cont], a_Z[:, cont]) < 1e-7
    assert np.all(np.delete(m._inducing_grad, cont, axis=1) == 0.0)
    assert abs(v - oo.sgpr_elbo(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])) < 1e-9 * abs(v)
    assert abs(elbo - v) < 1e-9 * abs(v)
    assert max_rel_err(g_ls[:4], a_ls[:4]) < 1e-7
    assert max_rel_err(g_var, a_var) < 1e-7
    assert abs(g_noise - a_noise) < 1e-7 * abs(a_noise)


def test_categorical_W_kappa_gradients_through_the_table_cotangent():
    """d ELBO / d (W, kappa) of a categorical sub-kernel: the backward tiles return the cotangent of the
    B table, training.py chains it (ortho_categorical_kernel.py:34-53); oracle = autograd with B(W, kappa)
    rebuilt in torch."""
    import torch

    from oak_b200.models import SGPR
    from oak_b200.training import discrete_parameter_gradients, freeze_unsupported, sgpr_elbo_and_grad
    from oak_b200.workloads import build_kernel

    cfg, ls, var, measures, ref = _mixed_cfg(400, 36, seed=13)
    kern = build_kernel(cfg)
    m = SGPR((cfg["X"], cfg["y"]), kernel=kern, inducing_variable=cfg["Z"], chunk=128)
    m.likelihood.variance.assign(cfg["noise"])
    freeze_unsupported(m)
    elbo, g_ls, g_var, g_noise = sgpr_elbo_and_grad(m)
    grads = discrete_parameter_gradients(m)
    cat = kern.kernels[5]
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    W = t(cfg["dims"][5]["W"]).clone().requires_grad_(True)
    kap = t(cfg["dims"][5]["kappa"]).clone().requires_grad_(True)
    p = t(cfg["dims"][5]["p"]).reshape(-1, 1)
    A = W @ W.T + torch.diag(kap)
    Ap = A @ p
    B = A - (Ap @ Ap.T) / (p.T @ Ap)[0, 0]
    meas = list(measures)
    meas[5] = ("table", B)
    val = go.sgpr_elbo(t(cfg["X"]), t(cfg["y"]), t(cfg["Z"]), t(ls), t(var), t(cfg["noise"]), meas)
    val.backward()
    assert abs(float(val.detach()) - elbo) < 1e-9 * abs(elbo)
    assert max_rel_err(grads[id(cat.W)], W.grad.numpy()) < 1e-7
    assert max_rel_err(grads[id(cat.kappa)].reshape(-1), kap.grad.numpy()) < 1e-7


def test_mixed_input_oak_model_trains_all_its_parameters():
    """Binary + categorical + empirical-measure + continuous inputs (the oak_model flow of
    model_utils.py:194-427: a flow on the plain continuous column, standardisation on the empirical one): every trainable parameter of the reference model --
    bounded lengthscales, order variances, noise, categorical W and kappa -- receives a gradient and
    BFGS improves the bound."""
    from oak_b200.model_utils import oak_model
    from oak_b200.training import optimise, trainable_parameters, training_loss_and_grad

    rng = np.random.default_rng(44)
    N = 400
    x_cat = rng.choice([0, 1, 2, 3], size=N, p=[0.2, 0.2, 0.3, 0.3])
    x_bin = rng.choice([0, 1], size=N, p=[0.8, 0.2])
    X = np.vstack([x_bin, x_cat, rng.standard_normal(N), np.round(3 * rng.standard_normal(N)) / 3]).T.astype(float)
    y = (np.sin(X[:, 2]) + X[:, 0] + 0.4 * (X[:, 1] == 2) + 0.3 * X[:, 3] + 0.05 * rng.standard_normal(N)).reshape(-1, 1)
    oak = oak_model(binary_feature=[0], categorical_feature=[1], max_interaction_depth=2,
                    empirical_measure=[3], sparse=True, num_inducing=50)
    oak.fit(X, y, optimise=False, initialise_inducing_points=False)
    params = trainable_parameters(oak.m)
    # 2 lengthscales, the base variance of the empirical-measure dim (oak_kernel.py:163-166 fixes it only
    # under the Gaussian measure), 3 order variances, noise, W and kappa
    assert len(params) == 2 + 1 + 3 + 1 + 2
    loss0, g = training_loss_and_grad(oak.m)
    assert g.shape[0] == 2 + 1 + 3 + 1 + 4 * 2 + 4 and np.all(np.isfinite(g)) and np.all(g != 0.0)
    optimise(oak.m, method="BFGS", maxiter=25)
    assert oak.m.training_loss() < loss0 - 10.0
    pred = oak.predict(X)
    assert float(np.sqrt(np.mean((pred - y[:, 0]) ** 2))) < 0.3
    sob = oak.get_sobol()
    assert abs(sob.sum() - 1) < 1e-12 and np.all(sob >= 0)


def test_base_variance_gradient_matches_autograd():
    """d/d s^2 of RBF sub-kernels (k~ is homogeneous of degree one in s^2) for Gaussian, empirical and
    no measure; oracle: autograd with k~ scaled by s^2."""
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg, ls, var, measures, ref = _mixed_cfg(200, 50, seed=14)
    s2 = np.array([1.3, 0.7, 2.1, 1.0])
    for d in range(4):
        cfg["dims"][d]["variance"] = float(s2[d])
    k = build_kernel(cfg)
    spec = k._make_spec()
    px, pz = _device.Points(spec, _device.to_device(cfg["X"])), _device.Points(spec, _device.to_device(cfg["Z"]))
    W = np.random.default_rng(2).standard_normal((200, 50))
    g = _device.gram_backward(spec, px, _device.to_device(W), px2=pz).cpu().numpy()
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    s2T = t(s2).clone().requires_grad_(True)
    ks = go._dim_values(t(cfg["X"]), t(cfg["Z"]), t(ls), measures)
    ks = [ks[d] * s2T[d] for d in range(4)] + ks[4:]
    (t(W) * go._esp_sum(ks, t(var))).sum().backward()
    assert max_rel_err(g[-6:-2], s2T.grad.numpy()) < 1e-10
    assert np.all(g[-2:] == 0.0)
    gd = _device.gram_diag_backward(spec, px, wscale=1.0).cpu().numpy()
    s2T = t(s2).clone().requires_grad_(True)
    kd = [torch.diagonal(v) for v in go._dim_values(t(cfg["X"]), t(cfg["X"]), t(ls), measures)]
    kd = [kd[d] * s2T[d] for d in range(4)] + kd[4:]
    go._esp_sum(kd, t(var)).sum().backward()
    assert max_rel_err(gd[-6:-2], s2T.grad.numpy()) < 1e-10
    spec.close()


def test_backward_tiles_every_measure_and_base_variances():
    """helpers.mixed_config: Gaussian, uniform, empirical, MOG and unconstrained RBF dims with s^2 != 1,
    a binary and a categorical dim.  Lengthscale, base-variance and order-variance gradients of
    sum W * K and of sum w * K_diag against autograd."""
    import torch

    from helpers import mixed_config
    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg = mixed_config(n=260, seed=5, depth=3)
    ref = build_oracle(cfg)
    dims = cfg["dims"]
    rbf = [0, 1, 2, 3, 6]
    ls = np.array([dims[i].get("lengthscale", 1.0) for i in range(7)])
    s2 = np.array([dims[i].get("variance", 1.0) for i in range(7)])
    var = np.array(cfg["variances"])
    measures = []
    for i, d in enumerate(dims):
        if d["type"] != "rbf":
            measures.append(("table", ref.dims[i].table()))
        elif d["measure"] is None:
            measures.append(("none",))
        else:
            measures.append(tuple(d["measure"]))
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    Xt, Zt = t(cfg["X"]), t(cfg["Z"])
    assert max_rel_err(go._esp_sum(go._dim_values(Zt, Xt, t(ls), measures, t(s2)), t(var)).numpy(),
                       ref.K(cfg["Z"], cfg["X"])) < 1e-12
    k = build_kernel(cfg)
    spec = k._make_spec()
    px, pz = _device.Points(spec, _device.to_device(cfg["X"])), _device.Points(spec, _device.to_device(cfg["Z"]))
    W = np.random.default_rng(9).standard_normal((260, cfg["Z"].shape[0]))
    g = _device.gram_backward(spec, px, _device.to_device(W), px2=pz).cpu().numpy()
    lsT, s2T, vT = (t(a).clone().requires_grad_(True) for a in (ls, s2, var))
    (t(W) * go._esp_sum(go._dim_values(Xt, Zt, lsT, measures, s2T), vT)).sum().backward()
    assert max_rel_err(g[rbf], lsT.grad.numpy()[rbf]) < 1e-10
    assert max_rel_err(g[7:11], vT.grad.numpy()) < 1e-10
    assert max_rel_err(g[-7:][rbf], s2T.grad.numpy()[rbf]) < 1e-10
    w = np.random.default_rng(10).standard_normal(260)
    gd = _device.gram_diag_backward(spec, px, wscale=1.0, w=_device.to_device(w, ndim=1)).cpu().numpy()
    lsT, s2T, vT = (t(a).clone().requires_grad_(True) for a in (ls, s2, var))
    kd = [torch.diagonal(v) for v in go._dim_values(Xt, Xt, lsT, measures, s2T)]
    (t(w) * go._esp_sum(kd, vT)).sum().backward()
    assert max_rel_err(gd[rbf], lsT.grad.numpy()[rbf]) < 1e-10
    assert max_rel_err(gd[7:11], vT.grad.numpy()) < 1e-10
    assert max_rel_err(gd[-7:][rbf], s2T.grad.numpy()[rbf]) < 1e-10
    spec.close()


def test_oak_model_with_gmm_measures_fits_and_trains():
    """oak_model(gmm_measure=...) (model_utils.py:286-300, 753-770): the mixture is estimated by sklearn on
    the host, the MOG-constrained kernels run and train on the device (no flow is needed for those inputs)."""
    from oak_b200.model_utils import oak_model
    from oak_b200.training import optimise

    rng = np.random.default_rng(21)
    N = 300
    X = np.column_stack([np.where(rng.random(N) < 0.4, rng.normal(-2, 0.5, N), rng.normal(1.5, 0.8, N)),
                         rng.normal(0, 1, N)])
    y = (np.sin(X[:, 0]) + 0.5 * X[:, 1] + 0.05 * rng.standard_normal(N)).reshape(-1, 1)
    oak = oak_model(max_interaction_depth=2, gmm_measure=[2, 1], use_normalising_flow=True, sparse=True, num_inducing=40)
    oak.fit(X, y, optimise=False, initialise_inducing_points=False)
    assert oak.estimated_gmm_measures[0] is not None and len(oak.estimated_gmm_measures[0].weights) == 2
    loss0 = oak.m.training_loss()
    optimise(oak.m, method="BFGS", maxiter=20)
    assert oak.m.training_loss() < loss0 - 10.0
    assert float(np.sqrt(np.mean((oak.predict(X) - y[:, 0]) ** 2))) < 0.3


def test_default_oak_model_with_normalising_flow_end_to_end():
    """The reference's default path: oak_model() (use_normalising_flow=True) on skewed positive inputs --
    flows fitted on the host (normalising_flow.py), GPR on the device, BFGS training, prediction and
    Sobol indices (cf. examples/uci and test_oak_model.py of the reference)."""
    from oak_b200.model_utils import oak_model

    rng = np.random.default_rng(5)
    N = 250
    X = np.column_stack([np.exp(0.5 * rng.standard_normal(N)), rng.gamma(3.0, 1.0, N), rng.standard_normal(N) ** 2 + 0.1])
    f = lambda A: np.log(A[:, 0]) ** 2 + 0.5 * A[:, 1] + 0.0 * A[:, 2]
    y = (f(X) + 0.05 * rng.standard_normal(N)).reshape(-1, 1)
    oak = oak_model(max_interaction_depth=2)
    oak.fit(X, y, optimise=False)
    assert all(fl is not None for fl in oak.input_flows)
    Xs = oak._transform_x(X)
    assert np.all(np.abs(Xs.mean(0)) < 0.15) and np.all(np.abs(Xs.std(0) - 1) < 0.15)
    loss0 = oak.m.training_loss()
    oak.optimise()
    assert oak.m.training_loss() < loss0 - 10.0
    Xt = np.column_stack([np.exp(0.5 * rng.standard_normal(100)), rng.gamma(3.0, 1.0, 100), rng.standard_normal(100) ** 2 + 0.1])
    rmse = float(np.sqrt(np.mean((oak.predict(Xt, clip=True) - f(Xt)) ** 2)))
    assert rmse < 0.5
    sob = oak.get_sobol()
    assert sob[0] + sob[1] > 0.8 and sob[2] < 0.05
    inv = oak._get_x_inverse_transformer(0)
    np.testing.assert_allclose(inv(Xs[:, 0]), X[:, 0], rtol=1e-8)
    ll = oak.get_loglik(Xt, f(Xt).reshape(-1, 1), clip=True)  # mean predictive log density (model_utils.py:445-460)
    assert np.isfinite(ll) and ll > -1.0


def test_oak_model_switches_to_sgpr_with_kmeans_inducing_points_above_1000_points():
    """N > 1000 -> sparse GP with k-means inducing points (model_utils.py:373-391, utils.py:555-574),
    flows on the host, training on the device."""
    from oak_b200.model_utils import oak_model
    from oak_b200.models import SGPR

    rng = np.random.default_rng(8)
    N = 1200
    X = np.column_stack([rng.gamma(2.0, 1.0, N), rng.standard_normal(N)])
    y = (np.sqrt(X[:, 0]) + np.sin(2 * X[:, 1]) + 0.05 * rng.standard_normal(N)).reshape(-1, 1)
    oak = oak_model(max_interaction_depth=2, num_inducing=40)
    oak.fit(X, y, optimise=True)
    assert isinstance(oak.m, SGPR) and oak.m.inducing_variable.Z.numpy().shape == (40, 2)
    rmse = float(np.sqrt(np.mean((oak.predict(X) - y[:, 0]) ** 2)))
    assert rmse < 0.2
    sob = oak.get_sobol()
    assert abs(sob.sum() - 1) < 1e-12 and sob[2] < 0.1  # little interaction in an additive target


# ---- gradients with respect to the row points (inducing points, zfixed=False) -------------------
@pytest.mark.parametrize("D,P", [(3, 1), (5, 2), (6, 3), (5, 4), (20, 3), (7, 6), (8, 8)])
def test_row_point_gradients_match_autograd(D, P):
    """d/dZ of sum W * K(Z, X) (and of K(Z, Z') in its first argument) by the row-gradient tiles vs autograd;
    the parameter gradients of the same call equal those of the plain backward tiles."""
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg, ls, var = _cfg(333, D, P, 77, seed=20 * D + P)
    k = build_kernel(cfg)
    spec = k._make_spec()
    rng = np.random.default_rng(4)
    Z = cfg["Z"] + 0.3 * rng.standard_normal(cfg["Z"].shape)
    px, pz = _device.Points(spec, _device.to_device(cfg["X"])), _device.Points(spec, _device.to_device(Z))
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    for pb, B in ((px, cfg["X"]), (None, Z)):
        W = rng.standard_normal((Z.shape[0], B.shape[0]))
        Wd = _device.to_device(W)
        g, gz = _device.gram_backward_rows(spec, pz, Wd, px2=pb)
        g0 = _device.gram_backward(spec, pz, Wd, px2=pb)
        assert max_rel_err(g.cpu().numpy(), g0.cpu().numpy()) < 1e-12
        ZT = t(Z).clone().requires_grad_(True)
        (t(W) * go.oak_K(ZT, t(B), t(ls), t(var))).sum().backward()
        assert max_rel_err(gz.cpu().numpy(), ZT.grad.numpy()) < 1e-10
        # accumulation into a caller's buffer with a pitch
        buf = torch.ones((Z.shape[0], D + 3), dtype=torch.float64, device="cuda")
        _device.gram_backward_rows(spec, pz, Wd, px2=pb, grad_rows=buf[:, :D])
        assert max_rel_err(buf[:, :D].cpu().numpy(), 1.0 + ZT.grad.numpy()) < 1e-10
        assert float((buf[:, D:] - 1.0).abs().max()) == 0.0
    spec.close()


def test_row_point_gradients_every_measure():
    """helpers.mixed_config: Gaussian, uniform, empirical, MOG and unconstrained RBF dims with s^2 != 1 get
    their d/dZ column; the binary and categorical columns get none (tf.cast / tf.gather)."""
    import torch

    from helpers import mixed_config
    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg = mixed_config(n=300, seed=6, depth=3)
    ref = build_oracle(cfg)
    dims = cfg["dims"]
    ls = np.array([dims[i].get("lengthscale", 1.0) for i in range(7)])
    s2 = np.array([dims[i].get("variance", 1.0) for i in range(7)])
    var = np.array(cfg["variances"])
    measures = []
    for i, d in enumerate(dims):
        if d["type"] != "rbf":
            measures.append(("table", ref.dims[i].table()))
        elif d["measure"] is None:
            measures.append(("none",))
        else:
            measures.append(tuple(d["measure"]))
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    rng = np.random.default_rng(12)
    Z = cfg["Z"].copy()
    Z[:, [0, 1, 2, 3, 6]] += 0.2 * rng.standard_normal((Z.shape[0], 5))
    k = build_kernel(cfg)
    spec = k._make_spec()
    px, pz = _device.Points(spec, _device.to_device(cfg["X"])), _device.Points(spec, _device.to_device(Z))
    W = rng.standard_normal((Z.shape[0], 300))
    g, gz = _device.gram_backward_rows(spec, pz, _device.to_device(W), px2=px)
    ZT = t(Z).clone().requires_grad_(True)
    (t(W) * go._esp_sum(go._dim_values(ZT, t(cfg["X"]), t(ls), measures, t(s2)), t(var))).sum().backward()
    gz = gz.cpu().numpy()
    assert max_rel_err(gz[:, [0, 1, 2, 3, 6]], ZT.grad.numpy()[:, [0, 1, 2, 3, 6]]) < 1e-10
    assert np.all(gz[:, [4, 5]] == 0.0)
    g0 = _device.gram_backward(spec, pz, _device.to_device(W), px2=px)
    assert max_rel_err(g.cpu().numpy(), g0.cpu().numpy()) < 1e-12
    spec.close()


@pytest.mark.parametrize("chunk,keep", [(128, True), (8192, True), (256, False)])
def test_sgpr_elbo_gradient_wrt_inducing_points_matches_autograd(chunk, keep):
    """zfixed=False (model_utils.py:98-101): d ELBO / d Z through Kuf and Kuu, with Kuf kept or recomputed."""
    from oak_b200.models import SGPR
    from oak_b200.training import freeze_unsupported, sgpr_elbo_and_grad
    from oak_b200.workloads import build_kernel

    cfg, ls, var = _cfg(700, 5, 3, 40, seed=7)
    Z = cfg["Z"] + 0.1 * np.random.default_rng(1).standard_normal(cfg["Z"].shape)
    m = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=Z, chunk=chunk)
    m.keep_kuf = keep
    m.likelihood.variance.assign(cfg["noise"])
    frozen = freeze_unsupported(m)
    assert m.inducing_variable.Z.trainable and all(p is not m.inducing_variable.Z for p in frozen)
    elbo, g_ls, g_var, g_noise = sgpr_elbo_and_grad(m)
    v, a_ls, a_var, a_noise, a_Z = go.value_and_grad(go.sgpr_elbo, cfg["X"], cfg["y"], Z, ls, var, cfg["noise"],
                                                     wrt_Z=True)
    assert abs(elbo - v) < 1e-9 * abs(v)
    assert max_rel_err(g_ls, a_ls) < 1e-7
    assert max_rel_err(g_var, a_var) < 1e-7
    assert abs(g_noise - a_noise) < 1e-7 * abs(a_noise)
    assert m._inducing_grad.shape == Z.shape
    assert max_rel_err(m._inducing_grad, a_Z) < 1e-7


def test_training_with_trainable_inducing_points_improves_the_bound():
    """create_model_oak(..., zfixed=False) + BFGS: Z is part of the trainable variables, the loss gradient
    agrees with central differences along a random direction, and the optimised bound beats zfixed=True's start."""
    from oak_b200.model_utils import create_model_oak
    from oak_b200.training import _assign_unconstrained, optimise, trainable_parameters, training_loss_and_grad

    rng = np.random.default_rng(3)
    X = rng.standard_normal((500, 3))
    y = (np.sin(2 * X[:, 0]) + X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(500)).reshape(-1, 1)
    Z0 = X[:25].copy()
    m = create_model_oak((X, y), max_interaction_depth=2, inducing_pts=Z0, optimise=False, zfixed=False)
    params = trainable_parameters(m)
    assert any(p is m.inducing_variable.Z for p in params)
    u0 = np.concatenate([np.asarray(p.unconstrained_variable, dtype=np.float64).reshape(-1) for p in params])
    loss0, g0 = training_loss_and_grad(m)
    assert g0.shape == u0.shape
    d = rng.standard_normal(u0.shape)
    d /= np.linalg.norm(d)
    h = 1e-5
    _assign_unconstrained(params, u0 + h * d)
    lp = training_loss_and_grad(m)[0]
    _assign_unconstrained(params, u0 - h * d)
    lm = training_loss_and_grad(m)[0]
    _assign_unconstrained(params, u0)
    fd = (lp - lm) / (2 * h)
    assert abs(fd - g0 @ d) < 1e-5 * max(1.0, abs(fd))
    res = optimise(m, maxiter=30)
    assert res.fun < loss0 - 1.0
    assert np.abs(m.inducing_variable.Z.numpy() - Z0).max() > 1e-3


# ---- share_var_across_orders=False (Duvenaud-style prod(1 + k_i)) ------------------------------
@pytest.mark.parametrize("sparse", [True, False])
def test_share_var_false_model_has_gradients_and_trains(sparse):
    """create_model_oak(share_var_across_orders=False): only variances[0] exists, the sub-kernels' own variances
    (RBF base variance, binary variance) are trainable (oak_kernel.py:163-166, 217-221).  The gradient of the
    training loss in the unconstrained variables is checked against central differences of the loss itself, then
    BFGS must improve it (the reference trains this model: test_oak_model / model_utils.py:395-427)."""
    from oak_b200.model_utils import create_model_oak
    from oak_b200.training import _assign_unconstrained, optimise, trainable_parameters, training_loss_and_grad

    rng = np.random.default_rng(21)
    n = 260
    X = rng.standard_normal((n, 3))
    X[:, 2] = (rng.random(n) < 0.4).astype(float)
    y = (np.sin(X[:, 0]) + 0.5 * X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
    y = (y - y.mean()) / y.std()
    model = create_model_oak((X, y), max_interaction_depth=2, inducing_pts=X[:30].copy() if sparse else None,
                             optimise=False, share_var_across_orders=False, p0=[None, None, 0.6], p=[None, None, None])
    assert len(model.kernel.variances) == 1
    params = trainable_parameters(model)
    # variances[0], 2 x (RBF variance, lengthscale), binary variance, likelihood variance
    assert len(params) == 1 + 4 + 1 + 1
    u0 = np.concatenate([np.asarray(p.unconstrained_variable, dtype=np.float64).reshape(-1) for p in params])
    loss0, g = training_loss_and_grad(model)
    assert np.isfinite(loss0) and g.shape == u0.shape and np.all(g != 0.0)
    h = 1e-5
    fd = np.zeros_like(u0)
    for i in range(u0.size):
        up, um = u0.copy(), u0.copy()
        up[i] += h
        um[i] -= h
        _assign_unconstrained(params, up)
        lp = model.training_loss()
        _assign_unconstrained(params, um)
        lm = model.training_loss()
        fd[i] = (lp - lm) / (2 * h)
    _assign_unconstrained(params, u0)
    assert max_rel_err(g, fd) < 1e-5
    res = optimise(model, method="BFGS", maxiter=8)
    assert res.fun < loss0 - 1e-3


def test_optimise_raises_when_the_initial_point_is_not_positive_definite():
    """A Cholesky failure at the FIRST evaluation is the model's problem, not a line-search trial step: gpflow's
    Scipy wrapper raises there, and so must optimise() instead of returning 'converged' with a zero gradient."""
    from oak_b200._cabi import OakNativeError
    from oak_b200.model_utils import create_model_oak
    from oak_b200.training import optimise

    rng = np.random.default_rng(2)
    X = rng.standard_normal((120, 2))
    y = rng.standard_normal((120, 1))
    X[7, 0] = np.nan  # poisons Kuu / Kuf: the factorisation fails at the initial point
    model = create_model_oak((X, y), max_interaction_depth=2, inducing_pts=X[:20].copy(), optimise=False)
    with pytest.raises((RuntimeError, OakNativeError)):
        optimise(model, method="BFGS", maxiter=3)
