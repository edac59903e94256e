"""GPU parity of the model-level quantities: GPR log marginal likelihood, SGPR ELBO, alpha, Sobol
indices, per-component predictions -- CUDA path (C ABI) vs the NumPy oracle, 1e-9 relative."""
import numpy as np
import pytest

from helpers import RTOL, build_oracle, max_rel_err, mixed_config
from oracle import oak_oracle as oo

pytestmark = pytest.mark.gpu


def _models(cfg, sparse):
    from oak_b200.models import GPR, SGPR
    from oak_b200.workloads import build_kernel

    k = build_kernel(cfg)
    if sparse:
        m = SGPR((cfg["X"], cfg["y"]), kernel=k, inducing_variable=cfg["Z"], chunk=cfg.get("chunk", 8192))
    else:
        m = GPR((cfg["X"], cfg["y"]), kernel=k)
    m.likelihood.variance.assign(cfg["noise"])
    return m


def _sobol_cfg(n=400, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.normal(0, 1, (n, 2))
    y = (X[:, 0] ** 2 + 2 * X[:, 1] + X[:, 0] * X[:, 1]).reshape(-1, 1)
    dims = [{"type": "rbf", "lengthscale": l, "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)} for l in (2.91, 9.20)]
    return dict(X=X, y=y, Z=X[:250].copy(), dims=dims, depth=2, variances=[0.76, 96.935, 128.27], share_var=True,
                noise=0.01)


def test_gpr_lml_and_alpha():
    cfg = mixed_config(n=300, seed=1, depth=3)
    m, ref = _models(cfg, False), build_oracle(cfg)
    lml = m.log_marginal_likelihood()
    lml_ref = oo.gpr_log_marginal_likelihood(ref, cfg["X"], cfg["y"], cfg["noise"])
    assert abs(lml - lml_ref) / abs(lml_ref) < RTOL
    alpha = m.sufficient_statistics().cpu().numpy()
    alpha_ref = oo.gpr_alpha(ref, cfg["X"], cfg["y"], cfg["noise"])
    assert max_rel_err(alpha, alpha_ref) < 1e-7  # conditioned by (K + noise I)^-1


@pytest.mark.parametrize("chunk", [64, 128, 8192])
def test_sgpr_elbo_and_alpha(chunk):
    cfg = mixed_config(n=700, seed=2, depth=2)
    cfg["chunk"] = chunk
    m, ref = _models(cfg, True), build_oracle(cfg)
    elbo = m.elbo()
    elbo_ref = oo.sgpr_elbo(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    assert abs(elbo - elbo_ref) / abs(elbo_ref) < RTOL
    assert m.maximum_log_likelihood_objective() == elbo
    alpha = m.sufficient_statistics().cpu().numpy()
    alpha_ref = oo.sgpr_alpha(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    assert max_rel_err(alpha, alpha_ref) < 1e-6


def test_sgpr_elbo_config_C_shape_small():
    from oak_b200.workloads import config_C

    cfg = config_C(n=5000, D=20, m=256, depth=3)
    m, ref = _models(cfg, True), build_oracle(cfg)
    elbo = m.elbo()
    elbo_ref = oo.sgpr_elbo(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    assert abs(elbo - elbo_ref) / abs(elbo_ref) < RTOL


def test_sgpr_virtual_sharding_sums_to_full_stats():
    """N-axis sharding: per-shard statistics add up to the unsharded ones."""
    import torch

    from oak_b200 import _device
    from oak_b200.parallel import partition_rows
    from oak_b200.workloads import build_kernel

    cfg = mixed_config(n=900, seed=4, depth=2)
    k = build_kernel(cfg)
    spec = k._make_spec()
    Xd, Zd, yd = _device.to_device(cfg["X"]), _device.to_device(cfg["Z"]), _device.to_device(cfg["y"])
    pz = _device.Points(spec, Zd)
    full = _device.sgpr_stats(spec, pz, _device.Points(spec, Xd), yd, chunk=128)
    for G in (2, 4, 8):
        acc = torch.zeros_like(full)
        for b, e in partition_rows(900, G):
            acc += _device.sgpr_stats(spec, pz, _device.Points(spec, Xd[b:e].contiguous()), yd[b:e].contiguous(),
                                      chunk=128)
        assert max_rel_err(acc.cpu().numpy(), full.cpu().numpy()) < 1e-13
    spec.close()


@pytest.mark.parametrize("sparse", [False, True])
def test_sobol_known_answer_and_oracle(sparse):
    """tests/test_sobol_oak_kernel.py:35-126 of the reference: Sobol ~ [2, 4, 1] for
    y = x0^2 + 2 x1 + x0 x1 at the fixed hyper-parameters, plus 1e-9 parity with the oracle."""
    from oak_b200.utils import compute_sobol_oak

    cfg = _sobol_cfg()
    m, ref = _models(cfg, sparse), build_oracle(cfg)
    idx, sob = compute_sobol_oak(m, 1.0, 0.0)
    assert idx == [[0], [1], [0, 1]]
    np.testing.assert_array_almost_equal(sob, np.array([2.0, 4.0, 1.0]), decimal=1)
    if sparse:
        a_ref = oo.sgpr_alpha(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
        _, sob_ref = oo.sobol_oak(ref, cfg["Z"], a_ref)
    else:
        a_ref = oo.gpr_alpha(ref, cfg["X"], cfg["y"], cfg["noise"])
        _, sob_ref = oo.sobol_oak(ref, cfg["X"], a_ref)
    # alpha carries the conditioning of the solve; the quadratic forms themselves are 1e-9 clean
    assert max_rel_err(sob, sob_ref) < 1e-6


def test_sobol_quadforms_with_oracle_alpha():
    """Isolates the Sobol tiles from the conditioning of alpha: same alpha on both sides."""
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg = mixed_config(n=120, seed=6, depth=2)
    cfg["dims"] = [d for i, d in enumerate(cfg["dims"]) if i not in (3, 6)]  # drop MOG + unconstrained
    cfg["X"] = np.delete(cfg["X"], [3, 6], axis=1)
    ref = build_oracle(cfg)
    k = build_kernel(cfg)
    Xc = cfg["X"][:60]
    alpha = np.random.default_rng(0).standard_normal((60, 1))
    comps, sob_ref = oo.sobol_oak(ref, Xc, alpha, delta=1.0, mu=0.0)
    spec = k._make_spec()
    Xd = _device.to_device(Xc)
    Ls = torch.stack([_device.sobol_L(spec, d, Xd, 1.0, 0.0) for d in range(len(cfg["dims"]))])
    scales = []
    for S in comps:
        first = cfg["dims"][S[0]]
        v = cfg["variances"][len(S)]
        scales.append(v if first["type"] == "binary" else v ** 2)
    sob = _device.sobol_quadforms(Ls, comps, scales, _device.to_device(alpha)).cpu().numpy()
    spec.close()
    assert max_rel_err(sob, sob_ref) < RTOL


def test_compute_L_entry_points():
    from oak_b200.utils import compute_L, compute_L_binary_kernel, compute_L_categorical_kernel

    rng = np.random.default_rng(3)
    X = rng.standard_normal((50, 2))
    assert max_rel_err(compute_L(X, 1.3, 2.5, 1, 1.0, 0.0), oo.L_gaussian(X[:, 1], 1.3, 2.5, 1.0, 0.0)) < RTOL
    Xb = (rng.random((40, 1)) < 0.77).astype(float)
    for p in (0.0, 0.77, 1.0):
        assert np.max(np.abs(compute_L_binary_kernel(Xb, p, 2.5, 0) - oo.L_binary(Xb[:, 0], p, 2.5))) < 1e-15
    Xc = rng.integers(0, 4, (30, 1)).astype(float)
    W, kappa, p = rng.uniform(0, 1, (4, 2)), np.ones(4), np.array([0.1, 0.2, 0.3, 0.4]).reshape(-1, 1)
    assert max_rel_err(compute_L_categorical_kernel(Xc, W, kappa, p, 1.7, 0),
                       oo.L_categorical(Xc[:, 0], W, kappa, p, 1.7)) < RTOL


def test_binary_L_identity():
    """tests/test_sobol.py:188-208 of the reference: L == K(X,0)K(0,X) p0 + K(X,1)K(1,X) p1."""
    from oak_b200.ortho_binary_kernel import OrthogonalBinary
    from oak_b200.utils import compute_L_binary_kernel

    for p in (0.0, 0.77, 1.0):
        X = np.random.default_rng(0).binomial(1, p, 300).reshape(-1, 1).astype(float)
        L = compute_L_binary_kernel(X, p, 1, 0)
        K = OrthogonalBinary(p0=p, active_dims=[0])
        x0, x1 = np.zeros((1, 1)), np.ones((1, 1))
        L1 = K(X, x0) @ K(x0, X) * p + K(X, x1) @ K(x1, X) * (1 - p)
        assert np.max(np.abs(L - L1)) < 1e-15


@pytest.mark.parametrize("sparse", [False, True])
def test_prediction_components_sum_to_prediction(sparse):
    """tests/test_utils.py:43-75 of the reference (sum of components == predict_f) + oracle parity."""
    from oak_b200.utils import get_model_sufficient_statistics, get_prediction_component

    cfg = mixed_config(n=260, seed=7, depth=2)
    cfg["variances"] = [1e-16, 1.0, 0.5]
    m, ref = _models(cfg, sparse), build_oracle(cfg)
    alpha = get_model_sufficient_statistics(m, get_L=False)
    comps = get_prediction_component(m, alpha, cfg["X"])
    total = np.sum(comps, axis=0)
    mean, var = m.predict_f(cfg["X"])
    np.testing.assert_allclose(total, mean[:, 0], rtol=1e-7, atol=1e-9)
    if sparse:
        mean_ref = oo.sgpr_predict_mean(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"], cfg["X"])
        Xc = cfg["Z"]
    else:
        mean_ref = oo.gpr_predict_mean(ref, cfg["X"], cfg["y"], cfg["noise"], cfg["X"])
        Xc = cfg["X"]
    assert max_rel_err(mean, mean_ref) < 1e-6
    comps_ref = oo.predict_components(ref, Xc, alpha, cfg["X"])
    assert max_rel_err(np.array(comps), np.array(comps_ref)) < RTOL
    assert np.all(var > -1e-8)


def test_kernel_components_sum_to_kernel(concrete_normalised_10_rows_data):
    """tests/test_oak_kernel.py:32-144 of the reference."""
    from oak_b200.oak_kernel import KernelComponenent, OAKKernel, get_list_representation
    from oak_b200.ortho_rbf_kernel import RBF

    X, _ = concrete_normalised_10_rows_data
    x_try = X[:, :2]
    k = OAKKernel([RBF, RBF], num_dims=2, max_interaction_depth=2, constrain_orthogonal=True)
    k.variances[0].assign(1.3)
    k.variances[1].assign(3.3)
    k.variances[2].assign(4.3)
    parts = [KernelComponenent(k, s)(x_try) for s in ([], [0], [1], [0, 1])]
    np.testing.assert_allclose(k(x_try), np.sum(parts, axis=0), rtol=1e-7)
    dparts = [KernelComponenent(k, s).K_diag(x_try) for s in ([], [0], [1], [0, 1])]
    np.testing.assert_allclose(k.K_diag(x_try), np.sum(dparts, axis=0), rtol=1e-7)
    sel, kl = get_list_representation(k, num_dims=2)
    assert sel == [[], [0], [1], [0, 1]]
    np.testing.assert_allclose(k.K_diag(X), np.diag(k(X)), rtol=1e-7)
    np.testing.assert_allclose(k(X), np.sum([c(X) for c in kl], axis=0), rtol=1e-7)


@pytest.mark.parametrize("num_dims", [3, 4])
def test_newton_girard_entry_point(num_dims):
    """tests/test_kernel_properties.py:70-86 of the reference."""
    from functools import reduce
    from itertools import combinations

    from oak_b200.oak_kernel import OAKKernel
    from oak_b200.ortho_rbf_kernel import RBF

    k = OAKKernel([RBF] * num_dims, num_dims=num_dims, max_interaction_depth=num_dims)
    xx = [np.random.randn(2, 2) for _ in range(num_dims)]
    result = k.compute_additive_terms(xx)
    hard = [np.ones((2, 2))] + [reduce(np.add, map(lambda x: np.prod(x, axis=0), combinations(xx, i)))
                                for i in range(1, num_dims + 1)]
    assert len(result) == num_dims + 1
    for r1, r2 in zip(result, hard):
        np.testing.assert_allclose(r1, r2, rtol=1e-7, atol=1e-12)


def test_oak_model_api_end_to_end():
    from oak_b200.model_utils import oak_model

    rng = np.random.default_rng(44)
    N = 300
    x_cat = rng.choice([0, 1, 2, 3], size=N, p=[0.2, 0.2, 0.3, 0.3])
    x_bin = rng.choice([0, 1], size=N, p=[0.8, 0.2])
    X = np.vstack([x_bin, x_cat, rng.standard_normal(N), rng.standard_normal(N)]).T.astype(float)
    y = (np.sin(X[:, 2]) + X[:, 0] + 0.05 * rng.standard_normal(N)).reshape(-1, 1)
    oak = oak_model(binary_feature=[0], categorical_feature=[1], max_interaction_depth=2,
                    empirical_measure=[3], sparse=True, num_inducing=40)
    oak.fit(X, y, optimise=False, initialise_inducing_points=False)
    assert not np.isnan(oak.m.elbo())
    pred = oak.predict(X)
    assert pred.shape == (N,)
    sob = oak.get_sobol()
    assert len(sob) == 4 + 6 and abs(sob.sum() - 1) < 1e-12 and np.all(sob >= 0)


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 6, 8])
def test_fused_gram_matvec_matches_matrix_product(depth):
    """oak_gram_matvec_f64 (prediction mean without the N* x M matrix) == K(X*, Z) @ alpha, and the oracle."""
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg = mixed_config(n=333, seed=20 + depth, depth=depth)
    k, ref = build_kernel(cfg), build_oracle(cfg)
    spec = k._make_spec()
    Xd, Zd = _device.to_device(cfg["X"]), _device.to_device(cfg["Z"])
    px, pz = _device.Points(spec, Xd), _device.Points(spec, Zd)
    alpha = np.random.default_rng(depth).standard_normal(cfg["Z"].shape[0])
    got = _device.gram_matvec(spec, px, pz, _device.to_device(alpha, ndim=1)).cpu().numpy()
    full = (_device.gram(spec, px, pz) @ _device.to_device(alpha, ndim=1)).cpu().numpy()
    want = ref.K(cfg["X"], cfg["Z"]) @ alpha
    assert max_rel_err(got, full) < 1e-12
    assert max_rel_err(got, want) < RTOL
    spec.close()


def test_sgpr_predict_mean_matches_predict_f_and_oracle():
    cfg = mixed_config(n=500, seed=31, depth=2)
    m, ref = _models(cfg, True), build_oracle(cfg)
    Xnew = mixed_config(n=170, seed=32, depth=2)["X"]
    mean = m.predict_mean(Xnew)
    mean_f, _ = m.predict_f(Xnew)
    want = oo.sgpr_predict_mean(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"], Xnew)
    assert max_rel_err(mean, mean_f) < 1e-10
    assert max_rel_err(mean, want) < 1e-6


def test_sobol_closed_form_terms_match_oracle_and_monte_carlo():
    """f1..f4 (oak/utils.py:116-165) evaluated by the device functions that build compute_L: against the NumPy
    oracle to rounding, and -- as the reference's tests/test_sobol.py:34-140 do -- against a Monte-Carlo estimate
    of the integrals they stand for (f1 = E_s[k(x,s) k(s,y)], f4 = E_s[k(x,s)] E_s'[k(s',y)] c, ...)."""
    from oak_b200 import utils
    from oracle import oak_oracle as oo

    rng = np.random.default_rng(3)
    x, y = rng.standard_normal(500), rng.standard_normal(500)
    sigma, l, delta, mu = 1.3, 0.8, 1.5, 0.4
    for name in ("f1", "f2", "f4"):
        got = getattr(utils, name)(x, y, sigma, l, delta, mu)
        assert max_rel_err(got, getattr(oo, name)(x, y, sigma, l, delta, mu)) < 1e-13
    assert max_rel_err(utils.f3(x, y, sigma, l, delta, mu), oo.f2(y, x, sigma, l, delta, mu)) < 1e-13
    assert utils.f1(x[:6].reshape(2, 3), y[:6].reshape(2, 3), sigma, l, delta, mu).shape == (2, 3)
    # Monte Carlo: f1(x, y) = E_{s ~ N(mu, delta^2)}[k(x, s) k(s, y)] with k the RBF of variance sigma^2
    s = mu + delta * rng.standard_normal(400_000)
    k = lambda a, b: sigma ** 2 * np.exp(-0.5 * (a - b) ** 2 / l ** 2)
    for i in range(3):
        mc = np.mean(k(x[i], s) * k(s, y[i]))
        assert abs(float(utils.f1(x[i], y[i], sigma, l, delta, mu)) - mc) < 5e-3 * sigma ** 4
    # L = f1 - f2 - f3 + f4 is what compute_L returns (utils.py:221-240)
    X = np.stack([x[:40], y[:40]], axis=1)
    L = utils.compute_L(X, l, sigma ** 2, 0, delta, mu)
    xi, xj = np.meshgrid(X[:, 0], X[:, 0], indexing="ij")
    want = sum(sgn * getattr(utils, f)(xi, xj, sigma, l, delta, mu) for f, sgn in (("f1", 1), ("f2", -1), ("f3", -1), ("f4", 1)))
    assert max_rel_err(L, want) < 1e-12
