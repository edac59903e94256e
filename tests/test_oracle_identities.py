"""The reference's own identity / known-answer tests for this path (starred
rows), ported onto the NumPy oracle.  CPU only."""
import numpy as np
import pytest

from helpers import build_oracle, max_rel_err, mixed_config
from oracle import oak_oracle as oo


def _kernels_1d():
    x3 = np.array([[0.1], [0.5], [0.5]])
    return [
        oo.RBFDim(1.0, 1.0, oo.Gaussian(0, 1)), oo.RBFDim(1.0, 1.0, oo.Uniform(0, 1)),
        oo.RBFDim(1.0, 1.0, oo.Empirical(x3)), oo.RBFDim(1.0, 1.0, oo.MOG([3.0, 2.0], [3.0, 10.0], [0.6, 0.4])),
        oo.RBFDim(1.0, 1.0, None), oo.BinaryDim(0.5),
    ]


@pytest.mark.parametrize("k", _kernels_1d())
def test_kernel_1d_diag_identity(k):
    """reference tests/test_kernel_properties.py:57-66"""
    X = np.array([[0.1], [0.5], [0.5]]) if not isinstance(k, oo.BinaryDim) else np.array([[0.0], [1.0], [1.0]])
    np.testing.assert_allclose(np.diag(k.K(X, X)), k.K_diag(X), rtol=1e-7)
    np.testing.assert_allclose(k.K(X), k.K(X, X), rtol=1e-12)


@pytest.mark.parametrize("num_dims", [3, 4, 8])
def test_newton_girard_vs_bruteforce_and_direct(num_dims):
    """reference tests/test_kernel_properties.py:70-86 (extended to the top order)"""
    rng = np.random.default_rng(num_dims)
    xx = [rng.standard_normal((2, 2)) for _ in range(num_dims)]
    a = oo.newton_girard(xx, num_dims)
    b = oo.esp_bruteforce(xx, num_dims)
    c = oo.esp_dp(xx, num_dims)
    for r1, r2, r3 in zip(a, b, c):
        np.testing.assert_allclose(r1, r2, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(r3, r2, rtol=1e-10, atol=1e-12)


def test_kernel_equals_sum_of_components(concrete_normalised_10_rows_data):
    """reference tests/test_oak_kernel.py:32-144"""
    X, _ = concrete_normalised_10_rows_data
    for depth, D in ((0, 1), (1, 1), (1, 2), (2, 2), (2, 7)):
        dims = [oo.RBFDim(1.0, 1.0, oo.Gaussian(0, 1)) for _ in range(D)]
        k = oo.OakOracle(dims, depth, [1.3, 3.3, 4.3][: depth + 1])
        x = X[:, :D]
        subs = oo.subsets(D, depth)
        if D == 2 and depth == 2:
            assert subs == [[], [0], [1], [0, 1]]
        np.testing.assert_allclose(k.K(x), sum(k.component_K(S, x) for S in subs), rtol=1e-7)
        np.testing.assert_allclose(k.K_diag(x), sum(k.component_K_diag(S, x) for S in subs), rtol=1e-7)
        np.testing.assert_allclose(k.K_diag(x), np.diag(k.K(x)), rtol=1e-7)


def test_mog_equals_gaussian():
    """reference tests/test_orthogonality.py:152-165"""
    kg = oo.RBFDim(10.0, 1.0, oo.Gaussian(3, 5))
    km = oo.RBFDim(10.0, 1.0, oo.MOG([3.0, 3.0], [5.0, 5.0], [0.2, 0.8]))
    xx = np.array([[-2], [2.0], [3.0]])
    np.testing.assert_allclose(kg.K(xx), km.K(xx), rtol=1e-7)


@pytest.mark.parametrize("meas", [oo.Gaussian(0.0, 1.0), oo.Gaussian(0.4, 2.0), oo.Uniform(-1.0, 2.0),
                                  oo.MOG([0.5, -1.0], [1.5, 0.7], [0.6, 0.4])])
def test_orthogonality_by_quadrature(meas):
    """int k~(x, s) p(s) ds = 0 and cov_X_s / var_s by quadrature (strengthens the 2-decimal Monte-Carlo
    checks of reference tests/test_orthogonality.py:27-76)."""
    from scipy import integrate, stats

    k = oo.RBFDim(0.8, 1.4, meas)
    if isinstance(meas, oo.Gaussian):
        pdf, lo, hi = (lambda s: stats.norm.pdf(s, meas.mu, np.sqrt(meas.var))), meas.mu - 12, meas.mu + 12
    elif isinstance(meas, oo.Uniform):
        pdf, lo, hi = (lambda s: 1.0 / (meas.b - meas.a)), meas.a, meas.b
    else:
        pdf = lambda s: sum(w * stats.norm.pdf(s, m, np.sqrt(v)) for m, v, w in zip(meas.means, meas.variances, meas.weights))
        lo, hi = -15, 15
    base = lambda x, s: 1.4 * np.exp(-0.5 * (x - s) ** 2 / 0.8 ** 2)
    for x0 in (-0.7, 0.3, 1.9):
        c_quad = integrate.quad(lambda s: base(x0, s) * pdf(s), lo, hi, epsabs=1e-13, epsrel=1e-13)[0]
        assert abs(c_quad - k.cov_X_s(np.array([[x0]]))[0, 0]) < 1e-10
        kt = lambda s: k.K(np.array([[x0]]), np.array([[s]]))[0, 0]
        assert abs(integrate.quad(lambda s: kt(s) * pdf(s), lo, hi, epsabs=1e-13, epsrel=1e-13)[0]) < 1e-10
    v_quad = integrate.quad(lambda s: k.cov_X_s(np.array([[s]]))[0, 0] * pdf(s), lo, hi, epsabs=1e-13, epsrel=1e-13)[0]
    assert abs(v_quad - k.var_s()) < 1e-10


def test_gaussian_L_by_quadrature():
    """L[i,j] = int k~(x_i,s) k~(s,x_j) p(s) ds for the Gaussian measure (utils.py:116-165, 221-240;
    the reference checks f1/f2/f4 by 1e-3 Monte Carlo, tests/test_sobol.py:34-140)."""
    from scipy import integrate, stats

    l, delta, mu = 1.3, 1.0, 0.0
    k = oo.RBFDim(l, 1.0, oo.Gaussian(mu, delta ** 2))
    xs = np.array([-1.1, 0.2, 0.9])
    L = oo.L_gaussian(xs, l, 1.0, delta, mu)
    for i, xi in enumerate(xs):
        for j, xj in enumerate(xs):
            f = lambda s: k.K(np.array([[xi]]), np.array([[s]]))[0, 0] * k.K(np.array([[s]]), np.array([[xj]]))[0, 0] \
                * stats.norm.pdf(s, mu, delta)
            assert abs(integrate.quad(f, -12, 12, epsabs=1e-13, epsrel=1e-13)[0] - L[i, j]) < 1e-10


@pytest.mark.parametrize("p", (0.0, 0.77, 1.0))
def test_binary_L_identity(p):
    """reference tests/test_sobol.py:188-208"""
    X = np.random.default_rng(0).binomial(1, p, 200).astype(float)
    L = oo.L_binary(X, p, 1)
    K = oo.BinaryDim(p)
    x0, x1, Xc = np.zeros((1, 1)), np.ones((1, 1)), X.reshape(-1, 1)
    L1 = K.K(Xc, x0) @ K.K(x0, Xc) * p + K.K(Xc, x1) @ K.K(x1, Xc) * (1 - p)
    assert np.max(np.abs(L - L1)) < 1e-16


@pytest.mark.parametrize("sparse", [False, True])
def test_sobol_known_answer(sparse):
    """reference tests/test_sobol_oak_kernel.py:35-126: Sobol ~ [2, 4, 1] for y = x0^2 + 2 x1 + x0 x1"""
    rng = np.random.default_rng(1)
    X = rng.normal(0, 1, (500, 2))
    Y = (X[:, 0] ** 2 + 2 * X[:, 1] + X[:, 0] * X[:, 1]).reshape(-1, 1)
    k = oo.OakOracle([oo.RBFDim(2.91), oo.RBFDim(9.20)], 2, [0.76, 96.935, 128.27])
    if sparse:
        Z = X[:300]
        idx, sob = oo.sobol_oak(k, Z, oo.sgpr_alpha(k, X, Y, Z, 0.01))
    else:
        idx, sob = oo.sobol_oak(k, X, oo.gpr_alpha(k, X, Y, 0.01))
    assert idx == [[0], [1], [0, 1]]
    np.testing.assert_array_almost_equal(sob, [2.0, 4.0, 1.0], decimal=1)


def test_empirical_sobol_equals_sample_variance():
    """reference tests/test_sobol_oak_kernel.py:129-152"""
    rng = np.random.default_rng(3)
    x = rng.normal(0, 1, (10, 1))
    y = x ** 2 + np.cos(x) + rng.normal(0, 0.1, (10, 1))
    dim = oo.RBFDim(1.0, 1.0, oo.Empirical(x, np.ones(x.shape) / 10))
    k = oo.OakOracle([dim], 1, [0.0, 1.0])
    alpha = oo.gpr_alpha(k, x, y, 1.0)
    var_samples = np.var(oo.gpr_predict_mean(k, x, y, 1.0, x))
    L = oo.L_empirical(x, np.ones(x.shape) / 10, dim, x[:, 0])
    np.testing.assert_array_almost_equal(var_samples, (alpha.T @ L @ alpha)[0, 0], decimal=5)


def test_components_sum_to_prediction():
    """reference tests/test_utils.py:43-75"""
    cfg = mixed_config(n=120, seed=9, depth=2)
    cfg["variances"] = [1e-16, 1.0, 0.5]
    k = build_oracle(cfg)
    alpha = oo.sgpr_alpha(k, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    total = np.sum(oo.predict_components(k, cfg["Z"], alpha, cfg["X"]), axis=0)
    np.testing.assert_allclose(total, oo.sgpr_predict_mean(k, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"], cfg["X"])[:, 0],
                               rtol=1e-7, atol=1e-9)


def test_cpu_baseline_twin_matches_oracle():
    """bench.py's timed CPU comparator computes the same numbers as the oracle."""
    import torch

    from oak_b200.workloads import config_D
    from oracle import cpu_baseline as cb

    cfg = config_D(300, 48)
    ref = build_oracle(cfg, expanded=True)
    X, Y, Z = (torch.as_tensor(cfg[k]) for k in ("X", "y", "Z"))
    assert max_rel_err(cb.gram(cfg, X).numpy(), ref.K(cfg["X"])) < 1e-12
    assert max_rel_err(cb.gram(cfg, Z, X).numpy(), ref.K(cfg["Z"], cfg["X"])) < 1e-12
    assert max_rel_err(cb.gram_diag(cfg, X).numpy(), ref.K_diag(cfg["X"])) < 1e-12
    e1, e2 = cb.sgpr_elbo(cfg, X, Y, Z, 0.01), oo.sgpr_elbo(ref, cfg["X"], cfg["y"], cfg["Z"], 0.01)
    assert abs(e1 - e2) < 1e-9 * abs(e2)


def test_longdouble_cross_check_of_gram():
    """float64 oracle vs an np.longdouble evaluation of the same formulas (Gaussian measure)."""
    rng = np.random.default_rng(7)
    D, P, n = 8, 4, 30
    X = rng.standard_normal((n, D))
    ls = rng.uniform(0.5, 3.0, D)
    var = [1.0, 1.0, 0.5, 0.25, 0.125]
    k = oo.OakOracle([oo.RBFDim(l, 1.0, oo.Gaussian(0, 1), expanded=False) for l in ls], P, var)
    K64 = k.K(X)
    Xl = X.astype(np.longdouble)
    mats = []
    for d in range(D):
        l = np.longdouble(ls[d])
        x = Xl[:, d:d + 1]
        base = np.exp(-(x - x.T) ** 2 / (2 * l * l))
        c = l / np.sqrt(l * l + 1) * np.exp(-x * x / (2 * (l * l + 1)))
        v = l / np.sqrt(l * l + 2)
        mats.append(base - (c @ c.T) / v)
    e = oo.esp_dp(mats, P)
    Kl = sum(np.longdouble(s) * t for s, t in zip(var, e))
    assert max_rel_err(K64, np.asarray(Kl, dtype=np.float64)) < 1e-12
