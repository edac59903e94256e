"""TEST INFRASTRUCTURE ONLY -- NumPy FP64 restatement of the reference hot path.

PARITY PIN STATUS: the reference (gpflow 2.2.1 / TensorFlow 2.11 / TFP 0.11, none of
them vendored or installable here) cannot execute in this container, and its test-suite
holds no golden Gram vectors.  This restatement is pinned instead by
  (1) ``tests/golden/*.npz`` -- outputs of the *reference's own Python sources*
      (``/root/reference/oak/*.py``, unmodified) executed over the NumPy stand-in for the
      TF/gpflow ops they call (``oracle/tf_shim``; script ``tests/golden/make_golden.py``),
  (2) every identity / known-answer test the reference ships for this path
      (SURVEY.md section 4, rows marked with a star), ported in ``tests/test_oracle_*.py``,
  (3) quadrature checks of the orthogonality constraint and of the Sobol ``L`` integrals.
The gpflow pieces (RBF, Kuf/Kuu, GPR.log_marginal_likelihood, SGPR.elbo) are restated from
the published gpflow 2.2.1 algorithm and anchored on the reference's call sites.

Every function cites the reference file:line it follows (paths relative to
``/root/reference``).  Operation order follows the reference (one dense matrix per input
dimension, ``pow``-based power sums, Newton-Girard on whole matrices) so that this file is
also the honest CPU comparator for ``bench.py --impl reference``.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
from scipy.special import erf as _erf

JITTER = 1e-6  # gpflow default_jitter(), used at oak/utils.py:185


# --------------------------------------------------------------------------------------
# measures  (oak/input_measures.py:16-78)
# --------------------------------------------------------------------------------------
@dataclass
class Gaussian:
    mu: float = 0.0
    var: float = 1.0


@dataclass
class Uniform:
    a: float = 0.0
    b: float = 1.0


@dataclass
class Empirical:
    location: np.ndarray  # (M,1)
    weights: Optional[np.ndarray] = None  # (M,1); default 1/M (input_measures.py:51-52)

    def __post_init__(self):
        self.location = np.asarray(self.location, dtype=np.float64).reshape(-1, 1)
        if self.weights is None:
            self.weights = np.full((self.location.shape[0], 1), 1.0 / self.location.shape[0])
        self.weights = np.asarray(self.weights, dtype=np.float64).reshape(-1, 1)
        assert np.isclose(self.weights.sum(), 1.0, atol=1e-6)  # input_measures.py:53-55


@dataclass
class MOG:
    means: np.ndarray
    variances: np.ndarray
    weights: np.ndarray

    def __post_init__(self):
        self.means = np.asarray(self.means).astype(float).reshape(-1)  # :75-76
        self.variances = np.asarray(self.variances).astype(float).reshape(-1)
        self.weights = np.asarray(self.weights, dtype=np.float64).reshape(-1)
        assert np.isclose(self.weights.sum(), 1.0, atol=1e-6)


# --------------------------------------------------------------------------------------
# gpflow SquaredExponential (un-vendored; call sites oak/ortho_rbf_kernel.py:107,116,169,176)
# --------------------------------------------------------------------------------------
def rbf_K(x, x2, lengthscale, variance, expanded=True):
    """gpflow 2.2.1 ``SquaredExponential.K``: ``variance * exp(-0.5 * square_distance(X/l, X2/l))``
    with the *expanded* squared distance ``|x|^2 + |y|^2 - 2 x.y``.
    ``expanded=False`` uses the direct ``(x-y)^2`` form (quirk 2 of SURVEY.md section 2.2)."""
    xs = x / lengthscale
    ys = xs if x2 is None else x2 / lengthscale
    if expanded:
        d = -2.0 * (xs @ ys.T)
        d = d + (np.sum(xs * xs, -1)[:, None] + np.sum(ys * ys, -1)[None, :])
    else:
        d = (xs[:, None, :] - ys[None, :, :]) ** 2
        d = d.sum(-1)
    return variance * np.exp(-0.5 * d)


# --------------------------------------------------------------------------------------
# per-dimension kernels
# --------------------------------------------------------------------------------------
@dataclass
class RBFDim:
    """``OrthogonalRBFKernel`` (oak/ortho_rbf_kernel.py:20-177); ``measure=None`` is the plain,
    unconstrained gpflow RBF used when ``constrain_orthogonal=False`` (oak/oak_kernel.py:191-210)."""

    lengthscale: float = 1.0
    variance: float = 1.0
    measure: object = field(default_factory=Gaussian)
    expanded: bool = True

    # ortho_rbf_kernel.py:47-152
    def cov_X_s(self, x):
        l, s2, m = self.lengthscale, self.variance, self.measure
        if isinstance(m, Uniform):  # :49-63
            return (
                s2 * l / (m.b - m.a) * np.sqrt(np.pi / 2)
                * (_erf((m.b - x) / np.sqrt(2) / l) - _erf((m.a - x) / np.sqrt(2) / l))
            )
        if isinstance(m, Gaussian):  # :82-92
            return s2 * l / np.sqrt(l ** 2 + m.var) * np.exp(-0.5 * ((x - m.mu) ** 2) / (l ** 2 + m.var))
        if isinstance(m, Empirical):  # :101-107
            return rbf_K(x, m.location, l, s2, self.expanded) @ m.weights
        if isinstance(m, MOG):  # :124-136
            tmp = np.exp(-0.5 * ((x - m.means) ** 2) / (l ** 2 + m.variances)) / np.sqrt(l ** 2 + m.variances)
            return s2 * l * (tmp @ m.weights.reshape(-1, 1))
        raise NotImplementedError  # :36-45

    def var_s(self):
        l, s2, m = self.lengthscale, self.variance, self.measure
        if isinstance(m, Uniform):  # :65-78
            y = (m.b - m.a) / np.sqrt(2) / l
            return 2.0 / ((m.b - m.a) ** 2) * s2 * l ** 2 * (np.sqrt(np.pi) * y * _erf(y) + np.exp(-np.square(y)) - 1.0)
        if isinstance(m, Gaussian):  # :94-97
            return s2 * l / np.sqrt(l ** 2 + 2 * m.var)
        if isinstance(m, Empirical):  # :109-120
            return float(np.squeeze(m.weights.T @ rbf_K(m.location, None, l, s2, self.expanded) @ m.weights))
        if isinstance(m, MOG):  # :138-152
            dists = np.square(m.means[:, None] - m.means[None, :])
            scales = np.square(l) + m.variances[:, None] + m.variances[None, :]
            tmp = s2 * l / np.sqrt(scales) * np.exp(-0.5 * dists / scales)
            return float(np.squeeze(m.weights[None, :] @ tmp @ m.weights[:, None]))
        raise NotImplementedError

    def K(self, x, x2=None):  # :157-172
        base = rbf_K(x, x2, self.lengthscale, self.variance, self.expanded)
        if self.measure is None:
            return base
        c = self.cov_X_s(x)
        c2 = c if x2 is None else self.cov_X_s(x2)
        return base - (c @ c2.T) / self.var_s()

    def K_diag(self, x):  # :174-177
        base = np.full(x.shape[0], float(np.squeeze(self.variance)))
        if self.measure is None:
            return base
        return base - np.square(self.cov_X_s(x)[:, 0]) / self.var_s()


def _as_index(x):
    """``tf.cast(X[..., 0], tf.int32)``: truncation toward zero (ortho_binary_kernel.py:47,51)."""
    return np.trunc(np.asarray(x)[..., 0]).astype(np.int32)


@dataclass
class BinaryDim:
    """``OrthogonalBinary`` (oak/ortho_binary_kernel.py:13-59)."""

    p0: float = 0.5
    variance: float = 1.0

    def table(self):  # :29-33
        p0, p1 = self.p0, 1.0 - self.p0
        return np.array([[p1 * p1, -p0 * p1], [-p0 * p1, p0 * p0]]) * self.variance

    def table_diag(self):  # :35-38
        p0, p1 = self.p0, 1.0 - self.p0
        return np.array([p1 * p1, p0 * p0]) * self.variance

    def K(self, x, x2=None):  # :40-53
        i = _as_index(x)
        j = i if x2 is None else _as_index(x2)
        return self.table()[j].T[i]

    def K_diag(self, x):  # :55-59
        return self.table_diag()[_as_index(x)]


@dataclass
class CategoricalDim:
    """``OrthogonalCategorical`` (oak/ortho_categorical_kernel.py:14-74)."""

    p: np.ndarray = None  # (C,1)
    W: np.ndarray = None  # (C,rank)
    kappa: np.ndarray = None  # (C,)
    variance: float = 1.0

    def __post_init__(self):
        self.p = np.asarray(self.p, dtype=np.float64).reshape(-1, 1)
        self.W = np.asarray(self.W, dtype=np.float64)
        self.kappa = np.asarray(self.kappa, dtype=np.float64).reshape(-1)

    def table(self):  # :34-42
        A = self.W @ self.W.T + np.diag(self.kappa)
        Ap = A @ self.p
        return (A - (Ap @ Ap.T) / (self.p.T @ Ap)[0]) * self.variance

    def table_diag(self):  # :44-53
        A = self.W @ self.W.T + np.diag(self.kappa)
        Ap = A @ self.p
        A_diag = np.sum(np.square(self.W), 1) + self.kappa
        return (A_diag - np.sum(np.square(Ap), 1) / (self.p.T @ Ap)[0]) * self.variance

    def K(self, x, x2=None):  # :55-68
        i = _as_index(x)
        j = i if x2 is None else _as_index(x2)
        return self.table()[j].T[i]

    def K_diag(self, x):  # :70-74
        return self.table_diag()[_as_index(x)]


# --------------------------------------------------------------------------------------
# composite kernel (oak/oak_kernel.py)
# --------------------------------------------------------------------------------------
def newton_girard(mats: Sequence[np.ndarray], depth: int) -> List[np.ndarray]:
    """``OAKKernel.compute_additive_terms`` (oak/oak_kernel.py:223-249): power sums
    ``s_p = sum_d k_d**p`` (p = 0..depth) then ``e_n = (1/n) sum_k (-1)^(k-1) e_{n-k} s_k``."""
    s = []
    for p in range(depth + 1):
        acc = np.power(mats[0], p)
        for k in mats[1:]:
            acc = acc + np.power(k, p)
        s.append(acc)
    e = [np.ones_like(mats[0])]
    for n in range(1, depth + 1):
        acc = None
        for k in range(1, n + 1):
            term = ((-1) ** (k - 1)) * e[n - k] * s[k]
            acc = term if acc is None else acc + term
        e.append((1.0 / n) * acc)
    return e


def esp_dp(mats: Sequence[np.ndarray], depth: int) -> List[np.ndarray]:
    """Cross-check: the direct recurrence ``e_n += k_d * e_{n-1}`` (not in the reference)."""
    e = [np.ones_like(mats[0])] + [np.zeros_like(mats[0]) for _ in range(depth)]
    for k in mats:
        for n in range(depth, 0, -1):
            e[n] = e[n] + k * e[n - 1]
    return e


def esp_bruteforce(mats: Sequence[np.ndarray], depth: int) -> List[np.ndarray]:
    """Brute force over ``itertools.combinations`` (tests/test_kernel_properties.py:80-83)."""
    out = [np.ones_like(mats[0])]
    for n in range(1, depth + 1):
        acc = np.zeros_like(mats[0])
        for comb in itertools.combinations(mats, n):
            acc = acc + np.prod(comb, axis=0)
        out.append(acc)
    return out


def subsets(num_dims: int, depth: int) -> List[List[int]]:
    """``get_list_representation`` ordering (oak/oak_kernel.py:338-364): ``[]`` then all
    combinations of each order 1..depth in lexicographic order."""
    out: List[List[int]] = [[]]
    if depth > 0:
        for n in range(1, depth + 1):
            out += [list(t) for t in itertools.combinations(range(num_dims), n)]
    return out


@dataclass
class OakOracle:
    """``OAKKernel`` (oak/oak_kernel.py:36-278). ``dims[d]`` acts on column ``active_dims[d]``."""

    dims: list
    depth: int
    variances: Sequence[float]  # sigma^2_0..depth (share_var) or [sigma^2_0]
    share_var_across_orders: bool = True
    active_dims: Optional[List[int]] = None

    def _col(self, X, d):
        j = d if self.active_dims is None else self.active_dims[d]
        return None if X is None else np.asarray(X, dtype=np.float64)[:, j : j + 1]

    def dim_matrices(self, X, X2=None):  # oak_kernel.py:252-254
        return [k.K(self._col(X, d), self._col(X2, d)) for d, k in enumerate(self.dims)]

    def _combine(self, e):  # oak_kernel.py:255-265
        if self.share_var_across_orders:
            out = self.variances[0] * e[0]
            for s2, t in zip(self.variances[1:], e[1:]):
                out = out + s2 * t
            return out
        out = self.variances[0] * e[0]
        for t in e[1:]:
            out = out + t
        return out

    def K(self, X, X2=None):  # oak_kernel.py:251-265
        return self._combine(newton_girard(self.dim_matrices(X, X2), self.depth))

    def K_diag(self, X):  # oak_kernel.py:267-278
        diags = [k.K_diag(self._col(X, d)) for d, k in enumerate(self.dims)]
        return self._combine(newton_girard(diags, self.depth))

    # KernelComponenent (oak/oak_kernel.py:281-335)
    def component_K(self, subset, X, X2=None):
        n = np.asarray(X).shape[0]
        n2 = n if X2 is None else np.asarray(X2).shape[0]
        if len(subset) == 0:
            return self.variances[0] * np.ones((n, n2))
        var_n = self.variances[len(subset)] if self.share_var_across_orders else 1.0
        mats = [self.dims[d].K(self._col(X, d), self._col(X2, d)) for d in sorted(subset)]
        return var_n * np.prod(mats, axis=0)

    def component_K_diag(self, subset, X):
        n = np.asarray(X).shape[0]
        if len(subset) == 0:
            return self.variances[0] * np.ones(n)
        var_n = self.variances[len(subset)] if self.share_var_across_orders else 1.0
        mats = [self.dims[d].K_diag(self._col(X, d)) for d in sorted(subset)]
        return var_n * np.prod(mats, axis=0)


# --------------------------------------------------------------------------------------
# gpflow objectives (un-vendored) and OAK's sufficient statistics
# --------------------------------------------------------------------------------------
def _chol(a):
    return np.linalg.cholesky(a)


def _trsm_lower(L, b):
    import scipy.linalg as sla

    return sla.solve_triangular(L, b, lower=True)


def gpr_log_marginal_likelihood(kern: OakOracle, X, Y, noise):
    """gpflow 2.2.1 ``GPR.log_marginal_likelihood`` (model built at oak/model_utils.py:159)."""
    n = X.shape[0]
    L = _chol(kern.K(X) + noise * np.eye(n))
    a = _trsm_lower(L, Y)
    per_col = -0.5 * np.sum(a * a, 0) - 0.5 * n * np.log(2 * np.pi) - np.sum(np.log(np.diag(L)))
    return float(np.sum(per_col))


def sgpr_pieces(kern: OakOracle, X, Y, Z, noise, jitter=JITTER):
    """The intermediate matrices shared by gpflow 2.2.1 ``SGPR.elbo`` and
    ``get_model_sufficient_statistics`` (oak/utils.py:180-204)."""
    m = Z.shape[0]
    kuf = kern.K(Z, X)  # Kuf(iv, kernel, X)   utils.py:184
    kuu = kern.K(Z) + jitter * np.eye(m)  # Kuu(iv, kernel, jitter)  utils.py:185
    sigma = np.sqrt(noise)
    L = _chol(kuu)
    A = _trsm_lower(L, kuf) / sigma
    AAT = A @ A.T
    B = AAT + np.eye(m)
    LB = _chol(B)
    Aerr = A @ Y
    c = _trsm_lower(LB, Aerr) / sigma
    return dict(kuf=kuf, kuu=kuu, L=L, A=A, AAT=AAT, B=B, LB=LB, Aerr=Aerr, c=c)


def sgpr_elbo(kern: OakOracle, X, Y, Z, noise, jitter=JITTER):
    """gpflow 2.2.1 ``SGPR.elbo`` (Titsias bound)."""
    n, r = Y.shape
    q = sgpr_pieces(kern, X, Y, Z, noise, jitter)
    kdiag = kern.K_diag(X)
    bound = -0.5 * n * r * np.log(2 * np.pi)
    bound += -r * np.sum(np.log(np.diag(q["LB"])))
    bound -= 0.5 * n * r * np.log(noise)
    bound += -0.5 * np.sum(np.square(Y)) / noise
    bound += 0.5 * np.sum(np.square(q["c"]))
    bound += -0.5 * r * np.sum(kdiag) / noise
    bound += 0.5 * r * np.sum(np.diag(q["AAT"]))
    return float(bound)


def sgpr_alpha(kern: OakOracle, X, Y, Z, noise, jitter=JITTER):
    """``get_model_sufficient_statistics`` SGPR branch (oak/utils.py:180-198)."""
    q = sgpr_pieces(kern, X, Y, Z, noise, jitter)
    tmp1 = np.linalg.solve(q["LB"].T, q["c"])
    return np.linalg.solve(q["L"].T, tmp1)


def gpr_alpha(kern: OakOracle, X, Y, noise):
    """``get_model_sufficient_statistics`` GPR branch (oak/utils.py:206-211)."""
    import scipy.linalg as sla

    Kt = kern.K(X) + np.eye(X.shape[0]) * noise
    L = _chol(Kt)
    return sla.cho_solve((L, True), Y)


def gpr_predict_mean(kern: OakOracle, X, Y, noise, Xnew):
    """Mean of gpflow ``GPR.predict_f`` (base_conditional)."""
    Lm = _chol(kern.K(X) + noise * np.eye(X.shape[0]))
    A = _trsm_lower(Lm, kern.K(X, Xnew))
    return A.T @ _trsm_lower(Lm, Y)


def sgpr_predict_mean(kern: OakOracle, X, Y, Z, noise, Xnew, jitter=JITTER):
    """Mean of gpflow 2.2.1 ``SGPR.predict_f``."""
    q = sgpr_pieces(kern, X, Y, Z, noise, jitter)
    tmp1 = _trsm_lower(q["L"], kern.K(Z, Xnew))
    tmp2 = _trsm_lower(q["LB"], tmp1)
    return tmp2.T @ q["c"]


# --------------------------------------------------------------------------------------
# Sobol indices (oak/utils.py:116-165, 221-435)
# --------------------------------------------------------------------------------------
# ---- SVGP + Bernoulli (the classification run, examples/uci/uci_classification_train.py:108-135) ------
# gpflow 2.2.1 is not vendored: restated from models/svgp.py (elbo, prior_kl), kullback_leiblers.py (gauss_kl with
# K=None and a diagonal q_sqrt), conditionals/util.py (base_conditional, white=True), likelihoods (Bernoulli,
# NDiagGHQuadrature with 20 points), posteriors.py (alpha = L^-T q_mu, Qinv = L^-T (I - diag q_sqrt^2) L^-1).
# Pinned by tests/golden/g11_svgp_classification.npz: the reference's own get_model_sufficient_statistics SVGP
# branch (oak/utils.py:174-179), compute_sobol_oak and get_prediction_component run over oracle/tf_shim.
def inv_logit(f, jitter=1e-3):  # uci_classification_train.py:43-45
    return 1.0 / (1.0 + np.exp(-np.asarray(f, dtype=np.float64))) * (1 - 2 * jitter) + jitter


def svgp_alpha(kern: OakOracle, Z, q_mu, jitter=JITTER):
    """``posterior.alpha`` of the SVGP branch of ``get_model_sufficient_statistics`` (oak/utils.py:174-177)."""
    import scipy.linalg as sla

    L = _chol(kern.K(Z) + jitter * np.eye(Z.shape[0]))
    return sla.solve_triangular(L, np.asarray(q_mu).reshape(-1, 1), lower=True, trans="T")


def svgp_L(kern: OakOracle, Z, q_sqrt, jitter=JITTER):
    """``cholesky(inv(posterior.Qinv[0]))`` (oak/utils.py:178-179)."""
    import scipy.linalg as sla

    L = _chol(kern.K(Z) + jitter * np.eye(Z.shape[0]))
    B = np.eye(Z.shape[0]) - np.diag(np.asarray(q_sqrt).reshape(-1) ** 2)
    LinvT_B = sla.solve_triangular(L, B, lower=True, trans="T")
    Qinv = sla.solve_triangular(L, LinvT_B.T, lower=True, trans="T")
    return _chol(np.linalg.inv(Qinv))


def svgp_predict_f(kern: OakOracle, Z, q_mu, q_sqrt, Xnew, jitter=JITTER):
    """Marginal mean and variance of the whitened conditional with a diagonal q(u)."""
    A = _trsm_lower(_chol(kern.K(Z) + jitter * np.eye(Z.shape[0])), kern.K(Z, Xnew))
    fvar = kern.K_diag(Xnew) - np.sum(A * A, 0)
    fmean = (A.T @ np.asarray(q_mu).reshape(-1, 1))[:, 0]
    LTA = A * np.asarray(q_sqrt).reshape(-1, 1)
    return fmean, fvar + np.sum(LTA * LTA, 0)


def _bernoulli_nodes(fmean, fvar, Y, invlink, n_gh):
    x, w = np.polynomial.hermite.hermgauss(n_gh)
    F = fmean[:, None] + np.sqrt(fvar)[:, None] * (x * np.sqrt(2.0))[None, :]
    p = invlink(F)
    return np.log(np.where(np.asarray(Y).reshape(-1, 1) == 1, p, 1 - p)), w / np.sqrt(np.pi)


def svgp_elbo(kern: OakOracle, X, Y, Z, q_mu, q_sqrt, invlink=inv_logit, num_data=None, n_gh=20, jitter=JITTER):
    fmean, fvar = svgp_predict_f(kern, Z, q_mu, q_sqrt, X, jitter)
    logp, w = _bernoulli_nodes(fmean, fvar, Y, invlink, n_gh)
    q_mu, q_sqrt = np.asarray(q_mu).reshape(-1), np.asarray(q_sqrt).reshape(-1)
    two_kl = np.sum(q_mu ** 2) - q_mu.size - np.sum(np.log(q_sqrt ** 2)) + np.sum(q_sqrt ** 2)
    scale = 1.0 if num_data is None else num_data / X.shape[0]
    return float(np.sum(logp * w) * scale - 0.5 * two_kl)


def svgp_predict_log_density(kern: OakOracle, Z, q_mu, q_sqrt, Xnew, Ynew, invlink=inv_logit, n_gh=20, jitter=JITTER):
    from scipy.special import logsumexp

    fmean, fvar = svgp_predict_f(kern, Z, q_mu, q_sqrt, Xnew, jitter)
    logp, w = _bernoulli_nodes(fmean, fvar, Ynew, invlink, n_gh)
    return logsumexp(logp + np.log(w)[None, :], axis=1)


def f1(x, y, sigma, l, delta, mu):  # utils.py:116-125
    return (
        sigma ** 4 * l / np.sqrt(l ** 2 + 2 * delta ** 2)
        * np.exp(-((x - y) ** 2) / (4 * l ** 2))
        * np.exp(-((mu - (x + y) / 2) ** 2) / (2 * delta ** 2 + l ** 2))
    )


def f2(x, y, sigma, l, delta, mu):  # utils.py:128-146
    M = 1 / (l ** 2) + 1 / (l ** 2 + delta ** 2)
    m = 1 / M * (mu / (l ** 2 + delta ** 2) + x / l ** 2)
    C = x ** 2 / (l ** 2) + mu ** 2 / (l ** 2 + delta ** 2) - m ** 2 * M
    return (
        sigma ** 4 * l * np.sqrt((l ** 2 + 2 * delta ** 2) / (delta ** 2 * M + 1))
        * np.exp(-C / 2) / (l ** 2 + delta ** 2)
        * np.exp(-((y - mu) ** 2) / (2 * (l ** 2 + delta ** 2)))
        * np.exp(-((m - mu) ** 2) / (2 * (1 / M + delta ** 2)))
    )


def f3(x, y, sigma, l, delta, mu):  # utils.py:149-151
    return f2(y, x, sigma, l, delta, mu)


def f4(x, y, sigma, l, delta, mu):  # utils.py:154-165
    return (
        sigma ** 4 * l ** 2 * (l ** 2 + 2 * delta ** 2)
        * np.sqrt((l ** 2 + delta ** 2) / (l ** 2 + 3 * delta ** 2))
        / ((l ** 2 + delta ** 2) ** 2)
        * np.exp(-((x - mu) ** 2 + (y - mu) ** 2) / (2 * (l ** 2 + delta ** 2)))
    )


def L_gaussian(xcol, l, variance, delta, mu):
    """``compute_L`` (oak/utils.py:221-240)."""
    n = xcol.shape[0]
    sigma = np.sqrt(variance)
    x = np.repeat(xcol, n)
    y = np.tile(xcol, n)
    L = f1(x, y, sigma, l, delta, mu) - f2(x, y, sigma, l, delta, mu) - f3(x, y, sigma, l, delta, mu) + f4(x, y, sigma, l, delta, mu)
    return L.reshape(n, n)


def L_binary(xcol, p0, variance):
    """``compute_L_binary_kernel`` (oak/utils.py:243-272) -- note the variance**1 scaling
    (quirk 1), replicated as-is."""
    assert 0 <= p0 <= 1
    n = xcol.shape[0]
    x = np.repeat(xcol, n)
    y = np.tile(xcol, n)
    p1 = 1 - p0
    L = variance * (
        p0 * (p1 ** 2 * (1 - x) - p0 * p1 * x) * (p1 ** 2 * (1 - y) - p0 * p1 * y)
        + p1 * (-p0 * p1 * (1 - x) + p0 ** 2 * x) * (-p0 * p1 * (1 - y) + p0 ** 2 * y)
    )
    return L.reshape(n, n)


def L_categorical(xcol, W, kappa, p, variance):
    """``compute_L_categorical_kernel`` (oak/utils.py:275-309)."""
    p = np.asarray(p, dtype=np.float64).reshape(-1, 1)
    assert np.abs(p.sum() - 1) < 1e-6
    B = CategoricalDim(p=p, W=W, kappa=kappa, variance=1.0).table() * variance
    idx = np.trunc(xcol).astype(np.int32)
    Kc = B[idx].T[np.arange(len(p))]  # (C, N): Kc[c, i] = B[x_i, c]
    return Kc.T @ (Kc * p)


def L_empirical(location, weights, kern: RBFDim, zcol):
    """``compute_L_empirical_measure`` (oak/utils.py:312-335)."""
    kxu = kern.K(np.asarray(location).reshape(-1, 1), np.asarray(zcol).reshape(-1, 1))
    w = np.asarray(weights).reshape(1, -1)
    return (w * kxu.T) @ kxu


def sobol_oak(kern: OakOracle, Xcond, alpha, delta=1.0, mu=0.0, share_var_across_orders=True, only=None):
    """``compute_sobol_oak`` (oak/utils.py:338-435). ``Xcond`` is Z for SGPR/SVGP and the
    training inputs for GPR (:361-364). Returns (list of subsets without the constant, values).
    ``only`` (test aid): indices into that list -- evaluate just those components (the reference
    rebuilds every L matrix per subset, which takes minutes at D=50 or depth 8)."""
    Xcond = np.asarray(Xcond, dtype=np.float64)
    n = Xcond.shape[0]
    comps = subsets(len(kern.dims), kern.depth)[1:]
    if only is not None:
        comps = [comps[i] for i in only]
    out = []
    for S in comps:
        L = np.ones((n, n))
        order = len(S)
        for j, d in enumerate(S):
            k = kern.dims[d]
            if share_var_across_orders:  # :376-380
                v = kern.variances[order] if j < 1 else 1
            else:
                v = k.variance if isinstance(k, RBFDim) else None  # :382 (quirk 5)
                if v is None:
                    raise AttributeError("base_kernel")
            col = kern._col(Xcond, d)[:, 0]
            if isinstance(k, RBFDim):
                if isinstance(k.measure, Empirical):  # :402-412
                    L = v ** 2 * L * L_empirical(k.measure.location, k.measure.weights, k, col)
                elif isinstance(k.measure, MOG):  # :413-414
                    raise NotImplementedError
                else:  # :386-400
                    L = L * L_gaussian(col, k.lengthscale, v, delta, mu)
            elif isinstance(k, BinaryDim):  # :416-418
                L = L * L_binary(col, k.p0, v)
            elif isinstance(k, CategoricalDim):  # :420-424
                L = L * L_categorical(col, k.W, k.kappa, k.p, v)
            else:
                raise NotImplementedError
        out.append(float((alpha.T @ L @ alpha)[0, 0]))  # :429-432
    return comps, out


def predict_components(kern: OakOracle, Xcond, alpha, X, share_var_across_orders=True):
    """``get_prediction_component`` (oak/utils.py:491-530)."""
    comps = subsets(len(kern.dims), kern.depth)[1:]
    out = []
    for S in comps:
        Kxx = np.ones((X.shape[0], alpha.shape[0]))
        for d in S:
            Kxx = Kxx * kern.dims[d].K(kern._col(X, d), kern._col(Xcond, d))
        if share_var_across_orders:
            Kxx = Kxx * kern.variances[len(S)]
        out.append((Kxx @ alpha)[:, 0])
    return out
