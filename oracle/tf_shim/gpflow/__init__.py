"""NumPy stand-in for the gpflow 2.2.1 names the reference hot path touches (test infrastructure).

Restated from the published gpflow 2.2.1 source: Parameter with the
softplus-positive transform, ``Kernel.__call__/slice``, ``SquaredExponential`` with the expanded
squared distance, ``Kuf``/``Kuu``, ``GPR.log_marginal_likelihood``, ``SGPR.elbo``/``predict_f``.
"""
import sys
import types

import numpy as np
from scipy import linalg as _sla

_JITTER = 1e-6


# ---- config / utilities ---------------------------------------------------------------------
def _eager(x):
    """ndarray answering ``.numpy()`` like the EagerTensor gpflow returns."""
    from tensorflow import EagerArray

    return np.asarray(x).view(EagerArray)


def default_float():
    return np.float64


def default_jitter():
    return _JITTER


class _Identity:
    forward = staticmethod(lambda u: u)
    inverse = staticmethod(lambda v: v)


class _Softplus:
    def __init__(self, lower=0.0):
        self.lower = lower

    def forward(self, u):
        return np.logaddexp(0.0, u) + self.lower

    def inverse(self, v):
        y = np.asarray(v, dtype=np.float64) - self.lower
        return y + np.log(-np.expm1(-y))


def positive(lower=None):
    return _Softplus(0.0 if lower is None else lower)


def to_default_float(x):
    return np.asarray(x, dtype=np.float64)


def print_summary(*a, **k):
    pass


class Parameter:
    __array_priority__ = 1000

    def __init__(self, value, transform=None, prior=None, trainable=True, dtype=None, name=None):
        self.transform = transform if transform is not None else _Identity()
        self.prior, self.trainable = prior, trainable
        if isinstance(value, Parameter):
            value = value.numpy()
        self._u = np.asarray(self.transform.inverse(np.asarray(value, dtype=np.float64)), dtype=np.float64)

    def numpy(self):
        return np.asarray(self.transform.forward(self._u), dtype=np.float64)

    def assign(self, v):
        self._u = np.asarray(self.transform.inverse(np.asarray(v, dtype=np.float64)), dtype=np.float64)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    shape = property(lambda self: self.numpy().shape)

    def __getitem__(self, i):
        return self.numpy()[i]

    def __len__(self):
        return len(self.numpy())

    def _b(op):  # noqa: N805
        def f(self, o):
            return getattr(self.numpy(), op)(np.asarray(o))

        return f

    __mul__ = _b("__mul__")
    __rmul__ = _b("__rmul__")
    __add__ = _b("__add__")
    __radd__ = _b("__radd__")
    __sub__ = _b("__sub__")
    __rsub__ = _b("__rsub__")
    __truediv__ = _b("__truediv__")
    __rtruediv__ = _b("__rtruediv__")
    __pow__ = _b("__pow__")
    __matmul__ = _b("__matmul__")
    __rmatmul__ = _b("__rmatmul__")

    def __neg__(self):
        return -self.numpy()

    def __float__(self):
        return float(np.squeeze(self.numpy()))


def _collect(obj, seen):
    out = []
    if id(obj) in seen:
        return out
    seen.add(id(obj))
    if isinstance(obj, Parameter):
        return [obj]
    if isinstance(obj, (list, tuple)):
        for o in obj:
            out += _collect(o, seen)
    elif isinstance(obj, Module):
        for k in sorted(vars(obj)):
            out += _collect(getattr(obj, k), seen)
    return out


class Module:
    @property
    def parameters(self):
        return tuple(_collect(self, set()))

    @property
    def trainable_parameters(self):
        return tuple(p for p in self.parameters if p.trainable)


def set_trainable(obj, flag):
    for p in _collect(obj, set()):
        p.trainable = flag


# ---- kernels ----------------------------------------------------------------------------------
class Kernel(Module):
    def __init__(self, active_dims=None, name=None):
        self._active_dims = self._norm(active_dims)

    @staticmethod
    def _norm(value):
        if value is None:
            value = slice(None, None, None)
        if not isinstance(value, slice):
            value = np.array(value, dtype=int)
        return value

    @property
    def active_dims(self):
        return self._active_dims

    @active_dims.setter
    def active_dims(self, value):
        self._active_dims = self._norm(value)

    def slice(self, X, X2=None):
        dims = self.active_dims
        X = np.asarray(X)
        X2 = None if X2 is None else np.asarray(X2)
        if isinstance(dims, slice):
            X = X[..., dims]
            if X2 is not None:
                X2 = X2[..., dims]
        else:
            X = np.take(X, dims, axis=-1)
            if X2 is not None:
                X2 = np.take(X2, dims, axis=-1)
        return X, X2

    def __call__(self, X, X2=None, *, full_cov=True, presliced=False):
        if (not full_cov) and (X2 is not None):
            raise ValueError("Ambiguous inputs: `not full_cov` and `X2` are not compatible.")
        if not presliced:
            X, X2 = self.slice(X, X2)
        if not full_cov:
            return self.K_diag(X)
        return self.K(X, X2)

    def __mul__(self, other):
        return Product([self, other])

    def __add__(self, other):
        return Sum([self, other])


class Product(Kernel):
    def __init__(self, kernels):
        super().__init__()
        self.kernels = []
        for k in kernels:
            self.kernels += k.kernels if isinstance(k, Product) else [k]

    def K(self, X, X2=None):
        return np.prod([k(X, X2) for k in self.kernels], axis=0)

    def K_diag(self, X):
        return np.prod([k(X, full_cov=False) for k in self.kernels], axis=0)


class Sum(Kernel):
    def __init__(self, kernels):
        super().__init__()
        self.kernels = []
        for k in kernels:
            self.kernels += k.kernels if isinstance(k, Sum) else [k]

    def K(self, X, X2=None):
        return np.sum([k(X, X2) for k in self.kernels], axis=0)

    def K_diag(self, X):
        return np.sum([k(X, full_cov=False) for k in self.kernels], axis=0)


class Constant(Kernel):
    def __init__(self, variance=1.0, active_dims=None):
        super().__init__(active_dims)
        self.variance = Parameter(variance, transform=positive())

    def K(self, X, X2=None):
        n2 = X.shape[0] if X2 is None else X2.shape[0]
        return np.full((X.shape[0], n2), float(self.variance))

    def K_diag(self, X):
        return np.full(X.shape[0], float(self.variance))


class SquaredExponential(Kernel):
    def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None, name=None):
        super().__init__(active_dims)
        self.variance = Parameter(variance, transform=positive())
        self.lengthscales = Parameter(lengthscales, transform=positive())

    def K(self, X, X2=None):
        l = np.asarray(self.lengthscales)
        Xs = np.asarray(X) / l
        if X2 is None:
            sq = np.sum(np.square(Xs), -1, keepdims=True)
            dist = -2 * (Xs @ Xs.T)
            dist = dist + (sq + sq.T)
        else:
            X2s = np.asarray(X2) / l
            dist = -2 * np.tensordot(Xs, X2s, [[-1], [-1]])
            dist = dist + (np.sum(np.square(Xs), -1)[:, None] + np.sum(np.square(X2s), -1)[None, :])
        return np.asarray(self.variance) * np.exp(-0.5 * dist)

    def K_diag(self, X):
        return np.full(np.shape(X)[:-1], float(np.squeeze(np.asarray(self.variance))))


RBF = SquaredExponential


# ---- inducing variables, likelihood, models -----------------------------------------------------
class InducingPoints(Module):
    def __init__(self, Z):
        self.Z = Parameter(np.asarray(Z, dtype=np.float64))

    def __len__(self):
        return self.Z.numpy().shape[0]

    num_inducing = property(lambda self: len(self))


class _Gaussian(Module):
    def __init__(self, variance=1.0):
        self.variance = Parameter(variance, transform=positive(lower=1e-6))


def Kuf(iv, kernel, X):
    return kernel(iv.Z.numpy(), X)


def Kuu(iv, kernel, jitter=0.0):
    Z = iv.Z.numpy()
    return kernel(Z) + jitter * np.eye(Z.shape[0])


class BayesianModel(Module):
    pass


class GPModel(BayesianModel):
    def __init__(self, data, kernel, mean_function=None):
        self.data = (np.asarray(data[0], dtype=np.float64), np.asarray(data[1], dtype=np.float64))
        self.kernel = kernel
        self.likelihood = _Gaussian()
        self.mean_function = lambda X: np.zeros((np.shape(X)[0], 1))


class GPR(GPModel):
    def log_marginal_likelihood(self):
        X, Y = self.data
        K = self.kernel(X)
        L = np.linalg.cholesky(K + float(self.likelihood.variance) * np.eye(X.shape[0]))
        a = _sla.solve_triangular(L, Y, lower=True)
        n = X.shape[0]
        return float(np.sum(-0.5 * np.sum(a * a, 0) - 0.5 * n * np.log(2 * np.pi) - np.sum(np.log(np.diag(L)))))

    maximum_log_likelihood_objective = log_marginal_likelihood

    def predict_f(self, Xnew):
        X, Y = self.data
        Lm = np.linalg.cholesky(self.kernel(X) + float(self.likelihood.variance) * np.eye(X.shape[0]))
        A = _sla.solve_triangular(Lm, self.kernel(X, Xnew), lower=True)
        mean = A.T @ _sla.solve_triangular(Lm, Y, lower=True)
        var = self.kernel(Xnew, full_cov=False) - np.sum(A * A, 0)
        return _eager(mean), _eager(var[:, None])


class SGPR(GPModel):
    def __init__(self, data, kernel, inducing_variable, mean_function=None):
        super().__init__(data, kernel, mean_function)
        self.inducing_variable = inducing_variable if isinstance(inducing_variable, InducingPoints) else InducingPoints(inducing_variable)

    def _common(self):
        X, Y = self.data
        m = len(self.inducing_variable)
        kuf = Kuf(self.inducing_variable, self.kernel, X)
        kuu = Kuu(self.inducing_variable, self.kernel, jitter=_JITTER)
        s2 = float(self.likelihood.variance)
        sigma = np.sqrt(s2)
        L = np.linalg.cholesky(kuu)
        A = _sla.solve_triangular(L, kuf, lower=True) / sigma
        AAT = A @ A.T
        LB = np.linalg.cholesky(AAT + np.eye(m))
        c = _sla.solve_triangular(LB, A @ Y, lower=True) / sigma
        return L, A, AAT, LB, c, s2

    def elbo(self):
        X, Y = self.data
        n, r = Y.shape
        L, A, AAT, LB, c, s2 = self._common()
        kdiag = self.kernel(X, full_cov=False)
        bound = -0.5 * n * r * np.log(2 * np.pi)
        bound += -r * np.sum(np.log(np.diag(LB)))
        bound -= 0.5 * n * r * np.log(s2)
        bound += -0.5 * np.sum(np.square(Y)) / s2
        bound += 0.5 * np.sum(np.square(c))
        bound += -0.5 * r * np.sum(kdiag) / s2
        bound += 0.5 * r * np.sum(np.diag(AAT))
        return float(bound)

    maximum_log_likelihood_objective = elbo

    def predict_f(self, Xnew):
        L, A, AAT, LB, c, s2 = self._common()
        Kus = Kuf(self.inducing_variable, self.kernel, Xnew)
        tmp1 = _sla.solve_triangular(L, Kus, lower=True)
        tmp2 = _sla.solve_triangular(LB, tmp1, lower=True)
        mean = tmp2.T @ c
        var = self.kernel(Xnew, full_cov=False) + np.sum(tmp2 * tmp2, 0) - np.sum(tmp1 * tmp1, 0)
        return _eager(mean), _eager(var[:, None])


class Bernoulli(Module):
    """gpflow 2.2.1 ``likelihoods.Bernoulli`` (restated): ScalarLikelihood with 20-point Gauss-Hermite quadrature
    (``NDiagGHQuadrature``: nodes of ``np.polynomial.hermite.hermgauss`` times sqrt 2, weights over sqrt pi)."""

    num_gauss_hermite_points = 20

    def __init__(self, invlink=None):
        from scipy.special import ndtr

        self.invlink = invlink if invlink is not None else (lambda f: ndtr(f) * (1 - 2e-3) + 1e-3)  # inv_probit

    def _log_prob(self, F, Y):
        p = np.asarray(self.invlink(F), dtype=np.float64)
        return np.log(np.where(Y == 1, p, 1 - p))      # logdensities.bernoulli

    def _quad_nodes(self, Fmu, Fvar):
        x, w = np.polynomial.hermite.hermgauss(self.num_gauss_hermite_points)
        F = Fmu[..., None] + np.sqrt(Fvar)[..., None] * (x * np.sqrt(2.0))
        return F, w / np.sqrt(np.pi)

    def variational_expectations(self, Fmu, Fvar, Y):
        F, w = self._quad_nodes(np.asarray(Fmu)[:, 0], np.asarray(Fvar)[:, 0])
        return np.sum(self._log_prob(F, np.asarray(Y)[:, :1]) * w, axis=-1)

    def predict_log_density(self, Fmu, Fvar, Y):
        from scipy.special import logsumexp

        F, w = self._quad_nodes(np.asarray(Fmu)[:, 0], np.asarray(Fvar)[:, 0])
        return logsumexp(self._log_prob(F, np.asarray(Y)[:, :1]) + np.log(w), axis=-1)


class _Posterior:
    """gpflow 2.2.1 ``posteriors.IndependentPosteriorSingleOutput._precompute`` (restated) for whiten=True and a
    diagonal q: alpha = L^-T q_mu, Qinv = L^-T (I - diag(q_sqrt^2)) L^-1 with L = chol(Kuu + jitter I)."""

    def __init__(self, model):
        Kmm = Kuu(model.inducing_variable, model.kernel, jitter=default_jitter())
        L = np.linalg.cholesky(Kmm)
        q_mu, q_sqrt = model.q_mu.numpy(), model.q_sqrt.numpy()
        self.alpha = _eager(_sla.solve_triangular(L, q_mu, lower=True, trans="T"))
        B = np.eye(L.shape[0]) - np.diag(q_sqrt[:, 0] ** 2)
        LinvT_B = _sla.solve_triangular(L, B, lower=True, trans="T")
        self.Qinv = _eager(_sla.solve_triangular(L, LinvT_B.T, lower=True, trans="T")[None])


class SVGP(GPModel):
    """gpflow 2.2.1 ``models.SVGP`` (restated) in the one configuration the reference builds
    (examples/uci/uci_classification_train.py:108-116): whiten=True, q_diag=True, one latent GP, zero mean."""

    def __init__(self, kernel, likelihood, inducing_variable, *, whiten=True, q_diag=False, q_mu=None, q_sqrt=None,
                 num_data=None, mean_function=None, num_latent_gps=1):
        assert whiten and q_diag and num_latent_gps == 1, "the shim restates the reference's configuration only"
        self.kernel, self.likelihood, self.num_data = kernel, likelihood, num_data
        self.inducing_variable = inducing_variable if isinstance(inducing_variable, InducingPoints) else \
            InducingPoints(inducing_variable)
        m = len(self.inducing_variable)
        self.q_mu = Parameter(np.zeros((m, 1)) if q_mu is None else q_mu)
        self.q_sqrt = Parameter(np.ones((m, 1)) if q_sqrt is None else q_sqrt, transform=positive())
        self.whiten, self.q_diag = whiten, q_diag
        self.mean_function = lambda X: np.zeros((np.shape(X)[0], 1))

    def posterior(self):
        return _Posterior(self)

    def prior_kl(self):
        # kullback_leiblers.gauss_kl(q_mu, q_sqrt, K=None) with a diagonal q_sqrt
        q_mu, q_sqrt = self.q_mu.numpy(), self.q_sqrt.numpy()
        two_kl = np.sum(q_mu ** 2) - q_sqrt.size - np.sum(np.log(q_sqrt ** 2)) + np.sum(q_sqrt ** 2)
        return 0.5 * two_kl

    def predict_f(self, Xnew, full_cov=False):
        # conditionals.util.base_conditional(Kmn, Kmm, Knn, f=q_mu, q_sqrt=diag, white=True)
        Kmm = Kuu(self.inducing_variable, self.kernel, jitter=default_jitter())
        Kmn = Kuf(self.inducing_variable, self.kernel, Xnew)
        Knn = np.asarray(self.kernel(Xnew, full_cov=False))
        Lm = np.linalg.cholesky(Kmm)
        A = _sla.solve_triangular(Lm, Kmn, lower=True)
        fvar = Knn - np.sum(A * A, 0)
        fmean = A.T @ self.q_mu.numpy()
        LTA = A * self.q_sqrt.numpy()
        fvar = fvar + np.sum(LTA * LTA, 0)
        return _eager(fmean), _eager(fvar[:, None])

    def elbo(self, data):
        X, Y = data
        kl = self.prior_kl()
        f_mean, f_var = self.predict_f(X)
        var_exp = self.likelihood.variational_expectations(f_mean, f_var, Y)
        scale = 1.0 if self.num_data is None else self.num_data / np.shape(X)[0]
        return float(np.sum(var_exp) * scale - kl)

    maximum_log_likelihood_objective = elbo

    def predict_log_density(self, data):
        X, Y = data
        f_mean, f_var = self.predict_f(X)
        return _eager(self.likelihood.predict_log_density(f_mean, f_var, Y))


class _Scipy:
    def minimize(self, *a, **k):
        raise NotImplementedError("optimisation is outside the golden-vector scope")


# ---- module tree ----------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_base_k = _mod("gpflow.kernels.base", Product=Product, Sum=Sum, Kernel=Kernel)
_statics = _mod("gpflow.kernels.statics", Constant=Constant)
kernels = _mod("gpflow.kernels", Kernel=Kernel, RBF=RBF, SquaredExponential=SquaredExponential, Constant=Constant,
               Product=Product, Sum=Sum, base=_base_k, statics=_statics)
utilities = _mod("gpflow.utilities", positive=positive, to_default_float=to_default_float,
                 print_summary=print_summary, set_trainable=set_trainable)
config = _mod("gpflow.config", default_float=default_float, default_jitter=default_jitter)
class _BaseModule(Module):
    def __init__(self, name=None):
        pass


base = _mod("gpflow.base", Parameter=Parameter, Module=_BaseModule)
inducing_variables = _mod("gpflow.inducing_variables", InducingPoints=InducingPoints)
_tm = _mod("gpflow.models.training_mixins", RegressionData=tuple)
models = _mod("gpflow.models", GPR=GPR, SGPR=SGPR, SVGP=SVGP, GPModel=GPModel, BayesianModel=BayesianModel,
              training_mixins=_tm)
_disp = _mod("gpflow.covariances.dispatch", Kuf=Kuf, Kuu=Kuu)
covariances = _mod("gpflow.covariances", dispatch=_disp, Kuf=Kuf, Kuu=Kuu)
optimizers = _mod("gpflow.optimizers", Scipy=_Scipy)
likelihoods = _mod("gpflow.likelihoods", Bernoulli=Bernoulli, Gaussian=_Gaussian)
