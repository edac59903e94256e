"""GPR / SGPR objectives on top of the fused OAK tiles.

Stand-ins for the gpflow 2.2.1 model classes the reference instantiates
(``oak/model_utils.py:149-159``): same attribute paths (``.data``, ``.kernel``,
``.likelihood.variance``, ``.inducing_variable.Z``) and methods (``log_marginal_likelihood``,
``elbo``, ``maximum_log_likelihood_objective``, ``training_loss``, ``predict_f``).  The Gram /
Kuf tiles, the SGPR statistics and the M^3 tail (a one-launch bordered Cholesky) run in ``liboak_b200.so``.
With ``distributed=True`` each rank holds a shard of (X, y) and the packed statistics are combined by one
all-reduce.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import _device, parallel
from ._gpflow_shim import (DEFAULT_JITTER, Gaussian, InducingPoints, Module, Parameter, collect_parameters,
                           scalar_of, value_of)


class GPModel(Module):
    def __init__(self, data, kernel, mean_function=None, noise_variance: float = 1.0):
        X, Y = data
        self.data = (X, Y)
        self.kernel = kernel
        if mean_function is not None:
            raise NotImplementedError("only the zero mean function is used by OAK (model_utils.py:152,159)")
        self.mean_function = lambda X: 0.0
        self.likelihood = Gaussian(noise_variance)
        self._Xd = None
        self._Yd = None

    # ---- data on the device (moved once) --------------------------------------------------
    def _device_data(self):
        if self._Xd is None:
            X, Y = self.data
            self._Xd = _device.to_device(X)
            self._Yd = _device.to_device(Y)
            if self._Yd.shape[1] != 1:
                raise NotImplementedError("single-output regression only (R = 1 on this path)")
        return self._Xd, self._Yd

    def _slice_for_kernel(self, Xd):
        return self.kernel.slice(Xd, None)[0].contiguous()

    def _sliced_training_inputs(self, Xd):
        """``_slice_for_kernel`` of the RESIDENT training inputs, kept across evaluations (the data and the kernel's
        active dimensions are fixed; re-gathering the columns was an index kernel and ~40 us of host time before
        the first tile of every objective evaluation)."""
        dims = self.kernel.active_dims
        sig = (dims.start, dims.stop, dims.step) if isinstance(dims, slice) else tuple(np.asarray(dims).tolist())
        key = (Xd.data_ptr(), tuple(Xd.shape), sig)
        cached = getattr(self, "_xs_cache", None)
        if cached is None or cached[0] != key:
            self._xs_cache = (key, self._slice_for_kernel(Xd))
        return self._xs_cache[1]

    def _check_training_discrete(self, Xs, dims):
        """Range check of the discrete columns of the RESIDENT training inputs: once per (data, kernel layout),
        not on every objective evaluation (it costs a device read-back)."""
        key = (Xs.data_ptr(), tuple(Xs.shape), tuple((d.column, d.count) for d in dims if d.type != 0))
        if getattr(self, "_discrete_checked", None) != key:
            self.kernel._check_discrete(Xs, dims)
            self._discrete_checked = key

    # ---- gpflow-style objective API --------------------------------------------------------
    def log_prior_density(self) -> float:
        total = 0.0
        for p in collect_parameters(self):
            if p.prior is not None and p.trainable:  # gpflow sums over trainable_parameters only
                total += float(np.sum(p.prior.log_prob(p.numpy())))
        return total

    def maximum_log_likelihood_objective(self) -> float:
        raise NotImplementedError

    def training_loss(self) -> float:
        return -(self.maximum_log_likelihood_objective() + self.log_prior_density())

    def training_loss_closure(self):
        return self.training_loss

    def predict_log_density(self, data):
        """gpflow ``GPModel.predict_log_density`` for the Gaussian likelihood: per point
        log N(y; f_mean, f_var + noise) (used by ``oak_model.get_loglik``, model_utils.py:445-460)."""
        X, Y = data
        mean, var = self.predict_f(X)
        mean, var = np.asarray(mean, dtype=np.float64), np.asarray(var, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64).reshape(mean.shape)
        s2 = var + scalar_of(self.likelihood.variance)
        return (-0.5 * np.log(2.0 * np.pi * s2) - 0.5 * (Y - mean) ** 2 / s2).sum(-1)

    def training_loss_and_grad(self):
        """(training_loss, d/d unconstrained trainables): what TensorFlow autodiff hands to the scipy
        optimiser in the reference (model_utils.py:168-175); see training.py."""
        from .training import training_loss_and_grad

        return training_loss_and_grad(self)


class GPR(GPModel):
    """Exact GP regression: ``log N(y; 0, K + noise I)`` (gpflow ``GPR``; model_utils.py:159)."""

    def __init__(self, data, kernel, mean_function=None, noise_variance: float = 1.0):
        super().__init__(data, kernel, mean_function, noise_variance)

    def _factorise(self):
        Xd, Yd = self._device_data()
        K = self.kernel._K_device(self._slice_for_kernel(Xd))
        lml, alpha = _device.gpr_finish(K, Yd, scalar_of(self.likelihood.variance))
        return K, lml, alpha  # K now holds the Cholesky factor (upper triangle, row-major)

    def log_marginal_likelihood(self) -> float:
        return float(self._factorise()[1].item())

    def maximum_log_likelihood_objective(self) -> float:
        return self.log_marginal_likelihood()

    def sufficient_statistics(self):
        """alpha = (K + noise I)^-1 y on the device, shape (N, 1) (oak/utils.py:206-211)."""
        return self._factorise()[2].reshape(-1, 1)

    def predict_f(self, Xnew, full_cov: bool = False) -> Tuple[object, object]:
        import torch

        if full_cov:
            raise NotImplementedError("marginal variances only (what the reference reads, model_utils.py:429-443)")
        host = _device.is_host(Xnew)
        Xd, _ = self._device_data()
        Xs = self._slice_for_kernel(Xd)
        Xn = self._slice_for_kernel(_device.to_device(Xnew))
        Kfac, _, alpha = self._factorise()
        spec = self.kernel._make_spec()
        try:
            self.kernel._check_discrete(Xn, spec._keep)
            pn, px = _device.Points(spec, Xn), _device.Points(spec, Xs)
            Kns = _device.gram(spec, pn, px)  # (N*, N)
            kdiag = _device.gram_diag(spec, pn)
        finally:
            spec.close()
        mean = Kns @ alpha.reshape(-1, 1)
        # var = k** - |L^-1 k*|^2 ; Kfac holds L^T in its upper triangle (library triangular solve)
        A = torch.linalg.solve_triangular(torch.triu(Kfac).T, Kns.T, upper=False)
        var = (kdiag - (A * A).sum(0)).reshape(-1, 1)
        return _device.from_device(mean, host), _device.from_device(var, host)


class SGPR(GPModel):
    """Titsias' sparse GP regression bound (gpflow 2.2.1 ``SGPR``; model_utils.py:150-155).

    Every evaluation factors Kuu first (``oak_sgpr_factor_f64``) and picks, on the device, between the
    un-whitened statistics Phi = Kuf Kuf^T (fast) and gpflow's operation order A = L^-1 Kuf, A A^T
    (``whiten_stats``: None = from the condition estimate of Kuu, True / False = forced); one device-to-host
    read per evaluation carries the bound and both Cholesky status codes.

    ``distributed=True`` declares that every rank holds its own SHARD of (X, y): the packed statistics are then
    summed over the ranks by one all-reduce and ``elbo`` / ``predict_f`` / ``sufficient_statistics`` become
    collective calls.  It is never inferred from ``torch.distributed`` being initialised: ranks that each hold
    the full data set (a generic DDP launch) would otherwise count it world-size times."""

    def __init__(self, data, kernel, inducing_variable, mean_function=None, noise_variance: float = 1.0,
                 chunk: int = 262144, distributed: bool = False, whiten_stats: Optional[bool] = None,
                 cond_threshold: float = 0.0):
        super().__init__(data, kernel, mean_function, noise_variance)
        if not isinstance(inducing_variable, InducingPoints):
            inducing_variable = InducingPoints(inducing_variable)
        self.inducing_variable = inducing_variable
        self.chunk = int(chunk)
        self.distributed = bool(distributed)
        if self.distributed and not parallel.is_distributed():
            raise ValueError("distributed=True needs an initialised torch.distributed process group (world size > 1)")
        self.whiten_stats = whiten_stats
        self.cond_threshold = float(cond_threshold)
        self.overlap_ctas = -1      # CTAs lent to the overlapped factorisation (-1 automatic, 0 serial)
        self.last_route = None      # route / condition estimate of the last evaluation (diagnostics)
        self.last_cond_estimate = None
        self.last_timings = {}

    def _Z_device(self):
        """Inducing points on the device; the upload is skipped while the host values are unchanged (they are
        fixed by default, model_utils.py:100-101)."""
        z = np.asarray(value_of(self.inducing_variable.Z), dtype=np.float64)
        cached = getattr(self, "_z_cache", None)
        if cached is None or cached[0].shape != z.shape or not np.array_equal(cached[0], z):
            self._z_cache = (z.copy(), self._slice_for_kernel(_device.to_device(z)))
        return self._z_cache[1]

    def _route(self) -> int:
        return _device.ROUTE_AUTO if self.whiten_stats is None else int(bool(self.whiten_stats))

    def _statistics(self, want_alpha: bool):
        """Returns (tail, factor, n_total): ``tail.out`` = [elbo, sum log diag LB, tr AAT, c^T c, info, info, route,
        cond]; nothing has been synchronised yet (``tail.host()`` does, and raises on a failed factorisation)."""
        Xd, Yd = self._device_data()
        Xs = self._sliced_training_inputs(Xd)
        Zs = self._Z_device()
        spec = self.kernel._make_spec()
        try:
            self._check_training_discrete(Xs, spec._keep)
            self.kernel._check_discrete(Zs, spec._keep)
            pz, px = _device.Points(spec, Zs), _device.Points(spec, Xs)
            # Kuu(iv, kernel) + jitter I, L = chol, L^-1, route flag -- before the statistics, all on the device
            # (one call: the factorisation is hidden behind the first chunk's Kuf tiles)
            fac, stats = _device.sgpr_factor_stats(spec, pz, px, Yd, DEFAULT_JITTER, route=self._route(),
                                                   cond_threshold=self.cond_threshold, chunk=self.chunk,
                                                   overlap_ctas=self.overlap_ctas)
            n_total = int(Xs.shape[0])
            if self.distributed:
                parallel.allreduce_sum_(stats)
                n_total = parallel.global_count(self, n_total)
        finally:
            spec.close()
        tail = _device.sgpr_finish2(fac, stats, n_total, scalar_of(self.likelihood.variance), want_alpha=want_alpha)
        return tail, fac, n_total

    def _record(self, o):
        self.last_route, self.last_cond_estimate = int(o[6]), float(o[7])

    def elbo(self) -> float:
        o = self._statistics(False)[0].host()
        self._record(o)
        return float(o[0])

    def maximum_log_likelihood_objective(self) -> float:
        return self.elbo()

    def sufficient_statistics(self):
        """alpha = L^-T LB^-T c, shape (M, 1) (oak/utils.py:195-198)."""
        tail = self._statistics(True)[0]
        self._record(tail.host())
        return tail.alpha.reshape(-1, 1)

    def predict_mean(self, Xnew):
        """Mean of ``predict_f`` only (what ``oak_model.predict`` uses, model_utils.py:429-443):
        Kus^T alpha through the fused Gram-matrix/vector tiles, no (M, N*) intermediate."""
        host = _device.is_host(Xnew)
        alpha = self.sufficient_statistics()
        Xn = self._slice_for_kernel(_device.to_device(Xnew))
        Zs = self._Z_device()
        spec = self.kernel._make_spec()
        try:
            self.kernel._check_discrete(Xn, spec._keep)
            mean = _device.gram_matvec(spec, _device.Points(spec, Xn), _device.Points(spec, Zs), alpha)
        finally:
            spec.close()
        return _device.from_device(mean.reshape(-1, 1), host)

    def predict_f(self, Xnew, full_cov: bool = False):
        import torch

        if full_cov:
            raise NotImplementedError("marginal variances only (what the reference reads, model_utils.py:429-443)")
        host = _device.is_host(Xnew)
        tail, fac, _ = self._statistics(True)
        self._record(tail.host())
        Xn = self._slice_for_kernel(_device.to_device(Xnew))
        Zs = self._Z_device()
        spec = self.kernel._make_spec()
        try:
            self.kernel._check_discrete(Xn, spec._keep)
            pn, pz = _device.Points(spec, Xn), _device.Points(spec, Zs)
            Kus = _device.gram(spec, pz, pn)  # (M, N*)
            kdiag = _device.gram_diag(spec, pn)
        finally:
            spec.close()
        mean = Kus.T @ tail.alpha.reshape(-1, 1)
        tmp1 = torch.linalg.solve_triangular(fac.L(), Kus, upper=False)
        tmp2 = torch.linalg.solve_triangular(tail.LB(), tmp1, upper=False)
        var = (kdiag + (tmp2 * tmp2).sum(0) - (tmp1 * tmp1).sum(0)).reshape(-1, 1)
        return _device.from_device(mean, host), _device.from_device(var, host)


class SVGP(Module):
    """Whitened sparse variational GP with a diagonal q(u) and a Bernoulli likelihood: the model of the
    reference's classification run, ``gpflow.models.SVGP(kernel, likelihood=Bernoulli(invlink=inv_logit),
    inducing_variable=Z, whiten=True, q_diag=True)`` trained full-batch with BFGS
    (examples/uci/uci_classification_train.py:108-124).  gpflow 2.2.1 semantics (``elbo(data)``,
    ``predict_f``, ``predict_log_density``, ``posterior().alpha``) are restated in csrc/oak_svgp.cu and
    ``training.svgp_elbo_and_grad``; Kuf / Kuu / K_diag come from the fused OAK tiles, the M x M and M x n
    triangular algebra from cuSOLVER / cuBLAS."""

    def __init__(self, kernel, likelihood, inducing_variable, *, mean_function=None, num_latent_gps: int = 1,
                 q_diag: bool = False, q_mu=None, q_sqrt=None, whiten: bool = True, num_data: Optional[int] = None,
                 chunk: int = 65536, distributed: bool = False):
        if not (whiten and q_diag):
            raise NotImplementedError("only whiten=True, q_diag=True (the reference's configuration) is built")
        if num_latent_gps != 1 or mean_function is not None:
            raise NotImplementedError("one latent GP with the zero mean function")
        from ._gpflow_shim import Bernoulli, positive

        if not isinstance(likelihood, Bernoulli):
            raise NotImplementedError("SVGP is built for the Bernoulli likelihood (classification)")
        self.kernel = kernel
        self.likelihood = likelihood
        if not isinstance(inducing_variable, InducingPoints):
            inducing_variable = InducingPoints(inducing_variable)
        self.inducing_variable = inducing_variable
        m = inducing_variable.num_inducing
        self.q_mu = Parameter(np.zeros((m, 1)) if q_mu is None else np.asarray(q_mu, dtype=np.float64).reshape(m, 1))
        self.q_sqrt = Parameter(np.ones((m, 1)) if q_sqrt is None else np.asarray(q_sqrt, dtype=np.float64).reshape(m, 1),
                                transform=positive())
        self.num_data = num_data
        self.mean_function = lambda X: 0.0
        self.chunk = int(chunk)
        # distributed=True: every rank passes its own SHARD of (X, Y) to elbo / training_loss; the data sums of the
        # bound and of its gradient are combined by one all-reduce (declared by the caller, never inferred)
        self.distributed = bool(distributed)
        if self.distributed and not parallel.is_distributed():
            raise ValueError("distributed=True needs an initialised torch.distributed process group (world size > 1)")
        self.data = None  # the reference assigns oak.m.data before the Sobol step (:145)

    _slice_for_kernel = GPModel._slice_for_kernel
    log_prior_density = GPModel.log_prior_density

    def _Z_device(self):
        return self._slice_for_kernel(_device.to_device(value_of(self.inducing_variable.Z)))

    # ---- objective ----------------------------------------------------------------------------
    def elbo(self, data) -> float:
        from .training import svgp_elbo_and_grad

        return svgp_elbo_and_grad(self, data, want_grad=False)[0]

    def maximum_log_likelihood_objective(self, data) -> float:
        return self.elbo(data)

    def training_loss(self, data) -> float:
        return -(self.elbo(data) + self.log_prior_density())

    def training_loss_closure(self, data):
        closure = lambda: self.training_loss(data)
        closure.model, closure.data = self, data  # picked up by training.optimise
        return closure

    def training_loss_and_grad(self, data):
        from .training import training_loss_and_grad

        return training_loss_and_grad(self, data)

    # ---- prediction -----------------------------------------------------------------------------
    def _conditional(self, Xnew):
        """(mean, var) of q(f) at Xnew on the device (gpflow ``base_conditional``, white=True)."""
        import torch

        Xn = self._slice_for_kernel(_device.to_device(Xnew))
        Zs = self._Z_device()
        q_mu = _device.to_device(self.q_mu.numpy(), ndim=1).reshape(-1)
        q_sqrt = _device.to_device(self.q_sqrt.numpy(), ndim=1).reshape(-1)
        spec = self.kernel._make_spec()
        try:
            self.kernel._check_discrete(Xn, spec._keep)
            self.kernel._check_discrete(Zs, spec._keep)
            pz = _device.Points(spec, Zs)
            Kuu = _device.gram(spec, pz)
            Kuu.diagonal().add_(DEFAULT_JITTER)
            L = torch.linalg.cholesky(Kuu)
            n = int(Xn.shape[0])
            mean = torch.empty(n, dtype=torch.float64, device=Xn.device)
            var = torch.empty(n, dtype=torch.float64, device=Xn.device)
            for c0 in range(0, n, self.chunk):
                c1 = min(c0 + self.chunk, n)
                pxc = _device.Points(spec, Xn[c0:c1])
                A = torch.linalg.solve_triangular(L, _device.gram(spec, pz, pxc), upper=False).contiguous()
                mean[c0:c1], var[c0:c1] = _device.svgp_moments(A, q_mu, q_sqrt, _device.gram_diag(spec, pxc))
        finally:
            spec.close()
        return mean, var

    def predict_f(self, Xnew, full_cov: bool = False):
        if full_cov:
            raise NotImplementedError("marginal variances only (what the reference reads, :128)")
        host = _device.is_host(Xnew)
        mean, var = self._conditional(Xnew)
        return _device.from_device(mean.reshape(-1, 1), host), _device.from_device(var.reshape(-1, 1), host)

    def predict_log_density(self, data):
        """log of the Gauss-Hermite predictive density of each (x, y) (gpflow ``predict_log_density``; :135)."""
        X, Y = data
        host = _device.is_host(X)
        mean, var = self._conditional(X)
        y = _device.to_device(Y, ndim=1).reshape(-1)
        lk = self.likelihood
        out = _device.bernoulli_quadrature(mean, var, y, lk.invlink.kind, lk.invlink.jitter,
                                           lk.num_gauss_hermite_points, want=("logdensity",))["logdensity"]
        return _device.from_device(out, host)

    def sufficient_statistics(self):
        """``posterior().alpha`` of the whitened model: L^-T q_mu, shape (M, 1) (oak/utils.py:174-177)."""
        import torch

        Zs = self._Z_device()
        spec = self.kernel._make_spec()
        try:
            Kuu = _device.gram(spec, _device.Points(spec, Zs))
        finally:
            spec.close()
        Kuu.diagonal().add_(DEFAULT_JITTER)
        L = torch.linalg.cholesky(Kuu)
        q_mu = _device.to_device(self.q_mu.numpy()).reshape(-1, 1)
        return torch.linalg.solve_triangular(L.T, q_mu, upper=True)
