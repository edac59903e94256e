"""Objective gradients and the hyper-parameter optimisation step.

The reference trains with ``gpflow.optimizers.Scipy().minimize(model.training_loss_closure(),
model.trainable_variables, method="BFGS")`` (oak/model_utils.py:168-175, 410-427) on the unconstrained
variables, gradients by TensorFlow autodiff through ``OAKKernel.K``.  Here the gradient of the
objective with respect to the kernel matrices is formed on the device (dense M x M / N x N algebra
through the cuSOLVER / cuBLAS bindings of torch -- plumbing), and contracted with dK/d theta by the
backward tiles of ``liboak_b200.so`` (``oak_gram_backward_f64``): no per-dimension derivative
matrix is ever formed.

Supported trainable parameters: RBF lengthscales under every measure (Gaussian -- the OAK default
after the normalising flow --, empirical, uniform, mixture of Gaussians, or none), the order variances sigma^2_0..P
(``share_var_across_orders=True``; with ``False`` only sigma^2_0 exists and the sub-kernels' own variances carry
the scale, oak_kernel.py:217-221), the likelihood variance, the base variances s^2 of the RBF sub-kernels
where the reference keeps them trainable, and W / kappa of the categorical sub-kernels (through the
cotangent of their B tables), and the inducing points Z of SGPR when ``zfixed=False``
(``oak_gram_backward_rows_f64``; the columns of discrete sub-kernels carry no gradient, as under
``tf.cast`` / ``tf.gather``).  Any other *trainable* parameter
raises ``NotImplementedError`` (set it non-trainable to keep it fixed).

SGPR (gpflow 2.2.1 ``SGPR.elbo``), with Phi = Kuf Kuf^T, b = Kuf y, s = sum K_diag,
Q = Kuu + jitter I, S = Q + Phi / noise:

    elbo = -N/2 log 2pi - 1/2 (log|S| - log|Q|) - N/2 log noise - y^T y / (2 noise)
           + b^T S^-1 b / (2 noise^2) - s / (2 noise) + tr(Q^-1 Phi) / (2 noise)

    d/dPhi = -S^-1/(2 noise) - S^-1 b b^T S^-1/(2 noise^3) + Q^-1/(2 noise)
    d/db   = S^-1 b / noise^2            d/ds = -1/(2 noise)
    d/dQ   = -S^-1/2 + Q^-1/2 - S^-1 b b^T S^-1/(2 noise^2) - Q^-1 Phi Q^-1/(2 noise)

(checked against torch autograd of the gpflow operation order in tests/test_gpu_training.py).
S^-1 and Q^-1 are never formed from S itself: with L = chol(Q), B = I + L^-1 Phi L^-T / noise (gpflow's
whitened matrix, always well conditioned), S^-1 = L^-T B^-1 L^-1 and log|S| - log|Q| = log|B|.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

from . import _cabi, _device, parallel
from ._gpflow_shim import DEFAULT_JITTER, Identity, Parameter, Sigmoid, Softplus, collect_parameters, scalar_of, value_of


# ---- d constrained / d unconstrained of the gpflow transforms -------------------------------
def _transform_grad(p: Parameter) -> np.ndarray:
    u = np.asarray(p.unconstrained_variable, dtype=np.float64)
    t = p.transform
    if isinstance(t, Identity):
        return np.ones_like(u)
    if isinstance(t, Softplus):
        return 1.0 / (1.0 + np.exp(-u))
    if isinstance(t, Sigmoid):
        s = 1.0 / (1.0 + np.exp(-u))
        return (t.high - t.low) * s * (1.0 - s)
    raise NotImplementedError(f"transform {type(t).__name__}")


def _prior_grad(p: Parameter) -> np.ndarray:
    """d log prior / d constrained value.  The reference only attaches Gamma(1, 0.2) to the order variances
    (oak/model_utils.py:163-165); a prior object may also bring its own ``grad_log_prob``."""
    from ._gpflow_shim import Gamma

    x = p.numpy()
    pr = p.prior
    if isinstance(pr, Gamma):
        return (pr.concentration - 1.0) / x - pr.rate
    if hasattr(pr, "grad_log_prob"):
        return np.asarray(pr.grad_log_prob(x), dtype=np.float64)
    raise NotImplementedError(f"gradient of the prior {type(pr).__name__} is not implemented (Gamma, or a prior with "
                              "grad_log_prob)")


def _lengthscale_parameter(sub):
    base = getattr(sub, "base_kernel", sub)
    return getattr(base, "lengthscales", None)


def _supported_parameters(model) -> Tuple[List[Parameter], List[Parameter], Parameter]:
    """(per-dimension lengthscale Parameters or None, order variances, noise -- None without a Gaussian likelihood)."""
    kern = model.kernel
    ls = [_lengthscale_parameter(k) for k in kern.kernels]
    return ls, list(kern.variances), getattr(model.likelihood, "variance", None)


def _variational_parameters(model) -> List[Parameter]:
    return [p for p in (getattr(model, "q_mu", None), getattr(model, "q_sqrt", None)) if isinstance(p, Parameter)]


def _base_variance_parameter(sub):
    """s^2 of an RBF sub-kernel when it is a Parameter (the reference fixes it to 1 only for the Gaussian
    measure with shared order variances, oak_kernel.py:163-166)."""
    if hasattr(sub, "kappa") or hasattr(sub, "p0"):
        return None
    base = getattr(sub, "base_kernel", sub)
    p = getattr(base, "variance", None)
    return p if isinstance(p, Parameter) else None


def _discrete_parameters(model) -> List[Parameter]:
    """W, kappa and (when it is a Parameter) variance of the categorical sub-kernels, variance of the
    binary ones: differentiated through the cotangent of their B tables."""
    out = []
    for k in model.kernel.kernels:
        if hasattr(k, "kappa") or hasattr(k, "p0"):
            for name in ("W", "kappa", "variance"):
                p = getattr(k, name, None)
                if isinstance(p, Parameter):
                    out.append(p)
    return out


def _inducing_parameter(model):
    iv = getattr(model, "inducing_variable", None)
    return getattr(iv, "Z", None)


def _all_supported_ids(model):
    ls, var, noise = _supported_parameters(model)
    bvar = [_base_variance_parameter(k) for k in model.kernel.kernels]
    z = _inducing_parameter(model)
    return ({id(p) for p in ls if p is not None} | {id(p) for p in var} | ({id(noise)} if noise is not None else set())
            | {id(p) for p in _discrete_parameters(model)} | {id(p) for p in bvar if p is not None}
            | ({id(z)} if z is not None else set()) | {id(p) for p in _variational_parameters(model)})


def freeze_unsupported(model) -> List[Parameter]:
    """Sets ``trainable=False`` on every parameter the backward tiles cannot differentiate; returns them."""
    ok = _all_supported_ids(model)
    frozen = []
    for p in collect_parameters(model):
        if p.trainable and id(p) not in ok:
            p.trainable = False
            frozen.append(p)
    return frozen


def _check_trainables(model, spec_dims):
    ls, var, noise = _supported_parameters(model)
    ok = _all_supported_ids(model)
    for p in collect_parameters(model):
        if p.trainable and id(p) not in ok:
            raise NotImplementedError(
                "gradient of a trainable parameter outside (RBF lengthscales, order / base variances, likelihood "
                f"variance, categorical W / kappa, inducing points) is not implemented ({p!r}); set it non-trainable (freeze_unsupported)")
    # share_var_across_orders=False (Duvenaud-style, oak_kernel.py:217-221, 262-265): the spec carries
    # sigma^2_n = 1 for n >= 1, only variances[0] exists as a Parameter and the base variances of the sub-kernels
    # are the trainable scales -- all of them already have gradients (order slot 0, base-variance and table slots)
    for p, d in zip(ls, spec_dims):
        if p is not None and p.trainable and d.type != _cabi.DIM_RBF:
            raise NotImplementedError("a lengthscale on a non-RBF sub-kernel cannot be differentiated")


def _chol_inverse(Q, what: str):
    """(L^-1, sum log diag L) of the symmetric positive definite device matrix ``Q`` = L L^T through
    ``oak_chol_f64`` with an identity border (csrc/oak_chol.cu).  Raises ``RuntimeError`` naming the failing leading
    minor, as ``torch.linalg.cholesky`` does (``optimise`` turns that into a rejected line-search point)."""
    torch = _device._torch()
    m = int(Q.shape[0])
    gap = (-m) % 8
    buf = torch.zeros((m, 2 * m + gap + (gap + 2 * m) % 2), dtype=torch.float64, device=Q.device)
    buf[:, :m] = Q                      # row j of the tensor = column j of the matrix; Q is symmetric
    buf[:, m + gap: 2 * m + gap].fill_diagonal_(1.0)
    info, logdet = _device.chol(buf, m, 2 * m, gap=gap, border_identity=True)
    k = int(info.item())
    if k != 0:
        raise RuntimeError(f"Cholesky: {what} is not positive definite (leading minor of order {k})")
    return buf[:, m + gap: 2 * m + gap], float(logdet.item())


# ---- objectives with gradients (constrained space) --------------------------------------------
def sgpr_elbo_and_grad(model) -> Tuple[float, np.ndarray, np.ndarray, float]:
    """(elbo, d/d lengthscales [num sub-kernels], d/d order variances [P+1], d/d noise)."""
    torch = _device._torch()
    Xd, Yd = model._device_data()
    Xs = model._sliced_training_inputs(Xd)
    Zs = model._Z_device()
    noise = scalar_of(model.likelihood.variance)
    kern = model.kernel
    spec = kern._make_spec()
    try:
        _check_trainables(model, spec._keep)
        kern._check_discrete(Xs, spec._keep)
        kern._check_discrete(Zs, spec._keep)
        pz, px = _device.Points(spec, Zs), _device.Points(spec, Xs)
        m, n_local = pz.n, px.n
        # Kuf is kept for the backward pass when it fits comfortably (8.2 GB at N = 10^6, M = 1024);
        # otherwise its chunks are recomputed in the second pass
        keep = getattr(model, "keep_kuf", None)
        if keep is None:
            keep = 8.0 * m * n_local < 48e9
        if keep:
            # the 8 GB buffer is kept on the model between evaluations (an optimiser calls this in a loop)
            stats, kuf_blocks, chunk_eff, model._kuf_store = _device.sgpr_stats(
                spec, pz, px, Yd, chunk=model.chunk, keep_kuf=True, kuf_store=getattr(model, "_kuf_store", None))
        else:
            stats = _device.sgpr_stats(spec, pz, px, Yd, chunk=model.chunk)
        n_total = n_local
        if model.distributed:
            parallel.allreduce_sum_(stats)
            n_total = parallel.global_count(model, n_total)
        Kuu = _device.gram(spec, pz)
        # statistics: Phi arrives as column-major lower == upper triangle of the row-major view
        U = torch.triu(stats[: m * m].view(m, m))
        Phi = U + torch.triu(U, 1).T
        b = stats[m * m: m * m + m].reshape(m, 1)
        s, yty = float(stats[m * m + m]), float(stats[m * m + m + 1])
        eye = torch.eye(m, dtype=torch.float64, device=Kuu.device)
        Q = Kuu + DEFAULT_JITTER * eye
        # whitened algebra (gpflow's operation order): B = I + L^-1 Phi L^-T / noise is always well
        # conditioned, S^-1 = L^-T B^-1 L^-1, Q^-1 = L^-T L^-1, log|S| - log|Q| = log|B|
        # both factorisations by the one-launch bordered Cholesky (identity border -> L^-1 comes with L: 0.41 ms
        # at M = 1024 against ~1 ms for cuSOLVER's potrf alone, plus the trsm / potri it replaces); this M x M
        # algebra is replicated on every rank, so it is what does not shrink with the rank count
        Linv, _ = _chol_inverse(Q, "Kuu + jitter I")
        Phiw = Linv @ Phi @ Linv.T
        B = eye + Phiw / noise
        LBinv, half_logdet = _chol_inverse(0.5 * (B + B.T), "B = I + L^-1 Phi L^-T / noise")
        Binv = LBinv.T @ LBinv
        Qi = Linv.T @ Linv
        Si = Linv.T @ Binv @ Linv
        Sib = Si @ b
        QiPhi = Qi @ Phi
        logdet = 2.0 * half_logdet
        bSb = float((b * Sib).sum())
        trQiPhi = float(torch.diagonal(QiPhi).sum())
        elbo = (-0.5 * n_total * math.log(2.0 * math.pi) - 0.5 * logdet - 0.5 * n_total * math.log(noise)
                - 0.5 * yty / noise + 0.5 * bSb / noise ** 2 - 0.5 * s / noise + 0.5 * trQiPhi / noise)
        SbbS = Sib @ Sib.T
        G_phi = -0.5 * Si / noise - SbbS / (2.0 * noise ** 3) + Qi / (2.0 * noise)
        g_b = Sib / noise ** 2
        G_Q = -0.5 * Si + 0.5 * Qi - SbbS / (2.0 * noise ** 2) - (QiPhi @ Qi) / (2.0 * noise)
        g_s = -0.5 / noise
        g_noise = (-0.5 * n_total / noise + 0.5 * yty / noise ** 2 - bSb / noise ** 3 + 0.5 * s / noise ** 2
                   - 0.5 * trQiPhi / noise ** 2 + 0.5 * float((Si * Phi).sum()) / noise ** 2
                   + 0.5 * float((Sib * (Phi @ Sib)).sum()) / noise ** 4)
        # second pass over the local points: W^T = 2 Kuf^T G_phi + y g_b^T, contracted by the backward tiles
        nout = int(_cabi.load().oak_backward_grad_count(spec.handle))  # lengthscales | variances | table blob
        grad = torch.zeros(nout, dtype=torch.float64, device=Kuu.device)
        G2 = (2.0 * G_phi).contiguous()
        zpar = _inducing_parameter(model)
        z_train = zpar is not None and zpar.trainable
        gZ = torch.zeros((m, spec.num_dims), dtype=torch.float64, device=Kuu.device) if z_train else None

        yvec = Yd.reshape(-1)
        gb_vec = g_b.reshape(-1).contiguous()
        if m % 2:  # the DMMA panel product wants an even pitch of its left factor
            G2p = torch.zeros((m, m + 1), dtype=torch.float64, device=Kuu.device)
            G2p[:, :m] = G2
            G2 = G2p[:, :m]

        def rows_chunk(Kc, c0, c1):
            # W = 2 G_phi Kuf + g_b y^T for one chunk (M x nc), rows = inducing points: one pass of the in-house
            # DMMA panel product (csrc/oak_pgemm.cu) with the rank-1 term in its epilogue
            if Kc.stride(0) % 2 == 0 and Kc.data_ptr() % 16 == 0 and (c0 % 2 == 0):
                Wc = _device.panel_gemm(G2, Kc, u=gb_vec, v=yvec[c0:c1])
            else:
                Wc = torch.addmm(g_b @ Yd[c0:c1].reshape(1, -1), G2, Kc)
            pxc = _device.Points(spec, Xs[c0:c1])
            if z_train:
                _device.gram_backward_rows(spec, pz, Wc, px2=pxc, grad=grad, grad_rows=gZ)
            else:
                _device.gram_backward(spec, pz, Wc, px2=pxc, grad=grad)

        if keep:
            for c, Kc in enumerate(kuf_blocks):
                rows_chunk(Kc, c * chunk_eff, c * chunk_eff + Kc.shape[1])
            del kuf_blocks
        elif z_train:
            chunk = max(64, (int(model.chunk) + 63) // 64 * 64)
            for c0 in range(0, n_local, chunk):
                c1 = min(c0 + chunk, n_local)
                rows_chunk(_device.gram(spec, pz, _device.Points(spec, Xs[c0:c1])), c0, c1)
        else:
            chunk = max(64, (int(model.chunk) + 63) // 64 * 64)
            for c0 in range(0, n_local, chunk):
                c1 = min(c0 + chunk, n_local)
                Kt = _device.gram(spec, px, pz, row_begin=c0, row_end=c1)  # (nc, M) = Kuf[:, c0:c1]^T
                Wt = torch.addmm(Yd[c0:c1].reshape(-1, 1) @ g_b.T, Kt, G2)
                _device.gram_backward(spec, px, Wt, px2=pz, row_begin=c0, row_end=c1, grad=grad)
                del Kt, Wt
        _device.gram_diag_backward(spec, px, wscale=g_s, grad=grad)
        if model.distributed:
            parallel.allreduce_sum_(grad)
            if z_train:
                parallel.allreduce_sum_(gZ)
        # Kuu term (replicated, added once)
        if z_train:
            # rows only differentiate the first argument of K(Z, Z): cotangent G_Q + G_Q^T, and half of the
            # parameter gradients it yields
            g_uu = torch.zeros_like(grad)
            _device.gram_backward_rows(spec, pz, (G_Q + G_Q.T).contiguous(), grad=g_uu, grad_rows=gZ)
            grad.add_(g_uu, alpha=0.5)
            # sub-kernel order -> columns of Z
            gz_full = np.zeros(zpar.numpy().shape, dtype=np.float64)
            gz_host = gZ.cpu().numpy()
            for i, d in enumerate(spec._keep):
                gz_full[:, d.column] += gz_host[:, i]
            model._inducing_grad = gz_full
        else:
            _device.gram_backward(spec, pz, G_Q.contiguous(), grad=grad)
        g = grad.cpu().numpy()
        layout = [_device.table_layout(spec, i) for i in range(spec.num_dims)]
    finally:
        spec.close()
    D, P1 = spec.num_dims, max(spec.depth, 1) + 1
    model._table_cotangent = (g[D + P1: len(g) - D].copy(), layout)
    model._base_variance_grad = g[len(g) - D:].copy()
    return elbo, g[:D].copy(), g[D: D + P1].copy(), float(g_noise)


def gpr_lml_and_grad(model) -> Tuple[float, np.ndarray, np.ndarray, float]:
    """(log marginal likelihood, d/d lengthscales, d/d order variances, d/d noise)."""
    torch = _device._torch()
    Xd, Yd = model._device_data()
    Xs = model._sliced_training_inputs(Xd)
    noise = scalar_of(model.likelihood.variance)
    kern = model.kernel
    spec = kern._make_spec()
    try:
        _check_trainables(model, spec._keep)
        kern._check_discrete(Xs, spec._keep)
        px = _device.Points(spec, Xs)
        n = px.n
        K = _device.gram(spec, px)
        K.diagonal().add_(noise)
        L = torch.linalg.cholesky(K)
        alpha = torch.cholesky_solve(Yd.reshape(-1, 1), L)
        lml = float(-0.5 * (Yd.reshape(-1, 1) * alpha).sum() - torch.log(torch.diagonal(L)).sum()
                    - 0.5 * n * math.log(2.0 * math.pi))
        Kinv = torch.cholesky_inverse(L)
        W = (0.5 * (alpha @ alpha.T - Kinv)).contiguous()
        g_noise = float(torch.diagonal(W).sum())
        g = _device.gram_backward(spec, px, W).cpu().numpy()
        layout = [_device.table_layout(spec, i) for i in range(spec.num_dims)]
    finally:
        spec.close()
    D, P1 = spec.num_dims, max(spec.depth, 1) + 1
    model._table_cotangent = (g[D + P1: len(g) - D].copy(), layout)
    model._base_variance_grad = g[len(g) - D:].copy()
    return lml, g[:D].copy(), g[D: D + P1].copy(), g_noise


def svgp_elbo_and_grad(model, data, want_grad: bool = True):
    """Whitened SVGP bound with diagonal q(u) and Bernoulli likelihood (gpflow 2.2.1 ``SVGP.elbo``):

        A = L^-1 Kuf,  mean = A^T q_mu,  var = K_diag - sum_m A^2 (1 - q_sqrt^2)
        elbo = scale * sum_i E_{N(mean_i, var_i)}[log p(y_i | f)] - KL,   scale = num_data / N (1 when num_data is None)
        KL = (sum q_mu^2 - M - sum log q_sqrt^2 + sum q_sqrt^2) / 2

    Returns (elbo, d/d lengthscales, d/d order variances, None); the gradients of q_mu / q_sqrt (and of Z when
    trainable) are left on the model.  Backward chain: Abar = dE/dA from ``oak_svgp_moments_backward_f64``,
    Kuf-bar = L^-T Abar, L-bar = -tril(L^-T Abar A^T), Kuu-bar by the Cholesky adjoint
    (1/2) L^-T (P + P^T) L^-1 with P = tril(L^T L-bar), diagonal halved; all contracted with dK/d theta by the
    backward tiles."""
    torch = _device._torch()
    X, Y = data
    Xs = model._slice_for_kernel(_device.to_device(X))
    y = _device.to_device(Y, ndim=1).reshape(-1)
    Zs = model._Z_device()
    kern = model.kernel
    lk = model.likelihood
    spec = kern._make_spec()
    try:
        if want_grad:
            _check_trainables(model, spec._keep)
        kern._check_discrete(Xs, spec._keep)
        kern._check_discrete(Zs, spec._keep)
        pz = _device.Points(spec, Zs)
        m, n = pz.n, int(Xs.shape[0])
        if y.numel() != n:
            raise ValueError("one label per input row")
        # N axis sharded over ranks (like the SGPR statistics): the minibatch scale refers to the global batch
        sharded = bool(getattr(model, "distributed", False))
        # (exchanged only when the minibatch scale needs it: the batch is an argument here, not model state)
        n_global = parallel.allreduce_int(n) if (sharded and model.num_data is not None) else n
        scale = 1.0 if model.num_data is None else float(model.num_data) / n_global
        dev = Xs.device
        q_mu = _device.to_device(model.q_mu.numpy(), ndim=1).reshape(-1)
        q_sqrt = _device.to_device(model.q_sqrt.numpy(), ndim=1).reshape(-1)
        Kuu = _device.gram(spec, pz)
        Kuu.diagonal().add_(DEFAULT_JITTER)
        L = torch.linalg.cholesky(Kuu)
        zpar = _inducing_parameter(model)
        z_train = want_grad and zpar is not None and zpar.trainable
        if want_grad:
            nout = int(_cabi.load().oak_backward_grad_count(spec.handle))
            grad = torch.zeros(nout, dtype=torch.float64, device=dev)
            gZ = torch.zeros((m, spec.num_dims), dtype=torch.float64, device=dev) if z_train else None
            G_AA = torch.zeros((m, m), dtype=torch.float64, device=dev)
            g_qmu = torch.zeros(m, dtype=torch.float64, device=dev)
            g_qsqrt = torch.zeros(m, dtype=torch.float64, device=dev)
        ve_sum = torch.zeros((), dtype=torch.float64, device=dev)
        chunk = max(64, int(model.chunk))
        for c0 in range(0, n, chunk):
            c1 = min(c0 + chunk, n)
            pxc = _device.Points(spec, Xs[c0:c1])
            A = torch.linalg.solve_triangular(L, _device.gram(spec, pz, pxc), upper=False).contiguous()
            mean, var = _device.svgp_moments(A, q_mu, q_sqrt, _device.gram_diag(spec, pxc))
            q = _device.bernoulli_quadrature(mean, var, y[c0:c1].contiguous(), lk.invlink.kind, lk.invlink.jitter,
                                             lk.num_gauss_hermite_points,
                                             want=("varexp", "gmean", "gvar") if want_grad else ("varexp",))
            ve_sum += q["varexp"].sum()
            if not want_grad:
                continue
            gm, gv = q["gmean"].mul_(scale), q["gvar"].mul_(scale)
            Abar = _device.svgp_moments_backward(A, q_mu, q_sqrt, gm, gv, g_qsqrt)
            g_qmu += A @ gm
            G_AA += Abar @ A.T
            Wc = torch.linalg.solve_triangular(L.T, Abar, upper=True).contiguous()  # L^-T Abar = d/dKuf
            if z_train:
                _device.gram_backward_rows(spec, pz, Wc, px2=pxc, grad=grad, grad_rows=gZ)
            else:
                _device.gram_backward(spec, pz, Wc, px2=pxc, grad=grad)
            _device.gram_diag_backward(spec, pxc, w=gv, grad=grad)
        if sharded:
            # one all-reduce of every data sum: variational expectation | d/dq_mu | d/dq_sqrt | A-bar A^T | kernel
            # gradients | row-point gradients (the KL term and the Kuu chain are replicated)
            parts = [ve_sum.reshape(1)]
            if want_grad:
                parts += [g_qmu, g_qsqrt, G_AA.reshape(-1), grad] + ([gZ.reshape(-1)] if z_train else [])
            packed = torch.cat(parts)
            parallel.allreduce_sum_(packed)
            off = 0

            def take(t):
                nonlocal off
                t.copy_(packed[off: off + t.numel()].view_as(t))
                off += t.numel()

            ve_sum = packed[0].clone()
            off = 1
            if want_grad:
                for t in [g_qmu, g_qsqrt, G_AA, grad] + ([gZ] if z_train else []):
                    take(t)
        kl = 0.5 * float((q_mu * q_mu).sum() - m - torch.log(q_sqrt * q_sqrt).sum() + (q_sqrt * q_sqrt).sum())
        elbo = scale * float(ve_sum) - kl
        if not want_grad:
            return elbo, None, None, None
        g_qmu -= q_mu
        g_qsqrt -= q_sqrt - 1.0 / q_sqrt
        Lbar = -torch.tril(torch.linalg.solve_triangular(L.T, G_AA, upper=True))
        P = torch.tril(L.T @ Lbar)
        P.diagonal().mul_(0.5)
        T = torch.linalg.solve_triangular(L.T, P + P.T, upper=True)             # L^-T (P + P^T)
        G_uu = 0.5 * torch.linalg.solve_triangular(L.T, T.T, upper=True).T     # ... L^-1
        G_uu = (0.5 * (G_uu + G_uu.T)).contiguous()
        if z_train:
            g_uu = torch.zeros_like(grad)
            _device.gram_backward_rows(spec, pz, (2.0 * G_uu).contiguous(), grad=g_uu, grad_rows=gZ)
            grad.add_(g_uu, alpha=0.5)
            gz_full = np.zeros(zpar.numpy().shape, dtype=np.float64)
            gz_host = gZ.cpu().numpy()
            for i, d in enumerate(spec._keep):
                gz_full[:, d.column] += gz_host[:, i]
            model._inducing_grad = gz_full
        else:
            _device.gram_backward(spec, pz, G_uu, grad=grad)
        g = grad.cpu().numpy()
        layout = [_device.table_layout(spec, i) for i in range(spec.num_dims)]
    finally:
        spec.close()
    D, P1 = spec.num_dims, max(spec.depth, 1) + 1
    model._table_cotangent = (g[D + P1: len(g) - D].copy(), layout)
    model._base_variance_grad = g[len(g) - D:].copy()
    model._variational_grads = {id(model.q_mu): g_qmu.cpu().numpy().reshape(-1, 1),
                                id(model.q_sqrt): g_qsqrt.cpu().numpy().reshape(-1, 1)}
    return elbo, g[:D].copy(), g[D: D + P1].copy(), None


def discrete_parameter_gradients(model) -> Dict[int, np.ndarray]:
    """Chains the table-blob cotangent of the last ``*_and_grad`` call to the categorical W / kappa /
    variance and binary variance Parameters (ortho_categorical_kernel.py:34-53, ortho_binary_kernel.py:29-38):
    a few C x C operations, differentiated with torch autograd on the host.  {id(Parameter): gradient}."""
    import torch

    g_tab, layout = model._table_cotangent
    out: Dict[int, np.ndarray] = {}
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    for i, k in enumerate(model.kernel.kernels):
        off, C_ = layout[i]
        if C_ == 0:
            continue
        GB = t(g_tab[off: off + C_ * C_].reshape(C_, C_))
        Gd = t(g_tab[off + C_ * C_: off + C_ * C_ + C_])
        var_p = getattr(k, "variance", None)
        var_t = t(scalar_of(var_p) if var_p is not None else 1.0).clone().requires_grad_(True)
        if hasattr(k, "kappa"):  # categorical
            W_t = t(value_of(k.W)).clone().requires_grad_(True)
            kap_t = t(value_of(k.kappa)).reshape(-1).clone().requires_grad_(True)
            p = t(k._p_vector()).reshape(-1, 1)
            A = W_t @ W_t.T + torch.diag(kap_t)
            Ap = A @ p
            pAp = (p.T @ Ap)[0, 0]
            B = (A - (Ap @ Ap.T) / pAp) * var_t
            Bd = ((W_t ** 2).sum(1) + kap_t - Ap[:, 0] ** 2 / pAp) * var_t
            ((GB * B).sum() + (Gd * Bd).sum()).backward()
            if isinstance(k.W, Parameter):
                out[id(k.W)] = W_t.grad.numpy().reshape(k.W.numpy().shape)
            if isinstance(k.kappa, Parameter):
                out[id(k.kappa)] = kap_t.grad.numpy().reshape(k.kappa.numpy().shape)
        else:  # binary
            p0 = float(k.p0)
            p1 = 1.0 - p0
            B1 = t([[p1 * p1, -p0 * p1], [-p0 * p1, p0 * p0]])
            ((GB * B1 * var_t).sum() + (Gd * torch.diagonal(B1) * var_t).sum()).backward()
        if isinstance(var_p, Parameter):
            out[id(var_p)] = np.full(var_p.numpy().shape, float(var_t.grad))
    return out


# ---- training loss on the unconstrained variables (gpflow's trainable_variables) -----------------
def trainable_parameters(model) -> List[Parameter]:
    return [p for p in collect_parameters(model) if p.trainable]


def training_loss_and_grad(model, data=None) -> Tuple[float, np.ndarray]:
    """-(objective + log prior) and its gradient w.r.t. the concatenated unconstrained trainables
    (``data``: the (X, Y) an SVGP's ``training_loss_closure(data)`` was given)."""
    from .models import SGPR, SVGP

    if isinstance(model, SVGP):
        data = data if data is not None else model.data
        if data is None:
            raise ValueError("an SVGP holds no data: pass (X, Y) or use model.training_loss_closure((X, Y))")
        val, g_ls, g_var, g_noise = svgp_elbo_and_grad(model, data)
    elif isinstance(model, SGPR):
        val, g_ls, g_var, g_noise = sgpr_elbo_and_grad(model)
    else:
        val, g_ls, g_var, g_noise = gpr_lml_and_grad(model)
    ls, var, noise = _supported_parameters(model)
    cgrad: Dict[int, np.ndarray] = {}
    for p, g in zip(ls, g_ls):
        if p is not None:
            cgrad[id(p)] = np.full(p.numpy().shape, g, dtype=np.float64)
    for p, g in zip(var, g_var):
        cgrad[id(p)] = np.full(p.numpy().shape, g, dtype=np.float64)
    if noise is not None:
        cgrad[id(noise)] = np.full(noise.numpy().shape, g_noise, dtype=np.float64)
    if _variational_parameters(model):
        cgrad.update(model._variational_grads)
    for k, g in zip(model.kernel.kernels, model._base_variance_grad):
        p = _base_variance_parameter(k)
        if p is not None:
            cgrad[id(p)] = np.full(p.numpy().shape, g, dtype=np.float64)
    if any(p.trainable for p in _discrete_parameters(model)):
        cgrad.update(discrete_parameter_gradients(model))
    zpar = _inducing_parameter(model)
    if zpar is not None and zpar.trainable:
        cgrad[id(zpar)] = model._inducing_grad
    loss = -(val + model.log_prior_density())
    parts = []
    for p in trainable_parameters(model):
        gc = cgrad[id(p)].copy()
        if p.prior is not None:
            gc = gc + _prior_grad(p)
        parts.append((-(gc) * _transform_grad(p)).reshape(-1))
    return float(loss), (np.concatenate(parts) if parts else np.zeros(0))


def _assign_unconstrained(params: List[Parameter], u: np.ndarray):
    off = 0
    for p in params:
        k = int(np.prod(p.unconstrained_variable.shape)) if np.ndim(p.unconstrained_variable) else 1
        p.unconstrained_variable = np.asarray(u[off: off + k], dtype=np.float64).reshape(np.shape(p.unconstrained_variable))
        off += k


def optimise(model, method: str = "BFGS", maxiter: int = 1000, data=None, **options):
    """``gpflow.optimizers.Scipy().minimize(model.training_loss_closure(), model.trainable_variables,
    method="BFGS")`` (oak/model_utils.py:168-175, 410-427) on the unconstrained variables.  ``model`` may be
    an ``SVGP.training_loss_closure(data)``, or an SVGP together with ``data``
    (examples/uci/uci_classification_train.py:119-124)."""
    if callable(model) and hasattr(model, "model"):
        model, data = model.model, model.data
    from scipy.optimize import minimize

    params = trainable_parameters(model)
    u0 = np.concatenate([np.asarray(p.unconstrained_variable, dtype=np.float64).reshape(-1) for p in params])

    state = {"finite": 0, "rejected": 0}

    def fun(u):
        _assign_unconstrained(params, u)
        try:
            out = training_loss_and_grad(model, data)
        except (RuntimeError, _cabi.OakNativeError) as exc:
            # a line-search TRIAL point whose Kuu / K + noise I is not numerically positive definite: report a
            # huge loss so that the step is shortened.  At the initial point (no finite loss seen yet) this is an
            # error of the model, not of the step: re-raise, as gpflow's Scipy wrapper would.
            if "positive" not in str(exc) and "Cholesky" not in str(exc):
                raise
            if state["finite"] == 0:
                raise
            state["rejected"] += 1
            return 1e25, np.zeros_like(u)
        if np.isfinite(out[0]):
            state["finite"] += 1
        return out

    res = minimize(fun, u0, jac=True, method=method, options=dict(maxiter=maxiter, **options))
    _assign_unconstrained(params, res.x)
    if state["rejected"]:
        import warnings

        warnings.warn(f"optimise: {state['rejected']} line-search trial point(s) rejected (matrix not positive definite)")
    res.rejected_trial_points = state["rejected"]
    return res
