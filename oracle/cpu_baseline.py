"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the timed CPU comparator.

A torch-CPU FP64 twin of ``oracle/oak_oracle.py`` with the reference's *unfused* structure:
one dense N x N2 matrix per input dimension kept alive (oak/oak_kernel.py:252-254), gpflow-style
expanded-distance RBF via matmul + outer-product correction (oak/ortho_rbf_kernel.py:163-172),
``pow``-based power sums and Newton-Girard on whole matrices (oak/oak_kernel.py:236-248),
variance-weighted sum (:256-260), gpflow 2.2.1 ``SGPR.elbo``.  TensorFlow
runs these as multi-threaded Eigen element-wise kernels; torch-CPU (MKL + OpenMP) on all host
cores is the closest stand-in available in this image.  Used by ``bench.py`` (``cpu_baseline`` and
``--impl reference``) and checked against the NumPy oracle in ``tests/test_oracle_identities.py``.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def tune_allocator():
    """Keep large temporaries on the heap instead of mmap/munmap-ing every N x N matrix (glibc):
    measured 5x faster for the unfused op sequence -- the comparator gets the benefit."""
    try:
        import ctypes

        libc = ctypes.CDLL("libc.so.6")
        libc.mallopt(-3, 1 << 30)  # M_MMAP_THRESHOLD
        libc.mallopt(-1, ctypes.c_int(-1))  # M_TRIM_THRESHOLD
    except Exception:
        pass


def _rbf(x, x2, l, s2):
    xs = x / l
    ys = xs if x2 is None else x2 / l
    d = -2.0 * (xs @ ys.T)
    d = d + ((xs * xs).sum(-1)[:, None] + (ys * ys).sum(-1)[None, :])
    return s2 * torch.exp(-0.5 * d)


def _cov_var(dc, x):
    l, s2 = dc["lengthscale"], dc.get("variance", 1.0)
    m = dc.get("measure", ("gaussian", 0.0, 1.0))
    if m[0] == "gaussian":
        mu, var = m[1], m[2]
        c = s2 * l / math.sqrt(l * l + var) * torch.exp(-0.5 * ((x - mu) ** 2) / (l * l + var))
        return c, s2 * l / math.sqrt(l * l + 2 * var)
    if m[0] == "empirical":
        loc = torch.as_tensor(np.asarray(m[1], dtype=np.float64).reshape(-1, 1))
        w = torch.as_tensor(np.asarray(m[2], dtype=np.float64).reshape(-1, 1))
        return _rbf(x, loc, l, s2) @ w, float((w.T @ _rbf(loc, None, l, s2) @ w).squeeze())
    raise NotImplementedError(m[0])


def _discrete_table(dc):
    if dc["type"] == "binary":
        p0, p1 = dc["p0"], 1.0 - dc["p0"]
        return torch.tensor([[p1 * p1, -p0 * p1], [-p0 * p1, p0 * p0]], dtype=torch.float64) * dc.get("variance", 1.0)
    W = torch.as_tensor(np.asarray(dc["W"], dtype=np.float64))
    kappa = torch.as_tensor(np.asarray(dc["kappa"], dtype=np.float64))
    p = torch.as_tensor(np.asarray(dc["p"], dtype=np.float64).reshape(-1, 1))
    A = W @ W.T + torch.diag(kappa)
    Ap = A @ p
    return (A - (Ap @ Ap.T) / (p.T @ Ap)[0]) * dc.get("variance", 1.0)


def dim_matrix(dc, x, x2=None):
    if dc["type"] == "rbf":
        base = _rbf(x, x2, dc["lengthscale"], dc.get("variance", 1.0))
        if dc.get("measure", ("gaussian", 0.0, 1.0)) is None:
            return base
        c, v = _cov_var(dc, x)
        c2 = c if x2 is None else _cov_var(dc, x2)[0]
        return base - torch.tensordot(c, c2.T, 1) / v
    B = _discrete_table(dc)
    i = x[:, 0].to(torch.int64)
    j = i if x2 is None else x2[:, 0].to(torch.int64)
    return B[j].T[i]


def additive_terms(mats, depth):
    s = []
    for p in range(depth + 1):
        acc = torch.pow(mats[0], p)
        for k in mats[1:]:
            acc = acc + torch.pow(k, p)
        s.append(acc)
    e = [torch.ones_like(mats[0])]
    for n in range(1, depth + 1):
        acc = None
        for k in range(1, n + 1):
            term = ((-1) ** (k - 1)) * e[n - k] * s[k]
            acc = term if acc is None else acc + term
        e.append((1.0 / n) * acc)
    return e


def gram(cfg, X, X2=None):
    """OAKKernel.K with the reference's op sequence (torch CPU tensors in, tensor out)."""
    mats = [dim_matrix(dc, X[:, d : d + 1], None if X2 is None else X2[:, d : d + 1])
            for d, dc in enumerate(cfg["dims"])]
    e = additive_terms(mats, cfg["depth"])
    var = cfg["variances"]
    if cfg.get("share_var", True):
        out = var[0] * e[0]
        for s2, t in zip(var[1:], e[1:]):
            out = out + s2 * t
        return out
    out = var[0] * e[0]
    for t in e[1:]:
        out = out + t
    return out


def gram_diag(cfg, X):
    diags = []
    for d, dc in enumerate(cfg["dims"]):
        x = X[:, d : d + 1]
        if dc["type"] == "rbf":
            base = torch.full((X.shape[0],), float(dc.get("variance", 1.0)), dtype=torch.float64)
            if dc.get("measure", ("gaussian", 0.0, 1.0)) is None:
                diags.append(base)
            else:
                c, v = _cov_var(dc, x)
                diags.append(base - c[:, 0] ** 2 / v)
        else:
            diags.append(torch.diagonal(_discrete_table(dc))[x[:, 0].to(torch.int64)])
    e = additive_terms(diags, cfg["depth"])
    var = cfg["variances"]
    out = var[0] * e[0]
    for n, t in enumerate(e[1:], start=1):
        out = out + (var[n] if cfg.get("share_var", True) else 1.0) * t
    return out


def sgpr_elbo(cfg, X, Y, Z, noise, jitter=1e-6):
    """gpflow 2.2.1 SGPR.elbo with the unfused OAK kernel."""
    n, r = Y.shape
    m = Z.shape[0]
    kdiag = gram_diag(cfg, X)
    kuf = gram(cfg, Z, X)
    kuu = gram(cfg, Z) + jitter * torch.eye(m, dtype=torch.float64)
    L = torch.linalg.cholesky(kuu)
    sigma = math.sqrt(noise)
    A = torch.linalg.solve_triangular(L, kuf, upper=False) / sigma
    AAT = A @ A.T
    B = AAT + torch.eye(m, dtype=torch.float64)
    LB = torch.linalg.cholesky(B)
    Aerr = A @ Y
    c = torch.linalg.solve_triangular(LB, Aerr, upper=False) / sigma
    bound = -0.5 * n * r * math.log(2 * math.pi)
    bound += -r * torch.log(torch.diagonal(LB)).sum()
    bound -= 0.5 * n * r * math.log(noise)
    bound += -0.5 * (Y ** 2).sum() / noise
    bound += 0.5 * (c ** 2).sum()
    bound += -0.5 * r * kdiag.sum() / noise
    bound += 0.5 * r * torch.diagonal(AAT).sum()
    return float(bound)
