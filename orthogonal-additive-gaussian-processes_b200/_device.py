"""Device plumbing between the host classes and the C ABI.

PyTorch is used for device memory, streams and ``torch.distributed`` only; every arithmetic
operation of the hot path happens inside ``liboak_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _cabi
from ._cabi import OakNativeError, Spec, check


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise OakNativeError("torch sees no CUDA device: the OAK B200 kernels have no CPU fallback")
    return torch


def stream_ptr() -> int:
    return int(_torch().cuda.current_stream().cuda_stream)


def to_device(X, ndim: int = 2):
    """NumPy / array-like / torch tensor -> contiguous float64 CUDA tensor."""
    torch = _torch()
    if isinstance(X, torch.Tensor):
        t = X.to(device="cuda", dtype=torch.float64)
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(X, dtype=np.float64))).to("cuda", non_blocking=False)
    if ndim == 2 and t.ndim == 1:
        t = t.reshape(-1, 1)
    return t.contiguous()


def is_host(X) -> bool:
    try:
        import torch

        if isinstance(X, torch.Tensor):
            return not X.is_cuda
    except Exception:  # pragma: no cover
        pass
    return True


def from_device(t, like_host: bool):
    """Returns NumPy when the caller handed in host data, the CUDA tensor otherwise."""
    return t.cpu().numpy() if like_host else t


def _p(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class Points:
    """Prepared-point block of ``n`` points for one spec (see csrc/oak_prepare.cu)."""

    _dbuf = None

    def backward_block(self):
        """Per-point d c^/dl block for the backward tiles (empirical / uniform / MOG dims), built on first use."""
        if not any(d.type == _cabi.DIM_RBF and d.measure in (_cabi.MEASURE_EMPIRICAL, _cabi.MEASURE_UNIFORM,
                                                           _cabi.MEASURE_MOG) for d in self.spec._keep):
            return None  # closed forms only: the library takes NULL
        if self._dbuf is None:
            torch = _torch()
            lib = _cabi.load()
            nbytes = int(lib.oak_backward_points_bytes(self.spec.handle, self.n))
            self._dbuf = torch.zeros(max(nbytes // 8, 1), dtype=torch.float64, device=self.buf.device)
            if self.n > 0:
                check(lib.oak_prepare_backward_f64(self.spec.handle, _p(self.buf), self.n, _p(self._dbuf),
                                                   C.c_void_p(stream_ptr())), "oak_prepare_backward_f64")
        return self._dbuf

    def __init__(self, spec: Spec, Xd):
        torch = _torch()
        self.n = int(Xd.shape[0])
        self.spec = spec
        nbytes = spec.points_bytes(self.n)
        self.buf = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=Xd.device)
        if self.n > 0:
            check(
                _cabi.load().oak_prepare_points_f64(
                    spec.handle, _p(Xd), self.n, int(Xd.stride(0)), _p(self.buf), C.c_void_p(stream_ptr())
                ),
                "oak_prepare_points_f64",
            )


def validate_discrete(Xd, columns: Sequence[int], counts: Sequence[int]):
    """``tf.gather`` on CPU raises for out-of-range indices; do the same before the table gather.  One fused
    min / max over all discrete columns and ONE read-back (models remember that their resident training data
    passed, see ``GPModel._check_training_discrete``)."""
    if Xd.shape[0] == 0 or not len(columns):
        return
    torch = _torch()
    v = Xd[:, list(columns)].trunc()
    lohi = torch.stack((v.amin(0), v.amax(0))).cpu().numpy()
    for j, (col, cnt) in enumerate(zip(columns, counts)):
        lo, hi = float(lohi[0, j]), float(lohi[1, j])
        if not (lo >= 0 and hi <= cnt - 1):  # NaN fails too
            raise ValueError(f"column {col}: category index outside [0, {cnt - 1}] (got [{lo}, {hi}])")


def gram(spec: Spec, px: Points, px2: Optional[Points] = None, row_begin: int = 0,
         row_end: Optional[int] = None, out=None):
    torch = _torch()
    n = px.n
    n2 = n if px2 is None else px2.n
    row_end = n if row_end is None else row_end
    rows = row_end - row_begin
    if out is None:
        out = torch.empty((rows, n2), dtype=torch.float64, device=px.buf.device)
    if rows > 0 and n2 > 0:
        check(
            _cabi.load().oak_gram_f64(
                spec.handle, _p(px.buf), n, _p(None if px2 is None else px2.buf), n2, row_begin, row_end,
                _p(out), int(out.stride(0)), C.c_void_p(stream_ptr()),
            ),
            "oak_gram_f64",
        )
    return out


def gram_lower(spec: Spec, px: Points, row_begin: int, row_end: int, out=None):
    """Rows [row_begin,row_end) x cols [0,row_end) of the symmetric Gram, lower-triangle tiles only."""
    torch = _torch()
    rows = row_end - row_begin
    if out is None:
        out = torch.empty((rows, row_end), dtype=torch.float64, device=px.buf.device)
    if rows > 0:
        check(
            _cabi.load().oak_gram_lower_f64(spec.handle, _p(px.buf), px.n, row_begin, row_end, _p(out),
                                            int(out.stride(0)), C.c_void_p(stream_ptr())),
            "oak_gram_lower_f64",
        )
    return out


def gram_lower_mirror(spec: Spec, px: Points, row_begin: int, row_end: int, out=None, out_t=None):
    """``gram_lower`` plus the mirror image of the strip's off-diagonal tiles: returns (K[rows, :row_end],
    Kt[:row_end, rows]) -- a rank's share of the full symmetric matrix."""
    torch = _torch()
    rows = row_end - row_begin
    if out is None:
        out = torch.empty((rows, row_end), dtype=torch.float64, device=px.buf.device)
    if out_t is None:
        out_t = torch.empty((row_end, rows), dtype=torch.float64, device=px.buf.device)
    if rows > 0:
        check(
            _cabi.load().oak_gram_lower_mirror_f64(spec.handle, _p(px.buf), px.n, row_begin, row_end, _p(out),
                                                   int(out.stride(0)), _p(out_t), int(out_t.stride(0)),
                                                   C.c_void_p(stream_ptr())),
            "oak_gram_lower_mirror_f64",
        )
    return out, out_t


def gram_matvec(spec: Spec, px: Points, px2: Points, alpha, out=None):
    """out[i] = sum_j K(px_i, px2_j) alpha_j, fused (no N x N2 matrix)."""
    torch = _torch()
    if out is None:
        out = torch.empty((px.n,), dtype=torch.float64, device=px.buf.device)
    a = alpha.reshape(-1).contiguous()
    assert a.numel() == px2.n
    if px.n > 0:
        check(
            _cabi.load().oak_gram_matvec_f64(spec.handle, _p(px.buf), px.n, 0, px.n, _p(px2.buf), px2.n, _p(a),
                                             _p(out), C.c_void_p(stream_ptr())),
            "oak_gram_matvec_f64",
        )
    return out


def gram_diag(spec: Spec, px: Points):
    torch = _torch()
    out = torch.empty((px.n,), dtype=torch.float64, device=px.buf.device)
    if px.n > 0:
        check(
            _cabi.load().oak_gram_diag_f64(spec.handle, _p(px.buf), px.n, _p(out), C.c_void_p(stream_ptr())),
            "oak_gram_diag_f64",
        )
    return out


def component_gram(spec: Spec, subset: Sequence[int], px: Points, px2: Optional[Points] = None):
    torch = _torch()
    n = px.n
    n2 = n if px2 is None else px2.n
    out = torch.empty((n, n2), dtype=torch.float64, device=px.buf.device)
    arr = (C.c_int32 * max(len(subset), 1))(*[int(s) for s in subset])
    if n > 0 and n2 > 0:
        check(
            _cabi.load().oak_component_gram_f64(
                spec.handle, arr, len(subset), _p(px.buf), n, _p(None if px2 is None else px2.buf), n2,
                _p(out), int(out.stride(0)), C.c_void_p(stream_ptr()),
            ),
            "oak_component_gram_f64",
        )
    return out


def component_diag(spec: Spec, subset: Sequence[int], px: Points):
    torch = _torch()
    out = torch.empty((px.n,), dtype=torch.float64, device=px.buf.device)
    arr = (C.c_int32 * max(len(subset), 1))(*[int(s) for s in subset])
    if px.n > 0:
        check(
            _cabi.load().oak_component_diag_f64(spec.handle, arr, len(subset), _p(px.buf), px.n, _p(out),
                                                C.c_void_p(stream_ptr())),
            "oak_component_diag_f64",
        )
    return out


def additive_terms(mats, depth: int):
    """e_0..e_depth of a stack of equally shaped arrays (num_mats, ...) -> (depth+1, ...)."""
    torch = _torch()
    m = mats.contiguous()
    num = int(m.shape[0])
    length = int(m[0].numel())
    out = torch.empty((depth + 1,) + tuple(m.shape[1:]), dtype=torch.float64, device=m.device)
    check(
        _cabi.load().oak_additive_terms_f64(_p(m), num, int(depth), length, _p(out), C.c_void_p(stream_ptr())),
        "oak_additive_terms_f64",
    )
    return out


def component_predict(spec: Spec, subsets: Sequence[Sequence[int]], px: Points, pcond: Points, alpha):
    """out[c, i] = sigma^2_|S_c| sum_j prod_{d in S_c} k_d(x_i, z_j) alpha_j."""
    torch = _torch()
    nc = len(subsets)
    max_order = max([len(s) for s in subsets] + [1])
    tab = np.full((nc, max_order), -1, dtype=np.int32)
    for c, s in enumerate(subsets):
        tab[c, : len(s)] = s
    d_tab = torch.as_tensor(tab).to(px.buf.device)
    out = torch.empty((nc, px.n), dtype=torch.float64, device=px.buf.device)
    a = alpha.reshape(-1).contiguous()
    if nc > 0 and px.n > 0:
        check(
            _cabi.load().oak_component_predict_f64(
                spec.handle, _p(d_tab), nc, max_order, _p(px.buf), px.n, _p(pcond.buf), pcond.n, _p(a),
                _p(out), C.c_void_p(stream_ptr()),
            ),
            "oak_component_predict_f64",
        )
    return out


# ---- backward tiles ---------------------------------------------------------------------
def gram_backward(spec: Spec, px: Points, W, px2: Optional[Points] = None, row_begin: int = 0,
                  row_end: Optional[int] = None, grad=None):
    """grad[:D] += sum W dK/d lengthscale_i, grad[D:] += sum W e_n for K(px[row_begin:row_end], px2)."""
    torch = _torch()
    lib = _cabi.load()
    n = px.n
    n2 = n if px2 is None else px2.n
    row_end = n if row_end is None else row_end
    nout = int(lib.oak_backward_grad_count(spec.handle))
    if grad is None:
        grad = torch.zeros(nout, dtype=torch.float64, device=px.buf.device)
    if grad.numel() != nout or not grad.is_contiguous():
        raise ValueError(f"gradient buffer must hold oak_backward_grad_count = {nout} contiguous doubles")
    if row_end > row_begin and n2 > 0:
        assert W.shape == (row_end - row_begin, n2) and W.stride(1) == 1
        work = torch.empty(max(int(lib.oak_gram_backward_work_bytes(spec.handle, max(n, n2))) // 8, 1),
                           dtype=torch.float64, device=px.buf.device)
        check(
            lib.oak_gram_backward_f64(spec.handle, _p(px.buf), _p(px.backward_block()), n, row_begin, row_end,
                                      _p(None if px2 is None else px2.buf),
                                      _p(None if px2 is None else px2.backward_block()), n2, _p(W),
                                      int(W.stride(0)), _p(grad), _p(work), C.c_void_p(stream_ptr())),
            "oak_gram_backward_f64",
        )
    return grad


def gram_backward_rows(spec: Spec, px: Points, W, px2: Optional[Points] = None, grad=None, grad_rows=None):
    """``gram_backward`` over all rows of ``px`` plus ``grad_rows[i, k] += sum_j W_ij dK(x_i, y_j)/dx_{i,k}``
    (k = sub-kernel index): the gradient with respect to the row points (inducing points)."""
    torch = _torch()
    lib = _cabi.load()
    n = px.n
    n2 = n if px2 is None else px2.n
    nout = int(lib.oak_backward_grad_count(spec.handle))
    if grad is None:
        grad = torch.zeros(nout, dtype=torch.float64, device=px.buf.device)
    if grad_rows is None:
        grad_rows = torch.zeros((n, spec.num_dims), dtype=torch.float64, device=px.buf.device)
    if grad.numel() != nout or not grad.is_contiguous():
        raise ValueError(f"gradient buffer must hold oak_backward_grad_count = {nout} contiguous doubles")
    if grad_rows.shape != (n, spec.num_dims) or grad_rows.stride(1) != 1 or grad_rows.dtype != torch.float64:
        raise ValueError("row-gradient buffer must be (n, num_dims) float64 with unit column stride")
    if n > 0 and n2 > 0:
        assert W.shape == (n, n2) and W.stride(1) == 1
        work = torch.empty(max(int(lib.oak_gram_backward_rows_work_bytes(spec.handle, n, n2)) // 8, 1),
                           dtype=torch.float64, device=px.buf.device)
        check(
            lib.oak_gram_backward_rows_f64(spec.handle, _p(px.buf), _p(px.backward_block()), n,
                                           _p(None if px2 is None else px2.buf),
                                           _p(None if px2 is None else px2.backward_block()), n2, _p(W),
                                           int(W.stride(0)), _p(grad), _p(grad_rows), int(grad_rows.stride(0)),
                                           _p(work), C.c_void_p(stream_ptr())),
            "oak_gram_backward_rows_f64",
        )
    return grad, grad_rows


def table_layout(spec: Spec, dim: int):
    """(offset, C) of sub-kernel ``dim``'s table inside the table-blob part of the gradient vector."""
    off, cnt = C.c_int32(0), C.c_int32(0)
    check(_cabi.load().oak_spec_table_layout(spec.handle, int(dim), C.byref(off), C.byref(cnt)), "oak_spec_table_layout")
    return int(off.value), int(cnt.value)


def gram_diag_backward(spec: Spec, px: Points, wscale: float = 1.0, w=None, grad=None):
    """grad += d/d theta of wscale * sum_i w_i K_diag(x_i)."""
    torch = _torch()
    lib = _cabi.load()
    nout = int(lib.oak_backward_grad_count(spec.handle))
    if grad is None:
        grad = torch.zeros(nout, dtype=torch.float64, device=px.buf.device)
    if grad.numel() != nout or not grad.is_contiguous():
        raise ValueError(f"gradient buffer must hold oak_backward_grad_count = {nout} contiguous doubles")
    if px.n > 0:
        work = torch.empty(max(int(lib.oak_gram_backward_work_bytes(spec.handle, px.n)) // 8, 1),
                           dtype=torch.float64, device=px.buf.device)
        check(
            lib.oak_gram_diag_backward_f64(spec.handle, _p(px.buf), _p(px.backward_block()), px.n, _p(w),
                                           float(wscale), _p(grad), _p(work), C.c_void_p(stream_ptr())),
            "oak_gram_diag_backward_f64",
        )
    return grad


# ---- SGPR / GPR -------------------------------------------------------------------------
# ---- input pipeline: normalising flow -------------------------------------------------------------
def flow_forward(x, offset: float, use_log: bool, shift: float, scale: float, skewness: float, tailweight: float,
                 out=None):
    """Transforms one (possibly strided) device column; ``out`` defaults to a new contiguous vector."""
    torch = _torch()
    assert x.ndim == 1 and x.dtype == torch.float64
    n = x.numel()
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=x.device)
    assert out.ndim == 1 and out.numel() == n
    if n > 0:
        check(_cabi.load().oak_flow_forward_f64(_p(x), n, int(x.stride(0)), float(offset), int(bool(use_log)),
                                                float(shift), float(scale), float(skewness), float(tailweight),
                                                _p(out), int(out.stride(0)), C.c_void_p(stream_ptr())),
              "oak_flow_forward_f64")
    return out


def flow_objective(x, offset: float, use_log: bool, theta, work=None):
    """[J, dJ/d theta] of the KL objective for theta = (log scale, shift, skewness, log tailweight): NumPy (5,)."""
    torch = _torch()
    lib = _cabi.load()
    n = x.numel()
    if work is None:
        work = torch.empty(max(int(lib.oak_flow_objective_work_bytes(n)) // 8, 1), dtype=torch.float64, device=x.device)
    out = torch.empty(5, dtype=torch.float64, device=x.device)
    check(lib.oak_flow_objective_f64(_p(x), n, int(x.stride(0)), float(offset), int(bool(use_log)), float(theta[0]),
                                     float(theta[1]), float(theta[2]), float(theta[3]), _p(out), _p(work),
                                     C.c_void_p(stream_ptr())), "oak_flow_objective_f64")
    return out.cpu().numpy()


def cuda_available() -> bool:
    """True on a machine with a CUDA device.  On such a machine a missing or broken ``liboak_b200.so`` is an ERROR
    (``_cabi.load`` raises): the host-side preprocessing must never drift onto a CPU path on a GPU box."""
    try:
        import torch

        if not torch.cuda.is_available():
            return False
    except ImportError:
        return False
    return _cabi.load().oak_device_count() > 0


def column_unique(Xd, col: int):
    """``np.unique(X[:, col], return_counts=True)`` of a device matrix: (values, counts) as NumPy arrays."""
    torch = _torch()
    lib = _cabi.load()
    assert Xd.ndim == 2 and Xd.dtype == torch.float64 and Xd.stride(1) == 1
    n = int(Xd.shape[0])
    vals = torch.empty(max(n, 1), dtype=torch.float64, device=Xd.device)
    counts = torch.empty(max(n, 1), dtype=torch.int64, device=Xd.device)
    num = torch.zeros(1, dtype=torch.int32, device=Xd.device)
    work = torch.empty(max(int(lib.oak_column_unique_work_bytes(n)) // 8 + 1, 1), dtype=torch.float64, device=Xd.device)
    check(lib.oak_column_unique_f64(_p(Xd), n, int(Xd.stride(0)), int(col), _p(vals), _p(counts), _p(num), _p(work),
                                    C.c_void_p(stream_ptr())), "oak_column_unique_f64")
    u = int(num.item())
    return vals[:u].cpu().numpy(), counts[:u].cpu().numpy()


def column_mean(Xd, col: int) -> float:
    torch = _torch()
    n = int(Xd.shape[0])
    out = torch.empty(1, dtype=torch.float64, device=Xd.device)
    work = torch.empty(n // 4096 + 2, dtype=torch.float64, device=Xd.device)
    check(_cabi.load().oak_column_mean_f64(_p(Xd), n, int(Xd.stride(0)), int(col), _p(out), _p(work),
                                           C.c_void_p(stream_ptr())), "oak_column_mean_f64")
    return float(out.item())


# ---- whitened SVGP / Bernoulli pieces ------------------------------------------------------------
_GH_CACHE = {}


def gauss_hermite(n_gh: int, device):
    """(x sqrt 2, w / sqrt pi) of ``np.polynomial.hermite.hermgauss`` on the device -- gpflow's own source of
    the nodes (gpflow/quadrature/gauss_hermite.py gh_points_and_weights)."""
    key = (int(n_gh), str(device))
    if key not in _GH_CACHE:
        torch = _torch()
        x, w = np.polynomial.hermite.hermgauss(int(n_gh))
        _GH_CACHE[key] = (torch.as_tensor(x * np.sqrt(2.0), dtype=torch.float64).to(device),
                          torch.as_tensor(w / np.sqrt(np.pi), dtype=torch.float64).to(device))
    return _GH_CACHE[key]


def svgp_moments(A, q_mu, q_sqrt, kdiag):
    """mean_i = A[:, i] . q_mu, var_i = kdiag_i - sum_r A_ri^2 (1 - q_sqrt_r^2) for A = L^-1 Kuf (M x n)."""
    torch = _torch()
    m, n = A.shape
    assert A.stride(1) == 1 and q_mu.numel() == m and q_sqrt.numel() == m and kdiag.numel() == n
    mean = torch.empty(n, dtype=torch.float64, device=A.device)
    var = torch.empty(n, dtype=torch.float64, device=A.device)
    check(_cabi.load().oak_svgp_moments_f64(_p(A), int(A.stride(0)), m, n, _p(q_mu), _p(q_sqrt), _p(kdiag), _p(mean),
                                            _p(var), C.c_void_p(stream_ptr())), "oak_svgp_moments_f64")
    return mean, var


def svgp_moments_backward(A, q_mu, q_sqrt, gmean, gvar, gq_sqrt):
    """Abar = q_mu gmean^T - 2 (1 - q_sqrt^2) A diag(gvar); gq_sqrt += 2 q_sqrt sum_i gvar_i A_ri^2."""
    torch = _torch()
    m, n = A.shape
    Abar = torch.empty((m, n), dtype=torch.float64, device=A.device)
    check(_cabi.load().oak_svgp_moments_backward_f64(_p(A), int(A.stride(0)), m, n, _p(q_mu), _p(q_sqrt), _p(gmean),
                                                     _p(gvar), _p(Abar), int(Abar.stride(0)), _p(gq_sqrt),
                                                     C.c_void_p(stream_ptr())), "oak_svgp_moments_backward_f64")
    return Abar


def bernoulli_quadrature(mean, var, y, link: int, jitter: float, n_gh: int, want=("varexp", "gmean", "gvar")):
    """Gauss-Hermite expectations of the Bernoulli log density; returns a dict of the requested arrays out of
    ("varexp", "gmean", "gvar", "logdensity")."""
    torch = _torch()
    n = mean.numel()
    gx, gw = gauss_hermite(n_gh, mean.device)
    out = {k: torch.empty(n, dtype=torch.float64, device=mean.device) for k in want}
    check(_cabi.load().oak_bernoulli_quadrature_f64(
        _p(mean), _p(var), _p(y), n, int(link), float(jitter), _p(gx), _p(gw), int(n_gh), _p(out.get("varexp")),
        _p(out.get("gmean")), _p(out.get("gvar")), _p(out.get("logdensity")), C.c_void_p(stream_ptr())),
        "oak_bernoulli_quadrature_f64")
    return out


def sgpr_stats(spec: Spec, pz: Points, px: Points, y, chunk: int = 262144, stats=None, keep_kuf: bool = False,
               kuf_store=None):
    """Accumulates Phi | Kuf y | sum K_diag | y^T y for the local points into ``stats``.
    ``keep_kuf``: also returns the list of per-chunk Kuf blocks (M x nc views) for a backward pass and
    the (nchunks, M, chunk) buffer holding them (pass it back as ``kuf_store`` to reuse the allocation)."""
    torch = _torch()
    lib = _cabi.load()
    m = pz.n
    count = int(lib.oak_sgpr_stats_count(m))
    if stats is None:
        stats = torch.zeros(count, dtype=torch.float64, device=pz.buf.device)
    chunk = int(max(64, min(chunk, max(px.n, 64))))
    chunk = (chunk + 63) // 64 * 64
    work = torch.empty(max(int(lib.oak_sgpr_stats_work_bytes(m, chunk)) // 8, 1), dtype=torch.float64,
                       device=pz.buf.device)
    yv = y.reshape(-1).contiguous()
    if keep_kuf:
        nchunks = max((px.n + chunk - 1) // chunk, 1)
        store = kuf_store
        if store is None or tuple(store.shape) != (nchunks, m, chunk) or store.device != pz.buf.device:
            store = torch.empty((nchunks, m, chunk), dtype=torch.float64, device=pz.buf.device)
        check(
            lib.oak_sgpr_stats_keep_f64(spec.handle, _p(pz.buf), m, _p(px.buf), _p(yv), px.n, chunk, _p(stats),
                                        _p(work), _p(store), C.c_void_p(stream_ptr())),
            "oak_sgpr_stats_keep_f64",
        )
        blocks = [store[c, :, : min(chunk, px.n - c * chunk)] for c in range(nchunks) if px.n - c * chunk > 0]
        return stats, blocks, chunk, store
    check(
        lib.oak_sgpr_stats_f64(spec.handle, _p(pz.buf), m, _p(px.buf), _p(yv), px.n, chunk, _p(stats),
                               _p(work), C.c_void_p(stream_ptr())),
        "oak_sgpr_stats_f64",
    )
    return stats


def sgpr_finish(Kuu, stats, n_total: int, noise: float, jitter: float, want_alpha=True):
    """Returns (out[4] = elbo, sum log diag LB, tr(AAT), c^T c ; alpha[M] or None). Overwrites inputs."""
    torch = _torch()
    lib = _cabi.load()
    m = int(Kuu.shape[0])
    out = torch.empty(4, dtype=torch.float64, device=Kuu.device)
    alpha = torch.empty(m, dtype=torch.float64, device=Kuu.device) if want_alpha else None
    work = torch.empty(max(int(lib.oak_sgpr_finish_work_bytes(m)) // 8, 1), dtype=torch.float64, device=Kuu.device)
    check(
        lib.oak_sgpr_finish_f64(_p(Kuu), _p(stats), m, int(n_total), float(noise), float(jitter), _p(out),
                                _p(alpha), _p(work), C.c_void_p(stream_ptr())),
        "oak_sgpr_finish_f64",
    )
    return out, alpha


# ---- SGPR, factor-first path (L = chol(Kuu) before the statistics; route chosen on the device) ----
ROUTE_AUTO, ROUTE_PHI, ROUTE_WHITENED = -1, 0, 1


class KuuFactor:
    """[L ; L^-T] of Kuu + jitter I with its header (see ``oak_sgpr_factor_f64`` in include/oak_b200.h)."""

    def __init__(self, buf, m: int):
        lib = _cabi.load()
        self.buf, self.m = buf, int(m)
        self.ld = int(lib.oak_sgpr_factor_ld(m))

    @property
    def _mat(self):
        return self.buf[: self.ld * self.m].view(self.m, self.ld)  # row r of this view = column r of the factor

    def L(self):
        """Lower-triangular factor as a (M, M) tensor (a copy)."""
        return _torch().tril(self._mat[:, : self.m].T)

    def Linv(self):
        """L^-1 (row-major view, exact zeros above the diagonal)."""
        return self._mat[:, self.ld // 2: self.ld // 2 + self.m]

    def header(self):
        """16 doubles: cond estimate, ||Kuu||_1, lambda_max(Kuu^-1) estimate, route, info, sum log diag L, threshold."""
        return self.buf[self.ld * self.m: self.ld * self.m + 16]


def sgpr_factor(spec: Spec, pz: Points, jitter: float, route: int = ROUTE_AUTO, cond_threshold: float = 0.0,
                buf=None) -> KuuFactor:
    torch = _torch()
    lib = _cabi.load()
    m = pz.n
    cnt = int(lib.oak_sgpr_factor_count(m))
    if buf is None or buf.numel() != cnt:
        buf = torch.empty(cnt, dtype=torch.float64, device=pz.buf.device)
    check(lib.oak_sgpr_factor_f64(spec.handle, _p(pz.buf), m, float(jitter), int(route), float(cond_threshold),
                                  _p(buf), C.c_void_p(stream_ptr())), "oak_sgpr_factor_f64")
    return KuuFactor(buf, m)


def sgpr_stats2(spec: Spec, pz: Points, px: Points, y, fac: KuuFactor, chunk: int = 262144, stats=None,
                keep_kuf: bool = False, kuf_store=None):
    """``sgpr_stats`` on the route stored in ``fac``: the first M*M entries are Phi (route 0) or Psi (route 1)."""
    torch = _torch()
    lib = _cabi.load()
    m = pz.n
    count = int(lib.oak_sgpr_stats_count(m))
    if stats is None:
        stats = torch.zeros(count, dtype=torch.float64, device=pz.buf.device)
    chunk = int(max(64, min(chunk, max(px.n, 64))))
    chunk = (chunk + 63) // 64 * 64
    work = torch.empty(max(int(lib.oak_sgpr_stats2_work_bytes(m, chunk)) // 8, 1), dtype=torch.float64,
                       device=pz.buf.device)
    yv = y.reshape(-1).contiguous()
    store = None
    if keep_kuf:
        nchunks = max((px.n + chunk - 1) // chunk, 1)
        store = kuf_store
        if store is None or tuple(store.shape) != (nchunks, m, chunk) or store.device != pz.buf.device:
            store = torch.empty((nchunks, m, chunk), dtype=torch.float64, device=pz.buf.device)
    check(lib.oak_sgpr_stats2_f64(spec.handle, _p(pz.buf), m, _p(fac.buf), _p(px.buf), _p(yv), px.n, chunk, _p(stats),
                                  _p(work), _p(store), C.c_void_p(stream_ptr())), "oak_sgpr_stats2_f64")
    if keep_kuf:
        blocks = [store[c, :, : min(chunk, px.n - c * chunk)] for c in range(nchunks) if px.n - c * chunk > 0]
        return stats, blocks, chunk, store
    return stats


def sgpr_factor_stats(spec: Spec, pz: Points, px: Points, y, jitter: float, route: int = ROUTE_AUTO,
                      cond_threshold: float = 0.0, chunk: int = 262144, stats=None, keep_kuf: bool = False,
                      kuf_store=None, overlap_ctas: int = -1, buf=None):
    """``sgpr_factor`` + ``sgpr_stats2`` in one call (``oak_sgpr_factor_stats_f64``): the factorisation of Kuu runs
    on a side stream next to the first chunk's Kuf tiles.  Returns (factor, stats) or, with ``keep_kuf``,
    (factor, stats, blocks, chunk, store)."""
    torch = _torch()
    lib = _cabi.load()
    m = pz.n
    dev = pz.buf.device
    cnt = int(lib.oak_sgpr_factor_count(m))
    if buf is None or buf.numel() != cnt:
        buf = torch.empty(cnt, dtype=torch.float64, device=dev)
    if stats is None:
        stats = torch.zeros(int(lib.oak_sgpr_stats_count(m)), dtype=torch.float64, device=dev)
    chunk = int(max(64, min(chunk, max(px.n, 64))))
    chunk = (chunk + 63) // 64 * 64
    work = torch.empty(max(int(lib.oak_sgpr_stats2_work_bytes(m, chunk)) // 8, 1), dtype=torch.float64, device=dev)
    yv = y.reshape(-1).contiguous()
    store = None
    nchunks = max((px.n + chunk - 1) // chunk, 1)
    if keep_kuf:
        store = kuf_store
        if store is None or tuple(store.shape) != (nchunks, m, chunk) or store.device != dev:
            store = torch.empty((nchunks, m, chunk), dtype=torch.float64, device=dev)
    check(lib.oak_sgpr_factor_stats_f64(spec.handle, _p(pz.buf), m, float(jitter), int(route), float(cond_threshold),
                                        _p(buf), _p(px.buf), _p(yv), px.n, chunk, _p(stats), _p(work), _p(store),
                                        int(overlap_ctas), C.c_void_p(stream_ptr())), "oak_sgpr_factor_stats_f64")
    fac = KuuFactor(buf, m)
    if keep_kuf:
        blocks = [store[c, :, : min(chunk, px.n - c * chunk)] for c in range(nchunks) if px.n - c * chunk > 0]
        return fac, stats, blocks, chunk, store
    return fac, stats


class SgprTail:
    """Result of ``sgpr_finish2``: ``out`` (8 doubles on the device), ``alpha`` and the factor of B."""

    def __init__(self, out, alpha, lb_buf, m):
        self.out, self.alpha, self.lb_buf, self.m = out, alpha, lb_buf, int(m)
        self.ldb = int(_cabi.load().oak_sgpr_lb_ld(m))

    def LB(self):
        return _torch().tril(self.lb_buf.view(self.m, self.ldb)[:, : self.m].T)

    def c(self):
        mp = (self.m + 7) // 8 * 8
        return self.lb_buf.view(self.m, self.ldb)[:, mp]

    def host(self):
        """The one read-back of an evaluation; raises when a factorisation failed."""
        o = self.out.cpu().numpy()
        if o[4] != 0:
            raise OakNativeError(f"oak_sgpr: Cholesky of Kuu failed (leading minor {int(o[4])} not positive definite)")
        if o[5] != 0:
            raise OakNativeError(f"oak_sgpr: Cholesky of B = A A^T + I failed (leading minor {int(o[5])} not positive "
                                 "definite)")
        return o


def sgpr_finish2(fac: KuuFactor, stats, n_total: int, noise: float, want_alpha: bool = True) -> SgprTail:
    torch = _torch()
    lib = _cabi.load()
    m = fac.m
    dev = fac.buf.device
    out = torch.empty(8, dtype=torch.float64, device=dev)
    alpha = torch.empty(m, dtype=torch.float64, device=dev) if want_alpha else None
    lb = torch.empty(int(lib.oak_sgpr_lb_ld(m)) * m, dtype=torch.float64, device=dev)
    work = torch.empty(max(int(lib.oak_sgpr_finish2_work_bytes(m)) // 8, 1), dtype=torch.float64, device=dev)
    check(lib.oak_sgpr_finish2_f64(_p(fac.buf), _p(stats), m, int(n_total), float(noise), _p(out), _p(alpha), _p(lb),
                                   _p(work), C.c_void_p(stream_ptr())), "oak_sgpr_finish2_f64")
    return SgprTail(out, alpha, lb, m)


def chol(A, n: int, rows: int, gap: int = 0, border_identity: bool = False):
    """In-place bordered Cholesky of a column-major matrix held in the 2-D tensor ``A`` (shape (n, ld): row j of
    the tensor is column j of the matrix).  Returns (info tensor[1] int32, logdet tensor[1])."""
    torch = _torch()
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    logdet = torch.zeros(1, dtype=torch.float64, device=A.device)
    check(_cabi.load().oak_chol_f64(_p(A), int(n), int(rows), int(gap), int(A.stride(0)), int(bool(border_identity)),
                                    _p(info), _p(logdet), C.c_void_p(stream_ptr())), "oak_chol_f64")
    return info, logdet


def panel_gemm(T, B, lower: bool = False, u=None, v=None, out=None):
    """out (M x n) = T (M x Kd) @ B (Kd x n) [+ u v^T] on the FP64 tensor cores (row-major, even pitches)."""
    torch = _torch()
    M, Kd = T.shape
    n = B.shape[1]
    assert B.shape[0] == Kd and T.stride(1) == 1 and B.stride(1) == 1
    if out is None:
        out = torch.empty((M, (n + 1) // 2 * 2), dtype=torch.float64, device=T.device)[:, :n]
    work = torch.zeros(8, dtype=torch.float64, device=T.device)
    check(_cabi.load().oak_panel_gemm_f64(_p(T), int(T.stride(0)), _p(B), int(B.stride(0)), _p(out), int(out.stride(0)),
                                          M, Kd, n, int(bool(lower)), _p(u), _p(v), _p(work),
                                          C.c_void_p(stream_ptr())), "oak_panel_gemm_f64")
    return out


def gpr_finish(K, y, noise: float):
    """Returns (lml tensor[1], alpha[n]); K is overwritten by its Cholesky factor."""
    torch = _torch()
    lib = _cabi.load()
    n = int(K.shape[0])
    lml = torch.empty(1, dtype=torch.float64, device=K.device)
    alpha = torch.empty(n, dtype=torch.float64, device=K.device)
    work = torch.empty(max(int(lib.oak_gpr_finish_work_bytes(n)) // 8, 1), dtype=torch.float64, device=K.device)
    yv = y.reshape(-1).contiguous()
    check(
        lib.oak_gpr_finish_f64(_p(K), _p(yv), n, float(noise), _p(lml), _p(alpha), _p(work),
                               C.c_void_p(stream_ptr())),
        "oak_gpr_finish_f64",
    )
    return lml, alpha


# ---- Sobol ------------------------------------------------------------------------------
def sobol_L(spec: Spec, dim: int, Xcond, delta: float, mu: float, out=None):
    torch = _torch()
    lib = _cabi.load()
    m = int(Xcond.shape[0])
    if out is None:
        out = torch.empty((m, m), dtype=torch.float64, device=Xcond.device)
    work = torch.empty(max(int(lib.oak_sobol_L_work_bytes(spec.handle, dim, m)) // 8, 1), dtype=torch.float64,
                       device=Xcond.device)
    check(
        lib.oak_sobol_L_f64(spec.handle, int(dim), _p(Xcond), m, int(Xcond.stride(0)), float(delta), float(mu),
                            _p(out), int(out.stride(0)), _p(work), C.c_void_p(stream_ptr())),
        "oak_sobol_L_f64",
    )
    return out


def sobol_gaussian_terms(x, y, sigma: float, lengthscale: float, delta: float, mu: float):
    """(4, n) device tensor [f1, f2, f3, f4](x_i, y_i) of oak/utils.py:116-165."""
    torch = _torch()
    n = x.numel()
    assert y.numel() == n and x.is_contiguous() and y.is_contiguous()
    out = torch.empty((4, n), dtype=torch.float64, device=x.device)
    check(_cabi.load().oak_sobol_gaussian_terms_f64(_p(x), _p(y), n, float(sigma), float(lengthscale), float(delta),
                                                    float(mu), _p(out), C.c_void_p(stream_ptr())),
          "oak_sobol_gaussian_terms_f64")
    return out


def sobol_quadforms(Lstack, subsets: Sequence[Sequence[int]], scale: Sequence[float], alpha):
    torch = _torch()
    nc = len(subsets)
    max_order = max([len(s) for s in subsets] + [1])
    tab = np.full((nc, max_order), -1, dtype=np.int32)
    for c, s in enumerate(subsets):
        tab[c, : len(s)] = s
    dev = Lstack.device
    d_tab = torch.as_tensor(tab).to(dev)
    d_scale = torch.as_tensor(np.asarray(scale, dtype=np.float64)).to(dev)
    out = torch.empty(nc, dtype=torch.float64, device=dev)
    a = alpha.reshape(-1).contiguous()
    lib = _cabi.load()
    work = torch.empty(max(int(lib.oak_sobol_quadforms_work_bytes(nc, int(Lstack.shape[1]))) // 8, 1),
                       dtype=torch.float64, device=dev)
    check(
        lib.oak_sobol_quadforms_f64(_p(Lstack), int(Lstack.shape[0]), int(Lstack.shape[1]), _p(d_tab), _p(d_scale), nc,
                                    max_order, _p(a), _p(out), _p(work), C.c_void_p(stream_ptr())),
        "oak_sobol_quadforms_f64",
    )
    return out


def measure_fp64_peak(seconds: float = 1.0) -> float:
    _torch()
    v = C.c_double(0.0)
    check(_cabi.load().oak_measure_fp64_peak(float(seconds), C.byref(v), C.c_void_p(stream_ptr())),
          "oak_measure_fp64_peak")
    return float(v.value)
