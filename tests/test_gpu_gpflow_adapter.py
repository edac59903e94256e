"""The guarded gpflow adapter (``oak_b200.gpflow_adapter``): a ``gpflow.kernels.Kernel``
subclass around the CUDA tiles.  Real gpflow / TensorFlow are not installable in this image, so ``oracle/tf_shim``
stands in for them (forward values only); the gradient callbacks behind ``tf.custom_gradient`` are plain NumPy
functions and are checked directly against autograd of the oracle -- only the TensorFlow wiring itself stays
untested."""
import os
import sys

import numpy as np
import pytest

from helpers import RTOL, build_oracle, load_golden, max_rel_err
from oracle import oak_grad_oracle as go
from oracle import oak_oracle as oo

pytestmark = pytest.mark.gpu
SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "tf_shim")


@pytest.fixture
def shim_gpflow():
    """``gpflow`` / ``tensorflow`` importable (the NumPy shim) for the duration of one test."""
    sys.path.insert(0, SHIM)
    try:
        import gpflow

        yield gpflow
    finally:
        sys.path.remove(SHIM)
        for name in [m for m in sys.modules if m.split(".")[0] in ("gpflow", "tensorflow", "tensorflow_probability")]:
            del sys.modules[name]


def test_adapter_reports_unavailable_without_gpflow():
    from oak_b200 import gpflow_adapter

    assert "gpflow" not in sys.modules
    assert gpflow_adapter.available() is False
    with pytest.raises(ImportError):
        gpflow_adapter.as_gpflow_kernel(object())


def test_adapter_kernel_is_a_gpflow_kernel_and_matches_the_reference_golden(shim_gpflow):
    from oak_b200 import gpflow_adapter
    from oak_b200.workloads import build_kernel

    gpflow = shim_gpflow
    assert gpflow_adapter.available()
    cfg, g = load_golden("g1_gaussian_d5_p3")
    shell = gpflow_adapter.as_gpflow_kernel(build_kernel(cfg))
    assert isinstance(shell, gpflow.kernels.Kernel)
    X, X2 = g["X"], g["X2"]
    assert max_rel_err(shell(X), g["K"]) < RTOL                      # gpflow's Kernel.__call__ protocol
    assert max_rel_err(shell(X, X2), g["K_cross"]) < RTOL
    assert max_rel_err(shell(X, full_cov=False), g["K_diag"]) < RTOL
    # the parameters gpflow sees are the kernel's: 5 x (lengthscale, base variance) + 4 order variances
    assert len(shell.oak_parameters) == 2 * 5 + 4
    # a changed gpflow Parameter reaches the tiles
    shell.oak_parameters[0].assign(0.37)
    cfg2 = dict(cfg, dims=[dict(d) for d in cfg["dims"]])
    cfg2["dims"][0]["lengthscale"] = 0.37
    assert max_rel_err(shell(X, X2), build_oracle(cfg2).K(X, X2)) < RTOL
    # inside a gpflow model class that is not this package's: the shim's GPR
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((X.shape[0], 1))
    m = gpflow.models.GPR((X, Y), kernel=shell)
    m.likelihood.variance.assign(0.1)
    lml_ref = oo.gpr_log_marginal_likelihood(build_oracle(cfg2), X, Y, 0.1)
    assert abs(m.log_marginal_likelihood() - lml_ref) < RTOL * abs(lml_ref)


@pytest.mark.parametrize("same", [True, False])
def test_adapter_gradient_callbacks_match_autograd(shim_gpflow, same):
    import torch

    from oak_b200 import gpflow_adapter
    from oak_b200.workloads import build_kernel

    rng = np.random.default_rng(3)
    D, P = 4, 3
    ls, var = rng.uniform(0.6, 2.0, D), rng.uniform(0.3, 1.2, P + 1)
    cfg = dict(dims=[{"type": "rbf", "lengthscale": float(l), "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)} for l in ls],
               depth=P, variances=list(var), share_var=True)
    shell = gpflow_adapter.as_gpflow_kernel(build_kernel(cfg))
    X, X2 = rng.standard_normal((70, D)), rng.standard_normal((45, D))
    dy = rng.standard_normal((70, 70 if same else 45))
    theta = [np.asarray(p.numpy()) for p in shell.oak_parameters]
    outs = shell._np_K_grad(X, X if same else X2, same, dy, *theta)
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731
    Xt, lst, vt = t(X).requires_grad_(True), t(ls).requires_grad_(True), t(var).requires_grad_(True)
    K = go.oak_K(Xt, Xt if same else t(X2), lst, vt)
    (K * t(dy)).sum().backward()
    assert max_rel_err(outs[0], Xt.grad.numpy()) < 1e-9
    got_ls = np.array([float(outs[1 + 2 * i]) for i in range(D)])       # slots: (lengthscale, base variance) per dim
    got_var = np.array([float(v) for v in outs[1 + 2 * D:]])
    assert max_rel_err(got_ls, lst.grad.numpy()) < 1e-9
    assert max_rel_err(got_var, vt.grad.numpy()) < 1e-9
    # K_diag
    w = rng.standard_normal(70)
    outs_d = shell._np_K_diag_grad(X, w, *theta)
    lst2, vt2 = t(ls).requires_grad_(True), t(var).requires_grad_(True)
    (go.oak_K_diag(t(X), lst2, vt2) * t(w)).sum().backward()
    assert max_rel_err(np.array([float(outs_d[2 * i]) for i in range(D)]), lst2.grad.numpy()) < 1e-9
    assert max_rel_err(np.array([float(v) for v in outs_d[2 * D:]]), vt2.grad.numpy()) < 1e-9
