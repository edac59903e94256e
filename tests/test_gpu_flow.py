"""Input pipeline on the device: the per-column normalising flow
(oak/normalising_flow.py, model_utils.py:179-191, 305-317) against the NumPy oracle."""
import numpy as np
import pytest

from helpers import max_rel_err
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log", [True, False])
@pytest.mark.parametrize("n", [1, 257, 100_003])
def test_flow_objective_and_gradient_match_oracle(log, n):
    from oak_b200 import _device

    rng = np.random.default_rng(n)
    x = np.exp(0.5 * rng.standard_normal(n)) + 2.0
    offset = x.min() - 1.0 if log else 0.0
    xd = _device.to_device(x, ndim=1)
    for theta in ([0.0, 0.0, 0.0, 0.0], [0.3, -0.8, 0.2, -0.15], [-0.4, -1.5, -0.3, 0.25]):
        out = _device.flow_objective(xd, offset, log, theta)
        J, g = fo.kl_objective_and_grad(x, offset, log, np.array(theta))
        assert abs(out[0] - J) < 1e-12 * max(1.0, abs(J))
        assert max_rel_err(out[1:], g) < 1e-11


def test_flow_objective_reads_strided_columns():
    import torch

    from oak_b200 import _device

    rng = np.random.default_rng(1)
    X = np.exp(0.4 * rng.standard_normal((5000, 3))) + 1.5
    Xd = _device.to_device(X)
    for c in range(3):
        out = _device.flow_objective(Xd[:, c], X[:, c].min() - 1.0, True, [0.1, -0.4, 0.05, 0.0])
        J, g = fo.kl_objective_and_grad(X[:, c], X[:, c].min() - 1.0, True, np.array([0.1, -0.4, 0.05, 0.0]))
        assert abs(out[0] - J) < 1e-12 * max(1.0, abs(J)) and max_rel_err(out[1:], g) < 1e-11
    assert torch.equal(Xd, _device.to_device(X))  # read-only


@pytest.mark.parametrize("log", [True, False])
def test_normalizer_fit_on_device_gaussianises_and_round_trips(log):
    from scipy import stats

    from oak_b200.normalising_flow import Normalizer

    rng = np.random.default_rng(0)
    x = np.exp(0.6 * rng.standard_normal(4000)) + 3.0
    n = Normalizer(x, log=log)
    j0 = n.KL_objective()
    res = n.fit()
    assert n.KL_objective() <= j0 + 1e-12 and np.isfinite(res.fun)
    # the optimum of the device objective is a stationary point of the oracle's
    theta = n._theta()
    J, g = fo.kl_objective_and_grad(x, n.offset, log, theta)
    assert abs(J - res.fun) < 1e-12 * max(1.0, abs(J)) and np.abs(g).max() < 1e-3
    y = n.bijector(x)
    par = (n.offset, log, float(n.shift.numpy()), float(n.scale.numpy()), float(n.skewness.numpy()),
           float(n.tailweight.numpy()))
    assert max_rel_err(y, fo.forward(x, *par)) < 1e-13
    np.testing.assert_allclose(n.bijector.inverse(y), x, rtol=1e-10)
    if log:
        assert abs(y.mean()) < 0.05 and abs(y.std() - 1.0) < 0.05
        assert stats.kstest(y, "norm")[1] > 0.01
        assert stats.kstest((x - x.mean()) / x.std(), "norm")[1] < 1e-6  # the raw column is far from normal


def test_apply_normalise_flow_transforms_columns_in_place_on_the_device():
    import torch

    from oak_b200 import _device
    from oak_b200.model_utils import apply_normalise_flow
    from oak_b200.normalising_flow import Normalizer

    rng = np.random.default_rng(2)
    X = np.exp(0.5 * rng.standard_normal((3000, 4))) + 1.0
    X[:, 2] = rng.integers(0, 3, 3000)  # a categorical column: no flow
    flows = [Normalizer(X[:, 0]), None, None, Normalizer(X[:, 3], log=False)]
    for f in (flows[0], flows[3]):
        f.fit(maxiter=20)
    out = apply_normalise_flow(X, flows)
    assert isinstance(out, np.ndarray) and out.shape == X.shape
    for c, f in enumerate(flows):
        if f is None:
            assert np.array_equal(out[:, c], X[:, c])
        else:
            par = (f.offset, f.log, float(f.shift.numpy()), float(f.scale.numpy()), float(f.skewness.numpy()),
                   float(f.tailweight.numpy()))
            assert max_rel_err(out[:, c], fo.forward(X[:, c], *par)) < 1e-13
    Xd = _device.to_device(X)
    out_d = apply_normalise_flow(Xd, flows)
    assert out_d.is_cuda and torch.equal(Xd, _device.to_device(X))  # the caller's tensor is left alone
    assert max_rel_err(out_d.cpu().numpy(), out) == 0.0
    assert np.array_equal(apply_normalise_flow(X, [None] * 4), X)


def test_device_flow_matches_the_references_normalizer_golden():
    """g8 (the reference's own Normalizer over the TF/TFP shim): the device transform and objective at the initial
    and the perturbed parameter settings."""
    import os

    from oak_b200 import _device

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g8_normalising_flow.npz"))
    x = g["x"]
    xd = _device.to_device(x, ndim=1)
    for tag, log in (("log", True), ("nolog", False)):
        offset = x.min() - 1.0 if log else 0.0
        for step in range(3):
            th = g[f"{tag}_theta_{step}"]
            y = _device.flow_forward(xd, offset, log, th[1], np.exp(th[0]), th[2], np.exp(th[3])).cpu().numpy()
            assert max_rel_err(y, g[f"{tag}_y_{step}"]) < 1e-13
            J = _device.flow_objective(xd, offset, log, th)[0]
            assert abs(J - float(g[f"{tag}_J_{step}"])) < 1e-12 * max(1.0, abs(J))


# ---- the rest of the input pipeline on the device: distinct values + counts, column means -------
@pytest.mark.parametrize("n", [1, 7, 1000, 2048, 2049, 50_000, 300_001])
def test_column_unique_matches_numpy_bit_for_bit(n):
    """oak_column_unique_f64 == np.unique(col, return_counts=True): the empirical-measure locations / weights
    (oak/model_utils.py:334-344) and the category frequencies (:736-739)."""
    from oak_b200 import _device

    rng = np.random.default_rng(n)
    X = np.zeros((n, 3))
    X[:, 0] = np.round(8 * rng.standard_normal(n)) / 8            # config D's empirical-measure column: few levels
    X[:, 0] = (X[:, 0] - X[:, 0].mean()) / (X[:, 0].std() if n > 1 else 1.0)
    X[:, 1] = rng.integers(0, 6, n).astype(float)                  # categorical levels
    X[:, 2] = rng.standard_normal(n)                               # all distinct
    if n > 5:
        X[3, 2], X[4, 2] = -0.0, 0.0                               # one value for np.unique
    Xd = _device.to_device(X)
    for col in range(3):
        vals, counts = _device.column_unique(Xd, col)
        ref_v, ref_c = np.unique(X[:, col], return_counts=True)
        assert vals.shape == ref_v.shape and np.array_equal(np.abs(vals), np.abs(ref_v)) and np.array_equal(vals == 0, ref_v == 0)
        assert np.array_equal(counts, ref_c)
        assert np.array_equal(counts / counts.sum(), ref_c / ref_c.sum())     # the weights, bit for bit
    b = (rng.random(n) < 0.3).astype(float)
    assert _device.column_mean(_device.to_device(b.reshape(-1, 1)), 0) == b.mean()   # exact for 0/1 columns


def test_feature_typing_and_empirical_measures_come_from_the_device():
    """_calculate_features (p0, p) and the empirical locations / weights of oak_model.fit equal the NumPy results."""
    from oak_b200 import model_utils as mu

    rng = np.random.default_rng(3)
    n = 4000
    X = np.zeros((n, 4))
    X[:, 0] = (rng.random(n) < 0.35).astype(float)
    X[:, 1] = rng.integers(0, 5, n).astype(float)
    X[:, 2] = rng.standard_normal(n)
    X[:, 3] = np.round(4 * rng.standard_normal(n)) / 4
    cont, binary, cat, p0, p = mu._calculate_features(X, categorical_feature=[1], binary_feature=[0])
    assert (cont, binary, cat) == ([2, 3], [0], [1])
    assert p0[0] == 1 - X[:, 0].mean() and p0[1] is None
    _, cnt = np.unique(X[:, 1], return_counts=True)
    assert np.array_equal(p[1], (cnt / n).reshape(-1, 1))
    stats = mu._ColumnStats(X)
    assert stats.Xd is not None                                    # the device path is the one that ran
    loc, c3 = stats.unique(3)
    ref_loc, ref_c = np.unique(X[:, 3], return_counts=True)
    assert np.array_equal(loc, ref_loc) and np.array_equal(c3, ref_c)
