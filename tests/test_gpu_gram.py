"""GPU parity: fused CUDA Gram tiles (through the C ABI) vs the NumPy oracle on the same inputs.

Tolerance: 1e-9 relative (north_star), measured as max|K - K_ref| / max|K_ref| and -- where no
entry is close to zero -- element-wise.
"""
import numpy as np
import pytest

from helpers import RTOL, build_oracle, max_rel_err, mixed_config

pytestmark = pytest.mark.gpu


def _product(cfg):
    from oak_b200.workloads import build_kernel

    return build_kernel(cfg)


def _gauss_cfg(n, D, depth, seed=0, share=True):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D))
    dims = [{"type": "rbf", "lengthscale": float(l), "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)}
            for l in rng.uniform(0.4, 3.0, D)]
    var = list(rng.uniform(0.2, 1.5, depth + 1)) if share else [0.7]
    return dict(X=X, dims=dims, depth=depth, variances=var, share_var=share)


@pytest.mark.parametrize("depth", [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 13])
@pytest.mark.parametrize("algo", [0, 1])
def test_gram_depths_vs_oracle(depth, algo):
    D = max(depth, 3) if depth < 9 else depth + 1
    cfg = _gauss_cfg(150, D, depth, seed=depth)
    k = _product(cfg)
    k.esp_algorithm = algo
    ref = build_oracle(cfg)
    X = cfg["X"]
    K = k.K(X)
    Kref = ref.K(X)
    assert K.shape == Kref.shape
    assert max_rel_err(K, Kref) < RTOL
    np.testing.assert_array_equal(K, K.T)  # mirrored tiles are bit-identical
    X2 = np.random.default_rng(99).standard_normal((77, D))
    assert max_rel_err(k.K(X, X2), ref.K(X, X2)) < RTOL
    assert max_rel_err(k.K_diag(X), ref.K_diag(X)) < RTOL


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 128, 129, 200, 513])
def test_gram_ragged_sizes(n):
    cfg = _gauss_cfg(n, 5, 3, seed=n)
    k, ref = _product(cfg), build_oracle(cfg)
    X = cfg["X"]
    assert max_rel_err(k.K(X), ref.K(X)) < RTOL
    X2 = np.random.default_rng(n + 1).standard_normal((max(n // 2, 1), 5))
    assert max_rel_err(k.K(X, X2), ref.K(X, X2)) < RTOL
    assert max_rel_err(k.K(X2, X), ref.K(X2, X)) < RTOL


def test_gram_empty_inputs():
    cfg = _gauss_cfg(10, 3, 2)
    k = _product(cfg)
    assert k.K(np.zeros((0, 3))).shape == (0, 0)
    assert k.K(cfg["X"], np.zeros((0, 3))).shape == (10, 0)
    assert k.K_diag(np.zeros((0, 3))).shape == (0,)


@pytest.mark.parametrize("depth", [1, 2, 3, 5])
@pytest.mark.parametrize("share", [True, False])
def test_gram_mixed_kernels_and_measures(depth, share):
    cfg = mixed_config(n=180, seed=depth, depth=depth, share=share)
    k, ref = _product(cfg), build_oracle(cfg)
    X = cfg["X"]
    assert max_rel_err(k.K(X), ref.K(X)) < RTOL
    assert max_rel_err(k.K(X, cfg["Z"]), ref.K(X, cfg["Z"])) < RTOL
    assert max_rel_err(k.K_diag(X), ref.K_diag(X)) < RTOL
    # the reference builds RBF with the expanded squared distance; same tolerance must hold
    ref_exp = build_oracle(cfg, expanded=True)
    assert max_rel_err(k.K(X), ref_exp.K(X)) < RTOL


def test_gram_many_dims_chunked():
    """D=50 exercises the dim-chunk pipeline (16 dims per stage) incl. a discrete tail."""
    rng = np.random.default_rng(5)
    cfg = _gauss_cfg(300, 50, 2, seed=5)
    X = cfg["X"]
    X[:, 48] = (rng.random(300) < 0.4)
    X[:, 49] = rng.integers(0, 3, 300)
    cfg["dims"][48] = {"type": "binary", "p0": 0.6, "variance": 1.0}
    cfg["dims"][49] = {"type": "categorical", "p": [0.3, 0.3, 0.4], "W": rng.uniform(0, 1, (3, 2)),
                       "kappa": np.ones(3), "variance": 1.0}
    k, ref = _product(cfg), build_oracle(cfg)
    assert max_rel_err(k.K(X), ref.K(X)) < RTOL
    assert max_rel_err(k.K(X[:100], X), ref.K(X[:100], X)) < RTOL


def test_gram_sub_kernels_standalone():
    """test_kernel_1d of the reference (tests/test_kernel_properties.py:57-66) + oracle parity."""
    from oak_b200.input_measures import EmpiricalMeasure, GaussianMeasure, MOGMeasure, UniformMeasure
    from oak_b200.ortho_binary_kernel import OrthogonalBinary
    from oak_b200.ortho_rbf_kernel import RBF, OrthogonalRBFKernel
    from oracle import oak_oracle as oo

    X = np.array([[0.1], [0.5], [0.5]])
    pairs = [
        (OrthogonalRBFKernel(RBF(), GaussianMeasure(0, 1)), oo.RBFDim(1.0, 1.0, oo.Gaussian(0, 1))),
        (OrthogonalRBFKernel(RBF(), UniformMeasure(0, 1)), oo.RBFDim(1.0, 1.0, oo.Uniform(0, 1))),
        (OrthogonalRBFKernel(RBF(), EmpiricalMeasure(X)), oo.RBFDim(1.0, 1.0, oo.Empirical(X))),
        (OrthogonalRBFKernel(RBF(lengthscales=10), MOGMeasure(np.array([3.0, 2.0]), np.array([3.0, 10.0]),
                                                              np.array([0.6, 0.4]))),
         oo.RBFDim(10.0, 1.0, oo.MOG([3.0, 2.0], [3.0, 10.0], [0.6, 0.4]))),
        (OrthogonalBinary(), oo.BinaryDim(0.5)),
        (RBF(lengthscales=0.3, variance=2.0), oo.RBFDim(0.3, 2.0, None)),
    ]
    for k, ref in pairs:
        Xi = X if not isinstance(k, OrthogonalBinary) else np.array([[0.0], [1.0], [1.0]])
        K = k.K(Xi, Xi)
        np.testing.assert_allclose(np.diag(K), k.K_diag(Xi), rtol=1e-12)
        np.testing.assert_allclose(K, k(Xi, Xi), rtol=0, atol=0)
        assert max_rel_err(K, ref.K(Xi, Xi)) < RTOL
        assert max_rel_err(k.K_diag(Xi), ref.K_diag(Xi)) < RTOL


def test_cov_X_s_and_var_s_closures():
    from oak_b200.input_measures import GaussianMeasure, UniformMeasure
    from oak_b200.ortho_rbf_kernel import RBF, OrthogonalRBFKernel
    from oracle import oak_oracle as oo

    x = np.random.default_rng(0).uniform(-1, 2, (40, 1))
    for meas, omeas in ((GaussianMeasure(0.3, 2.0), oo.Gaussian(0.3, 2.0)), (UniformMeasure(-1, 2), oo.Uniform(-1, 2))):
        k = OrthogonalRBFKernel(RBF(lengthscales=0.8, variance=1.7), meas)
        ref = oo.RBFDim(0.8, 1.7, omeas)
        assert abs(k.var_s() - ref.var_s()) / ref.var_s() < 1e-12
        assert max_rel_err(k.cov_X_s(x), ref.cov_X_s(x)) < 1e-12


def test_mog_equals_gaussian():
    """tests/test_orthogonality.py:152-165 of the reference."""
    from oak_b200.input_measures import GaussianMeasure, MOGMeasure
    from oak_b200.ortho_rbf_kernel import RBF, OrthogonalRBFKernel

    k_gmm = OrthogonalRBFKernel(RBF(lengthscales=10.0), MOGMeasure(np.array([3.0, 3.0]), np.array([5.0, 5.0]),
                                                                    np.array([0.2, 0.8])))
    k_g = OrthogonalRBFKernel(RBF(lengthscales=10.0), GaussianMeasure(3, 5))
    xx = np.array([[-2], [2.0], [3.0]])
    np.testing.assert_allclose(k_g.K(xx), k_gmm.K(xx), rtol=1e-7)


def test_row_range_sharding_is_bit_identical():
    """Row-block sharding: virtual ranks G in {2,4,8} reproduce the unsharded Gram."""
    import torch

    from oak_b200 import _device
    from oak_b200.parallel import partition_rows

    cfg = _gauss_cfg(1000, 6, 3, seed=3)
    k = _product(cfg)
    Xd = _device.to_device(cfg["X"])
    spec = k._make_spec()
    px = _device.Points(spec, Xd)
    full = _device.gram(spec, px)
    X2 = _device.to_device(np.random.default_rng(1).standard_normal((333, 6)))
    px2 = _device.Points(spec, X2)
    full2 = _device.gram(spec, px, px2)
    for G in (2, 4, 8):
        parts = [_device.gram(spec, px, None, b, e) for b, e in partition_rows(1000, G)]
        assert torch.equal(torch.cat(parts, 0), full)
        parts2 = [_device.gram(spec, px, px2, b, e) for b, e in partition_rows(1000, G)]
        assert torch.equal(torch.cat(parts2, 0), full2)
    spec.close()


def test_lower_trapezoid_strips_cover_the_lower_triangle():
    """oak_gram_lower_f64 over the folded strip assignment reproduces tril(K) bit for bit."""
    import torch

    from oak_b200 import _device
    from oak_b200.parallel import balanced_symmetric_rows

    n = 1100
    cfg = _gauss_cfg(n, 6, 4, seed=13)
    k = _product(cfg)
    spec = k._make_spec()
    px = _device.Points(spec, _device.to_device(cfg["X"]))
    full = _device.gram(spec, px)
    for G in (1, 2, 4):
        got = torch.zeros_like(full)
        rows_done = 0
        for strips in balanced_symmetric_rows(n, G):
            for b, e in strips:
                if e > b:
                    got[b:e, :e] = _device.gram_lower(spec, px, b, e)
                    rows_done += e - b
        assert rows_done == n
        assert torch.equal(torch.tril(got), torch.tril(full))
    spec.close()


@pytest.mark.parametrize("odd_pitch", [False, True])
def test_mirrored_strips_hold_the_whole_symmetric_matrix(odd_pitch):
    """oak_gram_lower_mirror_f64: every rank's strip plus its mirror image; together, bit for bit, the matrix the
    single-GPU symmetric call writes (the multi-GPU bench arm computes the same product as N = 1).  An odd
    pitch of the mirror block takes the direct-store path instead of the TMA stores."""
    import torch

    from oak_b200 import _device
    from oak_b200.parallel import balanced_symmetric_rows

    n = 1100
    cfg = _gauss_cfg(n, 6, 4, seed=14)
    k = _product(cfg)
    spec = k._make_spec()
    px = _device.Points(spec, _device.to_device(cfg["X"]))
    full = _device.gram(spec, px)
    for G in (1, 2, 4):
        got = torch.full_like(full, float("nan"))
        for strips in balanced_symmetric_rows(n, G):
            for b, e in strips:
                if e <= b:
                    continue
                K = torch.full((e - b, e), float("nan"), dtype=torch.float64, device="cuda")
                pitch = (e - b) + (1 if (e - b) % 2 == 0 else 0) if odd_pitch else (e - b + 1) // 2 * 2
                Ktb = torch.full((e, pitch), float("nan"), dtype=torch.float64, device="cuda")
                Kt = Ktb[:, : e - b]
                _device.gram_lower_mirror(spec, px, b, e, out=K, out_t=Kt)
                assert torch.isnan(Ktb[:, e - b:]).all()      # nothing outside the block
                blk = got[b:e, :e]
                got[b:e, :e] = torch.where(torch.isnan(K), blk, K)
                blk = got[:e, b:e]
                got[:e, b:e] = torch.where(torch.isnan(Kt), blk, Kt)
        assert not torch.isnan(got).any()
        assert torch.equal(got, full)
    spec.close()


def test_torch_inputs_stay_on_device():
    import torch

    cfg = _gauss_cfg(70, 4, 2)
    k, ref = _product(cfg), build_oracle(cfg)
    Xd = torch.as_tensor(cfg["X"]).cuda()
    K = k.K(Xd)
    assert isinstance(K, torch.Tensor) and K.is_cuda
    assert max_rel_err(K.cpu().numpy(), ref.K(cfg["X"])) < RTOL


def test_small_lengthscale_underflow_is_clean():
    """exp underflow (|x-y| / l huge) must give exact zeros for the RBF part, no NaN/garbage."""
    cfg = _gauss_cfg(100, 3, 2, seed=8)
    for d in cfg["dims"]:
        d["lengthscale"] = 0.001
    k, ref = _product(cfg), build_oracle(cfg)
    K, Kref = k.K(cfg["X"]), ref.K(cfg["X"])
    assert np.all(np.isfinite(K))
    assert max_rel_err(K, Kref) < RTOL


def test_invalid_category_raises():
    cfg = mixed_config(n=50, depth=2)
    k = _product(cfg)
    X = cfg["X"].copy()
    X[3, 5] = 7.0
    with pytest.raises(ValueError):
        k.K(X)


def test_gram_host_pipeline_matches_device():
    """oak_gram_host_f64 (host buffers, row-blocked, overlapped D2H) == device Gram."""
    import ctypes as C

    import torch

    from oak_b200 import _cabi, _device

    cfg = _gauss_cfg(700, 5, 3, seed=11)
    k, ref = _product(cfg), build_oracle(cfg)
    X = np.ascontiguousarray(cfg["X"])
    spec = k._make_spec()
    lib = _cabi.load()
    for X2 in (None, np.ascontiguousarray(np.random.default_rng(2).standard_normal((130, 5)))):
        cols = 700 if X2 is None else X2.shape[0]
        out = torch.empty((700, cols), dtype=torch.float64).pin_memory()
        wb = lib.oak_gram_host_work_bytes(spec.handle, 700, 0 if X2 is None else cols, 5, 128)
        work = torch.empty(wb // 8 + 1, dtype=torch.float64, device="cuda")
        rc = lib.oak_gram_host_f64(spec.handle, X.ctypes.data, 700, None if X2 is None else X2.ctypes.data,
                                   0 if X2 is None else cols, 5, out.data_ptr(), cols, 128, work.data_ptr(),
                                   C.c_void_p(_device.stream_ptr()))
        assert rc == 0, _cabi.last_error()
        assert max_rel_err(out.numpy(), ref.K(X, X2)) < RTOL
    spec.close()


def test_gram_host_lower_pipeline_halves_the_copies_and_can_mirror():
    """oak_gram_host_lower_f64: the symmetric Gram through host buffers, lower trapezoid row blocks only; with
    mirror=1 the host fills the upper triangle and the result is the full matrix; a row sub-range is a rank's share."""
    import ctypes as C

    import torch

    from oak_b200 import _cabi, _device

    n = 700
    cfg = _gauss_cfg(n, 5, 3, seed=12)
    k, ref = _product(cfg), build_oracle(cfg)
    X = np.ascontiguousarray(cfg["X"])
    Kref = ref.K(X)
    spec = k._make_spec()
    lib = _cabi.load()
    wb = lib.oak_gram_host_lower_work_bytes(spec.handle, n, 5, n, 128)
    work = torch.empty(wb // 8 + 1, dtype=torch.float64, device="cuda")
    for mirror in (0, 1):
        out = torch.full((n, n), float("nan"), dtype=torch.float64).pin_memory()
        rc = lib.oak_gram_host_lower_f64(spec.handle, X.ctypes.data, n, 5, 0, n, out.data_ptr(), n, 128, mirror,
                                         work.data_ptr(), C.c_void_p(_device.stream_ptr()))
        assert rc == 0, _cabi.last_error()
        got = out.numpy()
        assert max_rel_err(np.tril(got), np.tril(Kref)) < RTOL
        if mirror:
            assert max_rel_err(got, Kref) < RTOL and np.array_equal(got, got.T)
        else:
            assert np.isnan(got[0, n - 1])       # far right of the first block: never copied
    b, e = 256, 640                              # a rank's row range
    out = torch.full((e - b, e), float("nan"), dtype=torch.float64).pin_memory()
    rc = lib.oak_gram_host_lower_f64(spec.handle, X.ctypes.data, n, 5, b, e, out.data_ptr(), e, 128, 0,
                                     work.data_ptr(), C.c_void_p(_device.stream_ptr()))
    assert rc == 0, _cabi.last_error()
    got = out.numpy()
    mask = np.tril(np.ones((n, n), dtype=bool))[b:e, :e]
    assert np.abs(got[mask] - Kref[b:e, :e][mask]).max() / np.abs(Kref).max() < RTOL
    spec.close()


def test_fast_and_general_exp_bodies_agree_per_stage():
    """The clamp-free exp body is taken per 16-dim stage when every dim of the stage qualifies
    (s^2 == 1, bounded |a_i - b_j|).  D = 40: stage 0 has a tiny lengthscale (general body), stage 1
    has s^2 != 1 (general), stage 2 qualifies (fast).  All must match the oracle; and forcing the
    general body everywhere (min/max keys ignored) must agree with the default to rounding."""
    cfg = _gauss_cfg(300, 40, 3, seed=21)
    cfg["dims"][3]["lengthscale"] = 0.004
    cfg["dims"][20]["variance"] = 1.7
    k, ref = _product(cfg), build_oracle(cfg)
    X = cfg["X"]
    K, Kref = k.K(X), ref.K(X)
    assert np.all(np.isfinite(K))
    assert max_rel_err(K, Kref) < RTOL
    X2 = np.random.default_rng(5).standard_normal((170, 40))
    assert max_rel_err(k.K(X, X2), ref.K(X, X2)) < RTOL


def test_many_dims_beyond_the_fast_flag_words():
    """D = 300 > 256: dims past the flag words always take the general body."""
    cfg = _gauss_cfg(130, 300, 2, seed=22)
    k, ref = _product(cfg), build_oracle(cfg)
    assert max_rel_err(k.K(cfg["X"]), ref.K(cfg["X"])) < RTOL


def test_unaligned_and_odd_pitch_outputs_take_the_direct_store_path():
    """The mirrored tile leaves through TMA only for 16-byte aligned outputs with an even pitch;
    other outputs (views at an odd offset / odd leading dimension) use direct stores.  Both paths
    must produce the same bits."""
    import torch

    from oak_b200 import _device

    cfg = _gauss_cfg(448, 6, 4, seed=23)
    k = _product(cfg)
    spec = k._make_spec()
    px = _device.Points(spec, _device.to_device(cfg["X"]))
    base = _device.gram(spec, px)                      # aligned, even pitch -> TMA mirror
    big = torch.full((449, 451), float("nan"), dtype=torch.float64, device="cuda")
    view = big[1:, 1:449]                              # odd pitch (451), 8-byte offset
    _device.gram(spec, px, out=view)
    assert torch.equal(view, base)
    assert torch.isnan(big[0]).all() and torch.isnan(big[:, 0]).all() and torch.isnan(big[:, 449:]).all()
    big2 = torch.full((448, 450), float("nan"), dtype=torch.float64, device="cuda")
    view2 = big2[:, :448]                              # even pitch, aligned, wider than n -> TMA with ldk > n
    _device.gram(spec, px, out=view2)
    assert torch.equal(view2, base)
    assert torch.isnan(big2[:, 448:]).all()
    spec.close()


def test_non_finite_inputs_poison_only_their_rows_and_columns():
    cfg = _gauss_cfg(200, 4, 3, seed=24)
    k, ref = _product(cfg), build_oracle(cfg)
    X = cfg["X"].copy()
    X[17, 2] = np.nan
    K = k.K(X)
    bad = np.zeros(200, bool)
    bad[17] = True
    assert np.all(np.isnan(K[17, :])) and np.all(np.isnan(K[:, 17]))
    good = ~bad
    Kref = ref.K(cfg["X"])
    assert max_rel_err(K[np.ix_(good, good)], Kref[np.ix_(good, good)]) < RTOL
    # +-inf: exp(-inf) = 0 in the reference; the kernel must stay finite off the poisoned point
    X = cfg["X"].copy()
    X[5, 1] = np.inf
    K = k.K(X)
    good = np.ones(200, bool)
    good[5] = False
    assert np.all(np.isfinite(K[np.ix_(good, good)]))
    assert max_rel_err(K[np.ix_(good, good)], Kref[np.ix_(good, good)]) < RTOL
