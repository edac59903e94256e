"""GPU parity of the factor-first SGPR path: the one-launch bordered Cholesky, the DMMA panel product, the
device-side route choice and -- the point of it -- the ELBO on an ILL-CONDITIONED Kuu (long lengthscales,
cond(Kuu) ~ 2.6e8), where forming Phi = Kuf Kuf^T first is 2e-6 away from gpflow's operation order
(oak/utils.py:186-198) and the whitened route stays inside 1e-9."""
import numpy as np
import pytest

from helpers import RTOL, build_oracle, max_rel_err, mixed_config
from oracle import oak_oracle as oo

pytestmark = pytest.mark.gpu


def _spd(n, seed, cond_pow=2.0):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = 10.0 ** rng.uniform(-cond_pow, 1.0, n)
    return (Q * ev) @ Q.T


@pytest.mark.parametrize("n", [1, 5, 64, 65, 130, 200, 513, 1024, 1030])
@pytest.mark.parametrize("border", ["none", "identity", "rows"])
def test_bordered_cholesky_against_lapack(n, border):
    import scipy.linalg as sla
    import torch

    from oak_b200 import _device

    A = _spd(n, n)
    rng = np.random.default_rng(n + 1)
    nb = {"none": 0, "identity": n, "rows": 3}[border]
    gap = 0 if border == "none" else (-(n) % 8)
    R = np.eye(n) if border == "identity" else rng.standard_normal((nb, n))
    ld = n + gap + nb + 2
    buf = np.full((n, ld), np.nan)          # row j of the tensor = column j of the matrix
    buf[:, :n] = np.tril(A).T + np.triu(np.full((n, n), 7.0), 1).T  # upper triangle of the matrix: junk, never read
    if nb:
        buf[:, n + gap: n + gap + nb] = R.T
    d = torch.as_tensor(buf).cuda()
    info, logdet = _device.chol(d, n, n + nb, gap=gap, border_identity=(border == "identity"))
    out = d.cpu().numpy()
    assert int(info.item()) == 0
    L = np.linalg.cholesky(A)
    got_L = np.tril(out[:, :n].T)
    assert max_rel_err(got_L, L) < 1e-11
    assert abs(float(logdet.item()) - np.log(np.diag(L)).sum()) < 1e-11 * max(1.0, n)
    if nb:
        X = sla.solve_triangular(L, R.T, lower=True).T   # Border L^-T
        got = out[:, n + gap: n + gap + nb].T
        scale = np.abs(X).max()
        assert np.abs(got - X).max() / scale < 1e-10     # carries cond(L) ~ 1e2
    # untouched: the strict upper triangle of the symmetric block
    iu = np.triu_indices(n, 1)
    assert np.all(out[:, :n].T[iu] == 7.0)


def test_bordered_cholesky_reports_the_failing_minor():
    import torch

    from oak_b200 import _device

    n = 150
    A = _spd(n, 3)
    A[100, 100] = -1.0
    d = torch.as_tensor(np.tril(A).T.copy()).cuda()
    info, _ = _device.chol(d, n, n)
    assert int(info.item()) == 101
    A2 = _spd(n, 4)
    A2[70, 70] = np.nan
    d = torch.as_tensor(np.tril(A2).T.copy()).cuda()
    info, _ = _device.chol(d, n, n)
    assert int(info.item()) == 71


@pytest.mark.parametrize("M,Kd,n", [(8, 8, 2), (64, 64, 130), (200, 200, 1000), (256, 256, 4097), (1024, 1024, 3000),
                                    (130, 70, 515)])
@pytest.mark.parametrize("lower", [False, True])
def test_panel_gemm_against_torch(M, Kd, n, lower):
    import torch

    from oak_b200 import _device

    if lower and M != Kd:
        pytest.skip("triangular left factor is square")
    g = torch.Generator(device="cpu").manual_seed(M + n)
    T = torch.randn((M, Kd), generator=g, dtype=torch.float64)
    if lower:
        T = torch.tril(T)
    ldb = (n + 1) // 2 * 2 + 4
    Bfull = torch.randn((Kd, ldb), generator=g, dtype=torch.float64)
    u = torch.randn(M, generator=g, dtype=torch.float64)
    v = torch.randn(ldb, generator=g, dtype=torch.float64)
    Td, Bd = T.cuda(), Bfull.cuda()[:, :n]
    ref = T @ Bfull[:, :n]
    out = _device.panel_gemm(Td, Bd, lower=lower).cpu()
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-13
    out = _device.panel_gemm(Td, Bd, lower=lower, u=u.cuda(), v=v.cuda()[:n]).cpu()
    ref = ref + u[:, None] * v[None, :n]
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-13


def _ill_conditioned(seed=0, D=5, M=256, N=4000, lo=2.0, hi=6.0, depth=3):
    """D = 5, lengthscales ~ U(2, 6), M = 256, N = 4000: cond(Kuu) ~ 2.6e8."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    ls = rng.uniform(lo, hi, D)
    y = (np.sin(X).sum(1) / np.sqrt(D) + X[:, 0] * X[:, 1] + 0.1 * rng.standard_normal(N)).reshape(-1, 1)
    dims = [{"type": "rbf", "lengthscale": float(l), "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)} for l in ls]
    return dict(X=X, y=y, Z=X[:M].copy(), dims=dims, depth=depth, variances=[1.0, 1.0, 0.5, 0.25][: depth + 1],
                share_var=True, noise=0.01)


def _sgpr(cfg, **kw):
    from oak_b200.models import SGPR
    from oak_b200.workloads import build_kernel

    m = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=cfg.get("chunk", 1024), **kw)
    m.likelihood.variance.assign(cfg["noise"])
    return m


@pytest.mark.parametrize("seed", [0, 1])
def test_elbo_on_ill_conditioned_kuu_matches_gpflow_order(seed):
    cfg = _ill_conditioned(seed)
    ref = build_oracle(cfg)
    elbo_ref = oo.sgpr_elbo(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    m = _sgpr(cfg)                        # route from the condition estimate
    elbo = m.elbo()
    kuu = ref.K(cfg["Z"]) + 1e-6 * np.eye(cfg["Z"].shape[0])
    cond = np.linalg.cond(kuu)
    assert cond > 1e8
    assert m.last_route == 1, (m.last_route, m.last_cond_estimate)
    assert cond / 4 < m.last_cond_estimate < cond * 4
    assert abs(elbo - elbo_ref) / abs(elbo_ref) < RTOL, (elbo, elbo_ref)
    # the forced un-whitened route is what round 1 shipped: far outside the budget here
    err_phi = abs(_sgpr(cfg, whiten_stats=False).elbo() - elbo_ref) / abs(elbo_ref)
    assert err_phi > 10 * abs(elbo - elbo_ref) / abs(elbo_ref)
    alpha = m.sufficient_statistics().cpu().numpy()
    alpha_ref = oo.sgpr_alpha(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    assert max_rel_err(alpha, alpha_ref) < 1e-3   # alpha itself carries cond(Kuu) ~ 1e8 in any operation order


def test_long_lengthscales_of_the_reference_sobol_fixture():
    """Trained OAK lengthscales of 3-9 are normal (reference tests/test_sobol_oak_kernel.py:41-75)."""
    cfg = _ill_conditioned(seed=3, D=8, lo=3.0, hi=9.0)
    ref = build_oracle(cfg)
    elbo_ref = oo.sgpr_elbo(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    m = _sgpr(cfg)
    elbo = m.elbo()
    assert m.last_route == 1
    assert abs(elbo - elbo_ref) / abs(elbo_ref) < RTOL, (elbo, elbo_ref)


@pytest.mark.parametrize("whiten", [None, False, True])
@pytest.mark.parametrize("chunk", [64, 1024])
def test_both_routes_agree_with_the_oracle_when_well_conditioned(whiten, chunk):
    cfg = mixed_config(n=700, seed=2, depth=2)
    cfg["chunk"] = chunk
    ref = build_oracle(cfg)
    m = _sgpr(cfg, whiten_stats=whiten)
    elbo = m.elbo()
    elbo_ref = oo.sgpr_elbo(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    assert abs(elbo - elbo_ref) / abs(elbo_ref) < RTOL
    assert m.last_route == (1 if whiten else 0)
    alpha = m.sufficient_statistics().cpu().numpy()
    alpha_ref = oo.sgpr_alpha(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    assert max_rel_err(alpha, alpha_ref) < 1e-6
    Xn = cfg["X"][:50] + 0.01
    Xn[:, 4:6] = cfg["X"][:50, 4:6]
    mean, var = m.predict_f(Xn)
    mean_ref = oo.sgpr_predict_mean(ref, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"], Xn)
    assert max_rel_err(mean, mean_ref) < 1e-7
    assert np.all(var > -1e-9)
    assert max_rel_err(m.predict_mean(Xn), mean_ref) < 1e-7


def test_factor_buffer_holds_L_and_its_inverse():
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg = mixed_config(n=500, seed=5, depth=3)
    Z = cfg["X"][:203].copy()          # M = 203: not a multiple of 8 (padding rows) nor of 64 (ragged panel)
    k, ref = build_kernel(cfg), build_oracle(cfg)
    spec = k._make_spec()
    try:
        pz = _device.Points(spec, _device.to_device(Z))
        fac = _device.sgpr_factor(spec, pz, 1e-6)
        kuu = ref.K(Z) + 1e-6 * np.eye(len(Z))
        L = np.linalg.cholesky(kuu)
        assert max_rel_err(fac.L().cpu().numpy(), L) < 1e-10
        Linv = fac.Linv().cpu().numpy()
        assert np.all(np.triu(Linv, 1) == 0.0)
        assert np.abs(Linv @ L - np.eye(len(Z))).max() < 1e-7
        h = fac.header().cpu().numpy()
        cond = np.linalg.cond(kuu)
        assert cond / 4 < h[0] < 4 * cond and h[4] == 0
        assert abs(h[5] - np.log(np.diag(L)).sum()) < 1e-9
        assert h[3] == (1.0 if h[0] >= 3e5 else 0.0)
    finally:
        spec.close()


def test_failed_factorisation_is_reported_at_the_single_readback():
    from oak_b200._cabi import OakNativeError

    cfg = mixed_config(n=300, seed=7, depth=2)
    assert np.isfinite(_sgpr(cfg).elbo())
    cfg["X"] = cfg["X"].copy()
    cfg["X"][5, 0] = np.nan              # a NaN coordinate poisons Kuu / Kuf
    cfg["Z"] = cfg["X"][: len(cfg["Z"])].copy()
    with pytest.raises(OakNativeError, match="Cholesky"):
        _sgpr(cfg).elbo()


@pytest.mark.parametrize("route", [0, 1])
@pytest.mark.parametrize("overlap", [-1, 4, 8])
def test_overlapped_factorisation_is_bit_identical_to_the_two_serial_calls(route, overlap):
    """``oak_sgpr_factor_stats_f64``: the factorisation on a side stream next to the first chunk's Kuf tiles
    (fewer CTAs for both kernels) changes the schedule, not one bit of the factor, the statistics or the bound."""
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    cfg = mixed_config(n=2500, seed=11, depth=3)
    Z = cfg["X"][:300].copy()
    k = build_kernel(cfg)
    spec = k._make_spec()
    try:
        pz = _device.Points(spec, _device.to_device(Z))
        px = _device.Points(spec, _device.to_device(cfg["X"]))
        y = _device.to_device(cfg["y"])
        fac0 = _device.sgpr_factor(spec, pz, 1e-6, route=route)
        st0 = _device.sgpr_stats2(spec, pz, px, y, fac0, chunk=1024)
        out0 = _device.sgpr_finish2(fac0, st0, 2500, 0.05).host()
        for _ in range(3):   # repeated: the side stream's events and buffers are reused
            fac1, st1 = _device.sgpr_factor_stats(spec, pz, px, y, 1e-6, route=route, chunk=1024, overlap_ctas=overlap)
            out1 = _device.sgpr_finish2(fac1, st1, 2500, 0.05).host()
            assert torch.equal(fac0.buf[: fac0.ld * fac0.m], fac1.buf[: fac1.ld * fac1.m])
            assert torch.equal(st0, st1)
            assert np.array_equal(out0, out1)
        assert out1[6] == route
    finally:
        spec.close()
