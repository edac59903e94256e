"""Pins the NumPy oracle against the golden vectors produced by the reference's own sources
(tests/golden/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from helpers import build_oracle, load_golden, max_rel_err
from oracle import oak_oracle as oo

TIGHT = 1e-12  # same formulas, same op order (expanded RBF): rounding-level agreement
KERNEL_CASES = ["g1_gaussian_d5_p3", "g2_mixed_p2", "g3_no_share_var", "g4_unconstrained", "g9_full_depth_d8_p8"]


@pytest.mark.parametrize("name", KERNEL_CASES)
def test_oracle_kernel_matches_reference(name):
    cfg, g = load_golden(name)
    ref = build_oracle(cfg, expanded=True)
    X, X2 = g["X"], g["X2"]
    assert max_rel_err(ref.K(X), g["K"]) < TIGHT
    assert max_rel_err(ref.K(X, X2), g["K_cross"]) < TIGHT
    assert max_rel_err(ref.K_diag(X), g["K_diag"]) < TIGHT
    subsets = json.loads(str(g["subsets_json"]))
    assert subsets == oo.subsets(X.shape[1], cfg["depth"])
    for slot, ci in enumerate(g["component_index"]):
        S = subsets[int(ci)]
        assert max_rel_err(ref.component_K(S, X, X2), g["component_K"][slot]) < TIGHT
        assert max_rel_err(ref.component_K_diag(S, X), g["component_K_diag"][slot]) < TIGHT
    # the direct (x-y)^2 form used by the CUDA tiles agrees with the reference's expanded form
    assert max_rel_err(build_oracle(cfg, expanded=False).K(X, X2), g["K_cross"]) < 1e-11


def test_oracle_single_kernels_match_reference():
    cfgs, g = load_golden("g5_single_kernels")
    for name in ("gaussian", "uniform", "empirical", "mog"):
        k = build_oracle(cfgs[name], expanded=True).dims[0]
        xin = g["xe"] if name == "empirical" else g["x"]
        assert max_rel_err(k.K(xin, g["x2"]), g[f"{name}_K"]) < TIGHT
        assert max_rel_err(k.K_diag(xin), g[f"{name}_Kdiag"]) < TIGHT
        assert max_rel_err(k.cov_X_s(xin), g[f"{name}_cov"]) < TIGHT
        assert abs(k.var_s() - float(g[f"{name}_var"])) < TIGHT * abs(float(g[f"{name}_var"]))


def test_oracle_models_and_sobol_match_reference():
    cfg, g = load_golden("g6_models_sobol")
    ref = build_oracle(cfg, expanded=True)
    X, Y, Z, Xt, noise = g["X"], g["Y"], g["Z"], g["Xtest"], float(g["noise"])
    a_gpr = oo.gpr_alpha(ref, X, Y, noise)
    a_sgpr = oo.sgpr_alpha(ref, X, Y, Z, noise)
    assert max_rel_err(a_gpr, g["gpr_alpha"]) < 1e-9
    assert max_rel_err(a_sgpr, g["sgpr_alpha"]) < 1e-7
    idx, sob = oo.sobol_oak(ref, X, g["gpr_alpha"])
    assert idx == json.loads(str(g["gpr_sobol_index_json"]))
    assert max_rel_err(sob, g["gpr_sobol"]) < 1e-10
    _, sob = oo.sobol_oak(ref, Z, g["sgpr_alpha"])
    assert max_rel_err(sob, g["sgpr_sobol"]) < 1e-10
    assert max_rel_err(np.array(oo.predict_components(ref, X, g["gpr_alpha"], Xt)), g["gpr_components"]) < 1e-10
    assert max_rel_err(np.array(oo.predict_components(ref, Z, g["sgpr_alpha"], Xt)), g["sgpr_components"]) < 1e-10
    assert abs(oo.gpr_log_marginal_likelihood(ref, X, Y, noise) - float(g["restated_gpr_objective"])) < 1e-9
    assert abs(oo.sgpr_elbo(ref, X, Y, Z, noise) - float(g["restated_sgpr_objective"])) < 1e-8
    assert max_rel_err(oo.gpr_predict_mean(ref, X, Y, noise, Xt), g["restated_gpr_predict_mean"]) < 1e-9
    assert max_rel_err(oo.sgpr_predict_mean(ref, X, Y, Z, noise, Xt), g["restated_sgpr_predict_mean"]) < 1e-8
    # individual L builders
    assert max_rel_err(oo.L_gaussian(X[:, 0], 1.3, 2.0, 1.0, 0.0), g["L_gaussian"]) < TIGHT
    assert max_rel_err(oo.L_binary(X[:, 2], 0.6, 2.0), g["L_binary"]) < TIGHT
    p3 = np.asarray(cfg["dims"][3]["p"]).reshape(-1, 1)
    assert max_rel_err(oo.L_categorical(X[:, 3], np.array([[0.2, 0.9], [0.7, 0.1], [0.5, 0.6]]),
                                        np.array([1.0, 0.8, 1.2]), p3, 2.0), g["L_categorical"]) < TIGHT


def test_oracle_empirical_sobol_matches_reference():
    cfg, g = load_golden("g7_empirical_sobol")
    ref = build_oracle(cfg, expanded=True)
    _, sob = oo.sobol_oak(ref, g["Z"], g["alpha"])
    assert max_rel_err(sob, g["sobol"]) < 1e-10
    assert max_rel_err(oo.sgpr_alpha(ref, g["X"], g["Y"], g["Z"], float(g["noise"])), g["alpha"]) < 1e-6
    assert abs(oo.sgpr_elbo(ref, g["X"], g["Y"], g["Z"], float(g["noise"])) - float(g["restated_objective"])) < 1e-7 * abs(float(g["restated_objective"]))


def test_flow_oracle_matches_the_references_normalizer():
    """g8: oak/normalising_flow.py's own Normalizer (bijector chain order, offset, standardiser initialisation,
    KL_objective) executed over the TF/TFP shim, at the initial and two perturbed parameter settings, with and
    without the log step: oracle/flow_oracle.py reproduces y, log|dy/dx| and J; the product's host-side inverse
    round-trips."""
    import os

    from oak_b200.normalising_flow import Normalizer
    from oracle import flow_oracle as fo

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g8_normalising_flow.npz"))
    x = g["x"]
    for tag, log in (("log", True), ("nolog", False)):
        nz = Normalizer(x, log=log)
        offset = x.min() - 1.0 if log else 0.0
        assert nz.offset == offset
        # the product initialises the standardiser exactly as the reference does (:23-27)
        th0 = g[f"{tag}_theta_0"]
        np.testing.assert_allclose(nz._theta(), th0, rtol=1e-14, atol=1e-15)
        for step in range(3):
            th = g[f"{tag}_theta_{step}"]
            par = (offset, log, th[1], np.exp(th[0]), th[2], np.exp(th[3]))
            np.testing.assert_allclose(fo.forward(x, *par), g[f"{tag}_y_{step}"], rtol=1e-13, atol=1e-15)
            np.testing.assert_allclose(fo.forward_log_det_jacobian(x, *par), g[f"{tag}_ldj_{step}"], rtol=1e-13, atol=1e-14)
            J = fo.kl_objective_and_grad(x, offset, log, th)[0]
            assert abs(J - float(g[f"{tag}_J_{step}"])) < 1e-13 * max(1.0, abs(J))
        th = g[f"{tag}_theta_2"]
        for prm, v in zip((nz.scale, nz.shift, nz.skewness, nz.tailweight), th):
            prm.unconstrained_variable = np.asarray(v)
        np.testing.assert_allclose(nz.bijector.inverse(g[f"{tag}_y_2"]), x, rtol=1e-11)
        np.testing.assert_allclose(nz.bijector.forward_log_det_jacobian(x), g[f"{tag}_ldj_2"], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(g[f"{tag}_x_back"], x, rtol=1e-11)


def test_oak_model_pipeline_matches_the_references_model_utils():
    """g10: the reference's own oak_model (oak/model_utils.py fit, unmodified, over the shim) against this
    repository's: feature typing and the discrete measures p0 / p (:703-750), standardisation of X and y
    (:318-331), k-means inducing points with discrete columns (utils.py:555-574), kernel construction and the
    0.01 initial noise (:90-176).  Runs without a GPU: with flows off and optimise=False nothing is evaluated."""
    import warnings

    from oak_b200.model_utils import oak_model

    cfg, g = load_golden("g10_oak_model_pipeline")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oak = oak_model(max_interaction_depth=2, binary_feature=[0], categorical_feature=[1],
                        use_normalising_flow=False, sparse=True, num_inducing=12)
        oak.fit(g["X"], g["Y"], optimise=False)
    assert (oak.binary_index, oak.categorical_index, oak.continuous_index) == ([0], [1], [2, 3])
    assert max_rel_err(oak.X_scaled, g["X_scaled"]) < 1e-14
    assert max_rel_err(oak.Y_scaled, g["Y_scaled"]) < 1e-14
    assert max_rel_err(oak.m.inducing_variable.Z.numpy(), g["Z"]) < 1e-12
    assert abs(float(oak.m.likelihood.variance.numpy()) - float(g["noise"])) < 1e-15
    k = oak.m.kernel
    assert k.max_interaction_depth == cfg["depth"] and [float(v.numpy()) for v in k.variances] == cfg["variances"]
    assert abs(k.kernels[0].p0 - cfg["dims"][0]["p0"]) < 1e-15
    assert max_rel_err(k.kernels[1]._p_vector(), cfg["dims"][1]["p"]) < 1e-15
    for i in (2, 3):
        assert abs(float(k.kernels[i].base_kernel.lengthscales.numpy()) - cfg["dims"][i]["lengthscale"]) < 1e-14
        assert (k.kernels[i].measure.mu, k.kernels[i].measure.var) == (0.0, 1.0)


def test_oracle_svgp_chain_matches_the_references_classification_path():
    """g11: the reference's get_model_sufficient_statistics (SVGP branch), compute_sobol_oak and
    get_prediction_component on a whitened diagonal-q SVGP; the torch oracle that checks the CUDA gradients is
    held to the same vectors."""
    import torch

    from oracle import oak_grad_oracle as go

    cfg, g = load_golden("g11_svgp_classification")
    ref = build_oracle(cfg, expanded=True)
    X, Y, Z, Xt, q_mu, q_sqrt = g["X"], g["Y"], g["Z"], g["X_test"], g["q_mu"], g["q_sqrt"]
    assert max_rel_err(oo.inv_logit(np.linspace(-6, 6, 25)), g["inv_logit_of_grid"]) < TIGHT
    assert max_rel_err(oo.svgp_alpha(ref, Z, q_mu), g["alpha"]) < 1e-10
    assert max_rel_err(oo.svgp_L(ref, Z, q_sqrt), g["L"]) < 1e-8
    idx, sob = oo.sobol_oak(ref, Z, g["alpha"])
    assert idx == json.loads(str(g["sobol_index_json"]))
    assert max_rel_err(sob, g["sobol"]) < 1e-10
    assert max_rel_err(np.array(oo.predict_components(ref, Z, g["alpha"], Xt)), g["components"]) < 1e-10
    fm, fv = oo.svgp_predict_f(ref, Z, q_mu, q_sqrt, Xt)
    assert max_rel_err(fm, g["restated_predict_mean"][:, 0]) < 1e-10
    assert max_rel_err(fv, g["restated_predict_var"][:, 0]) < 1e-10
    assert abs(oo.svgp_elbo(ref, X, Y, Z, q_mu, q_sqrt) - float(g["restated_elbo"])) < 1e-10 * abs(float(g["restated_elbo"]))
    assert max_rel_err(oo.svgp_predict_log_density(ref, Z, q_mu, q_sqrt, Xt, Y[:15]), g["restated_predict_log_density"]) < 1e-10
    # constant term + components = latent mean (uci_classification_train.py:171-184)
    const = g["alpha"].sum() * cfg["variances"][0]
    assert max_rel_err(const + g["components"].sum(0), g["restated_predict_mean"][:, 0]) < 1e-9
    # the torch oracle (continuous dims only) on the two continuous columns of the same case
    cfg2 = dict(cfg)
    cfg2["dims"] = cfg["dims"][:2]
    ref2 = build_oracle(cfg2, expanded=False)
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
    ls = [d["lengthscale"] for d in cfg2["dims"]]
    a_t = go.svgp_alpha(t(Z[:, :2]), t(ls), t(cfg["variances"]), t(q_mu)).numpy()
    assert max_rel_err(a_t, oo.svgp_alpha(ref2, Z[:, :2], q_mu)) < 1e-9
    e_t = float(go.svgp_elbo(t(X[:, :2]), t(Y), t(Z[:, :2]), t(ls), t(cfg["variances"]), t(q_mu)[:, 0], t(q_sqrt)[:, 0]))
    assert abs(e_t - oo.svgp_elbo(ref2, X[:, :2], Y, Z[:, :2], q_mu, q_sqrt)) < 1e-9 * abs(e_t)
