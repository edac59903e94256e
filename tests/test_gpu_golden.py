"""GPU parity against the golden vectors produced by the reference's own sources
(tests/golden/*.npz, see make_golden.py): CUDA path through the C ABI, 1e-9 relative."""
import json

import numpy as np
import pytest

from helpers import max_elem_rel_err, RTOL, load_golden, max_rel_err

pytestmark = pytest.mark.gpu
KERNEL_CASES = ["g1_gaussian_d5_p3", "g2_mixed_p2", "g3_no_share_var", "g4_unconstrained", "g9_full_depth_d8_p8"]


@pytest.mark.parametrize("name", KERNEL_CASES)
@pytest.mark.parametrize("algo", [0, 1])
def test_cuda_kernel_matches_reference_golden(name, algo):
    from oak_b200.oak_kernel import KernelComponenent
    from oak_b200.workloads import build_kernel

    cfg, g = load_golden(name)
    k = build_kernel(cfg)
    k.esp_algorithm = algo
    X, X2 = g["X"], g["X2"]
    # element-wise, not norm-wise: every entry within 1e-9 of the reference's own output
    assert max_elem_rel_err(k.K(X), g["K"]) < RTOL
    assert max_elem_rel_err(k(X, X2), g["K_cross"]) < RTOL
    assert max_elem_rel_err(k(X, full_cov=False), g["K_diag"]) < RTOL
    subsets = json.loads(str(g["subsets_json"]))
    for slot, ci in enumerate(g["component_index"]):
        comp = KernelComponenent(k, subsets[int(ci)], share_var_across_orders=cfg["share_var"])
        assert max_rel_err(comp(X, X2), g["component_K"][slot]) < RTOL
        assert max_rel_err(comp.K_diag(X), g["component_K_diag"][slot]) < RTOL


def test_cuda_single_kernels_match_reference_golden():
    from oak_b200.workloads import build_kernel

    cfgs, g = load_golden("g5_single_kernels")
    for name in ("gaussian", "uniform", "empirical", "mog"):
        sub = build_kernel(cfgs[name]).kernels[0]
        sub.active_dims = [0]
        xin = g["xe"] if name == "empirical" else g["x"]
        assert max_rel_err(sub.K(xin, g["x2"]), g[f"{name}_K"]) < RTOL
        assert max_rel_err(sub.K_diag(xin), g[f"{name}_Kdiag"]) < RTOL
        assert max_rel_err(sub.cov_X_s(xin), g[f"{name}_cov"]) < RTOL
        assert abs(sub.var_s() - float(g[f"{name}_var"])) < RTOL * abs(float(g[f"{name}_var"]))


@pytest.mark.parametrize("tag", ["gpr", "sgpr"])
def test_cuda_models_and_sobol_match_reference_golden(tag):
    from oak_b200.models import GPR, SGPR
    from oak_b200.utils import (compute_L, compute_L_binary_kernel, compute_L_categorical_kernel,
                                compute_sobol_oak, get_model_sufficient_statistics, get_prediction_component)
    from oak_b200.workloads import build_kernel

    cfg, g = load_golden("g6_models_sobol")
    X, Y, Z, Xt, noise = g["X"], g["Y"], g["Z"], g["Xtest"], float(g["noise"])
    k = build_kernel(cfg)
    m = GPR((X, Y), kernel=k) if tag == "gpr" else SGPR((X, Y), kernel=k, inducing_variable=Z)
    m.likelihood.variance.assign(noise)
    alpha = get_model_sufficient_statistics(m, get_L=False)
    assert max_rel_err(alpha, g[f"{tag}_alpha"]) < (1e-8 if tag == "gpr" else 1e-6)  # conditioned by the solves
    obj = m.maximum_log_likelihood_objective()
    assert abs(obj - float(g[f"restated_{tag}_objective"])) < RTOL * abs(float(g[f"restated_{tag}_objective"]))
    idx, sob = compute_sobol_oak(m, 1.0, 0.0)
    assert idx == json.loads(str(g[f"{tag}_sobol_index_json"]))
    assert max_rel_err(sob, g[f"{tag}_sobol"]) < 1e-6
    # with the reference's own alpha the Sobol tiles and component predictions are 1e-9 clean
    comps = get_prediction_component(m, g[f"{tag}_alpha"], Xt)
    assert max_rel_err(np.array(comps), g[f"{tag}_components"]) < RTOL
    mean, _ = m.predict_f(Xt)
    assert max_rel_err(mean, g[f"restated_{tag}_predict_mean"]) < 1e-6
    if tag == "gpr":
        assert max_rel_err(compute_L(X, 1.3, 2.0, 0, 1.0, 0.0), g["L_gaussian"]) < RTOL
        assert max_rel_err(compute_L_binary_kernel(X, 0.6, 2.0, 2), g["L_binary"]) < RTOL
        p3 = np.asarray(cfg["dims"][3]["p"]).reshape(-1, 1)
        Lc = compute_L_categorical_kernel(X, np.array([[0.2, 0.9], [0.7, 0.1], [0.5, 0.6]]), np.array([1.0, 0.8, 1.2]),
                                          p3, 2.0, 3)
        assert max_rel_err(Lc, g["L_categorical"]) < RTOL


def test_cuda_sobol_quadforms_with_reference_alpha():
    """Sobol indices with alpha taken from the golden file: isolates the L tiles + quadratic forms."""
    import torch

    from oak_b200 import _device
    from oak_b200.workloads import build_kernel

    # g7 (empirical measure, variances 1e-3/90/15, lengthscales 2 and 5 on unit-scale data) is
    # ill-conditioned: k~ = k - c c'/v cancels ~4 digits, and the reference's own float64 result
    # (expanded squared distance) sits 1.2e-7 away from an np.longdouble evaluation of the same
    # formulas (1.43422282 vs 1.43422264; see DESIGN.md "conditioning").  The CUDA tiles use the
    # direct (x-y)^2 form and land between the two, so this case is held to 1e-6, not 1e-9.
    # g10: the reference's own oak_model pipeline (binary + categorical + two continuous inputs, k-means inducing
    # points); with ITS alpha the un-normalised indices are held to 1e-9
    for name, key_a, key_s, key_x, tol in (("g6_models_sobol", "sgpr_alpha", "sgpr_sobol", "Z", RTOL),
                                           ("g7_empirical_sobol", "alpha", "sobol", "Z", 1e-6),
                                           ("g10_oak_model_pipeline", "alpha", "sobol_raw", "Z", RTOL)):
        cfg, g = load_golden(name)
        k = build_kernel(cfg)
        spec = k._make_spec()
        Xc = _device.to_device(g[key_x])
        D = len(cfg["dims"])
        Ls = torch.stack([_device.sobol_L(spec, d, Xc, 1.0, 0.0) for d in range(D)])
        from oracle.oak_oracle import subsets as all_subsets

        comps = all_subsets(D, cfg["depth"])[1:]
        scales = []
        for S in comps:
            v = cfg["variances"][len(S)]
            scales.append(v if cfg["dims"][S[0]]["type"] == "binary" else v ** 2)
        sob = _device.sobol_quadforms(Ls, comps, scales, _device.to_device(g[key_a])).cpu().numpy()
        spec.close()
        assert max_rel_err(sob, g[key_s]) < tol


def test_cuda_oak_model_pipeline_matches_the_references_model_utils():
    """g10: predictions (plain and clipped) and normalised Sobol indices of the reference's own oak_model
    (oak/model_utils.py fit / predict / get_sobol over the shim) against this repository's oak_model on the GPU."""
    import warnings

    from oak_b200.model_utils import oak_model

    cfg, g = load_golden("g10_oak_model_pipeline")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oak = oak_model(max_interaction_depth=2, binary_feature=[0], categorical_feature=[1],
                        use_normalising_flow=False, sparse=True, num_inducing=12)
        oak.fit(g["X"], g["Y"], optimise=False)
    # the categorical W is drawn at random in the constructor (ortho_categorical_kernel.py:28): take the reference's
    oak.m.kernel.kernels[1].W.assign(np.asarray(cfg["dims"][1]["W"]))
    assert abs(oak.m.elbo() - float(g["restated_elbo"])) < 1e-8 * abs(float(g["restated_elbo"]))
    assert max_rel_err(oak.predict(g["X_test"]), g["y_pred"]) < 1e-8
    assert max_rel_err(oak.predict(g["X_test"], clip=True), g["y_pred_clip"]) < 1e-8
    sob = oak.get_sobol()
    assert json.loads(str(g["tuple_of_indices_json"])) == [[int(i) for i in t] for t in oak.tuple_of_indices]
    assert max_rel_err(sob, g["sobol"]) < 1e-7


def test_cuda_svgp_classification_chain_matches_the_references_golden():
    """g11: the SVGP model of examples/uci/uci_classification_train.py:108-160 -- the reference's own
    get_model_sufficient_statistics (SVGP branch), compute_sobol_oak and get_prediction_component -- against
    ``models.SVGP`` on the device."""
    from oak_b200._gpflow_shim import Bernoulli, inv_logit
    from oak_b200.models import SVGP
    from oak_b200.utils import compute_sobol_oak, get_model_sufficient_statistics, get_prediction_component
    from oak_b200.workloads import build_kernel

    cfg, g = load_golden("g11_svgp_classification")
    X, Y, Z, Xt = g["X"], g["Y"], g["Z"], g["X_test"]
    assert max_rel_err(inv_logit(np.linspace(-6, 6, 25)), g["inv_logit_of_grid"]) < 1e-15
    m = SVGP(kernel=build_kernel(cfg), likelihood=Bernoulli(invlink=inv_logit), inducing_variable=Z, whiten=True,
             q_diag=True, q_mu=g["q_mu"], q_sqrt=g["q_sqrt"])
    m.data = (X, Y)                          # as the script does before the Sobol step (:145)
    alpha, L = get_model_sufficient_statistics(m, get_L=True)
    assert max_rel_err(alpha, g["alpha"]) < RTOL
    assert max_rel_err(L, g["L"]) < 1e-7    # chol(inv(Qinv)): two inversions on top of cond(Kuu)
    idx, sob = compute_sobol_oak(m, 1.0, 0.0)
    assert idx == json.loads(str(g["sobol_index_json"]))
    assert max_rel_err(sob, g["sobol"]) < RTOL
    comps = get_prediction_component(m, g["alpha"], Xt)
    assert max_rel_err(np.array(comps), g["components"]) < RTOL
    elbo = m.elbo((X, Y))
    assert abs(elbo - float(g["restated_elbo"])) < RTOL * abs(float(g["restated_elbo"]))
    mean, var = m.predict_f(Xt)
    assert max_rel_err(mean, g["restated_predict_mean"]) < RTOL
    assert max_rel_err(var, g["restated_predict_var"]) < RTOL
    assert max_rel_err(m.predict_log_density((Xt, Y[:15])), g["restated_predict_log_density"]) < RTOL
