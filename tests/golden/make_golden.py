#!/usr/bin/env python
"""Generates tests/golden/*.npz by executing the REFERENCE'S OWN Python sources.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

TensorFlow / gpflow / TFP are not installable here, so the reference's unmodified modules
(``/root/reference/oak/{oak_kernel,ortho_*_kernel,input_measures,utils}.py``) are imported with
``oracle/tf_shim`` first on ``sys.path``: NumPy FP64 stand-ins for the ``tf.*`` / ``gpflow.*`` calls
those files make.  Everything written in the reference files (constrained kernels, measures,
Newton-Girard, the Sobol ``L`` builders, ``compute_sobol_oak``, ``get_model_sufficient_statistics``,
``get_prediction_component``) therefore runs as shipped; the gpflow pieces (RBF, GPR/SGPR
objectives) are the shim's restatement and are flagged ``restated_*`` in the files.

Each .npz holds the inputs, a JSON configuration in the ``workloads.build_kernel`` format and the
reference outputs.  The vectors are small on purpose (they travel with the repository).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(1, "/root/reference")

import gpflow  # noqa: E402  (the shim)
from oak.input_measures import EmpiricalMeasure, GaussianMeasure, MOGMeasure, UniformMeasure  # noqa: E402
from oak.oak_kernel import KernelComponenent, OAKKernel, get_list_representation  # noqa: E402
from oak.ortho_binary_kernel import OrthogonalBinary  # noqa: E402
from oak.ortho_categorical_kernel import OrthogonalCategorical  # noqa: E402
from oak.ortho_rbf_kernel import OrthogonalRBFKernel  # noqa: E402
from oak import utils as ref_utils  # noqa: E402


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    return o


def save(name, cfg, **arrays):
    np.savez(os.path.join(HERE, name + ".npz"), cfg=json.dumps(jsonable(cfg)), **arrays)
    print("wrote", name, {k: np.shape(v) for k, v in arrays.items()})


def cfg_from_kernel(k: OAKKernel):
    """workloads.build_kernel configuration describing the reference kernel's current parameters."""
    dims = []
    for sub in k.kernels:
        if isinstance(sub, OrthogonalRBFKernel):
            m = sub.measure
            if isinstance(m, GaussianMeasure):
                meas = ("gaussian", float(m.mu), float(m.var))
            elif isinstance(m, UniformMeasure):
                meas = ("uniform", float(m.a), float(m.b))
            elif isinstance(m, EmpiricalMeasure):
                meas = ("empirical", np.asarray(m.location).reshape(-1), np.asarray(m.weights).reshape(-1))
            else:
                meas = ("mog", m.means, m.variances, m.weights)
            dims.append({"type": "rbf", "lengthscale": float(np.squeeze(np.asarray(sub.base_kernel.lengthscales))),
                         "variance": float(np.squeeze(np.asarray(sub.base_kernel.variance))), "measure": meas})
        elif isinstance(sub, gpflow.kernels.RBF):
            dims.append({"type": "rbf", "lengthscale": float(np.squeeze(np.asarray(sub.lengthscales))),
                         "variance": float(np.squeeze(np.asarray(sub.variance))), "measure": None})
        elif isinstance(sub, OrthogonalBinary):
            dims.append({"type": "binary", "p0": float(sub.p0), "variance": float(np.squeeze(np.asarray(sub.variance)))})
        else:
            dims.append({"type": "categorical", "p": np.asarray(sub.p).reshape(-1), "W": np.asarray(sub.W),
                         "kappa": np.asarray(sub.kappa), "variance": float(np.squeeze(np.asarray(sub.variance)))})
    return {"dims": dims, "depth": int(k.max_interaction_depth),
            "variances": [float(v.numpy()) for v in k.variances], "share_var": bool(k.share_var_across_orders)}


def kernel_outputs(k, X, X2):
    D = X.shape[1]
    sel, _ = get_list_representation(k, num_dims=D) if k.share_var_across_orders else (None, None)
    if sel is None:  # the reference's get_list_representation indexes variances[n] (quirk): enumerate by hand
        import itertools
        sel = [[]] + [list(t) for o in range(1, k.max_interaction_depth + 1)
                      for t in itertools.combinations(range(D), o)]
    comps = [KernelComponenent(k, s_, share_var_across_orders=k.share_var_across_orders) for s_ in sel]
    out = dict(K=np.asarray(k(X)), K_cross=np.asarray(k(X, X2)), K_diag=np.asarray(k(X, full_cov=False)))
    pick = [0, 1, len(sel) // 2, len(sel) - 1]
    out["component_index"] = np.array(pick)
    out["component_K"] = np.stack([np.asarray(comps[i](X, X2)) for i in pick])
    out["component_K_diag"] = np.stack([np.asarray(comps[i].K_diag(X)) for i in pick])
    out["subsets_json"] = np.array(json.dumps(jsonable(sel)))
    return out


def main():
    rng = np.random.default_rng(2206)

    # --- G1: Gaussian measure, D=5, depth 3, bounded lengthscales -------------------------------
    D = 5
    X, X2 = rng.standard_normal((40, D)), rng.standard_normal((17, D))
    k = OAKKernel([gpflow.kernels.RBF] * D, num_dims=D, max_interaction_depth=3, constrain_orthogonal=True,
                  lengthscale_bounds=[1e-3, 1e3])
    for sub, l in zip(k.kernels, rng.uniform(0.4, 3.0, D)):
        sub.base_kernel.lengthscales.assign(l)
    for v, s in zip(k.variances, [0.7, 1.3, 0.5, 0.2]):
        v.assign(s)
    save("g1_gaussian_d5_p3", cfg_from_kernel(k), X=X, X2=X2, **kernel_outputs(k, X, X2))

    # --- G2: every sub-kernel kind through the OAKKernel constructor, depth 2 ----------------------
    n = 60
    Xm = np.zeros((n, 6))
    Xm[:, 0] = rng.standard_normal(n)
    Xm[:, 1] = np.round(4 * rng.standard_normal(n)) / 4
    Xm[:, 2] = rng.standard_normal(n) * 1.5 + 0.5
    Xm[:, 3] = (rng.random(n) < 0.35).astype(float)
    Xm[:, 4] = rng.integers(0, 4, n).astype(float)
    Xm[:, 5] = rng.standard_normal(n)
    loc, cnt = np.unique(Xm[:, 1], return_counts=True)
    p_cat = np.array([0.1, 0.2, 0.3, 0.4]).reshape(-1, 1)
    gmm = MOGMeasure(np.array([0.5, -1.0]), np.array([1.5, 0.7]), np.array([0.6, 0.4]))
    k = OAKKernel(
        [gpflow.kernels.RBF, gpflow.kernels.RBF, gpflow.kernels.RBF, None, None, gpflow.kernels.RBF],
        num_dims=6, max_interaction_depth=2, constrain_orthogonal=True,
        p0=[None, None, None, 0.65, None, None], p=[None, None, None, None, p_cat, None],
        empirical_locations=[None, loc.reshape(-1, 1), None, None, None, None],
        empirical_weights=[None, (cnt / cnt.sum()).reshape(-1, 1), None, None, None, None],
        gmm_measures=[None, None, gmm, None, None, None],
    )
    for i, l in zip((0, 1, 2, 5), (0.8, 1.4, 2.2, 0.5)):
        k.kernels[i].base_kernel.lengthscales.assign(l)
    k.kernels[1].base_kernel.variance.assign(1.3)  # empirical dims keep a trainable variance
    k.kernels[4].W.assign(rng.uniform(0, 1, (4, 2)))
    k.kernels[4].kappa.assign(rng.uniform(0.5, 1.5, 4))
    for v, s in zip(k.variances, [0.4, 1.1, 0.6]):
        v.assign(s)
    Xm2 = Xm[rng.permutation(n)[:23]] + 0.0
    Xm2[:, [0, 2, 5]] += 0.1 * rng.standard_normal((23, 3))
    save("g2_mixed_p2", cfg_from_kernel(k), X=Xm, X2=Xm2, **kernel_outputs(k, Xm, Xm2))

    # --- G3: share_var_across_orders=False (trainable base variances) ---------------------------------
    D = 3
    X, X2 = rng.standard_normal((30, D)), rng.standard_normal((11, D))
    k = OAKKernel([gpflow.kernels.RBF] * D, num_dims=D, max_interaction_depth=3, constrain_orthogonal=True,
                  share_var_across_orders=False)
    for sub, l, s2 in zip(k.kernels, (0.6, 1.7, 3.0), (0.5, 1.5, 2.5)):
        sub.base_kernel.lengthscales.assign(l)
        sub.base_kernel.variance.assign(s2)
    k.variances[0].assign(0.3)
    save("g3_no_share_var", cfg_from_kernel(k), X=X, X2=X2, **kernel_outputs(k, X, X2))

    # --- G4: unconstrained additive kernel -------------------------------------------------------------
    k = OAKKernel([gpflow.kernels.RBF] * D, num_dims=D, max_interaction_depth=2, constrain_orthogonal=False)
    for sub, l in zip(k.kernels, (0.9, 1.1, 2.0)):
        sub.lengthscales.assign(l)
    for v, s in zip(k.variances, [0.2, 0.9, 0.4]):
        v.assign(s)
    save("g4_unconstrained", cfg_from_kernel(k), X=X, X2=X2, **kernel_outputs(k, X, X2))

    # --- G5: single constrained kernels with every measure -------------------------------------------------
    x, x2 = rng.uniform(-1, 2, (25, 1)), rng.uniform(-1, 2, (9, 1))
    xe = np.round(rng.standard_normal((25, 1)) * 3) / 3
    le, ce = np.unique(xe, return_counts=True)
    singles = {
        "gaussian": OrthogonalRBFKernel(gpflow.kernels.RBF(lengthscales=0.7, variance=1.6), GaussianMeasure(0.3, 2.0)),
        "uniform": OrthogonalRBFKernel(gpflow.kernels.RBF(lengthscales=0.5, variance=0.8), UniformMeasure(-1.0, 2.0)),
        "empirical": OrthogonalRBFKernel(gpflow.kernels.RBF(lengthscales=1.2),
                                         EmpiricalMeasure(le.reshape(-1, 1), (ce / ce.sum()).reshape(-1, 1))),
        "mog": OrthogonalRBFKernel(gpflow.kernels.RBF(lengthscales=10.0),
                                   MOGMeasure(np.array([3.0, 2.0]), np.array([3.0, 10.0]), np.array([0.6, 0.4]))),
    }
    arrays = dict(x=x, x2=x2, xe=xe)
    for name, sk in singles.items():
        xin = xe if name == "empirical" else x
        arrays[f"{name}_K"] = np.asarray(sk.K(xin, x2))
        arrays[f"{name}_Kdiag"] = np.asarray(sk.K_diag(xin))
        arrays[f"{name}_cov"] = np.asarray(sk.cov_X_s(xin))
        arrays[f"{name}_var"] = np.asarray(sk.var_s())
    cfg5 = {name: cfg_from_kernel(type("K", (), {"kernels": [sk], "max_interaction_depth": 1,
                                                  "variances": [gpflow.Parameter(0.0), gpflow.Parameter(1.0)],
                                                  "share_var_across_orders": True})())
            for name, sk in singles.items()}
    save("g5_single_kernels", cfg5, **arrays)

    # --- G6: model level -- alpha, Sobol, per-component predictions (reference utils.py as shipped) ------------
    n = 90
    Xs = np.zeros((n, 4))
    Xs[:, 0] = rng.standard_normal(n)
    Xs[:, 1] = rng.standard_normal(n)
    Xs[:, 2] = (rng.random(n) < 0.4).astype(float)
    Xs[:, 3] = rng.integers(0, 3, n).astype(float)
    Y = (Xs[:, 0] ** 2 + 2 * Xs[:, 1] + Xs[:, 0] * Xs[:, 2] + 0.3 * Xs[:, 3] + 0.05 * rng.standard_normal(n)).reshape(-1, 1)
    Y = (Y - Y.mean()) / Y.std()
    p3 = np.array([np.mean(Xs[:, 3] == c) for c in range(3)]).reshape(-1, 1)

    def make_kernel():
        kk = OAKKernel([gpflow.kernels.RBF, gpflow.kernels.RBF, None, None], num_dims=4, max_interaction_depth=2,
                       constrain_orthogonal=True, p0=[None, None, float(1 - Xs[:, 2].mean()), None],
                       p=[None, None, None, p3], lengthscale_bounds=[1e-3, 1e3])
        kk.kernels[0].base_kernel.lengthscales.assign(1.3)
        kk.kernels[1].base_kernel.lengthscales.assign(2.4)
        kk.kernels[3].W.assign(np.array([[0.2, 0.9], [0.7, 0.1], [0.5, 0.6]]))
        kk.kernels[3].kappa.assign(np.array([1.0, 0.8, 1.2]))
        for v, s in zip(kk.variances, [0.5, 2.0, 0.7]):
            v.assign(s)
        return kk

    Z = Xs[:30].copy()
    Xtest = Xs[rng.permutation(n)[:20]] + 0.0
    Xtest[:, :2] += 0.2 * rng.standard_normal((20, 2))
    arrays = dict(X=Xs, Y=Y, Z=Z, Xtest=Xtest, noise=np.array(0.05))
    for tag in ("gpr", "sgpr"):
        kk = make_kernel()
        model = gpflow.models.GPR((Xs, Y), kernel=kk) if tag == "gpr" else \
            gpflow.models.SGPR((Xs, Y), kernel=kk, inducing_variable=gpflow.inducing_variables.InducingPoints(Z))
        model.likelihood.variance.assign(0.05)
        alpha = np.asarray(ref_utils.get_model_sufficient_statistics(model, get_L=False))
        idx, sobol = ref_utils.compute_sobol_oak(model, 1.0, 0.0)
        comps = ref_utils.get_prediction_component(model, alpha, Xtest)
        arrays[f"{tag}_alpha"] = alpha
        arrays[f"{tag}_sobol"] = np.array(sobol)
        arrays[f"{tag}_sobol_index_json"] = np.array(json.dumps(jsonable(idx)))
        arrays[f"{tag}_components"] = np.stack([np.asarray(c) for c in comps])
        arrays[f"restated_{tag}_objective"] = np.array(model.maximum_log_likelihood_objective())
        arrays[f"restated_{tag}_predict_mean"] = np.asarray(model.predict_f(Xtest)[0])
        cfg6 = cfg_from_kernel(kk)
    # the individual L builders of the reference
    arrays["L_gaussian"] = ref_utils.compute_L(Xs, 1.3, 2.0, 0, 1.0, 0.0)
    arrays["L_binary"] = ref_utils.compute_L_binary_kernel(Xs, 0.6, 2.0, 2)
    arrays["L_categorical"] = np.asarray(ref_utils.compute_L_categorical_kernel(
        Xs, np.array([[0.2, 0.9], [0.7, 0.1], [0.5, 0.6]]), np.array([1.0, 0.8, 1.2]), p3, 2.0, 3))
    save("g6_models_sobol", cfg6, **arrays)

    # --- G7: empirical-measure Sobol through compute_sobol_oak -----------------------------------------
    n = 50
    Xe = np.round(rng.standard_normal((n, 2)) * 4) / 4
    Ye = (Xe[:, 0] ** 2 + 2 * Xe[:, 1] + Xe[:, 0] * Xe[:, 1]).reshape(-1, 1)
    locs, ws = [], []
    for j in range(2):
        l_, c_ = np.unique(Xe[:, j], return_counts=True)
        locs.append(l_.reshape(-1, 1))
        ws.append((c_ / c_.sum()).reshape(-1, 1))
    kk = OAKKernel([gpflow.kernels.RBF] * 2, num_dims=2, max_interaction_depth=2, constrain_orthogonal=True,
                   empirical_locations=locs, empirical_weights=ws)
    kk.kernels[0].base_kernel.lengthscales.assign(2.0)
    kk.kernels[1].base_kernel.lengthscales.assign(5.0)
    for v, s in zip(kk.variances, [1e-3, 90.0, 15.0]):
        v.assign(s)
    model = gpflow.models.SGPR((Xe, Ye), kernel=kk, inducing_variable=gpflow.inducing_variables.InducingPoints(Xe[:20].copy()))
    model.likelihood.variance.assign(0.01)
    alpha = np.asarray(ref_utils.get_model_sufficient_statistics(model, get_L=False))
    idx, sobol = ref_utils.compute_sobol_oak(model, 1.0, 0.0)
    save("g7_empirical_sobol", cfg_from_kernel(kk), X=Xe, Y=Ye, Z=Xe[:20].copy(), noise=np.array(0.01), alpha=alpha,
         sobol=np.array(sobol), restated_objective=np.array(model.elbo()))

    # --- G8: the normalising flow as written in oak/normalising_flow.py (bijector chain and KL_objective are the
    # reference's; the individual TFP 0.11 bijectors are the shim's restatement) ----------------------------------
    from oak.normalising_flow import Normalizer

    xs = np.exp(0.6 * rng.standard_normal(64)) + 3.0
    out = {"x": xs}
    for tag, log in (("log", True), ("nolog", False)):
        nz = Normalizer(xs, log=log)
        sas, scale, shift = nz.bijector.bijectors[0], nz.bijector.bijectors[1], nz.bijector.bijectors[2]
        for step, bump in enumerate(([0.0, 0.0, 0.0, 0.0], [0.2, -0.3, 0.25, -0.1], [-0.3, 0.4, -0.2, 0.3])):
            # (log scale, shift, skewness, log tailweight) moved away from the initial standardiser
            if step:
                scale.scale.assign(float(scale.scale.numpy()) * np.exp(bump[0]))
                shift.shift.assign(float(shift.shift.numpy()) + bump[1])
                sas.skewness.assign(bump[2])
                sas.tailweight.assign(np.exp(bump[3]))
            out[f"{tag}_theta_{step}"] = np.array([np.log(float(scale.scale.numpy())), float(shift.shift.numpy()),
                                                  float(sas.skewness.numpy()), np.log(float(sas.tailweight.numpy()))])
            out[f"{tag}_y_{step}"] = np.asarray(nz.bijector(xs))
            out[f"{tag}_ldj_{step}"] = np.asarray(nz.bijector.forward_log_det_jacobian(xs, event_ndims=0))
            out[f"{tag}_J_{step}"] = np.array(float(nz.KL_objective()))
        out[f"{tag}_x_back"] = np.asarray(nz.bijector.inverse(nz.bijector(xs)))
    save("g8_normalising_flow", {"note": "oak/normalising_flow.py Normalizer over oracle/tf_shim (TFP 0.11 bijectors restated)"},
         **out)

    # --- G9: full interaction depth (config A's shape: D = 8, depth 8), where Newton-Girard cancels the most ----
    D = 8
    X9, X92 = rng.standard_normal((30, D)), rng.standard_normal((11, D))
    k9 = OAKKernel([gpflow.kernels.RBF] * D, num_dims=D, max_interaction_depth=D, constrain_orthogonal=True)
    for sub, l in zip(k9.kernels, rng.uniform(0.5, 2.5, D)):
        sub.base_kernel.lengthscales.assign(l)
    for v, s_ in zip(k9.variances, [0.5, 1.0, 0.8, 0.6, 0.4, 0.3, 0.2, 0.1, 0.05]):
        v.assign(s_)
    save("g9_full_depth_d8_p8", cfg_from_kernel(k9), X=X9, X2=X92, **kernel_outputs(k9, X9, X92))

    # --- G10: the reference's own oak_model pipeline (oak/model_utils.py fit / predict / get_sobol, unmodified):
    # feature typing, p0 / p, standardisation of X and y, k-means inducing points with discrete columns, kernel
    # construction.  Flows off (their optimiser is gpflow's), optimise=False (BFGS is gpflow's); the prediction
    # goes through the shim's restated SGPR.predict_f. ---------------------------------------------------------
    import warnings

    from oak import model_utils as ref_mu

    warnings.filterwarnings("ignore")
    N = 90
    Xp = np.zeros((N, 4))
    Xp[:, 0] = (rng.random(N) < 0.3).astype(float)
    Xp[:, 1] = rng.integers(0, 3, N).astype(float)
    Xp[:, 2] = rng.standard_normal(N) * 2.0 + 1.0
    Xp[:, 3] = rng.standard_normal(N) * 0.5 - 3.0
    Yp = (np.sin(Xp[:, 2]) + Xp[:, 0] + 0.5 * (Xp[:, 1] == 2) + 0.3 * Xp[:, 3] + 0.1 * rng.standard_normal(N)).reshape(-1, 1)
    Xt = Xp[:25] + np.array([0.0, 0.0, 0.3, -0.1])
    np.random.seed(7)
    oak = ref_mu.oak_model(max_interaction_depth=2, binary_feature=[0], categorical_feature=[1],
                           use_normalising_flow=False, sparse=True, num_inducing=12)
    oak.fit(Xp, Yp, optimise=False)
    sob = oak.get_sobol()
    # alpha and the un-normalised indices of the same model: lets the CUDA Sobol tiles be held to 1e-9 with the
    # reference's own alpha (the indices computed from the product's alpha carry cond(Kuu))
    alpha10 = np.asarray(ref_utils.get_model_sufficient_statistics(oak.m, get_L=False))
    _, sob_raw = ref_utils.compute_sobol_oak(oak.m, 1, 0, share_var_across_orders=True)
    save("g10_oak_model_pipeline", cfg_from_kernel(oak.m.kernel), X=Xp, Y=Yp, X_test=Xt, alpha=alpha10,
         sobol_raw=np.asarray(sob_raw, dtype=np.float64),
         X_scaled=np.asarray(oak.X_scaled), Y_scaled=np.asarray(oak.Y_scaled),
         Z=np.asarray(oak.m.inducing_variable.Z.numpy()), noise=np.array(float(oak.m.likelihood.variance.numpy())),
         y_pred=np.asarray(oak.predict(Xt)), y_pred_clip=np.asarray(oak.predict(Xt, clip=True)),
         sobol=np.asarray(sob), tuple_of_indices_json=np.array(json.dumps(jsonable(oak.tuple_of_indices))),
         restated_elbo=np.array(float(oak.m.elbo())))

    # --- G11: the classification chain of examples/uci/uci_classification_train.py:108-160 -- the reference's own
    # get_model_sufficient_statistics (SVGP branch, utils.py:174-179), compute_sobol_oak (:361) and
    # get_prediction_component (:514) on a whitened, diagonal-q SVGP with the script's own inv_logit.  gpflow's
    # SVGP / Bernoulli / posterior are the shim's restatement of gpflow 2.2.1 (flagged restated_*). -----------
    import ast

    script = open("/root/reference/examples/uci/uci_classification_train.py").read()
    fn = next(n for n in ast.parse(script).body if isinstance(n, ast.FunctionDef) and n.name == "inv_logit")
    import tensorflow as tf_shim  # noqa: E402  (the shim)

    ns = {"tf": tf_shim}
    exec(compile(ast.Module([fn], []), "uci_classification_train.py", "exec"), ns)   # the script's own definition
    inv_logit = ns["inv_logit"]
    n, m11 = 70, 16
    Xc = np.column_stack([rng.standard_normal(n), rng.standard_normal(n) * 0.7 + 0.2, (rng.random(n) < 0.4).astype(float)])
    Yc = (rng.random(n) < 1 / (1 + np.exp(-(1.5 * np.sin(Xc[:, 0]) + Xc[:, 1] * (Xc[:, 2] - 0.5))))).astype(float)[:, None]
    k11 = OAKKernel([gpflow.kernels.RBF, gpflow.kernels.RBF, None], num_dims=3, max_interaction_depth=3,
                    constrain_orthogonal=True, p0=[None, None, float(1 - Xc[:, 2].mean())], p=[None, None, None],
                    lengthscale_bounds=[1e-3, 1e3])
    k11.kernels[0].base_kernel.lengthscales.assign(0.9)
    k11.kernels[1].base_kernel.lengthscales.assign(1.7)
    for v, s_ in zip(k11.variances, [0.3, 1.2, 0.6, 0.2]):
        v.assign(s_)
    Z11 = Xc[rng.permutation(n)[:m11]].copy()
    Z11[:, :2] += 0.05 * rng.standard_normal((m11, 2))
    svgp = gpflow.models.SVGP(kernel=k11, likelihood=gpflow.likelihoods.Bernoulli(invlink=inv_logit),
                              inducing_variable=Z11, whiten=True, q_diag=True)
    svgp.q_mu.assign(0.8 * rng.standard_normal((m11, 1)))
    svgp.q_sqrt.assign(rng.uniform(0.3, 0.9, (m11, 1)))
    svgp.data = (Xc, Yc)                                  # as the script does before the Sobol step (:145)
    Xt11 = Xc[:15] + np.array([0.2, -0.1, 0.0])
    alpha11, L11 = ref_utils.get_model_sufficient_statistics(svgp, get_L=True)
    idx11, sobol11 = ref_utils.compute_sobol_oak(svgp, 1.0, 0.0)
    comps11 = ref_utils.get_prediction_component(svgp, np.asarray(alpha11), Xt11)
    fm, fv = svgp.predict_f(Xt11)
    save("g11_svgp_classification", cfg_from_kernel(k11), X=Xc, Y=Yc, Z=Z11, X_test=Xt11,
         q_mu=svgp.q_mu.numpy(), q_sqrt=svgp.q_sqrt.numpy(), alpha=np.asarray(alpha11), L=np.asarray(L11),
         sobol=np.asarray(sobol11, dtype=np.float64), sobol_index_json=np.array(json.dumps(jsonable(idx11))),
         components=np.stack([np.asarray(c) for c in comps11]),
         restated_elbo=np.array(svgp.elbo((Xc, Yc))), restated_predict_mean=np.asarray(fm),
         restated_predict_var=np.asarray(fv),
         restated_predict_log_density=np.asarray(svgp.predict_log_density((Xt11, Yc[:15]))),
         inv_logit_of_grid=np.asarray(inv_logit(np.linspace(-6, 6, 25))))


if __name__ == "__main__":
    main()
