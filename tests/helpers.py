"""Shared test helpers: build the oracle twin of a workload configuration, relative-error metric."""
import numpy as np

from oracle import oak_oracle as oo

RTOL = 1e-9  # north_star tolerance: relative error vs the reference arithmetic in FP64


def build_oracle(cfg, expanded=False):
    dims = []
    for dc in cfg["dims"]:
        if dc["type"] == "rbf":
            m = dc.get("measure", ("gaussian", 0.0, 1.0))
            if m is None:
                meas = None
            elif m[0] == "gaussian":
                meas = oo.Gaussian(m[1], m[2])
            elif m[0] == "uniform":
                meas = oo.Uniform(m[1], m[2])
            elif m[0] == "empirical":
                meas = oo.Empirical(np.asarray(m[1]).reshape(-1, 1), np.asarray(m[2]).reshape(-1, 1))
            else:
                meas = oo.MOG(m[1], m[2], m[3])
            dims.append(oo.RBFDim(dc["lengthscale"], dc.get("variance", 1.0), meas, expanded=expanded))
        elif dc["type"] == "binary":
            dims.append(oo.BinaryDim(dc["p0"], dc.get("variance", 1.0)))
        else:
            dims.append(oo.CategoricalDim(p=dc["p"], W=dc["W"], kappa=dc["kappa"], variance=dc.get("variance", 1.0)))
    return oo.OakOracle(dims, cfg["depth"], list(cfg["variances"]), cfg.get("share_var", True))


def max_rel_err(a, b):
    """max |a - b| / max(|b|_inf, tiny): error relative to the scale of the reference result."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def max_elem_rel_err(a, b, floor=None):
    """max ELEMENT-WISE relative error |a - b| / max(|b|, floor).  Entries of a constrained-kernel Gram pass through
    zero (k~ changes sign), so entries smaller than `floor` are compared at that absolute scale; the default floor
    is 1e-3 of the largest reference entry -- three orders tighter than the norm-wise ``max_rel_err``."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    if floor is None:
        floor = 1e-3 * max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def mixed_config(n=200, seed=0, depth=3, share=True):
    """Every sub-kernel kind and measure in one kernel (small)."""
    rng = np.random.default_rng(seed)
    X = np.zeros((n, 7))
    X[:, 0] = rng.standard_normal(n)
    X[:, 1] = rng.uniform(0, 1, n)
    X[:, 2] = np.round(4 * rng.standard_normal(n)) / 4
    X[:, 3] = rng.standard_normal(n) * 2 + 1
    X[:, 4] = (rng.random(n) < 0.3).astype(float)
    X[:, 5] = rng.integers(0, 5, n).astype(float)
    X[:, 6] = rng.standard_normal(n)
    loc, cnt = np.unique(X[:, 2], return_counts=True)
    dims = [
        {"type": "rbf", "lengthscale": 0.7, "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)},
        {"type": "rbf", "lengthscale": 0.4, "variance": 1.3, "measure": ("uniform", 0.0, 1.0)},
        {"type": "rbf", "lengthscale": 1.1, "variance": 0.8, "measure": ("empirical", loc, cnt / cnt.sum())},
        {"type": "rbf", "lengthscale": 2.0, "variance": 1.0,
         "measure": ("mog", [3.0, 2.0, -1.0], [3.0, 10.0, 0.5], [0.5, 0.3, 0.2])},
        {"type": "binary", "p0": 0.7, "variance": 1.0},
        {"type": "categorical", "p": np.array([0.1, 0.2, 0.3, 0.15, 0.25]), "W": rng.uniform(0, 1, (5, 2)),
         "kappa": rng.uniform(0.5, 1.5, 5), "variance": 1.2},
        {"type": "rbf", "lengthscale": 1.5, "variance": 2.0, "measure": None},
    ]
    var = [0.5, 1.0, 0.7, 0.3, 0.2, 0.1, 0.05, 0.02, 0.01][: depth + 1] if share else [0.5]
    y = (np.sin(X[:, 0]) + X[:, 4] + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
    return dict(X=X, y=y, Z=X[: max(n // 4, 3)].copy(), dims=dims, depth=depth, variances=var, share_var=share,
                noise=0.05)


def load_golden(name):
    """(cfg, arrays) of a golden file produced by tests/golden/make_golden.py from the reference."""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
    z = np.load(path, allow_pickle=False)
    cfg = json.loads(str(z["cfg"]))

    def fix(c):
        for d in c.get("dims", []):
            if d.get("measure") is not None:
                d["measure"] = tuple(np.asarray(v) if isinstance(v, list) else v for v in d["measure"])
            for key in ("p", "W", "kappa"):
                if key in d:
                    d[key] = np.asarray(d[key], dtype=np.float64)
        return c

    if "dims" in cfg:
        cfg = fix(cfg)
    else:
        cfg = {k: fix(v) for k, v in cfg.items()}
    return cfg, {k: z[k] for k in z.files if k != "cfg"}
