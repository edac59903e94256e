"""Model factory and the ``oak_model`` convenience API.

Drop-in for the hot-path entry points of the reference's ``oak/model_utils.py``:
``create_model_oak`` (:90-176) and ``oak_model.fit / predict / get_sobol`` (:194-524), with the same
argument names, defaults and error behaviour, the BFGS training step (``training.py``: gradients from
the backward tiles), ``save_model`` / ``load_model``.  The per-column normalising flow runs on the device
(``normalising_flow.py`` / csrc/oak_flow.cu: objective passes under scipy L-BFGS-B instead of TFP, forward
transform); the remaining one-off preprocessing stays on the host as in the reference: sklearn k-means
for the inducing points and sklearn Gaussian mixtures for the MOG measures; plotting is out of scope.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import _device
from ._gpflow_shim import Gamma, InducingPoints, set_trainable
from .input_measures import MOGMeasure
from .models import GPR, SGPR
from .oak_kernel import OAKKernel, get_list_representation
from .ortho_rbf_kernel import RBF
from .utils import compute_sobol_oak


def get_kmeans_centers(X: np.ndarray, K: int = 500) -> np.ndarray:
    """K-means centres used as inducing points (model_utils.py:31-41): scikit-learn's algorithm on the device
    (``kmeans.KMeans``)."""
    from .kmeans import kmeans_class

    np.random.seed(44)
    return kmeans_class()(n_clusters=K, random_state=0).fit(X).cluster_centers_


def save_model(model, filename) -> None:
    """Parameter values as one object array ``hyperparams`` in an ``.npz`` (model_utils.py:44-63): the trainable
    ones in ``model.trainable_parameters`` order (gpflow's tf.Module order, reproduced by
    ``_gpflow_shim.collect_parameters``), all of them for an SVGP."""
    import os
    from pathlib import Path

    from .models import SVGP

    filename = Path(filename)
    params = model.parameters if isinstance(model, SVGP) else model.trainable_parameters
    hyperparams = np.empty(len(params), dtype=object)
    for i, p in enumerate(params):
        hyperparams[i] = p.numpy()
    os.makedirs(filename.parents[0], exist_ok=True)
    np.savez(filename, hyperparams=hyperparams)


def load_model(model, filename, load_all_parameters: bool = False) -> None:
    """Inverse of ``save_model`` (model_utils.py:66-87)."""
    model_params = np.load(str(filename), allow_pickle=True)["hyperparams"]
    target = model.parameters if load_all_parameters else model.trainable_parameters
    if len(model_params) < len(target):
        raise ValueError(f"checkpoint holds {len(model_params)} parameters, the model needs {len(target)}")
    for i, p in enumerate(target):
        p.assign(model_params[i])


def create_model_oak(
    data,
    max_interaction_depth: int = 2,
    constrain_orthogonal: bool = True,
    inducing_pts: np.ndarray = None,
    optimise=False,
    zfixed=True,
    p0=None,
    p=None,
    lengthscale_bounds=None,
    empirical_locations: Optional[List[float]] = None,
    empirical_weights: Optional[List[float]] = None,
    use_sparsity_prior: bool = True,
    gmm_measures: Optional[List[MOGMeasure]] = None,
    share_var_across_orders: Optional[bool] = True,
):
    """GPR (or SGPR when ``inducing_pts`` is given) with an OAK kernel (model_utils.py:90-176)."""
    num_dims = np.shape(data[0])[1]
    if p0 is None:
        p0 = [None] * num_dims
    if p is None:
        p = [None] * num_dims
    base_kernels = [None] * num_dims
    for dim in range(num_dims):
        if (p0[dim] is None) and (p[dim] is None):
            base_kernels[dim] = RBF

    k = OAKKernel(
        base_kernels,
        num_dims=num_dims,
        max_interaction_depth=max_interaction_depth,
        constrain_orthogonal=constrain_orthogonal,
        p0=p0,
        p=p,
        lengthscale_bounds=lengthscale_bounds,
        empirical_locations=empirical_locations,
        empirical_weights=empirical_weights,
        gmm_measures=gmm_measures,
        share_var_across_orders=share_var_across_orders,
    )
    if inducing_pts is not None:
        model = SGPR(data, kernel=k, inducing_variable=InducingPoints(inducing_pts))
        if zfixed:
            set_trainable(model.inducing_variable, False)
    else:
        model = GPR(data, kernel=k)
    if use_sparsity_prior and share_var_across_orders:
        for prm in model.kernel.variances:  # Gamma(1, rate 0.2) prior (:163-165)
            prm.prior = Gamma(1.0, 0.2)
    # small initial noise to avoid the all-noise optimum (:167)
    model.likelihood.variance.assign(0.01)
    if optimise:
        # gpflow.optimizers.Scipy().minimize(training_loss_closure, trainable_variables, method="BFGS")
        # (model_utils.py:168-173): same optimiser, gradients from the backward tiles
        from .training import optimise as _optimise

        _optimise(model, method="BFGS")
    return model


def apply_normalise_flow(X, input_flows):
    """Each column through its flow, untouched where there is none (model_utils.py:179-191): the matrix goes
    to the device once and its flowed columns are rewritten in place by ``oak_flow_forward_f64`` (strided
    columns of the row-major layout the Gram tiles read).  NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor."""
    from . import _device

    if input_flows is None or all(f is None for f in input_flows):
        return np.array(X, dtype=np.float64) if _device.is_host(X) else X.clone()
    host = _device.is_host(X)
    Xd = _device.to_device(X)
    if not host:
        Xd = Xd.clone()
    for ii in range(Xd.shape[1]):
        f = input_flows[ii]
        if f is not None:
            col = Xd[:, ii]
            _device.flow_forward(col, f.offset, f.log, float(f.shift.numpy()), float(f.scale.numpy()),
                                 float(f.skewness.numpy()), float(f.tailweight.numpy()), out=col)
    return Xd.cpu().numpy() if host else Xd


def estimate_one_dim_gmm(K: int, X: np.ndarray) -> MOGMeasure:
    """Spherical Gaussian mixture fitted to one input column (model_utils.py:753-770): scikit-learn's EM on the device
    (``gmm.GaussianMixture1D``); the library's own class only on a machine without a CUDA device."""
    from .gmm import GaussianMixture1D, on_device

    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 1:
        raise ValueError(f"expected a 1-D array, got shape {X.shape}")
    assert K > 0
    if on_device():
        gm = GaussianMixture1D(n_components=int(K), random_state=0).fit(X)
    else:
        from sklearn.mixture import GaussianMixture

        gm = GaussianMixture(n_components=int(K), random_state=0, covariance_type="spherical").fit(X.reshape(-1, 1))
    assert np.allclose(gm.weights_.sum(), 1.0)
    return MOGMeasure(weights=gm.weights_, means=gm.means_.reshape(-1), variances=gm.covariances_)


class _ColumnStats:
    """Per-column statistics of the input pipeline (distinct values + counts, means).  With a CUDA device the matrix
    is moved to HBM once and every column is summarised there (csrc/oak_unique.cu: sort + run-length on
    order-preserving keys, bit-identical to NumPy); without one -- the CPU-only unit tests of this host logic -- the
    reference's own NumPy calls run.  This is one-off preprocessing, not the hot path."""

    def __init__(self, X):
        from . import _device

        self.X = np.asarray(X, dtype=np.float64)
        self.Xd = _device.to_device(self.X) if _device.cuda_available() else None

    def unique(self, j):
        if self.Xd is not None:
            from . import _device

            return _device.column_unique(self.Xd, j)
        return np.unique(self.X[:, j], return_counts=True)

    def mean(self, j):
        if self.Xd is not None:
            from . import _device

            return _device.column_mean(self.Xd, j)
        return self.X[:, j].mean()


def _calculate_features(X, categorical_feature, binary_feature):
    """Feature typing and discrete measures (model_utils.py:703-750): p0 = 1 - mean, p = frequencies."""
    if binary_feature is None and categorical_feature is None:
        return list(range(X.shape[1])), [], [], None, None
    stats = _ColumnStats(X)
    if binary_feature is not None and categorical_feature is not None:
        overlap = set(binary_feature).intersection(categorical_feature)
        if len(overlap) > 0:
            raise ValueError(f"Overlapping feature set {overlap}")
    binary_index, categorical_index, continuous_index, p0, p = [], [], [], [], []
    for j in range(X.shape[1]):
        if binary_feature is not None and j in binary_feature:
            p0.append(1 - stats.mean(j))
            p.append(None)
            binary_index.append(j)
        elif categorical_feature is not None and j in categorical_feature:
            p0.append(None)
            vals, counts = stats.unique(j)
            p.append((counts / len(X[:, j])).reshape(-1, 1))
            assert np.abs(p[-1].sum() - 1) < 1e-6
            categorical_index.append(j)
        else:
            p.append(None)
            p0.append(None)
            continuous_index.append(j)
    return continuous_index, binary_index, categorical_index, p0, p


class _Standardiser:
    def fit(self, A):
        self.mean_ = A.mean(0)
        self.var_ = A.var(0)
        self.scale_ = np.where(self.var_ > 0, np.sqrt(self.var_), 1.0)
        return self

    def transform(self, A):
        return (A - self.mean_) / self.scale_

    def inverse_transform(self, A):
        return A * self.scale_ + self.mean_


class oak_model:
    """OAK model with fitting, prediction and attribution utilities (model_utils.py:194-524)."""

    def __init__(
        self,
        max_interaction_depth=2,
        num_inducing=200,
        lengthscale_bounds=[1e-3, 1e3],
        binary_feature: Optional[List[int]] = None,
        categorical_feature: Optional[List[int]] = None,
        empirical_measure: Optional[List[int]] = None,
        use_sparsity_prior: bool = True,
        gmm_measure: Optional[List[int]] = None,
        sparse: bool = False,
        use_normalising_flow: bool = True,
        share_var_across_orders: bool = True,
    ):
        self.max_interaction_depth = max_interaction_depth
        self.num_inducing = num_inducing
        self.lengthscale_bounds = lengthscale_bounds
        self.binary_feature = binary_feature
        self.categorical_feature = categorical_feature
        self.use_sparsity_prior = use_sparsity_prior
        self.input_flows = None
        self.scaler_y = None
        self.Y_scaled = None
        self.X_scaled = None
        self.alpha = None
        self.continuous_index = None
        self.binary_index = None
        self.categorical_index = None
        self.empirical_measure = empirical_measure
        self.empirical_locations = None
        self.empirical_weights = None
        self.gmm_measure = gmm_measure
        self.estimated_gmm_measures = None
        self.sparse = sparse
        self.use_normalising_flow = use_normalising_flow
        self.share_var_across_orders = share_var_across_orders

    def fit(self, X, Y, optimise: bool = True, initialise_inducing_points: bool = True):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        self.xmin, self.xmax = X.min(0), X.max(0)
        self.num_dims = X.shape[1]
        (self.continuous_index, self.binary_index, self.categorical_index, p0, p) = _calculate_features(
            X, categorical_feature=self.categorical_feature, binary_feature=self.binary_feature
        )
        if self.empirical_measure is not None:
            if not set(self.empirical_measure).issubset(self.continuous_index):
                raise ValueError(
                    f"Empirical measure={self.empirical_measure} should only be used on non-binary/categorical "
                    f"inputs {self.continuous_index}"
                )
        if self.gmm_measure is not None:
            if len(self.gmm_measure) != self.num_dims:
                # the reference *returns* this error instead of raising it (:283); kept
                return ValueError(f"Must specify number of components for each inputs dimension 1..{X.shape[0]}")
            idx_gmm = np.flatnonzero(self.gmm_measure)
            if not set(idx_gmm).issubset(self.continuous_index):
                raise ValueError(
                    f"GMM measure on inputs {idx_gmm} should only be used on continuous inputs {self.continuous_index}"
                )
        self.estimated_gmm_measures = [None] * self.num_dims
        if self.gmm_measure is not None:  # one-off host preprocessing (:293-300)
            for i_dim in np.flatnonzero(self.gmm_measure):
                self.estimated_gmm_measures[i_dim] = estimate_one_dim_gmm(K=self.gmm_measure[i_dim], X=X[:, i_dim])
        self.empirical_locations = [None] * self.num_dims
        self.empirical_weights = [None] * self.num_dims
        self.input_flows = [None] * self.num_dims
        flow_dims = [i for i in self.continuous_index
                     if not (self.empirical_measure is not None and i in self.empirical_measure)
                     and self.estimated_gmm_measures[i] is None]  # (:306-311)
        if self.use_normalising_flow:  # one flow per remaining continuous input (:305-317), fitted on the device
            from .normalising_flow import Normalizer

            for i in flow_dims:
                n = Normalizer(X[:, i])
                n.fit()
                self.input_flows[i] = n
        self.alpha = None
        self.scaler_y = _Standardiser().fit(Y)
        self.Y_scaled = self.scaler_y.transform(Y)
        if self.empirical_measure is not None:
            self.scaler_X_empirical = _Standardiser().fit(X[:, self.empirical_measure])
        if not self.use_normalising_flow:
            self.scaler_X_continuous = _Standardiser().fit(X[:, self.continuous_index])
        self.X_scaled = self._transform_x(X)

        # empirical locations / weights from the scaled data (:334-344)
        if self.empirical_measure is not None:
            stats = _ColumnStats(self.X_scaled)
            for ii in self.empirical_measure:
                loc, cnt = stats.unique(ii)
                self.empirical_weights[ii] = (cnt / cnt.sum()).reshape(-1, 1)
                self.empirical_locations[ii] = loc.reshape(-1, 1)

        # the reference's consistency checks (:346-370).  The last one also rejects use_normalising_flow=False
        # together with an empirical measure: such a column is standardised twice by _transform_x (:468-475)
        assert np.allclose(self.X_scaled[:, self.binary_index], X[:, self.binary_index]), "Flow applied to binary inputs"
        assert np.allclose(self.X_scaled[:, self.categorical_index],
                           X[:, self.categorical_index]), "Flow applied to categorical inputs"
        if self.gmm_measure is not None:
            idx = np.flatnonzero(self.gmm_measure)
            assert np.allclose(self.X_scaled[:, idx], X[:, idx]), "Flow applied to GMM measure inputs"
        if self.empirical_measure is not None:
            back = np.stack([self._get_x_inverse_transformer(i)(self.X_scaled[:, i]) for i in self.empirical_measure],
                            axis=1)
            assert np.allclose(back, X[:, self.empirical_measure]), "Flow applied to empirical measure inputs"

        Z = None
        if X.shape[0] > 1000 or self.sparse:  # sparse GP switch (:373-374)
            if initialise_inducing_points:
                Z = self._kmeans_inducing(self.X_scaled, p0, p)
            else:
                Z = self.X_scaled[: self.num_inducing, :]
        self.m = create_model_oak(
            (self.X_scaled, self.Y_scaled),
            max_interaction_depth=self.max_interaction_depth,
            inducing_pts=Z,
            optimise=optimise,
            p0=p0,
            p=p,
            lengthscale_bounds=self.lengthscale_bounds,
            use_sparsity_prior=self.use_sparsity_prior,
            empirical_locations=self.empirical_locations,
            empirical_weights=self.empirical_weights,
            gmm_measures=self.estimated_gmm_measures,
            share_var_across_orders=self.share_var_across_orders,
        )

    def _kmeans_inducing(self, Xs, p0, p):
        """k-means inducing points (:377-391, utils.py:555-574): the continuous block on the device
        (``kmeans.KMeans``: scikit-learn's algorithm and random stream), discrete columns as the reference."""
        from .kmeans import kmeans_class

        KMeans = kmeans_class()
        if (p0 is None) and (p is None):
            return KMeans(n_clusters=self.num_inducing, random_state=0).fit(Xs).cluster_centers_
        from .utils import initialize_kmeans_with_categorical

        return initialize_kmeans_with_categorical(Xs, binary_index=self.binary_index,
                                                  categorical_index=self.categorical_index,
                                                  continuous_index=self.continuous_index, n_clusters=self.num_inducing)

    def _transform_x(self, X):
        X = apply_normalise_flow(np.asarray(X, dtype=np.float64), self.input_flows)
        if self.empirical_measure is not None:
            X[:, self.empirical_measure] = self.scaler_X_empirical.transform(X[:, self.empirical_measure])
        if not self.use_normalising_flow:
            X[:, self.continuous_index] = self.scaler_X_continuous.transform(X[:, self.continuous_index])
        return X

    def optimise(self, compile: bool = True):
        """BFGS on the training loss (:410-427); ``compile`` is accepted for signature parity."""
        from .training import optimise as _optimise

        self.alpha = None
        return _optimise(self.m, method="BFGS")

    def predict(self, X, clip=False):
        X = np.asarray(X, dtype=np.float64)
        Xs = self._transform_x(np.clip(X, self.xmin, self.xmax)) if clip else self._transform_x(X)
        # only the mean of predict_f is used (model_utils.py:441): take the fused mean when there is one
        y_pred = self.m.predict_mean(Xs) if hasattr(self.m, "predict_mean") else self.m.predict_f(Xs)[0]
        return self.scaler_y.inverse_transform(np.asarray(y_pred))[:, 0]

    def get_loglik(self, X, y, clip=False):
        """Mean predictive log density on (X, y) in the scaled output space (:445-460)."""
        X = np.asarray(X, dtype=np.float64)
        Xs = self._transform_x(np.clip(X, self.xmin, self.xmax)) if clip else self._transform_x(X)
        return float(np.mean(self.m.predict_log_density((Xs, self.scaler_y.transform(np.asarray(y, dtype=np.float64))))))

    def _get_x_inverse_transformer(self, i: int):
        """Inverse transformation of continuous feature ``i`` (:478-497)."""
        assert i in self.continuous_index
        if self.empirical_measure is not None and i in self.empirical_measure:
            j = self.empirical_measure.index(i)
            mean_i, std_i = self.scaler_X_empirical.mean_[j], np.sqrt(self.scaler_X_empirical.var_[j])
            return lambda x: x * std_i + mean_i
        if self.gmm_measure is not None and i in self.gmm_measure:  # (sic: membership test as in the reference)
            return None
        return self.input_flows[i].bijector.inverse

    def get_sobol(self, likelihood_variance=False):
        """Normalised Sobol index of each additive term (:499-524)."""
        delta, mu = 1, 0
        selected_dims, _ = get_list_representation(self.m.kernel, num_dims=self.num_dims)
        tuple_of_indices = selected_dims[1:]
        model_indices, sobols = compute_sobol_oak(self.m, delta, mu,
                                                  share_var_across_orders=self.share_var_across_orders)
        sobols = np.asarray(sobols)
        if likelihood_variance:
            normalised = sobols / (np.sum(sobols) + float(self.m.likelihood.variance.numpy()))
        else:
            normalised = sobols / np.sum(sobols)
        self.normalised_sobols = normalised
        self.tuple_of_indices = tuple_of_indices
        return normalised
