"""The scenarios of the reference's own tests/test_oak_model.py run against the B200 path: model creation with
every legal / illegal combination of feature typing and measures, which configurations support Sobol indices,
flows vs GMM measures, and the "better than the mean predictor" sanity check of the untrained model."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tiny_binary_data():
    """3 rows x 5 binary columns (the shape of the reference's binary_5D_data fixture, test_oak_model.py:91-97)."""
    rng = np.random.RandomState(42)
    X = rng.randint(0, 2, 15).reshape(3, 5).astype(float)
    return X, rng.randn(3, 1)


@pytest.mark.parametrize("sparsity_prior", [True, False])
@pytest.mark.parametrize("init_inducing", [True, False])
@pytest.mark.parametrize("sparse", [True, False])
@pytest.mark.parametrize("clip", [True, False])
def test_untrained_oak_model_beats_the_mean_predictor(sparsity_prior, init_inducing, sparse, clip):
    """test_oak_model.py:20-58: default flows, 80 training points, GPR or SGPR with 50 inducing points."""
    from oak_b200.model_utils import oak_model

    rng = np.random.RandomState(44)
    X = rng.normal(0, 1, (100, 3))
    y = X[:, 0] ** 2 + X[:, 1] + X[:, 1] * X[:, 2] + rng.normal(0, 0.01, 100)
    perm = np.random.RandomState(42).permutation(100)
    tr, te = perm[:80], perm[80:]
    oak = oak_model(num_inducing=50, max_interaction_depth=2, use_sparsity_prior=sparsity_prior, sparse=sparse)
    oak.fit(X[tr], y[tr, None], initialise_inducing_points=init_inducing, optimise=False)
    pred = oak.predict(X[te], clip=clip)
    assert pred.shape == (20,)
    assert np.mean((pred - y[te]) ** 2) < np.mean((y[te].mean() - y[te]) ** 2)


@pytest.mark.parametrize("depth", [1, 2])
@pytest.mark.parametrize("sparsity_prior", [True, False])
def test_binary_and_categorical_columns_give_a_finite_likelihood(depth, sparsity_prior):
    """test_oak_model.py:61-88: 20 points, one binary, one 4-category and one continuous column."""
    from oak_b200.model_utils import oak_model

    rng = np.random.RandomState(44)
    X = np.stack([rng.choice([0, 1], 20, p=[0.8, 0.2]), rng.choice([0, 1, 2, 3], 20, p=[0.2, 0.2, 0.3, 0.3]),
                  rng.randn(20)], axis=1).astype(float)
    Y = (np.sin(X[:, 2]) + rng.normal(0, 0.01, 20)).reshape(-1, 1)
    oak = oak_model(binary_feature=[0], categorical_feature=[1], max_interaction_depth=depth,
                    use_sparsity_prior=sparsity_prior)
    oak.fit(X, Y, optimise=False)
    assert np.isfinite(oak.m.log_marginal_likelihood())


@pytest.mark.parametrize("binary,categorical,gmm,empirical", [
    ([0], [1], [0, 0, 2, 3, 0], [4]),   # mixtures with 2 and 3 components (:100-110)
    ([0], [1], None, [2, 3]),
    ([0, 1], [2], [0, 0, 0, 2, 0], [4]),  # (:178-183)
    ([0, 1], [2], None, [3, 4]),
])
def test_legal_feature_and_measure_combinations_build(binary, categorical, gmm, empirical):
    from oak_b200.model_utils import oak_model

    X, Y = _tiny_binary_data()
    oak = oak_model(num_inducing=3, binary_feature=binary, categorical_feature=categorical, gmm_measure=gmm,
                    empirical_measure=empirical)
    oak.fit(X, Y, optimise=False)
    assert np.isfinite(oak.m.log_marginal_likelihood())


@pytest.mark.parametrize("binary,categorical,gmm,empirical", [
    ([0, 1], [1], [0] * 5, [3]),          # a column typed both binary and categorical (:203-213)
    ([0], [1], None, [0]),                # empirical measure on the binary column
    ([0], [1], None, [1]),                # ... on the categorical column
    ([0], [1], [2, 0, 0, 0, 0], [2, 4]),  # mixture measure on a discrete column
])
def test_illegal_feature_and_measure_combinations_raise_value_error(binary, categorical, gmm, empirical):
    from oak_b200.model_utils import oak_model

    X, Y = _tiny_binary_data()
    oak = oak_model(binary_feature=binary, categorical_feature=categorical, gmm_measure=gmm, empirical_measure=empirical)
    with pytest.raises(ValueError):
        oak.fit(X, Y, optimise=False)


@pytest.mark.parametrize("binary,categorical", [([0], [1]), ([0], [1, 3])])
def test_sobol_indices_of_mixed_models_are_non_negative(binary, categorical):
    """test_oak_model.py:137-159."""
    from oak_b200.model_utils import oak_model

    X, Y = _tiny_binary_data()
    cont = sorted(set(range(5)) - set(binary + categorical))
    X[:, cont] += np.random.RandomState(0).normal(0, 1, (3, len(cont)))
    oak = oak_model(binary_feature=binary, categorical_feature=categorical)
    oak.fit(X, Y, optimise=False)
    sobol = oak.get_sobol()
    assert len(sobol) == len(oak.tuple_of_indices) and np.all(sobol >= 0)


def test_sobol_with_a_mixture_measure_is_not_implemented():
    """test_oak_model.py:162-173 (utils.py:413-414)."""
    from oak_b200.model_utils import oak_model

    X, Y = _tiny_binary_data()
    X = X + np.random.RandomState(1).normal(0, 1, X.shape)
    oak = oak_model(gmm_measure=[0, 0, 3, 0, 0])
    oak.fit(X, Y, optimise=False)
    with pytest.raises(NotImplementedError):
        oak.get_sobol()


def test_mixture_measure_column_gets_no_flow():
    """test_oak_model.py:234-256: flows on the four continuous columns, a 2-component mixture (and no flow) on
    the last one, whose component means stay at its two raw values."""
    from oak_b200.input_measures import MOGMeasure
    from oak_b200.model_utils import oak_model

    X, Y = _tiny_binary_data()
    X[:, :-1] += np.random.RandomState(44).normal(0, 1, (3, 4))
    oak = oak_model(gmm_measure=[0, 0, 0, 0, 2])
    oak.fit(X, Y, optimise=False)
    assert oak.estimated_gmm_measures[:-1] == [None] * 4
    assert isinstance(oak.estimated_gmm_measures[-1], MOGMeasure)
    assert np.allclose(np.sort(np.asarray(oak.estimated_gmm_measures[-1].means).ravel()), [0.0, 1.0])
    assert oak.input_flows[-1] is None
    assert sum(f is not None for f in oak.input_flows[:-1]) == 4
