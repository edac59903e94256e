"""Device k-means (``oak_b200.kmeans.KMeans``, csrc/oak_kmeans.cu) against scikit-learn's ``KMeans`` -- the call the
reference makes for its inducing points (oak/model_utils.py:31-41, 376-383; oak/utils.py:549-552, 570-573).  The
device version follows the library's algorithm and consumes numpy's random stream identically, so the centres agree to
rounding on continuous data (no exact distance ties)."""
import numpy as np
import pytest

from helpers import max_rel_err

pytestmark = pytest.mark.gpu


def _blobs(n, d, centres, seed):
    rng = np.random.default_rng(seed)
    mu = rng.standard_normal((centres, d)) * 3.0
    return mu[rng.integers(0, centres, n)] + rng.standard_normal((n, d))


@pytest.mark.parametrize("n,d,k,seed", [(300, 2, 5, 0), (2000, 5, 20, 1), (5000, 20, 64, 2), (4097, 8, 200, 3),
                                        (1500, 33, 17, 4), (700, 70, 9, 5), (64, 3, 64, 6)])
def test_centres_match_scikit_learn(n, d, k, seed):
    from sklearn.cluster import KMeans as SkKMeans

    from oak_b200.kmeans import KMeans

    X = _blobs(n, d, max(3, k // 3), seed) if n > k else np.random.default_rng(seed).standard_normal((n, d))
    ref = SkKMeans(n_clusters=k, random_state=0).fit(X.copy())
    got = KMeans(n_clusters=k, random_state=0).fit(X)
    assert got.cluster_centers_.shape == (k, d)
    assert max_rel_err(got.cluster_centers_, ref.cluster_centers_) < 1e-9
    assert np.array_equal(got.labels_, ref.labels_)
    assert got.n_iter_ == ref.n_iter_


def test_seeding_consumes_the_random_stream_like_scikit_learn():
    from sklearn.cluster import kmeans_plusplus

    from oak_b200.kmeans import KMeans

    X = _blobs(3000, 6, 12, 7)
    Xc = X - X.mean(axis=0)
    _, idx_ref = kmeans_plusplus(Xc, 40, random_state=np.random.RandomState(0))
    km = KMeans(n_clusters=40, random_state=0, max_iter=1).fit(X)
    assert np.array_equal(km.seed_indices_, idx_ref)


def test_repeated_fits_are_bit_identical():
    from oak_b200.kmeans import KMeans

    X = _blobs(20000, 12, 30, 8)
    a = KMeans(n_clusters=128, random_state=0).fit(X).cluster_centers_
    b = KMeans(n_clusters=128, random_state=0).fit(X).cluster_centers_
    assert np.array_equal(a, b)


def test_inducing_point_helpers_use_the_device_kmeans():
    """get_kmeans_centers / initialize_kmeans_with_categorical (the reference's entry points) against the library."""
    from sklearn.cluster import KMeans as SkKMeans

    from oak_b200.model_utils import get_kmeans_centers
    from oak_b200.utils import initialize_kmeans_with_categorical

    rng = np.random.default_rng(9)
    X = np.column_stack([(rng.random(900) < 0.3).astype(float), rng.integers(0, 4, 900).astype(float),
                         _blobs(900, 3, 6, 10)])
    Z = initialize_kmeans_with_categorical(X, binary_index=[0], categorical_index=[1], continuous_index=[2, 3, 4],
                                           n_clusters=25)
    ref = SkKMeans(n_clusters=25, random_state=0).fit(X[:, [2, 3, 4]]).cluster_centers_
    assert max_rel_err(Z[:, 2:], ref) < 1e-9
    for col in (0, 1):
        ref_c = SkKMeans(n_clusters=25, random_state=0).fit(X[:, col][:, None]).cluster_centers_.astype(int)[:, 0]
        assert np.array_equal(Z[:, col], ref_c)
    Zc = get_kmeans_centers(X[:, 2:], 30)
    assert max_rel_err(Zc, SkKMeans(n_clusters=30, random_state=0).fit(X[:, 2:]).cluster_centers_) < 1e-9


def test_too_few_samples_raise_like_scikit_learn():
    from oak_b200.kmeans import KMeans

    with pytest.raises(ValueError, match="should be >= n_clusters"):
        KMeans(n_clusters=10, random_state=0).fit(np.zeros((4, 2)))


@pytest.mark.parametrize("n,K,seed", [(400, 2, 0), (3000, 3, 1), (20000, 5, 2), (1500, 1, 3)])
def test_one_dimensional_gaussian_mixture_matches_scikit_learn(n, K, seed):
    """The MOG input measure (oak/model_utils.py:753-770): GaussianMixture(K, random_state=0, spherical) on one column."""
    from sklearn.mixture import GaussianMixture

    from oak_b200.gmm import GaussianMixture1D
    from oak_b200.model_utils import estimate_one_dim_gmm

    rng = np.random.default_rng(seed)
    comp = rng.integers(0, max(K, 2), n)
    x = rng.standard_normal(n) * (0.4 + 0.3 * comp) + 2.5 * comp
    ref = GaussianMixture(n_components=K, random_state=0, covariance_type="spherical").fit(x.reshape(-1, 1))
    got = GaussianMixture1D(n_components=K, random_state=0).fit(x)
    assert got.n_iter_ == ref.n_iter_ and got.converged_ == ref.converged_
    assert max_rel_err(got.weights_, ref.weights_) < 1e-9
    assert max_rel_err(got.means_, ref.means_) < 1e-9
    assert max_rel_err(got.covariances_, ref.covariances_) < 1e-9
    mog = estimate_one_dim_gmm(K, x)
    assert max_rel_err(mog.means, ref.means_.reshape(-1)) < 1e-9
    assert max_rel_err(mog.variances, ref.covariances_) < 1e-9
    assert max_rel_err(mog.weights, ref.weights_) < 1e-9
