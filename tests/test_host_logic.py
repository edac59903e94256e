"""Host-side logic that needs no GPU: parameter transforms, kernel construction and its error
behaviour (mirrors the reference's checks), spec packing, row partitioning."""
import numpy as np
import pytest

from oak_b200 import _cabi
from oak_b200._gpflow_shim import Parameter, Sigmoid, collect_parameters, positive
from oak_b200.input_measures import EmpiricalMeasure, GaussianMeasure, MOGMeasure, UniformMeasure
from oak_b200.oak_kernel import KernelComponenent, OAKKernel, bounded_param, get_list_representation
from oak_b200.ortho_binary_kernel import OrthogonalBinary
from oak_b200.ortho_categorical_kernel import OrthogonalCategorical
from oak_b200.ortho_rbf_kernel import RBF, OrthogonalRBFKernel
from oak_b200.parallel import balanced_symmetric_rows, partition_rows


def test_parameter_transforms_round_trip():
    p = Parameter(1.0, transform=positive())
    assert abs(float(p) - 1.0) < 1e-15
    p.assign(1e-16)
    assert abs(float(p) - 1e-16) < 1e-25
    b = bounded_param(1e-3, 1e3, 1.0)
    assert abs(float(b) - 1.0) < 1e-12
    b.assign(500.0)
    assert abs(float(b) - 500.0) < 1e-9
    assert isinstance(b.transform, Sigmoid)


def test_oak_kernel_structure_matches_reference_constructor():
    p_cat = np.array([0.2, 0.3, 0.5]).reshape(-1, 1)
    loc = np.array([[0.1], [0.5], [0.9]])
    k = OAKKernel([RBF, RBF, None, None, RBF], num_dims=5, max_interaction_depth=2, constrain_orthogonal=True,
                  p0=[None, None, 0.3, None, None], p=[None, None, None, p_cat, None],
                  lengthscale_bounds=[1e-3, 1e3],
                  empirical_locations=[None, loc, None, None, None], empirical_weights=None,
                  gmm_measures=[None, None, None, None, MOGMeasure(np.array([0.0, 1.0]), np.array([1.0, 2.0]),
                                                                   np.array([0.5, 0.5]))])
    kinds = [type(s).__name__ for s in k.kernels]
    assert kinds == ["OrthogonalRBFKernel", "OrthogonalRBFKernel", "OrthogonalBinary", "OrthogonalCategorical",
                     "OrthogonalRBFKernel"]
    assert isinstance(k.kernels[0].measure, GaussianMeasure) and k.kernels[0].measure.var == 1
    assert isinstance(k.kernels[1].measure, EmpiricalMeasure)
    assert isinstance(k.kernels[4].measure, MOGMeasure)
    assert len(k.variances) == 3
    # Gaussian-measure dims get a constant unit variance, empirical / MOG dims keep a Parameter
    assert isinstance(k.kernels[0].base_kernel.variance, np.ndarray)
    assert isinstance(k.kernels[1].base_kernel.variance, Parameter)
    assert isinstance(k.kernels[0].base_kernel.lengthscales.transform, Sigmoid)
    specs = k._dim_specs()
    assert [s.type for s in specs] == [0, 0, 1, 2, 0]
    assert [s.column for s in specs] == [0, 1, 2, 3, 4]
    assert [s.measure for s in specs] == [_cabi.MEASURE_GAUSSIAN, _cabi.MEASURE_EMPIRICAL, 0, 0, _cabi.MEASURE_MOG]
    assert specs[3].count == 3 and specs[3].rank == 2 and specs[1].count == 3
    # trainable parameters: 3 lengthscales, 2 base variances, W, kappa, binary/cat variance are constants
    names = collect_parameters(k)
    assert len(names) == 3 + 2 + 2 + 3


def test_share_var_false_has_single_variance():
    k = OAKKernel([RBF] * 3, num_dims=3, max_interaction_depth=3, constrain_orthogonal=True,
                  share_var_across_orders=False)
    assert len(k.variances) == 1
    assert all(isinstance(s.base_kernel.variance, Parameter) for s in k.kernels)


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError):  # both empirical and GMM measure on one input (oak_kernel.py:136-138)
        OAKKernel([RBF], num_dims=1, max_interaction_depth=1, constrain_orthogonal=True,
                  empirical_locations=[np.array([[0.0], [1.0]])], empirical_weights=None,
                  gmm_measures=[MOGMeasure(np.array([0.0]), np.array([1.0]), np.array([1.0]))])
    with pytest.raises(AssertionError):  # duplicate active dims (:80-82)
        OAKKernel([RBF, RBF], num_dims=2, max_interaction_depth=1, active_dims=[[0], [0]])
    with pytest.raises(AssertionError):  # empirical locations without the orthogonal constraint (:192-197)
        OAKKernel([RBF], num_dims=1, max_interaction_depth=1, constrain_orthogonal=False,
                  empirical_locations=[np.array([[0.0]])])
    with pytest.raises(NotImplementedError):  # non-RBF base kernel (ortho_rbf_kernel.py:34-35)
        OrthogonalRBFKernel(OrthogonalBinary(), GaussianMeasure(0, 1))
    with pytest.raises(NotImplementedError):  # unknown measure (:36-45)
        OrthogonalRBFKernel(RBF(), object())
    with pytest.raises(AssertionError):  # weights must sum to one (input_measures.py:53-55)
        EmpiricalMeasure(np.zeros((3, 1)), np.ones((3, 1)))
    with pytest.raises(AssertionError):
        MOGMeasure(np.array([0.0, 1.0]), np.array([1.0, 1.0]), np.array([0.7, 0.7]))
    with pytest.raises(ValueError):  # depth beyond what the register-tiled kernels support
        OAKKernel([RBF] * 20, num_dims=20, max_interaction_depth=17)


def test_shape_validation_of_sub_kernels():
    k = OrthogonalRBFKernel(RBF(), UniformMeasure(0, 1))
    with pytest.raises(ValueError):
        k.K(np.zeros((4, 2)))
    with pytest.raises(ValueError):
        OrthogonalBinary().K(np.zeros((4, 2)))
    with pytest.raises(ValueError):
        OrthogonalCategorical(p=np.array([[0.5], [0.5]])).K_diag(np.zeros((4,)))
    with pytest.raises(ValueError):  # full_cov=False with X2 (gpflow Kernel.__call__)
        k(np.zeros((3, 1)), np.zeros((3, 1)), full_cov=False)


def test_list_representation_order():
    k = OAKKernel([RBF] * 4, num_dims=4, max_interaction_depth=3, constrain_orthogonal=True)
    sel, kl = get_list_representation(k, num_dims=4)
    assert sel[:6] == [[], [0], [1], [2], [3], [0, 1]]
    assert len(sel) == 1 + 4 + 6 + 4 and len(kl) == len(sel)
    assert all(isinstance(c, KernelComponenent) for c in kl)
    assert [len(c.kernels) for c in kl] == [len(s) for s in sel]
    k0 = OAKKernel([RBF], num_dims=1, max_interaction_depth=0, constrain_orthogonal=True)
    assert get_list_representation(k0, num_dims=1)[0] == [[]]


def test_spec_variance_count_is_validated():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs the no-device path to stay cheap")
    with pytest.raises((_cabi.OakNativeError, ValueError)):
        _cabi.Spec([_cabi.DimSpec(_cabi.DIM_RBF, 0)], 3, [1.0], share_var=True)


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (63, 2), (64, 2), (1000, 8), (65536, 8), (1_000_000, 8), (129, 3)])
def test_partition_rows_covers_and_aligns(n, world):
    parts = partition_rows(n, world)
    assert len(parts) == world
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
        assert e0 == b1
    for b, e in parts:
        assert b <= e and (b % 64 == 0 or b == n)
    sizes = [e - b for b, e in parts]
    assert max(sizes) - min(sizes) <= 64 or n < 64 * world


@pytest.mark.parametrize("n,world", [(1100, 2), (65536, 8), (4096, 4)])
def test_balanced_symmetric_rows_balance_the_triangle(n, world):
    strips = balanced_symmetric_rows(n, world)
    seen = sorted(s for r in strips for s in r if s[1] > s[0])
    assert seen[0][0] == 0 and seen[-1][1] == n
    for (b0, e0), (b1, e1) in zip(seen, seen[1:]):
        assert e0 == b1
    work = [sum((e - b) * (b + e + 1) / 2 for b, e in r) for r in strips]
    assert sum(work) == n * (n + 1) / 2
    if n >= 4096:
        assert max(work) / min(work) < 1.1


def test_save_and_load_model_round_trip(tmp_path):
    """save_model / load_model (model_utils.py:44-87): trainable parameters in traversal order."""
    import numpy as np

    from oak_b200.model_utils import create_model_oak, load_model, save_model

    rng = np.random.default_rng(0)
    X, y = rng.standard_normal((30, 3)), rng.standard_normal((30, 1))
    a = create_model_oak((X, y), max_interaction_depth=2, inducing_pts=X[:5].copy(), lengthscale_bounds=[1e-3, 1e3])
    for i, p in enumerate(a.trainable_parameters):
        p.assign(np.asarray(p.numpy()) * 0 + 0.3 + 0.1 * i)
    f = tmp_path / "ckpt" / "model.npz"
    save_model(a, f)
    b = create_model_oak((X, y), max_interaction_depth=2, inducing_pts=X[:5].copy(), lengthscale_bounds=[1e-3, 1e3])
    load_model(b, f)
    assert len(a.trainable_parameters) == len(b.trainable_parameters) == 3 + 3 + 1
    for p, q in zip(a.trainable_parameters, b.trainable_parameters):
        np.testing.assert_allclose(p.numpy(), q.numpy(), rtol=1e-12)


def test_normalising_flow_oracle_gradient_log_det_and_host_inverse():
    """oracle/flow_oracle.py (oak/normalising_flow.py:30-85 restated): analytic gradient of the KL objective vs
    central differences, log-det-Jacobian vs a numerical derivative; the product's host-side ``inverse`` /
    ``forward_log_det_jacobian`` (NumPy utilities of normalising_flow.Normalizer) agree with it.  The fit and
    the forward transform run on the device: tests/test_gpu_flow.py."""
    import numpy as np

    from oak_b200.normalising_flow import Normalizer
    from oracle import flow_oracle as fo

    rng = np.random.default_rng(0)
    x = np.exp(0.6 * rng.standard_normal(2000)) + 3.0  # log-normal, shifted
    for log in (True, False):
        n = Normalizer(x, log=log)
        assert n.offset == (x.min() - 1.0 if log else 0.0)
        theta = np.array([0.1, -0.2, 0.15, -0.1]) + np.array([
            float(n.scale.unconstrained_variable), float(n.shift.unconstrained_variable), 0.0, 0.0])
        J, g = fo.kl_objective_and_grad(x, n.offset, log, theta)
        for i in range(4):
            e = np.zeros(4)
            e[i] = 1e-6
            fd = (fo.kl_objective_and_grad(x, n.offset, log, theta + e)[0]
                  - fo.kl_objective_and_grad(x, n.offset, log, theta - e)[0]) / 2e-6
            assert abs(fd - g[i]) < 1e-6 * max(1.0, abs(g[i]))
        n.scale.unconstrained_variable, n.shift.unconstrained_variable = np.asarray(theta[0]), np.asarray(theta[1])
        n.skewness.unconstrained_variable, n.tailweight.unconstrained_variable = np.asarray(theta[2]), np.asarray(theta[3])
        par = (n.offset, log, theta[1], np.exp(theta[0]), theta[2], np.exp(theta[3]))
        y = fo.forward(x, *par)
        assert abs(J - (0.5 * np.mean(y * y) - np.mean(fo.forward_log_det_jacobian(x, *par)))) < 1e-12
        np.testing.assert_allclose(n.bijector.inverse(y), x, rtol=1e-10)
        h = 1e-6
        num = np.log((fo.forward(x[:50] + h, *par) - fo.forward(x[:50] - h, *par)) / (2 * h))
        np.testing.assert_allclose(fo.forward_log_det_jacobian(x[:50], *par), num, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(n.bijector.forward_log_det_jacobian(x[:50]), num, rtol=1e-6, atol=1e-7)


def test_default_symmetric_polynomial_scheme_is_the_direct_recurrence_everywhere():
    """The package default, the bench default and the header's enum agree: kernels pass OAK_ESP_DIRECT unless a
    kernel instance asks for the reference's power sums + Newton-Girard (`esp_algorithm = 0`)."""
    import os
    import re

    from oak_b200 import _native_kernel

    assert (_cabi.ESP_NEWTON_GIRARD, _cabi.ESP_DIRECT) == (0, 1)
    assert _native_kernel._DEFAULT_ALGORITHM == _cabi.ESP_DIRECT
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "oak_b200.h")).read()
    assert re.search(r"OAK_ESP_NEWTON_GIRARD\s*=\s*0", header) and re.search(r"OAK_ESP_DIRECT\s*=\s*1", header)
    bench = open(os.path.join(root, "bench.py")).read()
    assert re.search(r'"--algo", type=int, default=1', bench)
    k = OAKKernel([RBF] * 3, num_dims=3, max_interaction_depth=2, active_dims=[[0], [1], [2]])
    assert k.esp_algorithm is None  # instances follow the process-wide default until told otherwise


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours): one JSON line with the
    contract's keys, runnable without a GPU."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the arm must use all host cores regardless
    env1 = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--n-cpu", "512", "--elbo-m", "64"], capture_output=True, text=True,
                         timeout=600, env=env1)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["elbo_evals_per_s_scaled_to_n1e6"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # ranks other than 0 stay silent under torchrun
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_svgp_host_side_contract():
    """gpflow.models.SVGP stand-in (examples/uci/uci_classification_train.py:108-118): only the reference's
    configuration is built, the variational parameters start at the prior, the inverse links evaluate on the host."""
    from oak_b200._gpflow_shim import Bernoulli, Gaussian, inv_logit, inv_probit, set_trainable
    from oak_b200.models import SVGP
    from oak_b200.oak_kernel import OAKKernel
    from oak_b200.ortho_rbf_kernel import RBF
    from oak_b200.training import trainable_parameters

    k = OAKKernel([RBF] * 2, num_dims=2, max_interaction_depth=2, constrain_orthogonal=True)
    Z = np.zeros((5, 2))
    with pytest.raises(NotImplementedError):
        SVGP(kernel=k, likelihood=Bernoulli(invlink=inv_logit), inducing_variable=Z, whiten=False, q_diag=True)
    with pytest.raises(NotImplementedError):
        SVGP(kernel=k, likelihood=Bernoulli(invlink=inv_logit), inducing_variable=Z, whiten=True, q_diag=False)
    with pytest.raises(NotImplementedError):
        SVGP(kernel=k, likelihood=Gaussian(), inducing_variable=Z, whiten=True, q_diag=True)
    with pytest.raises(NotImplementedError):
        Bernoulli(invlink=lambda x: x)
    m = SVGP(kernel=k, likelihood=Bernoulli(invlink=inv_logit), inducing_variable=Z, whiten=True, q_diag=True)
    assert m.q_mu.numpy().shape == (5, 1) and np.all(m.q_mu.numpy() == 0.0)
    assert np.allclose(m.q_sqrt.numpy(), 1.0) and m.data is None
    n_before = len(trainable_parameters(m))
    set_trainable(m.inducing_variable, False)
    assert len(trainable_parameters(m)) == n_before - 1
    x = np.array([-50.0, -1.0, 0.0, 2.0, 50.0])
    assert np.allclose(inv_logit(x), (1.0 / (1.0 + np.exp(-x))) * 0.998 + 1e-3)
    assert abs(inv_logit(np.array([-800.0]))[0] - 1e-3) < 1e-15  # no overflow
    from scipy.stats import norm

    assert np.allclose(inv_probit(x), norm.cdf(x) * 0.998 + 1e-3)
    closure = m.training_loss_closure((np.zeros((3, 2)), np.zeros((3, 1))))
    assert closure.model is m and closure.data[0].shape == (3, 2)


@pytest.mark.parametrize("binary_index", [[0, 1, 2], [0, 2], [1]])
@pytest.mark.parametrize("n_cluster", [1, 20])
def test_kmeans_inducing_points_with_binary_columns(binary_index, n_cluster):
    """initialize_kmeans_with_binary / _categorical (oak/utils.py:533-574; reference tests/test_utils.py:17-40):
    host sklearn calls with the reference's seeds; binary columns come back as integers."""
    from oak_b200.utils import initialize_kmeans_with_binary, initialize_kmeans_with_categorical

    rng = np.random.RandomState(44)
    cont = sorted(set(range(3)) - set(binary_index))
    X = np.zeros((100, 3))
    for i in binary_index:
        X[:, i] = rng.binomial(1, 0.33, 100)
    for j in cont:
        X[:, j] = rng.normal(0, 4, 100)
    Z = initialize_kmeans_with_binary(X, binary_index, cont if cont else None, n_cluster)
    assert isinstance(Z, np.ndarray) and Z.shape == (n_cluster, 3)
    assert set(np.unique(Z[:, binary_index])) <= {0.0, 1.0}
    if cont and len(binary_index) == 1:
        Zc = initialize_kmeans_with_categorical(X, [], binary_index, cont, n_cluster)
        np.testing.assert_array_equal(Zc, Z)


def test_parameter_order_is_gpflows_and_checkpoints_follow_it(tmp_path):
    """The reference's checkpoints are positional (``hyperparams[i]`` <-> ``model.trainable_parameters[i]``,
    oak/model_utils.py:44-87), so the traversal order must be gpflow's: tf.Module walks attributes in sorted
    order, yields a module's own Parameters first and recurses into its sub-modules afterwards.  For an SGPR on
    an OAK kernel that is: inducing points, order variances, per sub-kernel (lengthscales, variance), noise.
    An SVGP's own q_mu / q_sqrt precede every sub-module, and all its parameters are saved (:53-56)."""
    from oak_b200._gpflow_shim import Bernoulli, inv_logit, set_trainable
    from oak_b200.model_utils import create_model_oak, load_model, save_model
    from oak_b200.models import SVGP

    rng = np.random.default_rng(1)
    X, y = rng.standard_normal((20, 2)), rng.standard_normal((20, 1))
    m = create_model_oak((X, y), max_interaction_depth=2, inducing_pts=X[:4].copy(), zfixed=False)
    k = m.kernel
    want = [m.inducing_variable.Z, *k.variances, k.kernels[0].base_kernel.lengthscales,
            k.kernels[1].base_kernel.lengthscales, m.likelihood.variance]
    got = list(m.trainable_parameters)
    assert len(got) == len(want) and all(a is b for a, b in zip(got, want))
    s = SVGP(kernel=k, likelihood=Bernoulli(invlink=inv_logit), inducing_variable=X[:4].copy(), whiten=True, q_diag=True)
    set_trainable(s.inducing_variable, False)
    allp = list(s.parameters)
    assert allp[0] is s.q_mu and allp[1] is s.q_sqrt and allp[2] is s.inducing_variable.Z and allp[3] is k.variances[0]
    s.q_mu.assign(rng.standard_normal((4, 1)))
    f = tmp_path / "svgp.npz"
    save_model(s, f)
    assert len(np.load(str(f), allow_pickle=True)["hyperparams"]) == len(allp)  # all parameters, fixed Z included
    s2 = SVGP(kernel=k, likelihood=Bernoulli(invlink=inv_logit), inducing_variable=np.zeros((4, 2)), whiten=True,
              q_diag=True)
    load_model(s2, f, load_all_parameters=True)
    np.testing.assert_allclose(s2.q_mu.numpy(), s.q_mu.numpy())
    np.testing.assert_allclose(s2.inducing_variable.Z.numpy(), X[:4])


def test_training_chain_host_pieces_against_finite_differences():
    """The host side of training.py, no GPU needed: d constrained / d unconstrained of the gpflow transforms, the
    Gamma-prior gradient (model_utils.py:163-167), and the chain from the cotangent of a categorical / binary B
    table (what the backward tiles return) to W, kappa and variance (ortho_categorical_kernel.py:34-53,
    ortho_binary_kernel.py:29-38)."""
    from types import SimpleNamespace

    from oak_b200._gpflow_shim import Gamma, Identity, Softplus
    from oak_b200.ortho_binary_kernel import OrthogonalBinary
    from oak_b200.ortho_categorical_kernel import OrthogonalCategorical
    from oak_b200.training import _prior_grad, _transform_grad, discrete_parameter_gradients
    from oracle import oak_oracle as oo

    h = 1e-6
    for tr, val in ((Identity(), 0.7), (Softplus(0.0), 0.3), (Softplus(1e-6), 2.5), (Sigmoid(0.1, 9.0), 4.2)):
        p = Parameter(val, transform=tr)
        u = float(p.unconstrained_variable)
        fd = (float(tr.forward(u + h)) - float(tr.forward(u - h))) / (2 * h)
        assert abs(float(_transform_grad(p)) - fd) < 1e-8 * max(1.0, abs(fd))
    prior = Gamma(1.0, 0.2)
    for conc, rate, x in ((1.0, 0.2, 0.9), (2.5, 1.3, 0.4)):
        p = Parameter(x, transform=positive(), prior=Gamma(conc, rate))
        fd = (float(p.prior.log_prob(x + h)) - float(p.prior.log_prob(x - h))) / (2 * h)
        assert abs(float(_prior_grad(p)) - fd) < 1e-7
    assert abs(float(_prior_grad(Parameter(1.0, transform=positive(), prior=prior))) + 0.2) < 1e-15

    rng = np.random.default_rng(5)
    kc = OrthogonalCategorical(p=np.array([0.2, 0.5, 0.3]).reshape(-1, 1), rank=2, active_dims=[0])
    kc.W.assign(rng.uniform(0.2, 1.0, (3, 2)))
    kc.kappa.assign(rng.uniform(0.5, 1.5, kc.kappa.numpy().shape))
    kc.variance.assign(1.7)
    kb = OrthogonalBinary(p0=0.35, active_dims=[1])
    kb.variance.assign(0.8)
    # cotangents of the two table blocks [C x C | C diagonal], laid out one after the other
    g_tab = rng.standard_normal(3 * 3 + 3 + 2 * 2 + 2)
    layout = [(0, 3), (12, 2)]
    model = SimpleNamespace(kernel=SimpleNamespace(kernels=[kc, kb]), _table_cotangent=(g_tab, layout))
    grads = discrete_parameter_gradients(model)

    def objective():
        dc = oo.CategoricalDim(p=kc._p_vector(), W=kc.W.numpy(), kappa=kc.kappa.numpy().reshape(-1),
                               variance=float(kc.variance.numpy()))
        db = oo.BinaryDim(kb.p0, float(kb.variance.numpy()))
        return ((g_tab[0:9].reshape(3, 3) * dc.table()).sum() + (g_tab[9:12] * dc.table_diag()).sum()
                + (g_tab[12:16].reshape(2, 2) * db.table()).sum() + (g_tab[16:18] * db.table_diag()).sum())

    for prm in (kc.W, kc.kappa, kc.variance, kb.variance):
        base = prm.numpy().copy()
        flat = base.reshape(-1)
        fd = np.zeros(flat.size)
        for i in range(flat.size):
            for sgn in (1, -1):
                v = flat.copy()
                v[i] += sgn * h
                prm.assign(v.reshape(base.shape))
                fd[i] += sgn * objective() / (2 * h)
        prm.assign(base)
        assert np.allclose(np.asarray(grads[id(prm)]).reshape(-1), fd, rtol=1e-6, atol=1e-8), prm


def test_fit_consistency_checks_follow_the_reference():
    """oak_model.fit keeps the reference's assertions (model_utils.py:346-370): with use_normalising_flow=False an
    empirical-measure column is standardised twice by _transform_x (:468-475) and the reference stops with
    "Flow applied to empirical measure inputs"; so does this implementation (no GPU is reached)."""
    from oak_b200.model_utils import oak_model

    rng = np.random.default_rng(0)
    X = np.stack([rng.standard_normal(50), np.round(3 * rng.standard_normal(50)) / 3], axis=1)
    Y = rng.standard_normal((50, 1))
    oak = oak_model(use_normalising_flow=False, empirical_measure=[1])
    with pytest.raises(AssertionError, match="empirical measure inputs"):
        oak.fit(X, Y, optimise=False)
