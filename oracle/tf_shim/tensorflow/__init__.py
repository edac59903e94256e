"""NumPy stand-in for the TensorFlow ops used by the reference hot path (test infrastructure)."""
import numpy as np
from scipy import linalg as _sla
from scipy.special import erf as _erf

float64, float32, int32, int64 = np.float64, np.float32, np.int32, np.int64
Tensor = np.ndarray


class _Dtypes:
    float64 = np.float64
    float32 = np.float32
    int32 = np.int32


dtypes = _Dtypes()


class EagerArray(np.ndarray):
    """ndarray that also answers ``.numpy()`` like a TF EagerTensor."""

    def numpy(self):
        return np.asarray(self)


def _a(x):
    return np.asarray(x)


def _w(x):
    return np.asarray(x).view(EagerArray)


def convert_to_tensor(x, dtype=None):
    return np.asarray(x, dtype=dtype)


def cast(x, dtype):
    x = _a(x)
    if np.issubdtype(dtype, np.integer):
        return np.trunc(x).astype(dtype)  # tf.cast float->int truncates toward zero
    return x.astype(dtype)


def ones(shape, dtype=np.float32):
    return np.ones(shape, dtype=dtype)


def zeros(shape, dtype=np.float32):
    return np.zeros(shape, dtype=dtype)


def ones_like(x):
    return np.ones_like(_a(x))


def eye(n, dtype=np.float32):
    return np.eye(n, dtype=dtype)


def pow(x, p):  # noqa: A001
    return np.power(_a(x), p)


def add(a, b):
    return _a(a) + _a(b)


def exp(x):
    return np.exp(_a(x))


def sqrt(x):
    return np.sqrt(_a(x))


def square(x):
    return np.square(_a(x))


def squeeze(x, axis=None):
    return np.squeeze(_a(x), axis=axis)


def transpose(x):
    return _w(np.transpose(_a(x)))


def reshape(x, shape):
    return np.reshape(_a(x), shape)


def shape(x):
    return np.array(_a(x).shape)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _a(a), _a(b)
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return _w(a @ b)


def tensordot(a, b, axes):
    return _w(np.tensordot(_a(a), _a(b), axes))


def gather(params, indices, axis=0):
    return np.take(_a(params), _a(indices), axis=axis)


def reduce_sum(x, axis=None, keepdims=False):
    return np.sum(_a(x), axis=axis, keepdims=keepdims)


def reduce_mean(x, axis=None):
    return _w(np.mean(_a(x), axis=axis))


def reduce_prod(x, axis=None):
    return np.prod(np.asarray(x), axis=axis)


def range(*args, dtype=None):  # noqa: A001
    return np.arange(*args, dtype=dtype)


def function(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def negative(x):
    return -_a(x)


def fill(dims, value):
    return np.full(dims, value)


def custom_gradient(f):
    """Forward value only (no autodiff in the shim)."""
    return lambda *a, **k: f(*a, **k)[0]


def numpy_function(func, inp, Tout, name=None):
    return func(*[np.asarray(i) for i in inp])


class math:  # noqa: N801
    erf = staticmethod(lambda x: _erf(_a(x)))
    log = staticmethod(lambda x: np.log(_a(x)))
    sigmoid = staticmethod(lambda x: _w(1.0 / (1.0 + np.exp(-_a(x)))))


class linalg:  # noqa: N801
    @staticmethod
    def matmul(a, b, transpose_a=False, transpose_b=False):
        return matmul(a, b, transpose_a, transpose_b)

    @staticmethod
    def diag(x):
        return np.diag(_a(x))

    @staticmethod
    def diag_part(x):
        return np.diagonal(_a(x))

    @staticmethod
    def cholesky(x):
        return np.linalg.cholesky(_a(x))

    @staticmethod
    def triangular_solve(L, b, lower=True):
        return _w(_sla.solve_triangular(_a(L), _a(b), lower=lower))

    @staticmethod
    def solve(A, b):
        return _w(np.linalg.solve(_a(A), _a(b)))

    @staticmethod
    def inv(A):
        return np.linalg.inv(_a(A))

    @staticmethod
    def cholesky_solve(L, b):
        return _w(_sla.cho_solve((_a(L), True), _a(b)))

    @staticmethod
    def adjoint(x):
        return np.swapaxes(_a(x), -1, -2)


class debugging:  # noqa: N801
    @staticmethod
    def assert_shapes(specs):
        """Checks ranks / literal dims / named-dim consistency like tf.debugging.assert_shapes."""
        names = {}
        for x, spec in specs:
            shp = tuple(np.shape(x))
            spec = list(spec)
            if spec and spec[0] is Ellipsis:
                spec = spec[1:]
                if len(shp) < len(spec):
                    raise ValueError(f"rank of {shp} smaller than {spec}")
                shp = shp[len(shp) - len(spec):]
            elif len(shp) != len(spec):
                raise ValueError(f"shape {shp} does not match {spec}")
            for d, s in zip(shp, spec):
                if isinstance(s, (int, np.integer)):
                    if d != s:
                        raise ValueError(f"shape {shp} does not match {spec}")
                elif isinstance(s, str):
                    if names.setdefault(s, d) != d:
                        raise ValueError(f"dimension {s} inconsistent: {names[s]} vs {d}")


class random:  # noqa: N801
    @staticmethod
    def uniform(shape, minval=0.0, maxval=1.0, dtype=np.float32):
        return np.random.uniform(minval, maxval, size=shape).astype(dtype)

    @staticmethod
    def set_seed(s):
        np.random.seed(s)
