"""NumPy stand-in for the two TFP names the reference hot path touches (test infrastructure)."""
import numpy as np


class _Sigmoid:
    def __init__(self, low=0.0, high=1.0):
        self.low, self.high = float(np.asarray(low)), float(np.asarray(high))

    def forward(self, u):
        return self.low + (self.high - self.low) / (1.0 + np.exp(-np.asarray(u, dtype=np.float64)))

    def inverse(self, v):
        y = (np.asarray(v, dtype=np.float64) - self.low) / (self.high - self.low)
        return np.log(y) - np.log1p(-y)


# ---- bijectors of the normalising flow (oak/normalising_flow.py), restated from TFP 0.11 (the reference's pin) ----
def _val(p):
    return np.asarray(p.numpy() if hasattr(p, "numpy") else p, dtype=np.float64)


class _Bijector:
    def __call__(self, x):
        return self.forward(x)


class _Exp(_Bijector):
    def forward(self, x):
        return _w(np.exp(np.asarray(x, dtype=np.float64)))

    def inverse(self, y):
        return _w(np.log(np.asarray(y, dtype=np.float64)))

    def forward_log_det_jacobian(self, x, event_ndims=0):
        return _w(np.asarray(x, dtype=np.float64))


class _Log(_Bijector):
    def forward(self, x):
        return _w(np.log(np.asarray(x, dtype=np.float64)))

    def inverse(self, y):
        return _w(np.exp(np.asarray(y, dtype=np.float64)))

    def forward_log_det_jacobian(self, x, event_ndims=0):
        return _w(-np.log(np.asarray(x, dtype=np.float64)))


class _Shift(_Bijector):
    def __init__(self, shift):
        self.shift = shift

    def forward(self, x):
        return _w(np.asarray(x, dtype=np.float64) + _val(self.shift))

    def inverse(self, y):
        return _w(np.asarray(y, dtype=np.float64) - _val(self.shift))

    def forward_log_det_jacobian(self, x, event_ndims=0):
        return _w(np.zeros(np.shape(x)))


class _Scale(_Bijector):
    def __init__(self, scale):
        self.scale = scale

    def forward(self, x):
        return _w(np.asarray(x, dtype=np.float64) * _val(self.scale))

    def inverse(self, y):
        return _w(np.asarray(y, dtype=np.float64) / _val(self.scale))

    def forward_log_det_jacobian(self, x, event_ndims=0):
        return _w(np.zeros(np.shape(x)) + np.log(np.abs(_val(self.scale))))


class _SinhArcsinh(_Bijector):
    """TFP 0.11: y = sinh((arcsinh(x) + skewness) * tailweight) (no tail-weight dependent multiplier yet)."""

    def __init__(self, skewness=0.0, tailweight=1.0):
        self.skewness, self.tailweight = skewness, tailweight

    def forward(self, x):
        return _w(np.sinh((np.arcsinh(np.asarray(x, dtype=np.float64)) + _val(self.skewness)) * _val(self.tailweight)))

    def inverse(self, y):
        return _w(np.sinh(np.arcsinh(np.asarray(y, dtype=np.float64)) / _val(self.tailweight) - _val(self.skewness)))

    def forward_log_det_jacobian(self, x, event_ndims=0):
        x = np.asarray(x, dtype=np.float64)
        w = (np.arcsinh(x) + _val(self.skewness)) * _val(self.tailweight)
        return _w(np.log(np.cosh(w)) - 0.5 * np.log1p(x * x) + np.log(_val(self.tailweight)))


class _Chain(_Bijector):
    """Chain([b0, ..., bn]).forward(x) = b0(b1(... bn(x)))."""

    def __init__(self, bijectors):
        self.bijectors = list(bijectors)

    def forward(self, x):
        for b in reversed(self.bijectors):
            x = b.forward(x)
        return _w(x)

    def inverse(self, y):
        for b in self.bijectors:
            y = b.inverse(y)
        return _w(y)

    def forward_log_det_jacobian(self, x, event_ndims=0):
        total = 0.0
        for b in reversed(self.bijectors):
            total = total + np.asarray(b.forward_log_det_jacobian(x, event_ndims))
            x = b.forward(x)
        return _w(total)


def _w(x):
    from tensorflow import EagerArray

    return np.asarray(x).view(EagerArray)


class bijectors:  # noqa: N801
    Sigmoid = _Sigmoid
    Exp, Log, Shift, Scale, SinhArcsinh, Chain = _Exp, _Log, _Shift, _Scale, _SinhArcsinh, _Chain


class distributions:  # noqa: N801
    class Gamma:
        def __init__(self, concentration, rate):
            self.concentration, self.rate = float(concentration), float(rate)
