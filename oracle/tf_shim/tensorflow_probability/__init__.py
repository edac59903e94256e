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


class bijectors:  # noqa: N801
    Sigmoid = _Sigmoid


class distributions:  # noqa: N801
    class Gamma:
        def __init__(self, concentration, rate):
            self.concentration, self.rate = float(concentration), float(rate)
