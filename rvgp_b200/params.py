"""Minimal stand-ins for the GPflow objects the reference touches: ``gpflow.Parameter`` with a
``gpflow.utilities.positive()`` transform (softplus, shifted by a lower bound that defaults to the module-global
default positive minimum AT CALL TIME), and ``gpflow.config.set_default_positive_minimum`` (RVGP/main.py:52,
RVGP/kernels.py:36-38)."""
import numpy as np

_default_positive_minimum = [0.0]          # gpflow's default


def set_default_positive_minimum(v):
    _default_positive_minimum[0] = float(v)


def default_positive_minimum():
    return _default_positive_minimum[0]


class positive:
    """softplus bijector, plus Shift(lower) when lower != 0 (gpflow/utilities/bijectors.py)."""

    def __init__(self, lower=None):
        self.lower = default_positive_minimum() if lower is None else float(lower)

    def forward(self, u):
        return self.lower + np.logaddexp(0.0, u)

    def inverse(self, y):
        y = float(y) - self.lower
        if not y > 0:
            raise ValueError("parameter value must be above its lower bound %g" % self.lower)
        return y + np.log(-np.expm1(-y))

    def dforward(self, u):
        if u < -700.0:                                   # exp(-u) would overflow; sigmoid underflows to 0 anyway
            return 0.0
        return 1.0 / (1.0 + np.exp(-u))


class Parameter:
    def __init__(self, value, transform=None, name=None, trainable=True):
        self.transform = transform or positive(0.0)
        self.name = name
        self.trainable = trainable
        self.unconstrained = float(self.transform.inverse(value))

    @property
    def value(self):
        return float(self.transform.forward(self.unconstrained))

    def numpy(self):
        return np.float64(self.value)

    def assign(self, v):
        self.unconstrained = float(self.transform.inverse(v))

    def __float__(self):
        return self.value

    def __repr__(self):
        return "<Parameter %s=%.6g lower=%g trainable=%s>" % (self.name, self.value, self.transform.lower, self.trainable)


def set_trainable(p, flag):
    p.trainable = bool(flag)
