import numpy as np
import tensorflow as tf


class Distribution:
    def __init__(self, dtype=None, reparameterization_type=None, validate_args=False, allow_nan_stats=True, parameters=None, name=None, **kw):
        self._dtype = dtype
        self._parameters = parameters
        self._name = name

    @property
    def dtype(self):
        return self._dtype

    @property
    def parameters(self):
        return self._parameters

    def log_prob(self, value, name=None, **kw):
        return self._log_prob(value, **kw)

    def prob(self, value, **kw):
        return tf.exp(self._log_prob(value, **kw))

    def cdf(self, value, **kw):
        return self._cdf(value, **kw)

    def log_survival_function(self, value, **kw):
        if hasattr(self, "_log_survival_function"):
            return self._log_survival_function(value, **kw)
        return tf.math.log1p(-self.cdf(value, **kw))   # TFP distribution.py default

    def sample(self, sample_shape=(), seed=None, name=None, **kw):
        n = int(np.prod(np.asarray(sample_shape, dtype=np.int64))) if np.ndim(sample_shape) else int(sample_shape)
        out = self._sample_n(n, seed=seed, **kw)
        if np.ndim(sample_shape) == 0 and not isinstance(sample_shape, (list, tuple)):
            return out
        return out

    def batch_shape_tensor(self):
        return self._batch_shape_tensor()

    @property
    def batch_shape(self):
        return self._batch_shape()

    @property
    def event_shape(self):
        return self._event_shape()
