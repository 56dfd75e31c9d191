"""tfp.distributions (stand-in, see ../../__init__.py)."""
import inspect

import numpy as np
import tensorflow as tf
from scipy import special

from . import distribution  # noqa: F401
from .distribution import Distribution

_a = tf._a
_res = tf._res


def _f64(x):
    return _a(x).astype(np.float64)


class Categorical(Distribution):
    def __init__(self, logits=None, probs=None, dtype=tf.int32, validate_args=False, allow_nan_stats=True, name="Categorical"):
        super().__init__(dtype=dtype, name=name)
        if logits is None:
            with np.errstate(divide="ignore"):
                logits = np.log(_a(probs))
        self._logits = _a(logits)

    def _log_prob(self, k):
        ls = _a(tf.nn.log_softmax(self._logits, -1))
        k = _a(k).astype(np.int64)
        ls_b, k_b = np.broadcast_arrays(ls, k[..., None]) if ls.ndim - 1 != k.ndim or ls.shape[:-1] != k.shape else (ls, k[..., None])
        if ls_b is ls:
            return _res(np.take_along_axis(ls, k[..., None], -1)[..., 0])
        shp = np.broadcast_shapes(ls.shape[:-1], k.shape)
        ls2 = np.broadcast_to(ls, shp + ls.shape[-1:])
        k2 = np.broadcast_to(k, shp)
        return _res(np.take_along_axis(ls2, k2[..., None], -1)[..., 0])

    def _sample_n(self, n, seed=None):
        lg = self._logits.reshape(-1, self._logits.shape[-1])
        out = _a(tf.random.categorical(lg, n, dtype=np.int32))   # [batch, n]
        out = out.T.reshape((n,) + self._logits.shape[:-1])
        return _res(out.astype(np.int32))

    def sample(self, sample_shape=(), seed=None, name=None):
        if isinstance(sample_shape, (list, tuple)) and len(sample_shape) == 0:
            return _res(_a(self._sample_n(1, seed))[0])
        n = int(_a(sample_shape).reshape(-1)[0]) if np.ndim(_a(sample_shape)) else int(_a(sample_shape))
        return self._sample_n(n, seed)


class Deterministic(Distribution):
    def __init__(self, loc, **kw):
        super().__init__(dtype=_a(loc).dtype)
        self.loc = _a(loc)

    def _log_prob(self, x):
        with np.errstate(divide="ignore"):
            return _res(np.log((_a(x) == self.loc).astype(np.float32)))

    def _sample_n(self, n, seed=None):
        return _res(np.broadcast_to(self.loc, (n,) + self.loc.shape).copy())


class Independent(Distribution):
    def __init__(self, distribution, reinterpreted_batch_ndims=None, **kw):   # noqa: A002
        super().__init__(dtype=distribution.dtype)
        self.distribution = distribution
        self.n = 0 if reinterpreted_batch_ndims is None else int(_a(reinterpreted_batch_ndims))

    def _log_prob(self, x):
        lp = _a(self.distribution.log_prob(x))
        for _ in range(self.n):
            lp = lp.sum(-1)
        return _res(lp)

    def _sample_n(self, n, seed=None):
        return self.distribution._sample_n(n, seed)


class NegativeBinomial(Distribution):
    def __init__(self, total_count, logits=None, probs=None, validate_args=False, allow_nan_stats=True, name="NegativeBinomial"):
        tc = _a(total_count)
        super().__init__(dtype=tc.dtype if tc.dtype.kind == "f" else np.float32)
        self.total_count = tc.astype(self.dtype)
        if logits is None:
            p = _f64(probs)
            with np.errstate(divide="ignore", invalid="ignore"):
                logits = (np.log(p) - np.log1p(-p)).astype(self.dtype)
        self.logits = _a(logits).astype(self.dtype)

    def _log_prob(self, x):
        x64, tc, lg = _f64(x), _f64(self.total_count), _f64(self.logits)
        with np.errstate(divide="ignore", invalid="ignore"):
            unnorm = tc * (-np.logaddexp(0.0, lg)) + x64 * (-np.logaddexp(0.0, -lg))
            lognorm = special.betaln(1.0 + x64, tc) + np.log(tc + x64)
        return _res((unnorm - lognorm).astype(self.dtype))

    def _cdf(self, x):
        x64, tc, lg = _f64(x), _f64(self.total_count), _f64(self.logits)
        with np.errstate(invalid="ignore"):
            return _res(special.betainc(tc, 1.0 + x64, 1.0 / (1.0 + np.exp(lg))).astype(self.dtype))


class BetaBinomial(Distribution):
    def __init__(self, total_count, concentration1, concentration0, validate_args=False, allow_nan_stats=True, name="BetaBinomial"):
        super().__init__(dtype=_a(concentration1).dtype)
        self.total_count, self.c1, self.c0 = _a(total_count), _a(concentration1), _a(concentration0)

    def _log_prob(self, x):
        n, c1, c0, x64 = _f64(self.total_count), _f64(self.c1), _f64(self.c0), _f64(x)
        with np.errstate(invalid="ignore"):
            comb = special.gammaln(n + 1.0) - special.gammaln(x64 + 1.0) - special.gammaln(n - x64 + 1.0)
            return _res((special.betaln(c1 + x64, c0 + n - x64) - special.betaln(c1, c0) + comb).astype(self.dtype))


class JointDistributionNamed(Distribution):
    """dict of distributions / callables whose parameter names are other entries; log_prob sums the parts."""

    def __init__(self, model, validate_args=False, name=None):
        super().__init__(dtype=None)
        self.model = model

    def _resolve(self, value):
        done, parts = {}, {}
        pending = dict(self.model)
        while pending:
            progressed = False
            for k in list(pending):
                m = pending[k]
                if callable(m) and not isinstance(m, Distribution):
                    names = list(inspect.signature(m).parameters)
                    if not all(nm in done for nm in names):
                        continue
                    m = m(*[value[nm] for nm in names])
                parts[k] = m
                done[k] = True
                del pending[k]
                progressed = True
            if not progressed:
                raise ValueError("JointDistributionNamed: unresolvable dependencies")
        return parts

    def _log_prob(self, value):
        parts = self._resolve(value)
        total = None
        for k, d in parts.items():
            lp = _a(d.log_prob(value[k]))
            total = lp if total is None else total + lp
        return _res(total)

    def _sample_n(self, n, seed=None):
        value, pending = {}, dict(self.model)
        while pending:
            progressed = False
            for k in list(pending):
                m = pending[k]
                if callable(m) and not isinstance(m, Distribution):
                    names = list(inspect.signature(m).parameters)
                    if not all(nm in value for nm in names):
                        continue
                    m = m(*[value[nm] for nm in names])
                value[k] = m.sample(n, seed=seed)
                del pending[k]
                progressed = True
            if not progressed:
                raise ValueError("JointDistributionNamed: unresolvable dependencies")
        return value

    def sample(self, sample_shape=(), seed=None, name=None):
        n = int(_a(sample_shape).reshape(-1)[0]) if np.ndim(_a(sample_shape)) else int(_a(sample_shape))
        return self._sample_n(n, seed)

    def log_prob_parts(self, value):
        return {k: d.log_prob(value[k]) for k, d in self._resolve(value).items()}
