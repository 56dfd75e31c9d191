"""oracle/shim_tf/tensorflow -- TEST INFRASTRUCTURE, not product code.

A NumPy-backed, eager stand-in for the slice of TensorFlow 2.3 the reference's two-group modules touch
(/root/reference/src/two_group/hygeia/{case_control_proposal_mappings,case_control_distributions,case_control_regime_model,
resampling_functions,smoothing_functions}.py), written from the TensorFlow API documentation.  TensorFlow / TFP are pinned by
the reference (src/two_group/requirements.txt:74,78) and cannot be installed here (no network); with this stand-in the
reference modules are imported UNMODIFIED, where they lie, and run on NumPy arrays, which pins our restatement
(oracle/tg_oracle.py) and the CUDA path to the reference's own code for: the proposal maps, the transition log-densities, the
hazard construction (including its fp32 fall-back value), optimal finite-state + systematic resampling, and the backward
kernel.  What it cannot reproduce: TensorFlow's random streams and the last bits of its fp32 special functions (lgamma,
betainc) -- values are computed in fp64 with SciPy and rounded to fp32 where TensorFlow would hold fp32.

Semantics kept: dtypes (fp32 stays fp32), TensorShape comparisons (`x.shape == []`), functional TensorArray, tf.while_loop /
tf.cond as Python control flow, tf.nest.map_structure over dicts / tuples / namedtuples.
"""
import collections as _collections
import contextlib as _contextlib

import numpy as _np

float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
bool = _np.bool_   # noqa: A001 (mirrors tf.bool)
newaxis = None

_pybool = __builtins__["bool"] if isinstance(__builtins__, dict) else __builtins__.bool


class TensorShape(tuple):
    def __new__(cls, dims=()):
        if dims is None:
            dims = ()
        if isinstance(dims, (int, _np.integer)):
            dims = (int(dims),)
        return super().__new__(cls, tuple(None if d is None else int(d) for d in dims))

    def _coerce(self, other):
        if isinstance(other, (int, _np.integer)):
            return (int(other),)
        try:
            return tuple(None if d is None else int(d) for d in other)
        except TypeError:
            return None

    def __eq__(self, other):
        return tuple(self) == self._coerce(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = tuple.__hash__

    def __getitem__(self, k):
        r = tuple.__getitem__(self, k)
        return TensorShape(r) if isinstance(k, slice) else r

    def as_list(self):
        return list(self)

    def __add__(self, other):
        return TensorShape(tuple(self) + tuple(self._coerce(other)))

    def __radd__(self, other):
        return TensorShape(tuple(self._coerce(other)) + tuple(self))

    @property
    def rank(self):
        return len(self)


class Tensor(_np.ndarray):
    """ndarray whose Python-level .shape is a TensorShape (the C-level shape is untouched)."""

    @property
    def shape(self):
        return TensorShape(_np.ndarray.shape.__get__(self))

    def numpy(self):
        return _np.asarray(self).view(_np.ndarray)

    def __bool__(self):
        return _pybool(_np.asarray(self).view(_np.ndarray).all()) if self.size == 1 else _np.ndarray.__bool__(self)

    def __hash__(self):
        return id(self)


def _a(x):
    """-> plain ndarray"""
    if isinstance(x, Tensor):
        return x.view(_np.ndarray)
    if isinstance(x, TensorShape):
        return _np.asarray(tuple(x), dtype=_np.int32)
    return _np.asarray(x)


def _t(x, dtype=None):
    a = _np.asarray(x)
    if dtype is not None:
        a = a.astype(dtype, copy=False)
    elif a.dtype == _np.float64 and not isinstance(x, _np.ndarray):
        a = a.astype(_np.float32)          # Python floats become fp32, as in TensorFlow
    elif a.dtype == _np.int64 and not isinstance(x, _np.ndarray):
        a = a.astype(_np.int32)
    return a.view(Tensor)


def _res(x, like=None):
    return _np.asarray(x).view(Tensor)


def _bin(a, b):
    """dtype rule of TF ops with a Python scalar operand: the scalar takes the tensor's dtype"""
    aa, bb = _a(a), _a(b)
    if not isinstance(a, _np.ndarray) and isinstance(b, _np.ndarray):
        aa = aa.astype(bb.dtype) if aa.dtype.kind in "fi" and bb.dtype.kind == "f" or aa.dtype.kind == bb.dtype.kind else aa
    if not isinstance(b, _np.ndarray) and isinstance(a, _np.ndarray):
        bb = bb.astype(aa.dtype) if bb.dtype.kind in "fi" and aa.dtype.kind == "f" or bb.dtype.kind == aa.dtype.kind else bb
    return aa, bb


def constant(v, dtype=None, shape=None):
    t = _t(v, dtype)
    return t if shape is None else _res(_np.broadcast_to(_a(t), tuple(shape)).copy())


def convert_to_tensor(v, dtype=None, name=None):
    return _t(v, dtype)


def cast(x, dtype):
    return _res(_a(x).astype(dtype))


def shape(x, out_type=int32):
    return _res(_np.asarray(_np.shape(_a(x)), dtype=out_type))


def _shape_tuple(shape_):
    if isinstance(shape_, (list, tuple)):
        return tuple(int(_a(v)) for v in shape_)
    return tuple(int(v) for v in _a(shape_).reshape(-1))


def zeros(shape_, dtype=float32):
    return _res(_np.zeros(_shape_tuple(shape_), dtype=dtype))


def ones(shape_, dtype=float32):
    return _res(_np.ones(_shape_tuple(shape_), dtype=dtype))


def zeros_like(x, dtype=None):
    return _res(_np.zeros_like(_a(x), dtype=dtype))


def ones_like(x, dtype=None):
    return _res(_np.ones_like(_a(x), dtype=dtype))


def eye(n, dtype=float32):
    return _res(_np.eye(int(n), dtype=dtype))


def range(*args, dtype=int32):   # noqa: A001
    return _res(_np.arange(*[int(_a(v)) for v in args], dtype=dtype))


def gather(params, indices, axis=0, batch_dims=0):
    p, i = _a(params), _a(indices)
    ax = int(_a(axis))
    return _res(_np.take(p, i.astype(_np.int64), axis=ax))


def stack(values, axis=0):
    vs = [_a(v) for v in values]
    dt = _np.result_type(*[v.dtype for v in vs])
    if all(v.dtype.kind in "iu" for v in vs):
        dt = _np.int32
    return _res(_np.stack([v.astype(dt) for v in vs], axis=axis))


def concat(values, axis):
    vs = [_a(v) for v in values]
    vs = [v.reshape(-1) if v.ndim == 0 else v for v in vs]
    return _res(_np.concatenate(vs, axis=int(_a(axis))))


def expand_dims(x, axis):
    return _res(_np.expand_dims(_a(x), int(_a(axis))))


def squeeze(x, axis=None):
    return _res(_np.squeeze(_a(x), axis=axis))


def reshape(x, shape_):
    return _res(_np.reshape(_a(x), _shape_tuple(shape_)))


def tile(x, multiples):
    return _res(_np.tile(_a(x), _shape_tuple(multiples)))


def broadcast_to(x, shape_):
    return _res(_np.broadcast_to(_a(x), _shape_tuple(shape_)).copy())


def transpose(x, perm=None):
    return _res(_np.transpose(_a(x), perm))


def where(cond_, x=None, y=None):
    if x is None:
        return _res(_np.argwhere(_a(cond_)))
    xa, ya = _bin(x, y)
    if xa.dtype != ya.dtype and xa.dtype.kind == "f" and ya.dtype.kind == "f":
        dt = xa.dtype if isinstance(x, _np.ndarray) else ya.dtype
        xa, ya = xa.astype(dt), ya.astype(dt)
    return _res(_np.where(_a(cond_), xa, ya))


def boolean_mask(x, mask, axis=None):
    return _res(_a(x)[_a(mask).astype(_np.bool_)])


def minimum(a, b):
    return _res(_np.minimum(*_bin(a, b)))


def maximum(a, b):
    return _res(_np.maximum(*_bin(a, b)))


def less(a, b):
    return _res(_np.less(*_bin(a, b)))


def exp(x):
    return _res(_np.exp(_a(x)))


def reduce_sum(x, axis=None, keepdims=False):
    return _res(_np.sum(_a(x), axis=axis, keepdims=keepdims, dtype=_a(x).dtype))


def reduce_max(x, axis=None, keepdims=False):
    return _res(_np.max(_a(x), axis=axis, keepdims=keepdims))


def reduce_logsumexp(x, axis=None, keepdims=False):
    a = _a(x)
    m = _np.max(a, axis=axis, keepdims=True)
    m0 = _np.where(_np.isfinite(m), m, 0).astype(a.dtype)
    with _np.errstate(divide="ignore"):
        r = _np.log(_np.sum(_np.exp(a - m0), axis=axis, keepdims=True, dtype=a.dtype)) + m0
    if not keepdims:
        r = _np.squeeze(r, axis=axis) if axis is not None else r.reshape(())
    return _res(r.astype(a.dtype))


def cumsum(x, axis=0, exclusive=False, reverse=False):
    a = _a(x)
    if reverse:
        a = _np.flip(a, axis)
    r = _np.cumsum(a, axis=axis, dtype=a.dtype)
    if exclusive:
        r = r - a
    if reverse:
        r = _np.flip(r, axis)
    return _res(r)


def argsort(values, axis=-1, direction="ASCENDING", stable=False):
    a = _a(values)
    # TensorFlow's argsort is a top_k on the GPU/CPU kernel: ties in index order; DESCENDING = ascending sort of -values
    if direction == "DESCENDING":
        return _res(_np.argsort(-a, axis=axis, kind="stable").astype(_np.int32))
    return _res(_np.argsort(a, axis=axis, kind="stable").astype(_np.int32))


def einsum(eq, *xs):
    return _res(_np.einsum(eq, *[_a(x) for x in xs]))


def ensure_shape(x, shape_):
    have = tuple(_np.shape(_a(x)))
    want = tuple(shape_)
    assert len(have) == len(want) and all(w is None or int(w) == h for h, w in zip(have, want)), (have, want)
    return x


def function(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


@_contextlib.contextmanager
def name_scope(name):
    yield name


def print(*a, **k):   # noqa: A001
    pass


def cond(pred, true_fn=None, false_fn=None, name=None):
    return true_fn() if _pybool(_a(pred).all()) else false_fn()


def while_loop(cond, body, loop_vars, parallel_iterations=10, maximum_iterations=None, **kw):   # noqa: A002
    vs = list(loop_vars)
    it = 0
    while _pybool(_a(cond(*vs)).all()):
        out = body(*vs)
        vs = list(out) if isinstance(out, (list, tuple)) else [out]
        it += 1
        if maximum_iterations is not None and it >= int(_a(maximum_iterations)):
            break
    if hasattr(loop_vars, "_fields"):
        return type(loop_vars)(*vs)
    return tuple(vs) if isinstance(loop_vars, tuple) else vs


class TensorArray:
    def __init__(self, dtype, size=0, dynamic_size=False, element_shape=None, clear_after_read=None, infer_shape=True, name=None, _items=None):
        self.dtype = dtype
        self._items = dict(_items) if _items is not None else {}
        self._size = int(_a(size))
        self._dynamic = dynamic_size

    def write(self, index, value):
        i = int(_a(index))
        new = TensorArray(self.dtype, max(self._size, i + 1) if self._dynamic else self._size, self._dynamic, _items=self._items)
        new._items[i] = _a(value).astype(self.dtype)
        return new

    def read(self, index):
        return _res(self._items[int(_a(index))])

    def size(self):
        return _res(_np.asarray(self._size, dtype=_np.int32))

    def stack(self):
        n = self._size if not self._dynamic else (max(self._items) + 1 if self._items else 0)
        if n == 0:
            return _res(_np.zeros((0,), dtype=self.dtype))
        return _res(_np.stack([self._items[i] for i in _np.arange(n)]))

    def gather(self, indices):
        return _res(_np.stack([self._items[int(i)] for i in _a(indices).reshape(-1)]))

    def unstack(self, value):
        v = _a(value)
        new = TensorArray(self.dtype, v.shape[0], self._dynamic)
        for i in _np.arange(v.shape[0]):
            new._items[int(i)] = v[i].astype(self.dtype)
        return new


class _Nest:
    @staticmethod
    def map_structure(fn, *structs, **kw):
        s0 = structs[0]
        if isinstance(s0, dict):
            return type(s0)((k, _Nest.map_structure(fn, *[s[k] for s in structs])) for k in s0)
        if isinstance(s0, tuple) and hasattr(s0, "_fields"):
            return type(s0)(*[_Nest.map_structure(fn, *[s[i] for s in structs]) for i in _np.arange(len(s0))])
        if isinstance(s0, (list, tuple)) and not isinstance(s0, TensorShape):
            return type(s0)(_Nest.map_structure(fn, *[s[i] for s in structs]) for i in _np.arange(len(s0)))
        return fn(*structs)

    @staticmethod
    def flatten(s):
        if isinstance(s, dict):
            return [x for k in sorted(s) for x in _Nest.flatten(s[k])]
        if isinstance(s, (list, tuple)) and not isinstance(s, TensorShape):
            return [x for v in s for x in _Nest.flatten(v)]
        return [s]


nest = _Nest()


class _Math:
    @staticmethod
    def log(x):
        with _np.errstate(divide="ignore", invalid="ignore"):
            return _res(_np.log(_a(_t(x) if not isinstance(x, _np.ndarray) else x)))

    @staticmethod
    def log1p(x):
        with _np.errstate(divide="ignore", invalid="ignore"):
            return _res(_np.log1p(_a(x)))

    @staticmethod
    def exp(x):
        return _res(_np.exp(_a(x)))

    @staticmethod
    def square(x):
        return _res(_np.square(_a(x)))

    @staticmethod
    def add(a, b):
        return _res(_np.add(*_bin(a, b)))

    @staticmethod
    def multiply(a, b):
        return _res(_np.multiply(*_bin(a, b)))

    @staticmethod
    def less(a, b):
        return _res(_np.less(*_bin(a, b)))

    @staticmethod
    def not_equal(a, b):
        return _res(_np.not_equal(*_bin(a, b)))

    @staticmethod
    def equal(a, b):
        return _res(_np.equal(*_bin(a, b)))

    @staticmethod
    def logical_and(a, b):
        return _res(_np.logical_and(_a(a), _a(b)))

    @staticmethod
    def logical_or(a, b):
        return _res(_np.logical_or(_a(a), _a(b)))

    @staticmethod
    def logical_not(a):
        return _res(_np.logical_not(_a(a)))

    @staticmethod
    def is_finite(x):
        return _res(_np.isfinite(_a(x)))

    @staticmethod
    def is_inf(x):
        return _res(_np.isinf(_a(x)))

    @staticmethod
    def is_nan(x):
        return _res(_np.isnan(_a(x)))

    cumsum = staticmethod(cumsum)
    reduce_sum = staticmethod(reduce_sum)
    reduce_logsumexp = staticmethod(reduce_logsumexp)
    minimum = staticmethod(minimum)
    maximum = staticmethod(maximum)

    @staticmethod
    def sigmoid(x):
        return _res((1.0 / (1.0 + _np.exp(-_a(x).astype(_np.float64)))).astype(_a(x).dtype))

    @staticmethod
    def log_sigmoid(x):
        a = _a(x).astype(_np.float64)
        return _res((-_np.logaddexp(0.0, -a)).astype(_a(x).dtype))

    @staticmethod
    def betainc(a, b, x):
        from scipy import special
        aa, bb, xx = _np.broadcast_arrays(_a(a), _a(b), _a(x))
        dt = _np.result_type(aa.dtype, bb.dtype, xx.dtype)
        return _res(special.betainc(aa.astype(_np.float64), bb.astype(_np.float64), xx.astype(_np.float64)).astype(dt))

    @staticmethod
    def lgamma(x):
        from scipy import special
        return _res(special.gammaln(_a(x).astype(_np.float64)).astype(_a(x).dtype))


math = _Math()
sigmoid = _Math.sigmoid


class _NN:
    @staticmethod
    def log_softmax(logits, axis=-1):
        a = _a(logits)
        return _res((a - _a(reduce_logsumexp(a, axis=axis, keepdims=True))).astype(a.dtype))

    @staticmethod
    def softmax(logits, axis=-1):
        a = _a(logits)
        with _np.errstate(invalid="ignore"):
            m = _np.max(a, axis=axis, keepdims=True)
            e = _np.exp(a - _np.where(_np.isfinite(m), m, 0))
            return _res((e / _np.sum(e, axis=axis, keepdims=True, dtype=a.dtype)).astype(a.dtype))

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(labels, logits):
        ls = _a(_NN.log_softmax(logits, -1))
        lab = _a(labels).astype(_np.int64)
        return _res(-_np.take_along_axis(ls, lab[..., None], -1)[..., 0])


nn = _NN()


class _Linalg:
    @staticmethod
    def set_diag(x, diagonal):
        a = _a(x).copy()
        d = _np.broadcast_to(_a(diagonal), a.shape[:-1]).astype(a.dtype)
        idx = _np.arange(a.shape[-1])
        a[..., idx, idx] = d
        return _res(a)


linalg = _Linalg()


class _Random:
    """tf.random.*: draws come from a hook so that a test can inject them (TensorFlow's own streams cannot be reproduced)."""
    uniform_hook = None        # callable(shape) -> ndarray of uniforms in [0, 1)
    categorical_hook = None    # callable(logits[batch, k], n) -> int array [batch, n]

    def uniform(self, shape_=(), minval=0.0, maxval=1.0, dtype=float32, seed=None):
        shp = tuple(int(v) for v in _a(shape_).reshape(-1))
        if self.uniform_hook is None:
            raise RuntimeError("shim_tf: tf.random.uniform called without an injected draw (set tensorflow.random.uniform_hook)")
        return _res(_np.asarray(self.uniform_hook(shp), dtype=dtype).reshape(shp))

    def categorical(self, logits, num_samples, dtype=int64, seed=None):
        if self.categorical_hook is None:
            raise RuntimeError("shim_tf: tf.random.categorical called without an injected draw")
        return _res(_np.asarray(self.categorical_hook(_a(logits), int(_a(num_samples))), dtype=dtype))


random = _Random()

from . import compat  # noqa: E402,F401
