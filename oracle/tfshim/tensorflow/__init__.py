"""numpy-backed stand-in for the handful of TensorFlow ops madflow's hot path uses.

TEST INFRASTRUCTURE ONLY.  TensorFlow is not installable offline, so the reference's own
Python sources (/root/reference/python_package/madflow/{wavefunctions_flow,phasespace,
parameters}.py and tests/mockup_debug_me.py) cannot be executed as they are.  This package
lets `tests/golden/make_golden.py` import those UNMODIFIED sources in the build container
and run them eagerly on numpy arrays, so the committed golden vectors come from the
reference's code and not from our restatement of it.

Only dtype-conversion semantics that change numbers are modelled with care:
  * a bare Python float/int (or list of them) handed to an op is first turned into a
    float32/int32 tensor, exactly like `tf.convert_to_tensor` does, before any cast --
    this is what makes `float_me(np.pi)`, `float_me(389379365.6)`, `float_me(1e-10)` and
    `SQH = float_me(tf.math.sqrt(0.5))` float32-rounded in the reference;
  * a Python scalar combined arithmetically with a tensor takes the tensor's dtype
    (that is numpy's behaviour too).
Everything else is a thin alias of the numpy function with the same meaning.
"""
import builtins
import numpy as np
from scipy import special as _sp

from . import math  # noqa: F401  (tensorflow.math)
from .math import (_t, sqrt, exp, square, pow, reduce_sum, reduce_prod, reduce_all,  # noqa: F401
                   maximum, minimum, tanh)

float64 = np.dtype("float64")
float32 = np.dtype("float32")
int32 = np.dtype("int32")
int64 = np.dtype("int64")
complex128 = np.dtype("complex128")
bool = np.dtype("bool")  # noqa: A001


class TensorSpec:
    def __init__(self, shape=None, dtype=None, name=None):
        self.shape, self.dtype, self.name = shape, dtype, name


class _Function:
    """Result of tf.function: calls straight through (eager)."""

    def __init__(self, fn, input_signature=None):
        self.python_function = fn
        self.input_signature = input_signature
        self.__doc__ = getattr(fn, "__doc__", None)
        self.__name__ = getattr(fn, "__name__", "fn")

    def __call__(self, *a, **k):
        return self.python_function(*a, **k)

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        bound = self.python_function.__get__(obj, objtype)
        return _Function(bound, self.input_signature)


def function(fn=None, input_signature=None, **_):
    if fn is None:
        return lambda f: _Function(f, input_signature)
    return _Function(fn, input_signature)


def executing_eagerly():
    return True


def convert_to_tensor(x, dtype=None):
    x = _t(x)
    return x if dtype is None else x.astype(dtype)


def cast(x, dtype):
    return np.asarray(_t(x)).astype(dtype)


def constant(x, dtype=None):
    return cast(x, dtype) if dtype is not None else _t(x)


def complex(real, imag):  # noqa: A001
    real, imag = np.asarray(_t(real)), np.asarray(_t(imag))
    out = np.empty(np.broadcast(real, imag).shape, dtype=np.complex128)
    out.real, out.imag = real, imag
    return out


def stack(values, axis=0):
    values = [np.asarray(_t(v)) for v in values]
    return np.stack(np.broadcast_arrays(*values) if len({v.shape for v in values}) > 1 else values, axis=axis)


def concat(values, axis=0):
    return np.concatenate([np.asarray(_t(v)) for v in values], axis=axis)


def expand_dims(x, axis):
    return np.expand_dims(_t(x), axis)


def transpose(x):
    return np.transpose(_t(x))


def reshape(x, shape):
    return np.reshape(_t(x), shape)


def where(cond, x=None, y=None):
    if x is None:
        return np.argwhere(cond)
    return np.where(cond, _t(x), _t(y))


def cond(pred, true_fn, false_fn):
    return true_fn() if builtins.bool(np.all(pred)) else false_fn()


def ones_like(x, dtype=None):
    return np.ones_like(_t(x), dtype=dtype)


def zeros_like(x, dtype=None):
    return np.zeros_like(_t(x), dtype=dtype)


def zeros(shape, dtype=float32):
    return np.zeros(shape, dtype=dtype)


def ones(shape, dtype=float32):
    return np.ones(shape, dtype=dtype)


def fill(dims, value):
    return np.full(dims, value)


def shape(x, out_type=int32):
    return np.asarray(np.shape(x), dtype=out_type)


def boolean_mask(x, mask, axis=0):
    return np.compress(mask, _t(x), axis=axis)


def gather(params, indices, axis=0, batch_dims=0):
    assert batch_dims == 0
    return np.take(_t(params), indices, axis=axis)


def scatter_nd(indices, updates, shape):
    out = np.zeros(tuple(int(s) for s in shape), dtype=np.asarray(updates).dtype)
    idx = np.asarray(indices)
    np.add.at(out, tuple(idx[:, k] for k in range(idx.shape[1])), updates)
    return out


def while_loop(cond, body, loop_vars, parallel_iterations=10, maximum_iterations=None):
    it = 0
    loop_vars = tuple(loop_vars)
    while builtins.bool(cond(*loop_vars)) and (maximum_iterations is None or it < maximum_iterations):
        loop_vars = tuple(body(*loop_vars))
        it += 1
    return loop_vars


def einsum(eq, *ops):
    return np.einsum(eq.replace(" ", ""), *[_t(o) for o in ops])


def logical_and(a, b):
    return np.logical_and(a, b)


class _Random:
    _rng = np.random.default_rng(0)

    def set_seed(self, seed):
        self._rng = np.random.default_rng(seed)

    def uniform(self, shape, minval=0.0, maxval=1.0, dtype=float32, seed=None):
        return (minval + (maxval - minval) * self._rng.random(tuple(shape))).astype(dtype)


random = _Random()


class _Backend:
    @staticmethod
    def batch_dot(x, y, axes=None):
        # keras: axes=k means (k, k): out[b,i,j] = sum_k x[b,i,k] * y[b,j,k] for rank-3 inputs
        assert axes == 2 and x.ndim == 3 and y.ndim == 3
        return np.einsum("bik,bjk->bij", x, y)


class _Keras:
    backend = _Backend()


keras = _Keras()


class _Config:
    @staticmethod
    def list_physical_devices(kind=None):
        return []


config = _Config()


def load_op_library(path):
    raise RuntimeError("tfshim: custom ops are not available")
