"""tensorflow.math stand-in (see package docstring).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
from scipy import special as _sp


def _t(x):
    """tf.convert_to_tensor dtype inference: bare Python numbers become 32-bit tensors."""
    if type(x) is float:
        return np.float32(x)
    if type(x) is int:
        return np.int32(x)
    if type(x) is complex:
        return np.complex128(x)
    if isinstance(x, (list, tuple)):
        if len(x) and all(type(v) is float or type(v) is int for v in x):
            if any(type(v) is float for v in x):
                return np.asarray(x, dtype=np.float32)
            return np.asarray(x, dtype=np.int32)
        if len(x) and all(isinstance(v, (list, tuple)) for v in x):
            flat = [w for v in x for w in v]
            if all(type(w) is float or type(w) is int for w in flat):
                dt = np.float32 if any(type(w) is float for w in flat) else np.int32
                return np.asarray(x, dtype=dt)
        return np.asarray(x)
    return x


def sqrt(x):
    return np.sqrt(_t(x))


def log(x):
    return np.log(_t(x))


def exp(x):
    return np.exp(_t(x))


def square(x):
    return np.square(_t(x))


def pow(x, y):  # noqa: A001
    x = _t(x)
    if type(y) in (float, int):  # python scalar takes the tensor's dtype
        return np.power(x, np.asarray(y, dtype=np.asarray(x).dtype))
    return np.power(x, _t(y))


def lgamma(x):
    return _sp.gammaln(_t(x))


def sign(x):
    return np.sign(_t(x))


def abs(x):  # noqa: A001
    return np.abs(_t(x))


def real(x):
    return np.real(_t(x))


def imag(x):
    return np.imag(_t(x))


def conj(x):
    return np.conj(_t(x))


def minimum(a, b):
    return np.minimum(_t(a), _t(b))


def maximum(a, b):
    return np.maximum(_t(a), _t(b))


def sin(x):
    return np.sin(_t(x))


def cos(x):
    return np.cos(_t(x))


def sinh(x):
    return np.sinh(_t(x))


def cosh(x):
    return np.cosh(_t(x))


def tanh(x):
    return np.tanh(_t(x))


def reduce_sum(x, axis=None, keepdims=False):
    if isinstance(x, (list, tuple)):
        x = np.stack([np.asarray(_t(v)) for v in x])
    return np.sum(x, axis=axis, keepdims=keepdims)


def reduce_prod(x, axis=None, keepdims=False):
    return np.prod(_t(x), axis=axis, keepdims=keepdims)


def reduce_all(x, axis=None):
    if isinstance(x, (list, tuple)):
        x = np.stack(x)
    return np.all(x, axis=axis)
