"""A numpy stand-in for the sliver of the Theano API that gccNMF/realtime/gccNMFProcessor.py uses.

TEST INFRASTRUCTURE ONLY (see gccnmf_oracle.py).  Theano is not installed in the build container and cannot be (no network),
so the reference's real-time class cannot be imported as it stands.  This module lets `oracle/make_golden.py` import and run the
UNMODIFIED reference class: `install()` registers a module named `theano` whose `shared`, `function` and `tensor.{tensor3, dot,
argmax, switch, exp, sum}` build a lazy expression graph evaluated with numpy -- nothing of the reference's own code is
restated.  What the stand-in decides, and Theano would decide the same way for this graph:
  * dtype promotion is numpy's: float32 (.) float32 -> float32, int64 - float32 -> float64 (Theano's scalar upcast gives
    float64 too), complex64 arithmetic stays complex64; the shared scalars are numpy.float32 values as the reference creates them;
  * `dot` on a 3-D by 2-D operand contracts the last axis with the first (tensor.dot = numpy.dot for these ranks).
What it cannot reproduce: the summation order inside Theano's BLAS call and its elementwise kernels -- fixtures made through
this module are therefore "the reference's code on numpy arithmetic", the closest executable statement of a13 there is here.
"""
import sys
import types

import numpy as np


class Node(object):
    """Lazy expression: `fn(env)` returns a numpy value; env maps input placeholders to arrays."""
    __array_priority__ = 1000.0       # numpy scalars / arrays on the left defer to Node.__r*__

    def __init__(self, fn):
        self._fn = fn

    def eval(self, env=None):
        return self._fn(env or {})

    # ---- structure
    def __getitem__(self, key):
        return Node(lambda env: self.eval(env)[key])

    @property
    def T(self):
        return Node(lambda env: self.eval(env).T)

    @property
    def real(self):
        return Node(lambda env: self.eval(env).real)

    def conj(self):
        return Node(lambda env: np.conj(self.eval(env)))

    # ---- arithmetic
    def _bin(self, other, op, swap=False):
        def run(env):
            a, b = self.eval(env), _value(other, env)
            return op(b, a) if swap else op(a, b)
        return Node(run)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, np.divide, True)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __pow__(self, o): return self._bin(o, np.power)
    def __lt__(self, o): return self._bin(o, np.less)
    def __neg__(self): return Node(lambda env: -self.eval(env))
    def __abs__(self): return Node(lambda env: np.abs(self.eval(env)))


def _value(x, env):
    return x.eval(env) if isinstance(x, Node) else x


class SharedVariable(Node):
    def __init__(self, value):
        self._value = value
        Node.__init__(self, lambda env: self._value)

    def get_value(self):
        return self._value

    def set_value(self, value):
        self._value = value


class Placeholder(Node):
    def __init__(self, name, dtype, ndim):
        self.name, self.dtype, self.ndim = name, np.dtype(dtype), ndim
        Node.__init__(self, lambda env: env[self])


def shared(value):
    return SharedVariable(value)


def function(inputs=(), outputs=()):
    inputs, outputs = list(inputs), list(outputs)

    def call(*args):
        env = {p: np.asarray(a, dtype=p.dtype) for p, a in zip(inputs, args)}
        return [np.asarray(o.eval(env)) for o in outputs]
    return call


def _tensor_module():
    t = types.ModuleType('theano.tensor')
    t.tensor3 = lambda name=None, dtype='float64': Placeholder(name, dtype, 3)
    t.dot = lambda a, b: Node(lambda env: np.dot(_value(a, env), _value(b, env)))
    t.argmax = lambda a, axis=None: Node(lambda env: np.argmax(_value(a, env), axis=axis))
    t.switch = lambda c, a, b: Node(lambda env: np.where(_value(c, env), _value(a, env), _value(b, env)))
    t.exp = lambda a: Node(lambda env: np.exp(_value(a, env)))
    t.sum = lambda a, axis=None, keepdims=False: Node(lambda env: np.sum(_value(a, env), axis=axis, keepdims=keepdims))
    return t


def install():
    """Registers the stand-in as `theano` (only when the real one is absent)."""
    if 'theano' in sys.modules:
        return sys.modules['theano']
    try:
        import theano  # noqa: F401
        return sys.modules['theano']
    except ImportError:
        pass
    m = types.ModuleType('theano')
    m.shared, m.function, m.tensor = shared, function, _tensor_module()
    m.__gccnmf_numpy_shim__ = True
    sys.modules['theano'] = m
    sys.modules['theano.tensor'] = m.tensor
    return m
