"""`jax.numpy` stand-in: NumPy float64 with an ndarray subclass that offers `.at[idx].set(v)`."""
import sys as _sys
import types as _types

import numpy as _np


class _AtIndexer:
    def __init__(self, arr, idx=None):
        self.arr, self.idx = arr, idx

    def __getitem__(self, idx):
        return _AtIndexer(self.arr, idx)

    def set(self, v):
        out = _np.array(self.arr, copy=True)
        out[self.idx] = v
        return out.view(Array)

    def add(self, v):
        out = _np.array(self.arr, copy=True)
        out[self.idx] += v
        return out.view(Array)


class Array(_np.ndarray):
    @property
    def at(self):
        return _AtIndexer(self)


ndarray = Array
newaxis = _np.newaxis
inf = _np.inf
pi = _np.pi
float32, float64, int32 = _np.float32, _np.float64, _np.int32


def _wrap(v):
    if isinstance(v, _np.ndarray):
        if v.dtype == _np.float32:
            v = v.astype(_np.float64)
        return v.view(Array)
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


def _lift(fn):
    def f(*a, **k):
        return _wrap(fn(*a, **k))
    f.__name__ = getattr(fn, "__name__", "f")
    return f


def array(obj, dtype=None):
    a = _np.array(obj, dtype=dtype)
    if a.dtype.kind == "f" or (a.dtype.kind in "iu" and dtype is None and not _all_int(obj)):
        a = a.astype(_np.float64)
    return a.view(Array)


def _all_int(obj):
    a = _np.asarray(obj)
    return a.dtype.kind in "iub"


def clip(a, min=None, max=None):
    return _wrap(_np.clip(a, min, max))


def array_str(a, max_line_width=None, precision=None, suppress_small=None):
    return _np.array_str(_np.asarray(a), max_line_width=10**9 if max_line_width in (None, _np.inf) else max_line_width,
                         precision=precision, suppress_small=suppress_small)


class _Linalg(_types.ModuleType):
    solve = staticmethod(_lift(_np.linalg.solve))
    inv = staticmethod(_lift(_np.linalg.inv))
    eigh = staticmethod(_lift(lambda a: tuple(_np.linalg.eigh(a))))
    cholesky = staticmethod(_lift(_np.linalg.cholesky))
    det = staticmethod(_lift(_np.linalg.det))
    slogdet = staticmethod(_lift(lambda a: tuple(_np.linalg.slogdet(a))))


linalg = _Linalg("jax.numpy.linalg")
_sys.modules["jax.numpy.linalg"] = linalg


def __getattr__(name):
    v = getattr(_np, name)
    return _lift(v) if callable(v) and not isinstance(v, type) else v
