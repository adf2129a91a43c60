"""`jax.lax.scan` stand-in: a Python loop that threads the carry and stacks the per-step outputs."""
import numpy as _np

from .numpy import Array


def _leaves_len(xs):
    if isinstance(xs, (tuple, list)):
        return _leaves_len(xs[0])
    return len(xs)


def _index(xs, t):
    if isinstance(xs, (tuple, list)):
        return tuple(_index(x, t) for x in xs)
    return xs[t]


def _stack(items):
    first = items[0]
    if isinstance(first, (tuple, list)):
        return tuple(_stack([it[i] for it in items]) for i in range(len(first)))
    return _np.stack([_np.asarray(it) for it in items]).view(Array)


def scan(f, init, xs, length=None, reverse=False):
    n = _leaves_len(xs) if xs is not None else length
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for t in order:
        carry, ys[t] = f(carry, _index(xs, t) if xs is not None else None)
    return carry, _stack(ys)
