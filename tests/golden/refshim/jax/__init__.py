"""Minimal float64 NumPy stand-in for the `jax` names the reference hot path imports (see ../README.md)."""
import numpy as _np

from . import numpy, lax, random, scipy  # noqa: F401
from .numpy import Array  # noqa: F401


def _tree_map(fn, tree):
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_map(fn, t) for t in tree) if not hasattr(tree, "_fields") else type(tree)(*[_tree_map(fn, t) for t in tree])
    return fn(tree)


def _stack_tree(items):
    first = items[0]
    if isinstance(first, (tuple, list)):
        cols = [_stack_tree([it[i] for it in items]) for i in range(len(first))]
        return type(first)(*cols) if hasattr(first, "_fields") else type(first)(cols)
    return _np.stack([_np.asarray(it) for it in items]).view(Array)


def vmap(fun, in_axes=0, out_axes=0):
    """Loop over the mapped leading axis and stack the outputs (in_axes: int/None or a tuple of them; axis 0 only)."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert all(a in (0, None) for a in axes) and out_axes == 0
        n = next(len(a) for a, ax in zip(args, axes) if ax == 0)
        outs = [fun(*[(a[i] if ax == 0 else a) for a, ax in zip(args, axes)]) for i in range(n)]
        return _stack_tree(outs)
    return mapped
