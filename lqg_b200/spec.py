"""LQG specification container -- mirrors the reference's ``lqg/spec.py:5-19`` (same field order).

Arrays are torch tensors shaped ``(T, rows, cols)`` like the reference's time-stacked jnp arrays, optionally with
one leading parameter-sample axis ``(S, T, rows, cols)`` (what ``jax.vmap`` adds in the reference).  The time
axis of a time-invariant model is a stride-0 ``expand`` (``lqg_b200.utils.time_stack``), so nothing is copied
T times and the kernels can tell that the spec is time-invariant.
"""
from typing import NamedTuple

import torch


class LQGSpec(NamedTuple):
    """(generalized) LQG specification"""

    Q: torch.Tensor
    q: torch.Tensor
    Qf: torch.Tensor
    qf: torch.Tensor
    P: torch.Tensor
    R: torch.Tensor
    r: torch.Tensor
    A: torch.Tensor
    B: torch.Tensor
    V: torch.Tensor
    F: torch.Tensor
    W: torch.Tensor
