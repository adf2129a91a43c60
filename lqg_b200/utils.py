"""Time stacking -- mirrors ``lqg/utils.py:6-35`` (``time_stack``, ``time_stack_spec``) without materialising copies."""
import weakref

import torch

from lqg_b200.spec import LQGSpec

# Tensors this module created as all-zero affine terms (q, P, r of time_stack_spec): lets the likelihood path know that
# P == 0 without reading device memory (a host synchronisation).  Identity-based; copies (.to()) fall back to a check.
_KNOWN_ZERO = {}   # id(tensor) -> weakref (tensors compare element-wise, so no WeakSet)


def _mark_zero(t: torch.Tensor):
    key = id(t)
    _KNOWN_ZERO[key] = weakref.ref(t, lambda _r, k=key: _KNOWN_ZERO.pop(k, None))


def is_known_zero(t: torch.Tensor) -> bool:
    r = _KNOWN_ZERO.get(id(t))
    return r is not None and r() is t


def time_stack(A: torch.Tensor, T: int) -> torch.Tensor:
    """``(..., r, c) -> (..., T, r, c)`` as a stride-0 view (reference: ``jnp.stack((A,) * T)``, utils.py:6-7)."""
    return A.unsqueeze(-3).expand(*A.shape[:-2], T, *A.shape[-2:])


def time_stack_spec(A, B, F, V, W, Q, R, T: int) -> LQGSpec:
    """Reference ``lqg/utils.py:10-35``: stack the matrices, zero affine terms, ``Qf = Q[-1]``."""
    batch = torch.broadcast_shapes(*[m.shape[:-2] for m in (A, B, F, V, W, Q, R)])
    A, B, F, V, W, Q, R = [m.expand(*batch, *m.shape[-2:]) for m in (A, B, F, V, W, Q, R)]
    state_dim, action_dim = Q.shape[-1], R.shape[-1]
    kw = dict(dtype=A.dtype, device=A.device)
    q = torch.zeros(*batch, 1, state_dim, **kw).expand(*batch, T, state_dim)
    P = torch.zeros(*batch, 1, action_dim, state_dim, **kw).expand(*batch, T, action_dim, state_dim)
    r = torch.zeros(*batch, 1, action_dim, **kw).expand(*batch, T, action_dim)
    for z in (q, P, r):
        _mark_zero(z)
    return LQGSpec(A=time_stack(A, T), B=time_stack(B, T), F=time_stack(F, T), V=time_stack(V, T),
                   W=time_stack(W, T), Q=time_stack(Q, T), R=time_stack(R, T),
                   q=q, Qf=Q, qf=q[..., -1, :], P=P, r=r)
