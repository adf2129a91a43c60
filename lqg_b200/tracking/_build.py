"""Small helpers for building model matrices from (possibly batched) parameter tensors."""
import functools
from itertools import chain

import torch


def default_device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def canon(params, dtype=None, device=None):
    """Broadcast scalar / tensor parameters to a common batch shape () or (S,).  Returns (list, batch, dtype, device)."""
    tens = [p for p in params if torch.is_tensor(p)]
    if device is None:
        device = tens[0].device if tens else default_device()
    if dtype is None:
        dtype = next((p.dtype for p in tens if p.is_floating_point()), torch.float32)
    # Python scalars become device scalars with a fill kernel (torch.full), not a host-to-device copy: model construction
    # stays legal inside a CUDA-graph capture (lqg_b200.graphs)
    vals = [p.to(device=device, dtype=dtype) if torch.is_tensor(p) else torch.full((), float(p), dtype=dtype, device=device)
            for p in params]
    vals = list(torch.broadcast_tensors(*vals))
    batch = vals[0].shape
    if len(batch) > 1:
        raise ValueError("parameters may carry at most one leading sample axis")
    return vals, batch, dtype, device


def _freeze(M):
    return tuple(_freeze(r) for r in M) if isinstance(M, (list, tuple)) else float(M)


@functools.lru_cache(maxsize=256)
def _const_cached(frozen, dtype, device):
    return torch.tensor(frozen, dtype=dtype, device=device)


def const(M, dtype, device):
    """Constant matrix on `device`; cached per (values, dtype, device) so that repeated model construction does no
    host-to-device copy (and is legal inside a CUDA-graph capture once warmed up).  Callers must not modify the result."""
    return _const_cached(_freeze(M), dtype, torch.device(device))


@functools.lru_cache(maxsize=64)
def _index_cached(idx, device):
    return torch.tensor(idx, dtype=torch.long, device=device)


def index(idx, device):
    """Index list as a cached device tensor (see `const`)."""
    return _index_cached(tuple(int(i) for i in idx), torch.device(device))


def diag(vals):
    """list of (batched) scalars -> (..., k, k) diagonal matrix."""
    return torch.diag_embed(torch.stack(vals, -1))


@functools.lru_cache(maxsize=256)
def _block_diag_cached(frozen, dim, dtype, device):
    return torch.block_diag(*[torch.tensor(frozen, dtype=dtype)] * dim).to(device)


def block_diag_const(block, dim, dtype, device):
    """dim copies of a constant block on the diagonal (cached like `const`; do not modify the result)."""
    return _block_diag_cached(_freeze(block), dim, dtype, torch.device(device))


def swap_dims(d, dim):
    """Reference lqg/tracking/subjective.py:7-12: observed (target, cursor) pairs of all axes first, the rest last."""
    idx = list(range(d))
    k = d // dim
    obs = [idx[k * i:k * i + 2] for i in range(dim)]
    un = [idx[k * i + 2:k * (i + 1)] for i in range(dim)]
    return list(chain(*(obs + un)))
