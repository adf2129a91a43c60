"""Small helpers for building model matrices from (possibly batched) parameter tensors."""
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
    vals = [p.to(device=device, dtype=dtype) if torch.is_tensor(p) else torch.tensor(float(p), dtype=dtype, device=device)
            for p in params]
    vals = list(torch.broadcast_tensors(*vals))
    batch = vals[0].shape
    if len(batch) > 1:
        raise ValueError("parameters may carry at most one leading sample axis")
    return vals, batch, dtype, device


def const(M, dtype, device):
    return torch.tensor(M, dtype=dtype, device=device)


def diag(vals):
    """list of (batched) scalars -> (..., k, k) diagonal matrix."""
    return torch.diag_embed(torch.stack(vals, -1))


def block_diag_const(block, dim, dtype, device):
    return torch.block_diag(*[torch.tensor(block, dtype=dtype, device=device)] * dim)


def swap_dims(d, dim):
    """Reference lqg/tracking/subjective.py:7-12: observed (target, cursor) pairs of all axes first, the rest last."""
    idx = list(range(d))
    k = d // dim
    obs = [idx[k * i:k * i + 2] for i in range(dim)]
    un = [idx[k * i + 2:k * (i + 1)] for i in range(dim)]
    return list(chain(*(obs + un)))
