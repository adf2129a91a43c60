"""Multi-GPU plumbing for the likelihood path: one process per GPU, `torch.distributed` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).

The path shards naturally: every (parameter sample, condition) system is independent, so ranks own
disjoint slices of the sample axis and no collective touches the data path.  The only exchange is the one the
inference loop needs: either an all-gather of the per-sample results `(S/G, 1 + P)` (parameter sweeps, vectorised
chains) or one all-reduce of `[sum ll, sum grad]` (a single parameter vector whose trials / conditions are sharded).
Both are a few KB to a couple of MB: latency-bound, so a plain NCCL call on the compute stream is the right tool.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of `total` items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_samples(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather per-sample results sharded with `shard_range` along dim 0 -> full `(total, ...)` on every rank."""
    rank, ws = world(group)
    if ws == 1:
        return local
    sizes = [shard_range(total, r, ws) for r in range(ws)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def sharded_value_and_grad(fn: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], theta: torch.Tensor, group=None,
                           gather: bool = True):
    """Evaluate `fn(theta_local) -> (ll[S_local], grad[S_local, P])` on this rank's slice of the parameter samples
    `theta[S, P]` and (optionally) all-gather the results.  Returns `(ll, grad)` for all S samples when `gather`,
    else the local slices."""
    rank, ws = world(group)
    lo, hi = shard_range(theta.shape[0], rank, ws)
    ll, g = fn(theta[lo:hi])
    if not gather:
        return ll, g
    packed = torch.cat([ll[:, None], g], 1)
    full = gather_samples(packed, theta.shape[0], group)
    return full[:, 0], full[:, 1:]


def trial_sharded_value_and_grad(fn: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], x: torch.Tensor, group=None):
    """Single parameter vector, many trials: every rank evaluates `fn(x_local) -> (sum ll, grad[P])` on its slice of the
    trials (the per-sample recursions are repeated on every rank) and ONE all-reduce sums `[ll, grad]`."""
    rank, ws = world(group)
    lo, hi = shard_range(x.shape[0], rank, ws)
    ll, g = fn(x[lo:hi])
    packed = torch.cat([ll.reshape(1), g.reshape(-1)])
    allreduce_sum(packed, group)
    return packed[0], packed[1:]
