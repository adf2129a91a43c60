"""World-size-2 gloo tests of the multi-GPU host logic (sharding + the path's single collective), on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lqg_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        S, P = 11, 3                                                   # uneven split: 6 + 5
        theta = torch.arange(S * P, dtype=torch.float64).reshape(S, P)

        def fn(th):                                                    # stand-in for the CUDA evaluation
            return th.sum(1) * 2.0, th * 3.0

        ll, g = parallel.sharded_value_and_grad(fn, theta)
        assert torch.equal(ll, theta.sum(1) * 2.0) and torch.equal(g, theta * 3.0)
        lo, hi = parallel.shard_range(S, rank, world_size)
        ll_loc, _ = parallel.sharded_value_and_grad(fn, theta, gather=False)
        assert ll_loc.shape[0] == hi - lo

        x = torch.arange(14, dtype=torch.float64).reshape(7, 2)       # 7 trials: 4 + 3

        def fn2(xl):
            return xl.sum(), xl.sum(0)

        tot, gt = parallel.trial_sharded_value_and_grad(fn2, x)
        assert tot.item() == x.sum().item() and torch.equal(gt, x.sum(0))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for total in (1, 5, 64, 65536 + 3):
        for ws in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_sample_and_trial_sharding_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_single_process_is_a_noop():
    t = torch.ones(3)
    assert parallel.allreduce_sum(t) is t
    assert parallel.world() == (0, 1)


def _worker_more_ranks_than_units(rank, world_size, port, out_dir):
    """3 ranks, 2 conditions (bench.py's c2 at 8 GPUs has 6 conditions): the rank without work contributes zeros and runs the
    SAME collectives as the others -- a rank-dependent number of all-reduces deadlocks NCCL (seen once at N=8)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        cond = torch.arange(2)
        w = torch.tensor([2.0, 5.0], dtype=torch.float64)

        def local(idx):
            if idx.numel() == 0:
                return torch.zeros((), dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
            g = torch.zeros(2, dtype=torch.float64)
            g[idx] = w[idx]
            return w[idx].sum(), g

        for _ in range(3):                                            # same count of collectives on every rank
            tot, g = parallel.trial_sharded_value_and_grad(local, cond)
            assert tot.item() == 7.0 and torch.equal(g, w)
        lo, hi = parallel.shard_range(2, rank, world_size)
        packed = torch.cat([w[lo:hi].sum().reshape(1), torch.zeros(2, dtype=torch.float64)]) if hi > lo else torch.zeros(3, dtype=torch.float64)
        assert parallel.allreduce_sum(packed)[0].item() == 7.0
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_more_ranks_than_work_units(tmp_path):
    port = _free_port()
    mp.spawn(_worker_more_ranks_than_units, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(3))
