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
