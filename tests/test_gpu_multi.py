"""Two-GPU parity (NCCL): the sample-sharded and the trial-sharded evaluation through lqg_b200.parallel reproduce the
single-GPU result.  Needs two visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped
on a one-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
NAMES = ("action_cost", "action_variability", "subj_noise", "subj_vel_noise", "sigma_target", "sigma_cursor")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=dev)
    try:
        import bench
        from lqg_b200 import parallel
        from lqg_b200.tracking import SubjectiveActor
        T, N, S = 300, 37, 45                                          # uneven splits: 23 + 22 samples, 19 + 18 trials
        X = torch.tensor(bench.make_data(N, T), device=dev)
        theta = torch.tensor(bench.make_theta(S, 41), device=dev)

        def per_sample(th, x):                                        # (ll[S_loc], grad[S_loc, 6]) through the public API
            th = th.detach().requires_grad_()
            m = SubjectiveActor(dim=2, T=T, device=dev, **{n: th[:, i] for i, n in enumerate(NAMES)})
            ll = m.log_likelihood(x).sum(-1)
            ll.sum().backward()
            return ll.detach(), th.grad

        ll1, g1 = per_sample(theta, X)                                # single-GPU reference, all samples, all trials
        # (1) parameter samples sharded over the ranks, results all-gathered
        ll2, g2 = parallel.sharded_value_and_grad(lambda th: per_sample(th, X), theta)
        assert ll2.shape == (S,) and g2.shape == (S, 6)
        assert torch.allclose(ll2, ll1, rtol=1e-6, atol=0) and torch.allclose(g2, g1, rtol=1e-6, atol=1e-6 * g1.abs().max().item())
        # (2) one parameter vector, its trials sharded over the ranks, ONE all-reduce of [sum ll, grad]
        th0 = theta[:1]

        def trial_shard(x):
            ll, g = per_sample(th0, x)
            return ll.sum(), g[0]

        tot, gt = parallel.trial_sharded_value_and_grad(trial_shard, X)
        assert torch.allclose(tot, ll1[0], rtol=1e-6)
        assert torch.allclose(gt, g1[0], rtol=1e-5, atol=1e-6 * g1[0].abs().max().item()), (gt, g1[0])
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sample_and_trial_sharding_two_gpus(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
