"""Latency regime: thread-per-sample vs warp-per-sample covariance kernels at small sample counts (configs c2 / c5):
python tools/bench_small_s.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lqg_b200 import abi
from lqg_b200.tracking import SubjectiveActor

dev = torch.device("cuda:0")
lib = abi.load_library()
N, T = 100, 1200
x = bench.make_data_gpu(N, T, dev)


def run(S, reps=5):
    theta = torch.tensor(bench.make_theta(S, 11), device=dev, requires_grad=True)

    def ev():
        theta.grad = None
        m = SubjectiveActor(dim=2, T=T, device=dev, **{n: theta[:, i] for i, n in enumerate(bench.PARAM_NAMES)})
        ll = m.log_likelihood(x).sum()
        ll.backward()
        return ll
    ev(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ev()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("| samples | thread-per-sample covariance kernels (ms) | warp-per-sample covariance kernels (ms) |\n|---:|---:|---:|")
for S in (8, 64, 512, 1024, 2048, 4096, 8192):
    lib.set_warp_cov_max_samples(0)
    t0 = run(S)
    lib.set_warp_cov_max_samples(1 << 30)
    t1 = run(S)
    print(f"| {S} | {t0:.2f} | {t1:.2f} |")
lib.set_warp_cov_max_samples(512)
