#!/usr/bin/env python
"""Print the launch timeline of one lqgk_loglik_vjp call (CUDA events recorded by the library around every launch).
Usage (on a GPU box): python tools/timeline.py [--samples 16384] [--streams 8]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from lqg_b200 import abi, runtime  # noqa: E402
from lqg_b200.tracking import SubjectiveActor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=16384)
ap.add_argument("--streams", type=int, default=1)
ap.add_argument("--min-ms", type=float, default=0.0)
ap.add_argument("--segments", type=int, default=6)
ap.add_argument("--pipe-max", type=int, default=16384)
ap.add_argument("--summary", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = abi.load_library()
lib.set_streams(a.streams)
lib.set_pipeline(a.pipe_max, a.segments)
N, T, S = 100, 1200, a.samples
X = bench.make_data_gpu(N, T, dev)
theta = torch.tensor(bench.make_theta(S, 11), device=dev)
m = SubjectiveActor(dim=2, T=T, device=dev, **{n: theta[:, i] for i, n in enumerate(bench.PARAM_NAMES)})._axis_system
xk = X.reshape(N, T + 1, 2, 2).permute(2, 0, 1, 3).reshape(2 * N, T + 1, 2).contiguous()
x_tm = lib.pack_obs(xk)
dims = abi.LqgkDims(S, 2 * N, T, 2, 3, 1, 2, 2)
act = {k: runtime._row_major(getattr(m.actor, k)[:, 0]) for k in abi.ACTOR_KEYS}
dyn = {k: runtime._row_major(getattr(m.dynamics, k)[:, 0]) for k in abi.DYN_KEYS}
ws = runtime.workspace(dev, lib.workspace_bytes(dims, abi.MODE_VJP, 0))
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    lib.loglik_vjp(dims, act, dyn, x_tm, ws=ws, stream=st)
torch.cuda.synchronize()
lib.profile_enable(True)
lib.loglik_vjp(dims, act, dyn, x_tm, ws=ws, stream=st)
torch.cuda.synchronize()
tl = lib.profile_timeline()
lib.profile_read()
print(f"{len(tl)} launches, span {max(e for _, _, e in tl):.2f} ms, busy sum {sum(e - s for _, s, e in tl):.2f} ms")
if a.summary:
    kinds = []
    for k, s, e in tl:
        if k not in kinds:
            kinds.append(k)
    for k in kinds:
        rows = [(s, e) for kk, s, e in tl if kk == k]
        print(f"  {k:12s} x{len(rows):2d}  first start {min(r[0] for r in rows):6.2f}  last end {max(r[1] for r in rows):6.2f}  busy {sum(r[1] - r[0] for r in rows):6.2f}")
    sys.exit(0)
for k, s, e in sorted(tl, key=lambda r: r[1]):
    if e - s >= a.min_ms:
        print(f"{s:8.2f} -> {e:8.2f}  ({e - s:6.2f} ms)  {k}")
