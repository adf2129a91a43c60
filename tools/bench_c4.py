"""Times config c4 (large-system path) alone and prints the per-kernel breakdown: python tools/bench_c4.py [S] [N] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
S, N, T = (int(a) for a in (sys.argv[1:4] + ["4096", "50", "600"][len(sys.argv) - 1:]))
val, ms, k = bench.bench_c4(torch.device("cuda:0"), S=S, N=N, T=T)
print(f"c4 S={S} N={N} T={T}: {ms:.1f} ms per fwd+grad evaluation, {val:.0f} trial-evals/s")
for name, v in sorted(k.items(), key=lambda kv: -kv[1]):
    print(f"  {name:14s} {v:9.2f} ms")
