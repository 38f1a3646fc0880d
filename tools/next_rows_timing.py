"""Timing of the 'next' rows at BASELINE shapes: per_cluster_confidence (1M query cells) and the no-Harmony
k-means fallback (500k reference cells, K=100)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from symphonypy_b200 import engine

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
N, d = 1_000_000, 30
X = torch.randn((N, d), generator=g, device=dev) + 5
for G in (1, 30, 1000):
    codes = torch.randint(0, G, (N,), generator=g, device=dev, dtype=torch.int32)
    engine.group_moments(X, codes, G)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        size, mean, sc = engine.group_moments(X, codes, G)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"group_moments N={N} d={d} G={G}: {ms:.2f} ms  ({2*N*d*d/ms/1e9:.2f} TF/s, {2*N*d*4/ms/1e6:.0f} GB/s)")
M, K = 500_000, 100
cent = torch.randn((50, d), generator=g, device=dev) * 2
Z = cent[torch.randint(0, 50, (M,), generator=g, device=dev)] + torch.randn((M, d), generator=g, device=dev)
engine.kmeans(Z[:10000].contiguous(), 10, n_init=1, seed=0)
torch.cuda.synchronize(); t0 = time.perf_counter()
C, inertia = engine.kmeans(Z, K, seed=0)
torch.cuda.synchronize()
print(f"kmeans M={M} d={d} K={K} n_init=10 max_iter=25: {time.perf_counter()-t0:.2f} s, inertia/M = {inertia/M:.3f}")
