"""Size-independent properties of one fit iteration at the north-star size (10 M x 512 on one GPU): every cluster gets
N // K jobs (the N % K leftovers land on worker 0, reference balancekmeans/__init__.py:88-89), the run is reproducible,
the cluster sums are linear (sum of centroids * counts == sum of rows).   PROF_ROWS=10000000 python tools/fullsize_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from generative_ranking_recommender_b200 import engine
from generative_ranking_recommender_b200.balancekmeans import KMeans

n = int(os.environ.get("PROF_ROWS", 10000000))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1234)
x = torch.empty((n, 512), device=dev)
for i in range(0, n, 1000000):
    x[i:i + 1000000] = torch.randn((min(1000000, n - i), 512), device=dev, generator=g)
for k in (128, 256):
    res = []
    for rep in range(2):
        np.random.seed(42)
        km = KMeans(n_clusters=k, device=dev, balanced=True)
        km.cluster_centers = km.initialize(x)
        score, assign, stats, shift = km._iterate(x, n)
        res.append((assign.clone(), km.cluster_centers.clone(), stats))
    a, c, stats = res[0]
    sizes = torch.bincount(a.long(), minlength=k)
    want = torch.full((k,), n // k, dtype=sizes.dtype, device=dev)
    if n % k:
        want[0] += n % k
    ok_sizes = bool(torch.equal(sizes, want)) if (n % k or stats.rounds < 1002) else None
    same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    lin = (c.double() * sizes.double().unsqueeze(1)).sum(0)
    tot = torch.zeros(512, dtype=torch.float64, device=dev)
    for i in range(0, n, 1000000):
        tot += x[i:i + 1000000].double().sum(0)
    rel = float((lin - tot).abs().max() / tot.abs().max())
    print(f"n={n} K={k}: rounds {stats.rounds} (passes {stats.passes}), sizes balanced {ok_sizes} (min {int(sizes.min())} max {int(sizes.max())}), "
          f"reproducible {same}, sum linearity rel err {rel:.1e}", flush=True)
