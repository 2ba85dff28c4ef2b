"""Short, deterministic workload for ncu: two fit iterations (score pass, auction, centroid update) at
1 M x 512, K = 128.  Usage on the GPU box:
  python tools/profile_iter.py && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches.csv python tools/profile_iter.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from generative_ranking_recommender_b200.balancekmeans import KMeans

n = int(os.environ.get("PROF_ROWS", 1000000))
k = int(os.environ.get("PROF_K", 128))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(1234)
x = torch.randn((n, 512), device=dev, generator=g)
np.random.seed(42)
km = KMeans(n_clusters=k, device=dev, balanced=True)
km.cluster_centers = km.initialize(x)
for it in range(int(os.environ.get("PROF_ITERS", 2))):
    score, assign, stats, shift = km._iterate(x, n)
    torch.cuda.synchronize()
    print(f"iteration {it}: rounds {stats.rounds} passes {stats.passes} cold {stats.cold_passes} "
          f"misses {stats.window_misses} shift {shift:.4f}")
    from generative_ranking_recommender_b200 import engine as _e
    _ws = _e.SCRATCH.get("auction", 256, dev)
    print("   threshold sink per worker-round (<64, <128, <250, >=250 keys):", _ws[:128].view(torch.int32)[17:21].tolist())
