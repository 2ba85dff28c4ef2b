"""In-situ breakdown of one fit iteration (CUDA events around the three phases) at a given size.
   PROF_ROWS=10000000 PROF_K=128 python tools/iter_breakdown.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from generative_ranking_recommender_b200 import engine
from generative_ranking_recommender_b200.balancekmeans import KMeans

n = int(os.environ.get("PROF_ROWS", 1000000)); k = int(os.environ.get("PROF_K", 128))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1234)
x = torch.empty((n, 512), device=dev)
for i in range(0, n, 1000000):
    x[i:i + 1000000] = torch.randn((min(1000000, n - i), 512), device=dev, generator=g)
np.random.seed(42)
km = KMeans(n_clusters=k, device=dev, balanced=True)
km.cluster_centers = km.initialize(x)
buf = None
peak = 6551.4
for it in range(int(os.environ.get("PROF_ITERS", 3))):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    score = engine.score_pass(x, km.cluster_centers, scores=True, argmin=True, counts=True, scores_out=buf)
    buf = score.scores_t
    ev[1].record()
    assign, stats = km._assign(x, score, n)
    ev[2].record()
    shift = km._update(x, assign, n)
    ev[3].record(); torch.cuda.synchronize()
    ts, ta, tu = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
    rounds = (stats.passes + 1) // 2
    bytes_per_vec = 8 * 512 + 2 * k + 2 * k * rounds          # SURVEY.md 8d composite bound
    bound_ms = bytes_per_vec * n / (peak * 1e9) * 1e3
    tot = ts + ta + tu
    print(f"n={n} K={k} it {it}: score {ts:.2f} ms, auction {ta:.2f} ms ({stats.passes} passes, {rounds} rounds, "
          f"{ta / max(rounds, 1) * 1e3:.0f} us/round, misses {stats.window_misses}, list rounds {stats.list_passes}), "
          f"update {tu:.2f} ms, total {tot:.2f} ms = {n / tot / 1e3:.1f} M vectors/s; composite HBM bound {bound_ms:.2f} ms "
          f"-> {bound_ms / tot:.2f} of roofline", flush=True)
