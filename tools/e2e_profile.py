"""Where the end-to-end train() time goes (bench.py's e2e leg with wall-clock timers around each stage).
   python tools/e2e_profile.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from generative_ranking_recommender_b200 import engine
from generative_ranking_recommender_b200 import hierarchical_rq_kmeans as H
from generative_ranking_recommender_b200.balancekmeans import KMeans

dev = torch.device("cuda:0")
n, DIM = int(os.environ.get("ROWS", 1000000)), 512
g = torch.Generator(device=dev); g.manual_seed(1)
x0 = torch.randn((n, DIM), device=dev, generator=g)
xh = torch.empty((n, DIM), dtype=torch.float32, pin_memory=True); xh.copy_(x0); x_np = xh.numpy(); del x0
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); T.setdefault(name, []).append(time.perf_counter() - t)
        return r
    return w
H.HierarchicalRQKMeans._h2d = staticmethod(timed("h2d", H.HierarchicalRQKMeans._h2d))
H.HierarchicalRQKMeans._train_layer_0 = timed("train_layer", H.HierarchicalRQKMeans._train_layer_0)
KMeans.fit_by_min_loss = timed("fit_by_min_loss", KMeans.fit_by_min_loss)
KMeans._iterate = timed("iterate", KMeans._iterate)
KMeans._draw = timed("draw(np.random.choice)", KMeans._draw)
engine.residual_normalise = timed("residual", engine.residual_normalise)
for rep in range(2):
    T.clear()
    cfg = H.HierarchicalRQKMeansConfig(layer_clusters=[128, 128, 256], need_clusters=[128, 128, 256], embedding_dim=DIM,
                                       group_dims=[DIM], hierarchical_weights=[[1.0]] * 3, iter_limit=20)
    np.random.seed(42); torch.manual_seed(42)
    m = H.HierarchicalRQKMeans(cfg, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = m.train(x_np, resume=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"run {rep}: train() {dt*1e3:.1f} ms, iterations {[len(s) for s in m.fit_stats]}")
    for k, v in T.items():
        print(f"   {k:28s} calls {len(v):3d} total {sum(v)*1e3:8.1f} ms  mean {np.mean(v)*1e3:7.2f}  max {max(v)*1e3:7.2f}")
    it = T["iterate"]
    print("   iterate ms:", " ".join(f"{t*1e3:.1f}" for t in it))
    print("   passes    :", " ".join(str(s["passes"]) for lv in m.fit_stats for s in lv))
