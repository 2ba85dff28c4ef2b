"""Multi-GPU check of the row-sharded fit (run under torchrun, one rank per GPU):
   * one sharded iteration == the unsharded iteration on the concatenated rows (assignment bit-identical,
     centroids to fp32 summation order), 
   * timing of sharded iterations (CUDA events, max over ranks).
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from generative_ranking_recommender_b200 import engine
from generative_ranking_recommender_b200.balancekmeans import KMeans

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shard = engine.ShardGroup()

def make(nloc, r, kind="iso"):
    g = torch.Generator(device=dev)
    g.manual_seed(100 + r)
    if kind == "iso":
        return torch.randn((nloc, 512), device=dev, generator=g)
    # clustered, unit-norm rows (the shape of a level-2 residual): dense fp16 score values, wide windows, window
    # misses - the partial HIST passes and their rank-major tie offsets get exercised
    gc = torch.Generator(device=dev)
    gc.manual_seed(99)
    c = torch.randn((96, 512), device=dev, generator=gc)
    x = c[torch.randint(0, 96, (nloc,), device=dev, generator=g)] + 0.35 * torch.randn((nloc, 512), device=dev, generator=g)
    return x / x.norm(dim=1, keepdim=True)

ok = True
for nloc, k, kind in ((30011, 64, "iso"), (100000, 128, "iso"), (150001, 256, "mix"), (60000, 128, "mix")):
    x = make(nloc, rank, kind)
    np.random.seed(7)
    km = KMeans(n_clusters=k, device=dev, balanced=True, shard=shard)
    km.cluster_centers = km.initialize(x)
    c0 = km.cluster_centers.clone()
    n_global = nloc * world
    score, assign, stats, shift = km._iterate(x, n_global)
    gathered = shard.all_gather(assign)                       # [world, nloc]
    if rank == 0:
        xa = torch.cat([make(nloc, r, kind) for r in range(world)])
        np.random.seed(7)
        k1 = KMeans(n_clusters=k, device=dev, balanced=True)
        k1.cluster_centers = k1.initialize(xa)
        same_init = torch.equal(k1.cluster_centers, c0)
        s1, a1, st1, sh1 = k1._iterate(xa, n_global)
        same_assign = torch.equal(a1, gathered.reshape(-1))
        cerr = (k1.cluster_centers - km.cluster_centers).abs().max().item() / k1.cluster_centers.abs().max().item()
        print(f"n={n_global} k={k} {kind}: misses {stats.window_misses}/{st1.window_misses}; init identical {same_init}; assignment identical {same_assign}; rounds {stats.rounds}/{st1.rounds} "
              f"passes {stats.passes}/{st1.passes}; centroid max rel diff {cerr:.2e}; shift {shift:.4f}/{sh1:.4f}", flush=True)
        ok &= same_init and same_assign and cerr < 1e-5
    dist.barrier()

# timing: 1 M rows per rank, K = 128
nloc, k = 1000000, 128
x = make(nloc, rank)
np.random.seed(11)
km = KMeans(n_clusters=k, device=dev, balanced=True, shard=shard)
km.cluster_centers = km.initialize(x)
buf = None
for it in range(2):
    score, *_ = km._iterate(x, nloc * world, buf)
    buf = score.scores_t
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
passes = []
for it in range(4):
    score, a, stats, sh = km._iterate(x, nloc * world, buf)
    passes.append(stats.passes)
e1.record(); dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 4], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"sharded iteration, {world} x {nloc} rows, K={k}: {t.item():.2f} ms/iteration, passes {passes}, "
          f"{nloc * world / t.item() * 1e3 / 1e6:.1f} M vectors/s", flush=True)
    print("DIST CHECK", "OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()
