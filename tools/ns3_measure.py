"""BASELINE.json configs 4 and 5 (and the 1-GPU leg of config 3), measured with device-generated S-mix rows:

  encode   encode-only throughput, 3-level ids of [128,128,256], 1 M .. 100 M vectors per job; rows sharded over
           the ranks (no exchange), generated on the device in chunks that fit HBM, CUDA-event time summed over the
           chunks, max over ranks.  `python tools/ns3_measure.py encode` or under torchrun for N ranks.
  deep     [256,256,256,256] on 50 M x 512 on ONE GPU (102 GB of rows + 26 GB of fp16 scores): timed fit iterations
           at level 0 and on a level-1-shaped residual, composite fraction of SURVEY.md 8(d).
  Each prints one JSON line."""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench as B
from generative_ranking_recommender_b200 import engine
from generative_ranking_recommender_b200.balancekmeans import KMeans

what = sys.argv[1] if len(sys.argv) > 1 else "encode"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
peak, how = B.measured_peaks()


def allmax(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if what == "encode":
    # centroids of a short real fit on 1 M rows (so the ids are not degenerate)
    x = B.synth_device("mix", 1 << 20, dev, 0)
    np.random.seed(42)
    torch.manual_seed(42)
    cs, cur = [], x
    for k in B.CLUSTERS:
        km = KMeans(n_clusters=k, device=dev, balanced=True)
        km.cluster_centers = km.initialize(cur)
        for _ in range(2):
            km._iterate(cur, len(cur))
        cs.append(km.cluster_centers.clone())
        ids = engine.score_pass(cur, km.cluster_centers, argmin=True).argmin
        cur = engine.residual_normalise(cur, ids, km.cluster_centers, [B.DIM])
    del x, cur
    torch.cuda.empty_cache()
    chunk_rows = 8 << 20                                   # 16 GiB of fp32 rows per chunk
    out = {"what": "encode-only throughput (engine.encode mode 0 = rqk_encode_fused: one tcgen05 kernel over the three levels), [128,128,256], S-mix rows generated on the device",
           "n_gpus": world, "sizes": []}
    for total in (1000000, 10000000, 50000000, 100000000):
        mine = total // world + (1 if rank < total % world else 0)
        ms, done, c = 0.0, 0, 0
        while done < mine:
            m = min(chunk_rows, mine - done)
            xc = B.synth_device("mix", m, dev, rank * 1000 + c)
            if c == 0:
                engine.encode(xc[:min(m, 1 << 16)], cs, B.CLUSTERS, [B.DIM], mode=0)       # warm-up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            ids = engine.encode(xc, cs, B.CLUSTERS, [B.DIM], mode=0)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
            done += m
            c += 1
            del xc, ids
        ms = allmax(ms)
        flops = 3 * sum(2 * k * (B.DIM + 2) for k in B.CLUSTERS) * total
        out["sizes"].append({"vectors": total, "ms": ms, "vectors_per_s": total / (ms * 1e-3),
                             "tf32_tflops_issued": flops / (ms * 1e-3) / 1e12 / 1.0,
                             "hbm_gbs_algorithmic": total * (4 * B.DIM + 8 * 3) / (ms * 1e-3) / 1e9})
        if rank == 0:
            print(f"encode {total}: {ms:.1f} ms, {total / (ms * 1e-3) / 1e6:.1f} M vectors/s", flush=True)
    if rank == 0:
        print(json.dumps(out), flush=True)

elif what == "deep":
    rows = int(os.environ.get("DEEP_ROWS", 50000000))
    ks = [256, 256, 256, 256]
    t0 = time.time()
    x = B.synth_device("mix", rows, dev, 0, chunk=1 << 22)
    torch.cuda.synchronize()
    out = {"what": f"[256,256,256,256] on {rows} x 512 S-mix rows, one B200", "rows": rows, "generate_s": time.time() - t0,
           "levels": []}
    np.random.seed(42)
    torch.manual_seed(42)
    for lvl in range(2):
        k = ks[lvl]
        km = KMeans(n_clusters=k, device=dev, balanced=True)
        km.cluster_centers = km.initialize(x)
        sc, _, _, _ = km._iterate(x, rows)                 # warm-up iteration
        buf = sc.scores_t
        stats = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        iters = 2
        for _ in range(iters):
            sc, a, st, _ = km._iterate(x, rows, buf)
            sc.counts.cpu()
            stats.append(st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        sizes = torch.bincount(a.long(), minlength=k)
        r_exec = [s.passes - s.cold_passes for s in stats]
        by = sum(B.composite_bytes(rows, k, r) for r in r_exec)
        ach = by / (ms * 1e-3) / 1e9
        rec = {"level": lvl, "k": k, "ms_per_iteration": ms / iters, "vectors_per_s": rows * iters / (ms * 1e-3),
               "rounds_executed": r_exec, "reference_rounds": [s.rounds for s in stats], "passes": [s.passes for s in stats],
               "window_misses": [s.window_misses for s in stats], "list_rounds": [s.list_passes for s in stats],
               "balanced": bool((sizes[1:] == rows // k).all() and sizes[0] == rows // k + rows % k),
               "composite": {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}}
        out["levels"].append(rec)
        print(json.dumps(rec), flush=True)
        if lvl == 0:
            ids = engine.score_pass(x, km.cluster_centers, argmin=True).argmin
            engine.residual_normalise(x, ids, km.cluster_centers, [B.DIM], out=x)
        del buf, sc
    out["hbm_gb_in_use"] = torch.cuda.max_memory_allocated(dev) / 1e9
    out["peak_source"] = how
    print(json.dumps(out), flush=True)

if world > 1:
    dist.destroy_process_group()
