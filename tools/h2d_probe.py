"""Host -> device copy of a 2 GB pageable fp32 matrix: torch's staged copy, page-locked source, and the chunked
uploader of engine.h2d_rows with several chunk sizes / thread counts.  python tools/h2d_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from generative_ranking_recommender_b200 import engine

dev = torch.device("cuda:0")
n = int(os.environ.get("PROBE_ROWS", 1000000))
x = np.random.default_rng(0).standard_normal((n, 512), dtype=np.float32)
torch.cuda.synchronize()


def timed(label, fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        del out
    print(f"{label:50s} {min(ts) * 1e3:8.1f} ms  ({x.nbytes / min(ts) / 1e9:5.1f} GB/s)  all {[round(t * 1e3) for t in ts]}", flush=True)


timed("torch.from_numpy(x).to(dev)  [pageable]", lambda: torch.from_numpy(x).to(dev))
xp = torch.empty((n, 512), dtype=torch.float32, pin_memory=True)
t0 = time.perf_counter()
xp.copy_(torch.from_numpy(x))
print(f"host copy into pinned (1 thread): {(time.perf_counter() - t0) * 1e3:.1f} ms")
timed("pinned .to(dev, non_blocking)", lambda: xp.to(dev, non_blocking=True))
del xp
for chunk_mb, nbuf, threads in [(32, 4, 4), (8, 8, 4), (16, 6, 6), (64, 3, 3), (16, 4, 1), (16, 8, 8)]:
    u = engine._Uploader()
    u.CHUNK_BYTES, u.NBUF, u.THREADS = chunk_mb << 20, nbuf, threads
    t0 = time.perf_counter()
    u._setup()
    setup = time.perf_counter() - t0
    timed(f"uploader chunk {chunk_mb} MB x{nbuf}, {threads} threads (setup {setup * 1e3:.0f} ms)", lambda: u.upload(x, dev))
    u.pool.shutdown()
