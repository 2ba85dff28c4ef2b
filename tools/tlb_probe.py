"""Does the time per row of the score pass / the HIST pass depend on the row pitch of S (ld * 2 bytes)?  With the
worker-major layout [K][ld] a 128-row tile touches K rows that are ld*2 bytes apart: K distinct 2 MB pages at N = 1 M."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generative_ranking_recommender_b200 import engine
dev = torch.device("cuda:0")
k = int(os.environ.get("PROF_K", 128))
for n in (62500, 125000, 250000, 500000, 1000000, 2000000):
    g = torch.Generator(device=dev); g.manual_seed(1)
    x = torch.randn((n, 512), device=dev, generator=g)
    c = x[torch.randperm(n, device=dev)[:k]].contiguous()
    for _ in range(2):
        sc = engine.score_pass(x, c, scores=True, argmin=True, counts=True)
    reps = max(2, 4000000 // n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sc = engine.score_pass(x, c, scores=True, argmin=True, counts=True, scores_out=sc.scores_t)
    e1.record(); torch.cuda.synchronize()
    t_score = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        engine.score_pass(x, c, scores=False, argmin=True, counts=True)
    e1.record(); torch.cuda.synchronize()
    t_arg = e0.elapsed_time(e1) / reps
    sess = engine.AuctionSession(sc.scores_t, n, n)
    sess.init(sc.minmax)
    ts = []
    for cyc in range(12):
        sess.do_pass(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sess.do_pass(2); b.record()
        sess.do_pass(4); sess.resolve()
        info = sess.poll()
        if cyc >= 3 and cyc % 2 == 0:
            ts.append(a.elapsed_time(b))
        if info.done: break
    ts = [t for t in ts if t > 0.005]
    th = sum(ts) / max(len(ts), 1)
    print(f"n={n:8d} K={k}: score+S {t_score*1e3:7.1f} us = {t_score*1e6/n:6.3f} ns/row | argmin only {t_arg*1e3:7.1f} us = {t_arg*1e6/n:6.3f} ns/row"
          f" | HIST {th*1e3:6.1f} us = {th*1e6/n:6.4f} ns/row ({2.0*k*n/th/1e6:6.0f} GB/s)", flush=True)
    del x, sc, sess
