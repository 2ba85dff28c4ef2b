"""Fused encoder (csrc/encode_fused.cu) against the level chain on resident rows: time, re-evaluated rows, rows whose
ids differ.  python tools/encode_fused_probe.py [rows] [codebook, e.g. 128,128,256]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from generative_ranking_recommender_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
cl = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "128,128,256").split(",")]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(1234)
c = torch.randn((1024, 512), device=dev, generator=g)
j = torch.randint(0, 1024, (n,), device=dev, generator=g)
x = (c[j] + 0.5 * torch.randn((n, 512), device=dev, generator=g)) / 512 ** 0.5          # S-mix
# centres of a quick unbalanced Lloyd fit per level (what a trained model's look like)
cs, cur = [], x.clone()
for k in cl:
    ck = cur[torch.randperm(n, device=dev, generator=g)[:k]].clone()
    for _ in range(3):
        ids = engine.score_pass(cur, ck, argmin=True).argmin
        s_, cnt_ = engine.centroid_accumulate(cur, ids, k)
        engine.centroid_finalize(s_, cnt_, ck)
    cs.append(ck)
    ids = engine.score_pass(cur, ck, argmin=True).argmin
    engine.residual_normalise(cur, ids, ck, [512], out=cur)
del cur
out = {}
for mode in (0, 1):
    for fused in (True, False):
        for _ in range(2):
            ids = engine.encode(x, cs, cl, [512], mode=mode, fused=fused)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ids = engine.encode(x, cs, cl, [512], mode=mode, fused=fused)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5
        out[(mode, fused)] = ids
        extra = f", {engine.encode_reevaluated_rows(dev)} rows re-evaluated" if fused else ""
        print(f"{'fused' if fused else 'chain'} mode {mode}: {t:7.3f} ms for {n} rows = {n / t / 1e3:7.1f} M vectors/s{extra}",
              flush=True)
    d = (out[(mode, True)] != out[(mode, False)]).any(dim=0)
    print(f"   mode {mode}: {int(d.sum())} of {n} rows differ between fused and chain; per level "
          f"{[(out[(mode, True)][l] != out[(mode, False)][l]).sum().item() for l in range(len(cl))]}", flush=True)
