"""Quick look at the round-2 kernels on small inputs (seconds): single-kernel encoder vs level chain in both modes, and
an auction on clustered unit-norm rows at K = 256 / 128, where the windows get coarse bins and the merge kernel refines
from the survivor lists (HIST passes == rounds).  python tools/small_smoke.py [encode|auction]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from generative_ranking_recommender_b200 import engine

dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(3)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "encode"):
    n, dim, cl = 3000, 128, [32, 64, 96, 256]
    x = torch.randn((n, dim), device=dev, generator=g)
    cs = [torch.randn((k, dim), device=dev, generator=g) * (1.0 if l == 0 else 0.2) for l, k in enumerate(cl)]
    for mode in (0, 1):
        a = engine.encode(x, cs, cl, [dim], mode=mode, fused=True)
        b = engine.encode(x, cs, cl, [dim], mode=mode, fused=False)
        torch.cuda.synchronize()
        print("encode mode", mode, "rows that differ", int((a != b).any(dim=0).sum()), "re-evaluated", engine.encode_reevaluated_rows(dev))
if what in ("all", "auction"):
    # clustered unit-norm rows at K = 256: coarse windows -> the list-based refine of the merge kernel
    n, k = 26000, 256
    c = torch.randn((48, 256), device=dev, generator=g)
    x = c[torch.randint(0, 48, (n,), device=dev, generator=g)] + 0.3 * torch.randn((n, 256), device=dev, generator=g)
    x = x / x.norm(dim=1, keepdim=True)
    cen = x[torch.randperm(n, device=dev, generator=g)[:k]].contiguous()
    sc = engine.score_pass(x, cen, scores=True, argmin=True, counts=True)          # K = 256: smem operands
    a, st = engine.auction(sc.scores_t, n, sc.minmax)
    sizes = torch.bincount(a.long(), minlength=k)
    print("auction rounds", st.rounds, "passes", st.passes, "HIST", st.cold_passes, "sizes", int(sizes.min()), int(sizes.max()))
    sc2 = engine.score_pass(x, cen[:128].contiguous(), scores=True, argmin=True)   # K = 128: A operand from tensor memory
    a2, st2 = engine.auction(sc2.scores_t, n, sc2.minmax)
    print("auction K=128 passes", st2.passes)
torch.cuda.synchronize()
print("done")
