"""Single-GPU reproduction of a 2-rank sharded auction (same data as tools/dist_check.py case 1):
which of {sharded protocol, unsharded driver} departs from the oracle, and do per-shard score passes equal
the full one bit for bit?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from generative_ranking_recommender_b200 import engine
from oracle import rqk_oracle as O

dev = torch.device("cuda", 0)
nloc, k, world = int(sys.argv[1]) if len(sys.argv) > 1 else 30011, int(sys.argv[2]) if len(sys.argv) > 2 else 64, 2

def make(r):
    g = torch.Generator(device=dev); g.manual_seed(100 + r)
    return torch.randn((nloc, 512), device=dev, generator=g)

xs = [make(r) for r in range(world)]
xa = torch.cat(xs)
n = nloc * world
np.random.seed(7)
idx = np.random.choice(n, k, replace=False)
c = xa[torch.from_numpy(idx).to(dev)].contiguous()
full = engine.score_pass(xa, c, scores=True, argmin=True, counts=True)
parts = [engine.score_pass(x, c, scores=True, argmin=True, counts=True) for x in xs]
S = full.scores_t[:, :n]
Sp = torch.cat([p.scores_t[:, :nloc] for p in parts], dim=1)
print("score shards == full:", torch.equal(S.view(torch.int16), Sp.view(torch.int16)),
      "differing entries:", int((S.view(torch.int16) != Sp.view(torch.int16)).sum()))
mmf = full.minmax
mmp = torch.stack([p.minmax for p in parts])
print("minmax full", mmf.tolist(), "parts", mmp.tolist())
a_un, st_un = engine.auction(full.scores_t, n, full.minmax)
ref = O.auction_lap_half_t(S.cpu().numpy().view(np.uint16))
print("unsharded == oracle:", np.array_equal(a_un.cpu().numpy().astype(np.int64), ref.assignment), st_un, ref.rounds)

mm = torch.stack([mmp[:, 0].max(), mmp[:, 1].min()]).to(torch.int32)
for sampled in (True, False):
    sess = []
    for r in range(world):
        sess.append(engine.AuctionSession(parts[r].scores_t, nloc, n)); sess[-1].init(mm)
    for it in range(3000):
        if sampled:
            allk = torch.cat([q.sample_collect(4096 // world) for q in sess], dim=1)
            for q in sess: q.sample_window(allk)
        for q in sess: q.do_pass(6)
        total = sum(q.reduce_block.clone() for q in sess)
        for q in sess:
            q.reduce_block.copy_(total); q.resolve()
        tt = torch.stack([q.tie_total.clone() for q in sess])
        for r, q in enumerate(sess):
            q.tie_offset(tt[:r].sum(0, dtype=torch.int32) if r else None)
        infos = [q.poll() for q in sess]
        if infos[0].done: break
    a = torch.cat([q.finalize() for q in sess]).cpu().numpy().astype(np.int64)
    bad = np.nonzero(a != ref.assignment)[0]
    print(f"sharded(sampled={sampled}) == oracle: {len(bad) == 0}; mismatches {len(bad)} first {bad[:8]}; passes {infos[0].passes} "
          f"misses {infos[0].window_misses} rounds {infos[0].rounds}")
