"""Size sweep of the auction's two round kernels timed alone through the C-ABI step functions (as bench.py's roofline
leg does): the intercept of time against rows is the size-independent part of a launch (histogram clear, merge of the
per-CTA histograms, dump, last-CTA resolve).  python tools/hist_fixed_probe.py [K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from generative_ranking_recommender_b200 import engine

k = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(7)
res = []
for n in (32768 + 5, 131072 + 5, 524288 + 5, 1000000, 2000000 + 5, 4000000 + 5):
    x = torch.randn((n, 512), device=dev, generator=g)
    c = x[torch.randperm(n, device=dev, generator=g)[:k]].contiguous()
    sc = engine.score_pass(x, c, scores=True, argmin=False)
    sess = engine.AuctionSession(sc.scores_t, n, n)
    sess.init(sc.minmax)
    t_s, t_h, t_b = [], [], []
    prev = sess.poll()
    FLOW = 16 if os.environ.get("PROBE_FLOW", "driver") == "driver" else 0     # 16: the kernels rqk_auction chains
    t_m = []
    for cyc in range(60):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        sess.do_pass(1 | FLOW)
        e[1].record()
        sess.do_pass(2 | FLOW)
        e[2].record()
        if FLOW:
            sess.do_pass(8 | FLOW)
        e[3].record()
        sess.do_pass(4 | FLOW)
        e[4].record()
        if not FLOW:
            sess.resolve()
        cur = sess.poll()
        if cur.done:
            break
        if cyc >= 4:
            if cur.cold_passes > prev.cold_passes:
                t_h.append(e[1].elapsed_time(e[2]))
                t_s.append(e[0].elapsed_time(e[1]))
                t_m.append(e[2].elapsed_time(e[3]))
            if cur.passes - cur.cold_passes > prev.passes - prev.cold_passes:
                t_b.append(e[3].elapsed_time(e[4]))
        prev = cur
    m = lambda v: 1e3 * sum(v) / max(len(v), 1)
    res.append((n, m(t_h), m(t_b), m(t_s), m(t_m)))
    print(f"n={n:8d} K={k}: HIST {m(t_h):7.1f} us ({2.0 * k * n / max(m(t_h), 1e-9) / 1e3:6.0f} GB/s), bid-list {m(t_b):6.1f} us, "
          f"sample {m(t_s):5.1f} us, merge+resolve {m(t_m):5.1f} us   ({len(t_h)} / {len(t_b)} launches)", flush=True)
    del x, sc, sess
a = np.array(res)
for name, col in (("HIST", 1), ("bid-list", 2)):
    sl, ic = np.polyfit(a[2:, 0], a[2:, col], 1)
    print(f"{name}: {ic:.1f} us + {sl * 1e6:.1f} us per 1 M rows (fit over the four largest sizes)")
