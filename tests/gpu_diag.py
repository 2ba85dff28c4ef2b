"""Bring-up diagnostics for the GPU box: each section runs in its own process under a timeout and
prints what it finds instead of stopping at the first mismatch.  `python tests/gpu_diag.py [section..]` (uses the oracle as its checker, hence it lives under tests/)"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SECTIONS = ["auction", "score_simt", "score_tc", "centroid", "residual", "encode", "fit", "big"]


def sec_auction():
    import numpy as np, torch
    from oracle import rqk_oracle as O
    from generative_ranking_recommender_b200 import engine
    from generative_ranking_recommender_b200.balancekmeans import _minmax_keys
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    cases = [(64, 4, 64), (130, 4, 64), (1000, 8, 64), (4096, 16, 64), (4100, 16, 64), (6000, 32, 128),
             (20000, 128, 128), (20010, 128, 128), (12800, 256, 64), (12900, 256, 64), (256, 256, 64), (300, 256, 64)]
    for n, k, dim in cases:
        x = O.synth_mix(n, dim, seed=n, modes=max(8, k))
        c = x[rng.choice(n, k, replace=False)]
        d = O.pairwise_distance_full(x, c, 100000)
        s = O.score_matrix_half_t(d)                      # [k, n] uint16
        t0 = time.time()
        ref = O.auction_lap_half_t(s)
        t_or = time.time() - t0
        ld = engine.pad_ld(n)
        st = torch.full((k, ld), float("-inf"), dtype=torch.float16)
        st[:, :n] = torch.from_numpy(s.view(np.float16))
        st = st.to(dev)
        mm = _minmax_keys(st[:, :n])
        t0 = time.time()
        a, stats = engine.auction(st, n, mm)
        torch.cuda.synchronize()
        t_gpu = time.time() - t0
        a = a.cpu().numpy().astype(np.int64)
        same = np.array_equal(a, ref.assignment)
        print(f"auction n={n} k={k}: match={same} diff={(a != ref.assignment).sum()} rounds gpu/oracle={stats.rounds}/{ref.rounds} "
              f"passes={stats.passes} cold={stats.cold_passes} miss={stats.window_misses} frozen={stats.frozen_exit} "
              f"amb={ref.ambiguous_rounds} eps={stats.eps}/{ref.eps} t_gpu={t_gpu*1e3:.1f}ms t_oracle={t_or:.2f}s", flush=True)


def _score_case(n, k, dim, simt, seed=0):
    import numpy as np, torch
    from oracle import rqk_oracle as O
    from generative_ranking_recommender_b200 import engine
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(seed)
    x = O.synth_mix(n, dim, seed=seed + 1, modes=max(8, k))
    c = x[rng.choice(n, k, replace=(k > n))].copy()
    c[k // 2:] += 0.01 * rng.standard_normal((k - k // 2, dim)).astype(np.float32)
    xd, cd = torch.from_numpy(x).to(dev), torch.from_numpy(c).to(dev)
    r = engine.score_pass(xd, cd, scores=True, argmin=True, best2=True, counts=True, dist=True, simt=simt)
    torch.cuda.synchronize()
    d = O.pairwise_distance_full(x, c, 100000)
    d64 = O.distance_exact64(x, c)
    s_ref = O.score_matrix_half_t(d)
    s = r.scores_t[:, :n].cpu().numpy().view(np.uint16)
    pad_ok = bool((r.scores_t[:, n:].cpu().numpy().view(np.uint16) == 0xFC00).all())
    frac = (s == s_ref).mean()
    # fp16 entries may differ by one ulp at rounding boundaries only
    sf = O.h2f(s); sr = O.h2f(s_ref)
    maxrel = np.max(np.abs(sf - sr) / np.maximum(np.abs(sr), 1e-3))
    dist = r.dist.cpu().numpy()
    scale = (x.astype(np.float64) ** 2).sum(1)[:, None] + (c.astype(np.float64) ** 2).sum(1)[None, :]
    err2 = np.max(np.abs(dist.astype(np.float64) ** 2 - d64 ** 2) / scale)
    am = r.argmin.cpu().numpy()
    ref_am = np.argmin(d, axis=1)
    gap = O.top2_relative_gap(d64)
    bad = (am != ref_am)
    bad_far = bad & (gap >= 1e-5)
    cnt = r.counts.cpu().numpy()
    mm = r.minmax.cpu().numpy()
    keys = np.where(s == 0x8000, 0, s).astype(np.int64)
    keys = np.where(keys & 0x8000, (~keys) & 0xFFFF, keys | 0x8000)
    b2 = r.best2.cpu().numpy()
    part = np.partition(dist, 1, axis=1)[:, :2] if k > 1 else np.concatenate([dist, np.full_like(dist, np.inf)], 1)
    print(f"score[{'simt' if simt else 'tc'}] n={n} k={k} dim={dim}: fp16 equal {frac:.6f} maxrel {maxrel:.2e} pad_ok={pad_ok} "
          f"d2 err {err2:.2e} argmin mismatches {bad.sum()} (outside near-ties {bad_far.sum()}) "
          f"counts_ok={np.array_equal(cnt, np.bincount(am, minlength=k))} minmax_ok={mm[0] == keys.max() and mm[1] == keys.min()} "
          f"best2_ok={np.allclose(b2, part, rtol=1e-6, atol=1e-7)}", flush=True)


def sec_score_simt():
    for n, k, dim in [(300, 5, 64), (5000, 16, 64), (4097, 128, 512), (3001, 256, 128)]:
        _score_case(n, k, dim, True)


def sec_score_tc():
    for n, k, dim in [(128, 16, 32), (300, 5, 64), (5000, 16, 64), (4097, 128, 512), (3001, 256, 128), (40000, 128, 512),
                      (20000, 256, 512)]:
        _score_case(n, k, dim, False)


def sec_centroid():
    import numpy as np, torch
    from oracle import rqk_oracle as O
    from generative_ranking_recommender_b200 import engine
    dev = torch.device("cuda:0")
    for n, k, dim in [(5000, 16, 64), (40001, 128, 512), (70000, 256, 512)]:
        x = O.synth_mix(n, dim, seed=3)
        rng = np.random.default_rng(1)
        a = rng.integers(0, k, n).astype(np.int32)
        a[a == 3] = 4                                    # cluster 3 empty
        c0 = x[:k].copy()
        xd = torch.from_numpy(x).to(dev)
        ad = torch.from_numpy(a).to(dev)
        sums, counts = engine.centroid_accumulate(xd, ad, k)
        sums2, _ = engine.centroid_accumulate(xd, ad, k)
        cd = torch.from_numpy(c0).to(dev)
        out, empty = engine.centroid_finalize(sums, counts, cd)
        torch.cuda.synchronize()
        ref = O.update_centers(x, a, c0, randrow=lambda m: 0)
        got = cd.cpu().numpy()
        nz = np.bincount(a, minlength=k) > 0
        err = np.max(np.abs(got[nz] - ref[nz]) / (np.abs(ref[nz]) + 1e-6))
        ref64 = np.stack([x[a == i].astype(np.float64).mean(0) if nz[i] else c0[i] for i in range(k)])
        err64 = np.max(np.abs(got[nz] - ref64[nz])) / np.max(np.abs(ref64))
        shift_ref = O.center_shift(np.where(nz[:, None], ref, c0), c0)
        print(f"centroid n={n} k={k}: counts_ok={np.array_equal(counts.cpu().numpy(), np.bincount(a, minlength=k))} "
              f"deterministic={torch.equal(sums, sums2)} max rel err vs numpy fp32 {err:.2e} vs fp64 {err64:.2e} "
              f"shift {out[0].item():.6f}/{shift_ref:.6f} n_empty={out[1].item()} empty_ok={empty.cpu().numpy()[3] == 1}", flush=True)


def sec_residual():
    import numpy as np, torch
    from oracle import rqk_oracle as O
    from generative_ranking_recommender_b200 import engine
    dev = torch.device("cuda:0")
    for n, k, dim, groups in [(3000, 16, 64, [64]), (3000, 16, 512, [512]), (1000, 8, 96, [32, 64]), (1000, 8, 512, [128, 384])]:
        x = O.synth_mix(n, dim, seed=5)
        c = x[:k].copy() * 0.9
        ids = np.random.default_rng(2).integers(0, k, n).astype(np.int32)
        ref = O.residual_normalised(x, ids, c, groups)
        got = engine.residual_normalise(torch.from_numpy(x).to(dev), torch.from_numpy(ids).to(dev),
                                        torch.from_numpy(c).to(dev), groups).cpu().numpy()
        print(f"residual n={n} dim={dim} groups={groups}: max abs err {np.max(np.abs(got - ref)):.2e}", flush=True)


def sec_encode():
    import numpy as np, torch
    from oracle import rqk_oracle as O
    from generative_ranking_recommender_b200 import engine
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(ROOT, "tests", "golden", "encode.npz"))
    x = O.synth_mix(int(g["n"]), int(g["dim"]), seed=int(g["seed"]), modes=int(g["modes"]))
    centers = [g["c0"], g["c1"], g["c2"]]
    cd = [torch.from_numpy(c).to(dev) for c in centers]
    xd = torch.from_numpy(x).to(dev)
    dim = int(g["dim"])
    for mode, key in ((0, "train_ids"), (1, "predict_ids")):
        ids = engine.encode(xd, cd, list(g["clusters"]), [dim], None, mode=mode).t().cpu().numpy()
        ok = np.cumprod(ids == g[key], axis=1).astype(bool)
        print(f"encode mode={mode}: cumulative agreement with reference {ok.mean(0)}", flush=True)


def sec_fit():
    import numpy as np, torch
    from oracle import rqk_oracle as O
    from generative_ranking_recommender_b200 import HierarchicalRQKMeans, HierarchicalRQKMeansConfig
    g = np.load(os.path.join(ROOT, "tests", "golden", "fit_stats.npz"))
    clusters = [int(c) for c in g["clusters"]]
    x = O.synth_mix(int(g["n"]), int(g["dim"]), seed=int(g["data_seed"]), modes=int(g["modes"]))
    dim = int(g["dim"])
    print("reference rows:", g["rows"][:, :5].tolist())
    for seed in (42, 43):
        np.random.seed(seed); torch.manual_seed(seed)
        cfg = HierarchicalRQKMeansConfig(layer_clusters=clusters, need_clusters=clusters, embedding_dim=dim,
                                         group_dims=[dim], hierarchical_weights=[[1.0]] * 3, iter_limit=int(g["iter_limit"]))
        m = HierarchicalRQKMeans(cfg, device=torch.device("cuda:0"))
        t0 = time.time()
        out = m.train(x, resume=False)
        dt = time.time() - t0
        ids = np.column_stack([t.numpy() for t in out["cluster_ids"]])
        st = O.collision_stats(ids)
        its = [len(s) for s in m.fit_stats]
        rounds = [sorted(set(i["rounds"] for i in s)) for s in m.fit_stats]
        print(f"fit seed={seed}: {st} iterations/level {its} rounds {rounds} time {dt:.2f}s", flush=True)


def sec_big():
    import numpy as np, torch
    from generative_ranking_recommender_b200 import engine
    from generative_ranking_recommender_b200.balancekmeans import KMeans
    dev = torch.device("cuda:0")
    for n, k in [(1000000, 128), (1048576, 128), (1000000, 256)]:
        g = torch.Generator(device=dev); g.manual_seed(1)
        x = torch.randn((n, 512), device=dev, generator=g)
        np.random.seed(0)
        km = KMeans(n_clusters=k, device=dev, balanced=True)
        c = km.initialize(x)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        sc = engine.score_pass(x, c, scores=True, argmin=True, counts=True)
        ev[1].record()
        a, stats = engine.auction(sc.scores_t, n, sc.minmax)
        ev[2].record()
        sums, counts = engine.centroid_accumulate(x, a, k)
        ev[3].record()
        torch.cuda.synchronize()
        sizes = torch.bincount(a.long(), minlength=k)
        print(f"big n={n} k={k}: score {ev[0].elapsed_time(ev[1]):.2f}ms auction {ev[1].elapsed_time(ev[2]):.2f}ms "
              f"(rounds {stats.rounds}, passes {stats.passes}, cold {stats.cold_passes}, miss {stats.window_misses}, frozen {stats.frozen_exit}) "
              f"centroid {ev[2].elapsed_time(ev[3]):.2f}ms sizes min/max {sizes.min().item()}/{sizes.max().item()} "
              f"size[0]={sizes[0].item()} jpw={n // k}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--run":
        globals()["sec_" + sys.argv[2]]()
        sys.exit(0)
    todo = sys.argv[1:] or SECTIONS
    for s in todo:
        print(f"===== {s} =====", flush=True)
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", s], timeout=int(os.environ.get("DIAG_TIMEOUT", "300")),
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(p.stdout[-6000:], flush=True)
            print(f"[{s}] exit {p.returncode} in {time.time() - t0:.1f}s", flush=True)
        except subprocess.TimeoutExpired as e:
            out = e.stdout if isinstance(e.stdout, str) else (e.stdout or b"").decode("utf-8", "replace")
            print(out[-4000:], flush=True)
            print(f"[{s}] TIMEOUT after {time.time() - t0:.1f}s", flush=True)
