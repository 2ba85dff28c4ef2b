"""
oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by importing the
UNMODIFIED reference from /root/reference (read-only) in the build container and recording what it
computes.  The reference cannot travel to the GPU box, so the fixtures are committed together with
this script.  Nothing here is imported by the product.

    python oracle/gen_golden.py            # regenerates every fixture (a few minutes, CPU)

What is recorded (SURVEY.md section 8c protocol):
  auction.npz      tie-free small score matrices -> reference assignment + number of topk rounds
                   (N%K==0, N%K!=0 -> 1002-round fallback, and the N<K argmin quirk)
  eps.npz          (max, min) fp16 pairs -> reference eps (balancekmeans/__init__.py:33-34)
  distance.npz     X, C -> pairwise_distance_full (fp32) and its (-D).half() bit pattern
  stage.npz        one teacher-forced fit_by_min_loss iteration: centres_t -> D, assignment,
                   centres_{t+1}, argmin counts, loss, shift
  encode.npz       centroids + seeded X -> ids of the train()-chain and of predict() (+10000 quirk)
  fit_stats.npz    full train() on S-mix for several seeds -> collision statistics (statistical pin)
  iter_limit.npz   _calculate_adaptive_iter_limit table
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from src.semantic_id_generator import balancekmeans as ref_bk  # noqa: E402
from src.semantic_id_generator import hierarchical_rq_kmeans as ref_h  # noqa: E402
from src.common.utils import set_seed  # noqa: E402
from oracle import rqk_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

# The reference's predict() calls torch.cuda.mem_get_info unconditionally (:252); this container
# has no driver.  Stub that one query (it only sizes batches; rows are independent).
torch.cuda.mem_get_info = lambda *a, **k: (64 << 30, 80 << 30)


class TopkCounter:
    """Counts torch.Tensor.topk calls made by the reference auction (= rounds)."""

    def __enter__(self):
        self.n = 0
        self._orig = torch.Tensor.topk
        outer = self

        def counted(t, *a, **k):
            outer.n += 1
            return outer._orig(t, *a, **k)

        torch.Tensor.topk = counted
        return self

    def __exit__(self, *exc):
        torch.Tensor.topk = self._orig


def gen_auction():
    rng = np.random.default_rng(20261018)
    scores, assigns, rounds, shapes = [], [], [], []
    tries = 0
    want = {"div": 24, "nondiv": 10, "small": 6}
    got = {"div": 0, "nondiv": 0, "small": 0}
    while any(got[k] < want[k] for k in want) and tries < 3000:
        tries += 1
        K = int(rng.choice([2, 3, 4, 5, 8, 16]))
        kind = ["div", "nondiv", "small"][tries % 3]
        if got[kind] >= want[kind]:
            continue
        if kind == "div":
            N = K * int(rng.integers(1, 13))
        elif kind == "nondiv":
            N = K * int(rng.integers(1, 9)) + int(rng.integers(1, K))
        else:
            N = int(rng.integers(1, K))
        sc = (-rng.random((N, K)) * float(rng.choice([1.0, 10.0, 100.0]))).astype(np.float32)
        res = O.auction_lap_half(sc)
        if N >= K and res.ambiguous_rounds != 0:
            continue  # reference output depends on topk heap order: not a valid golden vector
        with TopkCounter() as tc:
            a = ref_bk.auction_lap_half(torch.from_numpy(sc)).numpy().astype(np.int64)
        scores.append(sc.ravel())
        assigns.append(a)
        rounds.append(tc.n)
        shapes.append((N, K))
        got[kind] += 1
    np.savez_compressed(
        os.path.join(OUT, "auction.npz"),
        scores=np.concatenate(scores), assign=np.concatenate(assigns),
        rounds=np.array(rounds, np.int64), shapes=np.array(shapes, np.int64))
    print("auction:", got, "cases", len(shapes), "tries", tries)


def gen_eps():
    rng = np.random.default_rng(7)
    hi = torch.from_numpy(-rng.random(4000).astype(np.float32) * 3).half()
    lo = torch.from_numpy((-rng.random(4000).astype(np.float32) * 200 - 3)).half()
    lo[:500] = hi[:500] - torch.from_numpy(rng.random(500).astype(np.float32) * 0.01).half()
    out = []
    for a, b in zip(hi, lo):
        eps = (a - b) / 50
        eps = max(eps, torch.tensor(1e-04, dtype=torch.float16))
        out.append(eps)
    out = torch.stack(out)
    np.savez_compressed(os.path.join(OUT, "eps.npz"),
                        smax=hi.view(torch.int16).numpy().view(np.uint16),
                        smin=lo.view(torch.int16).numpy().view(np.uint16),
                        eps=out.view(torch.int16).numpy().view(np.uint16))
    print("eps: 4000 pairs")


def gen_distance():
    x = O.synth_mix(300, 64, seed=5)
    c = x[np.random.default_rng(1).choice(300, 40, replace=False)].copy()
    d = ref_bk.pairwise_distance_full(torch.from_numpy(x), torch.from_numpy(c)).numpy()
    s = (-torch.from_numpy(d)).half().view(torch.int16).numpy().view(np.uint16)
    # small-by-small goes down cdist's direct (non-mm) path
    xs, cs = x[:20], c[:10]
    ds = ref_bk.pairwise_distance_full(torch.from_numpy(xs), torch.from_numpy(cs)).numpy()
    np.savez_compressed(os.path.join(OUT, "distance.npz"), x=x, c=c, d=d, s_bits=s, ds=ds)
    print("distance:", d.shape, ds.shape)


def gen_stage():
    """One teacher-forced iteration of fit_by_min_loss (balancekmeans/__init__.py:308-346)."""
    n, dim, k = 4096, 64, 16
    x = O.synth_mix(n, dim, seed=11, modes=64)
    xt = torch.from_numpy(x)
    set_seed(3)
    km = ref_bk.KMeans(n_clusters=k, device=torch.device("cpu"), balanced=True)
    c0 = km.initialize(xt).clone()
    d = ref_bk.pairwise_distance_full(xt, c0, batch_size=100000)
    with TopkCounter() as tc:
        a = ref_bk.auction_lap_half(-d)
    c1 = c0.clone()
    for i in range(k):
        sel = torch.nonzero(a == i).squeeze()
        rows = torch.index_select(xt, 0, sel)
        c1[i] = rows.mean(dim=0)
    d2 = ref_bk.pairwise_distance_full(xt, c1, batch_size=100000)
    arg = torch.argmin(d2, dim=1)
    cnt = torch.bincount(arg, minlength=k)
    shift = torch.sum(torch.sqrt(torch.sum((c1 - c0) ** 2, dim=1)))
    np.savez_compressed(
        os.path.join(OUT, "stage.npz"), n=n, dim=dim, k=k, seed=11, modes=64,
        c0=c0.numpy(), s_bits=(-d).half().view(torch.int16).numpy().view(np.uint16),
        assign=a.numpy().astype(np.int64), rounds=tc.n, c1=c1.numpy(),
        argmin=arg.numpy().astype(np.int64), counts=cnt.numpy().astype(np.int64),
        shift=float(shift))
    print("stage: rounds", tc.n, "sizes", np.bincount(a.numpy(), minlength=k))


def gen_encode():
    n, dim = 6000, 64
    clusters = [8, 8, 16]
    x = O.synth_mix(n, dim, seed=21, modes=128)
    cfg = ref_h.HierarchicalRQKMeansConfig(layer_clusters=clusters, need_clusters=clusters,
                                           embedding_dim=dim, group_dims=[dim],
                                           hierarchical_weights=[[1.0]] * 3, iter_limit=10)
    set_seed(42)
    m = ref_h.HierarchicalRQKMeans(cfg, checkpoint_dir=None, device=torch.device("cpu"))
    out = m.train(x, resume=False)
    train_ids = np.column_stack([t.cpu().numpy() for t in out["cluster_ids"]]).astype(np.int64)
    centers = [c.cpu().numpy() for c in out["cluster_centers"]]
    pred_ids = m.predict(x).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "encode.npz"), n=n, dim=dim, seed=21, modes=128,
                        clusters=np.array(clusters), c0=centers[0], c1=centers[1], c2=centers[2],
                        train_ids=train_ids, predict_ids=pred_ids)
    agree = (train_ids == pred_ids).mean(0)
    print("encode: train-vs-predict agreement per level", agree)


def gen_fit_stats():
    n, dim = 8192, 64
    clusters = [16, 16, 32]
    rows = []
    for seed in (42, 43, 44, 45, 46):
        x = O.synth_mix(n, dim, seed=1234, modes=256)
        cfg = ref_h.HierarchicalRQKMeansConfig(layer_clusters=clusters, need_clusters=clusters,
                                               embedding_dim=dim, group_dims=[dim],
                                               hierarchical_weights=[[1.0]] * 3, iter_limit=20)
        set_seed(seed)
        m = ref_h.HierarchicalRQKMeans(cfg, checkpoint_dir=None, device=torch.device("cpu"))
        out = m.train(x, resume=False)
        ids = np.column_stack([t.cpu().numpy() for t in out["cluster_ids"]])
        st = O.collision_stats(ids)
        per = [np.bincount(ids[:, l], minlength=clusters[l]) for l in range(3)]
        rows.append([seed, st["unique_ids"], st["colliding_ids"], st["songs_in_collision"],
                     st["max_collision"]] + [int(p.min()) for p in per] + [int(p.max()) for p in per])
        print("fit_stats seed", seed, rows[-1])
    np.savez_compressed(os.path.join(OUT, "fit_stats.npz"), n=n, dim=dim, data_seed=1234, modes=256,
                        clusters=np.array(clusters), iter_limit=20,
                        rows=np.array(rows, dtype=np.int64),
                        columns=np.array(["seed", "unique_ids", "colliding_ids", "songs_in_collision",
                                          "max_collision", "min0", "min1", "min2", "max0", "max1", "max2"]))


def gen_iter_limit():
    rows = []
    f = ref_h.HierarchicalRQKMeans._calculate_adaptive_iter_limit
    for n in (100, 4999, 5000, 9999, 20000, 49999, 60000, 100000, 499999, 500000, 999999, 1000000,
              10000000, 50000000):
        for k in (16, 128, 256, 257, 512, 513, 1280):
            for layer in (0, 1, 2, 3):
                for base in (20, 50, 100):
                    for sub in (False, True):
                        rows.append([n, k, layer, base, int(sub), f(n, k, layer, base, sub)])
    np.savez_compressed(os.path.join(OUT, "iter_limit.npz"), rows=np.array(rows, dtype=np.int64))
    print("iter_limit:", len(rows))


def gen_middle():
    """A hybrid config whose middle layer takes the recursive strategy (layer_clusters != need_clusters there, direct
    layers either side): the unmodified reference's train() and predict() on a small mixture."""
    n, d = 4096, 32
    x = O.synth_mix(n, d, seed=11, modes=64).astype(np.float32)
    cfg = ref_h.HierarchicalRQKMeansConfig(layer_clusters=[8, 64, 16], need_clusters=[8, 8, 16], embedding_dim=d,
                                           group_dims=[d], hierarchical_weights=[[1.0]] * 3, iter_limit=20)
    set_seed(42)
    m = ref_h.HierarchicalRQKMeans(cfg, device=torch.device("cpu"))
    out = m.train(x, resume=False)
    pred = m.predict(x)
    np.savez_compressed(os.path.join(OUT, "middle.npz"), x=x,
                        c0=out["cluster_centers"][0].numpy(), c1=out["cluster_centers"][1].numpy(),
                        c2=out["cluster_centers"][2].numpy(),
                        train_ids=np.stack([t.numpy() for t in out["cluster_ids"]]), predict_ids=pred)
    print("middle:", [tuple(c.shape) for c in out["cluster_centers"]], "unique ids",
          len({tuple(r) for r in np.stack([t.numpy() for t in out["cluster_ids"]]).T.tolist()}))


def gen_io():
    """The reference driver's own CSV reader, id assembly, jsonl writer and statistics on a small case with the
    awkward inputs (short rows, non-numeric fields, wrong dimension, quotes / non-ASCII / duplicate song ids)."""
    import json
    import tempfile
    from src.semantic_id_generator import train_semantic_ids as ref_t
    cls = ref_t.SemanticIDTrainer
    rng = np.random.default_rng(3)
    D, L = 6, 3
    names = ["s1", "song,with,commas", 'quo"te', "дом", "s1", "tab\tid", "7", "s8", "", "s10"]
    lines = []
    for i, nm in enumerate(names):
        vec = ["%.6g" % v for v in rng.standard_normal(D)]
        if i == 2:
            vec = vec[:-1]                      # wrong dimension
        if i == 5:
            vec[1] = "abc"                      # non-numeric
        row = [nm] + vec
        import csv as _csv, io as _io
        bio = _io.StringIO()
        _csv.writer(bio).writerow(row)
        lines.append(bio.getvalue())
    lines.insert(4, "lonely\r\n")              # fewer than two fields
    csv_text = "".join(lines)
    tmp = tempfile.mkdtemp()
    csv_path = os.path.join(tmp, "v.csv")
    open(csv_path, "w", encoding="utf-8", newline="").write(csv_text)
    fake = types.SimpleNamespace(rqkmeans_config=types.SimpleNamespace(embedding_dim=D, layer_clusters=[4, 4, 8],
                                                                       need_clusters=[4, 4, 8]))
    fake.config = types.SimpleNamespace(data=types.SimpleNamespace(song_vectors_file=csv_path))
    ids, emb = cls.load_song_vectors(fake, None)
    ids7, emb7 = cls.load_song_vectors(fake, 7)
    n = len(ids)
    cluster_ids = [torch.from_numpy(rng.integers(0, k, n).astype(np.int64)) for k in (4, 4, 8)]
    cluster_ids[2][1] = cluster_ids[2][0]; cluster_ids[1][1] = cluster_ids[1][0]; cluster_ids[0][1] = cluster_ids[0][0]
    sem = cls._generate_semantic_ids(fake, ids, {"cluster_ids": cluster_ids})
    out_path = os.path.join(tmp, "o", "song_semantic_ids.jsonl")
    fake.config = types.SimpleNamespace(data=types.SimpleNamespace(semantic_ids_file=out_path))
    cls._save_semantic_ids(fake, sem)
    stats = cls._generate_statistics(fake, sem)
    np.savez_compressed(os.path.join(OUT, "io.npz"), csv=np.frombuffer(csv_text.encode("utf-8"), dtype=np.uint8),
                        song_ids=np.array(json.dumps(ids)), emb=emb.numpy(), song_ids7=np.array(json.dumps(ids7)),
                        cluster_ids=torch.stack(cluster_ids).numpy(),
                        jsonl=np.frombuffer(open(out_path, "rb").read(), dtype=np.uint8),
                        stats=np.array(json.dumps(stats)))
    print("io:", n, "songs,", len(sem), "distinct ids written")


def gen_simplified():
    """A full run of the reference's second entry point, SimplifiedHierarchicalRQ.train (simplified_semantic_id_
    generator.py:176-245), on CPU: un-normalised residuals (:78-96), recursive middle layer with `inf` masks
    (:98-174), last layer = two balanced KMeans.fit + dynamic match matrix (:247-309) + masked prediction (:311-331).
    Recorded: every stage's centres and ids, and the temporary sub-centres each (l1, l2) group's match-matrix row
    was built from (KMeans.fit is chaotic: the tests teacher-force on these)."""
    import tempfile
    from src.semantic_id_generator import simplified_semantic_id_generator as ref_s
    n, dim = 3000, 32
    x = O.synth_mix(n, dim, seed=31, modes=48)
    lc, nc = [4, 8, 16], [4, 4, 8]
    cfg = ref_h.HierarchicalRQKMeansConfig(layer_clusters=lc, need_clusters=nc, embedding_dim=dim,
                                           group_dims=[dim], hierarchical_weights=[[1.0]] * 3, iter_limit=12)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "v.csv")
    with open(path, "w") as f:
        for i in range(n):
            f.write(f"s{i}," + ",".join(repr(float(v)) for v in x[i]) + "\n")
    fits = []
    orig_fit = ref_bk.KMeans.fit

    def recording_fit(self, *a, **k):
        r = orig_fit(self, *a, **k)
        fits.append(self.cluster_centers.detach().cpu().numpy().copy())
        return r

    ref_bk.KMeans.fit = recording_fit
    try:
        set_seed(42)
        m = ref_s.SimplifiedHierarchicalRQ(cfg)
        m.train(path)
    finally:
        ref_bk.KMeans.fit = orig_fit
    ids = np.array([m.semantic_ids[f"s{i}"] for i in range(n)], dtype=np.int64)
    # order of KMeans.fit calls: need[0] middle sub-fits, 2 candidate fits, then one per (l1, l2) group with > n_need rows
    sub = fits[nc[0] + 2:]
    groups = []
    for i in range(nc[0]):
        for j in range(nc[1]):
            cnt = int(((ids[:, 0] == i) & (ids[:, 1] == j)).sum())
            groups.append(cnt)
    big = [g for g, c in enumerate(groups) if c > nc[2]]
    assert len(sub) == len(big), (len(sub), len(big))
    assert min(groups) > 0, "pick a seed without empty (l1, l2) groups: their rows are host-RNG draws"
    xl = np.loadtxt(path, delimiter=",", usecols=range(1, dim + 1), dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "simplified.npz"), x=xl, ids=ids,
                        c0=m.trained_kmeans_models[0].cluster_centers.cpu().numpy(),
                        c_mid=m.middle_layer_centers.cpu().numpy(), c_last=m.final_layer_centers.cpu().numpy(),
                        match=m.dynamic_match_matrix.numpy().astype(np.uint8), group_sizes=np.array(groups),
                        sub_groups=np.array(big), sub_centers=np.stack(sub),
                        layer_clusters=np.array(lc), need_clusters=np.array(nc), iter_limit=12)
    print("simplified:", n, "rows,", len(big), "groups with a temporary fit,", len({tuple(r) for r in ids.tolist()}), "distinct ids")


def gen_last_layer():
    """A full HierarchicalRQKMeans.train + predict of the unmodified reference with the PROD config's SHAPE
    (direct first layer, recursive middle layer, last layer = two balanced KMeans.fit + match matrix,
    hierarchical_rq_kmeans.py:754-837, :906-1086, :1235-1305) at a size the CPU finishes in minutes.  Recorded: every
    layer's centres and ids, the match matrix, and the sub-centres (argument of the torch.cdist call, :1014) each
    (l1, l2) group's matrix row was built from - the sub-fits and the random subsets are host-RNG / chaotic, so the
    tests teacher-force on them."""
    n, dim = 3000, 32
    x = O.synth_mix(n, dim, seed=33, modes=48)
    lc, nc = [4, 16, 16], [4, 4, 8]
    cfg = ref_h.HierarchicalRQKMeansConfig(layer_clusters=lc, need_clusters=nc, embedding_dim=dim, group_dims=[dim],
                                           hierarchical_weights=[[1.0]] * 3, iter_limit=20)
    subs = []
    orig_cdist = torch.cdist

    def recording_cdist(a, b, *args, **kw):
        if sys._getframe(1).f_code.co_name == "_assign_last_match_matrix":      # not the sub-fits' own distance calls
            subs.append(a.detach().cpu().numpy().copy())
        return orig_cdist(a, b, *args, **kw)

    set_seed(42)
    m = ref_h.HierarchicalRQKMeans(cfg, device=torch.device("cpu"))
    orig_assign = m._assign_last_match_matrix

    def wrapped(*a, **k):
        torch.cdist = recording_cdist          # only the match-matrix builder's direct calls (KMeans uses F.cdist too)
        try:
            subs.clear()
            out = orig_assign(*a, **k)
            wrapped.subs = list(subs)
            return out
        finally:
            torch.cdist = orig_cdist

    m._assign_last_match_matrix = wrapped
    out = m.train(x, resume=False)
    ids = np.column_stack([t.cpu().numpy() for t in out["cluster_ids"]]).astype(np.int64)
    match = np.array(m.match_matrices[0], dtype=np.uint8)
    sizes = [int(((ids[:, 0] == i) & (ids[:, 1] == j)).sum()) for i in range(nc[0]) for j in range(nc[1])]
    nonempty = [g_ for g_, c in enumerate(sizes) if c > 0]
    sub = wrapped.subs
    assert len(sub) == len(nonempty), (len(sub), len(nonempty))
    assert all(s_.shape[0] == nc[2] for s_ in sub), "every group large enough for need[-1] sub-centres"
    pred = m.predict(x)
    np.savez_compressed(os.path.join(OUT, "last_layer.npz"), x=x, train_ids=ids, predict_ids=pred.astype(np.int64),
                        c0=out["cluster_centers"][0].cpu().numpy(), c_mid=out["cluster_centers"][1].cpu().numpy(),
                        c_last=out["cluster_centers"][2].cpu().numpy(), match=match, group_sizes=np.array(sizes),
                        sub_groups=np.array(nonempty), sub_centers=np.stack(sub), layer_clusters=np.array(lc),
                        need_clusters=np.array(nc), iter_limit=20)
    print("last_layer:", n, "rows,", len(nonempty), "non-empty groups,", len({tuple(r) for r in ids.tolist()}), "distinct ids,",
          "predict agrees with train on", float((pred == ids).all(1).mean()))


if __name__ == "__main__":
    import contextlib
    import io
    which = sys.argv[1:] or ["auction", "eps", "distance", "stage", "encode", "fit_stats", "iter_limit", "io", "middle", "simplified", "last_layer"]
    for w in which:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):   # the reference prints every iteration
            globals()["gen_" + w]()
        print(buf.getvalue().strip().splitlines()[-1] if buf.getvalue().strip() else w)
