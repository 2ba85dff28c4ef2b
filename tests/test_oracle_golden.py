"""The oracle (oracle/rqk_oracle.py + auction_oracle.c) against fixtures recorded from the unmodified
reference by oracle/gen_golden.py.  CPU only."""
import os

import numpy as np
import pytest

from oracle import rqk_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_fp16_emulation_matches_numpy_half():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200000).astype(np.float32) * s
                        for s in (1e-8, 1e-5, 1e-3, 1.0, 100.0, 30000.0, 70000.0)])
    x = np.concatenate([x, np.array([0.0, -0.0, np.inf, -np.inf, 65504.0, 65519.9, 65520.0], np.float32)])
    assert np.array_equal(O.f2h_bits(x), x.astype(np.float16).view(np.uint16))
    bits = np.arange(65536, dtype=np.uint16)
    f = O.h2f(bits)
    ref = bits.view(np.float16).astype(np.float32)
    ok = (f == ref) | (np.isnan(f) & np.isnan(ref))
    assert ok.all()


def test_fp16_add_sub_match_numpy():
    rng = np.random.default_rng(1)
    a = (rng.standard_normal(300000) * 40).astype(np.float16)
    b = (rng.standard_normal(300000) * rng.choice([1e-3, 1.0, 300.0], 300000)).astype(np.float16)
    out = np.empty(a.shape, np.uint16)
    O.lib().rqk_oracle_hsub(a.view(np.uint16).ctypes.data, b.view(np.uint16).ctypes.data, out.ctypes.data, a.size)
    assert np.array_equal(out, (a - b).view(np.uint16))
    O.lib().rqk_oracle_hadd(a.view(np.uint16).ctypes.data, b.view(np.uint16).ctypes.data, out.ctypes.data, a.size)
    assert np.array_equal(out, (a + b).view(np.uint16))


def test_eps_matches_reference(golden_dir):
    g = _load(golden_dir, "eps.npz")
    got = np.array([O.lib().rqk_oracle_eps(int(a), int(b)) for a, b in zip(g["smax"], g["smin"])], np.uint16)
    assert np.array_equal(got, g["eps"])


def test_auction_matches_reference_on_tie_free_inputs(golden_dir):
    g = _load(golden_dir, "auction.npz")
    so = ao = 0
    n_nondiv = n_small = 0
    for (n, k), rounds in zip(g["shapes"], g["rounds"]):
        sc = g["scores"][so:so + n * k].reshape(n, k)
        ref = g["assign"][ao:ao + n]
        so += n * k
        ao += n
        res = O.auction_lap_half(sc)
        assert np.array_equal(res.assignment, ref), (n, k)
        if n >= k:
            assert res.ambiguous_rounds == 0
            assert res.rounds == rounds, (n, k, res.rounds, rounds)
            if n % k:
                n_nondiv += 1
                assert res.rounds == 1002 and res.fallback_used
        else:
            n_small += 1
            assert rounds == 0                                        # N<K quirk: no topk at all
    assert n_nondiv >= 5 and n_small >= 3


def test_fast_auction_variant_is_the_literal_one(golden_dir):
    """`rqk_oracle_auction_half_t_fast` (no materialised bid matrix, hardware fp16 conversion; what the
    BASELINE-size GPU tests use) against the literal restatement, bit for bit - assignment, rounds, the number
    of canonical-tie decisions, the fallback flag, eps - and against the reference's own golden vectors."""
    rng = np.random.default_rng(3)

    def both(s):
        a, b = O.auction_lap_half_t(s), O.auction_lap_half_t(s, fast=True)
        assert np.array_equal(a.assignment, b.assignment)
        assert (a.rounds, a.ambiguous_rounds, a.fallback_used, a.eps) == (b.rounds, b.ambiguous_rounds, b.fallback_used, b.eps)
        return a

    regimes = set()
    for n, k, dim in [(64, 4, 16), (130, 4, 16), (1000, 8, 32), (4100, 16, 32), (4096, 16, 32), (6001, 32, 32),
                      (2600, 256, 16), (300, 256, 16), (256, 256, 16)]:
        x = O.synth_mix(n, dim, seed=n, modes=max(8, k))
        c = x[rng.choice(n, k, replace=False)]
        r = both(O.score_matrix_half_t(O.pairwise_distance_full(x, c)))
        regimes.add((r.rounds == 1002, r.ambiguous_rounds > 0))
    assert regimes >= {(True, True), (False, True), (True, False)}
    for n, k in [(3000, 16), (3003, 16), (1024, 8)]:                     # few distinct values: ties decide everything
        both(-(rng.integers(0, 12, size=(k, n)) * 0.25).astype(np.float16).view(np.uint16))
    both(np.full((8, 1024), np.float16(-1.5)).view(np.uint16))
    g = _load(golden_dir, "auction.npz")
    so = ao = 0
    for (n, k), rounds in zip(g["shapes"], g["rounds"]):
        sc = g["scores"][so:so + n * k].reshape(n, k)
        ref = g["assign"][ao:ao + n]
        so += n * k
        ao += n
        if n >= k:
            res = O.auction_lap_half_t(O.score_matrix_half_t(-sc), fast=True)
            assert np.array_equal(res.assignment, ref) and res.rounds == rounds


def test_distance_matches_reference(golden_dir):
    g = _load(golden_dir, "distance.npz")
    d = O.pairwise_distance_full(g["x"], g["c"])
    # cancellation regime (x is itself a centre): the fp32 error lives in d^2, ~1e-7 * (|x|^2+|c|^2)
    scale = (g["x"] ** 2).sum(1)[:, None] + (g["c"] ** 2).sum(1)[None, :]
    assert (np.abs(d * d - g["d"] * g["d"]) <= 1e-6 * scale).all()
    s = O.score_matrix_half_t(d).T
    assert (s == g["s_bits"]).mean() >= 0.999
    ds = O.pairwise_distance_full(g["x"][:20], g["c"][:10])
    np.testing.assert_allclose(ds, g["ds"], rtol=1e-6, atol=1e-7)
    d64 = O.distance_exact64(g["x"], g["c"])
    assert (np.abs(d.astype(np.float64) ** 2 - d64 ** 2) <= 1e-6 * scale).all()


def test_stage_iteration_matches_reference(golden_dir):
    g = _load(golden_dir, "stage.npz")
    x = O.synth_mix(int(g["n"]), int(g["dim"]), seed=int(g["seed"]), modes=int(g["modes"]))
    c0 = g["c0"]
    d = O.pairwise_distance_full(x, c0, batch_size=100000)
    s = O.score_matrix_half_t(d)
    frac = (s.T == g["s_bits"]).mean()
    assert frac >= 0.9995, frac   # numpy BLAS vs MKL: last-bit differences flip a few fp16 roundings
    # teacher-forced: auction on the REFERENCE's own fp16 matrix.  Ties make element-wise equality
    # with torch.topk's heap order impossible, so compare what is invariant.
    res = O.auction_lap_half_t(np.ascontiguousarray(g["s_bits"].T))
    ref_a = g["assign"]
    k = int(g["k"])
    assert res.rounds == int(g["rounds"]) or abs(res.rounds - int(g["rounds"])) <= 3
    sz, ref_sz = np.bincount(res.assignment, minlength=k), np.bincount(ref_a, minlength=k)
    jpw = len(x) // k
    assert sz.min() >= jpw - 1 and ref_sz.min() >= jpw - 1
    dsum = d[np.arange(len(x)), res.assignment].sum(dtype=np.float64)
    ref_dsum = d[np.arange(len(x)), ref_a].sum(dtype=np.float64)
    assert abs(dsum - ref_dsum) / ref_dsum < 1e-3
    assert (res.assignment != ref_a).mean() < 0.05
    # centroid update, counts, shift: teacher-forced on the reference's assignment
    c1 = O.update_centers(x, ref_a, c0)
    np.testing.assert_allclose(c1, g["c1"], rtol=1e-5, atol=1e-7)
    arg = O.predict(x, g["c1"])
    assert (arg == g["argmin"]).mean() >= 0.9995
    assert abs(O.center_shift(g["c1"], c0) - float(g["shift"])) <= 1e-5 * float(g["shift"])


def test_encode_chain_and_predict_match_reference(golden_dir):
    g = _load(golden_dir, "encode.npz")
    x = O.synth_mix(int(g["n"]), int(g["dim"]), seed=int(g["seed"]), modes=int(g["modes"]))
    centers = [g["c0"], g["c1"], g["c2"]]
    dim = int(g["dim"])
    w = [[1.0]] * 3
    ids = np.column_stack(O.encode_train_chain(x, centers, [dim], w))
    # a flip at level l excuses the later levels of that vector
    ok = np.cumprod(ids == g["train_ids"], axis=1).astype(bool)
    assert ok[:, 0].mean() >= 0.9995 and ok[:, 2].mean() >= 0.998, ok.mean(0)
    pred = O.predict_hierarchy(x, centers, list(g["clusters"]), [dim], w)
    okp = np.cumprod(pred == g["predict_ids"], axis=1).astype(bool)
    assert okp[:, 0].mean() >= 0.9995 and okp[:, 2].mean() >= 0.995, okp.mean(0)
    # the +10000 quirk really is exercised: predict() differs from train() ids at level 2
    assert (g["predict_ids"][:, 1] != g["train_ids"][:, 1]).any()


def test_adaptive_iter_limit_table(golden_dir):
    rows = _load(golden_dir, "iter_limit.npz")["rows"]
    for n, k, layer, base, sub, want in rows:
        assert O.adaptive_iter_limit(int(n), int(k), int(layer), int(base), bool(sub)) == int(want)
    assert [O.adaptive_iter_limit(1000000, c, l, 20) for l, c in enumerate([128, 128, 256])] == [30, 30, 27]
    assert [O.adaptive_iter_limit(100000, c, l, 20) for l, c in enumerate([128, 128, 256])] == [20, 20, 18]


@pytest.mark.timeout(600)
def test_fit_statistics_overlap_reference(golden_dir):
    """End-to-end statistical parity: the canonical-tie oracle, run from the same
    seeds, must land in the reference's own spread."""
    import torch

    g = _load(golden_dir, "fit_stats.npz")
    rows = g["rows"]
    clusters = [int(c) for c in g["clusters"]]
    x = O.synth_mix(int(g["n"]), int(g["dim"]), seed=int(g["data_seed"]), modes=int(g["modes"]))
    dim = int(g["dim"])
    mine = []
    for seed in rows[:2, 0]:
        np.random.seed(int(seed))
        torch.manual_seed(int(seed))
        ids, _, _ = O.train_direct(x, clusters, [dim], [[1.0]] * 3, iter_limit=int(g["iter_limit"]))
        st = O.collision_stats(np.column_stack(ids))
        mine.append([st["unique_ids"], st["colliding_ids"], st["songs_in_collision"], st["max_collision"]])
    mine = np.array(mine, dtype=np.float64)
    ref = rows[:, 1:5].astype(np.float64)
    n = float(g["n"])
    # unique-id count within 1.5 % of N of the reference mean; collisions of the same order
    assert abs(mine[:, 0].mean() - ref[:, 0].mean()) < 0.015 * n, (mine, ref)
    assert 0.5 * ref[:, 1].min() <= mine[:, 1].mean() <= 2.0 * ref[:, 1].max(), (mine, ref)
    assert mine[:, 3].max() <= 2 * ref[:, 3].max() + 2


def test_middle_layer_restatement_against_reference_run(golden_dir):
    """Recursive middle layer (hierarchical_rq_kmeans.py:671-752, :839-904) and predict() over it (:1175-1233),
    teacher-forced on the centres and ids of an unmodified reference run (oracle/gen_golden.py middle)."""
    g = _load(golden_dir, "middle.npz")
    x, c0, c1, c2, tid = g["x"], g["c0"], g["c1"], g["c2"], g["train_ids"]
    assert c1.shape == (64, 32)
    res0 = O.residual_normalised(x, tid[0], c0, [32])
    assert np.array_equal(O.predict(x, c0), tid[0])
    raw, res1 = O.reassign_middle_layer(res0, c1, tid[0], 8, 8, [32])
    assert np.array_equal(raw // 8, tid[0]) and np.array_equal(raw % 8, tid[1])
    assert np.array_equal(O.predict(res1, c2), tid[2])
    pred = O.predict_hierarchy(x, [c0, c1, c2], [8, 8, 16], [32], [[1.0]] * 3)
    assert np.array_equal(pred, g["predict_ids"])
    # the reference's predict() differs from its own train() ids after a recursive layer (residual from the first
    # parent's block, :577 after :1231) - the restatement must reproduce that, not "fix" it
    assert (pred[:, 2] != tid[2]).mean() > 0.2


def test_simplified_generator_restatement_against_reference_run(golden_dir):
    """SimplifiedHierarchicalRQ (the reference's second entry point): given the centres of an unmodified
    reference run, the restated stages - un-normalised residuals, inf-masked middle layer, greedy dynamic match
    matrix, masked last-layer prediction - give the reference's ids and match matrix."""
    g = _load(golden_dir, "simplified.npz")
    x, ids, need = g["x"], g["ids"], [int(v) for v in g["need_clusters"]]
    ids0 = O.predict(x, g["c0"])
    assert np.array_equal(ids0, ids[:, 0])
    res1 = O.simplified_residual(x, ids0, g["c0"])
    raw, res2 = O.simplified_middle_predict(res1, g["c_mid"], ids0, need[0], need[1])
    assert np.array_equal(raw // need[1], ids0) and np.array_equal(raw % need[1], ids[:, 1])
    assert len(g["sub_groups"]) == need[0] * need[1]                  # every group went through a temporary fit
    for row, (grp, sub) in enumerate(zip(g["sub_groups"], g["sub_centers"])):
        assert np.array_equal(O.simplified_match_row(sub, g["c_last"], need[2]), g["match"][grp]), grp
    assert (g["match"].sum(1) == need[2]).all()
    last = O.simplified_predict_with_matrix(res2, ids[:, 0], ids[:, 1], g["c_last"], g["match"], need[1])
    assert np.array_equal(last, ids[:, 2])
    assert ids[:, 2].max() >= need[2]                                  # raw candidate indices, not remapped (:229)


def test_last_layer_match_matrix_restatement_against_reference_run(golden_dir):
    """The PROD config's shape (direct first layer, recursive middle layer, last layer = two balanced fits + match
    matrix) on the centres and sub-centres of an unmodified reference run: match matrix, train ids (positions inside
    the group's allowed set) and predict() ids (raw candidate indices: the lookup-index bug of :1248)."""
    g = _load(golden_dir, "last_layer.npz")
    x, tid, need, dim = g["x"], g["train_ids"], [int(v) for v in g["need_clusters"]], g["x"].shape[1]
    ids0 = O.predict(x, g["c0"])
    assert np.array_equal(ids0, tid[:, 0])
    res0 = O.residual_normalised(x, ids0, g["c0"], [dim])
    raw, res1 = O.reassign_middle_layer(res0, g["c_mid"], ids0, need[0], need[1], [dim])
    assert np.array_equal(raw % need[1], tid[:, 1])
    for grp, sub in zip(g["sub_groups"], g["sub_centers"]):
        assert np.array_equal(O.last_layer_match_row(sub, g["c_last"], need[2]), g["match"][grp]), grp
    before = tid[:, 0] * need[0] + tid[:, 1]                                   # :824
    raw_last = O.last_layer_reassign(res1, g["c_last"], before, g["match"])
    assert np.array_equal(O.merge_match_ids(g["match"], raw_last, before), tid[:, 2])
    pred = O.predict_hierarchy(x, [g["c0"], g["c_mid"], g["c_last"]], need, [dim], [[1.0]] * 3,
                               match_matrices=[g["match"].tolist()])
    assert np.array_equal(pred, g["predict_ids"])
    assert g["predict_ids"][:, 2].max() >= need[2]                            # raw candidate ids at predict time


def test_residual_free_encode_identity():
    """The algebra of the single-kernel encoder (csrc/encode_fused.cu), on the CPU: ids from corrected dot products
    with the ORIGINAL rows and Gram tables of the centres equal the literal chain (predict + normalised residual per
    level) - everywhere except rows whose top-2 gap is at rounding level (the chain works on fp32 residuals)."""
    n, dim, cl = 6000, 96, [16, 24, 32, 8]
    x = O.synth_mix(n, dim, seed=13, modes=80)
    rng = np.random.default_rng(2)
    cur, cs = x, []
    for k in cl:
        c = cur[rng.integers(0, n, (k, 3))].mean(axis=1).astype(np.float32)
        cs.append(c)
        cur = O.residual_normalised(cur, O.predict(cur, c), c, [dim])
    chain = O.encode_train_chain(x, cs, [dim], [[1.0]] * len(cl))
    ids, gaps = O.encode_train_chain_gram(x, cs)
    alive = np.ones(n, bool)
    for l in range(len(cl)):
        diff = alive & (ids[l] != chain[l])
        assert (gaps[l][diff] < 1e-5).all(), (l, int(diff.sum()))
        assert diff.sum() <= 3
        alive &= ~diff
    assert alive.sum() >= n - 6
