"""
oracle/rqk_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + the C auction in auction_oracle.c) of the reference's semantic-ID hot
path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module; the product package (generative_ranking_recommender_b200) never
does and fails loudly without its CUDA library instead.

Every function cites the reference lines it restates (paths relative to /root/reference/).

Parity pin: the reference has no tests or golden vectors of its own, so the pin is
tests/golden/*.npz, produced by oracle/gen_golden.py from the UNMODIFIED reference imported
in-process in the build container (script and fixtures committed; tests/test_oracle_golden.py
checks this module against them).  Where the reference itself is not reproducible by an
independent implementation (torch.topk order among equal fp16 values) the
fixtures are restricted to inputs where that freedom is provably not exercised
(`ambiguous_rounds == 0`) and the statistical protocol of SURVEY.md section 8c applies otherwise.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librqk_oracle.so")
_lib = None


class _AuctionInfo(ctypes.Structure):
    _fields_ = [
        ("rounds", ctypes.c_int64),
        ("ambiguous_rounds", ctypes.c_int64),
        ("fallback_used", ctypes.c_int64),
        ("eps_bits", ctypes.c_uint16),
        ("smax_bits", ctypes.c_uint16),
        ("smin_bits", ctypes.c_uint16),
    ]


def build(force: bool = False) -> str:
    """Compile auction_oracle.c with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "auction_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        L.rqk_oracle_auction_half.argtypes = [p, i64, i64, p, i64, ctypes.POINTER(_AuctionInfo)]
        L.rqk_oracle_auction_half.restype = ctypes.c_int
        L.rqk_oracle_auction_half_t.argtypes = [p, i64, i64, p, i64, ctypes.POINTER(_AuctionInfo)]
        L.rqk_oracle_auction_half_t.restype = ctypes.c_int
        L.rqk_oracle_auction_half_t_fast.argtypes = [p, i64, i64, p, i64, ctypes.POINTER(_AuctionInfo)]
        L.rqk_oracle_auction_half_t_fast.restype = ctypes.c_int
        for name in ("rqk_oracle_f2h", "rqk_oracle_h2f"):
            getattr(L, name).argtypes = [p, p, i64]
            getattr(L, name).restype = None
        for name in ("rqk_oracle_hsub", "rqk_oracle_hadd"):
            getattr(L, name).argtypes = [p, p, p, i64]
            getattr(L, name).restype = None
        L.rqk_oracle_num_threads.restype = ctypes.c_int
        L.rqk_oracle_set_threads.argtypes = [ctypes.c_int]
        L.rqk_oracle_eps.argtypes = [ctypes.c_uint16, ctypes.c_uint16]
        L.rqk_oracle_eps.restype = ctypes.c_uint16
        _lib = L
    return _lib


def num_threads() -> int:
    """Host threads the C auction uses (OpenMP over workers; results do not depend on it)."""
    return int(lib().rqk_oracle_num_threads())


def set_threads(n: int) -> None:
    lib().rqk_oracle_set_threads(int(n))


# --------------------------------------------------------------------------------------------
# distance: balancekmeans/__init__.py:576-603 (pairwise_distance_full) over torch.cdist (:596)
# --------------------------------------------------------------------------------------------

def cdist_mm(x: np.ndarray, c: np.ndarray) -> np.ndarray:
    """ATen `_euclidean_dist` (the path torch.cdist takes for p=2 when either side has > 25 rows):
    cat([-2x, |x|^2, 1]) @ cat([c, 1, |c|^2])^T, clamp_min(0), sqrt -- all in fp32.
    SURVEY.md E5: bit-identical to torch on this container's MKL; numpy's BLAS may differ in the
    last bits, so comparisons against it carry a tolerance."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    xn = (x * x).sum(-1, keepdims=True, dtype=np.float32)
    cn = (c * c).sum(-1, keepdims=True, dtype=np.float32)
    xa = np.concatenate([x * np.float32(-2.0), xn, np.ones_like(xn)], axis=1)
    ca = np.concatenate([c, np.ones_like(cn), cn], axis=1)
    r = xa @ ca.T
    np.maximum(r, np.float32(0.0), out=r)
    return np.sqrt(r, dtype=np.float32)


def cdist_direct(x: np.ndarray, c: np.ndarray) -> np.ndarray:
    """torch.cdist's non-mm path (both sides <= 25 rows): sqrt(sum((x-c)^2))."""
    x = np.asarray(x, dtype=np.float32)
    c = np.asarray(c, dtype=np.float32)
    d = x[:, None, :] - c[None, :, :]
    return np.sqrt((d * d).sum(-1, dtype=np.float32), dtype=np.float32)


def pairwise_distance_full(x: np.ndarray, c: np.ndarray, batch_size: int = 10000) -> np.ndarray:
    """balancekmeans/__init__.py:576-603: Euclidean (not squared) distance, row batches."""
    x = np.asarray(x, dtype=np.float32)
    c = np.asarray(c, dtype=np.float32)
    out = np.zeros((len(x), len(c)), dtype=np.float32)
    for i in range(0, len(x), batch_size):
        b = x[i:i + batch_size]
        if len(b) > 25 or len(c) > 25:
            out[i:i + batch_size] = cdist_mm(b, c)
        else:
            out[i:i + batch_size] = cdist_direct(b, c)
    return out


def distance_exact64(x: np.ndarray, c: np.ndarray) -> np.ndarray:
    """fp64 distance of the fp32 inputs: the yardstick for near-tie exclusion."""
    x = np.asarray(x, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    d2 = (x * x).sum(-1)[:, None] + (c * c).sum(-1)[None, :] - 2.0 * (x @ c.T)
    return np.sqrt(np.maximum(d2, 0.0))


def top2_relative_gap(d64: np.ndarray) -> np.ndarray:
    """(second smallest - smallest) / smallest per row, fp64; inf when the smallest is 0."""
    part = np.partition(d64, 1, axis=1)[:, :2]
    lo, hi = part[:, 0], part[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        g = (hi - lo) / lo
    g[lo == 0] = np.inf
    return g


# --------------------------------------------------------------------------------------------
# fp16 helpers (bit patterns as uint16)
# --------------------------------------------------------------------------------------------

def f2h_bits(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().rqk_oracle_f2h(x.ctypes.data, out.ctypes.data, x.size)
    return out


def h2f(bits: np.ndarray) -> np.ndarray:
    bits = np.ascontiguousarray(bits, dtype=np.uint16)
    out = np.empty(bits.shape, dtype=np.float32)
    lib().rqk_oracle_h2f(bits.ctypes.data, out.ctypes.data, bits.size)
    return out


def score_matrix_half_t(dist: np.ndarray) -> np.ndarray:
    """((-D).half()).T.contiguous() of __init__.py:29,40 as fp16 bit patterns [K, N]."""
    return np.ascontiguousarray(f2h_bits(-np.asarray(dist, dtype=np.float32)).T)


# --------------------------------------------------------------------------------------------
# balanced assignment: balancekmeans/__init__.py:12-140 (auction_lap_half)
# --------------------------------------------------------------------------------------------

@dataclass
class AuctionResult:
    assignment: np.ndarray   # int64 [N], worker (cluster) per job
    rounds: int              # number of topk evaluations
    ambiguous_rounds: int    # (round, worker) pairs where the canonical tie rule decided
    fallback_used: bool      # counter > 1000 dump onto worker 0 fired
    eps: float


def auction_lap_half(scores: np.ndarray, max_rounds: int = 0) -> AuctionResult:
    """scores = [N, K] fp32 job-by-worker scores (the reference passes -distance).  Canonical tie
    rule: lowest job index first (see auction_oracle.c header)."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n, k = scores.shape
    assign = np.empty(n, dtype=np.int64)
    info = _AuctionInfo()
    rc = lib().rqk_oracle_auction_half(scores.ctypes.data, n, k, assign.ctypes.data, max_rounds,
                                       ctypes.byref(info))
    if rc not in (0, -2):
        raise RuntimeError(f"oracle auction failed rc={rc}")
    eps = float(h2f(np.array([info.eps_bits], dtype=np.uint16))[0])
    return AuctionResult(assign, int(info.rounds), int(info.ambiguous_rounds),
                         bool(info.fallback_used), eps)


def auction_lap_half_t(s_bits_t: np.ndarray, max_rounds: int = 0, fast: bool = False) -> AuctionResult:
    """Same, on an already rounded and transposed [K, N] fp16 bit matrix (teacher-forced input for
    the GPU auction: identical fp16 in, identical int64 out).  fast=True: the bookkeeping-only variant
    `rqk_oracle_auction_half_t_fast` (no materialised bid matrix, hardware fp16 conversion), pinned bit for
    bit to the literal one by tests/test_oracle_golden.py; the BASELINE-size GPU tests use it."""
    s = np.ascontiguousarray(s_bits_t, dtype=np.uint16)
    k, n = s.shape
    assert n >= k
    assign = np.empty(n, dtype=np.int64)
    info = _AuctionInfo()
    fn = lib().rqk_oracle_auction_half_t_fast if fast else lib().rqk_oracle_auction_half_t
    rc = fn(s.ctypes.data, k, n, assign.ctypes.data, max_rounds, ctypes.byref(info))
    if rc not in (0, -2):
        raise RuntimeError(f"oracle auction failed rc={rc}")
    eps = float(h2f(np.array([info.eps_bits], dtype=np.uint16))[0])
    return AuctionResult(assign, int(info.rounds), int(info.ambiguous_rounds),
                         bool(info.fallback_used), eps)


# --------------------------------------------------------------------------------------------
# KMeans engine: balancekmeans/__init__.py:223-534
# --------------------------------------------------------------------------------------------

def initialize(x: np.ndarray, k: int) -> np.ndarray:
    """__init__.py:240-256 -- draws from NumPy's GLOBAL legacy RNG exactly like the reference, so
    `np.random.seed(s)` before the call lines the sequences up."""
    n = len(x)
    idx = np.random.choice(n, k, replace=(k > n))
    return np.array(x[idx], dtype=np.float32, copy=True)


def _default_randrow(n: int) -> int:
    # __init__.py:322 uses torch.randint on the global CPU generator; borrowed for the draw only.
    import torch
    return int(torch.randint(n, (1,)))


def update_centers(x: np.ndarray, assign: np.ndarray, centers: np.ndarray,
                   randrow: Callable[[int], int] = _default_randrow) -> np.ndarray:
    """__init__.py:314-324: mean of the members; an empty cluster takes one random data row."""
    k = len(centers)
    out = np.array(centers, dtype=np.float32, copy=True)
    for i in range(k):
        sel = np.nonzero(assign == i)[0]
        rows = x[sel]
        if rows.shape[0] == 0:
            rows = x[[randrow(len(x))]]
        out[i] = rows.mean(axis=0, dtype=np.float32)
    return out


def overflow_loss(counts: np.ndarray, target_nodes_num: int) -> int:
    """__init__.py:333-336."""
    c = np.asarray(counts, dtype=np.int64)
    return int(np.maximum(c - int(target_nodes_num), 0).sum())


def center_shift(new: np.ndarray, old: np.ndarray) -> float:
    """__init__.py:343-346: sum over clusters of the L2 norm of the move (fp32)."""
    d = np.asarray(new, np.float32) - np.asarray(old, np.float32)
    return float(np.sqrt((d * d).sum(1, dtype=np.float32), dtype=np.float32).sum(dtype=np.float32))


@dataclass
class FitTrace:
    iterations: int
    rounds: List[int]
    losses: List[int]
    shifts: List[float]
    best_iteration: int


def fit_by_min_loss(x: np.ndarray, k: int, target_nodes_num: int, iter_limit: int = 0,
                    tol: float = 1e-3, balanced: bool = True,
                    randrow: Callable[[int], int] = _default_randrow,
                    dist_fn=None) -> Tuple[np.ndarray, FitTrace]:
    """__init__.py:259-365 (euclidean, half=False)."""
    dist_fn = dist_fn or (lambda a, b: pairwise_distance_full(a, b, batch_size=100000))
    x = np.asarray(x, dtype=np.float32)
    centers = initialize(x, k)                                            # :295
    it = 0
    best, min_loss, best_it = None, float("inf"), -1
    tr = FitTrace(0, [], [], [], -1)
    while True:
        if it > 0 and it % 10 == 0:                                       # :305-306
            centers = initialize(x, k)
        d = dist_fn(x, centers)                                           # :308
        if balanced:
            res = auction_lap_half(-d)                                    # :310
            a = res.assignment
            tr.rounds.append(res.rounds)
        else:
            a = np.argmin(d, axis=1)                                      # :312
        prev = centers.copy()                                             # :314
        centers = update_centers(x, a, centers, randrow)                  # :315-324
        d2 = dist_fn(x, centers)                                          # :327
        cnt = np.bincount(np.argmin(d2, axis=1), minlength=k)             # :328-329
        loss = overflow_loss(cnt, target_nodes_num)                       # :333-336
        if loss <= min_loss:                                              # :338-341 (ties: later wins)
            min_loss, best, best_it = loss, centers.copy(), it
        shift = center_shift(centers, prev)                               # :343-346
        tr.losses.append(loss)
        tr.shifts.append(shift)
        it += 1
        if shift ** 2 < tol:                                              # :359
            break
        if iter_limit != 0 and it >= iter_limit:                          # :361
            break
    tr.iterations, tr.best_iteration = it, best_it
    return best, tr


def predict(x: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """KMeans.predict(balanced=False), __init__.py:489-534: argmin (first index on ties)."""
    d = pairwise_distance_full(x, centers, batch_size=100000)
    return np.argmin(d, axis=1).astype(np.int64)


# --------------------------------------------------------------------------------------------
# hierarchy: hierarchical_rq_kmeans.py
# --------------------------------------------------------------------------------------------

def apply_weights(x: np.ndarray, group_dims: Sequence[int], weights: Sequence[float]) -> np.ndarray:
    """hierarchical_rq_kmeans.py:583-604."""
    w = np.ones(x.shape[1], dtype=np.float32)
    cur = 0
    for g, wt in zip(group_dims, weights):
        w[cur:cur + g] = np.float32(wt)
        cur += g
    return (np.asarray(x, np.float32) * w[None, :]).astype(np.float32)


def residual_normalised(x: np.ndarray, ids: np.ndarray, centers: np.ndarray,
                        group_dims: Sequence[int]) -> np.ndarray:
    """hierarchical_rq_kmeans.py:1088-1128: r = x - C[id]; per dim-group r /= (|r|_2 + 1e-8)."""
    r = np.asarray(x, np.float32) - np.asarray(centers, np.float32)[np.asarray(ids, np.int64)]
    cur = 0
    for g in group_dims:
        blk = r[:, cur:cur + g]
        nrm = np.sqrt((blk * blk).sum(1, keepdims=True, dtype=np.float32), dtype=np.float32)
        blk /= (nrm + np.float32(1e-8))
        cur += g
    return r


def adaptive_iter_limit(num_samples: int, n_clusters: int, layer: int, base: int = 100,
                        is_sub_cluster: bool = False) -> int:
    """hierarchical_rq_kmeans.py:288-366."""
    spc = num_samples / max(n_clusters, 1)
    if is_sub_cluster:
        if num_samples < 5000:
            it = 15
        elif num_samples < 10000:
            it = 20
        elif num_samples < 20000:
            it = 25
        else:
            it = 30
        if spc < 50:
            it = max(10, int(it * 0.8))
        elif spc > 200:
            it = int(it * 1.2)
        return max(10, it)
    if num_samples < 5000:
        it = max(10, int(base * 0.2))
    elif num_samples < 10000:
        it = max(15, int(base * 0.3))
    elif num_samples < 50000:
        it = max(30, int(base * 0.5))
    elif num_samples < 100000:
        it = max(50, int(base * 0.7))
    elif num_samples < 500000:
        it = base
    elif num_samples < 1000000:
        it = int(base * 1.2)
    else:
        it = int(base * 1.5)
    if n_clusters > 512:
        it = int(it * 1.3)
    elif n_clusters > 256:
        it = int(it * 1.15)
    if layer > 1:
        it = max(10, int(it * 0.9))
    if spc < 50:
        it = int(it * 1.2)
    return max(10, it)


def train_direct(x: np.ndarray, clusters: Sequence[int], group_dims: Sequence[int],
                 weights: Sequence[Sequence[float]], iter_limit: int = 100):
    """HierarchicalRQKMeans.train with layer_clusters == need_clusters (every layer takes
    `_train_layer_0`): hierarchical_rq_kmeans.py:368-537 + :606-669."""
    cur = np.asarray(x, dtype=np.float32)
    ids_all, centers_all, traces = [], [], []
    for layer, k in enumerate(clusters):
        xw = apply_weights(cur, group_dims, weights[layer])                # :428
        target = 1
        for i, c in enumerate(clusters):                                   # :625-628
            if i != layer:
                target *= c
        iters = adaptive_iter_limit(len(xw), k, layer, iter_limit)        # :631
        centers, tr = fit_by_min_loss(xw, k, target, iters, balanced=True)  # :637-648
        ids = predict(xw, centers)                                         # :654
        res = residual_normalised(xw, ids, centers, group_dims)            # :660
        ids_all.append(ids)
        centers_all.append(centers)
        traces.append(tr)
        cur = res                                                          # :501-503
    return ids_all, centers_all, traces


def reassign_middle_layer(xw: np.ndarray, centers: np.ndarray, prev_ids: np.ndarray, pre_need: int, cur_need: int,
                          group_dims: Sequence[int]):
    """_reassign_clusters_middle_layer_with_residuals, hierarchical_rq_kmeans.py:839-904: argmin over ALL
    pre_need * cur_need centres with +10000 (fp32) outside the parent's block; the residual uses the raw id."""
    d = pairwise_distance_full(xw, centers, 100000)
    mask = np.zeros_like(d)
    for ci in range(pre_need):                                             # :876-881
        rows = prev_ids == ci
        if rows.any():
            mask[rows, ci * cur_need:(ci + 1) * cur_need] = 1.0
    d = d + np.float32(10000.0) * (np.float32(1.0) - mask)                 # :884
    raw = np.argmin(d, axis=1).astype(np.int64)                            # :886
    return raw, residual_normalised(xw, raw, centers, group_dims)          # :898


def train_middle_layer(xw: np.ndarray, prev_ids: np.ndarray, need_clusters: Sequence[int], layer: int,
                       group_dims: Sequence[int], iter_limit: int = 100):
    """_train_middle_layer, hierarchical_rq_kmeans.py:671-752: one balanced fit of need[layer] centres inside
    every cluster of the previous layer (sub-cluster iteration budget), then the block-masked reassignment.
    Returns (centres [pre*cur, D], ids in [0, cur), residual)."""
    cur_need, pre_need = need_clusters[layer], need_clusters[layer - 1]
    target = 1
    for idx, c in enumerate(need_clusters):                                # :699-703 (layers AFTER this one)
        if idx > layer:
            target *= c
    centers = []
    for i in range(pre_need):                                              # :709-733
        rows = np.where(prev_ids == i)[0]
        iters = adaptive_iter_limit(len(rows), cur_need, layer, iter_limit, True)
        c, _ = fit_by_min_loss(xw[rows], cur_need, target, iters, balanced=True)
        centers.append(c)
    centers = np.concatenate(centers, axis=0)                              # :736
    raw, res = reassign_middle_layer(xw, centers, prev_ids, pre_need, cur_need, group_dims)
    return centers, raw % cur_need, res                                    # :750


def encode_train_chain(x: np.ndarray, centers_list: Sequence[np.ndarray],
                       group_dims: Sequence[int], weights: Sequence[Sequence[float]]):
    """The ids `train()` emits for given centroids: per level KMeans.predict + normalised residual."""
    cur = np.asarray(x, np.float32)
    out = []
    for layer, c in enumerate(centers_list):
        xw = apply_weights(cur, group_dims, weights[layer])
        ids = predict(xw, c)
        out.append(ids)
        cur = residual_normalised(xw, ids, c, group_dims)
    return out


def encode_train_chain_gram(x: np.ndarray, centers_list: Sequence[np.ndarray]):
    """The identity behind csrc/encode_fused.cu, restated in fp64 (unit weights, one dim-group): the ids of
    encode_train_chain WITHOUT ever forming a residual.  With r_0 = x and r_{l+1} = (r_l - C_l[id_l]) / s_l,
    s_l = |r_l - C_l[id_l]| + 1e-8 (hierarchical_rq_kmeans.py:1111-1122), every score is a corrected dot product with
    the original row:  r_m . C_m[k] = (..((x . C_m[k] - G_0m[id_0][k]) / s_0 - G_1m[id_1][k]) / s_1 ..) / s_{m-1},
    G_lm = C_l C_m^T,  |r_{l+1}|^2 = (d_l / s_l)^2,  and d_l is the distance the level's argmin has just produced.
    Returns (ids per level, relative top-2 gap of d^2 per level) - tests compare with the literal chain."""
    x64 = np.asarray(x, np.float64)
    cs = [np.asarray(c, np.float64) for c in centers_list]
    n = len(x64)
    rows = np.arange(n)
    rn2 = (x64 * x64).sum(1)
    ids, gaps, inv_s = [], [], []
    for m, cm in enumerate(cs):
        dot = x64 @ cm.T
        for l in range(m):
            dot = (dot - (cs[l] @ cm.T)[ids[l]]) * inv_s[l][:, None]
        d2 = np.maximum(rn2[:, None] + (cm * cm).sum(1)[None, :] - 2.0 * dot, 0.0)
        i = np.argmin(d2, axis=1)
        part = np.partition(d2, 1, axis=1) if d2.shape[1] > 1 else np.concatenate([d2, d2 + 1], 1)
        gaps.append((part[:, 1] - part[:, 0]) / np.maximum(part[:, 1], 1e-300))
        d = np.sqrt(d2[rows, i])
        s = d + 1e-8
        ids.append(i.astype(np.int64))
        inv_s.append(1.0 / s)
        rn2 = (d / s) ** 2
    return ids, gaps


def predict_hierarchy(x: np.ndarray, centers_list: Sequence[np.ndarray], need_clusters: Sequence[int],
                      group_dims: Sequence[int], weights: Sequence[Sequence[float]],
                      match_matrices: Optional[list] = None) -> np.ndarray:
    """HierarchicalRQKMeans.predict, hierarchical_rq_kmeans.py:539-581 with _predict_layer_0
    (:1146-1173), _predict_middle_layer (:1175-1233, incl. the +10000 fp32 quirk) and
    _predict_last_layer (:1235-1305; match-matrix lookup index bug A13 kept).  Note :577: the
    residual handed to the next level is computed from the UNWEIGHTED current data."""
    cur = np.asarray(x, np.float32)
    L = len(centers_list)
    all_ids: List[np.ndarray] = []
    match_matrices = match_matrices or []
    for layer in range(L):
        c = np.asarray(centers_list[layer], np.float32)
        xw = apply_weights(cur, group_dims, weights[layer])
        d = pairwise_distance_full(xw, c)                                   # batch 10000 default
        if layer == 0:
            ids = np.argmin(d, axis=1)
        elif layer == L - 1:
            mm = match_matrices[layer - 1] if layer - 1 < len(match_matrices) else []
            before = all_ids[-2] * need_clusters[layer - 2] + all_ids[-1]   # :1256
            if mm:
                m = np.array(mm, dtype=np.float32)[before]
                d = d + np.float32(10000.0) * (np.float32(1.0) - m)
            ids = np.argmin(d, axis=1)
            if mm:
                remap = []
                for row in mm:
                    r, cnt = {}, 0
                    for col, v in enumerate(row):
                        if v == 1:
                            r[col] = cnt
                            cnt += 1
                    remap.append(r)
                ids = np.array([remap[b][i] for b, i in zip(before, ids)], dtype=np.int64)
        else:
            prev = all_ids[layer - 1]
            pre_need, cur_need = need_clusters[layer - 1], need_clusters[layer]
            mask = np.zeros_like(d)
            for ci in range(pre_need):                                      # :1211-1216
                rows = prev == ci
                if rows.any():
                    mask[rows, ci * cur_need:(ci + 1) * cur_need] = 1.0
            d = d + np.float32(10000.0) * (np.float32(1.0) - mask)          # :1219, fp32
            ids = np.argmin(d, axis=1) % cur_need                           # :1221,:1231
        ids = ids.astype(np.int64)
        all_ids.append(ids)
        if layer < L - 1:
            cur = residual_normalised(cur, ids, c, group_dims)              # :577 (unweighted cur)
    return np.column_stack(all_ids)


# --------------------------------------------------------------------------------------------
# last layer of the PROD-shaped config: two balanced fits + match matrix (hierarchical_rq_kmeans.py:754-1086)
# --------------------------------------------------------------------------------------------

def last_layer_match_row(sub_centers: np.ndarray, centers: np.ndarray, need: int,
                         randint: Callable[[int], int] = np.random.randint) -> np.ndarray:
    """_assign_last_match_matrix, :1017-1050, for one (l1, l2) group: `torch.cdist(sub_centers, centres)` (the mm
    path: the candidates are more than 25), every sub-centre takes its nearest unused candidate (first index on
    ties), random fill up to `need`.  uint8 [2K]."""
    k2 = centers.shape[0]
    d = pairwise_distance_full(sub_centers, centers, batch_size=1 << 30)
    row = np.zeros(k2, dtype=np.uint8)
    used: List[int] = []
    for j in range(min(len(sub_centers), need)):
        dr = d[j].copy()
        dr[used] = np.inf
        m = int(np.argmin(dr))
        row[m] = 1
        used.append(m)
    while len(set(used)) < need:
        m = randint(k2)
        while m in used:
            m = randint(k2)
        row[m] = 1
        used.append(m)
    return row


def last_layer_reassign(xw: np.ndarray, centers: np.ndarray, before: np.ndarray, match: np.ndarray) -> np.ndarray:
    """_reassign_clusters_last_layer_with_residuals, :906-966: argmin of d + 10000 * (1 - match[before]) in fp32;
    raw candidate index."""
    d = pairwise_distance_full(xw, centers)
    m = np.asarray(match, dtype=np.float32)[np.asarray(before, np.int64)]
    return np.argmin(d + np.float32(10000.0) * (np.float32(1.0) - m), axis=1).astype(np.int64)


def merge_match_ids(match: np.ndarray, raw: np.ndarray, before: np.ndarray) -> np.ndarray:
    """_merge_match_matrix_cluster_ids, :1054-1086: position of the raw id among the ones of the group's row."""
    mm = np.asarray(match, dtype=np.int64)
    pos = np.cumsum(mm == 1, axis=1) - 1
    assert (mm[before, raw] == 1).all(), "KeyError in the reference"
    return pos[before, raw]


# --------------------------------------------------------------------------------------------
# SimplifiedHierarchicalRQ: simplified_semantic_id_generator.py (second entry point, README.md:192-195)
# --------------------------------------------------------------------------------------------

def simplified_residual(x: np.ndarray, ids: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """:78-96 / :160-164: batch - centres[ids], NOT normalised."""
    return np.asarray(x, np.float32) - np.asarray(centers, np.float32)[np.asarray(ids, np.int64)]


def simplified_middle_predict(x: np.ndarray, centers: np.ndarray, prev_ids: np.ndarray, pre_need: int, n_need: int):
    """:139-174: distances to ALL pre_need * n_need centres, +inf outside the parent's block, argmin; the id
    reported is raw % n_need, the residual uses the raw id.  Returns (raw ids, residual)."""
    d = pairwise_distance_full(x, centers)
    mask = np.full_like(d, np.inf)
    for j in range(pre_need):                                              # :150-154
        rows = prev_ids == j
        if rows.any():
            mask[rows, j * n_need:(j + 1) * n_need] = 0
    raw = np.argmin(d + mask, axis=1).astype(np.int64)                     # :156-157
    return raw, simplified_residual(x, raw, centers)


def simplified_match_row(sub_centers: np.ndarray, candidates: np.ndarray, n_need: int,
                         randint: Callable[[int], int] = np.random.randint) -> np.ndarray:
    """:284-305 for one (l1, l2) group: every sub-centre takes its nearest not-yet-taken candidate (greedy, in
    sub-centre order); random fill up to n_need; uint8 [n_candidates] of 0 / 1."""
    n_cand = candidates.shape[0]
    dm = np.linalg.norm(sub_centers[:, np.newaxis, :] - candidates[np.newaxis, :, :], axis=2)
    sel = set()
    for k in range(len(sub_centers)):
        for ci in np.argsort(dm[k]):
            if ci not in sel:
                sel.add(ci)
                break
    while len(sel) < n_need:
        r = randint(n_cand)
        if r not in sel:
            sel.add(r)
    row = np.zeros(n_cand, dtype=np.uint8)
    row[list(sel)] = 1
    return row


def simplified_predict_with_matrix(x: np.ndarray, prev1: np.ndarray, prev2: np.ndarray, candidates: np.ndarray,
                                   match: np.ndarray, n_prev2: int) -> np.ndarray:
    """:311-331: dist[match[group] == 0] = inf; argmin (ids index the 2K candidates)."""
    d = pairwise_distance_full(x, candidates)
    g = np.asarray(prev1, np.int64) * n_prev2 + np.asarray(prev2, np.int64)
    d[np.asarray(match)[g] == 0] = np.inf
    return np.argmin(d, axis=1).astype(np.int64)


# --------------------------------------------------------------------------------------------
# collision statistics: debug_collisions.py:27-61, train_semantic_ids.py:300-303
# --------------------------------------------------------------------------------------------

def collision_stats(ids: np.ndarray) -> dict:
    """ids [N, L] -> #unique tuples, #tuples shared by >1 song, #songs involved, worst group."""
    ids = np.asarray(ids, dtype=np.int64)
    _, counts = np.unique(ids, axis=0, return_counts=True)
    coll = counts[counts > 1]
    return {
        "unique_ids": int(len(counts)),
        "colliding_ids": int(len(coll)),
        "songs_in_collision": int(coll.sum()),
        "max_collision": int(counts.max()) if len(counts) else 0,
    }


# --------------------------------------------------------------------------------------------
# synthetic inputs of SURVEY.md section 8d
# --------------------------------------------------------------------------------------------

def synth_iso(n: int, d: int = 512, seed: int = 1234) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((n, d), dtype=np.float32)


def synth_mix(n: int, d: int = 512, seed: int = 1234, modes: int = 1024) -> np.ndarray:
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((modes, d), dtype=np.float32)
    j = rng.integers(0, modes, n)
    x = (c[j] + np.float32(0.5) * rng.standard_normal((n, d), dtype=np.float32)) / np.float32(np.sqrt(d))
    return x.astype(np.float32)
