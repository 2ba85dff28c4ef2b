"""Helpers of the parity tests (test infrastructure; imports the oracle).

Near-tie rule (north-star): an integer code may differ from the reference's only for vectors whose reference top-2
distance gap is below 1e-5 relative; such vectors are EXCLUDED and COUNTED, and the count is printed and recorded
(pytest `record_property`), never hidden behind an agreement threshold."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from oracle import rqk_oracle as O

NEAR_TIE = 1e-5


def near_tie_rows(x: np.ndarray, c: np.ndarray, mask_block: Optional[np.ndarray] = None, block: int = 0) -> np.ndarray:
    """Rows of x whose argmin over the centres c could flip under a 1e-5 relative change of the distances (fp64).
    Plain levels: second-smallest distance within 1e-5 relative of the smallest.
    predict()'s masked levels (hierarchical_rq_kmeans.py:1210-1219: fl32(d + 10000) outside the parent's block
    [mask_block*block, (mask_block+1)*block)): the +10000 grid (ulp ~1e-3) decides, so a row is a near-tie if any
    other centre can reach the winner's rounded value when both distances move by 1e-5 relative."""
    d = O.distance_exact64(x, c)
    if mask_block is None:
        return O.top2_relative_gap(d) < NEAR_TIE
    k = d.shape[1]
    cols = np.arange(k)[None, :]
    inside = (cols >= mask_block[:, None] * block) & (cols < (mask_block[:, None] + 1) * block)
    off = np.where(inside, 0.0, 10000.0)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    q = f32(f32(d) + off)
    w = np.argmin(q, axis=1)
    rows = np.arange(len(d))
    worst_w = f32(f32(d[rows, w] * (1 + NEAR_TIE)) + off[rows, w])
    best_other = f32(f32(d * (1 - NEAR_TIE)) + off)
    lower = cols < w[:, None]
    can = np.where(lower, best_other <= worst_w[:, None], best_other < worst_w[:, None])
    can[rows, w] = False
    return can.any(axis=1)


def chain_mismatches(x: np.ndarray, centers: Sequence[np.ndarray], ids_got: np.ndarray, ids_ref: np.ndarray,
                     group_dims: Sequence[int], weights: Optional[Sequence[Sequence[float]]] = None,
                     predict_mode: bool = False, needs: Optional[Sequence[int]] = None) -> dict:
    """Compares multi-level ids [N, L] level by level, teacher-forced on the REFERENCE ids: a row is judged at the
    first level where it differs (a flip there excuses the later levels of that row) and is excused only if it is a
    near-tie at that level.  Returns {"excluded": [per level], "bad": [per level], "rows": N}."""
    n, L = ids_ref.shape
    weights = weights or [[1.0] * len(group_dims)] * L
    alive = np.ones(n, dtype=bool)
    cur = np.asarray(x, np.float32)
    excluded, bad = [], []
    for l in range(L):
        c = np.asarray(centers[l], np.float32)
        xw = O.apply_weights(cur, group_dims, weights[l])
        diff = alive & (ids_got[:, l] != ids_ref[:, l])
        rows = np.nonzero(diff)[0]
        if len(rows):
            if predict_mode and 0 < l < L - 1:
                near = near_tie_rows(xw[rows], c, ids_ref[rows, l - 1], int(needs[l]))
            else:
                near = near_tie_rows(xw[rows], c)
            excluded.append(int(near.sum()))
            bad.append(int((~near).sum()))
        else:
            excluded.append(0)
            bad.append(0)
        alive &= ~diff
        if l < L - 1:
            # train(): the residual lives in the weighted space (:428, :660); predict(): :577 takes it from the
            # unweighted data
            base = cur if predict_mode else xw
            cur = O.residual_normalised(base, ids_ref[:, l], c, group_dims)
    return {"excluded": excluded, "bad": bad, "rows": n}


def report(record_property, name: str, res: dict):
    line = f"{name}: {res['rows']} rows, near-ties excluded per level {res['excluded']}, mismatches outside {res['bad']}"
    print(line)
    if record_property is not None:
        record_property(name, line)


def count_mismatch_report(am_got: np.ndarray, d32: np.ndarray, x: np.ndarray, c: np.ndarray) -> dict:
    """Single level: GPU argmin vs the oracle's fp32 argmin, near-ties by the fp64 gap."""
    ref = np.argmin(d32, axis=1)
    rows = np.nonzero(am_got != ref)[0]
    near = near_tie_rows(x[rows], c) if len(rows) else np.zeros(0, bool)
    return {"rows": len(ref), "excluded": [int(near.sum())], "bad": [int((~near).sum())]}
