"""Pure host-side arithmetic of the row-sharded fit.  No device code here, so the
world_size > 1 logic is testable with the gloo backend on CPU (tests/test_dist_gloo.py)."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch


def global_rows(sizes: Sequence[int], rank: int) -> Tuple[int, int]:
    """(total rows, first global row of `rank`) for contiguous row blocks in rank order."""
    sizes = [int(s) for s in sizes]
    return sum(sizes), sum(sizes[:rank])


def owned_rows(indices: np.ndarray, row0: int, n_local: int) -> Tuple[np.ndarray, np.ndarray]:
    """Which of the globally drawn centroid row `indices` live on this rank.
    Returns (positions in `indices`, local row numbers).  Duplicate indices (sampling with replacement
    when K > N, balancekmeans/__init__.py:250-251) are all kept."""
    indices = np.asarray(indices, dtype=np.int64)
    local = (indices >= row0) & (indices < row0 + n_local)
    return np.nonzero(local)[0].astype(np.int64), (indices[local] - row0).astype(np.int64)


def scatter_owned(out: torch.Tensor, positions: np.ndarray, rows: torch.Tensor) -> torch.Tensor:
    """out[positions] = rows; every other row stays zero, so a SUM all-reduce over ranks reconstructs
    X[indices] exactly (one non-zero contribution per row)."""
    if len(positions):
        out[torch.from_numpy(positions).to(out.device)] = rows
    return out


def rank_tie_offsets(totals: torch.Tensor, rank: int) -> torch.Tensor:
    """totals [world, k] = per-rank number of values equal to each worker's threshold.  Jobs are ordered
    rank-major, so the canonical lowest-job-index tie rule gives rank r an offset of sum_{r' < r}."""
    if rank == 0:
        return torch.zeros(totals.shape[1], dtype=torch.int32, device=totals.device)
    return totals[:rank].sum(dim=0, dtype=torch.int32)
