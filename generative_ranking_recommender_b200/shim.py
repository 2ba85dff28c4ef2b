"""Puts this package behind the reference's own import surfaces WITHOUT touching the reference tree
(INTEGRATION.md section 1): after `install()`, the reference's drivers

    src/semantic_id_generator/train_semantic_ids.py:30             from src.semantic_id_generator.hierarchical_rq_kmeans import ...
    config.py:9                                                    (same module)
    src/semantic_id_generator/simplified_semantic_id_generator.py:18   from src.semantic_id_generator.balancekmeans import ...

resolve those two module names to `generative_ranking_recommender_b200.{hierarchical_rq_kmeans,balancekmeans}`;
every other reference module (`config`, `src.common.utils`, the drivers themselves) is imported from the reference
tree as it is.  The drivers re-insert their project root at sys.path[0] when they are imported
(`train_semantic_ids.py:23-26`), so a shim DIRECTORY placed first on sys.path would lose; registering the two
modules in `sys.modules` cannot."""
from __future__ import annotations

import sys

_NAMES = ("src.semantic_id_generator.hierarchical_rq_kmeans", "src.semantic_id_generator.balancekmeans")


def install(reference_root: str | None = None) -> None:
    """Call before importing any reference driver.  `reference_root` (the checkout that holds `config.py` and
    `src/`) is appended to sys.path if it is not there yet."""
    from . import balancekmeans, hierarchical_rq_kmeans
    if reference_root and reference_root not in sys.path:
        sys.path.append(reference_root)
    sys.modules[_NAMES[0]] = hierarchical_rq_kmeans
    sys.modules[_NAMES[1]] = balancekmeans


def uninstall() -> None:
    for n in _NAMES:
        sys.modules.pop(n, None)
