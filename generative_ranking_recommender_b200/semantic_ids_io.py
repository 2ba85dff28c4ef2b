"""The caller side of the semantic-ID path: the data formats either
side of `HierarchicalRQKMeans.train`.  Same files, byte for byte, as the reference's driver
(`src/semantic_id_generator/train_semantic_ids.py`), written without its per-song Python work where that is
possible without changing a byte.

  load_song_vectors      train_semantic_ids.py:85-130   song_vectors.csv  -> (song_ids, [N, D] array)
  generate_semantic_ids  train_semantic_ids.py:209-237  train() result    -> {song_id: [id_1 .. id_L]}
  save_semantic_ids      train_semantic_ids.py:239-264  -> song_semantic_ids.jsonl
  semantic_id_statistics train_semantic_ids.py:292-331  per-layer usage statistics (training_statistics.json)
  collision_statistics   debug_collisions.py:27-61      id tuples shared by more than one song
"""
from __future__ import annotations

import csv
import json
import os
from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch


def load_song_vectors(vector_file: str, embedding_dim: int, max_samples: Optional[int] = None,
                      layer_clusters: Optional[Sequence[int]] = None) -> Tuple[List[str], torch.Tensor]:
    """train_semantic_ids.py:85-130.  Rows `song_id,v_1,...,v_D`; rows with fewer than two fields, a non-numeric
    vector or another dimension are skipped; `max_samples` counts ROWS READ (skipped ones included, :103).
    Returns float32 (fp16 if any layer has more than 512 clusters, :125-127)."""
    if not os.path.isfile(vector_file):
        raise FileNotFoundError(f"Song vector file not found: {vector_file}")
    song_ids: List[str] = []
    rows: List[np.ndarray] = []
    with open(vector_file, "r", encoding="utf-8") as f:
        for i, row in enumerate(csv.reader(f)):
            if max_samples and i >= max_samples:
                break
            if len(row) < 2:
                continue
            try:
                embed = np.array(row[1:], dtype=np.float32)
            except ValueError:
                continue
            if embed.shape[0] == embedding_dim:
                song_ids.append(row[0])
                rows.append(embed)
    if not song_ids:
        raise ValueError("No valid data with the correct embedding dimension found in the CSV file.")
    t = torch.from_numpy(np.vstack(rows))
    use_half = bool(layer_clusters) and any(n > 512 for n in layer_clusters)
    return song_ids, (t.half() if use_half else t.float())


def _id_matrix(train_result: Dict) -> np.ndarray:
    cols = []
    for c in train_result["cluster_ids"]:
        cols.append(c.detach().cpu().numpy() if isinstance(c, torch.Tensor) else np.asarray(c))
    return np.column_stack(cols).astype(np.int64)


def generate_semantic_ids(song_ids: Sequence, train_result: Dict) -> Dict:
    """train_semantic_ids.py:209-237: {song_id: [int]*L} in first-occurrence order; a repeated song_id keeps its
    first position and takes the ids of its LAST occurrence (dict assignment)."""
    m = _id_matrix(train_result).tolist()        # one C-level conversion instead of N*L .item() calls
    return {sid: row for sid, row in zip(song_ids, m)}


def save_semantic_ids(semantic_ids: Dict, output_file: str) -> int:
    """train_semantic_ids.py:239-264: one `{"song_id": ..., "semantic_ids": [...]}` object per line with
    json.dumps' default separators.  Returns the number of unique id tuples (:262).  The line is assembled from
    json.dumps(song_id) and the integers, which is what json.dumps(data) emits for this object."""
    d = os.path.dirname(output_file)
    if d:
        os.makedirs(d, exist_ok=True)
    unique = set()
    with open(output_file, "w", encoding="utf-8") as f:
        buf = []
        for song_id, ids in semantic_ids.items():
            buf.append('{"song_id": %s, "semantic_ids": [%s]}\n' % (json.dumps(song_id), ", ".join(map(str, ids))))
            unique.add(tuple(ids))
            if len(buf) >= 65536:
                f.write("".join(buf))
                buf.clear()
        f.write("".join(buf))
    return len(unique)


def semantic_id_statistics(semantic_ids: Dict, need_clusters: Sequence[int]) -> Dict:
    """train_semantic_ids.py:292-331."""
    rows = list(semantic_ids.values())
    stats = {"total_songs": len(semantic_ids), "unique_semantic_ids": len(set(tuple(r) for r in rows)),
             "layer_statistics": []}
    m = np.asarray(rows, dtype=np.int64).reshape(len(rows), -1)
    for layer in range(len(need_clusters)):
        ids = m[:, layer]
        counts = np.bincount(ids)
        stats["layer_statistics"].append({
            "layer": layer + 1, "unique_clusters": int(len(np.unique(ids))), "expected_clusters": need_clusters[layer],
            "min_cluster_id": int(ids.min()), "max_cluster_id": int(ids.max()),
            "cluster_distribution": {"min": int(counts.min()), "max": int(counts.max()),
                                     "mean": float(counts.mean()), "std": float(counts.std())}})
    return stats


def collision_statistics(source: Union[str, Dict, Iterable]) -> Dict:
    """debug_collisions.py:27-61.  `source`: a jsonl path (lines that do not parse or lack a key are skipped, like
    the script), a {song_id: ids} dict, or an iterable of (song_id, ids).  Collisions sorted worst first."""
    reverse = defaultdict(list)
    if isinstance(source, str):
        with open(source, "r", encoding="utf-8") as f:
            for line in f:
                try:
                    item = json.loads(line)
                    reverse[tuple(item["semantic_ids"])].append(item["song_id"])
                except (json.JSONDecodeError, KeyError):
                    continue
    else:
        it = source.items() if isinstance(source, dict) else source
        for song_id, ids in it:
            reverse[tuple(ids)].append(song_id)
    collisions = [(k, v) for k, v in reverse.items() if len(v) > 1]
    collisions.sort(key=lambda kv: len(kv[1]), reverse=True)
    return {"unique_semantic_ids": len(reverse), "colliding_ids": len(collisions),
            "songs_in_collision": sum(len(v) for _, v in collisions),
            "worst_collision": len(collisions[0][1]) if collisions else 0,
            "collisions": collisions}
