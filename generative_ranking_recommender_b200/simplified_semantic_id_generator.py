"""Drop-in for the reference's second entry point, `src/semantic_id_generator/simplified_semantic_id_generator.py`
(citations are to that file): `SimplifiedHierarchicalRQ` with the same constructor, `train(data_path, data_limit)`,
`save_model / load_model / save_semantic_ids`, attributes and file formats, on the sm_100a engine.

What it does differently from `HierarchicalRQKMeans` (and why it is a separate class in the reference too):
  * residuals are NOT normalised                                   (:78-96, :160-164)
  * the middle layer masks foreign blocks with `inf`, not +10000   (:145-158)  -> exact argmin inside the parent's block
  * middle / last layers use `KMeans.fit` (no min-loss selection)  (:127-129, :208-212)
  * the last layer builds a "dynamic match matrix": for every (l1, l2) group the need[-1] candidates (of the
    2 * layer_clusters[-1] centres of two balanced fits) nearest to the group's own sub-centres, picked greedily
    without repetition, and predicts inside that set                (:247-331)

The data stays on the GPU from the first layer to the last; distances, balanced fits, residuals and the masked
argmin are library kernels.  The greedy matching itself is host logic in the reference as well (numpy on the CPU
whatever the device, :284-305, with host RNG draws in group order) and is kept as such, statement for statement,
so that a seeded run consumes the generators exactly like the reference.
"""
from __future__ import annotations

import json
import os.path as osp
import pickle
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import engine
from .balancekmeans import KMeans, pairwise_distance_full  # noqa: F401  (re-exported like the reference, :18)
from .hierarchical_rq_kmeans import HierarchicalRQKMeans, HierarchicalRQKMeansConfig
from .semantic_ids_io import load_song_vectors

__all__ = ["SimplifiedHierarchicalRQ"]


class SimplifiedHierarchicalRQ:
    """:22-37."""

    def __init__(self, config: HierarchicalRQKMeansConfig):
        self.config = config
        self.device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
        self.trained_kmeans_models: List[Optional[KMeans]] = []
        self.dynamic_match_matrix = None      # last layer
        self.final_layer_centers = None       # last layer
        self.middle_layer_centers = None      # recursive layers
        print(f"Initialized SimplifiedHierarchicalRQ on device: {self.device}")

    # ---- data ------------------------------------------------------------------------------------
    def _load_data(self, data_path: str, limit: int = None) -> Tuple[List[str], torch.Tensor]:
        """:39-76.  Same rows kept / skipped as the reference; fp16 when any layer has more than 512 clusters."""
        if not osp.isfile(data_path):
            raise FileNotFoundError(f"The specified data file was not found: {data_path}")
        return load_song_vectors(data_path, self.config.embedding_dim, limit, self.config.layer_clusters)

    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        if self.device.type != "cuda":
            raise engine._lib.RqkError(f"device {self.device}: SimplifiedHierarchicalRQ runs on CUDA sm_100a only "
                                       "(no CPU fallback)")
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    # ---- residuals ---------------------------------------------------------------------------------
    def _get_residuals(self, data: torch.Tensor, kmeans: KMeans) -> torch.Tensor:
        """:78-96: data - centres[predict(data)], not normalised.  Stays on the device."""
        x = self._dev(data)
        ids = engine.score_pass(x, self._dev(kmeans.cluster_centers), argmin=True).argmin
        return engine.residual_plain(x, ids, self._dev(kmeans.cluster_centers))

    # ---- middle layer --------------------------------------------------------------------------------
    def _train_middle_layer(self, data: torch.Tensor, prev_cluster_ids: torch.Tensor, layer_idx: int):
        """:98-174.  Returns (ids in [0, need), residual)."""
        n_clusters = self.config.layer_clusters[layer_idx]
        n_need = self.config.need_clusters[layer_idx]
        prev_n_need = self.config.need_clusters[layer_idx - 1]
        use_half = n_clusters > 512
        x = self._dev(data)
        prev = prev_cluster_ids.to(self.device).long()
        order = torch.argsort(prev, stable=True)
        counts = torch.bincount(prev, minlength=prev_n_need).cpu().tolist()
        all_sub_centers, start = [], 0
        for i in range(prev_n_need):                                               # :109-133
            rows = order[start:start + counts[i]]
            start += counts[i]
            if counts[i] == 0:
                all_sub_centers.append(torch.zeros(n_need, x.shape[1], device=self.device))   # :112-117
                continue
            sub_data = engine.gather_rows(x, rows)
            if counts[i] < n_need:                                                 # :124-126
                sample_indices = np.random.choice(counts[i], n_need, replace=True)
                sub_centers = sub_data[torch.from_numpy(sample_indices).to(self.device)]
            else:                                                                  # :127-130
                sub_kmeans = KMeans(n_clusters=n_need, device=self.device, balanced=True)
                sub_kmeans.fit(X=sub_data, iter_limit=self.config.iter_limit, half=use_half, tqdm_flag=False)
                sub_centers = sub_kmeans.cluster_centers
            all_sub_centers.append(sub_centers)
        combined_centers = torch.cat(all_sub_centers).to(self.device)              # :136-137
        self.middle_layer_centers = combined_centers
        # :145-164: `dist += inf outside the parent's block; argmin` is the exact argmin inside the block
        raw = HierarchicalRQKMeans._reassign_middle_layer(x, combined_centers, prev, prev_n_need, n_need)
        residuals = engine.residual_plain(x, raw, combined_centers)
        return (raw.long() % n_need), residuals                                    # :170-174

    # ---- training --------------------------------------------------------------------------------------
    def train(self, data_path: str, data_limit: int = None):
        """:176-245."""
        song_ids, embeddings = self._load_data(data_path, limit=data_limit)
        self._train_on(song_ids, embeddings)

    def _train_on(self, song_ids: List[str], embeddings: torch.Tensor):
        current_data = self._dev(embeddings)
        L = len(self.config.layer_clusters)
        ids_per_layer: List[torch.Tensor] = []
        previous_level_ids = None
        for layer_idx in range(L):
            print(f"--- Training Layer {layer_idx + 1}/{L} ---")
            n_clusters = self.config.layer_clusters[layer_idx]
            use_half = n_clusters > 512
            if layer_idx == 0:                                                     # :192-201
                kmeans = KMeans(n_clusters=n_clusters, device=self.device, balanced=True)
                target_nodes = np.prod(self.config.need_clusters[1:])
                kmeans.fit_by_min_loss(X=current_data, target_nodes_num=target_nodes, iter_limit=self.config.iter_limit,
                                       half=use_half)
                self.trained_kmeans_models.append(kmeans)
                cluster_ids = kmeans.predict(current_data)
            elif layer_idx < L - 1:                                                # :203-208
                cluster_ids, residuals = self._train_middle_layer(current_data, previous_level_ids, layer_idx)
                current_data = residuals
                self.trained_kmeans_models.append(None)
            else:                                                                  # :210-229
                kmeans_part1 = KMeans(n_clusters=n_clusters, device=self.device, balanced=True)
                kmeans_part1.fit(X=current_data, iter_limit=20, half=use_half)
                kmeans_part2 = KMeans(n_clusters=n_clusters, device=self.device, balanced=True)
                kmeans_part2.fit(X=current_data, iter_limit=20, half=use_half)
                candidate_centers = torch.cat([kmeans_part1.cluster_centers, kmeans_part2.cluster_centers], dim=0)
                self.final_layer_centers = candidate_centers
                self.trained_kmeans_models.append(None)
                prev_ids_l1, prev_ids_l2 = ids_per_layer[layer_idx - 2], ids_per_layer[layer_idx - 1]
                self.dynamic_match_matrix = self._get_dynamic_match_matrix(current_data, prev_ids_l1, prev_ids_l2,
                                                                           candidate_centers)
                cluster_ids = self._predict_with_dynamic_matrix(current_data, prev_ids_l1, prev_ids_l2,
                                                                candidate_centers, self.dynamic_match_matrix)
            cluster_ids = cluster_ids.long().cpu()
            ids_per_layer.append(cluster_ids)
            previous_level_ids = cluster_ids
            if layer_idx == 0:                                                     # :238-241
                current_data = self._get_residuals(current_data, self.trained_kmeans_models[0])
        # :231-236, vectorised: {song_id: [id per layer]} in first-occurrence order, last occurrence's ids
        m = torch.stack(ids_per_layer, dim=1).tolist()
        self.semantic_ids: Dict[str, List[int]] = {sid: row for sid, row in zip(song_ids, m)}
        self._ids_per_layer = ids_per_layer
        print("Training complete.")

    # ---- last layer ------------------------------------------------------------------------------------
    @staticmethod
    def _match_row(sub_centers: np.ndarray, candidates: np.ndarray, n_need: int) -> List[int]:
        """:284-305 for one (l1, l2) group, statement for statement (host numpy, host RNG for the fill)."""
        n_candidates = candidates.shape[0]
        dist_matrix = np.linalg.norm(sub_centers[:, np.newaxis, :] - candidates[np.newaxis, :, :], axis=2)
        selected_indices = set()
        row_match = [0] * n_candidates
        for k in range(len(sub_centers)):
            for candidate_idx in np.argsort(dist_matrix[k]):
                if candidate_idx not in selected_indices:
                    selected_indices.add(candidate_idx)
                    break
        while len(selected_indices) < n_need:
            rand_idx = np.random.randint(n_candidates)
            if rand_idx not in selected_indices:
                selected_indices.add(rand_idx)
        for idx in selected_indices:
            row_match[idx] = 1
        return row_match

    def _get_dynamic_match_matrix(self, data, prev_ids_l1, prev_ids_l2, candidate_centers) -> torch.Tensor:
        """:247-309.  float32 [need[-3] * need[-2], 2 * layer_clusters[-1]] of 0 / 1."""
        n_prev1, n_prev2, n_need = self.config.need_clusters[-3], self.config.need_clusters[-2], self.config.need_clusters[-1]
        n_candidates = candidate_centers.shape[0]
        x = self._dev(data)
        cand_np = candidate_centers.cpu().numpy()
        group = (prev_ids_l1.long() * n_prev2 + prev_ids_l2.long()).to(self.device)
        order = torch.argsort(group, stable=True)                      # rows of a group in ascending row order
        counts = torch.bincount(group, minlength=n_prev1 * n_prev2).cpu().tolist()
        match_matrix, start = [], 0
        for g in range(n_prev1 * n_prev2):                             # (i, j) in the reference's order: g = i * n_prev2 + j
            rows = order[start:start + counts[g]]
            start += counts[g]
            if counts[g] == 0:                                         # :267-268
                sub_centers = cand_np[np.random.choice(n_candidates, n_need, replace=False)]
            elif counts[g] <= n_need:                                  # :269-270
                sub_centers = engine.gather_rows(x, rows).cpu().numpy()
            else:                                                      # :271-276
                temp_kmeans = KMeans(n_clusters=n_need, device=self.device, balanced=True)
                temp_kmeans.fit(X=engine.gather_rows(x, rows), iter_limit=20, tqdm_flag=False)
                sub_centers = temp_kmeans.cluster_centers.cpu().numpy()
            match_matrix.append(self._match_row(sub_centers, cand_np, n_need))
        return torch.tensor(match_matrix, dtype=torch.float32)

    def _predict_with_dynamic_matrix(self, data, prev_ids_l1, prev_ids_l2, candidate_centers, match_matrix,
                                     batch_size: int = 1 << 20) -> torch.Tensor:
        """:311-331: nearest ALLOWED candidate of the row's (l1, l2) group; ids index the 2K candidates."""
        n_prev2 = self.config.need_clusters[-2]
        x = self._dev(data)
        cand = self._dev(candidate_centers)
        group = (prev_ids_l1.long() * n_prev2 + prev_ids_l2.long()).to(self.device)
        allow = (match_matrix != 0).to(torch.uint8)
        out = []
        for i in range(0, len(x), batch_size):      # the N x 2K fp32 matrix exists one batch at a time, as in the reference
            dist = engine.score_pass(x[i:i + batch_size], cand, argmin=False, dist=True).dist
            out.append(engine.masked_argmin(dist, group[i:i + batch_size], allow))
        return torch.cat(out).long()

    # ---- persistence -----------------------------------------------------------------------------------
    def save_model(self, path: str):
        """:333-343 (same pickle keys; tensors as they are)."""
        with open(path, "wb") as f:
            pickle.dump({
                "config": self.config,
                "trained_kmeans_models": [km.cluster_centers if km else None for km in self.trained_kmeans_models],
                "dynamic_match_matrix": self.dynamic_match_matrix,
                "final_layer_centers": self.final_layer_centers,
            }, f)

    @classmethod
    def load_model(cls, path: str):
        """:345-368."""
        with open(path, "rb") as f:
            checkpoint = pickle.load(f)
        model = cls(checkpoint["config"])
        model.trained_kmeans_models = []
        for centers in checkpoint["trained_kmeans_models"]:
            if centers is not None:
                model.trained_kmeans_models.append(KMeans(n_clusters=centers.shape[0], cluster_centers=centers,
                                                          device=model.device))
            else:
                model.trained_kmeans_models.append(None)
        model.dynamic_match_matrix = checkpoint["dynamic_match_matrix"]
        model.final_layer_centers = checkpoint["final_layer_centers"]
        return model

    def save_semantic_ids(self, output_file: str):
        """:370-387: one {"song_id": ..., "semantic_ids": [...]} object per line."""
        if not hasattr(self, "semantic_ids"):
            print("No semantic IDs generated yet. Run train() or predict() first.")
            return
        unique_ids = set()
        with open(output_file, "w", encoding="utf-8") as f:
            for song_id, ids in self.semantic_ids.items():
                f.write(json.dumps({"song_id": song_id, "semantic_ids": ids}) + "\n")
                unique_ids.add(tuple(ids))
        print(f"Saved {len(self.semantic_ids)} total IDs.")
        print(f"Found {len(unique_ids)} unique semantic IDs.")
