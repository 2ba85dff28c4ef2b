"""The reference's own driver (`train_semantic_ids.py`) running against this package through the import shim
(INTEGRATION.md section 1).  The reference tree exists only in the build container, so the test skips elsewhere; without a
GPU the run must end in the product's "no CPU fallback" error - raised from OUR train(), reached through the
reference's unmodified driver code - and never in a silent CPU fit."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "semantic_id_generator")), reason="reference tree not present")
def test_reference_driver_runs_against_the_shim(tmp_path, monkeypatch):
    import torch

    import generative_ranking_recommender_b200 as pkg
    from generative_ranking_recommender_b200 import shim
    from generative_ranking_recommender_b200._lib import RqkError

    monkeypatch.chdir(tmp_path)                                   # Config() creates its output directories in the cwd
    saved = {k: v for k, v in sys.modules.items() if k == "config" or k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    path0 = list(sys.path)
    try:
        shim.install(REF)
        import src.semantic_id_generator.train_semantic_ids as drv      # the reference's file, unmodified
        assert drv.__file__.startswith(REF)
        assert drv.HierarchicalRQKMeans is pkg.HierarchicalRQKMeans
        assert drv.HierarchicalRQKMeansConfig is pkg.HierarchicalRQKMeansConfig
        import config as ref_config
        assert ref_config.__file__.startswith(REF)
        assert isinstance(ref_config.H_RQ_KMEANS_TEST, pkg.HierarchicalRQKMeansConfig)
        import src.semantic_id_generator.simplified_semantic_id_generator as simp
        from generative_ranking_recommender_b200 import balancekmeans
        assert simp.KMeans is balancekmeans.KMeans and simp.pairwise_distance_full is balancekmeans.pairwise_distance_full

        cfg = ref_config.Config()
        cfg.data.song_vectors_file = str(tmp_path / "vectors.csv")
        cfg.data.semantic_ids_file = str(tmp_path / "out" / "song_semantic_ids.jsonl")
        rng = np.random.default_rng(0)
        with open(cfg.data.song_vectors_file, "w") as f:
            for i in range(200):
                f.write(f"s{i}," + ",".join(f"{v:.5f}" for v in rng.standard_normal(512)) + "\n")
        trainer = drv.SemanticIDTrainer(cfg, use_test_config=True)
        ids, vec = trainer.load_song_vectors(max_samples=100)
        assert len(ids) == 100 and tuple(vec.shape) == (100, 512)
        if not torch.cuda.is_available():
            with pytest.raises(RqkError, match="no CPU fallback"):
                trainer.train(resume=False)
    finally:
        shim.uninstall()
        for k in [k for k in sys.modules if k == "config" or k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = path0
