"""Host-side mirror of the reference interface: config validation, checkpoint files, iteration budget,
error behaviour, and the rule that nothing computes without the CUDA library.  CPU only."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from generative_ranking_recommender_b200 import (CheckpointManager, HierarchicalRQKMeans, HierarchicalRQKMeansConfig,
                                                   hierarchicalRqClusterParams)
from generative_ranking_recommender_b200._lib import RqkError
from generative_ranking_recommender_b200.balancekmeans import KMeans, pairwise_cosine


def _cfg(**kw):
    base = dict(layer_clusters=[128, 128, 256], need_clusters=[128, 128, 256], embedding_dim=512)
    base.update(kw)
    return HierarchicalRQKMeansConfig(**base)


def test_config_normalisation_matches_reference():
    c = _cfg()
    assert c.group_dims == [512] and c.hierarchical_weights == [[1.0]] * 3 and c.iter_limit == 100
    c = _cfg(group_dims=512, hierarchical_weights=0.3)          # scalar weight -> uniform 1/len(groups)
    assert c.hierarchical_weights == [[1.0]] * 3
    c = _cfg(group_dims=[128, 384])
    assert c.hierarchical_weights == [[0.5, 0.5]] * 3


@pytest.mark.parametrize("kw,msg", [
    (dict(group_dims=[100, 100]), "Sum of group_dims 200 must equal embedding_dim 512"),
    (dict(hierarchical_weights=[[1.0]]), "Length of hierarchical_weights 1 must equal length of layer_clusters 3"),
    (dict(group_dims=[256, 256], hierarchical_weights=[[1.0]] * 3), "Length of hierarchical_weights[0] 1 must equal length of group_dims 2"),
])
def test_config_errors_verbatim(kw, msg):
    with pytest.raises(ValueError) as e:
        _cfg(**kw)
    assert msg in str(e.value)


def test_legacy_params_class():
    p = hierarchicalRqClusterParams()
    assert p.layer_clusters == [128, 256, 256] and p.need_clusters == [128, 128, 128] and p.group_dims == [1024]


def test_adaptive_iter_limit_table(golden_dir):
    rows = np.load(os.path.join(golden_dir, "iter_limit.npz"))["rows"]
    f = HierarchicalRQKMeans._calculate_adaptive_iter_limit
    for n, k, layer, base, sub, want in rows:
        assert f(int(n), int(k), int(layer), int(base), bool(sub)) == int(want)


def test_checkpoint_manager_files_and_resume_rule(tmp_path):
    cm = CheckpointManager(str(tmp_path / "ck"))
    assert cm.get_last_completed_layer() == -1
    ids = torch.arange(10)
    res = torch.randn(10, 4)
    cen = torch.randn(3, 4)
    cm.save_layer_checkpoint(0, ids, res, cen, None)
    cm.save_layer_checkpoint(2, ids, res, cen, None)            # stale higher layer after a gap is ignored (A10)
    assert cm.get_last_completed_layer() == 0
    assert not (tmp_path / "ck" / "layer_0_checkpoint.tmp").exists()
    raw = pickle.load(open(tmp_path / "ck" / "layer_0_checkpoint.pkl", "rb"))
    assert set(raw) == {"layer", "cluster_ids", "residual_data", "cluster_centers", "match_matrix"}
    assert isinstance(raw["cluster_ids"], np.ndarray) and raw["cluster_ids"].dtype == np.int64
    assert raw["residual_data"].dtype == np.float32 and raw["cluster_centers"].shape == (3, 4)
    back = cm.load_layer_checkpoint(0, torch.device("cpu"))
    assert torch.equal(back["cluster_ids"], ids) and torch.equal(back["residual_data"], res)
    with pytest.raises(ValueError):
        cm.save_layer_checkpoint(1, ids, res, None, None)       # validation: centres must not be None
    assert not (tmp_path / "ck" / "layer_1_checkpoint.pkl").exists()
    cm.save_metadata({"num_layers": 3, "x": np.float32(1.5)})
    assert cm.load_metadata()["num_layers"] == 3
    cm.clear_checkpoints()
    assert cm.get_last_completed_layer() == -1 and cm.load_metadata() is None


def test_checkpoint_files_are_per_rank_when_rows_are_sharded(tmp_path):
    """Rows sharded over ranks: every rank writes and resumes ITS row block (one shared file name would be a
    write/rename race and hand every rank the same block); rank 0 alone writes the metadata."""
    d = str(tmp_path / "ck")
    cms = [CheckpointManager(d, rank=r, world=2) for r in range(2)]
    blocks = [torch.randn(6 + r, 4) for r in range(2)]
    for r, cm in enumerate(cms):
        cm.save_layer_checkpoint(0, torch.arange(6 + r), blocks[r], torch.randn(3, 4), None)
        cm.save_metadata({"num_samples": 6 + r})
    assert sorted(os.listdir(d)) == ["checkpoint_metadata.json", "layer_0_rank0_checkpoint.pkl",
                                     "layer_0_rank1_checkpoint.pkl"]
    assert cms[0].load_metadata()["num_samples"] == 6
    for r, cm in enumerate(cms):
        assert cm.get_last_completed_layer() == 0
        assert torch.equal(cm.load_layer_checkpoint(0, torch.device("cpu"))["residual_data"], blocks[r])
    assert CheckpointManager(d).get_last_completed_layer() == -1     # an unsharded run does not pick them up
    cms[1].clear_checkpoints()
    assert cms[0].get_last_completed_layer() == 0 and cms[1].get_last_completed_layer() == -1


def test_resume_rejects_a_checkpoint_of_another_row_count(tmp_path):
    cfg = _cfg(embedding_dim=32, layer_clusters=[4, 4], need_clusters=[4, 4])
    cm = CheckpointManager(str(tmp_path / "ck"))
    cm.save_layer_checkpoint(0, torch.zeros(50, dtype=torch.int64), torch.zeros(50, 32), torch.zeros(4, 32), None)
    m = HierarchicalRQKMeans(cfg, checkpoint_dir=str(tmp_path / "ck"), device=torch.device("cpu"))
    with pytest.raises(RuntimeError, match="Checkpoint holds 50 rows but train\\(\\) was given 64"):
        m.train(np.zeros((64, 32), np.float32), resume=True)


def test_model_save_load_formats(tmp_path):
    m = HierarchicalRQKMeans(_cfg(), device=torch.device("cpu"))
    m.cluster_centers_list = [torch.randn(128, 512), torch.randn(128, 512), torch.randn(256, 512)]
    m.is_trained = True
    m.save_model(str(tmp_path / "model"))
    cfg = json.load(open(tmp_path / "model" / "config.json"))
    assert cfg["layer_clusters"] == [128, 128, 256] and cfg["group_dims"] == [512] and cfg["iter_limit"] == 100
    centres = pickle.load(open(tmp_path / "model" / "cluster_centers.pkl", "rb"))
    assert [c.shape for c in centres] == [(128, 512), (128, 512), (256, 512)] and centres[0].dtype == np.float32
    assert not (tmp_path / "model" / "match_matrices.pkl").exists()
    m2 = HierarchicalRQKMeans(_cfg(iter_limit=5), device=torch.device("cpu"))
    assert not m2.is_trained
    m2.load_model(str(tmp_path / "model"))
    assert m2.is_trained and m2.config.iter_limit == 100
    assert torch.equal(m2.cluster_centers_list[2], m.cluster_centers_list[2])
    assert m2.get_training_status() == {"is_trained": True, "last_completed_layer": -1, "total_layers": 3,
                                        "can_resume": False}


def test_errors_match_reference():
    m = HierarchicalRQKMeans(_cfg(), device=torch.device("cpu"))
    with pytest.raises(ValueError, match="Input dimension 64 does not match config embedding_dim 512"):
        m.train(np.zeros((10, 64), np.float32))
    with pytest.raises(RuntimeError, match="Model not trained"):
        m.predict(np.zeros((10, 512), np.float32))
    m.is_trained, m.cluster_centers_list = True, [torch.zeros(128, 512)]
    with pytest.raises(ValueError, match="Input dimension 64"):
        m.predict(np.zeros((10, 64), np.float32))
    with pytest.raises(NotImplementedError):
        KMeans(4, device=torch.device("cuda:0")).fit(torch.zeros(8, 32), distance="manhattan")
    with pytest.raises(NotImplementedError):
        pairwise_cosine(None, None)


def test_there_is_no_cpu_path():
    """A CPU device (or a box without CUDA) must raise, never fall back to torch or the oracle."""
    m = HierarchicalRQKMeans(_cfg(embedding_dim=32, layer_clusters=[4], need_clusters=[4]), device=torch.device("cpu"))
    with pytest.raises(RqkError, match="no CPU fallback"):
        m.train(np.zeros((64, 32), np.float32))
    with pytest.raises(RqkError, match="no CPU fallback"):
        KMeans(4, device=torch.device("cpu"), balanced=True).fit_by_min_loss(torch.zeros(64, 32), 16, iter_limit=1)
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            KMeans(4, device=torch.device("cuda:0"), balanced=True).predict(torch.zeros(8, 32))


def test_product_never_imports_the_oracle():
    import generative_ranking_recommender_b200 as pkg
    root = os.path.dirname(pkg.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "rqk_oracle" not in src, f


def test_speculative_draws_keep_the_numpy_stream_exact():
    """KMeans._draw hands out np.random.choice(N, K, replace=False) of the global generator and computes the next
    draw ahead on a thread; used or dropped, the stream every other consumer sees must be NumPy's own."""
    n = 250000                                   # above the threshold that turns speculation on
    np.random.seed(7)
    want = [np.random.choice(n, k, replace=False) for k in (8, 8, 16)]
    mid = np.random.randint(1 << 30)
    want.append(np.random.choice(n, 4, replace=False))
    after = np.random.randint(1 << 30)
    np.random.seed(7)
    km8, km16, km4 = (KMeans(n_clusters=k, device=torch.device("cpu")) for k in (8, 16, 4))
    got = [km8._draw(n), km8._draw(n), km16._draw(n)]          # 2nd and 3rd come from the speculative thread
    assert np.random.randint(1 << 30) == mid                    # ... which the generator state reflects exactly
    got.append(km4._draw(n))                                    # speculation invalidated by the randint: recomputed
    assert np.random.randint(1 << 30) == after
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n,k,seed", [(1, 1, 0), (10, 10, 1), (1000, 7, 2), (65537, 128, 3), (300000, 256, 4)])
def test_legacy_choice_is_numpy_global_choice(n, k, seed):
    """The library's host-side seed draw must be np.random.choice(n, k, replace=False) on the global legacy
    generator: same indices, same generator state afterwards (reference: balancekmeans/__init__.py:247-253)."""
    from generative_ranking_recommender_b200.balancekmeans import legacy_choice
    np.random.seed(seed)
    np.random.rand(seed + 3)                        # a state in the middle of a block of 624
    want1 = np.random.choice(n, k, replace=False)
    want2 = np.random.choice(n, k, replace=False)
    tail = np.random.rand(4)
    np.random.seed(seed)
    np.random.rand(seed + 3)
    got1 = legacy_choice(n, k)
    got2, st = legacy_choice(n, k, np.random.get_state())     # explicit-state form leaves the global alone
    assert np.array_equal(got1, want1) and np.array_equal(got2, want2)
    np.random.set_state(st)
    assert np.array_equal(np.random.rand(4), tail)


# ------------------------------------------------------------------------------------------------
# fit_by_min_loss loop control (reference balancekmeans/__init__.py:304-364), deterministic: the device work is
# scripted, so only the control flow is under test - the loss of iteration i is read one iteration later from the
# fused score pass, through an extra pass when a re-initialisation (:305-306) or the end of the loop intervenes
# ------------------------------------------------------------------------------------------------
def _reference_control(losses, shifts, tol, iter_limit):
    """:304-364 restated: (index of the iteration whose centroids are returned, number of iterations run)."""
    best, min_loss, it = None, float("inf"), 0
    while True:
        if losses[it] <= min_loss:                                         # :338 (ties: the later iteration wins)
            min_loss, best = losses[it], it
        it += 1
        if shifts[it - 1] ** 2 < tol:                                      # :359
            break
        if iter_limit != 0 and it >= iter_limit:                           # :361
            break
    return best, it


@pytest.mark.parametrize("case", [
    dict(losses=[9, 7, 7, 8, 5, 5, 6, 9, 9, 9, 4, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 3, 3, 9], limit=24),   # two re-inits
    dict(losses=[5, 4, 3, 2, 1, 1, 1, 1, 1, 1, 7, 7], limit=12),          # best sits right before a re-init
    dict(losses=[5, 5, 5, 5, 5, 5, 5, 5, 5, 5], limit=10),                # all equal: the last one wins; loop ends ON a boundary
    dict(losses=[3, 2, 9, 9, 9, 9], limit=0, stop_at=3),                  # tol stop, unlimited iterations
    dict(losses=[0, 4, 4], limit=3),                                      # zero loss first
    dict(losses=list(range(30, 0, -1)), limit=30),                        # monotone: the final extra pass decides
])
def test_fit_by_min_loss_loop_control_is_the_reference(case, monkeypatch):
    from generative_ranking_recommender_b200 import balancekmeans as bk
    from generative_ranking_recommender_b200 import engine

    losses, limit = case["losses"], case["limit"]
    shifts = [1.0] * len(losses)
    if "stop_at" in case:
        shifts[case["stop_at"]] = 1e-3                                     # shift^2 = 1e-6 < tol
    target = 100

    def counts_for(tag: float) -> torch.Tensor:
        # centroids produced by iteration t are tagged t >= 0; seed rows are tagged < 0 and would read as loss 0
        # ("always best") if the loop ever mistook the fused pass over fresh seeds for a loss evaluation
        loss = losses[int(tag)] if tag >= 0 else 0
        return torch.tensor([target + loss, 0], dtype=torch.int32)

    class Scripted(bk.KMeans):
        it, inits = 0, 0

        def _global_rows(self, n_local):
            return n_local, 0

        def _draw(self, n):
            return np.zeros(2, dtype=np.int64)

        def _rows(self, X, idx, n_global, row0):
            self.inits += 1
            return torch.full((2, 1), -float(self.inits))

        def _iterate(self, X, n_global, scores_buf=None):
            counts = counts_for(float(self.cluster_centers[0, 0]))          # fused pass: counts of the CURRENT centres
            self.cluster_centers = torch.full((2, 1), float(self.it))       # ... then the update
            self.it += 1
            return engine.ScoreResult(counts=counts), None, None, shifts[self.it - 1]

    monkeypatch.setattr(bk, "_cuda_device", lambda d: torch.device("cpu"))
    monkeypatch.setattr(bk.engine, "score_pass",
                        lambda X, c, **kw: engine.ScoreResult(counts=counts_for(float(c[0, 0]))))
    km = Scripted(n_clusters=2, device=torch.device("cpu"), balanced=True)
    km.trace_fit = True
    km.fit_by_min_loss(torch.zeros(8, 1), target_nodes_num=target, tol=1e-3, tqdm_flag=False, iter_limit=limit)
    want_best, want_iters = _reference_control(losses, shifts, 1e-3, limit)
    assert km.it == want_iters
    assert int(km.cluster_centers[0, 0]) == want_best and km.min_loss == losses[want_best]
    assert km.inits == 1 + (want_iters - 1) // 10                           # :295 + one re-init per 10 iterations (:305)
    assert [t["loss"] for t in km.last_fit_trace] == losses[:want_iters]    # every iteration's loss was evaluated, once
    assert [t["reinit"] for t in km.last_fit_trace] == [i > 0 and i % 10 == 0 for i in range(want_iters)]
