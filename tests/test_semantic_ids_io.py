"""the data formats either side of train() — CSV in, jsonl out, usage and
collision statistics — against a fixture recorded from the reference's own driver (oracle/gen_golden.py io)."""
import json
import os

import numpy as np
import pytest
import torch

from generative_ranking_recommender_b200 import semantic_ids_io as IO

G = os.path.join(os.path.dirname(__file__), "golden", "io.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(G)


def test_csv_reader_matches_reference(g, tmp_path):
    p = tmp_path / "v.csv"
    p.write_bytes(g["csv"].tobytes())
    ids, emb = IO.load_song_vectors(str(p), 6, layer_clusters=[4, 4, 8])
    assert ids == json.loads(str(g["song_ids"])) and emb.dtype == torch.float32
    assert np.array_equal(emb.numpy(), g["emb"])
    ids7, _ = IO.load_song_vectors(str(p), 6, max_samples=7)          # the limit counts rows read, skipped ones too
    assert ids7 == json.loads(str(g["song_ids7"]))
    assert IO.load_song_vectors(str(p), 6, layer_clusters=[4, 1280])[1].dtype == torch.float16      # :125-127
    with pytest.raises(ValueError):
        IO.load_song_vectors(str(p), 4)
    with pytest.raises(FileNotFoundError):
        IO.load_song_vectors(str(tmp_path / "missing.csv"), 6)


def test_jsonl_bytes_and_statistics_match_reference(g, tmp_path):
    ids = json.loads(str(g["song_ids"]))
    result = {"cluster_ids": [torch.from_numpy(c) for c in g["cluster_ids"]]}
    sem = IO.generate_semantic_ids(ids, result)
    out = tmp_path / "o" / "song_semantic_ids.jsonl"
    unique = IO.save_semantic_ids(sem, str(out))
    assert out.read_bytes() == g["jsonl"].tobytes()                   # byte for byte, duplicate song id collapsed
    want = json.loads(str(g["stats"]))
    assert unique == want["unique_semantic_ids"]
    assert IO.semantic_id_statistics(sem, [4, 4, 8]) == want
    col = IO.collision_statistics(str(out))
    assert col["unique_semantic_ids"] == unique and col["colliding_ids"] >= 1
    assert col == IO.collision_statistics(sem)


def test_jsonl_line_is_json_dumps_for_awkward_ids(tmp_path):
    sem = {'a"b\\c': [1, 2, 3], "дом\n": [0, 0, 0], 17: [5, 6, 7], "": [1, 2, 3]}
    out = tmp_path / "x.jsonl"
    IO.save_semantic_ids(sem, str(out))
    want = "".join(json.dumps({"song_id": k, "semantic_ids": v}) + "\n" for k, v in sem.items())
    assert out.read_text(encoding="utf-8") == want


def test_collision_statistics_definition(tmp_path):
    p = tmp_path / "c.jsonl"
    p.write_text('{"song_id": "a", "semantic_ids": [1, 2]}\nnot json\n{"song_id": "b", "semantic_ids": [1, 2]}\n'
                 '{"song_id": "c"}\n{"song_id": "d", "semantic_ids": [3, 4]}\n{"song_id": "e", "semantic_ids": [1, 2]}\n'
                 '{"song_id": "f", "semantic_ids": [5, 5]}\n{"song_id": "g", "semantic_ids": [5, 5]}\n')
    c = IO.collision_statistics(str(p))
    assert (c["unique_semantic_ids"], c["colliding_ids"], c["songs_in_collision"], c["worst_collision"]) == (3, 2, 5, 3)
    assert c["collisions"][0] == ((1, 2), ["a", "b", "e"])
