"""Host side of the self-play output contract (matrix0_b200/records.py): the shard writer used when no DataManager is passed."""
import os

import numpy as np
import pytest


@pytest.mark.parametrize("level", [None, 1, 6])
def test_write_game_npz_roundtrip_and_atomic_name(tmp_path, level):
    from matrix0_b200.records import write_game_npz
    rs = np.random.RandomState(3)
    T = 7
    game = {"s": (rs.rand(T, 19, 8, 8) < 0.1).astype(np.float32), "pi": rs.rand(T, 4672).astype(np.float32), "z": np.ones((T,), np.float32),
            "legal_mask": (rs.rand(T, 4672) < 0.01).astype(np.uint8), "meta_moves": np.array([T], np.int32),
            "meta_result": np.array([0.25], np.float32), "meta_resigned": np.array([0], np.int8), "meta_draw": np.array([0], np.int8),
            "meta_avg_policy_entropy": np.array([1.5], np.float32), "meta_avg_sims": np.array([800.0], np.float32)}
    path = write_game_npz(str(tmp_path), game, worker_id=2, game_id=5, compresslevel=level)
    assert os.path.basename(path).startswith("selfplay_w2_g5_") and path.endswith(".npz")
    assert [f for f in os.listdir(tmp_path) if f.endswith(".tmp")] == []      # written under a temporary name, then renamed
    with np.load(path) as f:                                                 # the reference's readers use np.load (data_manager.py)
        assert set(f.files) == set(game)
        for k, v in game.items():
            assert f[k].dtype == v.dtype and np.array_equal(f[k], v), k
