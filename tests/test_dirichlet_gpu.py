"""Root Dirichlet noise (MCTS._add_dirichlet, azchess/mcts.py:955-992): the caller-noise path against the oracle under the same
np.random stream, and the moments of the device's own Dirichlet sampler."""
import numpy as np
import pytest
import torch

import chess
from conftest import random_playout_boards
from oracle.backends import HashBackend
from oracle.mcts_ref import RefConfig, RefMCTS

pytestmark = pytest.mark.gpu

KW = dict(cpuct=2.5, cpuct_start=3.0, cpuct_end=2.0, cpuct_plies=40, fpu_reduction=0.1, draw_penalty=-0.05, legal_softmax=True,
          selection_jitter=0.05, inference_batch_size=32, no_instant_backtrack=True, enable_entropy_noise=False,
          dirichlet_alpha=0.3, dirichlet_frac=0.25, dirichlet_plies=30)


def test_dirichlet_with_reference_rng_stream_matches_oracle():
    """np.random.seed(s) before run(): the drop-in draws np.random.dirichlet([alpha] * k) exactly like the reference (same global stream);
    mixing into the priors happens on the device.  Jitter is pinned to random.random() == 0.5 on both sides; visit counts, pi, value bit-exact.
    ply >= dirichlet_plies switches the noise off (mcts.py:373-376)."""
    from matrix0_b200.mcts import MCTS, MCTSConfig
    boards = random_playout_boards(6, 100, seed=77)[::7][:12]
    for i, b in enumerate(boards):
        for ply in (3, 30):
            sims = 200
            c = MCTSConfig(num_threads=1, enable_memory_cleanup=False, playout_random_frac=0.0, num_simulations=sims, **KW)
            m = MCTS(c, None, device="cuda", inference_backend=HashBackend(1.0, seed=40 + i), deterministic=False, max_nodes=32768)
            m.set_random_streams(jitter=np.full(200_000, 0.5))
            np.random.seed(500 + i)
            vc1, pi1, v1 = m.run(b, ply=ply)
            ref = RefMCTS(RefConfig(num_simulations=sims, playout_random_frac=0.0, **KW), HashBackend(1.0, seed=40 + i), jitter_value=0.5)
            np.random.seed(500 + i)
            vc2, pi2, v2 = ref.run(b.copy(), ply=ply)
            assert [(mv.uci(), n) for mv, n in vc1.items()] == [(mv.uci(), n) for mv, n in vc2.items()], (b.fen(), ply)
            assert pi1.tobytes() == pi2.tobytes() and v1 == v2
            pr1 = [ch.prior for ch in m._last_root.children.values()]
            pr2 = [ch.prior for ch in ref._last_root.children.values()]
            np.testing.assert_allclose(pr1, pr2, rtol=1e-6, atol=1e-9)
            m._engine.close()


def test_device_dirichlet_sampler_moments(golden_dir):
    """noise == NULL: the kernel draws Dirichlet(alpha) itself (Marsaglia-Tsang gamma variates).  eta is recovered from the priors before
    and after (p' = 0.75 p + 0.25 eta, no clamp active) over 8192 roots of the start position (k = 20): E[eta_j] = 1/k,
    Var[eta_j] = (1/k)(1 - 1/k) / (k alpha + 1), sum_j eta_j = 1."""
    from matrix0_b200.engine import SearchEngine
    from matrix0_b200.mcts import MCTSConfig
    G, k, alpha = 8192, 20, 0.3
    eng = SearchEngine(G, max_nodes=64, tt_capacity=128, max_depth=16, hist_cap=4)
    cfg = MCTSConfig(num_simulations=1, **KW)
    eng.configure(cfg, deterministic=False, seed=12345)
    eng.set_boards([chess.Board()] * G, with_history=False)
    eng.begin()
    logits = torch.zeros((G, 4672), dtype=torch.float32, device="cuda")
    values = torch.zeros((G,), dtype=torch.float32, device="cuda")
    eng.expand_backup(logits, values)
    eng.result(with_pi=False)
    before = eng.res_prior[:, :k].clone()
    assert torch.allclose(before, torch.full_like(before, 1.0 / k))
    eng.add_dirichlet(None, None)
    eng.result(with_pi=False)
    eta = ((eng.res_prior[:, :k] - 0.75 * before) / 0.25).cpu().numpy()
    assert np.abs(eta.sum(1) - 1.0).max() < 1e-9 and eta.min() >= -1e-12
    mean, var = eta.mean(), eta.var()
    want_var = (1.0 / k) * (1 - 1.0 / k) / (k * alpha + 1)
    assert abs(mean - 1.0 / k) < 1e-9                       # exactly 1/k per row by construction
    assert abs(var / want_var - 1.0) < 0.03, (var, want_var)
    # marginal of one component: Beta(alpha, (k-1) alpha); compare the empirical CDF at a few points with numpy's own sampler
    ref = np.random.RandomState(0).dirichlet([alpha] * k, size=200000)[:, 0]
    for q in (1e-4, 1e-2, 0.05, 0.2, 0.5):
        assert abs((eta[:, 3] < q).mean() - (ref < q).mean()) < 0.02, q
    # two games / two launches never share a draw
    assert not np.allclose(eta[0], eta[1])
    eng.close()
