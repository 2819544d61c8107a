"""Closed loop on the GPU: batched search (tree kernels + native evaluator, nothing leaves the device) against
the oracle search fed by the same network, and the device game loop's invariants."""
import numpy as np
import pytest
import torch

import chess
from conftest import random_playout_boards
from oracle.mcts_ref import RefConfig, RefMCTS
from test_oracle_nn import load_case

pytestmark = pytest.mark.gpu

MCTS_KW = dict(cpuct=2.5, cpuct_start=3.0, cpuct_end=2.0, cpuct_plies=40, fpu_reduction=0.1, draw_penalty=-0.05, legal_softmax=True,
               selection_jitter=0.05, inference_batch_size=96, no_instant_backtrack=True, dirichlet_frac=0.25, dirichlet_plies=30,
               playout_random_frac=0.05, enable_entropy_noise=True)


def small_net(golden_dir, precision="fp32"):
    from matrix0_b200.model import PolicyValueNet
    g, cfg, sd = load_case(golden_dir, "small")
    net = PolicyValueNet(cfg, device="cuda", precision=precision)
    net.load_state_dict(sd, strict=True)
    return net


def test_batched_search_matches_oracle_per_game(golden_dir):
    """Every game of a lock-step batched search == the reference algorithm run alone on that game (fresh MCTS),
    with the evaluator being the same CUDA fp32 network for both (rows are evaluated independently)."""
    from matrix0_b200.selfplay import SelfPlayEngine
    net = small_net(golden_dir)
    boards = random_playout_boards(6, 90, seed=17)[::5][:24]
    boards += [chess.Board("6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1"), chess.Board("4k3/8/8/8/8/8/8/4K2R w K - 148 90")]
    G = len(boards)
    sims = 300
    cfg = {"mcts": dict(MCTS_KW, num_simulations=sims), "selfplay": {"num_simulations": sims, "opening_random_plies": 0}}
    sp = SelfPlayEngine(net, cfg, games=G, deterministic=True, seed=1, precision="fp32", max_nodes=8192)
    sp.engine.set_boards(boards)
    sp.begin_move()
    for _ in range(sp.batches_per_move()):
        sp.search_step()
    eng = sp.engine
    eng.result(with_pi=True)
    cnt = eng.res_count.cpu().numpy()
    moves = eng.res_moves.cpu().numpy().view(np.uint16)
    visits = eng.res_visits.cpu().numpy()
    root_q = eng.res_root_q.cpu().numpy()
    pi = eng.res_pi.cpu().numpy()
    st, _ = eng.status()
    assert int(st.abs().sum()) == 0
    kw = {k: v for k, v in MCTS_KW.items()}
    for g, b in enumerate(boards):
        ref = RefMCTS(RefConfig(**{**kw, "num_simulations": sims, "dirichlet_frac": 0.0, "enable_entropy_noise": False,
                                   "playout_random_frac": 0.0}), net, jitter_value=0.5)
        vc, rpi, v = ref.run(b.copy(), ply=0)
        got = [(int(moves[g, j]), int(visits[g, j])) for j in range(int(cnt[g]))]
        exp = [(m.from_square | (m.to_square << 6) | ((m.promotion or 0) << 12), n) for m, n in vc.items()]
        assert got == exp, (b.fen(), got, exp)
        assert root_q[g] == v, b.fen()
        assert pi[g].tobytes() == rpi.tobytes(), b.fen()
    c = sp.counters()
    assert c["sims"] == G * sims


def test_device_game_loop_invariants(golden_dir):
    from matrix0_b200.selfplay import SelfPlayEngine
    net = small_net(golden_dir)
    G, sims = 96, 64
    cfg = {"mcts": dict(MCTS_KW, num_simulations=sims, inference_batch_size=32),
           "selfplay": {"num_simulations": sims, "opening_random_plies": 12, "max_game_len": 14, "temperature_start": 1.2, "temperature_end": 0.3,
                        "temperature_moves": 40, "resign_threshold": -0.85, "min_resign_plies": 50}}
    sp = SelfPlayEngine(net, cfg, games=G, deterministic=False, seed=5, precision="fp32", max_nodes=4096)
    sp.start()
    for _ in range(16):
        sp.play_move()
    torch.cuda.synchronize()
    c = sp.counters()
    assert c["positions_played"] == 16 * G
    lo, hi = int(sims * 0.95) * 16 * G, int(sims * 1.05) * 16 * G
    assert lo <= c["sims"] <= hi                       # playout-cap randomisation, mcts.py:380-385
    fin = sp.finished_games()
    assert len(fin) >= G                               # max_game_len 14 -> every slot finished at least one game
    assert c["games_finished"] == len(fin)
    for f in fin:
        assert 0 <= f["moves"] <= 14 and -1.0 <= f["result"] <= 1.0
        if f["reason"] == "max_game_len":
            assert f["moves"] == 14
        if f["reason"] in ("stalemate", "insufficient_material", "fifty_moves", "repetition"):
            assert f["result"] == 0.0
        if f["reason"] == "checkmate":
            assert abs(f["result"]) == 1.0
    st, nc = sp.engine.status()
    assert int(st.abs().sum()) == 0 and int(nc.max()) <= 4096
