"""The device game loop (csrc/selfplay_kernels.cu) against the oracle's restatement of selfplay_worker's loop
(oracle/selfplay_ref.GameLoopRef: internal.py:382-600 + draw.py), fed the device's own search results and the same np.random draw:
move sampling by temperature, the resign rule, heuristic draw adjudication, game ends and results, restart budget."""
import numpy as np
import pytest
import torch

import chess
from oracle.selfplay_ref import GameLoopRef
from test_selfplay_gpu import MCTS_KW, small_net

pytestmark = pytest.mark.gpu


def _move(code):
    code = int(code)
    return chess.Move(code & 63, (code >> 6) & 63, promotion=((code >> 12) & 7) or None)


@pytest.mark.parametrize("variant", ["draw_heuristics", "resign"])
def test_device_game_loop_equals_reference_loop(golden_dir, variant):
    from matrix0_b200.selfplay import SelfPlayEngine, resolve_draw_config
    net = small_net(golden_dir)
    G, sims = 24, 24
    sp_cfg = {"num_simulations": sims, "opening_random_plies": 0, "max_game_len": 26, "temperature_start": 1.2, "temperature_end": 0.3,
              "temperature_moves": 8, "resign_threshold": -2.0, "min_resign_plies": 50}
    cfg = {"mcts": dict(MCTS_KW, num_simulations=sims, inference_batch_size=8), "selfplay": sp_cfg}
    if variant == "draw_heuristics":
        cfg["draw"] = {"enabled": True, "min_plies": 20, "window": 10, "min_unique": 3, "halfmove_cap": 40}
        sp_cfg["draw"] = {"min_plies": 6, "window": 6, "min_unique": 6, "halfmove_cap": 7, "material_draw_threshold": 74}   # selfplay.draw wins
        sp_cfg["low_visit_threshold"] = 12
    else:
        sp_cfg.update(resign_threshold=0.02, min_resign_plies=3, resign_window=4, resign_consecutive_bad=2, resign_min_entropy=9.0,
                      resign_value_margin=0.05)
    draw_cfg = resolve_draw_config(cfg)
    sp = SelfPlayEngine(net, cfg, games=G, deterministic=False, seed=11, precision="fp32")
    eng = sp.engine
    sp.start()
    loops = [GameLoopRef(sp_cfg, draw_cfg) for _ in range(G)]
    gen = torch.Generator(device="cuda").manual_seed(3)
    reasons, n_fin = {}, 0
    for ply in range(60):
        sp.begin_move()
        for _ in range(sp.batches_per_move()):
            sp.search_step()
        eng.result(with_pi=True)
        cnt = eng.res_count.cpu().numpy()
        mv = eng.res_moves.cpu().numpy().view(np.uint16)
        vis = eng.res_visits.cpu().numpy()
        pi = eng.res_pi.cpu().numpy()
        rq, rn = eng.res_root_q.cpu().numpy(), eng.res_root_n.cpu().numpy()
        u = torch.rand((G,), dtype=torch.float64, device="cuda", generator=gen)
        sp.set_sampling_uniforms(u)
        sp.end_move()
        played = sp.moves_played.cpu().numpy().view(np.uint16)
        fin = {f["slot"]: f for f in sp.finished_games()}
        uh = u.cpu().numpy()
        for g in range(G):
            L = loops[g]
            k = int(cnt[g])
            moves = [_move(c) for c in mv[g, :k]]
            assert moves == list(L.board.legal_moves), (ply, g)                      # the device plays the game the oracle replays
            v = float(rq[g]) if rn[g] > 0 else 0.0
            idx = L.play(moves, [int(x) for x in vis[g, :k]], pi[g], v, float(uh[g]))
            reason = "resign" if idx is None else None
            if idx is not None:
                assert int(played[g]) == int(mv[g, idx]), (ply, g, L.board.fen())       # same move from the same draw
                reason = L.ended()
            if reason is None:
                assert g not in fin, (ply, g, fin.get(g))
                continue
            f = fin[g]
            assert f["reason"] == reason and f["moves"] == L.n_states, (ply, g, f, reason, L.n_states)
            assert abs(f["result"] - np.float32(L.z)) < 1e-6, (f, L.z)
            reasons[reason] = reasons.get(reason, 0) + 1
            n_fin += 1
            loops[g] = GameLoopRef(sp_cfg, draw_cfg)                                   # the slot restarts from the start position
            assert loops[g].ended() is None
    assert n_fin >= G
    if variant == "draw_heuristics":
        assert reasons.get("draw_adjudicated", 0) > 0, reasons
    else:
        assert reasons.get("resign", 0) > 0, reasons


def test_start_budget_plays_every_started_game_to_its_end(golden_dir):
    """selfplay_worker's `games` argument (internal.py:326): exactly that many games are started; none is discarded."""
    from matrix0_b200.selfplay import SelfPlayEngine
    net = small_net(golden_dir)
    G, sims, want = 16, 16, 27
    cfg = {"mcts": dict(MCTS_KW, num_simulations=sims, inference_batch_size=8),
           "selfplay": {"num_simulations": sims, "opening_random_plies": 4, "max_game_len": 9, "resign_threshold": -2.0}}
    sp = SelfPlayEngine(net, cfg, games=G, deterministic=False, seed=2, precision="fp32")
    sp.start(want)
    fin = []
    for _ in range(60):
        if sp.active_games() == 0:
            break
        sp.play_move()
        fin += sp.finished_games()
    assert sp.active_games() == 0 and len(fin) == want
    assert sorted(f["moves"] for f in fin)[-1] <= 9 and any(f["reason"] == "max_game_len" for f in fin)
    sp.check_status()


def test_temperature_overflow_falls_back_to_uniform(golden_dir):
    """visits ** (1 / T) overflowing float32 (T = 0.1, > ~7000 visits on one move): the reference's probabilities become NaN and it picks
    a uniformly random legal move (internal.py:715-731); kernel and oracle restatement take the same branch with the same draw."""
    from matrix0_b200.selfplay import SelfPlayEngine
    from oracle.selfplay_ref import sample_move_from_counts
    net = small_net(golden_dir)
    G, sims = 2, 60000
    cfg = {"mcts": dict(MCTS_KW, num_simulations=sims, inference_batch_size=96, playout_random_frac=0.0, dirichlet_frac=0.0),
           "selfplay": {"num_simulations": sims, "opening_random_plies": 2, "max_game_len": 50, "temperature_start": 0.1, "temperature_end": 0.1,
                        "temperature_moves": 0, "resign_threshold": -2.0}}
    sp = SelfPlayEngine(net, cfg, games=G, deterministic=True, seed=4, precision="fp32")
    sp.start()
    sp.begin_move()
    for _ in range(sp.batches_per_move()):
        sp.search_step()
    eng = sp.engine
    eng.result(with_pi=False)
    cnt, mv, vis = eng.res_count.cpu().numpy(), eng.res_moves.cpu().numpy().view(np.uint16), eng.res_visits.cpu().numpy()
    u = torch.tensor([0.51, 0.07], dtype=torch.float64, device="cuda")
    sp.set_sampling_uniforms(u)
    sp.end_move()
    played = sp.moves_played.cpu().numpy().view(np.uint16)
    overflowed = 0
    for g in range(G):
        k = int(cnt[g])
        v = [int(x) for x in vis[g, :k]]
        with np.errstate(over="ignore"):
            overflowed += int(not np.isfinite((np.array(v, dtype=np.float32) ** np.float32(10.0)).sum()))
        idx = sample_move_from_counts(list(range(k)), v, 0.1, float(u[g]))
        assert int(played[g]) == int(mv[g, idx]), (g, v)
    assert overflowed > 0
