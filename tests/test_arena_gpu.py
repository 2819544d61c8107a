"""Arena path (azchess/arena.py:59-126) against the oracle: every game of ArenaEngine's lock-step batch is replayed move by move with
the oracle search (RefMCTS) driven by the evaluator of the side to move -- A when (white to move) == (A has white), arena.py:66-71 --
deterministic settings (selection_jitter 0 -> the reference's random.random() pinned to 0.5, argmax moves after temp_plies = 0)."""
import numpy as np
import pytest
import torch

import chess
from oracle.mcts_ref import RefConfig, RefMCTS
from oracle.selfplay_ref import should_adjudicate_draw
from test_oracle_nn import load_case
from oracle import nn_ref
from matrix0_b200.model import parameter_shapes

pytestmark = pytest.mark.gpu


def test_arena_games_equal_oracle_replay_with_the_movers_network(golden_dir):
    from matrix0_b200.arena import ArenaEngine
    from matrix0_b200.model import PolicyValueNet
    g, cfg, sd_a = load_case(golden_dir, "small")
    sd_b = nn_ref.make_state_dict(parameter_shapes(cfg), seed=2)                    # a different network for B
    net_a = PolicyValueNet(cfg, device="cuda", precision="fp32"); net_a.load_state_dict(sd_a, strict=True)
    net_b = PolicyValueNet(cfg, device="cuda", precision="fp32"); net_b.load_state_dict(sd_b, strict=True)
    G, sims, max_moves = 8, 24, 9
    kw = dict(cpuct=2.5, cpuct_start=3.0, cpuct_end=2.0, cpuct_plies=40, fpu_reduction=0.1, draw_penalty=-0.05, legal_softmax=True,
              inference_batch_size=8, no_instant_backtrack=True, dirichlet_frac=0.0, playout_random_frac=0.0, enable_entropy_noise=False)
    cfg_dict = {"mcts": dict(kw, selection_jitter=0.05), "selfplay": {"selection_jitter": 0.0}}        # the arena reads selfplay.selection_jitter
    arena = ArenaEngine(net_a, net_b, cfg_dict, num_sims=sims, temperature=0.0, temp_plies=0, max_moves=max_moves, concurrent_games=G, seed=4,
                        precision="fp32", deterministic=True)
    games = arena.games
    games.a_is_white.copy_(torch.tensor([i % 2 == 0 for i in range(G)], device="cuda"))            # game idx = slot (arena.py:66)
    games.start()
    boards = [chess.Board() for _ in range(G)]
    alive = [True] * G
    results = {}
    for ply in range(max_moves + 1):
        if not any(alive):
            break
        games.begin_move()
        for _ in range(games.batches_per_move()):
            games.search_step()
        eng = games.engine
        eng.result(with_pi=False)
        cnt, mv, vis = eng.res_count.cpu().numpy(), eng.res_moves.cpu().numpy().view(np.uint16), eng.res_visits.cpu().numpy()
        games.end_move()
        played = games.moves_played.cpu().numpy().view(np.uint16)
        fin = {f["slot"]: f for f in games.finished_games()}
        for s in range(G):
            if not alive[s]:
                continue
            b = boards[s]
            a_white = s % 2 == 0
            net = net_a if (b.turn == chess.WHITE) == a_white else net_b                              # the mover's evaluator
            ref = RefMCTS(RefConfig(num_simulations=sims, selection_jitter=0.0, **kw), net, jitter_value=0.5)
            vc, _, _ = ref.run(b.copy(), ply=ply)
            got = [(int(mv[s, j]), int(vis[s, j])) for j in range(int(cnt[s]))]
            exp = [(m.from_square | (m.to_square << 6) | ((m.promotion or 0) << 12), n) for m, n in vc.items()]
            assert got == exp, (ply, s, b.fen())
            best = max(vc.items(), key=lambda kv: kv[1])[0]                                        # arena.py:116 (first maximum)
            assert int(played[s]) == (best.from_square | (best.to_square << 6) | ((best.promotion or 0) << 12)), (ply, s)
            b.push(best)
            over = b.is_game_over(claim_draw=True) or len(b.move_stack) >= max_moves or should_adjudicate_draw(b, list(b.move_stack), {})
            if over:
                assert s in fin, (ply, s, b.fen())
                res = b.result(claim_draw=True) if b.is_game_over(claim_draw=True) else "1/2-1/2"   # arena.py:118-126
                z = float(fin[s]["result"]) if fin[s]["reason"] == "checkmate" else 0.0
                assert {"1-0": 1.0, "0-1": -1.0}.get(res, 0.0) == z, (res, fin[s])
                results[s] = (1.0 + z) / 2.0 if a_white else (1.0 - z) / 2.0
                alive[s] = False
            else:
                assert s not in fin, (ply, s, fin.get(s))
    assert len(results) == G and games.rows_a > 0 and games.rows_b > 0
