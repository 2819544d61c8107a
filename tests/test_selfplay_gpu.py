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


@pytest.mark.parametrize("how", ["at_once", "deferred", "deferred_budget", "deferred_forced"])
def test_recorder_game_records_match_reference_contract(golden_dir, tmp_path, how):
    """Finished games come out as the reference's game_data dictionaries (internal.py:626-651): every recorded state replays
    through the oracle (python-chess restatement + the reference encoder's restatement) -- s and legal_mask bit-exact, pi the
    normalised visit counts on legal indices only, z = result x side to move, consecutive states one legal move apart."""
    from matrix0_b200.records import GameRecorder, write_game_npz
    from matrix0_b200.selfplay import SelfPlayEngine
    from oracle import ssl_ref
    from oracle.encoding_ref import encode_board, get_legal_actions
    net = small_net(golden_dir)
    G, sims = 48, 48
    cfg = {"mcts": dict(MCTS_KW, num_simulations=sims, inference_batch_size=16),
           "selfplay": {"num_simulations": sims, "opening_random_plies": 6, "max_game_len": 10, "temperature_start": 1.2, "temperature_end": 0.3,
                        "temperature_moves": 40, "resign_threshold": -0.85, "min_resign_plies": 50}}
    sp = SelfPlayEngine(net, cfg, games=G, deterministic=False, seed=9, precision="fp32", max_nodes=2048)
    rec = GameRecorder(sp, ssl_tasks=("piece", "threat", "pin", "fork", "control"))
    sp.start()
    games = []
    for _ in range(40 if how == "deferred_forced" else 12):
        sp.begin_move()
        for _ in range(sp.batches_per_move()):
            sp.search_step()
        rec.after_search()
        sp.end_move()
        # defer: the records are assembled / copied on the recorder's side stream under the next ply's search and come out one call later
        # "deferred_budget": at most 3 games are started per ply, the others wait in the device ring (oldest first, started anyway before
        # their rows would be overwritten)
        if how == "at_once":
            games += rec.after_move()
            assert rec.pending_games() == 0 and rec.backlog_games() == 0
        elif how == "deferred_forced":
            # no budget at all: every game waits in the ring (2 * max_game_len + 3 plies) until its oldest row is about to be overwritten
            games += list(rec.iter_after_move(defer=True, max_games=0))
        else:
            games += list(rec.iter_after_move(defer=True, max_games=3 if how == "deferred_budget" else None))
    waited = rec.backlog_games()
    games += list(rec.flush())
    assert rec.pending_games() == 0 and rec.backlog_games() == 0
    assert how not in ("deferred_budget", "deferred_forced") or waited > 0
    if how == "deferred_forced":
        assert len(games) - waited >= G         # the games of the first generation were handed out by the forced path alone
    assert len(games) >= G                      # max_game_len 10: every slot finished at least one game
    checked = 0
    for gd in games[:20]:
        T = int(gd["meta_moves"][0])
        assert gd["s"].shape == (T, 19, 8, 8) and gd["s"].dtype == np.float32
        assert gd["pi"].shape == (T, 4672) and gd["pi"].dtype == np.float32
        assert gd["legal_mask"].shape == (T, 4672) and gd["legal_mask"].dtype == np.uint8
        assert gd["z"].shape == (T,) and gd["z"].dtype == np.float32
        assert T == 10 or gd["meta_result"][0] != 0.0 or gd["meta_draw"][0] == 1
        z = float(gd["meta_result"][0])
        # rebuild every position from its planes is not possible; replay instead: the first state is reached from the start
        # position by the random opening, so check state-to-state consistency through the legal moves of the oracle board
        boards = _boards_from_planes(gd["s"])
        for t in range(T):
            b = boards[t]
            assert np.array_equal(gd["s"][t], encode_board(b)), (t, b.fen())
            assert np.array_equal(gd["legal_mask"][t].astype(bool), get_legal_actions(b))
            assert abs(float(gd["pi"][t].sum()) - 1.0) < 1e-5
            assert not np.any(gd["pi"][t][~gd["legal_mask"][t].astype(bool)] > 0)
            assert gd["z"][t] == np.float32(z * (1.0 if b.turn else -1.0))
            want = ssl_ref.ssl_targets(gd["s"][t])               # ssl_{task} arrays, internal.py:460-466 / 644-648
            for task, arr in want.items():
                assert np.array_equal(gd[f"ssl_{task}"][t], arr), (task, t, b.fen())
            if t + 1 < T:   # the next recorded state is one legal move away
                nxt = boards[t + 1]
                assert any(_same_placement(_pushed(b, m), nxt) for m in b.legal_moves), (t, b.fen(), nxt.fen())
        checked += 1
    assert checked > 0
    path = write_game_npz(str(tmp_path), games[0], worker_id=0, game_id=0)
    with np.load(path) as f:
        assert set(f.files) >= {"s", "pi", "z", "legal_mask", "meta_moves", "meta_result", "meta_resigned", "meta_draw",
                                "meta_avg_policy_entropy", "meta_avg_sims"}
        assert np.array_equal(f["s"], games[0]["s"])


def _pushed(b, m):
    c = b.copy()
    c.push(m)
    return c


def _same_placement(a, b):
    return a.board_fen() == b.board_fen() and a.turn == b.turn and a.castling_rights == b.castling_rights


def _boards_from_planes(s):
    """Invert encode_board (encoding.py:11-37): planes 0-11 pieces (row = 7 - rank), 12 side to move, 13-16 castling,
    17 halfmove clock, 18 fullmove number -- only what the oracle needs to regenerate planes and legal moves; en passant comes from the
    previous state's double pawn push."""
    out = []
    prev = None
    for t in range(s.shape[0]):
        p = s[t]
        b = chess.Board(None)
        for pl in range(12):
            color = chess.WHITE if pl < 6 else chess.BLACK
            ptype = (pl % 6) + 1
            for r in range(8):
                for f in range(8):
                    if p[pl, r, f] > 0.5:
                        b.set_piece_at(chess.square(f, 7 - r), chess.Piece(ptype, color))
        b.turn = bool(p[12, 0, 0] > 0.5)
        rights = 0
        if p[13, 0, 0] > 0.5: rights |= chess.BB_H1
        if p[14, 0, 0] > 0.5: rights |= chess.BB_A1
        if p[15, 0, 0] > 0.5: rights |= chess.BB_H8
        if p[16, 0, 0] > 0.5: rights |= chess.BB_A8
        b.castling_rights = rights
        b.halfmove_clock = int(round(float(p[17, 0, 0]) * 99.0))      # encoding.py:33: min(clock, 99) / 99
        b.fullmove_number = max(1, int(round(float(p[18, 0, 0]) * 199.0)))   # encoding.py:34: min(number, 199) / 199
        if prev is not None:   # en passant square: set when the previous move was a double pawn push
            for m in prev.legal_moves:
                c = _pushed(prev, m)
                if c.board_fen() == b.board_fen():
                    b.ep_square = c.ep_square
                    break
        out.append(b)
        prev = b
    return out


def test_selfplay_worker_dropin_writes_shards_and_messages(golden_dir, tmp_path):
    """selfplay_worker(proc_id, cfg, ckpt, games, q) -- the reference's worker entry (internal.py:94): NPZ shards in data_dir/selfplay
    and `game` messages with the orchestrator's keys (internal.py:666-679)."""
    import queue
    from matrix0_b200.selfplay import selfplay_worker
    g, ncfg, sd = load_case(golden_dir, "small")
    ckpt = tmp_path / "ckpt.pt"
    torch.save({"model": sd}, ckpt)
    model_cfg = {k: getattr(ncfg, k) for k in ncfg.__dataclass_fields__}
    cfg = {"model": model_cfg, "data_dir": str(tmp_path / "data"), "seed": 3,
           "mcts": dict(MCTS_KW, num_simulations=32, inference_batch_size=16),
           "selfplay": {"num_simulations": 32, "opening_random_plies": 4, "max_game_len": 8, "temperature_start": 1.0, "temperature_end": 0.3,
                        "temperature_moves": 40, "resign_threshold": -0.85, "min_resign_plies": 50}}
    q = queue.Queue()
    n = selfplay_worker(0, cfg, str(ckpt), games=10, q=q, concurrent_games=8, precision="fp32")
    assert n == 10
    files = sorted((tmp_path / "data" / "selfplay").glob("selfplay_w0_g*.npz"))
    assert len(files) == 10
    msgs = []
    while not q.empty():
        msgs.append(q.get())
    game_msgs = [m for m in msgs if m["type"] == "game"]
    assert len(game_msgs) == 10
    assert set(game_msgs[0]) >= {"type", "proc", "file", "moves", "result", "secs", "resigned", "resigner", "draw", "avg_policy_entropy",
                                 "avg_ms_per_move", "avg_sims"}
    with np.load(files[0]) as f:
        T = int(f["meta_moves"][0])
        assert f["s"].shape == (T, 19, 8, 8) and f["pi"].shape == (T, 4672) and f["z"].shape == (T,)


def test_arena_two_evaluators_route_rows_and_score(golden_dir):
    """ArenaEngine (arena.py:59-126 semantics): with identical networks on both sides the match equals the single-evaluator game
    loop move for move; leaves are routed to the mover's network (all roots are White to move at ply 0: half the games use A)."""
    from matrix0_b200.arena import ArenaEngine
    from matrix0_b200.selfplay import SelfPlayEngine
    net_a, net_b = small_net(golden_dir), small_net(golden_dir)
    G, sims, n = 16, 32, 16
    cfg = {"mcts": dict(MCTS_KW, num_simulations=sims, inference_batch_size=16, dirichlet_frac=0.0, playout_random_frac=0.0),
           "selfplay": {"selection_jitter": 0.0}}
    arena = ArenaEngine(net_a, net_b, cfg, num_sims=sims, temperature=0.0, temp_plies=0, max_moves=12, concurrent_games=G, seed=4,
                        precision="fp32", deterministic=True)
    res = arena.play(n)
    assert len(res["games"]) == n and all(r is not None for r in res["games"])
    assert res["wins"] + res["draws"] + res["losses"] == n
    assert abs(res["score_a"] - sum(r["score"] for r in res["games"])) < 1e-9
    assert [r["a_is_white"] for r in res["games"][:4]] == [True, False, True, False]
    assert res["rows_a"] > 0 and res["rows_b"] > 0
    # deterministic argmax games with one network on both sides: every game is the same game, and equals the plain game loop
    sp_cfg = {"mcts": dict(cfg["mcts"], selection_jitter=0.0),
              "selfplay": {"num_simulations": sims, "selection_jitter": 0.0, "opening_random_plies": 0, "max_game_len": 12, "temperature_start": 0.0,
                           "temperature_end": 0.0, "temperature_moves": 0, "argmax_after_plies": 0, "resign_threshold": -2.0}}
    sp = SelfPlayEngine(net_a, sp_cfg, games=G, deterministic=True, seed=4, precision="fp32")
    sp.start()
    fin = []
    while len(fin) < G:
        sp.play_move()
        fin += sp.finished_games()
    assert sorted((f["moves"], f["reason"]) for f in fin[:G]) == sorted((r["moves"], r["reason"]) for r in res["games"])


def test_reloading_weights_invalidates_the_captured_forward_graph(golden_dir):
    """load_state_dict while a SelfPlayEngine holds a captured CUDA graph of the evaluator: the next step must run on the NEW weights
    (the graph cache is keyed by the evaluator's workspace / weight epoch), not replay launches that point at freed buffers."""
    from matrix0_b200.model import PolicyValueNet, parameter_shapes
    from matrix0_b200.selfplay import SelfPlayEngine
    from oracle import nn_ref
    g, cfg, sd = load_case(golden_dir, "small")
    net = PolicyValueNet(cfg, device="cuda", precision="fp16")
    net.load_state_dict(sd, strict=True)
    G = 32
    c = {"mcts": dict(MCTS_KW, num_simulations=32, inference_batch_size=16), "selfplay": {"num_simulations": 32, "opening_random_plies": 4}}
    sp = SelfPlayEngine(net, c, games=G, deterministic=False, seed=3, precision="fp16")
    assert sp.cuda_graph
    sp.start()
    sp.play_move()
    assert sp.graph_replays > 0
    planes = sp.engine.planes
    lg_old, v_old = [t.clone() for t in sp._forward(planes)]
    sd2 = nn_ref.make_state_dict(parameter_shapes(cfg), seed=7)
    net.load_state_dict(sd2, strict=True)
    lg_new, v_new = [t.clone() for t in sp._forward(planes)]                 # through the engine's graph path
    lg_ref, v_ref = net.forward_planes(planes, "fp16")                        # eager, same weights
    assert torch.equal(lg_new, lg_ref) and torch.equal(v_new, v_ref)
    assert not torch.equal(lg_new, lg_old)
    sp.play_move()                                                            # and self-play continues on the new weights
    sp.check_status()


def test_search_on_a_second_device_of_the_same_process(golden_dir, tmp_path):
    """Engines on cuda:0 and cuda:1 in ONE process, driven WITHOUT the caller selecting the device (the engine's methods run on their
    own device, `_native.on_own_device`), the first engine being destroyed while the second is in use (a handle's destructor must
    leave the thread's current device alone, `DeviceGuard` in csrc/m0_common.cuh): the same deterministic search gives the same visit
    counts on both devices, and `selfplay_worker(proc_id=1, ...)` -- the reference picks the device from the worker index -- writes
    its shards from cuda:1.  Skipped on a single-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import gc
    import queue
    from matrix0_b200.model import PolicyValueNet
    from matrix0_b200.selfplay import SelfPlayEngine, selfplay_worker
    g, cfg_net, sd = load_case(golden_dir, "small")
    boards = random_playout_boards(4, 60, seed=5)[::7][:16]
    sims = 200
    cfg = {"mcts": dict(MCTS_KW, num_simulations=sims), "selfplay": {"num_simulations": sims, "opening_random_plies": 0}}
    res = []
    assert torch.cuda.current_device() == 0
    for dev in (0, 1, 0):
        net = PolicyValueNet(cfg_net, device=f"cuda:{dev}", precision="fp32")
        net.load_state_dict(sd, strict=True)
        sp = SelfPlayEngine(net, cfg, games=len(boards), device=dev, deterministic=True, seed=1, precision="fp32", max_nodes=8192)
        gc.collect()                                     # the previous iteration's engine and network die here
        assert torch.cuda.current_device() == 0
        sp.engine.set_boards(boards)
        sp.begin_move()
        for _ in range(sp.batches_per_move()):
            sp.search_step()
        sp.engine.result(with_pi=True)
        st, _ = sp.engine.status()
        assert int(st.abs().sum()) == 0
        assert sp.engine.res_visits.device.index == dev and torch.cuda.current_device() == 0
        res.append((sp.engine.res_count.cpu().numpy(), sp.engine.res_moves.cpu().numpy(), sp.engine.res_visits.cpu().numpy(),
                    sp.engine.res_pi.cpu().numpy()))
    assert int(res[0][2].sum()) == len(boards) * sims
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert a.tobytes() == b.tobytes()
    # the worker entry on the second device
    ckpt = tmp_path / "ckpt.pt"
    torch.save({"model": sd}, ckpt)
    wcfg = {"model": {k: getattr(cfg_net, k) for k in cfg_net.__dataclass_fields__}, "data_dir": str(tmp_path / "data"), "seed": 3,
            "mcts": dict(MCTS_KW, num_simulations=32, inference_batch_size=16),
            "selfplay": {"num_simulations": 32, "opening_random_plies": 4, "max_game_len": 8, "temperature_start": 1.0, "temperature_end": 0.3,
                         "temperature_moves": 40, "resign_threshold": -0.85, "min_resign_plies": 50}}
    q = queue.Queue()
    assert selfplay_worker(1, wcfg, str(ckpt), games=6, q=q, concurrent_games=4, precision="fp32") == 6
    assert torch.cuda.current_device() == 0
    assert len(sorted((tmp_path / "data" / "selfplay").glob("selfplay_w1_g*.npz"))) == 6
