"""GPU parity of the search AS SHIPPED (per-simulation selection jitter, entropy noise, child pruning, direct-model priors):
matrix0_b200.mcts.MCTS in its stochastic mode, fed the SAME streams of random draws a seeded reference run makes, against the
committed outputs of the UNMODIFIED reference (tests/golden/mcts_stochastic_golden.json).  Bars: visit counts, pi, root value and
child Q bit-exact; priors within 1e-6 relative (float32 softmax: CUDA expf vs torch CPU exp), exact for the zero-logit cases."""
import numpy as np
import pytest
import torch

import chess
import stochastic_cases as S
from conftest import random_playout_boards
from oracle.backends import HashBackend
from oracle.mcts_ref import RefConfig, RefMCTS
from test_oracle_nn import load_case

pytestmark = pytest.mark.gpu


def make_gpu(cfg, backend, sims, direct=False, **kw):
    from matrix0_b200.mcts import MCTS, MCTSConfig
    c = MCTSConfig(num_threads=1, enable_memory_cleanup=False, dirichlet_frac=0.0, playout_random_frac=0.0, num_simulations=sims, **cfg)
    return MCTS(c, None, device="cuda", inference_backend=backend, deterministic=False, max_nodes=65536, direct_model_priors=direct, **kw)


def test_reference_stochastic_goldens():
    d = S.load()
    kinds = {}
    for c in d["cases"]:
        jit, nrm = S.streams(c["seed"], c["expect"]["jitter_used"] + 8, c["expect"]["normal_used"] + 8)
        m = make_gpu(d["configs"][c["cfg"]], S.backend_of(c), c["sims"], direct=(c["cfg"] == "direct"))
        m.set_random_streams(jit, nrm)
        vc, pi, v = m.run(S.board_of(c), ply=c["ply"])
        S.check(c, vc, pi, v, m._last_root, prior_rtol=0.0 if c["backend"][1] == 0.0 else 1e-6)
        kinds[c["cfg"]] = kinds.get(c["cfg"], 0) + 1
        m._engine.close()
    assert sum(kinds.values()) >= 70 and set(kinds) == {"selfplay", "selfplay_b32", "arena_full", "prune", "direct", "default_full"}, kinds


def test_virtual_loss_throughput_mode_goldens():
    """Throughput mode (north_star: "PUCT selection with virtual loss"): the in-flight marking of the reference's own MCTS._select
    (inflight_counts, mcts.py:889-890, :922-923) applied inside every mini-batch, so that the simulations of a batch spread over
    distinct leaves (142-254 evaluator rows per 160-300 simulations in these cases).  Goldens: the UNMODIFIED _select driven with one
    dict per mini-batch (tests/golden/make_stochastic_golden.py vl), same seeded draws; visit counts, pi, value, child Q bit-exact."""
    d = S.load("mcts_vl_golden.json")
    assert d["virtual_loss_batches"] and len(d["cases"]) >= 20
    for c in d["cases"]:
        jit, nrm = S.streams(c["seed"], c["expect"]["jitter_used"] + 8, c["expect"]["normal_used"] + 8)
        m = make_gpu(d["configs"][c["cfg"]], S.backend_of(c), c["sims"], virtual_loss_batches=True)
        m.set_random_streams(jit, nrm)
        vc, pi, v = m.run(S.board_of(c), ply=c["ply"])
        S.check(c, vc, pi, v, m._last_root, prior_rtol=0.0 if c["backend"][1] == 0.0 else 1e-6)
        assert m._engine.counters()["nn_evals"] >= c["expect"]["distinct_rows"]          # (+ the root evaluation is counted by begin)
        m._engine.close()


def test_stream_exhaustion_is_reported():
    d = S.load()
    c = d["cases"][1]
    jit, nrm = S.streams(c["seed"], 50, 8)
    m = make_gpu(d["configs"][c["cfg"]], S.backend_of(c), c["sims"])
    m.set_random_streams(jit, nrm)
    with pytest.raises(RuntimeError, match="exhausted"):
        m.run(S.board_of(c), ply=c["ply"])


def test_device_generator_statistics():
    """Without supplied streams the draws come from the counter-based device generator: runs differ between seeds, visit counts sum to
    the simulation budget, and the search still concentrates on what a jitter-free search prefers."""
    kw = dict(cpuct=2.5, fpu_reduction=0.1, legal_softmax=True, selection_jitter=0.05, inference_batch_size=32, enable_entropy_noise=True)
    b = chess.Board("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
    outs = []
    for seed in (1, 2):
        m = make_gpu(kw, HashBackend(1.0, seed=5), 256, seed=seed)
        vc, pi, v = m.run(b, ply=40)
        assert sum(vc.values()) == 256 and abs(float(pi.sum()) - 1.0) < 1e-5
        outs.append(list(vc.values()))
    assert outs[0] != outs[1]


def test_as_shipped_batched_engine_matches_oracle_per_game(golden_dir):
    """SelfPlayEngine(search_mode='as_shipped'): every game of the lock-step batched search (compact evaluator batches over the distinct
    leaves of all games) == the oracle run alone on that game with the same streams of draws and the same CUDA fp32 evaluator."""
    from matrix0_b200.model import PolicyValueNet
    from matrix0_b200.selfplay import SelfPlayEngine
    g, cfg, sd = load_case(golden_dir, "small")
    net = PolicyValueNet(cfg, device="cuda", precision="fp32")
    net.load_state_dict(sd, strict=True)
    boards = random_playout_boards(5, 80, seed=23)[::6][:14]
    boards += [chess.Board("6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1"), chess.Board("4k3/8/8/8/8/8/8/4K2R w K - 148 90")]
    G, sims = len(boards), 200
    kw = dict(cpuct=2.5, cpuct_start=3.0, cpuct_end=2.0, cpuct_plies=40, fpu_reduction=0.1, draw_penalty=-0.05, legal_softmax=True,
              selection_jitter=0.05, inference_batch_size=48, no_instant_backtrack=True, enable_entropy_noise=True)
    c = {"mcts": dict(kw, num_simulations=sims, dirichlet_frac=0.0, playout_random_frac=0.0), "selfplay": {"num_simulations": sims, "opening_random_plies": 0}}
    sp = SelfPlayEngine(net, c, games=G, deterministic=False, seed=3, precision="fp32", search_mode="as_shipped", forward_rows=64)
    NJ, NN = 200_000, 40_000
    st = [S.streams(700 + i, NJ, NN) for i in range(G)]
    jit = torch.from_numpy(np.stack([s[0] for s in st])).cuda()
    nrm = torch.from_numpy(np.stack([s[1] for s in st])).cuda()
    sp.engine.set_streams(jit, nrm)
    sp.engine.set_boards(boards)
    sp.begin_move()
    for _ in range(sp.batches_per_move()):
        sp.search_step()
    eng = sp.engine
    eng.result(with_pi=True)
    cnt, moves, visits = eng.res_count.cpu().numpy(), eng.res_moves.cpu().numpy().view(np.uint16), eng.res_visits.cpu().numpy()
    root_q, pi = eng.res_root_q.cpu().numpy(), eng.res_pi.cpu().numpy()
    status, _ = eng.status()
    assert int(status.abs().sum()) == 0
    rows = 0
    for gi, b in enumerate(boards):
        ref = RefMCTS(RefConfig(num_simulations=sims, dirichlet_frac=0.0, playout_random_frac=0.0, **kw), net, jitter_value=None,
                      jitter_stream=st[gi][0], normal_stream=st[gi][1])
        vc, rpi, v = ref.run(b.copy(), ply=0)
        got = [(int(moves[gi, j]), int(visits[gi, j])) for j in range(int(cnt[gi]))]
        exp = [(m.from_square | (m.to_square << 6) | ((m.promotion or 0) << 12), n) for m, n in vc.items()]
        assert got == exp, (b.fen(), got, exp)
        assert root_q[gi] == v and pi[gi].tobytes() == rpi.tobytes(), b.fen()
        rows += ref.distinct_rows
    cn = sp.counters()
    assert cn["sims"] == G * sims and cn["nn_evals"] == rows     # one evaluator row per distinct leaf, as the oracle counts them
