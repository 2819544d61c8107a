"""Shared helpers: replay the committed reference MCTS golden cases against any implementation."""
import json
import os

import chess
from oracle.backends import ConstantBackend, HashBackend

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcts_golden.json")


def load_cases():
    d = json.load(open(GOLDEN))
    return d["configs"], d["cases"]


def make_backend(spec):
    return ConstantBackend(spec[1]) if spec[0] == "const" else HashBackend(scale=spec[1], seed=spec[2])


def board_from(fen, moves):
    b = chess.Board(fen)
    for u in moves:
        b.push(chess.Move.from_uci(u))
    return b


def check_result(run, board, ply, expect, prior_rtol=0.0, tag=""):
    """run(board, ply) -> (visit_counts, pi, value, last_root) or raises RuntimeError."""
    if "error" in expect:
        try:
            run(board, ply)
        except RuntimeError as e:
            assert expect["error"] in str(e), (tag, str(e))
            return
        raise AssertionError(f"{tag}: expected RuntimeError({expect['error']})")
    vc, pi, v, root = run(board, ply)
    got = [[m.uci(), int(n)] for m, n in vc.items()]
    assert got == expect["visits"], (tag, got, expect["visits"])
    assert v == expect["value"], (tag, v, expect["value"])
    import numpy as np
    nz = [[int(i), float(pi[i])] for i in np.nonzero(pi)[0]]
    assert nz == expect["pi_nonzero"], tag
    if root is not None and "child_q" in expect:
        assert int(root.n) == expect["root_n"], tag
        ch = list(root.children.values())
        assert [float(c.q) for c in ch] == expect["child_q"], tag
        pr = [float(c.prior) for c in ch]
        if prior_rtol == 0.0:
            assert pr == expect["child_prior"], tag
        else:
            np.testing.assert_allclose(pr, expect["child_prior"], rtol=prior_rtol, atol=1e-9, err_msg=tag)


def run_all(make_mcts, prior_rtol=0.0, kinds=None):
    """make_mcts(cfg_dict, backend, sims) -> object with run(board, ply=) and _last_root."""
    cfgs, cases = load_cases()
    n = 0
    for ci, c in enumerate(cases):
        if kinds and c["kind"] not in kinds:
            continue
        if c["kind"] in ("fresh", "game_fresh"):
            m = make_mcts(cfgs[c["cfg"]], make_backend(c["backend"]), c["sims"])

            def run(b, ply, m=m):
                vc, pi, v = m.run(b, ply=ply)
                return vc, pi, v, m._last_root
            check_result(run, board_from(c["fen"], c["moves"]), c["ply"], c["expect"], prior_rtol, f"case {ci} {c['kind']} {c['cfg']}")
            n += 1
        else:
            k = 1 if c["kind"] == "persistent" else 2
            ms = [make_mcts(cfgs[c["cfg"]], make_backend(c["backend"]), c["sims"]) for _ in range(k)]
            for s in c["sequence"]:
                m = ms[s["ply"] % k]

                def run(b, ply, m=m):
                    vc, pi, v = m.run(b, ply=ply)
                    return vc, pi, v, m._last_root
                check_result(run, board_from(c["fen"], s["moves"]), s["ply"], s["expect"], prior_rtol,
                             f"case {ci} {c['kind']} {c['cfg']} ply {s['ply']}")
                n += 1
    return n
