"""Shared helpers for the stochastic-search goldens (tests/golden/mcts_stochastic_golden.json, generated from the UNMODIFIED
reference by tests/golden/make_stochastic_golden.py): the seeded streams of random draws and the result comparison."""
import json
import os
import random

import numpy as np

import chess
from oracle.backends import HashBackend

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcts_stochastic_golden.json")


def load(name="mcts_stochastic_golden.json"):
    return json.load(open(os.path.join(os.path.dirname(GOLDEN), name)))


def streams(seed, n_jitter, n_normal):
    """jitter = the values random.random() returns under random.seed(seed); normal = np.random.RandomState(seed).normal(0, 0.1, n)."""
    rj = random.Random(seed)
    return (np.array([rj.random() for _ in range(n_jitter)], dtype=np.float64), np.random.RandomState(seed).normal(0, 0.1, n_normal))


def board_of(case):
    b = chess.Board(case["fen"])
    for u in case["moves"]:
        b.push(chess.Move.from_uci(u))
    return b


def backend_of(case):
    return HashBackend(scale=case["backend"][1], seed=case["backend"][2])


def check(case, vc, pi, v, root, prior_rtol):
    e = case["expect"]
    tag = (case["cfg"], case["backend"], case["fen"], len(case["moves"]))
    assert [[m.uci(), int(n)] for m, n in vc.items()] == e["visits"], tag
    assert v == e["value"], tag
    assert [[int(i), float(pi[i])] for i in np.nonzero(pi)[0]] == e["pi_nonzero"], tag
    ch = list(root.children.values())
    assert int(root.n) == e["root_n"], tag
    assert [float(c.q) for c in ch] == e["child_q"], tag
    pr = [float(c.prior) for c in ch]
    if prior_rtol == 0.0:
        assert pr == e["child_prior"], tag
    else:
        # with entropy noise a prior is (softmax + noise) renormalised: the <= 2 ulp float32 softmax difference is an ABSOLUTE
        # error of ~1e-8 on a value that the noise may have cancelled down to 1e-4, hence the absolute term
        np.testing.assert_allclose(pr, e["child_prior"], rtol=prior_rtol, atol=2e-8, err_msg=str(tag))
