"""GPU parity of the shared-memory evaluation server (matrix0_b200.inference.run_inference_server on the native
evaluator, m0_net_forward through the C ABI): what a worker reads from its mailbox equals what the evaluator returns for the
same rows directly, whatever the server batched them with; a search driven through the mailbox equals the same search driven
by the evaluator directly; the fp32 path stays within 1e-4 relative of the reference outputs (committed goldens)."""
import threading

import numpy as np
import pytest
import torch

import chess
from test_oracle_nn import load_case

pytestmark = pytest.mark.gpu

PLANES, POLICY = 19, 4672


class Server:
    def __init__(self, cfg, sd, caps, precision):
        from dataclasses import asdict
        from matrix0_b200 import inference as m0inf
        self.m0inf = m0inf
        self.resources = [m0inf.setup_shared_memory_for_worker(i, PLANES, POLICY, c) for i, c in enumerate(caps)]
        self.stop, self.ready = threading.Event(), threading.Event()
        self.error = None
        self.t = threading.Thread(target=self._run, args=(asdict(cfg), sd, precision), daemon=True)

    def _run(self, cfg, sd, precision):
        try:
            self.m0inf.run_inference_server("cuda", cfg, sd, self.stop, self.ready, self.resources, precision=precision)
        except Exception as e:
            self.error = e

    def __enter__(self):
        self.t.start()
        assert self.ready.wait(60), self.error          # orchestrator.py:464 allows 60 s
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(30)
        assert not self.t.is_alive() and self.error is None, self.error

    def client(self, i):
        return self.m0inf.InferenceClient(self.resources[i])


def direct_net(cfg, sd, precision):
    from matrix0_b200.model import PolicyValueNet
    net = PolicyValueNet(cfg, device="cuda", precision=precision)
    net.load_state_dict(sd, strict=True)
    return net


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_mailbox_equals_direct_forward(golden_dir, precision):
    """Concurrent workers with ragged requests; rows are evaluated independently of their batch (bit-exact on the fp32
    path, tests/test_nn_gpu.py), so each worker's answer must equal the direct forward of its own rows."""
    g, cfg, sd = load_case(golden_dir, "small")
    net = direct_net(cfg, sd, precision)
    caps = [96, 8, 32, 1]
    errors, worst, direct_lock = [], [0.0], threading.Lock()      # one evaluator handle is driven by one thread at a time

    def worker(srv, i):
        try:
            rng = np.random.default_rng(40 + i)
            client = srv.client(i)
            for rows in [1, caps[i], max(1, caps[i] // 3), 1]:
                x = (rng.random((rows, PLANES, 8, 8)) < 0.2).astype(np.float32)
                p, v = client.infer_np(x if rows > 1 else x[0])
                p, v = p.copy(), v.copy()
                with direct_lock:
                    pr, vr = net.infer_np(x)
                assert p.shape == (rows, POLICY) and v.shape == (rows,)
                if precision == "fp32":
                    assert np.array_equal(p, pr) and np.array_equal(v, vr)
                else:
                    worst[0] = max(worst[0], float(np.abs(p - pr).max()), float(np.abs(v - vr).max()))
                    assert np.allclose(p, pr, rtol=0, atol=2e-3 * max(1.0, float(np.abs(pr).max()))) and np.allclose(v, vr, atol=2e-3)
        except Exception as e:
            errors.append((i, repr(e)))

    with Server(cfg, sd, caps, precision) as srv:
        ts = [threading.Thread(target=worker, args=(srv, i)) for i in range(len(caps))]
        [t.start() for t in ts]
        [t.join(120) for t in ts]
    assert not errors, errors


def test_fp32_server_vs_reference_golden(golden_dir):
    g, cfg, sd = load_case(golden_dir, "small")
    x = g["small_x"]
    with Server(cfg, sd, [max(8, len(x))], "fp32") as srv:
        p, v = srv.client(0).infer_np(x)
        for got, ref in ((p, g["small_logits"]), (v, g["small_values"])):
            assert float(np.abs(got - ref).max()) <= 1e-4 * max(float(np.abs(ref).max()), 1e-6)


def test_search_through_mailbox_equals_direct(golden_dir):
    """The reference's worker wiring (internal.py:305-310): MCTS(cfg, None, inference_backend=InferenceClient(res)).
    One worker, so the server's batches are the search's own batches and the result is bit-exact."""
    from matrix0_b200.mcts import MCTS, MCTSConfig
    g, cfg, sd = load_case(golden_dir, "small")
    net = direct_net(cfg, sd, "fp16")
    kw = dict(num_simulations=200, inference_batch_size=32, num_threads=1, enable_memory_cleanup=False, dirichlet_frac=0.0,
              enable_entropy_noise=False, playout_random_frac=0.0, legal_softmax=True, selection_jitter=0.0)
    boards = [chess.Board(), chess.Board("r1bqkbnr/pppp1ppp/2n5/4p3/4P3/5N2/PPPP1PPP/RNBQKB1R w KQkq - 2 3")]
    with Server(cfg, sd, [32], "fp16") as srv:
        for b in boards:
            a = MCTS(MCTSConfig(**kw), None, device="cuda", inference_backend=srv.client(0), deterministic=True).run(b.copy(), ply=40)
            d = MCTS(MCTSConfig(**kw), None, device="cuda", inference_backend=net, deterministic=True).run(b.copy(), ply=40)
            assert [(m.uci(), n) for m, n in a[0].items()] == [(m.uci(), n) for m, n in d[0].items()]
            assert a[1].tobytes() == d[1].tobytes() and a[2] == d[2]
