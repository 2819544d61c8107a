"""Drop-in for ``azchess/mcts.py``: same ``MCTSConfig`` fields and ``MCTS`` call surface
(SURVEY.md section 8b), with the tree living on the GPU and every search step executed by the
warp-per-game CUDA kernels of ``csrc/tree_kernels.cu``.

``MCTS(cfg, model, device, inference_backend)`` accepts the reference's argument orders.  The
evaluator is either any object with ``infer_np`` (the reference's inference-backend seam,
``selfplay/inference.py:585``; test fakes such as ``ConstantBackend``) or a native evaluator
exposing ``forward_planes(planes_cuda) -> (logits_cuda, values_cuda)`` (``matrix0_b200.model``).

Two search modes.  ``deterministic=True`` (or env ``MATRIX0_DETERMINISTIC=1``) reproduces the reference run with
``random.random`` patched to 0.5 and noise off, bit for bit (visit counts): every simulation of a mini-batch then
selects the same leaf and the engine backs it up with the batch's multiplicity.  Otherwise the search runs
as the reference ships it: ``(random.random() - 0.5) * jitter`` on every PUCT score (``mcts.py:893-897``) drawn per
child per simulation, so a mini-batch collects many different leaves, each with its own evaluator row; entropy noise
on near-uniform priors (``mcts.py:170-186``), Dirichlet noise and playout-cap randomisation as configured.  The draws
come from a counter-based device generator; ``set_random_streams`` substitutes caller-supplied draws (the values
``random.random()`` / ``np.random.normal(0, 0.1)`` would return) so that a run can be compared with the reference
under a seeded RNG -- visit counts are then bit-exact (tests/test_mcts_stochastic_gpu.py).
"""
from __future__ import annotations

import logging
import os
import random
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .boards import POLICY_SIZE, code_to_move
from .engine import SearchEngine

logger = logging.getLogger(__name__)


@dataclass
class MCTSConfig:
    """Field-for-field mirror of ``azchess/mcts.py:61-107`` (names and defaults are the API)."""
    num_simulations: int = 800
    cpuct: float = 2.5
    dirichlet_alpha: float = 0.3
    dirichlet_frac: float = 0.25
    dirichlet_plies: int = 16
    tt_capacity: int = 2_000_000
    selection_jitter: float = 0.01
    tt_cleanup_frequency: int = 5000
    tt_memory_limit_mb: int = 2048
    fpu: float = 0.5
    fpu_reduction: float = 0.15
    parent_q_init: bool = True
    draw_penalty: float = -0.1
    virtual_loss: float = 1.0
    cpuct_start: Optional[float] = None
    cpuct_end: Optional[float] = None
    cpuct_plies: int = 0
    cpuct_c_base: Optional[float] = None
    cpuct_c_init: Optional[float] = None
    value_from_white: bool = False
    max_children: int = 0
    min_child_prior: float = 0.0
    legal_softmax: bool = False
    encoder_cache: bool = True
    tt_cleanup_interval_s: int = 5
    no_instant_backtrack: bool = True
    enable_memory_cleanup: bool = True
    memory_cleanup_threshold_mb: int = 1024
    max_tree_nodes: int = 100000
    num_threads: int = 6
    parallel_simulations: bool = True
    inference_batch_size: int = 96
    simulation_batch_size: int = 96
    tree_parallelism: bool = True
    playout_random_frac: float = 0.0
    enable_entropy_noise: bool = True

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "MCTSConfig":
        """``mcts.py:109-117``: unknown keys are dropped with a warning."""
        known = set(cls.__dataclass_fields__.keys())
        unknown = set(data.keys()) - known
        if unknown:
            logger.warning(f"Unknown MCTSConfig keys: {sorted(unknown)}")
        return cls(**{k: v for k, v in data.items() if k in known})


class _ChildView:
    __slots__ = ("n", "q", "prior", "move", "move_idx")

    def __init__(self, move, n, q, prior, move_idx):
        self.move, self.n, self.q, self.prior, self.move_idx = move, n, q, prior, move_idx


class _RootView:
    """What ``arena.py:737-747`` reads from ``mcts._last_root``: ``.children`` (Move -> .n/.q/.prior), ``.n``, ``.q``."""
    __slots__ = ("children", "n", "q")

    def __init__(self, children, n, q):
        self.children, self.n, self.q = children, n, q


class MCTS:
    def __init__(self, model_or_cfg, cfg_or_model, device: str = "cuda", inference_backend=None, num_threads: int = None,
                 *, deterministic: Optional[bool] = None, max_nodes: Optional[int] = None, seed: Optional[int] = None,
                 direct_model_priors: Optional[bool] = None, virtual_loss_batches: bool = False):
        # both argument orders, mcts.py:270-277
        if isinstance(model_or_cfg, MCTSConfig):
            cfg, model = model_or_cfg, cfg_or_model
        else:
            model, cfg = model_or_cfg, cfg_or_model
        import torch
        self.cfg = cfg
        self.model = model
        self.inference_backend = inference_backend
        dev = torch.device(device) if isinstance(device, str) else device
        if dev.type != "cuda":
            # "cpu"/"mps" strings come from reference configs; this engine only exists on the GPU
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.num_threads = num_threads if num_threads is not None else getattr(cfg, "num_threads", 1)
        if deterministic is None:
            deterministic = os.environ.get("MATRIX0_DETERMINISTIC", "") in ("1", "true", "yes")
        self.deterministic = bool(deterministic)
        if max_nodes is None:
            max_nodes = int(min(max(getattr(cfg, "max_tree_nodes", 100000), 1024), 1 << 20))
        if direct_model_priors is None:
            direct_model_priors = os.environ.get("MATRIX0_DIRECT_MODEL_PRIORS", "") in ("1", "true", "yes")
        # SURVEY Q3 switch: the reference WITHOUT an inference backend expands non-root leaves from raw logits (mcts.py:697-703)
        self.direct_model_priors = bool(direct_model_priors) and bool(getattr(cfg, "legal_softmax", False))
        # throughput mode: the reference's in-flight marking (mcts.py:889-890, :922-923 -- code it ships but never calls with a dict,
        # SURVEY Q2b) spreads the simulations of a mini-batch over distinct leaves, each with its own evaluator row
        self.virtual_loss_batches = bool(virtual_loss_batches) and not self.deterministic
        self._max_batch = int(getattr(cfg, "inference_batch_size", None) or getattr(cfg, "simulation_batch_size", 96))
        if self._max_batch <= 0:
            self._max_batch = 96
        with torch.cuda.device(dev):
            self._engine = SearchEngine(1, max_nodes=max_nodes, device=dev.index if dev.index is not None else torch.cuda.current_device())
            self._engine.configure(cfg, self.deterministic, seed if seed is not None else random.getrandbits(63),
                                   raw_logit_priors=self.direct_model_priors, virtual_loss=self.virtual_loss_batches)
            if not self.deterministic:
                self._engine.enable_multi(self._max_batch, virtual_loss=self.virtual_loss_batches)
        rows = 1 if self.deterministic else self._max_batch
        self._logits = torch.zeros((rows, POLICY_SIZE), dtype=torch.float32, device=dev)
        self._values = torch.zeros((rows,), dtype=torch.float32, device=dev)
        # (an even number of rows: the tensor-core evaluator works on pairs of boards)
        self._sample_planes = None if self.deterministic else torch.zeros((rows + (rows & 1), 19, 8, 8), dtype=torch.float32, device=dev)
        self._nn_cache: Dict[Tuple[int, int], float] = {}
        self.simulations_run = 0
        self._last_sims_run = 0
        self._last_root = None
        self.tt_hits = 0
        self.tt_misses = 0
        self.executor = None

    # ---- evaluator plumbing -----------------------------------------------------------------------
    def _policy_size(self) -> int:
        return int(getattr(getattr(self.model, "cfg", None), "policy_size", POLICY_SIZE))

    def _evaluate_pending(self, rows: int) -> None:
        """Evaluate the engine's pending leaf: `rows` identical rows go to an ``infer_np`` backend, as the
        reference's collected batch would contain; a native evaluator runs on the device."""
        import torch
        eng = self._engine
        if self.inference_backend is not None:
            planes = eng.planes[0].cpu().numpy()
            batch = np.ascontiguousarray(np.repeat(planes[None], max(1, rows), axis=0))
            policies, values = self.inference_backend.infer_np(batch)
            if policies is None or values is None:
                raise RuntimeError("Inference backend returned None results")
            policies = np.asarray(policies, dtype=np.float32)
            values = np.asarray(values, dtype=np.float32).reshape(-1)
            if policies.ndim == 1:
                policies = policies[None]
            if len(policies) != max(1, rows) or len(values) != max(1, rows):
                raise RuntimeError(f"Inference result shape mismatch: expected {rows}, got policies={len(policies)}, values={len(values)}")
            self._logits.copy_(torch.from_numpy(np.ascontiguousarray(policies[:1, :POLICY_SIZE])), non_blocking=False)
            self._values.copy_(torch.from_numpy(values[:1].copy()), non_blocking=False)
        elif hasattr(self.model, "forward_planes"):
            logits, values = self._native_forward(eng.planes)
            self._logits.copy_(logits.reshape(1, -1)[:, :POLICY_SIZE].float())
            self._values.copy_(values.reshape(-1)[:1].float())
        else:
            raise RuntimeError("MCTS needs an inference_backend with infer_np() or a native evaluator with forward_planes(); "
                               "matrix0_b200 has no PyTorch/CPU fallback evaluator")
        eng.expand_backup(self._logits, self._values)

    def _native_forward(self, planes):
        fwd = getattr(self.model, "forward_planes_graphed", None)
        return fwd(planes) if fwd is not None else self.model.forward_planes(planes)

    def set_random_streams(self, jitter=None, normal=None) -> None:
        """Caller-supplied draws for the stochastic search (parity testing): ``jitter`` = the values ``random.random()`` would
        return (one per child per visited node, selection order, mcts.py:893-897), ``normal`` = the ``np.random.normal(0, 0.1)``
        values of the entropy noise (k per noisy expansion, expansion order, mcts.py:181).  1-D float64 arrays; ``None`` returns to
        the device generator.  A run that exhausts a stream raises."""
        import torch
        if self.deterministic:
            raise RuntimeError("set_random_streams: the deterministic mode draws nothing")

        def dev(a):
            return None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(1, -1)).to(self.device)
        self._engine.set_streams(dev(jitter), dev(normal))

    def _evaluate_samples(self, n_samples: int) -> None:
        """The collected mini-batch (mcts.py:571-670).  An ``infer_np`` backend receives one row per SAMPLE in collection order,
        duplicates included, exactly the batch tensor the reference stacks (:604); each sample then backs up the value of its own
        row.  A native evaluator gets one row per distinct leaf."""
        import torch
        eng = self._engine
        if self.inference_backend is not None:
            eng.multi_encode(0, 1, 0, 2, self._sample_planes)
            batch = np.ascontiguousarray(self._sample_planes[:n_samples].cpu().numpy())
            policies, values = self.inference_backend.infer_np(batch)
            if policies is None or values is None:
                raise RuntimeError("Inference backend returned None results")
            policies = np.asarray(policies, dtype=np.float32)
            values = np.asarray(values, dtype=np.float32).reshape(-1)
            if policies.ndim == 1:
                policies = policies[None]
            if len(policies) != n_samples or len(values) != n_samples:
                raise RuntimeError(f"Inference result shape mismatch: expected {n_samples}, got policies={len(policies)}, values={len(values)}")
            self._logits[:n_samples].copy_(torch.from_numpy(np.ascontiguousarray(policies[:, :POLICY_SIZE])))
            self._values[:n_samples].copy_(torch.from_numpy(values.copy()))
            eng.expand_backup_multi(0, 1, self._logits, self._values, 0, per_sample=True)
        elif hasattr(self.model, "forward_planes"):
            rows = int(eng.row_base[1])
            eng.multi_encode(0, 1, 0, 0, self._sample_planes)
            # the whole (fixed) sample buffer: one cached CUDA graph serves every mini-batch, and a forward of <= 96 rows costs the same
            # as one of 2 (a single round of every kernel); rows past `rows` hold planes of earlier batches and are ignored
            half = getattr(self.model, "precision", "fp32") != "fp32"      # (the fp32 SIMT path runs eagerly, on the live rows only)
            logits, values = self._native_forward(self._sample_planes if half else self._sample_planes[:rows + (rows & 1)])
            self._logits[:rows].copy_(logits[:rows, :POLICY_SIZE].float())
            self._values[:rows].copy_(values.reshape(-1)[:rows].float())
            eng.expand_backup_multi(0, 1, self._logits, self._values, 0, per_sample=False)
        else:
            raise RuntimeError("MCTS needs an inference_backend with infer_np() or a native evaluator with forward_planes(); "
                               "matrix0_b200 has no PyTorch/CPU fallback evaluator")

    def _infer(self, board) -> Tuple[np.ndarray, float]:
        """``mcts.py:995-1221`` single-position evaluation (used by tests and for the reused-root value)."""
        from .encoding import encode_board
        enc = encode_board(board)[None]
        if self.inference_backend is not None:
            p, v = self.inference_backend.infer_np(enc)
            p = np.asarray(p, dtype=np.float32).reshape(1, -1)
            v = np.asarray(v, dtype=np.float32).reshape(-1)
        elif hasattr(self.model, "forward_planes"):
            import torch
            lg, vv = self.model.forward_planes(torch.from_numpy(enc).to(self.device))
            p, v = lg.float().cpu().numpy().reshape(1, -1), vv.float().cpu().numpy().reshape(-1)
        else:
            raise RuntimeError("no evaluator")
        if p.shape != (1, self._policy_size()):
            raise ValueError(f"Policy shape mismatch: got {p.shape}, expected {(1, self._policy_size())}")
        if not np.isfinite(p).all():
            raise ValueError("Policy logits contain NaN/Inf values")
        if not np.isfinite(v).all():
            raise ValueError("Value contains NaN/Inf values")
        v = np.clip(v, -1.0, 1.0)
        if bool(getattr(self.cfg, "value_from_white", False)) and not board.turn:
            v = -v
        return p[0], float(v[0])

    # ---- MCTS.run: mcts.py:318-512 ---------------------------------------------------------------------
    def run(self, board, num_simulations: Optional[int] = None, ply: Optional[int] = None):
        import torch
        try:
            with torch.cuda.device(self.device):
                return self._run(board, num_simulations, ply)
        except Exception as e:
            logger.error(f"MCTS run error: {e}")
            raise RuntimeError(f"MCTS run failed: {e}") from e

    def _run(self, board, num_simulations, ply):
        eng = self._engine
        st, nc = eng.status()
        if int(st[0]) != 0 or int(nc[0]) > eng.max_nodes - 4096:
            # the reference bounds its tree with LRU/RSS-driven cleanups (mcts.py:1285-1458); here the
            # pool is fixed, so a nearly full pool drops the reuse table (documented in DESIGN.md)
            eng.reset()
            self._nn_cache.clear()
        eng.set_boards([board])
        eng.begin()
        info = int(eng.info[0])
        if info & 1:
            return {}, np.zeros(self._policy_size(), dtype=np.float32), float(eng.term_value[0])
        v_root = 0.0
        if info & 2:
            self.tt_misses += 1
            if self.inference_backend is not None:
                # the reference validates the root evaluation in _infer (NaN / shape -> error)
                planes = eng.planes[0].cpu().numpy()[None]
                p, v = self.inference_backend.infer_np(planes)
                p = np.asarray(p, dtype=np.float32).reshape(1, -1)
                v = np.asarray(v, dtype=np.float32).reshape(-1)
                if p.shape[1] != self._policy_size():
                    raise ValueError(f"Policy shape mismatch: got {p.shape}, expected {(1, self._policy_size())}")
                if not np.isfinite(p).all() or not np.isfinite(v).all():
                    raise ValueError("Policy logits / value contain NaN/Inf values")
                import torch
                self._logits.copy_(torch.from_numpy(np.ascontiguousarray(p[:, :POLICY_SIZE])))
                self._values.copy_(torch.from_numpy(v[:1].copy()))
                eng.expand_backup(self._logits, self._values)
                v_root = float(np.clip(v[0], -1.0, 1.0))
            else:
                self._evaluate_pending(1)
                v_root = float(self._values[0].clamp(-1, 1))
            if bool(getattr(self.cfg, "value_from_white", False)) and not board.turn:
                v_root = -v_root
        else:
            self.tt_hits += 1
            if self.inference_backend is not None:
                # reused root: the reference re-evaluates it once per key for the fallback value (mcts.py:360-371)
                key = board._transposition_key() if hasattr(board, "_transposition_key") else None
                if key not in self._nn_cache:
                    _, v_root = self._infer(board)
                    if len(self._nn_cache) >= 10000:
                        self._nn_cache.pop(next(iter(self._nn_cache)))
                    self._nn_cache[key] = v_root
                else:
                    v_root = self._nn_cache[key]

        # Dirichlet noise at the root, gated by ply (mcts.py:373-376, :955-992)
        dirichlet_plies = getattr(self.cfg, "dirichlet_plies", None)
        if not self.deterministic and (dirichlet_plies is None or ply is None or ply < int(dirichlet_plies)):
            self._add_dirichlet()

        sims_to_run = num_simulations if num_simulations is not None else self.cfg.num_simulations
        frac = float(getattr(self.cfg, "playout_random_frac", 0.0))
        if not self.deterministic and frac > 0.0 and sims_to_run > 0:  # mcts.py:380-385
            low = int(max(1, sims_to_run * (1.0 - frac)))
            high = int(max(low, sims_to_run * (1.0 + frac)))
            sims_to_run = random.randint(low, high)

        max_batch = self._max_batch
        total, done = int(max(0, sims_to_run)), 0
        failures = attempts = 0
        while done < total:  # mcts.py:535-740
            batch_n = min(max_batch, total - done)
            if self.deterministic:
                eng.select(batch_n)
                m = int(eng.pending_counts()[0])
            else:
                eng.select_multi(batch_n)
                m = int(eng.n_samples[0])
            if m > 0:
                attempts += 1
                try:
                    if self.deterministic:
                        self._evaluate_pending(m)
                    else:
                        self._evaluate_samples(m)
                    failures = 0
                except (TimeoutError, RuntimeError) as err:
                    if self.inference_backend is None:
                        raise
                    failures += 1
                    if failures >= 3 and attempts >= 3:  # mcts.py:626-634
                        raise RuntimeError(f"MCTS inference completely failed: {failures} consecutive failures out of "
                                           f"{attempts} attempts. Last error: {err}") from err
                    logger.warning(f"Inference failed for batch of {m} positions: {err}. Skipping batch and continuing.")
            done += batch_n
        self.simulations_run += total

        eng.result(with_pi=True)
        k = int(eng.res_count[0])
        moves = eng.res_moves[0, :k].cpu().numpy().view(np.uint16)
        visits = eng.res_visits[0, :k].cpu().numpy()
        qs = eng.res_q[0, :k].cpu().numpy()
        priors = eng.res_prior[0, :k].cpu().numpy()
        pi = eng.res_pi[0].cpu().numpy()
        root_n = int(eng.res_root_n[0])
        root_q = float(eng.res_root_q[0])
        st, _ = eng.status()
        if int(st[0]) & 3:
            raise RuntimeError("search tree capacity exhausted (raise max_nodes / max_tree_nodes)")
        if int(st[0]) & 16:
            raise RuntimeError("set_random_streams: a supplied stream of random draws was exhausted during the search")
        mv_objs = [code_to_move(int(c)) for c in moves]
        visit_counts = {m: int(n) for m, n in zip(mv_objs, visits)}
        if sum(visit_counts.values()) == 0:  # mcts.py:433-463
            raise RuntimeError(f"MCTS search failed: zero visits after {sims_to_run} simulations. "
                               f"Root node has {k} children but none were visited.")
        policy_size = self._policy_size()
        if policy_size != POLICY_SIZE:
            out = np.zeros(policy_size, dtype=np.float32)
            out[:min(policy_size, POLICY_SIZE)] = pi[:min(policy_size, POLICY_SIZE)]
            pi = out
        self._last_sims_run = sims_to_run
        self._last_root = _RootView({m: _ChildView(m, int(n), float(q), float(p), None)
                                     for m, n, q, p in zip(mv_objs, visits, qs, priors)}, root_n, root_q)
        return visit_counts, pi, (root_q if root_n > 0 else float(v_root))

    def _add_dirichlet(self) -> None:
        """``mcts.py:955-992``.  The variates come from ``np.random.dirichlet`` exactly like the reference
        (same global RNG stream); mixing them into the priors happens on the device."""
        import torch
        eng = self._engine
        if float(self.cfg.dirichlet_frac) <= 0:
            return
        eng.result(with_pi=False)
        k = int(eng.res_count[0])
        if k <= 0:
            return
        noise = np.zeros((1, 256), dtype=np.float64)
        noise[0, :k] = np.random.dirichlet([self.cfg.dirichlet_alpha] * k)
        eng.add_dirichlet(torch.from_numpy(noise).to(self.device))

    # ---- housekeeping (mcts.py:1460-1501) -------------------------------------------------------------
    def get_memory_usage(self) -> Dict[str, int]:
        _, nc = self._engine.status()
        return {"tt_entries": int(nc[0]), "nn_cache_entries": len(self._nn_cache), "simulations_run": self.simulations_run,
                "device_bytes": self._engine.bytes}

    def reset(self):
        self._engine.reset()
        self._nn_cache.clear()
        self.simulations_run = 0
        self._last_sims_run = 0
        self._last_root = None
        self.tt_hits = 0
        self.tt_misses = 0

    def shutdown(self) -> None:
        self.executor = None
