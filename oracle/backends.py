"""TEST INFRASTRUCTURE ONLY -- deterministic stand-ins for the inference backend seam
(``infer_np``, azchess/selfplay/inference.py:585; the reference's own fakes are
tests/test_utils.py:38-50 ConstantBackend and tests/test_mcts_logits.py:9-18 DummyBackend)."""
from __future__ import annotations

import hashlib

import numpy as np


class ConstantBackend:
    """reference tests/test_utils.py:38-50: zero logits, constant value."""

    def __init__(self, value: float = 0.0):
        self.value = value
        self.calls = 0

    def infer_np(self, batch):
        batch = np.asarray(batch)
        if batch.ndim == 3:
            batch = batch[None]
        self.calls += 1
        b = batch.shape[0]
        return np.zeros((b, 4672), dtype=np.float32), np.full((b,), self.value, dtype=np.float32)


class HashBackend:
    """Stateless pseudo-network: logits and value are a pure function of the planes bytes, so the
    reference and the engine receive identical outputs for identical positions whatever the call
    order.  `scale` sets the logit spread (0.02 ~ a random-init net; 2.0 ~ a trained one)."""

    def __init__(self, scale: float = 1.0, seed: int = 0):
        self.scale, self.seed = float(scale), int(seed)
        self.calls = 0
        self.rows = 0
        self._cache = {}

    def _one(self, planes: np.ndarray):
        key = planes.tobytes()
        hit = self._cache.get(key)
        if hit is None:
            h = hashlib.sha256(key + self.seed.to_bytes(8, "little")).digest()
            rs = np.random.RandomState(int.from_bytes(h[:4], "little"))
            logits = (rs.standard_normal(4672) * self.scale).astype(np.float32)
            value = np.float32(np.tanh(rs.standard_normal() * 0.6))
            hit = (logits, value)
            self._cache[key] = hit
        return hit

    def infer_np(self, batch):
        batch = np.ascontiguousarray(np.asarray(batch, dtype=np.float32))
        if batch.ndim == 3:
            batch = batch[None]
        self.calls += 1
        self.rows += batch.shape[0]
        outs = [self._one(batch[i]) for i in range(batch.shape[0])]
        return np.stack([o[0] for o in outs]), np.array([o[1] for o in outs], dtype=np.float32)
