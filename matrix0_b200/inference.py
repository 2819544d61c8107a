"""Shared-memory evaluation server and client: the drop-in for ``azchess/selfplay/inference.py``.

Reference seam (SURVEY §8a S4, §8b "Server entry"):

* ``setup_shared_memory_for_worker(worker_id, planes, policy_size, max_batch_size)`` (``inference.py:18-35``) -- one dict
  per worker: ``request_tensor f32[max,planes,8,8]``, ``response_policy_tensor f32[max,policy]``,
  ``response_value_tensor f32[max,1]``, ``batch_size_tensor i32[1]`` (all in shared memory) and two ``multiprocessing.Event``s.
  The layout is kept byte for byte, so reference workers (their ``InferenceClient``) can talk to this server and this client
  can talk to the reference server (``tests/test_inference_cpu.py`` checks both directions where the reference is mounted).
* ``InferenceClient(resources).infer_np(arr)`` (``inference.py:578-681``) -- same shapes, time-outs, retries and exceptions.
* ``run_inference_server(device, model_cfg, model_state_dict, stop_event, server_ready_event, shared_memory_resources)``
  (``inference.py:101-574``) -- same signature; must set ``server_ready_event`` once the evaluator is loaded
  (``orchestrator.py:464`` waits 60 s for it).

What is different underneath (B200-first):

* the evaluator is the native tensor-core ``PolicyValueNet`` of this package (``m0_net_forward`` through the C ABI); there is
  no eager-PyTorch or CPU evaluator behind this entry point -- without the CUDA library ``run_inference_server`` raises
  ``NativeLibraryError`` and never sets the ready event;
* the workers' shared-memory tensors are page-locked in place (``cudaHostRegister``), so every request is ONE asynchronous
  copy from the worker's own shared tensor into its slice of the device batch and every response ONE copy back: the
  reference's ``torch.cat`` on the host, the staging ``.cpu()`` tensors and the per-worker ``copy_`` (two host copies each
  way) are gone;
* a sweep serves EVERY worker whose request is pending in one forward (rows of different workers are independent in the
  evaluator, so a worker's answer does not depend on how the server batched it); the reference's "half the target batch"
  heuristic (``inference.py:372-405``), which defers small requests behind a large one, is not reproduced;
* the 1 ms sleep of the reference's polling loop (``inference.py:331``) is kept only while the server is idle.
"""
from __future__ import annotations

import logging
import os
import time
from multiprocessing import Event
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native

logger = logging.getLogger(__name__)

RESOURCE_KEYS = ("request_tensor", "response_policy_tensor", "response_value_tensor", "request_event", "response_event",
                 "batch_size_tensor")


def setup_shared_memory_for_worker(worker_id: int, planes: int, policy_size: int, max_batch_size: int) -> Dict[str, Any]:
    """``inference.py:18-35``: the per-worker mailbox (``worker_id`` is unused there as well)."""
    import torch
    return {
        "request_tensor": torch.zeros((max_batch_size, planes, 8, 8), dtype=torch.float32).share_memory_(),
        "response_policy_tensor": torch.zeros((max_batch_size, policy_size), dtype=torch.float32).share_memory_(),
        "response_value_tensor": torch.zeros((max_batch_size, 1), dtype=torch.float32).share_memory_(),
        "request_event": Event(),
        "response_event": Event(),
        "batch_size_tensor": torch.tensor([0], dtype=torch.int32).share_memory_(),
    }


# --------------------------------------------------------------------------------------------------------------------
# client
# --------------------------------------------------------------------------------------------------------------------
class InferenceClient:
    """``inference.py:578-681``.  ``infer_np`` accepts ``(C,H,W)`` or ``(B,C,H,W)`` and returns
    ``(policy f32[B,policy_size], value f32[B])``; the policy array is a VIEW of the shared response tensor, valid until the
    next call (as in the reference)."""

    MAX_RETRIES = 2                 # inference.py:634
    RETRY_PAUSE_S = 0.1             # inference.py:668
    _timeout_scale = 1.0            # tests shrink the reference's 5-15 s time-outs; production leaves it at 1

    def __init__(self, resources: Dict[str, Any]):
        self.res = resources
        self.logger = logging.getLogger(__name__)

    @classmethod
    def request_timeout(cls, batch_size: int) -> float:
        """``inference.py:599-614``: 10 s for one row, 7.5 s up to 8, 5 s up to 32, then 5·(1 + B/64), capped at 15 s
        (12 / 18 / 24 s and a 30 s cap when ``MATRIX0_FAST_TIMEOUTS`` is switched off)."""
        fast = os.environ.get("MATRIX0_FAST_TIMEOUTS", "1").lower() in ("1", "true", "yes")
        base = 5.0 if fast else 12.0
        if batch_size == 1:
            t = base * 2.0
        elif batch_size <= 8:
            t = base * 1.5
        elif batch_size <= 32:
            t = base
        else:
            t = base * (1.0 + batch_size / 64.0)
        return min(t, 15.0 if fast else 30.0) * cls._timeout_scale

    def infer_np(self, arr_batch: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        import torch
        if arr_batch.ndim == 3:
            arr_batch = np.expand_dims(arr_batch, 0)
        if arr_batch.ndim != 4:
            self.logger.error(f"Invalid input shape for inference: {arr_batch.shape}")
            raise ValueError("InferenceClient expects (B,C,H,W) or (C,H,W)")
        batch_size = int(arr_batch.shape[0])
        if arr_batch.dtype != np.float32:
            arr_batch = arr_batch.astype(np.float32, copy=False)
        timeout = self.request_timeout(batch_size)
        res = self.res
        # Every failure below -- a request larger than the mailbox, a time-out after the retries, a malformed answer --
        # leaves as RuntimeError("Failed to copy data to shared memory: ...") chained to its cause, because the
        # reference's outermost handler (inference.py:676-681) wraps its own TimeoutError / RuntimeError as well.
        try:
            res["request_tensor"][:batch_size] = torch.from_numpy(np.ascontiguousarray(arr_batch))
            res["batch_size_tensor"][0] = batch_size
            res["request_event"].set()
            for attempt in range(self.MAX_RETRIES + 1):
                try:
                    if not res["response_event"].wait(timeout=timeout):
                        raise TimeoutError(f"Inference timeout after {timeout}s")
                    policy = res["response_policy_tensor"][:batch_size].numpy()
                    value = res["response_value_tensor"][:batch_size].numpy()
                    res["response_event"].clear()
                    if policy.shape[0] != batch_size or value.shape[0] != batch_size:
                        raise ValueError(f"Response shape mismatch: policy={policy.shape}, value={value.shape}, "
                                         f"expected_batch_size={batch_size}")
                    return policy, value.flatten()
                except TimeoutError:
                    if attempt >= self.MAX_RETRIES:
                        msg = (f"Inference timeout after {timeout}s for batch size {batch_size} "
                               f"(final attempt after {self.MAX_RETRIES + 1} retries)")
                        self.logger.error(msg)
                        raise TimeoutError(msg)
                    self.logger.warning(f"Inference timeout (attempt {attempt + 1}/{self.MAX_RETRIES + 1}), retrying...")
                    # as in the reference the request is NOT re-posted: both events are cleared and the wait repeats
                    res["request_event"].clear()
                    res["response_event"].clear()
                    time.sleep(self.RETRY_PAUSE_S * self._timeout_scale)
                except Exception as e:
                    if attempt >= self.MAX_RETRIES:
                        msg = f"Inference failed after {self.MAX_RETRIES + 1} attempts: {e}"
                        self.logger.error(msg)
                        raise RuntimeError(msg) from e
                    self.logger.warning(f"Inference error (attempt {attempt + 1}/{self.MAX_RETRIES + 1}): {e}, retrying...")
                    time.sleep(self.RETRY_PAUSE_S * self._timeout_scale)
        except Exception as e:
            msg = f"Failed to copy data to shared memory: {e}"
            self.logger.error(msg)
            raise RuntimeError(msg) from e
        raise AssertionError("unreachable")  # pragma: no cover


# --------------------------------------------------------------------------------------------------------------------
# server: the mailbox sweep (host logic, evaluator-agnostic) ...
# --------------------------------------------------------------------------------------------------------------------
Request = Tuple[int, int, int]      # (worker, rows, first row of the worker's slice in the sweep's batch)


def collect_requests(resources: Sequence[Dict[str, Any]]) -> List[Request]:
    """One sweep over the mailboxes: every worker whose request event is set is taken (event cleared, as
    ``inference.py:315-318``), a row count <= 0 is a spurious wake-up and ignored (``:360-365``), a row count above the
    mailbox capacity is clamped and written back (``:367-373``)."""
    out: List[Request] = []
    first = 0
    for w, res in enumerate(resources):
        ev = res["request_event"]
        if not ev.is_set():
            continue
        ev.clear()
        rows = int(res["batch_size_tensor"][0])
        if rows <= 0:
            continue
        cap = int(res["request_tensor"].shape[0])
        if rows > cap:
            logger.warning(f"Worker {w} batch_size {rows} > capacity {cap}; clamping")
            rows = cap
            res["batch_size_tensor"][0] = cap
        out.append((w, rows, first))
        first += rows
    return out


def serve_shared_memory(resources: Sequence[Dict[str, Any]], evaluate: Callable[[Sequence[Dict[str, Any]], List[Request]], None],
                        stop_event: Any, idle_sleep_s: float = 0.001) -> int:
    """The server loop.  ``evaluate(resources, requests)`` must fill ``response_policy_tensor[:rows]`` and
    ``response_value_tensor[:rows]`` of every listed worker; the loop then raises their response events.  If ``evaluate``
    raises, the failed workers' events are cleared (``inference.py:555-566``: their clients time out and retry) and the
    loop carries on.  Returns the number of sweeps that served at least one request."""
    sweeps = 0
    while not stop_event.is_set():
        requests = collect_requests(resources)
        if not requests:
            time.sleep(idle_sleep_s)
            continue
        try:
            evaluate(resources, requests)
        except Exception as e:
            logger.error(f"Error in batch processing: {e}", exc_info=True)
            for w, _, _ in requests:
                resources[w]["request_event"].clear()
                resources[w]["response_event"].clear()
            continue
        for w, _, _ in requests:
            resources[w]["response_event"].set()
        sweeps += 1
    return sweeps


# --------------------------------------------------------------------------------------------------------------------
# ... and the native evaluator behind it
# --------------------------------------------------------------------------------------------------------------------
class NativeBatchEvaluator:
    """Device side of the server: one device batch sized for all mailboxes together, the workers' shared tensors page-locked
    in place, one copy in and one copy out per request on the server's stream, one ``m0_net_forward`` per sweep."""

    def __init__(self, device: str, model_cfg: dict, model_state_dict: Optional[Dict[str, Any]],
                 resources: Sequence[Dict[str, Any]], precision: Optional[str] = None):
        import torch
        _native.load_library()                       # NativeLibraryError when the CUDA library is missing: no fallback
        if not torch.cuda.is_available():
            raise _native.NativeLibraryError("run_inference_server needs a CUDA device; matrix0_b200 has no CPU evaluator")
        from .model import PolicyValueNet
        dev = torch.device(device)
        if dev.type != "cuda":
            # the orchestrator passes its configured device string ("mps", "cpu", "auto" on the reference's machines)
            logger.warning(f"device {device!r} requested; the native evaluator runs on cuda:{torch.cuda.current_device()}")
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.precision = precision or os.environ.get("M0_SERVER_PRECISION", "fp16")
        self.model = PolicyValueNet.from_config(model_cfg, device=str(dev), precision=self.precision).eval()
        if model_state_dict:
            missing, unexpected = self.model.load_state_dict(model_state_dict, strict=False)
            if missing:
                logger.warning(f"Missing keys during load (initialized from defaults): {len(missing)} keys: {sorted(missing)}")
            if unexpected:
                logger.warning(f"Unexpected keys during load (ignored): {len(unexpected)} keys")
        else:
            logger.warning("No model state_dict provided, using random weights.")   # inference.py:231
        planes = {int(r["request_tensor"].shape[1]) for r in resources} or {int(self.model.cfg.planes)}
        widths = {int(r["response_policy_tensor"].shape[1]) for r in resources} or {int(self.model.cfg.policy_size)}
        if len(planes) != 1 or len(widths) != 1 or widths != {int(self.model.cfg.policy_size)}:
            raise ValueError(f"mailbox shapes (planes {planes}, policy {widths}) do not match the model "
                             f"(policy_size {self.model.cfg.policy_size})")
        capacity = max(1, sum(int(r["request_tensor"].shape[0]) for r in resources))
        with torch.cuda.device(dev):
            self.stream = torch.cuda.Stream()
            self.batch = torch.zeros((capacity, planes.pop(), 8, 8), dtype=torch.float32, device=dev)
        self.registered: List[int] = []
        for r in resources:
            for k in ("request_tensor", "response_policy_tensor", "response_value_tensor"):
                self._page_lock(r[k])
        self.forwards = 0
        self.rows = 0

    def _page_lock(self, t) -> None:
        """cudaHostRegister on a shared-memory tensor: copies to / from it become real asynchronous DMA.  A refusal (e.g. a
        locked-memory limit in a container) is not an error: the copies below are then ordinary pageable copies."""
        import torch
        try:
            rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
            if int(rc) == 0:
                self.registered.append(t.data_ptr())
            else:
                logger.warning(f"cudaHostRegister returned {int(rc)}; pageable copies for this mailbox")
        except Exception as e:                        # torch raises on a CUDA error code
            logger.warning(f"cudaHostRegister failed ({e}); pageable copies for this mailbox")

    def close(self) -> None:
        import torch
        for p in self.registered:
            try:
                torch.cuda.cudart().cudaHostUnregister(p)
            except Exception:
                pass
        self.registered = []

    def __call__(self, resources: Sequence[Dict[str, Any]], requests: List[Request]) -> None:
        import torch
        total = requests[-1][1] + requests[-1][2]
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            for w, rows, first in requests:
                self.batch[first:first + rows].copy_(resources[w]["request_tensor"][:rows], non_blocking=True)
            logits, values = self.model.forward_planes(self.batch[:total])
            values = values.unsqueeze(1)
            for w, rows, first in requests:
                resources[w]["response_policy_tensor"][:rows].copy_(logits[first:first + rows], non_blocking=True)
                resources[w]["response_value_tensor"][:rows].copy_(values[first:first + rows], non_blocking=True)
            self.stream.synchronize()                 # the answers are in shared memory before any response event is set
        self.forwards += 1
        self.rows += total


def run_inference_server(device: str, model_cfg: dict, model_state_dict: Optional[Dict[str, Any]], stop_event: Any,
                         server_ready_event: Any, shared_memory_resources: List[Dict[str, Any]],
                         precision: Optional[str] = None) -> None:
    """``inference.py:101-574`` on the native evaluator.  Errors while building the evaluator are logged AND re-raised (the
    reference only logs them and lets the orchestrator's 60 s ready time-out report the failure)."""
    logger.info(f"Inference server starting on device: {device}")
    logger.info(f"Available workers: {len(shared_memory_resources)}")
    try:
        evaluator = NativeBatchEvaluator(device, model_cfg, model_state_dict, shared_memory_resources, precision)
    except Exception as e:
        logger.error(f"Failed to create the evaluator: {e}")
        raise
    try:
        if shared_memory_resources:                   # first-use work (weight conversion, workspaces) before "ready"
            import torch
            with torch.cuda.device(evaluator.device), torch.cuda.stream(evaluator.stream):
                evaluator.model.forward_planes(evaluator.batch)
                evaluator.stream.synchronize()
        server_ready_event.set()
        logger.info("Inference server ready")
        sweeps = serve_shared_memory(shared_memory_resources, evaluator, stop_event)
        logger.info(f"Inference server shutting down after {sweeps} sweeps, {evaluator.rows} rows in {evaluator.forwards} forwards")
    finally:
        evaluator.close()
