"""Multi-GPU plumbing of the self-play engine: games are independent, so ranks only share
(1) the evaluator parameters, broadcast once from rank 0, and (2) a few counters, reduced for reporting.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used for nothing else; the search path has no collective."""
from __future__ import annotations

from typing import Dict, Iterable, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `total` game ids owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Distinct, reproducible RNG stream per rank."""
    return (int(seed) + 7919 * int(rank)) & 0x7FFFFFFFFFFFFFFF


def broadcast_parameters(params: Dict[str, "object"], src: int = 0) -> None:
    """Make every rank hold rank `src`'s parameters: ONE broadcast of one packed buffer (the reference loads one checkpoint per phase,
    orchestrator.py:373-388).  Tensors are flattened in name order into a float32 buffer on the source rank and copied back in place."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    names = sorted(params)
    if not names:
        return
    ref = params[names[0]]
    total = sum(int(params[n].numel()) for n in names)
    buf = torch.empty((total,), dtype=torch.float32, device=ref.device)
    if dist.get_rank() == src:
        off = 0
        for n in names:
            k = int(params[n].numel())
            buf[off:off + k].copy_(params[n].reshape(-1).to(torch.float32))
            off += k
    dist.broadcast(buf, src=src)
    off = 0
    for n in names:
        k = int(params[n].numel())
        params[n].copy_(buf[off:off + k].reshape(params[n].shape).to(params[n].dtype))
        off += k


def reduce_scalars(values: Iterable[float], op: str = "sum", device=None):
    """All-reduce a short list of python numbers (float64) and return them as a list."""
    import torch
    import torch.distributed as dist
    vals = [float(v) for v in values]
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return vals
    t = torch.tensor(vals, dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return [float(x) for x in t.cpu()]
