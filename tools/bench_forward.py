"""BASELINE configs[2]: ResNet-24 inference forward alone, batch sweep, CUDA events.

    python tools/bench_forward.py [--precision fp16|bf16|fp32] [--batches 256,512,...] [--out gpurun_out/forward_sweep.json]

Methodology of the reference's tools/bench_inference.py:33-47 (warm-up, then timed iterations bracketed by CUDA events on the
launching stream); inputs are real encoded positions (binary planes) produced on the device by the random-playout kernel; weights
are random-init (deterministic seed).  positions/s and the fraction of the measured bf16 peaks are reported per batch size.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matrix0_b200 import _native  # noqa: E402
from bench_selfplay import FLOP_PER_POSITION, reference_cfg  # noqa: E402
from matrix0_b200.model import PolicyValueNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--batches", default="256,512,1024,2048,4096,8192")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/forward_sweep.json")
    args = ap.parse_args()
    lib = _native.lib()
    cfg = reference_cfg(800)
    net = PolicyValueNet.from_config(cfg["model"], device="cuda:0", precision=args.precision, seed=0)
    stream = torch.cuda.current_stream()
    peaks = {"burst": 1674.3, "sustained": 1403.5}
    try:
        mp = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        peaks = {"burst": float(mp.get("bf16_tflops_burst", mp.get("bf16_tflops", 1674.3))), "sustained": float(mp.get("bf16_tflops_sustained", 1403.5))}
    except (OSError, ValueError):
        pass
    res = []
    for B in [int(b) for b in args.batches.split(",")]:
        pos = torch.empty((B, 9), dtype=torch.int64, device="cuda")
        _native.check(lib.m0_random_playouts(pos.data_ptr(), B, 7, 80, stream.cuda_stream))
        planes = torch.empty((B, 19, 8, 8), dtype=torch.float32, device="cuda")
        _native.check(lib.m0_encode_planes(pos.data_ptr(), B, planes.data_ptr(), stream.cuda_stream))
        for _ in range(args.warmup):
            net.forward_planes(planes, args.precision)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.iters):
            net.forward_planes(planes, args.precision)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        tf = FLOP_PER_POSITION * B / (ms * 1e-3) / 1e12
        res.append({"batch": B, "ms": ms, "positions_per_s": B / (ms * 1e-3), "tflops": tf, "frac_burst": tf / peaks["burst"],
                    "frac_sustained": tf / peaks["sustained"], "precision": args.precision})
        print(res[-1], flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
