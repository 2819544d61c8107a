"""Micro-benchmark of the dominant kernel (tcgen05 implicit-GEMM 3x3 convolution 320 -> 320) with CUDA events."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matrix0_b200 import _native

lib = _native.lib()
res = []
shapes = [(4096, 320, 320, 9), (8192, 320, 320, 9), (1024, 320, 320, 9), (4096, 320, 320, 1), (296, 320, 320, 9)]
if len(sys.argv) > 1:
    shapes = [(int(sys.argv[1]), 320, 320, int(sys.argv[2]) if len(sys.argv) > 2 else 9)]
for boards, cin, n, taps in shapes:
    act = torch.randn((boards, 8, 8, cin), device="cuda").to(torch.bfloat16)
    w = torch.randn((n, taps * cin), device="cuda").to(torch.bfloat16)
    out = torch.empty((boards * 64, n), dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    for _ in range(3):
        _native.check(lib.m0_tc_conv(act.data_ptr(), w.data_ptr(), boards, cin, n, taps, out.data_ptr(), s.cuda_stream))
    torch.cuda.synchronize()
    reps = 3 if len(sys.argv) > 1 else 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(reps):
        _native.check(lib.m0_tc_conv(act.data_ptr(), w.data_ptr(), boards, cin, n, taps, out.data_ptr(), s.cuda_stream))
    b.record(s)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    flops = 2.0 * boards * 64 * n * taps * cin
    res.append({"boards": boards, "cin": cin, "n": n, "taps": taps, "ms": ms, "tflops": flops / ms / 1e9})
    print(res[-1], flush=True)
json.dump(res, open("gpurun_out/tc_conv_bench.json", "w"))
