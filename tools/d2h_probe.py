"""Host-side rates the game recorder depends on, measured on the box: page-locked allocation, device-to-host copy bandwidth into
page-locked and pageable memory, and the same copy while a GEMM loop keeps the SMs busy on another stream.
    python tools/d2h_probe.py [--out gpurun_out/d2h_probe.json]"""
import argparse
import json
import os
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/d2h_probe.json")
    a = ap.parse_args()
    out = {}
    torch.cuda.init()
    x = torch.empty((1 << 30,), dtype=torch.uint8, device="cuda")
    x.fill_(1)
    torch.cuda.synchronize()
    for gb in (0.25, 1.0, 2.0):
        n = int(gb * (1 << 30))
        t = time.perf_counter()
        h = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
        out[f"pinned_alloc_{gb}GB_s"] = time.perf_counter() - t
        t = time.perf_counter()
        h2 = torch.empty((n,), dtype=torch.uint8)
        h2.fill_(0)
        out[f"pageable_alloc_touch_{gb}GB_s"] = time.perf_counter() - t
        del h, h2
    t = time.perf_counter()
    h = torch.empty((1 << 30,), dtype=torch.uint8, pin_memory=True)
    out["pinned_alloc_1GB_cached_s"] = time.perf_counter() - t
    for name, dst in (("pinned", h), ("pageable", torch.empty((1 << 30,), dtype=torch.uint8))):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t = time.perf_counter()
            dst.copy_(x, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t)
        out[f"d2h_{name}_GBps"] = 1.073741824 / best
    # the same pinned copy on a side stream while a GEMM loop runs
    a_ = torch.randn((8192, 8192), dtype=torch.bfloat16, device="cuda")
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(200):
        a_ @ a_
    with torch.cuda.stream(side):
        e0.record(side)
        h.copy_(x, non_blocking=True)
        e1.record(side)
    torch.cuda.synchronize()
    out["d2h_pinned_under_gemm_GBps"] = 1.073741824 / (e0.elapsed_time(e1) * 1e-3)
    # host memcpy rate (what a per-game copy out of a staging buffer would cost)
    src = h.numpy()
    t = time.perf_counter()
    dst = src.copy()
    out["host_memcpy_GBps"] = 1.073741824 / (time.perf_counter() - t)
    out["host_cores"] = os.cpu_count()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
