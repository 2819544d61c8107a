"""Throughput / latency of the shared-memory evaluation server (matrix0_b200.inference.run_inference_server):
W worker threads, each posting B-row requests back to back through its mailbox for a few seconds against the R24 evaluator.
    python tools/bench_inference_server.py [--workers 16] [--rows 96] [--seconds 5]
Prints one JSON line: rows/s through the mailboxes (host buffers in, host buffers out), mean / p99 request latency, rows per
forward.  The reference's own server cannot be timed beside it: it does not start on a non-"mps" device (DESIGN Q13)."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=16)
    ap.add_argument("--rows", type=int, default=96)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    from matrix0_b200 import inference as m0inf
    from bench_selfplay import reference_cfg
    cfg = reference_cfg(800)["model"]                      # the R24 of BASELINE configs[3]
    res = [m0inf.setup_shared_memory_for_worker(i, 19, 4672, args.rows) for i in range(args.workers)]
    stop, ready = threading.Event(), threading.Event()
    err = []

    def serve():
        try:
            m0inf.run_inference_server("cuda", cfg, None, stop, ready, res, precision=args.precision)
        except Exception as e:  # surfaced below
            err.append(e)
            ready.set()

    st = threading.Thread(target=serve, daemon=True)
    st.start()
    assert ready.wait(120) and not err, err
    lat = [[] for _ in range(args.workers)]
    go, deadline = threading.Event(), [0.0]

    def worker(i):
        c = m0inf.InferenceClient(res[i])
        x = (np.random.default_rng(i).random((args.rows, 19, 8, 8)) < 0.15).astype(np.float32)
        c.infer_np(x)                                     # warm-up
        go.wait()
        while time.perf_counter() < deadline[0]:
            t = time.perf_counter()
            c.infer_np(x)
            lat[i].append(time.perf_counter() - t)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(args.workers)]
    [t.start() for t in ts]
    time.sleep(1.0)
    t0 = time.perf_counter()
    deadline[0] = t0 + args.seconds
    go.set()
    [t.join() for t in ts]
    wall = time.perf_counter() - t0
    stop.set()
    st.join(30)
    all_lat = np.sort(np.concatenate([np.asarray(l) for l in lat]))
    n_req = int(all_lat.size)
    print(json.dumps({"metric": "evaluation-server rows/s through shared-memory mailboxes", "value": n_req * args.rows / wall,
                      "unit": "positions/s", "workers": args.workers, "rows_per_request": args.rows, "requests": n_req,
                      "latency_ms_mean": float(all_lat.mean() * 1e3), "latency_ms_p99": float(all_lat[int(0.99 * (n_req - 1))] * 1e3),
                      "precision": args.precision, "model": "R24 320ch/24 blocks, random init",
                      "h2d_bytes_per_request": args.rows * 19 * 64 * 4, "d2h_bytes_per_request": args.rows * (4672 + 1) * 4}))


if __name__ == "__main__":
    main()
