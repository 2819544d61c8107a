"""End-to-end rate of the drop-in worker entry: selfplay_worker(proc_id, cfg, ckpt, games, q) (reference internal.py:94) from the call to
the last NPZ shard on disk -- search, game loop, record assembly and the compressed shard writes, everything a self-play phase of the
reference's orchestrator waits for.

    python tools/worker_throughput.py [--games 8192] [--concurrent 4096] [--mode as_shipped|collapsed] [--max-game-len 200]
                                      [--writers 8] [--level 1] [--dir /tmp/m0_worker] [--out gpurun_out/worker.json]

The shards are read back (np.load of a sample, schema + game count) and deleted.
"""
import argparse
import glob
import json
import os
import queue
import shutil
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bench_selfplay import reference_cfg  # noqa: E402
from matrix0_b200 import selfplay as m0_selfplay  # noqa: E402
from matrix0_b200.selfplay import selfplay_worker  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=8192)
    ap.add_argument("--concurrent", type=int, default=4096)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--mode", default="collapsed")
    ap.add_argument("--max-game-len", type=int, default=200)
    ap.add_argument("--writers", type=int, default=None)
    ap.add_argument("--level", type=int, default=None, help="deflate level of the shards (default: np.savez_compressed like the reference)")
    ap.add_argument("--dir", default="/tmp/m0_worker")
    ap.add_argument("--out", default="gpurun_out/worker.json")
    a = ap.parse_args()
    cfg = reference_cfg(a.sims)
    cfg["selfplay"]["max_game_len"] = a.max_game_len
    cfg["data_dir"] = a.dir
    cfg["model"]["self_supervised"] = True
    cfg["model"]["ssl_tasks"] = ["piece", "threat", "pin", "fork", "control"]
    if a.writers is not None:
        cfg["selfplay"]["writer_threads"] = a.writers
    if a.level is not None:
        cfg["selfplay"]["npz_compresslevel"] = a.level
    shutil.rmtree(a.dir, ignore_errors=True)
    q = queue.Queue()
    t0 = time.perf_counter()
    n = selfplay_worker(0, cfg, None, games=a.games, q=q, concurrent_games=a.concurrent, precision="fp16", search_mode=a.mode)
    total = time.perf_counter() - t0
    msgs = []
    while not q.empty():
        msgs.append(q.get())
    games = [m for m in msgs if m["type"] == "game"]
    plies = sum(m["moves"] for m in games)
    files = sorted(glob.glob(os.path.join(a.dir, "selfplay", "*.npz")))
    size = sum(os.path.getsize(f) for f in files)
    bad = []
    for f in files[:: max(1, len(files) // 16)]:
        with np.load(f) as z:
            T = int(z["meta_moves"][0])
            if z["s"].shape != (T, 19, 8, 8) or z["pi"].shape != (T, 4672) or z["legal_mask"].shape != (T, 4672) or z["ssl_piece"].shape != (T, 13, 8, 8):
                bad.append(f)
            if abs(float(z["pi"].sum()) - T) > 1e-3 * T:
                bad.append(f + ": pi")
    # the worker's own clock starts after the network is built and the evaluator graph is captured, and stops when the last shard is on disk
    stats = dict(m0_selfplay.LAST_WORKER_STATS)
    play = float(stats.get("seconds_play_and_write", total))
    out = {"mode": a.mode, "games_requested": a.games, "games_written": n, "shards_on_disk": len(files), "game_messages": len(games),
           "concurrent_games": a.concurrent, "sims_per_move": a.sims, "max_game_len": a.max_game_len, "plies_written": plies,
           "seconds_total_call": total, "seconds_playing_and_writing": play, "games_per_s": len(games) / play, "positions_per_s": plies / play,
           "sims_per_s": plies * a.sims / play, "bytes_written": size, "writer_threads": cfg["selfplay"].get("writer_threads", "default"),
           "npz_compresslevel": a.level, "where_the_time_went": stats, "host_cores": os.cpu_count(), "unreadable_or_wrong_shards": bad,
           "includes": "selfplay_worker end to end: network construction and graph capture (total only), search, game loop, device-side "
                       "record assembly, D2H, NPZ shards with SSL targets written by the writer threads, queue messages"}
    shutil.rmtree(a.dir, ignore_errors=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))
    if bad or len(files) != n:
        sys.exit(1)


if __name__ == "__main__":
    main()
