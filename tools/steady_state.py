"""Steady-state self-play throughput INCLUDING what the worker path does per move beyond the search (SelfPlayEngine + GameRecorder:
root positions / visit counts to the host every ply, finished games assembled into the reference's game_data arrays, slots restarting):
    python tools/steady_state.py [--games 4096] [--moves 12] [--mode collapsed|as_shipped] [--out profiles/...json]
Wall-clock over >= 10 whole moves after one warm-up move (the recorder's host work is part of what is measured)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench_selfplay import reference_cfg  # noqa: E402
from matrix0_b200.model import PolicyValueNet  # noqa: E402
from matrix0_b200.records import GameRecorder  # noqa: E402
from matrix0_b200.selfplay import SelfPlayEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=4096)
    ap.add_argument("--moves", type=int, default=12)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--mode", default="collapsed")
    ap.add_argument("--max-game-len", type=int, default=200)
    ap.add_argument("--out", default="gpurun_out/steady_state.json")
    a = ap.parse_args()
    cfg = reference_cfg(a.sims)
    cfg["selfplay"]["max_game_len"] = a.max_game_len
    net = PolicyValueNet.from_config(cfg["model"], device="cuda:0", precision="fp16", seed=0)
    sp = SelfPlayEngine(net, cfg, games=a.games, device=0, deterministic=False, seed=1234, precision="fp16", search_mode=a.mode)
    rec = GameRecorder(sp, ssl_tasks=("piece", "threat", "pin", "fork", "control"))
    sp.warm_up_forward()
    sp.start()

    def move():
        sp.begin_move()
        for _ in range(sp.batches_per_move()):
            sp.search_step()
        rec.after_search()
        sp.end_move()
        return list(rec.iter_after_move(defer=True))     # assembled / copied on a side stream under the next ply's search

    move()
    torch.cuda.synchronize()
    c0 = sp.counters()
    t0 = time.perf_counter()
    per_move, games, plies = [], 0, 0
    for _ in range(a.moves):
        t = time.perf_counter()
        fin = move()
        torch.cuda.synchronize()
        per_move.append(time.perf_counter() - t)
        games += len(fin)
        plies += sum(int(g["meta_moves"][0]) for g in fin)
    dt = time.perf_counter() - t0
    c1 = sp.counters()
    d = {k: c1[k] - c0[k] for k in c1}
    sp.check_status()
    out = {"mode": a.mode, "games": a.games, "moves": a.moves, "seconds": dt, "sims_per_s": d["sims"] / dt, "positions_per_s": d["positions_played"] / dt,
           "nn_rows_per_s": (d["nn_evals"] + a.games * a.moves) / dt, "seconds_per_move": per_move, "games_finished_and_assembled": games,
           "plies_in_finished_games": plies, "max_game_len": a.max_game_len,
           "includes": "GameRecorder.after_search (positions, policy indices, visit counts into the device ring every ply), iter_after_move (finished games -> s / pi / z / "
                       "legal_mask / ssl_* arrays assembled on the device, one D2H per array and group), slot restarts with 12 random opening plies"}
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
