"""Whole games, start to end: every slot plays ONE complete self-play game (opening plies, search, temperature sampling, resign / draw
adjudication / max_game_len -- the reference's selfplay_worker loop, internal.py:326-687) with the game recorder attached, and every
finished game's arrays are checked against the reference's NPZ schema (internal.py:626-651) before they are dropped.

    python tools/full_games.py [--games 4096] [--sims 800] [--mode collapsed|as_shipped] [--out gpurun_out/full_games.json] [--npz-dir DIR --npz-keep 4]

What it shows beyond the 10-move steady-state run: tree reuse over a whole game (node pool / transposition table / history capacity
never exhausted: SelfPlayEngine.check_status() after every ply), the end-of-game paths (which reasons end games with a random-init
net and how long they are), and games / positions per second over complete games including the recorder's host work.
"""
import argparse
import collections
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench_selfplay import reference_cfg  # noqa: E402
from matrix0_b200.model import PolicyValueNet  # noqa: E402
from matrix0_b200.records import GameRecorder, write_game_npz  # noqa: E402
from matrix0_b200.selfplay import SelfPlayEngine  # noqa: E402

SSL_SHAPES = {"ssl_piece": None, "ssl_threat": None, "ssl_pin": None, "ssl_fork": None, "ssl_control": None}
META = {"meta_moves": np.int32, "meta_result": np.float32, "meta_resigned": np.int8, "meta_draw": np.int8,
        "meta_avg_policy_entropy": np.float32, "meta_avg_sims": np.float32}


def check_game(g, sims_lo, sims_hi):
    """Schema and invariants of one game_data dictionary; returns a list of violations (empty = fine)."""
    bad = []
    T = int(g["meta_moves"][0])
    for k, dt in META.items():
        if k not in g or g[k].shape != (1,) or g[k].dtype != dt:
            bad.append(f"{k}: {g.get(k, None) if k not in g else (g[k].shape, g[k].dtype)}")
    for k, shape, dt in (("s", (T, 19, 8, 8), np.float32), ("pi", (T, 4672), np.float32), ("z", (T,), np.float32), ("legal_mask", (T, 4672), np.uint8)):
        if g[k].shape != shape or g[k].dtype != dt:
            bad.append(f"{k}: {g[k].shape} {g[k].dtype}")
    if bad or T == 0:
        return bad + ([] if T else ["empty game"])
    for k in SSL_SHAPES:
        if k not in g or g[k].shape[0] != T or g[k].dtype != np.float32:
            bad.append(f"{k} missing / wrong leading dimension")
    rows = g["pi"].sum(1, dtype=np.float64)
    if np.abs(rows - 1.0).max() > 1e-4:
        bad.append(f"pi rows do not sum to 1 (max deviation {np.abs(rows - 1.0).max():.3g})")
    if ((g["pi"] > 0) & (g["legal_mask"] == 0)).any():
        bad.append("pi has mass on an illegal move")
    if (g["legal_mask"].sum(1) == 0).any():
        bad.append("a recorded position has no legal move")
    if not np.isin(g["s"][:, :12], (0.0, 1.0)).all():
        bad.append("piece planes are not binary")
    z = float(g["meta_result"][0])
    # decisive games: +-1; length cap / heuristic adjudication without a formal result: the last search value (internal.py:587-599)
    if not (-1.0 <= z <= 1.0) or not np.isin(g["z"], (np.float32(-z), np.float32(z))).all():
        bad.append(f"z / meta_result inconsistent ({z})")
    if T > 1 and z != 0.0 and not (g["z"][1:] == -g["z"][:-1]).all():
        bad.append("z does not alternate with the side to move")
    if int(g["meta_draw"][0]) != (1 if z == 0.0 else 0):
        bad.append("meta_draw")
    if not (sims_lo <= float(g["meta_avg_sims"][0]) <= sims_hi):
        bad.append(f"meta_avg_sims {float(g['meta_avg_sims'][0])} outside [{sims_lo}, {sims_hi}]")
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=4096)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--mode", default="collapsed")
    ap.add_argument("--max-game-len", type=int, default=200)
    ap.add_argument("--max-seconds", type=float, default=900.0)
    ap.add_argument("--npz-dir", default=None)
    ap.add_argument("--npz-keep", type=int, default=4)
    ap.add_argument("--out", default="gpurun_out/full_games.json")
    a = ap.parse_args()
    cfg = reference_cfg(a.sims)
    cfg["selfplay"]["max_game_len"] = a.max_game_len
    net = PolicyValueNet.from_config(cfg["model"], device="cuda:0", precision="fp16", seed=0)
    sp = SelfPlayEngine(net, cfg, games=a.games, device=0, deterministic=False, seed=1234, precision="fp16", search_mode=a.mode)
    rec = GameRecorder(sp, ssl_tasks=("piece", "threat", "pin", "fork", "control"))
    sp.warm_up_forward()
    frac = float(sp.mcfg.playout_random_frac)
    # a game's plies are searched with the playout-cap budget (mcts.py:380-385); tree reuse lets a root arrive with visits of earlier searches
    sims_lo, sims_hi = 0.9 * a.sims * (1.0 - frac), 64.0 * a.sims
    sp.start(a.games)            # every slot plays exactly one game
    torch.cuda.synchronize()
    c0 = sp.counters()
    t0 = time.perf_counter()
    reasons, results, lengths = collections.Counter(), collections.Counter(), []
    violations, done, kept, per_move, active_curve = [], 0, 0, [], []
    t_assemble = t_check = 0.0
    def consume(games_iter):
        nonlocal done, kept, t_assemble, t_check
        it = iter(games_iter)
        while True:
            ta = time.perf_counter()
            g = next(it, None)                   # the recorder waits for a group's device-to-host copies inside next()
            t_assemble += time.perf_counter() - ta
            if g is None:
                return
            done += 1
            T = int(g["meta_moves"][0])
            lengths.append(T)
            zr = float(g["meta_result"][0])
            results["white" if zr == 1.0 else "black" if zr == -1.0 else "draw" if zr == 0.0 else "last search value (no formal result)"] += 1
            reasons["resigned" if int(g["meta_resigned"][0]) else ("draw" if int(g["meta_draw"][0]) else "decisive")] += 1
            tc = time.perf_counter()
            bad = check_game(g, sims_lo, sims_hi)
            t_check += time.perf_counter() - tc
            if bad and len(violations) < 20:
                violations.append({"game": done, "plies": T, "violations": bad})
            if a.npz_dir and kept < a.npz_keep:
                write_game_npz(a.npz_dir, g, 0, done)
                kept += 1

    while done < a.games:
        t = time.perf_counter()
        sp.begin_move()
        for _ in range(sp.batches_per_move()):
            sp.search_step()
        rec.after_search()
        sp.end_move()
        sp.check_status()
        consume(rec.iter_after_move(defer=True))
        torch.cuda.synchronize()
        per_move.append(time.perf_counter() - t)
        active = sp.active_games()
        active_curve.append(active)
        if active == 0:
            break
        if time.perf_counter() - t0 > a.max_seconds:
            violations.append({"aborted": f"time limit {a.max_seconds}s with {active} games still running"})
            break
    consume(rec.flush())
    dt = time.perf_counter() - t0
    c1 = sp.counters()
    d = {k: c1[k] - c0[k] for k in c1}
    st, _ = sp.engine.status()
    L = np.asarray(lengths if lengths else [0])
    out = {"mode": a.mode, "games": a.games, "sims_per_move": a.sims, "max_game_len": a.max_game_len, "seconds": dt, "engine_plies": len(per_move),
           "games_finished": done, "games_per_s": done / dt, "seconds_in_record_assembly": t_assemble, "seconds_in_this_tools_checks": t_check,
           "assembly_positions_per_s": float(L.sum()) / max(t_assemble, 1e-9), "positions_per_s_without_the_checks": float(L.sum()) / (dt - t_check), "positions_recorded": int(L.sum()), "positions_per_s": float(L.sum()) / dt,
           "sims_per_s": d.get("sims", 0) / dt, "searched_plies_per_game": {"min": int(L.min()), "median": float(np.median(L)), "mean": float(L.mean()), "max": int(L.max())},
           "results": dict(results), "ends": dict(reasons),
           "status_bits_or": int(st.max()) if st.numel() else 0, "schema_violations": violations,
           "seconds_per_engine_ply": {"first": per_move[0], "median": float(np.median(per_move)), "max": float(np.max(per_move))},
           "active_games_every_20_plies": active_curve[::20], "counters": d,
           "includes": "search + GameRecorder (per-ply records kept in HBM, finished games assembled on the device into s / pi / z / legal_mask / ssl_* / meta_* and checked), "
                       "whole games from the 12 random opening plies to their end; slots go idle after their game (the engine keeps stepping all "
                       "slots, so games/s here is a lower bound of the continuous-restart rate in steady_state.py)"}
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))
    if violations:
        sys.exit(1)


if __name__ == "__main__":
    main()
