"""Self-play output contract of the reference: per finished game the arrays ``selfplay_worker`` hands to
``DataManager.add_selfplay_data`` (azchess/selfplay/internal.py:447-466, 612-651):

    s          float32[T, 19, 8, 8]   encode_board of the position before every searched move       (internal.py:447)
    pi         float32[T, 4672]       MCTS visit distribution returned by MCTS.run                   (internal.py:408, mcts.py:828-849)
    z          float32[T]             game result (White's view) x (+1 White / -1 Black to move)     (internal.py:612-616)
    legal_mask uint8[T, 4672]         MoveEncoder.get_legal_actions                                  (internal.py:454)
    meta_moves int32[1], meta_result float32[1], meta_resigned int8[1], meta_draw int8[1],
    meta_avg_policy_entropy float32[1], meta_avg_sims float32[1]                                    (internal.py:632-637)

The games run on the GPU (selfplay.SelfPlayEngine); per ply only the packed root positions (72 B) and the root children's
(policy index, visit count) pairs leave the device.  When a game ends, its positions go back through the encode kernel in one
batch (planes + legal masks) and ``pi`` is rebuilt as ``n / total`` (float64 divide, float32 store -- mcts.py:846).
``ssl_{task}`` arrays (internal.py:460-466, 644-648) come from the SSL target kernel (``m0_ssl_targets``) for the tasks in ``ssl_tasks``.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import numpy as np

from . import _native


class GameRecorder:
    """Collects the training records of the games a SelfPlayEngine plays.

        rec = GameRecorder(sp)
        sp.start()
        while ...:
            sp.begin_move(); [sp.search_step() ...]
            rec.after_search()          # root positions + visit counts of this ply -> pinned host memory
            sp.end_move()
            for game in rec.after_move():   # finished games as the reference's game_data dictionaries
                np.savez_compressed(path, **game)
    """

    def __init__(self, sp, keep_plies: int = 1024, ssl_tasks=()):
        import torch
        self.sp = sp
        self.G = sp.G
        self._torch = torch
        dev = sp.device
        self._pos_dev = torch.empty((self.G, 9), dtype=torch.int64, device=dev)
        self._idx_dev = torch.empty((self.G, 256), dtype=torch.int16, device=dev)
        self._cnt_dev = torch.empty((self.G,), dtype=torch.int32, device=dev)
        self._plies: List[Dict[str, np.ndarray]] = []   # one entry per searched ply of the engine (all slots)
        self._ply0 = 0                                   # absolute index of self._plies[0]
        self._start = np.zeros((self.G,), dtype=np.int64)   # absolute ply index at which the current game of a slot started
        self._sims = np.zeros((self.G,), dtype=np.float64)
        self.keep_plies = int(keep_plies)
        self.ssl_tasks = tuple(t for t in ssl_tasks if t in ("piece", "threat", "pin", "fork", "control"))   # model.ssl_tasks (internal.py:251-256)

    def after_search(self) -> None:
        """Call after the last search step of a ply and before SelfPlayEngine.end_move()."""
        torch = self._torch
        eng = self.sp.engine
        lib = _native.lib()
        st = _native.current_stream()
        eng.result(with_pi=False)
        _native.check(lib.m0_games_get_positions(eng._h, self._pos_dev.data_ptr(), st), "m0_games_get_positions")
        # policy index of every root child (child order = legal-move order = m0_legal_moves order)
        _native.check(lib.m0_legal_moves(self._pos_dev.data_ptr(), self.G, None, self._idx_dev.data_ptr(), self._cnt_dev.data_ptr(), st), "m0_legal_moves")
        rec = {"pos": self._pos_dev.cpu().numpy().copy(), "idx": self._idx_dev.cpu().numpy().view(np.uint16).copy(),
               "visits": eng.res_visits.cpu().numpy().copy(), "count": eng.res_count.cpu().numpy().copy(),
               "root_n": eng.res_root_n.cpu().numpy().copy()}
        self._plies.append(rec)

    def after_move(self) -> List[Dict[str, np.ndarray]]:
        """Call after SelfPlayEngine.end_move(); returns the game_data dictionaries of the games that just ended."""
        return list(self.iter_after_move())

    def iter_after_move(self):
        """after_move() as a generator: the finished games come one at a time and are assembled in groups of <= 65,536 positions, so a ply
        on which thousands of games end (all slots reaching max_game_len together) never holds more than one group's arrays
        (~2 GB of planes / masks / pi) on the host.  Exhaust it before the next after_search()."""
        now = self._ply0 + len(self._plies)          # absolute index one past the ply just played
        todo = []
        for fin in self.sp.finished_games():
            slot = fin["slot"]
            first = int(self._start[slot])
            T = now - first
            self._start[slot] = now
            if T <= 0 or first < self._ply0:
                continue  # the game began before the retained window (keep_plies too small): skipped, never truncated
            todo.append((slot, first, fin))
        group, rows = [], 0
        for item in todo:                                  # bounded device / host staging: <= 65,536 positions per encode launch
            T = now - item[1]
            if group and rows + T > 65536:
                yield from self._assemble_batch(group, now)
                group, rows = [], 0
            group.append(item)
            rows += T
        if group:
            yield from self._assemble_batch(group, now)
        # drop plies no live game needs any more
        lo = int(self._start.min())
        drop = max(0, min(lo - self._ply0, len(self._plies)))
        if len(self._plies) - drop > self.keep_plies:
            drop = len(self._plies) - self.keep_plies
        if drop:
            del self._plies[:drop]
            self._ply0 += drop

    def _assemble_batch(self, todo, now: int) -> List[Dict[str, np.ndarray]]:
        """game_data dictionaries of all games that ended this ply: ONE encode launch (planes + legal masks) and one SSL launch over the
        concatenated positions of all of them, one D2H copy per array, then per-game views."""
        torch = self._torch
        lib = _native.lib()
        rows_of = [[self._plies[i - self._ply0] for i in range(first, now)] for _, first, _ in todo]
        lens = [len(r) for r in rows_of]
        pos = np.concatenate([np.stack([r["pos"][slot] for r in rows]) for (slot, _, _), rows in zip(todo, rows_of)])     # [sum T, 9]
        N = pos.shape[0]
        dpos = torch.from_numpy(pos).to(self.sp.device)
        planes = torch.empty((N, 19, 8, 8), dtype=torch.float32, device=self.sp.device)
        mask = torch.empty((N, 4672), dtype=torch.uint8, device=self.sp.device)
        _native.check(lib.m0_encode_positions(dpos.data_ptr(), N, planes.data_ptr(), mask.data_ptr(), None, None, None, _native.current_stream()),
                      "m0_encode_positions")
        planes_h, mask_h = planes.cpu().numpy(), mask.cpu().numpy()
        ssl_h = {}
        if self.ssl_tasks:
            from .encoding import ssl_targets_device
            maps = ssl_targets_device(dpos)
            ssl_h = {t: maps[t].cpu().numpy() for t in self.ssl_tasks}
        out, off = [], 0
        for (slot, first, fin), rows, T in zip(todo, rows_of, lens):
            pi = np.zeros((T, 4672), dtype=np.float32)
            sims = np.empty((T,), dtype=np.float64)
            for t, r in enumerate(rows):
                k = int(r["count"][slot])
                n = r["visits"][slot, :k].astype(np.float64)
                tot = n.sum()
                if tot > 0:
                    pi[t, r["idx"][slot, :k].astype(np.int64)] = (n / tot).astype(np.float32)   # mcts.py:840-847
                sims[t] = tot
            turns = np.where((pos[off:off + T, 8] & 1) != 0, 1.0, -1.0).astype(np.float32)       # packed state word: bit 0 = side to move
            z = float(fin["result"])
            out.append({
                **{f"ssl_{t}": ssl_h[t][off:off + T].copy() for t in ssl_h},
                "s": planes_h[off:off + T].copy(), "pi": pi, "z": (z * turns).astype(np.float32), "legal_mask": mask_h[off:off + T].copy(),
                "meta_moves": np.array([T], dtype=np.int32), "meta_result": np.array([z], dtype=np.float32),
                "meta_resigned": np.array([1 if fin["resigned"] else 0], dtype=np.int8), "meta_draw": np.array([1 if z == 0.0 else 0], dtype=np.int8),
                "meta_avg_policy_entropy": np.array([fin["avg_policy_entropy"]], dtype=np.float32),
                "meta_avg_sims": np.array([float(sims.mean()) if T else 0.0], dtype=np.float32),
            })
            off += T
        return out


def write_game_npz(directory: str, game: Dict[str, np.ndarray], worker_id: int, game_id: int) -> str:
    """np.savez_compressed with the reference's shard naming (data_manager.py:198-243: selfplay_w{worker}_g{game}_{timestamp}.npz);
    the SQLite bookkeeping of DataManager is the caller's (out of scope here)."""
    import time
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, f"selfplay_w{worker_id}_g{game_id}_{int(time.time() * 1000)}.npz")
    np.savez_compressed(path, **game)
    return path
