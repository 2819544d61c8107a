"""Drop-in for ``azchess/encoding.py`` backed by the CUDA encode kernels.

Same names, argument meaning and error behaviour as the reference:
``encode_board`` (encoding.py:11-37), ``move_to_index`` (:113-150), ``MoveEncoder`` (:153-273),
``move_encoder`` (:277), ``build_horizontal_flip_permutation`` / ``build_rotate180_permutation``
(:310-386), ``POLICY_SHAPE`` (:51).  The batch functions (``encode_boards``, ``legal_moves_batch``)
are the form the engine itself uses; the single-board functions are thin wrappers that upload one
packed position, launch the same kernel and read the result back.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .boards import MAX_MOVES, PLANES, POLICY_SIZE, POSITION_WORDS, boards_to_raw, code_to_move, move_to_code

POLICY_SHAPE = (8, 8, 73)
LEGACY_POLICY_SIZE = 4672
USE_AZ1858 = False

RAY_DIRS: Tuple[Tuple[int, int], ...] = ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (1, -1), (-1, 1), (-1, -1))
KNIGHT_DELTAS: Tuple[Tuple[int, int], ...] = ((-2, -1), (-2, 1), (-1, -2), (-1, 2), (1, -2), (1, 2), (2, -1), (2, 1))
UNDERPROMO_PIECES: Tuple[int, ...] = (2, 3, 4)  # knight, bishop, rook (python-chess piece types)


# ---- device helpers ---------------------------------------------------------------------------
def upload_positions(boards: Sequence, device=None):
    """Boards -> packed positions ``uint64[n, 9]`` resident on the GPU (``m0_positions_pack``)."""
    import torch
    lib = _native.lib()
    device = torch.device("cuda" if device is None else device)
    raw = torch.from_numpy(boards_to_raw(boards).view(np.int64)).to(device, non_blocking=False)
    pos = torch.empty((len(boards), POSITION_WORDS), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        _native.check(lib.m0_positions_pack(raw.data_ptr(), len(boards), pos.data_ptr(), _native.current_stream()),
                      "m0_positions_pack")
    return pos


def encode_positions_device(pos, planes: bool = True, mask: bool = True, moves: bool = False):
    """Run the fused encode kernel on packed positions already in HBM; returns device tensors."""
    import torch
    lib = _native.lib()
    n = pos.shape[0]
    dev = pos.device
    t_planes = torch.empty((n, PLANES, 8, 8), dtype=torch.float32, device=dev) if planes else None
    t_mask = torch.empty((n, POLICY_SIZE), dtype=torch.uint8, device=dev) if mask else None
    t_moves = torch.empty((n, MAX_MOVES), dtype=torch.int16, device=dev) if moves else None
    t_idx = torch.empty((n, MAX_MOVES), dtype=torch.int16, device=dev) if moves else None
    t_cnt = torch.empty((n,), dtype=torch.int32, device=dev) if moves else None
    with torch.cuda.device(dev):
        _native.check(lib.m0_encode_positions(pos.data_ptr(), n, _native.ptr(t_planes), _native.ptr(t_mask),
                                              _native.ptr(t_moves), _native.ptr(t_idx), _native.ptr(t_cnt),
                                              _native.current_stream()), "m0_encode_positions")
    return t_planes, t_mask, t_moves, t_idx, t_cnt


def ssl_targets_device(pos):
    """SSL target maps (azchess/ssl_algorithms.py create_enhanced_ssl_targets) of packed positions already in HBM:
    dict of float32 device tensors {"piece": [n,13,8,8], "threat" / "pin" / "fork" / "control": [n,8,8]}."""
    import torch
    lib = _native.lib()
    n = pos.shape[0]
    out = {"piece": torch.empty((n, 13, 8, 8), dtype=torch.float32, device=pos.device)}
    for k in ("threat", "pin", "fork", "control"):
        out[k] = torch.empty((n, 8, 8), dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        _native.check(lib.m0_ssl_targets(pos.data_ptr(), n, out["piece"].data_ptr(), out["threat"].data_ptr(), out["pin"].data_ptr(),
                                         out["fork"].data_ptr(), out["control"].data_ptr(), _native.current_stream()), "m0_ssl_targets")
    return out


def create_enhanced_ssl_targets(boards: Sequence, device=None) -> dict:
    """Batch form of ``ChessSSLAlgorithms.create_enhanced_ssl_targets`` (ssl_algorithms.py:502) taking boards instead of planes."""
    pos = upload_positions(boards, device)
    return {k: v.cpu().numpy() for k, v in ssl_targets_device(pos).items()}


def encode_boards(boards: Sequence, device=None) -> Tuple[np.ndarray, np.ndarray]:
    """Batch form of ``encode_board`` + ``get_legal_actions``: (float32[n,19,8,8], bool[n,4672])."""
    if len(boards) == 0:
        return np.zeros((0, PLANES, 8, 8), np.float32), np.zeros((0, POLICY_SIZE), bool)
    pos = upload_positions(boards, device)
    planes, mask, _, _, _ = encode_positions_device(pos, True, True, False)
    return planes.cpu().numpy(), mask.cpu().numpy().astype(bool)


def legal_moves_batch(boards: Sequence, device=None) -> List[List[Tuple[int, int]]]:
    """Per board: [(move_code, policy_index), ...] in python-chess generation order."""
    if len(boards) == 0:
        return []
    pos = upload_positions(boards, device)
    _, _, moves, idx, cnt = encode_positions_device(pos, False, False, True)
    moves = moves.cpu().numpy().view(np.uint16)
    idx = idx.cpu().numpy().view(np.uint16)
    cnt = cnt.cpu().numpy()
    return [[(int(moves[i, k]), int(idx[i, k])) for k in range(int(cnt[i]))] for i in range(len(boards))]


# ---- reference API ----------------------------------------------------------------------------
def encode_board(board, planes: int = 19) -> np.ndarray:
    """``azchess/encoding.py:11-37``: float32 [19, 8, 8], row = 7 - rank, col = file."""
    if planes != PLANES:
        raise ValueError(f"Expected {planes} planes, got {PLANES}")
    pos = upload_positions([board])
    out, _, _, _, _ = encode_positions_device(pos, True, False, False)
    return out[0].cpu().numpy()


def move_to_index(board, move) -> int:
    """``azchess/encoding.py:113-150``; raises ``ValueError`` for an illegal move like the reference."""
    if USE_AZ1858:
        raise NotImplementedError("1858 move_to_index not yet implemented")
    code = move_to_code(move)
    for c, i in legal_moves_batch([board])[0]:
        if c == code:
            return i
    raise ValueError(f"Illegal move: {move}")


@dataclass
class MoveEncoder:
    """``azchess/encoding.py:153-273`` -- encode/decode utilities and masks."""

    def __post_init__(self):
        self._cache: Dict[Tuple[object, object], int] = {}
        self._cache_hits = 0
        self._cache_misses = 0

    def encode_move(self, board, move) -> int:
        try:
            cache_key = (board._transposition_key(), move)
        except Exception:
            cache_key = None
        if cache_key is not None and cache_key in self._cache:
            self._cache_hits += 1
            return self._cache[cache_key]
        self._cache_misses += 1
        try:
            result = move_to_index(board, move)
        except ValueError:
            fen = board.fen() if hasattr(board, "fen") else "?"
            raise ValueError(f"Illegal move: {move} in position {fen}")
        if cache_key is not None:
            self._cache[cache_key] = result
        return result

    def decode_move(self, board, action_idx: int):
        import chess
        if USE_AZ1858:
            raise NotImplementedError("1858 decode_move not yet implemented")
        if not (0 <= action_idx < 4672):
            raise ValueError("action_idx out of range")
        from_sq, off = divmod(int(action_idx), 73)
        fr, ff = from_sq >> 3, from_sq & 7
        legal = [code_to_move(c) for c, _ in legal_moves_batch([board])[0]]

        def mk(dr: int, df: int, steps: int = 1, promo: Optional[int] = None):
            tr, tf = fr + dr * steps, ff + df * steps
            if not (0 <= tr < 8 and 0 <= tf < 8):
                return chess.Move.null()
            p = promo
            if promo is None:
                piece = board.piece_at(from_sq)
                if piece and piece.piece_type == chess.PAWN and (tr == 0 or tr == 7):
                    p = chess.QUEEN
            return chess.Move(from_sq, tr * 8 + tf, p)

        if off < 56:
            dr, df = RAY_DIRS[off // 7]
            mv = mk(dr, df, (off % 7) + 1)
        elif off < 64:
            dr, df = KNIGHT_DELTAS[off - 56]
            mv = mk(dr, df)
        else:
            u = off - 64
            dirs = ((1, 0), (1, -1), (1, 1)) if board.turn else ((-1, 0), (-1, 1), (-1, -1))
            dr, df = dirs[u % 3]
            mv = mk(dr, df, 1, UNDERPROMO_PIECES[u // 3])
        if mv in legal:
            return mv
        for lm in legal:
            if lm.from_square == from_sq and lm.to_square == mv.to_square:
                return lm
        for lm in legal:
            if lm.from_square == from_sq:
                return lm
        return chess.Move.null()

    def get_cache_stats(self) -> Dict[str, float]:
        total = self._cache_hits + self._cache_misses
        return {"hits": self._cache_hits, "misses": self._cache_misses, "total": total,
                "hit_rate": self._cache_hits / max(total, 1), "cache_size": len(self._cache)}

    def get_legal_actions(self, board) -> np.ndarray:
        pos = upload_positions([board])
        _, mask, _, _, _ = encode_positions_device(pos, False, True, False)
        return mask[0].cpu().numpy().astype(bool)

    def validate_encoding(self, board) -> bool:
        try:
            for code, idx in legal_moves_batch([board])[0]:
                m = code_to_move(code)
                m2 = self.decode_move(board, idx)
                if m.from_square != m2.from_square or m.to_square != m2.to_square:
                    return False
        except Exception:
            return False
        return True

    def get_action_statistics(self, board) -> Dict[str, float]:
        mask = self.get_legal_actions(board)
        return {"total_actions": LEGACY_POLICY_SIZE, "legal_actions": int(mask.sum()),
                "illegal_actions": int((~mask).sum()), "legal_ratio": float(mask.sum() / float(mask.size))}


move_encoder = MoveEncoder()


def _swap_pairs(pairs) -> np.ndarray:
    perm = np.arange(73, dtype=np.int64)
    for a, b in pairs:
        perm[a], perm[b] = b, a
    return np.ascontiguousarray(perm)


def build_horizontal_flip_permutation() -> np.ndarray:
    """``encoding.py:310-348``: mirror files -- E<->W, NE<->NW, SE<->SW, knight pairs, left/right underpromotions."""
    pairs = []
    for step in range(7):
        pairs += [(2 * 7 + step, 3 * 7 + step), (4 * 7 + step, 5 * 7 + step), (6 * 7 + step, 7 * 7 + step)]
    pairs += [(56 + o, 57 + o) for o in (0, 2, 4, 6)]
    pairs += [(b + 1, b + 2) for b in (64, 67, 70)]
    return _swap_pairs(pairs)


def build_rotate180_permutation() -> np.ndarray:
    """``encoding.py:351-386``: rotate by 180 degrees -- N<->S, E<->W, NE<->SW, NW<->SE, knight opposites."""
    pairs = []
    for step in range(7):
        pairs += [(0 * 7 + step, 1 * 7 + step), (2 * 7 + step, 3 * 7 + step),
                  (4 * 7 + step, 7 * 7 + step), (5 * 7 + step, 6 * 7 + step)]
    pairs += [(56, 63), (57, 62), (58, 61), (59, 60)]
    pairs += [(b + 1, b + 2) for b in (64, 67, 70)]
    return _swap_pairs(pairs)
