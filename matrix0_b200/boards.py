"""Host-side marshalling of python-chess style boards into the engine's packed positions.

A board is anything exposing python-chess's public ``Board`` attributes (``pawns``, ``knights``,
``bishops``, ``rooks``, ``queens``, ``kings``, ``occupied_co``, ``turn``, ``castling_rights``,
``ep_square``, ``halfmove_clock``, ``fullmove_number``).  The raw record mirrors those fields one
to one; cleaning of castling rights (``Board.clean_castling_rights``) and packing happen on the
device in ``m0_positions_pack``.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

RAW_WORDS = 10
POSITION_WORDS = 9
MAX_MOVES = 256
POLICY_SIZE = 4672
PLANES = 19


def board_to_raw(board, out: np.ndarray | None = None) -> np.ndarray:
    """One board -> uint64[10] raw record (see csrc/encode_kernels.cu: pack_positions_kernel)."""
    if out is None:
        out = np.empty(RAW_WORDS, dtype=np.uint64)
    ep = board.ep_square
    misc = (1 if board.turn else 0) | ((255 if ep is None else int(ep)) << 8) \
        | (min(int(board.halfmove_clock), 0xFFFF) << 16) | (min(int(board.fullmove_number), 0xFFFF) << 32)
    out[0] = board.pawns
    out[1] = board.knights
    out[2] = board.bishops
    out[3] = board.rooks
    out[4] = board.queens
    out[5] = board.kings
    out[6] = board.occupied_co[True]
    out[7] = board.occupied_co[False]
    out[8] = int(board.castling_rights) & 0xFFFFFFFFFFFFFFFF
    out[9] = misc
    return out


def boards_to_raw(boards: Sequence) -> np.ndarray:
    raw = np.empty((len(boards), RAW_WORDS), dtype=np.uint64)
    for i, b in enumerate(boards):
        board_to_raw(b, raw[i])
    return raw


def move_to_code(move) -> int:
    """python-chess Move -> packed u16 (from | to<<6 | promotion<<12)."""
    return int(move.from_square) | (int(move.to_square) << 6) | ((int(move.promotion) if move.promotion else 0) << 12)


def code_to_move(code: int):
    """Packed u16 -> chess.Move (needs an importable ``chess`` module, as the reference does)."""
    import chess
    promo = (code >> 12) & 7
    return chess.Move(code & 63, (code >> 6) & 63, promo if promo else None)
