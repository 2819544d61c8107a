"""TEST INFRASTRUCTURE ONLY -- minimal ``chess.pgn`` restatement (python-chess, un-vendored).

Covers what ``azchess/selfplay/internal.py:36-63`` uses (``read_game``, ``Game.board``,
``Game.mainline_moves``) and what the oracle pinning tests need to replay the reference's
``data/eval_games/*.pgn`` by SAN.  No variations, NAGs or comments beyond skipping them.
"""
from __future__ import annotations

import re
from typing import Dict, Iterator, List, Optional

import chess

_TAG = re.compile(r'^\[([A-Za-z0-9_]+)\s+"(.*)"\]\s*$')
_RESULTS = {"1-0", "0-1", "1/2-1/2", "*"}


class Game:
    def __init__(self) -> None:
        self.headers: Dict[str, str] = {}
        self.san_moves: List[str] = []
        self.result_token: Optional[str] = None

    def board(self) -> chess.Board:
        fen = self.headers.get("FEN")
        return chess.Board(fen) if fen else chess.Board()

    def mainline_sans(self) -> List[str]:
        return list(self.san_moves)

    def mainline_moves(self) -> Iterator[chess.Move]:
        board = self.board()
        for san in self.san_moves:
            move = board.parse_san(san)
            board.push(move)
            yield move


def _tokenize(movetext: str) -> List[str]:
    movetext = re.sub(r"\{[^}]*\}", " ", movetext)
    movetext = re.sub(r";[^\n]*", " ", movetext)
    depth = 0
    out = []
    for ch in movetext:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth = max(0, depth - 1)
        elif depth == 0:
            out.append(ch)
    return "".join(out).split()


def read_game(handle) -> Optional[Game]:
    game = Game()
    found = False
    movetext: List[str] = []
    in_moves = False
    while True:
        pos = handle.tell() if hasattr(handle, "tell") else None
        line = handle.readline()
        if not line:
            break
        s = line.strip()
        if not s:
            if in_moves:
                break
            continue
        if s.startswith("%"):
            continue
        m = _TAG.match(s)
        if m and not in_moves:
            game.headers[m.group(1)] = m.group(2)
            found = True
            continue
        if m and in_moves:
            if pos is not None:
                handle.seek(pos)
            break
        in_moves = True
        found = True
        movetext.append(s)
    if not found:
        return None
    for tok in _tokenize(" ".join(movetext)):
        if tok in _RESULTS:
            game.result_token = tok
            break
        tok = re.sub(r"^\d+\.+", "", tok)
        if not tok or tok.startswith("$"):
            continue
        tok = tok.rstrip("!?")
        if tok:
            game.san_moves.append(tok)
    return game
