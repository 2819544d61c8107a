"""Oracle digests for BASELINE configs[1] at its full size (1 Mi synthetic positions).

Run in the build container:  python tests/golden/make_encode_digest.py   ->  tests/golden/encode_digest.json

The positions are the ones ``m0_random_playouts(seed=1234, max_plies=120)`` writes on the device (the bench's input).  They are
regenerated here on the CPU by the host build of the very same function (tests/hostcheck, ``random_playout_position`` in
csrc/chess_core.cuh); SHA-256 of the packed records per 4096-position chunk pins device == host for all 256 chunks.  The first
ORACLE_CHUNKS chunks (131,072 positions) then go through the ORACLE (oracle/encoding_ref.py on the oracle chess shim, itself pinned
against the unmodified azchess/encoding.py): SHA-256 per chunk of the float32 planes, the uint8 legal mask, the legal-move list in
python-chess order, the policy indices and the move counts.  The GPU test hashes the kernels' outputs the same way."""
from __future__ import annotations

import ctypes
import hashlib
import json
import multiprocessing as mp
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
SEED, MAX_PLIES, N_TOTAL, CHUNK, ORACLE_CHUNKS = 1234, 120, 1 << 20, 4096, 32


def hostcheck():
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "_hostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, os.path.join(d, "hostcheck.cpp")])
    return ctypes.CDLL(so)


def host_positions(hc, first, n):
    out = np.zeros((n, 9), dtype=np.uint64)
    hc.hc_random_playouts(out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), first, n, ctypes.c_uint64(SEED), MAX_PLIES)
    return out


def board_from_packed(w):
    """Packed position (csrc/chess_core.cuh: 8 bitboards + state word) -> oracle chess.Board."""
    from oracle import chess_shim  # noqa: F401
    import chess
    b = chess.Board(None)
    b.pawns, b.knights, b.bishops, b.rooks, b.queens, b.kings = (int(x) for x in w[:6])
    b.occupied_co[chess.WHITE], b.occupied_co[chess.BLACK] = int(w[6]), int(w[7])
    b.occupied = int(w[6]) | int(w[7])
    st = int(w[8])
    b.turn = bool(st & 1)
    cr = (st >> 1) & 15
    b.castling_rights = ((chess.BB_H1 if cr & 1 else 0) | (chess.BB_A1 if cr & 2 else 0) | (chess.BB_H8 if cr & 4 else 0) | (chess.BB_A8 if cr & 8 else 0))
    ep = (st >> 5) & 127
    b.ep_square = None if ep > 63 else ep
    b.halfmove_clock = (st >> 16) & 0xFFFF
    b.fullmove_number = (st >> 32) & 0xFFFF
    return b


def oracle_chunk(pos):
    from oracle import encoding_ref as E
    n = pos.shape[0]
    planes = np.zeros((n, 19, 8, 8), dtype=np.float32)
    mask = np.zeros((n, 4672), dtype=np.uint8)
    moves = np.zeros((n, 256), dtype=np.uint16)
    idx = np.zeros((n, 256), dtype=np.uint16)
    cnt = np.zeros((n,), dtype=np.int32)
    for i in range(n):
        b = board_from_packed(pos[i])
        planes[i] = E.encode_board(b)
        mask[i] = E.get_legal_actions(b)
        lm = E.legal_moves_and_indices(b)
        cnt[i] = len(lm)
        for k, (c, j) in enumerate(lm):
            moves[i, k], idx[i, k] = c, j
    return {k: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
            for k, a in (("planes", planes), ("mask", mask), ("moves", moves), ("idx", idx), ("counts", cnt))}


def main():
    hc = hostcheck()
    pos = host_positions(hc, 0, N_TOTAL)
    pos_sha = [hashlib.sha256(pos[c * CHUNK:(c + 1) * CHUNK].tobytes()).hexdigest() for c in range(N_TOTAL // CHUNK)]
    distinct = len(np.unique(pos.view([("", pos.dtype)] * 9)))
    print("positions:", N_TOTAL, "distinct:", distinct)
    with mp.Pool(os.cpu_count() or 1) as pool:
        digs = pool.map(oracle_chunk, [pos[c * CHUNK:(c + 1) * CHUNK] for c in range(ORACLE_CHUNKS)])
    out = {"seed": SEED, "max_plies": MAX_PLIES, "n_total": N_TOTAL, "chunk": CHUNK, "distinct_positions": distinct,
           "position_sha256": pos_sha, "oracle_chunks": ORACLE_CHUNKS,
           "oracle_sha256": {k: [d[k] for d in digs] for k in ("planes", "mask", "moves", "idx", "counts")}}
    json.dump(out, open(os.path.join(HERE, "encode_digest.json"), "w"))
    print("oracle digests for", ORACLE_CHUNKS * CHUNK, "positions written")


if __name__ == "__main__":
    main()
