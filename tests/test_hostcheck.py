"""CPU pre-flight of the PRODUCT's chess core: matrix0_b200/csrc/chess_core.cuh compiled with g++
(tests/hostcheck/hostcheck.cpp, a test harness, not a product path) against the oracle.  Lets the
GPU-less build container catch logic errors before GPU time is spent; the real parity tests are
the `-m gpu` ones that call the CUDA kernels through the C ABI."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import chess
from conftest import ROOT, random_playout_boards
from matrix0_b200.boards import board_to_raw, move_to_code
from oracle import encoding_ref as E

HC_DIR = os.path.join(ROOT, "tests", "hostcheck")
U64P = ctypes.POINTER(ctypes.c_uint64)


@pytest.fixture(scope="module")
def hc():
    # M0_HOSTCHECK_SANITIZE=1 (tests/hostcheck/run_sanitized.sh): the same harness under AddressSanitizer + UBSan -- the stand-in for
    # compute-sanitizer (closed on the GPU pool) as far as the code shared by host and device goes
    sanitize = os.environ.get("M0_HOSTCHECK_SANITIZE", "") == "1"
    so = os.path.join(HC_DIR, "_hostcheck_san.so" if sanitize else "_hostcheck.so")
    src = os.path.join(HC_DIR, "hostcheck.cpp")
    core = os.path.join(ROOT, "matrix0_b200", "csrc", "chess_core.cuh")
    ssl = os.path.join(ROOT, "matrix0_b200", "csrc", "ssl_core.cuh")
    mg = os.path.join(ROOT, "matrix0_b200", "csrc", "movegen_warp.cuh")
    sm = os.path.join(ROOT, "matrix0_b200", "csrc", "search_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in (src, core, ssl, mg, sm)):
        flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer"] if sanitize else ["-O2"]
        subprocess.check_call(["g++", *flags, "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, src])
    return ctypes.CDLL(so)


def pack(hc, b):
    raw = board_to_raw(b)
    out = np.zeros(9, dtype=np.uint64)
    misc = int(raw[9])
    ep = (misc >> 8) & 255
    hc.hc_pack(raw.ctypes.data_as(U64P), ctypes.c_uint64(int(raw[8])), misc & 1, -1 if ep > 63 else ep,
               (misc >> 16) & 0xFFFF, (misc >> 32) & 0xFFFF, out.ctypes.data_as(U64P))
    return out


def test_core_matches_oracle(hc):
    boards = random_playout_boards(40, 160, seed=5)
    boards += [chess.Board(f) for f in ["8/8/8/3pP3/8/8/8/8 w - d6 0 2", "8/8/8/8/k2Pp2Q/8/8/3K4 b - d3 0 1",
                                        "4k3/8/8/8/8/8/8/R3K2R w KQkq - 0 1",
                                        "R6R/3Q4/1Q4Q1/4Q3/2Q4Q/Q4Q2/pp1Q4/kBNN1KB1 w - - 0 1"]]
    keys = {}
    for b in boards:
        pos = pack(hc, b)
        mv = np.zeros(256, dtype=np.uint16)
        idx = np.zeros(256, dtype=np.int16)
        n = hc.hc_legal_moves(pos.ctypes.data_as(U64P), mv.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
                              idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)))
        exp = E.legal_moves_and_indices(b)
        assert n == len(exp), b.fen()
        assert mv[:n].tolist() == [c for c, _ in exp], b.fen()
        assert idx[:n].tolist() == [i for _, i in exp], b.fen()
        pl = np.zeros((19, 8, 8), dtype=np.float32)
        hc.hc_planes(pos.ctypes.data_as(U64P), pl.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        assert pl.tobytes() == E.encode_board(b).tobytes(), b.fen()
        assert bool(hc.hc_has_legal_ep(pos.ctypes.data_as(U64P))) == b.has_legal_en_passant()
        assert bool(hc.hc_insufficient(pos.ctypes.data_as(U64P))) == b.is_insufficient_material()
        legal = list(b.legal_moves)
        for m in legal[:: max(1, len(legal) // 4)]:
            out = np.zeros(9, dtype=np.uint64)
            fl = (ctypes.c_int * 2)()
            z, r = b.is_zeroing(m), b._reduces_castling_rights(m)
            hc.hc_push(pos.ctypes.data_as(U64P), ctypes.c_uint16(move_to_code(m)), out.ctypes.data_as(U64P), fl)
            b2 = b.copy()
            b2.push(m)
            assert (out == pack(hc, b2)).all(), (b.fen(), m.uci())
            assert (bool(fl[0]), bool(fl[1])) == (z, r)
            k = np.zeros(2, dtype=np.uint64)
            hc.hc_key(out.ctypes.data_as(U64P), k.ctypes.data_as(U64P))
            tk, kk = b2._transposition_key(), (int(k[0]), int(k[1]))
            assert keys.setdefault(tk, kk) == kk
    assert len(set(keys.values())) == len(keys)  # distinct transposition keys -> distinct hashes


def test_ssl_core_matches_reference_goldens(hc, golden_dir):
    """matrix0_b200/csrc/ssl_core.cuh (bitboard form of azchess/ssl_algorithms.py) on the host vs the golden maps generated from the
    unmodified reference module."""
    g = np.load(os.path.join(golden_dir, "ssl_golden.npz"))
    for i, fen in enumerate(g["fens"]):
        b = chess.Board(str(fen))
        pos = pack(hc, b)
        out = np.zeros((17, 8, 8), dtype=np.float32)
        hc.hc_ssl(pos.ctypes.data_as(U64P), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        assert np.array_equal(out[:13], g["piece"][i]), fen
        assert np.array_equal(out[13], g["threat"][i]), fen
        assert np.array_equal(out[14], g["pin"][i]), fen
        assert np.array_equal(out[15], g["fork"][i]), fen
        assert np.array_equal(out[16], g["control"][i]), fen


WEIRD_FENS = [
    "8/8/8/3pP3/8/8/8/8 w - d6 0 2",                                   # kingless, en passant
    "8/8/8/8/k2Pp2Q/8/8/3K4 b - d3 0 1",                               # en passant skewered along the rank
    "4k3/8/8/8/8/8/8/R3K2R w KQkq - 0 1",                              # dirty castling rights
    "R6R/3Q4/1Q4Q1/4Q3/2Q4Q/Q4Q2/pp1Q4/kBNN1KB1 w - - 0 1",            # 218 moves
    "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",   # Kiwipete
    "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
    "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",
    "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
    "4k3/8/8/8/8/8/4r3/R3K2R w KQ - 0 1",                              # castling while in check
    "4k3/8/8/8/8/8/5r2/R3K2R w KQ - 0 1",                              # castling through an attacked square
    "8/8/8/2k5/3Pp3/8/8/4K2B b - d3 0 1",                              # en passant removes the checker
    "8/8/3k4/8/3Pp3/8/8/3RK3 b - d3 0 1",                              # pinned en passant capturer
    "k7/8/8/8/8/8/8/K6K w - - 0 1", "kk6/8/8/8/8/8/8/KR5K w - - 0 1",  # several kings of one colour
    "4k3/P6P/8/8/8/8/p6p/4K3 w - - 0 1", "4k3/P6P/8/8/8/8/p6p/4K3 b - - 0 1",   # promotions
    "3rk3/8/8/8/8/8/3B4/3K4 w - - 0 1", "4k3/8/8/8/7b/8/5P2/4K3 w - - 0 1",     # pins
    "4k3/8/8/8/8/5n2/8/4K2r w - - 0 1",                                # double check
    "r3k2r/8/8/8/8/8/8/4K3 b kq - 0 1", "r3k2r/8/8/8/8/8/4R3/4K3 b kq - 0 1",   # black castling: free, in check
    "r3k2r/8/8/8/8/8/5R2/4K3 b kq - 0 1", "r3k2r/8/8/8/8/8/3R4/4K3 b kq - 0 1",   # ... through f8 / d8 attacked
    "r3k2r/8/8/8/8/8/1R6/4K3 b kq - 0 1", "4k3/1r6/8/8/8/8/8/R3K2R w KQ - 0 1",   # b-file attacked: long castling stays legal
    "r3k2r/8/8/8/8/8/6R1/4K3 b kq - 0 1", "4k3/8/8/8/8/8/8/RN2K1NR w KQ - 0 1",   # g8 attacked; pieces in the way
]


def test_legal_set_equals_ordered_generator(hc):
    """The per-piece SET interface of chess_core.cuh (the warp-cooperative mask path of the encode kernel) yields exactly
    the moves of the ordered generator: constructed edge positions, oracle playouts, and a 300 k-position C++ sweep."""
    boards = [chess.Board(f) for f in WEIRD_FENS] + random_playout_boards(30, 200, seed=77)
    for b in boards:
        assert hc.hc_legal_set_differs(pack(hc, b).ctypes.data_as(U64P)) == 0, b.fen()
    hc.hc_legal_set_sweep.restype = ctypes.c_long
    bad = ctypes.c_long(0)
    first = np.zeros(9, dtype=np.uint64)
    seen = 0
    for i, f in enumerate([chess.STARTING_FEN, WEIRD_FENS[4], WEIRD_FENS[6], WEIRD_FENS[7], WEIRD_FENS[14]]):
        seen += hc.hc_legal_set_sweep(pack(hc, chess.Board(f)).ctypes.data_as(U64P), ctypes.c_uint64(1000 + i), 600, 300,
                                      ctypes.byref(bad), first.ctypes.data_as(U64P))
        assert bad.value == 0, (f, [hex(int(x)) for x in first])
    assert seen > 300_000


def test_host_playouts_match_committed_position_digests(hc, golden_dir):
    """The host build of random_playout_position reproduces the committed digests of the first chunks (the GPU test checks all 256
    against the device kernel)."""
    import hashlib
    import json
    d = json.load(open(os.path.join(golden_dir, "encode_digest.json")))
    n = 4 * d["chunk"]
    out = np.zeros((n, 9), dtype=np.uint64)
    hc.hc_random_playouts(out.ctypes.data_as(U64P), 0, n, ctypes.c_uint64(d["seed"]), d["max_plies"])
    for c in range(4):
        assert hashlib.sha256(out[c * d["chunk"]:(c + 1) * d["chunk"]].tobytes()).hexdigest() == d["position_sha256"][c]


def test_warp_ordered_generator_matches_python_chess_order(hc, golden_dir):
    """csrc/movegen_warp.cuh (one own piece per lane, list placed by suffix sums) with its lanes simulated on the host: the SAME ordered
    list as the oracle's python-chess generation order -- random playouts, the constructed edge positions (checks, double checks, pins,
    en passant, promotions, castling), and 20,000 of the bench's synthetic positions against the single-thread generator."""
    boards = random_playout_boards(60, 200, seed=77) + [chess.Board(f) for f in WEIRD_FENS]
    u16p = ctypes.POINTER(ctypes.c_uint16)
    n_fallback = n_check = 0
    for b in boards:
        pos = pack(hc, b)
        mv = np.zeros(512, dtype=np.uint16)
        chk = ctypes.c_int(0)
        n = hc.hc_legal_moves_warp(pos.ctypes.data_as(U64P), mv.ctypes.data_as(u16p), ctypes.byref(chk))
        if n < 0:
            n_fallback += 1
            continue
        exp = [c for c, _ in E.legal_moves_and_indices(b)]
        assert n == len(exp) and mv[:n].tolist() == exp, b.fen()
        assert bool(chk.value) == b.is_check(), b.fen()
        n_check += b.is_check()
    assert n_check > 20 and n_fallback < len(WEIRD_FENS)
    N = 20000
    pos = np.zeros((N, 9), dtype=np.uint64)
    hc.hc_random_playouts(pos.ctypes.data_as(U64P), 0, N, ctypes.c_uint64(99), 160)
    a, bb, idx = np.zeros(512, dtype=np.uint16), np.zeros(512, dtype=np.uint16), np.zeros(512, dtype=np.int16)
    for i in range(N):
        chk = ctypes.c_int(0)
        n1 = hc.hc_legal_moves_warp(pos[i].ctypes.data_as(U64P), a.ctypes.data_as(u16p), ctypes.byref(chk))
        n2 = hc.hc_legal_moves(pos[i].ctypes.data_as(U64P), bb.ctypes.data_as(u16p), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)))
        assert n1 == n2 and (a[:n1] == bb[:n1]).all(), i


def test_search_math_is_numpy_and_python_arithmetic(hc):
    """csrc/search_math.cuh on the host (-ffp-contract=off, like the *_rn intrinsics on the device): numpy's pairwise float32 / float64
    `sum()` (mcts.py:206 `lp / lp.sum()`, :184 `dist.sum()`) bit for bit at every length the search meets (legal-move lists of 1 ... 300
    entries and a few longer ones, the whole 4,672-entry policy), the PUCT score and the repeated backup against the reference's Python expressions
    (mcts.py:878-881, :946-953), and Python's min / max clamp (:948) including its NaN behaviour."""
    import math
    import struct
    rng = np.random.default_rng(5)
    f32p, f64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    hc.hc_pairwise_f32.restype = ctypes.c_float
    hc.hc_pairwise_f64.restype = ctypes.c_double
    hc.hc_puct_score.restype = ctypes.c_double
    hc.hc_clip_unit.restype = ctypes.c_double
    bits32 = lambda x: struct.pack("<f", x)
    bits64 = lambda x: struct.pack("<d", x)
    lengths = list(range(1, 301)) + [511, 512, 513, 777, 900, 1024]
    for n in lengths:
        for scale in (1.0, 1e-3):
            a = (rng.random(n, dtype=np.float32) * np.float32(scale)).astype(np.float32)
            a[rng.integers(0, n)] *= np.float32(37.5)
            assert bits32(hc.hc_pairwise_f32(a.ctypes.data_as(f32p), n, 0)) == bits32(float(np.add.reduce(a))), n
            d = a.astype(np.float64) + rng.normal(0, 0.1, n)
            assert bits64(hc.hc_pairwise_f64(d.ctypes.data_as(f64p), n, 0)) == bits64(float(np.add.reduce(d))), n
    for n in (1023, 1025, 2048, 4672, 7000, 8192):
        a = rng.random(n, dtype=np.float32)
        assert bits32(hc.hc_pairwise_f32(a.ctypes.data_as(f32p), n, 1)) == bits32(float(a.sum())), n
        d = rng.normal(0, 1, n)
        assert bits64(hc.hc_pairwise_f64(d.ctypes.data_as(f64p), n, 1)) == bits64(float(d.sum())), n
    for _ in range(2000):
        q, cpuct, prior = float(rng.uniform(-1, 1)), float(rng.uniform(0.5, 4.0)), float(np.float32(rng.random()))
        parent, child_n = int(rng.integers(1, 100000)), int(rng.integers(0, 5000))
        sq = math.sqrt(parent)
        exp = q + cpuct * prior * (sq / (1.0 + child_n))                     # mcts.py:878-881
        got = hc.hc_puct_score(ctypes.c_double(q), ctypes.c_double(cpuct), ctypes.c_double(prior), ctypes.c_double(sq), child_n)
        assert bits64(got) == bits64(exp)
    for _ in range(300):
        n0, w0, v, times = int(rng.integers(0, 1000)), float(rng.uniform(-50, 50)), float(rng.uniform(-1, 1)), int(rng.integers(1, 97))
        n, w, q = ctypes.c_int(n0), ctypes.c_double(w0), ctypes.c_double(0.0)
        hc.hc_backup_repeated(ctypes.byref(n), ctypes.byref(w), ctypes.byref(q), ctypes.c_double(v), times)
        en, ew = n0, w0
        for _k in range(times):                                              # mcts.py:946-953, one sample at a time
            en += 1
            ew += v
            eq = ew / en
        assert n.value == en and bits64(w.value) == bits64(ew) and bits64(q.value) == bits64(eq)
    for x in (0.3, -0.3, 1.0, -1.0, 1.5, -7.0, float("inf"), float("-inf"), float("nan"), -0.0):
        exp = max(-1.0, min(1.0, x))                                         # mcts.py:948
        got = hc.hc_clip_unit(ctypes.c_double(x))
        assert (math.isnan(exp) and math.isnan(got)) or bits64(got) == bits64(exp), x
