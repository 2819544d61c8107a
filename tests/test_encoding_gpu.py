"""GPU parity (bit-exact): CUDA encode / legal-move / mask kernels through the C ABI vs the oracle
and the committed golden vectors produced by the unmodified reference."""
import os

import numpy as np
import pytest

import chess
from conftest import random_playout_boards
from oracle import encoding_ref as E

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def enc():
    from matrix0_b200 import encoding
    return encoding


def test_golden_vectors(enc, golden_dir):
    import torch
    g = np.load(os.path.join(golden_dir, "encoding_golden.npz"))
    raw = torch.from_numpy(g["raw"].view(np.int64)).cuda()
    from matrix0_b200 import _native
    pos = torch.empty((raw.shape[0], 9), dtype=torch.int64, device="cuda")
    _native.check(_native.lib().m0_positions_pack(raw.data_ptr(), raw.shape[0], pos.data_ptr(), _native.current_stream()))
    planes, mask, moves, idx, cnt = enc.encode_positions_device(pos, True, True, True)
    planes = planes.cpu().numpy()
    n = raw.shape[0]
    bits = np.packbits(planes[:, :12].astype(np.uint8).reshape(n, -1), axis=1)
    assert (bits == g["piece_bits"]).all()
    assert set(np.unique(planes[:, :12])) <= {0.0, 1.0}
    assert planes[:, 12:].tobytes() == np.broadcast_to(g["const_planes"][:, :, None, None], (n, 7, 8, 8)).astype(np.float32).tobytes()
    cnt = cnt.cpu().numpy()
    assert (cnt == g["counts"]).all()
    moves = moves.cpu().numpy().view(np.uint16)
    idx = idx.cpu().numpy().view(np.uint16)
    mask = mask.cpu().numpy()
    for i in range(n):
        k = int(cnt[i])
        assert (moves[i, :k] == g["moves"][i, :k]).all(), g["fens"][i]
        assert (idx[i, :k] == g["idx"][i, :k]).all(), g["fens"][i]
        exp = np.zeros(4672, dtype=np.uint8)
        exp[g["idx"][i, :k]] = 1
        assert (mask[i] == exp).all(), g["fens"][i]


def test_random_playouts_vs_oracle(enc):
    boards = random_playout_boards(25, 200, seed=77)
    planes, mask = enc.encode_boards(boards)
    mi = enc.legal_moves_batch(boards)
    for i, b in enumerate(boards):
        assert planes[i].tobytes() == E.encode_board(b).tobytes(), b.fen()
        assert (mask[i] == E.get_legal_actions(b)).all(), b.fen()
        assert mi[i] == E.legal_moves_and_indices(b), b.fen()


@pytest.mark.parametrize("n", [0, 1, 31, 127, 128, 129, 1000])
def test_ragged_batch_sizes(enc, n):
    boards = random_playout_boards(8, 200, seed=3)[:n]
    if n > len(boards):
        boards = (boards * (n // max(1, len(boards)) + 1))[:n]
    planes, mask = enc.encode_boards(boards)
    assert planes.shape == (n, 19, 8, 8) and mask.shape == (n, 4672)
    for i in (0, n // 2, n - 1) if n else ():
        assert planes[i].tobytes() == E.encode_board(boards[i]).tobytes()
        assert (mask[i] == E.get_legal_actions(boards[i])).all()


def test_reference_api_semantics(enc):
    b = chess.Board()
    assert enc.encode_board(b).shape == (19, 8, 8)
    m = enc.move_encoder.get_legal_actions(b)
    assert m.dtype == bool and m.shape == (4672,) and m.sum() == 20          # ref tests/test_encoding.py:73-84
    with pytest.raises(ValueError):                                            # ref tests/test_encoding.py:120-130
        enc.move_to_index(b, chess.Move.from_uci("a1a8"))
    with pytest.raises(ValueError):
        enc.move_encoder.encode_move(b, chess.Move.from_uci("a1a8"))
    with pytest.raises(ValueError):
        enc.encode_board(b, planes=18)
    kiwi = chess.Board("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
    for mv in kiwi.legal_moves:                                                # ref tests/test_encoding.py:64-71
        a = enc.move_encoder.encode_move(kiwi, mv)
        assert a == E.move_to_index(kiwi, mv)
        assert enc.move_encoder.decode_move(kiwi, a) == mv
    assert enc.move_encoder.validate_encoding(kiwi)
    ep = chess.Board("8/8/8/3pP3/8/8/8/8 w - d6 0 2")                          # kingless board accepted
    assert 0 <= enc.move_to_index(ep, chess.Move.from_uci("e5d6")) < 4672
    hp = enc.build_horizontal_flip_permutation()
    assert len(hp) == 73 and (hp[hp] == np.arange(73)).all()                   # ref tests/test_encoding_random.py:20-27
    rp = enc.build_rotate180_permutation()
    assert (rp[rp] == np.arange(73)).all()


def test_full_size_properties(enc):
    """Config-2 size (1M positions): size-independent invariants instead of an oracle sweep."""
    import torch
    from matrix0_b200.boards import boards_to_raw
    from matrix0_b200 import _native
    boards = random_playout_boards(60, 200, seed=123)
    raw = torch.from_numpy(boards_to_raw(boards).view(np.int64)).cuda()
    reps = (1 << 20) // raw.shape[0] + 1
    raw = raw.repeat(reps, 1)[: 1 << 20].contiguous()
    n = raw.shape[0]
    pos = torch.empty((n, 9), dtype=torch.int64, device="cuda")
    _native.check(_native.lib().m0_positions_pack(raw.data_ptr(), n, pos.data_ptr(), _native.current_stream()))
    planes, mask, moves, idx, cnt = enc.encode_positions_device(pos, True, True, True)
    # mask popcount == number of legal moves (policy indices are unique per position)
    assert torch.equal(mask.sum(dim=1, dtype=torch.int32), cnt)
    # piece planes sum == popcount of the occupancy
    occ = (pos[:, 6] | pos[:, 7])
    pc = torch.zeros(n, dtype=torch.int32, device="cuda")
    for bit in range(64):
        pc += ((occ >> bit) & 1).to(torch.int32)
    assert torch.equal(planes[:, :12].sum(dim=(1, 2, 3)).to(torch.int32), pc)
    # every replica of the same position produced identical rows
    k = len(boards)
    assert torch.equal(planes[:k], planes[k:2 * k]) and torch.equal(mask[:k], mask[-(n % k or k) - k:-(n % k or k)] if False else mask[k:2 * k])
    # idempotence: a second launch gives the same bytes
    planes2, mask2, _, _, _ = enc.encode_positions_device(pos, True, True, False)
    assert torch.equal(planes, planes2) and torch.equal(mask, mask2)


def test_ssl_targets_match_reference_goldens(golden_dir):
    """m0_ssl_targets (azchess/ssl_algorithms.py create_enhanced_ssl_targets as a bitboard kernel) vs the maps generated from the
    unmodified reference module, bit-exact; ragged batch sizes."""
    import os
    import chess
    from matrix0_b200.encoding import create_enhanced_ssl_targets
    g = np.load(os.path.join(golden_dir, "ssl_golden.npz"))
    boards = [chess.Board(str(f)) for f in g["fens"]]
    for n in (len(boards), 1, 129, 0):
        t = create_enhanced_ssl_targets(boards[:n])
        for k in ("piece", "threat", "pin", "fork", "control"):
            assert t[k].dtype == np.float32
            assert np.array_equal(t[k], g[k][:n]), (k, n)


def test_set_kernel_edge_positions(enc):
    """The half-warp-per-position kernel (legal moves as a per-piece set, danger map, constant-mask castling) on constructed
    edge positions -- kingless, several kings of one colour, castling in / through check, pinned and check-removing en passant,
    promotions, double check, 218 moves -- against the oracle AND against the ordered thread-per-position kernel."""
    import torch
    from test_hostcheck import WEIRD_FENS
    boards = [chess.Board(f) for f in WEIRD_FENS]
    planes, mask = enc.encode_boards(boards)                                  # planes + mask only: the set kernel
    for i, b in enumerate(boards):
        assert planes[i].tobytes() == E.encode_board(b).tobytes(), b.fen()
        assert (mask[i] == E.get_legal_actions(b)).all(), b.fen()
    pos = enc.upload_positions(boards)
    p2, m2, _, _, cnt = enc.encode_positions_device(pos, True, True, True)    # with the lists: the ordered kernel
    assert np.array_equal(p2.cpu().numpy(), planes) and np.array_equal(m2.cpu().numpy().astype(bool), mask.astype(bool))
    assert torch.equal(m2.sum(dim=1, dtype=torch.int32), cnt)
    m3 = enc.encode_positions_device(pos, False, True, False)[1]              # mask alone
    assert torch.equal(m3, m2)


def test_config2_full_size_against_oracle_digests(enc, golden_dir):
    """BASELINE configs[1] as written: the 1 Mi positions of the bench (m0_random_playouts(seed 1234, <= 120 plies); 1,019,010 of them
    distinct).  (a) the device playouts equal the host build of the same function on all 256 chunks (SHA-256 of the packed records);
    (b) the half-warp set kernel (planes + mask, the benchmarked one) and the ordered thread-per-position kernel agree byte for byte on
    all 1 Mi positions; (c) planes, mask, ordered move list, policy indices and counts of the first 131,072 positions hash to the digests
    the ORACLE produced in the build container (tests/golden/make_encode_digest.py)."""
    import hashlib
    import json
    import torch
    from matrix0_b200 import _native
    d = json.load(open(os.path.join(golden_dir, "encode_digest.json")))
    n, chunk = d["n_total"], d["chunk"]
    pos = torch.empty((n, 9), dtype=torch.int64, device="cuda")
    _native.check(_native.lib().m0_random_playouts(pos.data_ptr(), n, d["seed"], d["max_plies"], _native.current_stream()), "m0_random_playouts")
    ph = pos.cpu().numpy()
    for c in range(n // chunk):
        assert hashlib.sha256(ph[c * chunk:(c + 1) * chunk].tobytes()).hexdigest() == d["position_sha256"][c], c
    assert len(np.unique(ph.view([("", ph.dtype)] * 9))) == d["distinct_positions"] >= 1_000_000
    planes, mask, _, _, _ = enc.encode_positions_device(pos, True, True, False)            # the set kernel (the bench's kernel)
    no = d["oracle_chunks"] * chunk
    for lo in range(0, n, 1 << 18):                                                        # ordered kernel in slices of 256 Ki
        hi = min(n, lo + (1 << 18))
        p2, m2, mv, ix, cnt = enc.encode_positions_device(pos[lo:hi], True, True, True)
        assert torch.equal(p2, planes[lo:hi]) and torch.equal(m2, mask[lo:hi]), lo
        assert torch.equal(m2.sum(dim=1, dtype=torch.int32), cnt)
        if lo < no:
            k = min(hi, no) - lo
            ar = torch.arange(256, device="cuda")[None, :] < cnt[:k, None]                 # entries past the count are unspecified
            mvh = torch.where(ar, mv[:k], torch.zeros_like(mv[:k])).cpu().numpy().view(np.uint16)
            ixh = torch.where(ar, ix[:k], torch.zeros_like(ix[:k])).cpu().numpy().view(np.uint16)
            pl, mk, cn = planes[lo:lo + k].cpu().numpy(), mask[lo:lo + k].cpu().numpy(), cnt[:k].cpu().numpy()
            for c in range(k // chunk):
                gc = lo // chunk + c
                s = slice(c * chunk, (c + 1) * chunk)
                for name, arr in (("planes", pl[s]), ("mask", mk[s]), ("moves", mvh[s]), ("idx", ixh[s]), ("counts", cn[s])):
                    assert hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest() == d["oracle_sha256"][name][gc], (name, gc)
