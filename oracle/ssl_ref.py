"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of ``azchess/ssl_algorithms.py`` for ONE position at a time, which is how
``selfplay_worker`` calls it (``create_enhanced_ssl_targets`` on a batch of one, internal.py:460-466).

Pinned by tests/test_oracle_ssl.py against tests/golden/ssl_golden.npz (generated from the unmodified reference module by
tests/golden/make_ssl_golden.py).  All maps live in PLANE coordinates (row = 7 - rank, col = file, encoding.py:40-46); the
reference's quirks are kept as they are:
  * "white pawns attack towards increasing row" (ssl_algorithms.py:116-118, 448-451), i.e. towards rank 1 on the real board;
  * the pin map is identically zero: `is_own_piece` and `has_slider_beyond` (:320-340) are one-hot masks of two DIFFERENT squares,
    so their elementwise AND is empty (confirmed on constructed pin positions in the golden file).
"""
from __future__ import annotations

import numpy as np

KNIGHT = [(-2, -1), (-2, 1), (-1, -2), (-1, 2), (1, -2), (1, 2), (2, -1), (2, 1)]
KING = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
DIAG = [(-1, -1), (-1, 1), (1, -1), (1, 1)]
ORTHO = [(-1, 0), (1, 0), (0, -1), (0, 1)]


def _shift(m: np.ndarray, dr: int, dc: int) -> np.ndarray:
    """ssl_algorithms.py:70-80: roll and clear what wrapped around."""
    s = np.zeros_like(m)
    r0, r1 = max(0, dr), min(8, 8 + dr)
    c0, c1 = max(0, dc), min(8, 8 + dc)
    if r0 < r1 and c0 < c1:
        s[r0:r1, c0:c1] = m[r0 - dr:r1 - dr, c0 - dc:c1 - dc]
    return s


def _rays(src: np.ndarray, dirs, occ: np.ndarray) -> np.ndarray:
    """ssl_algorithms.py:82-93 / :482-492: blocking-aware ray accumulation."""
    att = np.zeros((8, 8), dtype=np.float32)
    for dr, dc in dirs:
        f = src.astype(np.float32).copy()
        for _ in range(1, 8):
            f = _shift(f, dr, dc)
            att += f
            f = f * (~occ).astype(np.float32)
    return att


def _attack_counts(p: np.ndarray):
    """white / black attack-count maps shared by detect_threats_batch (:95-137) and calculate_square_control_batch (:440-495)."""
    occ = p[:12].sum(axis=0) > 0
    w, b = p[0:6], p[6:12]
    wa = _shift(w[0], 1, -1) + _shift(w[0], 1, 1)
    ba = _shift(b[0], -1, -1) + _shift(b[0], -1, 1)
    for d in KNIGHT:
        wa = wa + _shift(w[1], *d)
        ba = ba + _shift(b[1], *d)
    for d in KING:
        wa = wa + _shift(w[5], *d)
        ba = ba + _shift(b[5], *d)
    wa = wa + _rays(w[2] + w[4], DIAG, occ) + _rays(w[3] + w[4], ORTHO, occ)
    ba = ba + _rays(b[2] + b[4], DIAG, occ) + _rays(b[3] + b[4], ORTHO, occ)
    return wa.astype(np.float32), ba.astype(np.float32)


def piece_targets(planes: np.ndarray) -> np.ndarray:
    """_create_piece_targets (:537-557): 12 piece planes + empty squares, 13 x 8 x 8."""
    out = np.zeros((13, 8, 8), dtype=np.float32)
    out[:12] = (planes[:12] > 0).astype(np.float32)
    out[12] = (planes[:12].sum(axis=0) == 0).astype(np.float32)
    return out


def threat_target(planes: np.ndarray) -> np.ndarray:
    """detect_threats_batch (:51-143): squares attacked by the side NOT to move, clamped to {0, 1}."""
    wa, ba = _attack_counts(planes)
    stm_white = planes[12, 0, 0] > 0.5
    return np.clip(ba if stm_white else wa, 0.0, 1.0).astype(np.float32)


def pin_target(planes: np.ndarray) -> np.ndarray:
    """detect_pins_batch (:256-346): see the module docstring -- always zero."""
    return np.zeros((8, 8), dtype=np.float32)


def fork_target(planes: np.ndarray) -> np.ndarray:
    """detect_forks_batch (:348-421): own N/B/R/Q/K squares that attack >= 2 enemy pieces (first blocker on a ray counts if enemy)."""
    p = planes[:12].astype(np.float32)
    occ = p.sum(axis=0) > 0
    stm_white = planes[12, 0, 0] > 0.5
    own, enemy = (p[0:6], p[6:12]) if stm_white else (p[6:12], p[0:6])
    enemy_any = enemy.sum(axis=0) > 0
    count = np.zeros((8, 8), dtype=np.float32)
    for dr, dc in KNIGHT:
        count += ((own[1] > 0) & (_shift(enemy_any.astype(np.float32), -dr, -dc) > 0)).astype(np.float32)
    for dr, dc in KING:
        count += ((own[5] > 0) & (_shift(enemy_any.astype(np.float32), -dr, -dc) > 0)).astype(np.float32)

    def slide(origins, dirs):
        nonlocal count
        ob = origins > 0
        for dr, dc in dirs:
            for s in range(1, 8):
                back = _shift(enemy_any.astype(np.float32), -dr * s, -dc * s) > 0
                blocked = np.zeros((8, 8), dtype=bool)
                for t in range(1, s):
                    blocked |= _shift(occ.astype(np.float32), -dr * t, -dc * t) > 0
                count += (ob & back & ~blocked).astype(np.float32)
    slide(own[2], DIAG)
    slide(own[3], ORTHO)
    slide(own[4], DIAG)
    slide(own[4], ORTHO)
    tactical = (own[1] + own[2] + own[3] + own[4] + own[5]) > 0
    return ((count >= 2.0) & tactical).astype(np.float32)


def control_target(planes: np.ndarray) -> np.ndarray:
    """calculate_square_control_batch (:423-500): sign(white attackers - black attackers)."""
    wa, ba = _attack_counts(planes)
    return np.sign(wa - ba).astype(np.float32)


def ssl_targets(planes: np.ndarray):
    """create_enhanced_ssl_targets (:502-535) for one position: dict of float32 arrays."""
    return {"piece": piece_targets(planes), "threat": threat_target(planes), "pin": pin_target(planes), "fork": fork_target(planes),
            "control": control_target(planes)}
