// Self-supervised target maps of azchess/ssl_algorithms.py (create_enhanced_ssl_targets, :502-535) as bitboard logic.
//
// Everything here lives in PLANE coordinates, the layout of encode_board (encoding.py:40-46): bit r*8 + c with row r = 7 - rank and
// column c = file, i.e. the byte-swapped python-chess bitboard.  The reference works on the encoded planes with shifted copies
// (`_shift`, :70-80) and per-square float counters; here a shift is one 64-bit shift plus a column mask and a counter is five
// bit-sliced bitboards.  Reference behaviour that looks odd on a chess board is kept as is (oracle/ssl_ref.py lists it): white
// pawns "attack towards increasing row", and the pin map is identically zero.
#pragma once
#include "chess_core.cuh"

namespace m0 {

struct SslMasks {
  u64 piece[13];   // 12 piece planes (white P N B R Q K, black P N B R Q K) + empty squares   (_create_piece_targets, :537-557)
  u64 threat;      // squares attacked by the side not to move                                (detect_threats_batch, :51-143)
  u64 fork;        // own N/B/R/Q/K squares attacking >= 2 enemy pieces                       (detect_forks_batch, :348-421)
  u64 ctrl_pos;    // more white than black attackers  (+1)                                   (calculate_square_control_batch, :423-500)
  u64 ctrl_neg;    // more black than white attackers  (-1)
};                 // pin map (detect_pins_batch, :256-346) is all zero

// _shift(mask, dr, dc): new[r][c] = old[r - dr][c - dc], nothing wraps around (:70-80)
M0_HD u64 ssl_shift(u64 b, int dr, int dc) {
  const int s = dr * 8 + dc;
  u64 r = s >= 0 ? (s < 64 ? b << s : 0) : (-s < 64 ? b >> (-s) : 0);
  const u64 col = 0x0101010101010101ull;
  u64 bad = 0;
  if (dc > 0) for (int k = 0; k < dc; ++k) bad |= col << k;          // columns 0 .. dc-1 can only hold wrapped bits
  if (dc < 0) for (int k = 0; k < -dc; ++k) bad |= col << (7 - k);   // columns 8+dc .. 7
  return r & ~bad;
}

struct SslCounter {   // per-square counter 0..31 as bit slices
  u64 b[5];
  M0_HD void clear() { for (int i = 0; i < 5; ++i) b[i] = 0; }
  M0_HD void add(u64 x) {
    for (int i = 0; i < 5; ++i) {
      const u64 carry = b[i] & x;
      b[i] ^= x;
      x = carry;
    }
  }
  M0_HD u64 any() const { return b[0] | b[1] | b[2] | b[3] | b[4]; }
  M0_HD u64 at_least_two() const { return b[1] | b[2] | b[3] | b[4]; }
};

M0_HD u64 ssl_plane_bb(u64 bb) {   // python-chess bitboard (a1 = bit 0) -> plane coordinates
#ifdef __CUDA_ARCH__
  return __byte_perm((unsigned)(bb >> 32), 0, 0x0123) | ((u64)__byte_perm((unsigned)bb, 0, 0x0123) << 32);
#else
  return __builtin_bswap64(bb);
#endif
}

// blocking-aware ray accumulation (:82-93, :482-492): every step of every direction adds the shifted frontier
M0_HD void ssl_add_rays(SslCounter& cnt, u64 src, const int (*dirs)[2], int ndirs, u64 occ) {
  for (int d = 0; d < ndirs; ++d) {
    u64 f = src;
    for (int step = 1; step < 8; ++step) {
      f = ssl_shift(f, dirs[d][0], dirs[d][1]);
      cnt.add(f);
      f &= ~occ;
      if (!f) break;   // nothing left to propagate: the remaining steps add zero
    }
  }
}

M0_HD void ssl_attack_counts(const u64* w, const u64* b, u64 occ, SslCounter& wa, SslCounter& ba) {
  const int KN[8][2] = {{-2, -1}, {-2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2}, {2, -1}, {2, 1}};
  const int KG[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
  const int DG[4][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
  const int OR[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
  wa.clear();
  ba.clear();
  wa.add(ssl_shift(w[0], 1, -1)); wa.add(ssl_shift(w[0], 1, 1));      // :116-118 (towards increasing row)
  ba.add(ssl_shift(b[0], -1, -1)); ba.add(ssl_shift(b[0], -1, 1));    // :128-130
  for (int d = 0; d < 8; ++d) { wa.add(ssl_shift(w[1], KN[d][0], KN[d][1])); ba.add(ssl_shift(b[1], KN[d][0], KN[d][1])); }
  for (int d = 0; d < 8; ++d) { wa.add(ssl_shift(w[5], KG[d][0], KG[d][1])); ba.add(ssl_shift(b[5], KG[d][0], KG[d][1])); }
  ssl_add_rays(wa, w[2] | w[4], DG, 4, occ); ssl_add_rays(wa, w[3] | w[4], OR, 4, occ);
  ssl_add_rays(ba, b[2] | b[4], DG, 4, occ); ssl_add_rays(ba, b[3] | b[4], OR, 4, occ);
}

// own sliders hitting an enemy piece as the FIRST occupied square of a ray (:400-412)
M0_HD void ssl_add_sliding_hits(SslCounter& cnt, u64 origins, const int (*dirs)[2], int ndirs, u64 enemy, u64 occ) {
  if (!origins) return;
  for (int d = 0; d < ndirs; ++d) {
    u64 blocked = 0;
    for (int s = 1; s < 8; ++s) {
      const u64 back = ssl_shift(enemy, -dirs[d][0] * s, -dirs[d][1] * s);
      cnt.add(origins & back & ~blocked);
      blocked |= ssl_shift(occ, -dirs[d][0] * s, -dirs[d][1] * s);
    }
  }
}

M0_HD void ssl_masks(const Position& p, SslMasks& m) {
  const u64 pt[6] = {p.pawns, p.knights, p.bishops, p.rooks, p.queens, p.kings};
  u64 w[6], b[6];
  u64 occ = 0;
  for (int i = 0; i < 6; ++i) {
    w[i] = ssl_plane_bb(pt[i] & p.occ_w);
    b[i] = ssl_plane_bb(pt[i] & p.occ_b);
    m.piece[i] = w[i];
    m.piece[6 + i] = b[i];
    occ |= w[i] | b[i];
  }
  m.piece[12] = ~occ;
  SslCounter wa, ba;
  ssl_attack_counts(w, b, occ, wa, ba);
  const bool stm_white = pos_turn(p) != 0;
  m.threat = stm_white ? ba.any() : wa.any();
  // sign(white - black): compare the bit-sliced counters from the top bit down
  u64 gt = 0, lt = 0;
  for (int i = 4; i >= 0; --i) {
    const u64 undecided = ~(gt | lt);
    gt |= undecided & wa.b[i] & ~ba.b[i];
    lt |= undecided & ~wa.b[i] & ba.b[i];
  }
  m.ctrl_pos = gt;
  m.ctrl_neg = lt;
  // forks of the side to move
  const u64* own = stm_white ? w : b;
  u64 enemy = 0;
  for (int i = 0; i < 6; ++i) enemy |= stm_white ? b[i] : w[i];
  const int KN[8][2] = {{-2, -1}, {-2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2}, {2, -1}, {2, 1}};
  const int KG[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
  const int DG[4][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
  const int OR[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
  SslCounter fc;
  fc.clear();
  for (int d = 0; d < 8; ++d) fc.add(own[1] & ssl_shift(enemy, -KN[d][0], -KN[d][1]));
  for (int d = 0; d < 8; ++d) fc.add(own[5] & ssl_shift(enemy, -KG[d][0], -KG[d][1]));
  ssl_add_sliding_hits(fc, own[2], DG, 4, enemy, occ);
  ssl_add_sliding_hits(fc, own[3], OR, 4, enemy, occ);
  ssl_add_sliding_hits(fc, own[4], DG, 4, enemy, occ);
  ssl_add_sliding_hits(fc, own[4], OR, 4, enemy, occ);
  m.fork = fc.at_least_two() & (own[1] | own[2] | own[3] | own[4] | own[5]);
}

}  // namespace m0
