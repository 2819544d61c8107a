// Bitboard chess core for the B200 self-play engine: position representation, legal-move
// generation in python-chess's generation ORDER, make-move, 4672-way policy index, the
// transposition key hash and the game-end predicates.  Everything is a single-thread inline
// function so that one lane of a warp-per-game kernel (or one thread of the thread-per-position
// encode kernel) can run it; the kernels around it do the coalesced I/O.
//
// Behavioural contract (what the reference gets from python-chess + azchess/encoding.py):
//   * move order      = chess.Board.generate_legal_moves -- this is the
//                       Node.children order of azchess/mcts.py:140,214-223 and decides every
//                       PUCT tie-break (mcts.py:901).
//   * policy index    = azchess/encoding.py:113-150 (move_to_index), layout from_sq*73 + off.
//   * planes          = azchess/encoding.py:11-46 (encode_board).
//   * transposition   = chess.Board._transposition_key (mcts.py:342,919,1343).
//   * terminal value  = azchess/mcts.py:1223-1229 on top of Board.is_game_over().
//
// The code is plain C++ guarded by M0_HD so that tests/hostcheck can compile the very same
// functions with g++ and compare them with the oracle on the CPU-only build container; that host
// build is a test harness only and is never loaded by the product.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define M0_HD __host__ __device__ __forceinline__
#define M0_HD_NOINLINE __host__ __device__ __noinline__
#else
#define M0_HD inline
#define M0_HD_NOINLINE
#endif

namespace m0 {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

// ---- constants ---------------------------------------------------------------------------------
static constexpr u64 BB_ALL = 0xFFFFFFFFFFFFFFFFull;
static constexpr u64 FILE_A = 0x0101010101010101ull;
static constexpr u64 FILE_B = FILE_A << 1;
static constexpr u64 FILE_G = FILE_A << 6;
static constexpr u64 FILE_H = FILE_A << 7;
static constexpr u64 RANK_1 = 0xFFull;
static constexpr u64 RANK_3 = 0xFFull << 16;
static constexpr u64 RANK_4 = 0xFFull << 24;
static constexpr u64 RANK_5 = 0xFFull << 32;
static constexpr u64 RANK_6 = 0xFFull << 40;
static constexpr u64 RANK_8 = 0xFFull << 56;
static constexpr u64 DIAG_MAIN = 0x8040201008040201ull;  // a1-h8
static constexpr u64 DIAG_ANTI = 0x0102040810204080ull;  // h1-a8
static constexpr u64 DARK_SQUARES = 0xAA55AA55AA55AA55ull;
static constexpr u64 LIGHT_SQUARES = 0x55AA55AA55AA55AAull;

enum { SQ_A1 = 0, SQ_C1 = 2, SQ_D1 = 3, SQ_E1 = 4, SQ_F1 = 5, SQ_G1 = 6, SQ_H1 = 7,
       SQ_A8 = 56, SQ_C8 = 58, SQ_D8 = 59, SQ_E8 = 60, SQ_F8 = 61, SQ_G8 = 62, SQ_H8 = 63 };
enum { PT_NONE = 0, PT_PAWN = 1, PT_KNIGHT = 2, PT_BISHOP = 3, PT_ROOK = 4, PT_QUEEN = 5, PT_KING = 6 };
// castling bits inside Position::state
enum { CR_WK = 1, CR_WQ = 2, CR_BK = 4, CR_BQ = 8 };
static constexpr int EP_NONE = 64;
static constexpr int MAX_MOVES = 256;
static constexpr int POLICY_SIZE = 4672;

// A move is packed as from | to<<6 | promotion<<12 (promotion = python-chess piece type, 0 = none).
typedef u16 Move;
M0_HD Move make_move(int from, int to, int promo = 0) { return (Move)(from | (to << 6) | (promo << 12)); }
M0_HD int move_from(Move m) { return m & 63; }
M0_HD int move_to(Move m) { return (m >> 6) & 63; }
M0_HD int move_promo(Move m) { return (m >> 12) & 7; }

// ---- position ----------------------------------------------------------------------------------
// state bits: [0] turn (1 = white)  [1..4] clean castling rights (WK,WQ,BK,BQ)
//             [5..11] ep square (0..63, 64 = none; set after EVERY double push like python-chess)
//             [16..31] halfmove clock  [32..47] fullmove number
struct Position {
  u64 pawns, knights, bishops, rooks, queens, kings, occ_w, occ_b;
  u64 state;
};
static constexpr int POSITION_WORDS = 9;

M0_HD int pos_turn(const Position& p) { return (int)(p.state & 1); }
M0_HD int pos_castling(const Position& p) { return (int)((p.state >> 1) & 15); }
M0_HD int pos_ep(const Position& p) { return (int)((p.state >> 5) & 127); }
M0_HD int pos_halfmove(const Position& p) { return (int)((p.state >> 16) & 0xFFFF); }
M0_HD int pos_fullmove(const Position& p) { return (int)((p.state >> 32) & 0xFFFF); }
M0_HD u64 pack_state(int turn, int castling, int ep, int halfmove, int fullmove) {
  return (u64)(turn & 1) | ((u64)(castling & 15) << 1) | ((u64)(ep & 127) << 5) |
         ((u64)(halfmove & 0xFFFF) << 16) | ((u64)(fullmove & 0xFFFF) << 32);
}
M0_HD u64 pos_occ(const Position& p) { return p.occ_w | p.occ_b; }
M0_HD u64 pos_us(const Position& p) { return pos_turn(p) ? p.occ_w : p.occ_b; }
M0_HD u64 pos_them(const Position& p) { return pos_turn(p) ? p.occ_b : p.occ_w; }

// ---- bit helpers -------------------------------------------------------------------------------
M0_HD int popcnt(u64 b) {
#if defined(__CUDA_ARCH__)
  return __popcll(b);
#else
  return __builtin_popcountll(b);
#endif
}
M0_HD int msb(u64 b) {  // b != 0
#if defined(__CUDA_ARCH__)
  return 63 - __clzll((long long)b);
#else
  return 63 - __builtin_clzll(b);
#endif
}
M0_HD int lsb(u64 b) {  // b != 0
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)b) - 1;
#else
  return __builtin_ctzll(b);
#endif
}
M0_HD u64 brev64(u64 b) {
#if defined(__CUDA_ARCH__)
  return __brevll(b);
#else
  b = __builtin_bswap64(b);
  b = ((b >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((b & 0x0F0F0F0F0F0F0F0Full) << 4);
  b = ((b >> 2) & 0x3333333333333333ull) | ((b & 0x3333333333333333ull) << 2);
  b = ((b >> 1) & 0x5555555555555555ull) | ((b & 0x5555555555555555ull) << 1);
  return b;
#endif
}
M0_HD u64 sq_bb(int sq) { return 1ull << sq; }

// ---- line masks and attacks (no lookup tables: arithmetic masks + hyperbola quintessence) -------
M0_HD u64 rank_mask(int sq) { return RANK_1 << (sq & 56); }
M0_HD u64 file_mask(int sq) { return FILE_A << (sq & 7); }
M0_HD u64 diag_mask(int sq) {
  int d = (sq & 7) - (sq >> 3);
  return d >= 0 ? (DIAG_MAIN >> (8 * d)) : (DIAG_MAIN << (8 * -d));
}
M0_HD u64 anti_mask(int sq) {
  int k = (sq & 7) + (sq >> 3);
  return k <= 7 ? (DIAG_ANTI >> (8 * (7 - k))) : (DIAG_ANTI << (8 * (k - 7)));
}
// sliding attacks along one line (mask includes the slider square) for a given occupancy
M0_HD u64 line_attacks(u64 occ, u64 mask, u64 s) {
  u64 m = mask & ~s;
  u64 o = occ & m;
  u64 f = o - s;
  u64 r = brev64(brev64(o) - brev64(s));
  return (f ^ r) & m;
}
M0_HD u64 rank_attacks(int sq, u64 occ) { return line_attacks(occ, rank_mask(sq), sq_bb(sq)); }
M0_HD u64 file_attacks(int sq, u64 occ) { return line_attacks(occ, file_mask(sq), sq_bb(sq)); }
M0_HD u64 diag_attacks(int sq, u64 occ) {
  u64 s = sq_bb(sq);
  return line_attacks(occ, diag_mask(sq), s) | line_attacks(occ, anti_mask(sq), s);
}
// rays on an empty board: line_attacks(0, mask, s) == mask & ~s
M0_HD u64 rook_rays_empty(int sq) { return (rank_mask(sq) | file_mask(sq)) & ~sq_bb(sq); }
M0_HD u64 bishop_rays_empty(int sq) { return (diag_mask(sq) | anti_mask(sq)) & ~sq_bb(sq); }
M0_HD u64 knight_attacks_bb(u64 b) {
  return ((b << 17) & ~FILE_A) | ((b << 15) & ~FILE_H) | ((b << 10) & ~(FILE_A | FILE_B)) |
         ((b << 6) & ~(FILE_G | FILE_H)) | ((b >> 17) & ~FILE_H) | ((b >> 15) & ~FILE_A) |
         ((b >> 10) & ~(FILE_G | FILE_H)) | ((b >> 6) & ~(FILE_A | FILE_B));
}
M0_HD u64 king_attacks_bb(u64 b) {
  return ((b << 1) & ~FILE_A) | ((b >> 1) & ~FILE_H) | (b << 8) | (b >> 8) | ((b << 9) & ~FILE_A) |
         ((b << 7) & ~FILE_H) | ((b >> 7) & ~FILE_A) | ((b >> 9) & ~FILE_H);
}
// squares attacked by a pawn of `color` (1 = white) standing on b
M0_HD u64 pawn_attacks_bb(int color, u64 b) {
  return color ? (((b << 7) & ~FILE_H) | ((b << 9) & ~FILE_A)) : (((b >> 7) & ~FILE_A) | ((b >> 9) & ~FILE_H));
}
// full line through a and b (edge to edge, both squares included) or 0 when not aligned: chess.ray
M0_HD u64 ray_through(int a, int b) {
  if (a == b) return 0;
  u64 bb = sq_bb(b);
  if (diag_mask(a) & bb) return diag_mask(a);
  if (anti_mask(a) & bb) return anti_mask(a);
  if (rank_mask(a) & bb) return rank_mask(a);
  if (file_mask(a) & bb) return file_mask(a);
  return 0;
}
// squares strictly between a and b when aligned: chess.between
M0_HD u64 between_bb(int a, int b) {
  u64 bb = ray_through(a, b) & ((BB_ALL << a) ^ (BB_ALL << b));
  return bb & (bb - 1);
}

M0_HD int piece_type_at(const Position& p, int sq) {
  u64 m = sq_bb(sq);
  if (!(pos_occ(p) & m)) return PT_NONE;
  if (p.pawns & m) return PT_PAWN;
  if (p.knights & m) return PT_KNIGHT;
  if (p.bishops & m) return PT_BISHOP;
  if (p.rooks & m) return PT_ROOK;
  if (p.queens & m) return PT_QUEEN;
  return PT_KING;
}

// chess.Board.attacks_mask
M0_HD u64 attacks_from(const Position& p, int sq) {
  u64 s = sq_bb(sq);
  if (s & p.pawns) return pawn_attacks_bb((s & p.occ_w) ? 1 : 0, s);
  if (s & p.knights) return knight_attacks_bb(s);
  if (s & p.kings) return king_attacks_bb(s);
  u64 occ = pos_occ(p), a = 0;
  if (s & (p.bishops | p.queens)) a = diag_attacks(sq, occ);
  if (s & (p.rooks | p.queens)) a |= rank_attacks(sq, occ) | file_attacks(sq, occ);
  return a;
}
// chess.Board._attackers_mask(color, square, occupied)
M0_HD u64 attackers_of(const Position& p, int color, int sq, u64 occ) {
  u64 s = sq_bb(sq);
  u64 qr = p.queens | p.rooks, qb = p.queens | p.bishops;
  u64 a = (king_attacks_bb(s) & p.kings) | (knight_attacks_bb(s) & p.knights) |
          ((rank_attacks(sq, occ) | file_attacks(sq, occ)) & qr) | (diag_attacks(sq, occ) & qb) |
          (pawn_attacks_bb(!color, s) & p.pawns);
  return a & (color ? p.occ_w : p.occ_b);
}

// chess.Board.clean_castling_rights() for a raw rook-square mask (standard chess), as CR_* bits
M0_HD int clean_castling_bits(const Position& p, u64 raw_rights) {
  u64 c = raw_rights & p.rooks;
  u64 w = c & RANK_1 & p.occ_w & (sq_bb(SQ_A1) | sq_bb(SQ_H1));
  u64 b = c & RANK_8 & p.occ_b & (sq_bb(SQ_A8) | sq_bb(SQ_H8));
  if (!(p.occ_w & p.kings & sq_bb(SQ_E1))) w = 0;
  if (!(p.occ_b & p.kings & sq_bb(SQ_E8))) b = 0;
  int bits = 0;
  if (w & sq_bb(SQ_H1)) bits |= CR_WK;
  if (w & sq_bb(SQ_A1)) bits |= CR_WQ;
  if (b & sq_bb(SQ_H8)) bits |= CR_BK;
  if (b & sq_bb(SQ_A8)) bits |= CR_BQ;
  return bits;
}
M0_HD u64 castling_rook_mask(int bits) {
  return ((bits & CR_WK) ? sq_bb(SQ_H1) : 0) | ((bits & CR_WQ) ? sq_bb(SQ_A1) : 0) |
         ((bits & CR_BK) ? sq_bb(SQ_H8) : 0) | ((bits & CR_BQ) ? sq_bb(SQ_A8) : 0);
}

// ---- legal move generation in python-chess order ------------------------------------------------
struct MoveGenCtx {
  const Position* p;
  int us;        // 1 = white
  u64 occ, ours, theirs;
  int king;      // msb(kings & ours) or -1
  u64 blockers;  // _slider_blockers(king)
  Move* out;
  int n;
};

// chess.Board._slider_blockers
M0_HD u64 slider_blockers(const Position& p, int king, u64 occ, u64 ours, u64 theirs) {
  u64 rq = p.rooks | p.queens, bq = p.bishops | p.queens;
  u64 snipers = (rook_rays_empty(king) & rq) | (bishop_rays_empty(king) & bq);
  u64 blockers = 0;
  u64 s = snipers & theirs;
  while (s) {
    int sn = msb(s);
    s ^= sq_bb(sn);
    u64 b = between_bb(king, sn) & occ;
    if (b && (b & (b - 1)) == 0) blockers |= b;
  }
  return blockers & ours;
}

// chess.Board.pin_mask(turn, square) given the king square
M0_HD u64 pin_mask_for(const Position& p, int king, int sq, u64 occ, u64 theirs) {
  u64 sm = sq_bb(sq);
  u64 rq = p.rooks | p.queens, bq = p.bishops | p.queens;
  for (int k = 0; k < 3; ++k) {
    u64 rays = k == 0 ? (file_mask(king) & ~sq_bb(king)) : (k == 1 ? (rank_mask(king) & ~sq_bb(king)) : bishop_rays_empty(king));
    if (rays & sm) {
      u64 snipers = rays & (k == 2 ? bq : rq) & theirs;
      while (snipers) {
        int sn = msb(snipers);
        snipers ^= sq_bb(sn);
        if ((between_bb(sn, king) & (occ | sm)) == sm) return ray_through(king, sn);
      }
      break;
    }
  }
  return BB_ALL;
}

// chess.Board._ep_skewered
M0_HD bool ep_skewered(const Position& p, int us, int king, int capturer, int ep, u64 occ, u64 theirs) {
  int last_double = ep + (us ? -8 : 8);
  u64 occupancy = (occ & ~sq_bb(last_double) & ~sq_bb(capturer)) | sq_bb(ep);
  if (rank_attacks(king, occupancy) & theirs & (p.rooks | p.queens)) return true;
  if (diag_attacks(king, occupancy) & theirs & (p.bishops | p.queens)) return true;
  return false;
}

// chess.Board._is_safe
M0_HD bool move_is_safe(const MoveGenCtx& c, int from, int to, bool is_castle, bool is_ep) {
  const Position& p = *c.p;
  if (from == c.king) {
    if (is_castle) return true;
    return attackers_of(p, !c.us, to, c.occ) == 0;
  }
  if (is_ep) {
    return (pin_mask_for(p, c.king, from, c.occ, c.theirs) & sq_bb(to)) != 0 &&
           !ep_skewered(p, c.us, c.king, from, to, c.occ, c.theirs);
  }
  return !(c.blockers & sq_bb(from)) || (ray_through(from, to) & sq_bb(c.king)) != 0;
}

M0_HD void emit(MoveGenCtx& c, int from, int to, int promo, bool is_castle, bool is_ep) {
  if (c.king >= 0 && !move_is_safe(c, from, to, is_castle, is_ep)) return;
  if (c.n < MAX_MOVES) c.out[c.n] = make_move(from, to, promo);
  c.n++;
}
M0_HD void emit_pawn(MoveGenCtx& c, int from, int to) {
  int r = to >> 3;
  if (r == 0 || r == 7) {
    emit(c, from, to, PT_QUEEN, false, false);
    emit(c, from, to, PT_ROOK, false, false);
    emit(c, from, to, PT_BISHOP, false, false);
    emit(c, from, to, PT_KNIGHT, false, false);
  } else {
    emit(c, from, to, 0, false, false);
  }
}

// chess.Board.generate_pseudo_legal_ep
M0_HD void gen_pseudo_ep(MoveGenCtx& c, u64 from_mask, u64 to_mask) {
  const Position& p = *c.p;
  int ep = pos_ep(p);
  // python: `if not self.ep_square` -- square 0 (a1) is falsy too, harmless for real ep squares
  if (ep == EP_NONE || ep == 0 || !(sq_bb(ep) & to_mask)) return;
  if (sq_bb(ep) & c.occ) return;
  u64 capturers = p.pawns & c.ours & from_mask & pawn_attacks_bb(!c.us, sq_bb(ep)) & (c.us ? RANK_5 : RANK_4);
  while (capturers) {
    int f = msb(capturers);
    capturers ^= sq_bb(f);
    emit(c, f, ep, 0, false, true);
  }
}

M0_HD bool attacked_for_king(const Position& p, int us, u64 path, u64 occ) {
  while (path) {
    int s = msb(path);
    path ^= sq_bb(s);
    if (attackers_of(p, !us, s, occ)) return true;
  }
  return false;
}

// chess.Board.generate_castling_moves (standard chess; king e1g1 / e1c1), h-side before a-side.
// Writes the king square and up to two destination squares; returns the number of castling moves.
// Note for legal_castling(): with the king on its e-file home square, the rooks in the corners (the only rights
// clean_castling_bits keeps) and the side NOT in check, the per-square tests below -- which lift the king, and for the
// destination also move the rook -- see the same attackers as the plain attack map under the full occupancy: the only
// line through two back-rank squares is the back rank, an attacker on it beyond the king would give check, the corners
// have nothing behind them, and the rook's new square shields what the king shielded.
M0_HD int castling_moves(const Position& p, int us, u64 occ, u64 ours, u64 from_mask, u64 to_mask, int* ksq_out, int* to_out) {
  u64 backrank = us ? RANK_1 : RANK_8;
  u64 king = ours & p.kings & backrank & from_mask;
  king &= (0 - king);
  if (!king) return 0;
  int ksq = msb(king);
  *ksq_out = ksq;
  u64 bb_c = (FILE_A << 2) & backrank, bb_d = (FILE_A << 3) & backrank;
  u64 bb_f = (FILE_A << 5) & backrank, bb_g = (FILE_A << 6) & backrank;
  u64 cand = castling_rook_mask(pos_castling(p)) & backrank & to_mask;
  int n = 0;
  while (cand) {
    int rs = msb(cand);
    cand ^= sq_bb(rs);
    u64 rook = sq_bb(rs);
    bool a_side = rook < king;
    u64 king_to = a_side ? bb_c : bb_g;
    u64 rook_to = a_side ? bb_d : bb_f;
    u64 king_path = between_bb(ksq, msb(king_to));
    u64 rook_path = between_bb(rs, msb(rook_to));
    bool blocked = ((occ ^ king ^ rook) & (king_path | rook_path | king_to | rook_to)) != 0;
    bool attacked;
    if (blocked) attacked = true;
    else attacked = attacked_for_king(p, us, king_path | king, occ ^ king) ||
                    attacked_for_king(p, us, king_to, occ ^ king ^ rook ^ rook_to);
    if (!attacked) {
      // _from_chess960: e1->h1 becomes e1g1, e1->a1 becomes e1c1 (king on e-file in standard chess)
      int to = rs;
      if (ksq == SQ_E1 && rs == SQ_H1) to = SQ_G1;
      else if (ksq == SQ_E1 && rs == SQ_A1) to = SQ_C1;
      else if (ksq == SQ_E8 && rs == SQ_H8) to = SQ_G8;
      else if (ksq == SQ_E8 && rs == SQ_A8) to = SQ_C8;
      if (n < 2) to_out[n] = to;
      n++;
    }
  }
  return n < 2 ? n : 2;
}
M0_HD void gen_castling(MoveGenCtx& c, u64 from_mask, u64 to_mask) {
  int ksq = 0, to[2];
  int n = castling_moves(*c.p, c.us, c.occ, c.ours, from_mask, to_mask, &ksq, to);
  for (int i = 0; i < n; ++i) emit(c, ksq, to[i], 0, true, false);
}

// chess.Board.generate_pseudo_legal_moves with the _is_safe filter applied at emission
M0_HD void gen_pseudo(MoveGenCtx& c, u64 from_mask, u64 to_mask) {
  const Position& p = *c.p;
  // 1. non-pawn pieces, from-squares high to low, targets high to low
  u64 non_pawns = c.ours & ~p.pawns & from_mask;
  while (non_pawns) {
    int f = msb(non_pawns);
    non_pawns ^= sq_bb(f);
    u64 t = attacks_from(p, f) & ~c.ours & to_mask;
    while (t) {
      int to = msb(t);
      t ^= sq_bb(to);
      emit(c, f, to, 0, false, false);
    }
  }
  // 2. castling
  if (from_mask & p.kings) gen_castling(c, from_mask, to_mask);
  // 3-6. pawns
  u64 pawns = p.pawns & c.ours & from_mask;
  if (!pawns) return;
  u64 cap = pawns;
  while (cap) {
    int f = msb(cap);
    cap ^= sq_bb(f);
    u64 t = pawn_attacks_bb(c.us, sq_bb(f)) & c.theirs & to_mask;
    while (t) {
      int to = msb(t);
      t ^= sq_bb(to);
      emit_pawn(c, f, to);
    }
  }
  u64 single, dbl;
  if (c.us) {
    single = (pawns << 8) & ~c.occ;
    dbl = (single << 8) & ~c.occ & (RANK_3 | RANK_4);
  } else {
    single = (pawns >> 8) & ~c.occ;
    dbl = (single >> 8) & ~c.occ & (RANK_6 | RANK_5);
  }
  single &= to_mask;
  dbl &= to_mask;
  while (single) {
    int to = msb(single);
    single ^= sq_bb(to);
    emit_pawn(c, to + (c.us ? -8 : 8), to);
  }
  while (dbl) {
    int to = msb(dbl);
    dbl ^= sq_bb(to);
    emit(c, to + (c.us ? -16 : 16), to, 0, false, false);
  }
  if (pos_ep(p) != EP_NONE && pos_ep(p) != 0) gen_pseudo_ep(c, from_mask, to_mask);
}

// chess.Board.generate_legal_moves.  Returns the number of legal moves (may exceed MAX_MOVES only
// for absurd kingless boards; entries past MAX_MOVES are dropped).  *checkers_out gets the checkers.
M0_HD int generate_legal_moves(const Position& p, Move* out, u64* checkers_out = nullptr) {
  MoveGenCtx c;
  c.p = &p;
  c.us = pos_turn(p);
  c.occ = pos_occ(p);
  c.ours = pos_us(p);
  c.theirs = pos_them(p);
  c.out = out;
  c.n = 0;
  u64 king_mask = p.kings & c.ours;
  u64 checkers = 0;
  if (king_mask) {
    c.king = msb(king_mask);
    c.blockers = slider_blockers(p, c.king, c.occ, c.ours, c.theirs);
    checkers = attackers_of(p, !c.us, c.king, c.occ);
    if (checkers) {
      // chess.Board._generate_evasions
      u64 sliders = checkers & (p.bishops | p.rooks | p.queens);
      u64 attacked = 0;
      while (sliders) {
        int ch = msb(sliders);
        sliders ^= sq_bb(ch);
        attacked |= ray_through(c.king, ch) & ~sq_bb(ch);
      }
      u64 t = king_attacks_bb(sq_bb(c.king)) & ~c.ours & ~attacked;
      while (t) {
        int to = msb(t);
        t ^= sq_bb(to);
        emit(c, c.king, to, 0, false, false);
      }
      int checker = msb(checkers);
      if (sq_bb(checker) == checkers) {
        u64 target = between_bb(c.king, checker) | checkers;
        gen_pseudo(c, ~p.kings, target);
        int ep = pos_ep(p);
        if (ep != EP_NONE && ep != 0 && !(sq_bb(ep) & target)) {
          int last_double = ep + (c.us ? -8 : 8);
          if (last_double == checker) gen_pseudo_ep(c, BB_ALL, BB_ALL);
        }
      }
    } else {
      gen_pseudo(c, BB_ALL, BB_ALL);
    }
  } else {
    c.king = -1;
    c.blockers = 0;
    gen_pseudo(c, BB_ALL, BB_ALL);
  }
  if (checkers_out) *checkers_out = checkers;
  return c.n;
}

// ---- the legal moves as a SET (no order): the same moves as generate_legal_moves, organised per piece so that
// the lanes of a warp can each take one piece of a position (encode_kernels.cu, legal-mask path).  The ordered
// generator above stays the definition; tests/hostcheck and the GPU tests compare the two on every position.
struct LegalCtx {
  int us, king;             // king = msb(kings & ours) or -1
  u64 occ, ours, theirs;
  u64 blockers;             // _slider_blockers(king)
  u64 checkers;
  u64 to_mask;              // targets of non-king pieces: all squares, or between(king, checker) | checker in single check
  u64 king_cand;            // king destinations BEFORE the attacked-square test (evasions exclude the checking rays)
  bool others_move;         // false in double check
  bool ep_open;             // an en-passant capture is generated at all (ep square set, empty, and admitted by the evasion rule)
  int ep;
};
// `checkers_in` (optional): the opponent's pieces that attack the king, when the caller already has them
M0_HD LegalCtx make_legal_ctx(const Position& p, const u64* checkers_in = nullptr) {
  LegalCtx c;
  c.us = pos_turn(p);
  c.occ = pos_occ(p);
  c.ours = pos_us(p);
  c.theirs = pos_them(p);
  c.blockers = 0;
  c.checkers = 0;
  c.to_mask = BB_ALL;
  c.king_cand = 0;
  c.others_move = true;
  c.ep = pos_ep(p);
  c.ep_open = c.ep != EP_NONE && c.ep != 0 && !(sq_bb(c.ep) & c.occ);
  u64 king_mask = p.kings & c.ours;
  c.king = king_mask ? msb(king_mask) : -1;
  if (c.king < 0) return c;
  c.blockers = slider_blockers(p, c.king, c.occ, c.ours, c.theirs);
  c.checkers = checkers_in ? *checkers_in : attackers_of(p, !c.us, c.king, c.occ);
  c.king_cand = king_attacks_bb(sq_bb(c.king)) & ~c.ours;
  if (c.checkers) {
    u64 sliders = c.checkers & (p.bishops | p.rooks | p.queens);
    while (sliders) {
      int ch = msb(sliders);
      sliders ^= sq_bb(ch);
      c.king_cand &= ~(ray_through(c.king, ch) & ~sq_bb(ch));
    }
    if (c.checkers & (c.checkers - 1)) {
      c.others_move = false;
      c.to_mask = 0;
      c.ep_open = false;
    } else {
      int checker = msb(c.checkers);
      c.to_mask = between_bb(c.king, checker) | c.checkers;
      if (c.ep_open && !(sq_bb(c.ep) & c.to_mask) && (c.ep + (c.us ? -8 : 8)) != checker) c.ep_open = false;
    }
  }
  return c;
}
// attack set of the piece on `sq` and whether it reaches the king: OR-ed over the opponent's pieces these give the
// danger map (every square the opponent attacks under the full occupancy) and the checkers
M0_HD u64 enemy_attacks(const Position& p, int sq, u64 king_bb, u64* checkers) {
  u64 a = attacks_from(p, sq);
  if (a & king_bb) *checkers |= sq_bb(sq);
  return a;
}
// attacks_from with the knight / king steps read from two 64-entry tables (filled with knight_attacks_bb / king_attacks_bb)
M0_HD u64 attacks_from_lut(const Position& p, int sq, const u64* knight_tab, const u64* king_tab) {
  u64 s = sq_bb(sq);
  if (s & p.pawns) return pawn_attacks_bb((s & p.occ_w) ? 1 : 0, s);
  if (s & p.knights) return knight_tab[sq];
  if (s & p.kings) return king_tab[sq];
  u64 occ = pos_occ(p), a = 0;
  if (s & (p.bishops | p.queens)) a = diag_attacks(sq, occ);
  if (s & (p.rooks | p.queens)) a |= rank_attacks(sq, occ) | file_attacks(sq, occ);
  return a;
}
// THE king may step to `to` (a bit of c.king_cand)
M0_HD bool king_step_safe(const Position& p, const LegalCtx& c, int to) { return attackers_of(p, !c.us, to, c.occ) == 0; }
// Destinations of the own piece on `from`, without castling, en passant and the steps of THE king (c.king).
// For a pawn the set holds captures and pushes; a destination on the first / last rank stands for four promotions.
// `attacks` = attacks_from(p, from); not read for a pawn
M0_HD u64 piece_targets_given(const Position& p, const LegalCtx& c, int from, u64 attacks) {
  u64 fb = sq_bb(from);
  if (from == c.king || !c.others_move) return 0;
  if (c.checkers && (p.kings & fb)) return 0;          // evasions are generated with from_mask = ~kings
  u64 t;
  if (p.pawns & fb) {
    u64 single = (c.us ? (fb << 8) : (fb >> 8)) & ~c.occ;
    u64 dbl = (c.us ? (single << 8) : (single >> 8)) & ~c.occ & (c.us ? (RANK_3 | RANK_4) : (RANK_6 | RANK_5));
    t = ((pawn_attacks_bb(c.us, fb) & c.theirs) | single | dbl) & c.to_mask;
  } else {
    t = attacks & ~c.ours & c.to_mask;
  }
  if (c.king >= 0 && (c.blockers & fb)) t &= ray_through(from, c.king);   // _is_safe for a pinned piece
  return t;
}
M0_HD u64 piece_targets(const Position& p, const LegalCtx& c, int from) {
  return piece_targets_given(p, c, from, (p.pawns & sq_bb(from)) ? 0 : attacks_from(p, from));
}
// the pawn on `from` may capture en passant
M0_HD bool ep_capture_legal(const Position& p, const LegalCtx& c, int from) {
  if (!c.ep_open) return false;
  u64 capturers = p.pawns & c.ours & pawn_attacks_bb(!c.us, sq_bb(c.ep)) & (c.us ? RANK_5 : RANK_4);
  if (!(capturers & sq_bb(from))) return false;
  if (c.king < 0) return true;
  return (pin_mask_for(p, c.king, from, c.occ, c.theirs) & sq_bb(c.ep)) != 0 &&
         !ep_skewered(p, c.us, c.king, from, c.ep, c.occ, c.theirs);
}
// castling moves of the position (king square, up to two destinations); _is_safe passes them for THE king and applies
// the pinned-piece rule to any other own king (boards with several kings)
// castling moves of the position through castling_moves() with its per-square attack tests
M0_HD int legal_castling_exact(const Position& p, const LegalCtx& c, int* ksq_out, int* to_out) {
  if (c.checkers) return 0;
  int to[2];
  int n = castling_moves(p, c.us, c.occ, c.ours, BB_ALL, BB_ALL, ksq_out, to), m = 0;
  for (int i = 0; i < n; ++i) {
    int ksq = *ksq_out;
    bool ok = c.king < 0 || ksq == c.king || !(c.blockers & sq_bb(ksq)) || (ray_through(ksq, to[i]) & sq_bb(c.king)) != 0;
    if (ok) to_out[m++] = to[i];
  }
  return m;
}
// With the danger map (squares the opponent attacks under the full occupancy), a single own king on its home square and
// the side not in check, castling_moves() reduces to constant masks: f/g (b/c/d) empty, e/f/g (c/d/e) not attacked -- see
// the note at castling_moves() for why lifting the king and moving the rook does not change the attackers.
// Returns -1 when the board is not of that kind (several own kings, king elsewhere): use legal_castling_exact then.
M0_HD int legal_castling_fast(const Position& p, const LegalCtx& c, u64 danger, int* ksq_out, int* to_out) {
  if (c.checkers) return 0;
  const u64 own_kings = p.kings & c.ours;
  if (!own_kings || (own_kings & (own_kings - 1)) || c.king != (c.us ? SQ_E1 : SQ_E8)) return -1;
  const int bits = pos_castling(p), sh = c.us ? 0 : 56;
  int m = 0;
  *ksq_out = c.king;
  if ((bits & (c.us ? CR_WK : CR_BK)) && !(c.occ & (0x60ull << sh)) && !(danger & (0x70ull << sh))) to_out[m++] = c.king + 2;
  if ((bits & (c.us ? CR_WQ : CR_BQ)) && !(c.occ & (0x0Eull << sh)) && !(danger & (0x1Cull << sh))) to_out[m++] = c.king - 2;
  return m;
}
M0_HD int legal_castling(const Position& p, const LegalCtx& c, int* ksq_out, int* to_out, const u64* danger = nullptr) {
  int m = danger ? legal_castling_fast(p, c, *danger, ksq_out, to_out) : -1;
  return m >= 0 ? m : legal_castling_exact(p, c, ksq_out, to_out);
}

// chess.Board.has_legal_en_passant (used by the transposition key and is_irreversible):
// any(generate_legal_ep()) = some pseudo-legal ep capture that is not is_into_check().
M0_HD bool has_legal_ep(const Position& p) {
  int ep = pos_ep(p);
  if (ep == EP_NONE || ep == 0) return false;
  int us = pos_turn(p);
  u64 occ = pos_occ(p), ours = pos_us(p), theirs = pos_them(p);
  if (sq_bb(ep) & occ) return false;
  u64 capturers = p.pawns & ours & pawn_attacks_bb(!us, sq_bb(ep)) & (us ? RANK_5 : RANK_4);
  if (!capturers) return false;
  u64 king_mask = p.kings & ours;
  if (!king_mask) return true;  // is_into_check() is False without a king
  int king = msb(king_mask);
  u64 checkers = attackers_of(p, !us, king, occ);
  if (checkers) {
    // the capture must be one of _generate_evasions(): single checker, and it either blocks /
    // captures on the target mask or removes the checking pawn that just double-pushed
    if (checkers & (checkers - 1)) return false;
    int checker = msb(checkers);
    u64 target = between_bb(king, checker) | checkers;
    int last_double = ep + (us ? -8 : 8);
    if (!(sq_bb(ep) & target) && last_double != checker) return false;
  }
  while (capturers) {
    int f = msb(capturers);
    capturers ^= sq_bb(f);
    if ((pin_mask_for(p, king, f, occ, theirs) & sq_bb(ep)) && !ep_skewered(p, us, king, f, ep, occ, theirs)) return true;
  }
  return false;
}

// ---- make move: chess.Board.push ----------------------------------------------------------------
M0_HD void remove_piece(Position& p, int sq) {
  u64 m = ~sq_bb(sq);
  p.pawns &= m; p.knights &= m; p.bishops &= m; p.rooks &= m; p.queens &= m; p.kings &= m;
  p.occ_w &= m; p.occ_b &= m;
}
M0_HD void set_piece(Position& p, int sq, int pt, int color) {
  remove_piece(p, sq);
  u64 m = sq_bb(sq);
  switch (pt) {
    case PT_PAWN: p.pawns |= m; break;
    case PT_KNIGHT: p.knights |= m; break;
    case PT_BISHOP: p.bishops |= m; break;
    case PT_ROOK: p.rooks |= m; break;
    case PT_QUEEN: p.queens |= m; break;
    case PT_KING: p.kings |= m; break;
    default: return;
  }
  if (color) p.occ_w |= m; else p.occ_b |= m;
}

// chess.Board.is_zeroing on the standard (e1g1-style) move
M0_HD bool is_zeroing(const Position& p, Move mv) {
  u64 touched = sq_bb(move_from(mv)) ^ sq_bb(move_to(mv));
  return (touched & p.pawns) || (touched & pos_them(p));
}

// flags describing what the pushed move did (used for repetition bookkeeping)
struct PushInfo {
  bool zeroing;          // pawn move or capture
  bool reduced_castling; // castling rights changed
};

M0_HD PushInfo push_move(Position& p, Move mv) {
  int us = pos_turn(p);
  int from = move_from(mv), to = move_to(mv), promo = move_promo(mv);
  int castling = pos_castling(p);
  int ep_prev = pos_ep(p);
  int halfmove = pos_halfmove(p) + 1;
  int fullmove = pos_fullmove(p) + (us ? 0 : 1);
  PushInfo info;
  // _to_chess960: king e1g1/e1c1 with no rook on the target becomes king-takes-rook
  int to960 = to;
  if (from == SQ_E1 && (p.kings & sq_bb(SQ_E1))) {
    if (to == SQ_G1 && !(p.rooks & sq_bb(SQ_G1))) to960 = SQ_H1;
    else if (to == SQ_C1 && !(p.rooks & sq_bb(SQ_C1))) to960 = SQ_A1;
  } else if (from == SQ_E8 && (p.kings & sq_bb(SQ_E8))) {
    if (to == SQ_G8 && !(p.rooks & sq_bb(SQ_G8))) to960 = SQ_H8;
    else if (to == SQ_C8 && !(p.rooks & sq_bb(SQ_C8))) to960 = SQ_A8;
  }
  to = to960;
  {
    u64 touched = sq_bb(from) ^ sq_bb(to);
    info.zeroing = (touched & p.pawns) || (touched & pos_them(p));
  }
  if (info.zeroing) halfmove = 0;
  u64 from_bb = sq_bb(from), to_bb = sq_bb(to);
  int pt = piece_type_at(p, from);
  remove_piece(p, from);
  int captured = piece_type_at(p, to);
  // castling rights (rook-square mask semantics)
  u64 rights = castling_rook_mask(castling);
  u64 rights0 = rights;
  rights &= ~to_bb & ~from_bb;
  if (pt == PT_KING) {
    rights &= us ? ~RANK_1 : ~RANK_8;
  } else if (captured == PT_KING) {
    if (us && (to >> 3) == 7) rights &= ~RANK_8;
    else if (!us && (to >> 3) == 0) rights &= ~RANK_1;
  }
  info.reduced_castling = rights != rights0;
  int new_ep = EP_NONE;
  if (pt == PT_PAWN) {
    int diff = to - from;
    if (diff == 16 && (from >> 3) == 1) new_ep = from + 8;
    else if (diff == -16 && (from >> 3) == 6) new_ep = from - 8;
    else if (to == ep_prev && ep_prev != EP_NONE && (diff == 7 || diff == 9 || diff == -7 || diff == -9) && !captured) {
      remove_piece(p, ep_prev + (us ? -8 : 8));
    }
  }
  if (promo) pt = promo;
  bool castle = pt == PT_KING && ((us ? p.occ_w : p.occ_b) & to_bb);
  if (castle) {
    bool a_side = (to & 7) < (from & 7);
    remove_piece(p, from);
    remove_piece(p, to);
    if (a_side) {
      set_piece(p, us ? SQ_C1 : SQ_C8, PT_KING, us);
      set_piece(p, us ? SQ_D1 : SQ_D8, PT_ROOK, us);
    } else {
      set_piece(p, us ? SQ_G1 : SQ_G8, PT_KING, us);
      set_piece(p, us ? SQ_F1 : SQ_F8, PT_ROOK, us);
    }
  } else {
    set_piece(p, to, pt, us);
  }
  int bits = 0;
  if (rights & sq_bb(SQ_H1)) bits |= CR_WK;
  if (rights & sq_bb(SQ_A1)) bits |= CR_WQ;
  if (rights & sq_bb(SQ_H8)) bits |= CR_BK;
  if (rights & sq_bb(SQ_A8)) bits |= CR_BQ;
  p.state = pack_state(!us, bits, new_ep, halfmove, fullmove);
  return info;
}

// ---- policy index: azchess/encoding.py:113-150 ---------------------------------------------------
// Pure function of (from, to, promotion, side to move).  Returns -1 when no slot exists.
M0_HD int policy_index(Move mv, int white_to_move) {
  int from = move_from(mv), to = move_to(mv), promo = move_promo(mv);
  int dr = (to >> 3) - (from >> 3), df = (to & 7) - (from & 7);
  // knight deltas, order of encoding.py:70-72
  int adr = dr < 0 ? -dr : dr, adf = df < 0 ? -df : df;
  if ((adr == 2 && adf == 1) || (adr == 1 && adf == 2)) {
    int k;
    if (dr == -2) k = df == -1 ? 0 : 1;
    else if (dr == -1) k = df == -2 ? 2 : 3;
    else if (dr == 1) k = df == -2 ? 4 : 5;
    else k = df == -1 ? 6 : 7;
    return from * 73 + 56 + k;
  }
  if (promo == PT_KNIGHT || promo == PT_BISHOP || promo == PT_ROOK) {
    int dir = -1;
    if (white_to_move) {           // dirs ((1,0),(1,-1),(1,1)), encoding.py:102
      if (dr == 1) dir = df == 0 ? 0 : (df == -1 ? 1 : (df == 1 ? 2 : -1));
    } else {                       // dirs ((-1,0),(-1,1),(-1,-1)), encoding.py:104
      if (dr == -1) dir = df == 0 ? 0 : (df == 1 ? 1 : (df == -1 ? 2 : -1));
    }
    if (dir >= 0) return from * 73 + 64 + (promo - PT_KNIGHT) * 3 + dir;
  }
  if (dr == 0 || df == 0 || adr == adf) {
    int step = adr > adf ? adr : adf;
    int sdr = dr == 0 ? 0 : (dr > 0 ? 1 : -1), sdf = df == 0 ? 0 : (df > 0 ? 1 : -1);
    // RAY_DIRS order N,S,E,W,NE,NW,SE,SW as (d_rank,d_file), encoding.py:60-69
    int d;
    if (sdf == 0) d = sdr > 0 ? 0 : 1;
    else if (sdr == 0) d = sdf > 0 ? 2 : 3;
    else if (sdr > 0) d = sdf > 0 ? 4 : 5;
    else d = sdf > 0 ? 6 : 7;
    if (step >= 1 && step <= 7) return from * 73 + d * 7 + (step - 1);
  }
  return -1;
}

// ---- transposition key: chess.Board._transposition_key -----------------------------------------
struct Key128 { u64 lo, hi; };
M0_HD bool key_eq(const Key128& a, const Key128& b) { return a.lo == b.lo && a.hi == b.hi; }
M0_HD u64 mix64(u64 x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
// 128-bit hash of (8 bitboards, turn, clean castling rights, ep square iff a legal ep capture exists).
// Clocks are NOT part of the key (python-chess), so positions that differ only in clocks merge.
M0_HD Key128 position_key(const Position& p, bool* ep_legal_out = nullptr) {
  bool epl = has_legal_ep(p);
  if (ep_legal_out) *ep_legal_out = epl;
  int ep = epl ? pos_ep(p) : EP_NONE;
  u64 w[9] = {p.pawns, p.knights, p.bishops, p.rooks, p.queens, p.kings, p.occ_w, p.occ_b,
              (u64)pos_turn(p) | ((u64)pos_castling(p) << 1) | ((u64)ep << 5)};
  u64 a = 0x9E3779B97F4A7C15ull, b = 0xC2B2AE3D27D4EB4Full;
  for (int i = 0; i < 9; ++i) {
    a = mix64(a ^ w[i]) + 0x9E3779B97F4A7C15ull * (u64)(i + 1);
    b = mix64((b + w[i]) * 0xD6E8FEB86659FD93ull) ^ (a >> 17);
  }
  Key128 k;
  k.lo = mix64(a ^ (b << 1));
  k.hi = mix64(b ^ (a >> 3) ^ 0xA0761D6478BD642Full);
  if (k.lo == 0) k.lo = 1;  // (0, 0) is the empty-slot marker of the TT and its low word the claim flag of concurrent inserts
  return k;
}

// ---- game-end predicates --------------------------------------------------------------------------
M0_HD bool has_insufficient_material(const Position& p, int color) {
  u64 own = color ? p.occ_w : p.occ_b, other = color ? p.occ_b : p.occ_w;
  if (own & (p.pawns | p.rooks | p.queens)) return false;
  if (own & p.knights) return popcnt(own) <= 2 && !(other & ~p.kings & ~p.queens);
  if (own & p.bishops) {
    bool same_color = !(p.bishops & DARK_SQUARES) || !(p.bishops & LIGHT_SQUARES);
    return same_color && !p.pawns && !p.knights;
  }
  return true;
}
M0_HD bool is_insufficient_material(const Position& p) {
  return has_insufficient_material(p, 1) && has_insufficient_material(p, 0);
}

// ---- synthetic positions: seeded random playouts -----------------------------------------------------
// Position i of a batch = plies_i = hash(seed, i) % (max_plies + 1) uniformly random legal moves from the standard start
// position (stops early when no legal move exists or material is insufficient): azchess/utils/board.py:7-38 (random_board) as a
// pure function of (seed, i), so that the device kernel and the host check (tests/hostcheck) produce the same positions.
M0_HD Position start_position_std() {
  Position p;
  p.pawns = 0x00FF00000000FF00ull; p.knights = 0x4200000000000042ull; p.bishops = 0x2400000000000024ull;
  p.rooks = 0x8100000000000081ull; p.queens = 0x0800000000000008ull; p.kings = 0x1000000000000010ull;
  p.occ_w = 0x000000000000FFFFull; p.occ_b = 0xFFFF000000000000ull;
  p.state = pack_state(1, CR_WK | CR_WQ | CR_BK | CR_BQ, EP_NONE, 0, 1);
  return p;
}
M0_HD Position random_playout_position(u64 seed, int i, int max_plies) {
  Position p = start_position_std();
  u64 rng = mix64(seed ^ (0x9E3779B97F4A7C15ull * (u64)(i + 1)));
  int plies = (int)(rng % (u64)(max_plies + 1));
  Move mv[MAX_MOVES];
  for (int k = 0; k < plies; ++k) {
    if (is_insufficient_material(p)) break;
    int m = generate_legal_moves(p, mv);
    if (m == 0) break;
    if (m > MAX_MOVES) m = MAX_MOVES;
    rng = mix64(rng + 0x9E3779B97F4A7C15ull);
    push_move(p, mv[(int)(rng % (u64)m)]);
  }
  return p;
}

// ---- board planes: azchess/encoding.py:11-46 -------------------------------------------------------
// plane p (0..11) bitboard in python-chess orientation; value of the 7 constant planes 12..18
M0_HD u64 piece_plane_bb(const Position& p, int plane) {
  u64 side = plane < 6 ? p.occ_w : p.occ_b;
  int t = plane % 6;
  u64 bb = t == 0 ? p.pawns : t == 1 ? p.knights : t == 2 ? p.bishops : t == 3 ? p.rooks : t == 4 ? p.queens : p.kings;
  return bb & side;
}
M0_HD float const_plane_value(const Position& p, int plane) {
  int cr = pos_castling(p);
  switch (plane) {
    case 12: return pos_turn(p) ? 1.0f : 0.0f;
    case 13: return (cr & CR_WK) ? 1.0f : 0.0f;
    case 14: return (cr & CR_WQ) ? 1.0f : 0.0f;
    case 15: return (cr & CR_BK) ? 1.0f : 0.0f;
    case 16: return (cr & CR_BQ) ? 1.0f : 0.0f;
    case 17: { int h = pos_halfmove(p); if (h > 99) h = 99; return (float)((double)h / 99.0); }
    default: { int f = pos_fullmove(p); if (f > 199) f = 199; return (float)((double)f / 199.0); }
  }
}
// value of planes[plane][row][col] with row = 7 - rank, col = file (encoding.py:43-45)
M0_HD float plane_value(const Position& p, int plane, int row, int col) {
  if (plane < 12) return ((piece_plane_bb(p, plane) >> ((7 - row) * 8 + col)) & 1) ? 1.0f : 0.0f;
  return const_plane_value(p, plane);
}

}  // namespace m0
