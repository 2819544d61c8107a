// Warp-cooperative legal-move generation in python-chess ORDER (chess.Board.generate_legal_moves):
// one own piece per lane, the ordered list assembled by prefix sums over the lanes.
//
// The tree kernels need the ordered list of every leaf (Node.children order = legal-move order decides PUCT tie-breaks,
// azchess/mcts.py:140, :901).  generate_legal_moves() of chess_core.cuh produces it on ONE lane while 31 lanes idle; here
//   1. every lane takes one opposing piece: OR-reduction of the attack sets = the danger map (squares the king may not step on)
//      and the checkers;
//   2. every lane takes one own piece (lane j = j-th lowest square of the side to move) and computes its legal destinations as
//      bitboards per category -- the legal moves as a SET (LegalCtx / piece_targets_given / ep_capture_legal, the functions the
//      half-warp encode kernel uses; pins and evasion masks instead of a per-move safety test);
//   3. python-chess emits [king evasions when in check] [non-pawn pieces, from-squares high to low, targets high to low] [castling,
//      h-side first] [pawn captures, from high to low, promotions Q R B N] [single pushes by target high to low] [double pushes]
//      [en passant, capturers high to low]: within a category the order is descending lane order, so a lane's write offset is the
//      category base plus the SUFFIX sum of the counts of the lanes above it (five shuffle steps on packed counters);
//   4. every lane writes its own moves.
// The per-lane pieces are plain M0_HD functions, so tests/hostcheck simulates the 32 lanes on the CPU and compares the list with
// the ordered single-thread generator on every test position.  Boards outside standard chess (no king / several own kings / more
// than 32 pieces of a colour) take the single-lane generator.
#pragma once
#include "chess_core.cuh"

namespace m0 {

struct LaneGen {
  int from;        // own piece of this lane, -1: none
  u64 kng;         // steps of THE king (all of them legal); evasions when in check
  u64 np;          // destinations of a non-pawn, non-king piece
  u64 cap, sgl, dbl;   // pawn captures, single and double pushes
  int ep;          // 1: this pawn may capture en passant
};
struct LaneCounts { u32 w0, w1; };   // w0 = A | C << 16 ; w1 = S | D << 8 | E << 16 | K << 24

M0_HD bool warp_movegen_supported(const Position& p) {
  const u64 ours = pos_us(p), kings = p.kings & ours;
  return kings && !(kings & (kings - 1)) && popcnt(ours) <= 32 && popcnt(pos_them(p)) <= 32;
}

// destinations of the own piece standing on `from` (c and danger are the same for all lanes)
M0_HD LaneGen lane_generate(const Position& p, const LegalCtx& c, u64 danger, int from) {
  LaneGen g;
  g.from = from;
  g.kng = g.np = g.cap = g.sgl = g.dbl = 0;
  g.ep = 0;
  if (from < 0) return g;
  const u64 fb = sq_bb(from);
  if (from == c.king) {
    g.kng = c.king_cand & ~danger;
  } else if (p.pawns & fb) {
    const u64 t = piece_targets_given(p, c, from, 0);
    g.cap = t & c.theirs;
    const u64 push = t & ~c.theirs;
    g.sgl = push & (c.us ? (fb << 8) : (fb >> 8));
    g.dbl = push & ~g.sgl;
    g.ep = ep_capture_legal(p, c, from) ? 1 : 0;
  } else {
    g.np = piece_targets_given(p, c, from, attacks_from(p, from));
  }
  return g;
}

M0_HD int promo_weighted(u64 targets) {   // a destination on the first / last rank stands for four promotions
  const u64 back = RANK_1 | RANK_8;
  return popcnt(targets & ~back) + 4 * popcnt(targets & back);
}
// category counts of a lane.  In check the king's steps form their own leading category, otherwise the king is one of the
// non-pawn pieces (scan_reversed(our & ~pawns) reaches it at its square)
M0_HD LaneCounts lane_counts(const LaneGen& g, bool in_check) {
  LaneCounts n;
  const int k = popcnt(g.kng);
  const int a = popcnt(g.np) + (in_check ? 0 : k);
  n.w0 = (u32)a | ((u32)promo_weighted(g.cap) << 16);
  n.w1 = (u32)promo_weighted(g.sgl) | ((u32)popcnt(g.dbl) << 8) | ((u32)g.ep << 16) | ((u32)(in_check ? k : 0) << 24);
  return n;
}

M0_HD void put_move(Move* out, int& at, int from, int to, int promo) {
  if (at < MAX_MOVES) out[at] = make_move(from, to, promo);
  ++at;
}
M0_HD void put_targets(Move* out, int& at, int from, u64 t, bool pawn) {
  while (t) {
    const int to = msb(t);
    t ^= sq_bb(to);
    if (pawn && ((to >> 3) == 0 || (to >> 3) == 7)) {
      put_move(out, at, from, to, PT_QUEEN);
      put_move(out, at, from, to, PT_ROOK);
      put_move(out, at, from, to, PT_BISHOP);
      put_move(out, at, from, to, PT_KNIGHT);
    } else {
      put_move(out, at, from, to, 0);
    }
  }
}
// write the lane's moves; base_* = first index of each category, above_* = moves of the lanes ABOVE this one in that category
M0_HD void lane_emit(const LaneGen& g, bool in_check, int ep_sq, Move* out, int base_a, int base_c, int base_s, int base_d, int base_e,
                     u32 above0, u32 above1) {
  if (g.from < 0) return;
  if (g.kng) {
    int at = in_check ? 0 : base_a + (int)(above0 & 0xFFFFu);
    put_targets(out, at, g.from, g.kng, false);
  }
  if (g.np) {
    int at = base_a + (int)(above0 & 0xFFFFu);
    put_targets(out, at, g.from, g.np, false);
  }
  if (g.cap) {
    int at = base_c + (int)(above0 >> 16);
    put_targets(out, at, g.from, g.cap, true);
  }
  if (g.sgl) {
    int at = base_s + (int)(above1 & 0xFFu);
    put_targets(out, at, g.from, g.sgl, true);
  }
  if (g.dbl) {
    int at = base_d + (int)((above1 >> 8) & 0xFFu);
    put_targets(out, at, g.from, g.dbl, false);
  }
  if (g.ep) {
    int at = base_e + (int)((above1 >> 16) & 0xFFu);
    put_move(out, at, g.from, ep_sq, 0);
  }
}

#if defined(__CUDACC__)
// j-th set bit of b (j < popcount)
__device__ __forceinline__ int nth_set_bit64(u64 b, int j) {
  u32 w = (u32)b;
  int pos = 0, c = __popc(w);
  if (j >= c) { j -= c; w = (u32)(b >> 32); pos = 32; }
  c = __popc(w & 0xFFFFu); if (j >= c) { j -= c; w >>= 16; pos += 16; }
  c = __popc(w & 0xFFu);   if (j >= c) { j -= c; w >>= 8;  pos += 8; }
  c = __popc(w & 0xFu);    if (j >= c) { j -= c; w >>= 4;  pos += 4; }
  c = __popc(w & 0x3u);    if (j >= c) { j -= c; w >>= 2;  pos += 2; }
  return pos + (j >= (int)(w & 1u) ? 1 : 0);
}

// All 32 lanes call this with the same position; `out` is this warp's list (shared memory).  Returns the number of legal moves
// (entries past MAX_MOVES are dropped) and, through *in_check_out, whether the side to move is in check.  Ends with __syncwarp().
__device__ __forceinline__ int warp_generate_legal_moves(const Position& p, Move* out, int* in_check_out, int lane) {
  const unsigned FULLW = 0xFFFFFFFFu;
  if (!warp_movegen_supported(p)) {
    int n = 0, chk = 0;
    if (lane == 0) {
      u64 checkers;
      n = generate_legal_moves(p, out, &checkers);
      chk = checkers != 0;
    }
    n = __shfl_sync(FULLW, n, 0);
    *in_check_out = __shfl_sync(FULLW, chk, 0);
    __syncwarp();
    return n;
  }
  const u64 ours = pos_us(p), theirs = pos_them(p);
  const u64 king_bb = p.kings & ours;
  u64 danger = 0, checkers = 0;
  if (lane < popcnt(theirs)) {
    const int e = nth_set_bit64(theirs, lane);
    danger = attacks_from(p, e);
    if (danger & king_bb) checkers = sq_bb(e);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    danger |= __shfl_xor_sync(FULLW, danger, o);
    checkers |= __shfl_xor_sync(FULLW, checkers, o);
  }
  const LegalCtx c = make_legal_ctx(p, &checkers);
  const bool in_check = checkers != 0;
  const LaneGen g = lane_generate(p, c, danger, lane < popcnt(ours) ? nth_set_bit64(ours, lane) : -1);
  const LaneCounts mine = lane_counts(g, in_check);
  // inclusive suffix sums over the lanes (lane 31 = highest square comes first in every category)
  u32 s0 = mine.w0, s1 = mine.w1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t0 = __shfl_down_sync(FULLW, s0, o), t1 = __shfl_down_sync(FULLW, s1, o);
    if (lane + o < 32) { s0 += t0; s1 += t1; }
  }
  const u32 tot0 = __shfl_sync(FULLW, s0, 0), tot1 = __shfl_sync(FULLW, s1, 0);
  int n_castle = 0, ksq = 0, cto[2] = {0, 0};
  if (!in_check) {   // the same for all lanes (cheap: constant masks against the danger map)
    n_castle = legal_castling(p, c, &ksq, cto, &danger);
  }
  const int nK = (int)(tot1 >> 24), nA = (int)(tot0 & 0xFFFFu), nC = (int)(tot0 >> 16), nS = (int)(tot1 & 0xFFu), nD = (int)((tot1 >> 8) & 0xFFu),
            nE = (int)((tot1 >> 16) & 0xFFu);
  const int base_a = nK, base_z = base_a + nA, base_c = base_z + n_castle, base_s = base_c + nC, base_d = base_s + nS, base_e = base_d + nD;
  lane_emit(g, in_check, c.ep, out, base_a, base_c, base_s, base_d, base_e, s0 - mine.w0, s1 - mine.w1);
  if (lane == 0)
    for (int i = 0; i < n_castle; ++i) {
      int at = base_z + i;
      put_move(out, at, ksq, cto[i], 0);
    }
  *in_check_out = in_check ? 1 : 0;
  __syncwarp();
  return base_e + nE;
}
#endif

}  // namespace m0
