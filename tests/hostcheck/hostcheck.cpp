// TEST HARNESS ONLY -- compiles matrix0_b200/csrc/chess_core.cuh with g++ so that the CPU-only
// build container can compare the product's chess logic with the oracle before any GPU time is
// spent.  Built by tests/test_hostcheck.py into tests/hostcheck/_hostcheck.so; never imported by
// matrix0_b200 (the product has no CPU path and raises when the CUDA library is missing).
#include "../../matrix0_b200/csrc/chess_core.cuh"
#include "../../matrix0_b200/csrc/ssl_core.cuh"
#include "../../matrix0_b200/csrc/movegen_warp.cuh"
#include "../../matrix0_b200/csrc/search_math.cuh"
#include <string.h>
using namespace m0;

static Position load(const uint64_t* w) {
  Position p;
  p.pawns = w[0]; p.knights = w[1]; p.bishops = w[2]; p.rooks = w[3]; p.queens = w[4]; p.kings = w[5];
  p.occ_w = w[6]; p.occ_b = w[7]; p.state = w[8];
  return p;
}
static void store(const Position& p, uint64_t* w) {
  w[0] = p.pawns; w[1] = p.knights; w[2] = p.bishops; w[3] = p.rooks; w[4] = p.queens; w[5] = p.kings;
  w[6] = p.occ_w; w[7] = p.occ_b; w[8] = p.state;
}

extern "C" {
// raw python-chess fields -> packed 9-word position (cleans the castling rights)
void hc_pack(const uint64_t* bbs8, uint64_t raw_castling, int turn, int ep, int halfmove, int fullmove, uint64_t* out9) {
  Position p;
  p.pawns = bbs8[0]; p.knights = bbs8[1]; p.bishops = bbs8[2]; p.rooks = bbs8[3]; p.queens = bbs8[4]; p.kings = bbs8[5];
  p.occ_w = bbs8[6]; p.occ_b = bbs8[7];
  p.state = 0;
  int bits = clean_castling_bits(p, raw_castling);
  p.state = pack_state(turn, bits, ep < 0 ? EP_NONE : ep, halfmove, fullmove);
  store(p, out9);
}
int hc_legal_moves(const uint64_t* pos9, uint16_t* moves, int16_t* idx) {
  Position p = load(pos9);
  Move mv[MAX_MOVES];
  int n = generate_legal_moves(p, mv);
  int m = n < MAX_MOVES ? n : MAX_MOVES;
  for (int i = 0; i < m; ++i) { moves[i] = mv[i]; idx[i] = (int16_t)policy_index(mv[i], pos_turn(p)); }
  return n;
}
void hc_push(const uint64_t* pos9, uint16_t mv, uint64_t* out9, int* flags) {
  Position p = load(pos9);
  PushInfo pi = push_move(p, mv);
  store(p, out9);
  flags[0] = pi.zeroing; flags[1] = pi.reduced_castling;
}
void hc_key(const uint64_t* pos9, uint64_t* key2) {
  Key128 k = position_key(load(pos9));
  key2[0] = k.lo; key2[1] = k.hi;
}
// the positions m0_random_playouts(seed, max_plies) writes, computed on the host (same function as the device kernel)
void hc_random_playouts(uint64_t* out, int first, int n, uint64_t seed, int max_plies) {
  for (int i = 0; i < n; ++i) store(random_playout_position(seed, first + i, max_plies), out + (size_t)i * 9);
}
// The warp-cooperative ordered generator (movegen_warp.cuh) with its 32 lanes simulated one after the other: the same per-lane
// functions (lane_generate / lane_counts / lane_emit) and the same suffix-sum placement as the device wrapper.
// Returns -1 when the board takes the single-lane fallback on the device.
int hc_legal_moves_warp(const uint64_t* pos9, uint16_t* moves, int* in_check_out) {
  Position p = load(pos9);
  if (!warp_movegen_supported(p)) return -1;
  const u64 ours = pos_us(p), theirs = pos_them(p), king_bb = p.kings & ours;
  u64 danger = 0, checkers = 0;
  for (u64 t = theirs; t; t &= t - 1) {
    const int e = lsb(t);
    const u64 a = attacks_from(p, e);
    danger |= a;
    if (a & king_bb) checkers |= sq_bb(e);
  }
  const LegalCtx c = make_legal_ctx(p, &checkers);
  const bool in_check = checkers != 0;
  LaneGen g[32];
  LaneCounts n[32];
  u64 rest = ours;
  for (int lane = 0; lane < 32; ++lane) {
    int from = -1;
    if (rest) { from = lsb(rest); rest &= rest - 1; }
    g[lane] = lane_generate(p, c, danger, from);
    n[lane] = lane_counts(g[lane], in_check);
  }
  u32 s0[33], s1[33];
  s0[32] = s1[32] = 0;
  for (int lane = 31; lane >= 0; --lane) { s0[lane] = s0[lane + 1] + n[lane].w0; s1[lane] = s1[lane + 1] + n[lane].w1; }
  int n_castle = 0, ksq = 0, cto[2] = {0, 0};
  if (!in_check) n_castle = legal_castling(p, c, &ksq, cto, &danger);
  const u32 tot0 = s0[0], tot1 = s1[0];
  const int nK = (int)(tot1 >> 24), nA = (int)(tot0 & 0xFFFFu), nC = (int)(tot0 >> 16), nS = (int)(tot1 & 0xFFu), nD = (int)((tot1 >> 8) & 0xFFu),
            nE = (int)((tot1 >> 16) & 0xFFu);
  const int base_a = nK, base_z = base_a + nA, base_c = base_z + n_castle, base_s = base_c + nC, base_d = base_s + nS, base_e = base_d + nD;
  for (int lane = 0; lane < 32; ++lane) lane_emit(g[lane], in_check, c.ep, moves, base_a, base_c, base_s, base_d, base_e, s0[lane + 1], s1[lane + 1]);
  for (int i = 0; i < n_castle; ++i) { int at = base_z + i; put_move(moves, at, ksq, cto[i], 0); }
  *in_check_out = in_check ? 1 : 0;
  return base_e + nE;
}
int hc_has_legal_ep(const uint64_t* pos9) { return has_legal_ep(load(pos9)); }
int hc_insufficient(const uint64_t* pos9) { return is_insufficient_material(load(pos9)); }
void hc_planes(const uint64_t* pos9, float* out) {
  Position p = load(pos9);
  for (int pl = 0; pl < 19; ++pl)
    for (int r = 0; r < 8; ++r)
      for (int c = 0; c < 8; ++c) out[(pl * 8 + r) * 8 + c] = plane_value(p, pl, r, c);
}
// 17 maps x 64 floats in plane coordinates: piece[13], threat, pin, fork, control (ssl_core.cuh)
void hc_ssl(const uint64_t* pos9, float* out) {
  SslMasks m;
  ssl_masks(load(pos9), m);
  for (int i = 0; i < 64; ++i) {
    for (int k = 0; k < 13; ++k) out[k * 64 + i] = (m.piece[k] >> i) & 1 ? 1.0f : 0.0f;
    out[13 * 64 + i] = (m.threat >> i) & 1 ? 1.0f : 0.0f;
    out[14 * 64 + i] = 0.0f;
    out[15 * 64 + i] = (m.fork >> i) & 1 ? 1.0f : 0.0f;
    out[16 * 64 + i] = (m.ctrl_pos >> i) & 1 ? 1.0f : ((m.ctrl_neg >> i) & 1 ? -1.0f : 0.0f);
  }
}
}

// legal mask through the per-piece SET interface (what the warp-cooperative encode kernel evaluates, one piece per lane)
static int legal_set_mask(const Position& p, uint8_t* mask) {
  memset(mask, 0, POLICY_SIZE);
  u64 danger = 0, checkers = 0, theirs = pos_them(p), kings = p.kings & pos_us(p);
  u64 king_bb = kings ? sq_bb(msb(kings)) : 0;
  while (theirs) {
    int e = lsb(theirs);
    theirs &= theirs - 1;
    danger |= enemy_attacks(p, e, king_bb, &checkers);
  }
  LegalCtx c = make_legal_ctx(p, &checkers);
  int wtm = pos_turn(p), n = 0;
  u64 own = c.ours;
  while (own) {
    int from = lsb(own);
    own &= own - 1;
    if (from == c.king) {
      u64 t = c.king_cand & ~danger;
      while (t) {
        int to = lsb(t);
        t &= t - 1;
        mask[policy_index(make_move(from, to, 0), wtm)] = 1; n++;
      }
      continue;
    }
    u64 t = piece_targets(p, c, from);
    bool pawn = (p.pawns >> from) & 1;
    while (t) {
      int to = lsb(t);
      t &= t - 1;
      if (pawn && ((to >> 3) == 0 || (to >> 3) == 7)) {
        for (int pr = PT_KNIGHT; pr <= PT_QUEEN; ++pr) { mask[policy_index(make_move(from, to, pr), wtm)] = 1; n++; }
      } else {
        mask[policy_index(make_move(from, to, 0), wtm)] = 1; n++;
      }
    }
    if (pawn && ep_capture_legal(p, c, from)) { mask[policy_index(make_move(from, c.ep, 0), wtm)] = 1; n++; }
  }
  int ksq = 0, to[2];
  int nc = legal_castling(p, c, &ksq, to, &danger);
  for (int i = 0; i < nc; ++i) { mask[policy_index(make_move(ksq, to[i], 0), wtm)] = 1; n++; }
  return n;
}
static int ordered_mask(const Position& p, uint8_t* mask) {
  memset(mask, 0, POLICY_SIZE);
  Move mv[MAX_MOVES];
  int n = generate_legal_moves(p, mv);
  int m = n < MAX_MOVES ? n : MAX_MOVES;
  for (int i = 0; i < m; ++i) mask[policy_index(mv[i], pos_turn(p))] = 1;
  return n;
}
extern "C" {
// 0 when the SET interface gives the same move count and mask as the ordered generator
int hc_legal_set_differs(const uint64_t* pos9) {
  static uint8_t a[POLICY_SIZE], b[POLICY_SIZE];
  Position p = load(pos9);
  int na = legal_set_mask(p, a), nb = ordered_mask(p, b);
  return (na != nb) || memcmp(a, b, POLICY_SIZE) != 0;
}
// random playouts from `start9`: returns the number of positions visited, *bad = positions where the two disagree
long hc_legal_set_sweep(const uint64_t* start9, uint64_t seed, int games, int max_plies, long* bad, uint64_t* first_bad9) {
  static uint8_t a[POLICY_SIZE], b[POLICY_SIZE];
  long seen = 0;
  *bad = 0;
  u64 rng = seed;
  for (int g = 0; g < games; ++g) {
    Position p = load(start9);
    for (int ply = 0; ply < max_plies; ++ply) {
      int na = legal_set_mask(p, a), nb = ordered_mask(p, b);
      seen++;
      if (na != nb || memcmp(a, b, POLICY_SIZE) != 0) { if (!*bad) store(p, first_bad9); (*bad)++; }
      Move mv[MAX_MOVES];
      int n = generate_legal_moves(p, mv);
      if (n == 0 || is_insufficient_material(p) || pos_halfmove(p) >= 150) break;
      rng = mix64(rng + 0x9E3779B97F4A7C15ull);
      push_move(p, mv[(int)(rng % (u64)n)]);
    }
  }
  return seen;
}
}

// csrc/search_math.cuh: the single-rounding arithmetic of the search (numpy's pairwise sums, the PUCT score, repeated backups)
extern "C" {
float hc_pairwise_f32(const float* a, int n, int big) { return big ? np_pairwise_sum_f32_big(a, n) : np_pairwise_sum_f32(a, n); }
double hc_pairwise_f64(const double* a, int n, int big) { return big ? np_pairwise_sum_f64_big(a, n) : np_pairwise_sum_f64(a, n); }
double hc_puct_score(double q, double cpuct, double prior, double sqrt_parent_visits, int child_n) {
  return puct_score(q, cpuct, prior, sqrt_parent_visits, child_n);
}
void hc_backup_repeated(int* n, double* w, double* q, double v, int times) { backup_repeated(*n, *w, *q, v, times); }
double hc_clip_unit(double x) { return py_clip_unit(x); }
}
