// Position -> (planes, legal moves, policy indices, legal mask) kernels.
//
// Replaces azchess/encoding.py:11-46 (encode_board), :113-150 (move_to_index) and :243-253
// (MoveEncoder.get_legal_actions) for whole batches of positions resident in HBM.
//
// HBM-bound byte work: per position 72 B are read and 4,864 B of float32
// planes + 4,672 B of uint8 mask are written.  Layout: positions are packed [n][9] u64; a block of
// 128 threads stages its 128 positions (9,216 contiguous bytes) through shared memory with
// coalesced 8-byte loads, every thread then runs the ordered legal-move generator on one position
// (integer ALU work, hidden under the stores of the other resident blocks), and the warps write
// the outputs cooperatively with 16-byte vector stores so that every store instruction covers
// 512 contiguous bytes.
#include "chess_core.cuh"
#include "m0_common.cuh"

namespace m0 {

static constexpr int ENC_THREADS = 128;

// ---- raw python-chess fields -> packed position ------------------------------------------------
// raw[i] = {pawns, knights, bishops, rooks, queens, kings, occ_white, occ_black, castling_rights
//           (rook-square mask), turn | ep<<8 (255 = none) | halfmove<<16 | fullmove<<32}
__global__ void pack_positions_kernel(const u64* __restrict__ raw, int n, u64* __restrict__ pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64* r = raw + (size_t)i * 10;
  Position p;
  p.pawns = r[0]; p.knights = r[1]; p.bishops = r[2]; p.rooks = r[3]; p.queens = r[4]; p.kings = r[5];
  p.occ_w = r[6]; p.occ_b = r[7];
  u64 misc = r[9];
  int turn = (int)(misc & 1);
  int ep = (int)((misc >> 8) & 255);
  if (ep > 63) ep = EP_NONE;
  int half = (int)((misc >> 16) & 0xFFFF), full = (int)((misc >> 32) & 0xFFFF);
  p.state = pack_state(turn, clean_castling_bits(p, r[8]), ep, half, full);
  store_position(pos + (size_t)i * POSITION_WORDS, p);
}


// ---- synthetic positions: seeded random playouts from the start position ---------------------------
// Mirrors azchess/utils/board.py:7-38 (random_board) for whole batches: position i is reached by
// playing plies_i = hash(seed, i) % (max_plies + 1) uniformly random legal moves from the standard
// start position (stops early when no legal move exists or material is insufficient).
__global__ void random_playouts_kernel(u64* __restrict__ pos, int n, u64 seed, int max_plies) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Position p;
  p.pawns = 0x00FF00000000FF00ull; p.knights = 0x4200000000000042ull; p.bishops = 0x2400000000000024ull;
  p.rooks = 0x8100000000000081ull; p.queens = 0x0800000000000008ull; p.kings = 0x1000000000000010ull;
  p.occ_w = 0x000000000000FFFFull; p.occ_b = 0xFFFF000000000000ull;
  p.state = pack_state(1, CR_WK | CR_WQ | CR_BK | CR_BQ, EP_NONE, 0, 1);
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
  store_position(pos + (size_t)i * POSITION_WORDS, p);
}

// Fused kernel of the encode + legal-mask microbenchmark.  Any of planes / mask / moves may be null.
__global__ void __launch_bounds__(ENC_THREADS)
encode_positions_kernel(const u64* __restrict__ pos, int n, float* __restrict__ planes, u8* __restrict__ mask,
                        u16* __restrict__ moves, u16* __restrict__ idx, int* __restrict__ counts) {
  __shared__ u64 s_pos[ENC_THREADS * POSITION_WORDS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = blockIdx.x * ENC_THREADS;
  const int nb = min(ENC_THREADS, n - base);
  // stage the block's positions: 9 coalesced 8-byte loads per thread
  for (int w = tid; w < nb * POSITION_WORDS; w += ENC_THREADS) s_pos[w] = ld_global_nc_u64(pos + (size_t)base * POSITION_WORDS + w);
  __syncthreads();

  // per-thread ordered legal-move generation (policy indices kept in thread-local memory)
  u16 my_idx[MAX_MOVES];
  int my_n = 0;
  if (tid < nb) {
    Position p = load_position(s_pos + tid * POSITION_WORDS);
    Move mv[MAX_MOVES];
    my_n = generate_legal_moves(p, mv);
    if (my_n > MAX_MOVES) my_n = MAX_MOVES;
    const int wtm = pos_turn(p);
    for (int k = 0; k < my_n; ++k) my_idx[k] = (u16)policy_index(mv[k], wtm);
    if (counts) counts[base + tid] = my_n;
    if (moves) {
      u16* mrow = moves + (size_t)(base + tid) * MAX_MOVES;
      for (int k = 0; k < my_n; ++k) mrow[k] = mv[k];
    }
    if (idx) {
      u16* irow = idx + (size_t)(base + tid) * MAX_MOVES;
      for (int k = 0; k < my_n; ++k) irow[k] = my_idx[k];
    }
  }

  // planes: each warp walks its 32 positions; all lanes store one float4 per step
  if (planes) {
    for (int j = 0; j < 32; ++j) {
      int t = warp * 32 + j;
      if (t >= nb) break;
      Position p = load_position(s_pos + t * POSITION_WORDS);
      warp_write_planes(p, planes + (size_t)(base + t) * (19 * 64), lane);
    }
  }
  // mask: zero-fill each row with 16-byte stores, then every thread scatters its own legal bytes.
  // Both hit L2 before write-back, so DRAM sees each row once.
  if (mask) {
    for (int j = 0; j < 32; ++j) {
      int t = warp * 32 + j;
      if (t >= nb) break;
      uint4* row = reinterpret_cast<uint4*>(mask + (size_t)(base + t) * POLICY_SIZE);
      for (int c = lane; c < POLICY_SIZE / 16; c += 32) row[c] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    if (tid < nb) {
      u8* row = mask + (size_t)(base + tid) * POLICY_SIZE;
      for (int k = 0; k < my_n; ++k) row[my_idx[k]] = 1;
    }
  }
}

}  // namespace m0

using namespace m0;

extern "C" {

int m0_positions_pack(const uint64_t* d_raw, int n, uint64_t* d_pos, void* stream) {
  if (n <= 0) return 0;
  pack_positions_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_raw, n, d_pos);
  return m0_check_launch("m0_positions_pack");
}

int m0_random_playouts(uint64_t* d_pos, int n, uint64_t seed, int max_plies, void* stream) {
  if (n <= 0) return 0;
  if (max_plies < 0) { m0_set_error("m0_random_playouts: max_plies < 0"); return M0_ERR_ARG; }
  random_playouts_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_pos, n, seed, max_plies);
  return m0_check_launch("m0_random_playouts");
}

int m0_encode_positions(const uint64_t* d_pos, int n, float* d_planes, uint8_t* d_mask, uint16_t* d_moves,
                        uint16_t* d_idx, int32_t* d_counts, void* stream) {
  if (n <= 0) return 0;
  int blocks = (n + ENC_THREADS - 1) / ENC_THREADS;
  encode_positions_kernel<<<blocks, ENC_THREADS, 0, (cudaStream_t)stream>>>(d_pos, n, d_planes, d_mask, d_moves, d_idx, d_counts);
  return m0_check_launch("m0_encode_positions");
}

int m0_encode_planes(const uint64_t* d_pos, int n, float* d_planes, void* stream) {
  return m0_encode_positions(d_pos, n, d_planes, nullptr, nullptr, nullptr, nullptr, stream);
}

int m0_legal_mask(const uint64_t* d_pos, int n, uint8_t* d_mask, void* stream) {
  return m0_encode_positions(d_pos, n, nullptr, d_mask, nullptr, nullptr, nullptr, stream);
}

int m0_legal_moves(const uint64_t* d_pos, int n, uint16_t* d_moves, uint16_t* d_idx, int32_t* d_counts, void* stream) {
  return m0_encode_positions(d_pos, n, nullptr, nullptr, d_moves, d_idx, d_counts, stream);
}

}  // extern "C"
