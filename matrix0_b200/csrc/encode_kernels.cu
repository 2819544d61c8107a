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
// 512 contiguous bytes.  That kernel (encode_positions_kernel) serves every call that wants the ORDERED
// move / index lists.  Planes and mask alone -- the microbenchmark and encode_boards() -- go through
// encode_mask_planes_kernel: a half warp per position, one own piece per lane (the legal moves as a set,
// chess_core.cuh LegalCtx), the mask row assembled as 4,672 bits in shared memory and expanded to bytes on
// the way out, so that the 32 lanes of a warp no longer wait for the slowest of 32 unrelated move lists
// (ncu of the thread-per-position kernel: 1.3 k issued warp instructions per position at 25 % occupancy,
// issue-bound at 58 % of the HBM roofline).
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
  const Position p = random_playout_position(seed, i, max_plies);
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


// ---- planes + mask, half warp per position ----------------------------------------------------------
static constexpr int ENCW_THREADS = 128;                   // 8 half warps (the table fills below assume 128)
static constexpr int ENCW_HALVES = ENCW_THREADS / 16;
static constexpr int ENCW_POS_PER_HALF = 8;
static constexpr int ENCW_POS_PER_BLOCK = ENCW_HALVES * ENCW_POS_PER_HALF;
static constexpr int MASK_BIT_WORDS = 148;                 // 4,672 bits = 146 words, padded to 37 x 16 bytes
static constexpr int MASK_CHUNKS = POLICY_SIZE / 16;       // 292 16-byte chunks per mask row

// policy index of (from, to) without under-promotion: encoding.py:113-150 is a pure function of the two squares then
// (a queen promotion shares the slot of the plain move).  Filled once per device by build_move_index_tab_kernel.
__device__ u16 g_move_index_tab[64 * 64];

__global__ void build_move_index_tab_kernel() {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 64) return;
  int from = i >> 6, to = i & 63;
  int v = from == to ? -1 : policy_index(make_move(from, to, 0), 1);
  g_move_index_tab[i] = v < 0 ? (u16)0xFFFF : (u16)v;
}

__device__ __forceinline__ void st_global_cs_u4(uint4* ptr, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// four mask bits -> four bytes of 0 / 1
__device__ __forceinline__ u32 spread4(u32 nibble) { return (nibble * 0x00204081u) & 0x01010101u; }
// j-th set bit of b (j < popcount): halving steps on population counts, no loop
__device__ __forceinline__ int nth_set_bit(u64 b, int j) {
  u32 w = (u32)b;
  int pos = 0, c = __popc(w);
  if (j >= c) { j -= c; w = (u32)(b >> 32); pos = 32; }
  c = __popc(w & 0xFFFFu); if (j >= c) { j -= c; w >>= 16; pos += 16; }
  c = __popc(w & 0xFFu);   if (j >= c) { j -= c; w >>= 8;  pos += 8; }
  c = __popc(w & 0xFu);    if (j >= c) { j -= c; w >>= 4;  pos += 4; }
  c = __popc(w & 0x3u);    if (j >= c) { j -= c; w >>= 2;  pos += 2; }
  return pos + (j >= (int)(w & 1u) ? 1 : 0);
}
__device__ __forceinline__ void set_mask_bit(u32* bits, int idx) { atomicOr(bits + (idx >> 5), 1u << (idx & 31)); }

// 16 lanes write the 19x8x8 float32 planes of one position: 16 float4 chunks per plane, one per lane
__device__ __forceinline__ void half_write_planes(const Position& p, float* __restrict__ out, int hl) {
  float4* o4 = reinterpret_cast<float4*>(out) + hl;
  // row = 7 - rank (encoding.py:43-45), four files per chunk: the lane's nibble lies in one 32-bit half of a bitboard
  const bool hi = (hl >> 1) < 4;
  const int sh = ((7 - (hl >> 1)) & 3) * 8 + (hl & 1) * 4;
  auto half32 = [&](u64 bb) { return hi ? (u32)(bb >> 32) : (u32)bb; };
  const u32 side[2] = {half32(p.occ_w), half32(p.occ_b)};
  const u32 kind[6] = {half32(p.pawns), half32(p.knights), half32(p.bishops), half32(p.rooks), half32(p.queens), half32(p.kings)};
#pragma unroll
  for (int plane = 0; plane < 12; ++plane) {              // plane order of piece_plane_bb: white P N B R Q K, black P N B R Q K
    u32 bits = ((kind[plane % 6] & side[plane / 6]) >> sh) & 15u;
    st_global_cs_f4(o4 + plane * 16, make_float4((bits & 1) ? 1.0f : 0.0f, (bits & 2) ? 1.0f : 0.0f,
                                                 (bits & 4) ? 1.0f : 0.0f, (bits & 8) ? 1.0f : 0.0f));
  }
#pragma unroll
  for (int plane = 12; plane < 19; ++plane) {
    float f = const_plane_value(p, plane);
    st_global_cs_f4(o4 + plane * 16, make_float4(f, f, f, f));
  }
}

// 8 blocks per SM (64 registers, 32 warps): measured 2.02 ms against 2.27 ms at the natural 92 registers / 5 blocks on the
// same box; 10 and 12 blocks spill and are slower (tools/scratch/enc_variants.py)
#ifndef ENCW_MIN_BLOCKS
#define ENCW_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(ENCW_THREADS, ENCW_MIN_BLOCKS)
encode_mask_planes_kernel(const u64* __restrict__ pos, int n, float* __restrict__ planes, u8* __restrict__ mask) {
  __shared__ __align__(16) u16 s_tab[64 * 64];
  __shared__ __align__(16) u32 s_bits[ENCW_HALVES][MASK_BIT_WORDS];
  __shared__ u64 s_knight[64], s_king[64];
  __shared__ uint2 s_spread[256];                          // eight mask bits -> eight bytes of 0 / 1
  const int tid = threadIdx.x, hl = tid & 15, half = tid >> 4;
  const unsigned hmask = 0xFFFFu << (tid & 16);            // the 16 lanes that share a position
  if (mask) {
    for (int b = tid; b < 256; b += ENCW_THREADS) s_spread[b] = make_uint2(spread4(b & 15), spread4(b >> 4));
    if (tid < 64) s_knight[tid] = knight_attacks_bb(sq_bb(tid));
    else s_king[tid - 64] = king_attacks_bb(sq_bb(tid - 64));
    const uint4* src = reinterpret_cast<const uint4*>(g_move_index_tab);
    uint4* dst = reinterpret_cast<uint4*>(s_tab);
    for (int i = tid; i < 64 * 64 * 2 / 16; i += ENCW_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  u32* bits = s_bits[half];
  for (int k = 0; k < ENCW_POS_PER_HALF; ++k) {
    const int t = blockIdx.x * ENCW_POS_PER_BLOCK + k * ENCW_HALVES + half;
    if (t >= n) break;                                     // uniform within the half warp
    const u64* w = pos + (size_t)t * POSITION_WORDS;       // 72 bytes, broadcast loads
    Position p;
    p.pawns = ld_global_nc_u64(w + 0); p.knights = ld_global_nc_u64(w + 1); p.bishops = ld_global_nc_u64(w + 2);
    p.rooks = ld_global_nc_u64(w + 3); p.queens = ld_global_nc_u64(w + 4); p.kings = ld_global_nc_u64(w + 5);
    p.occ_w = ld_global_nc_u64(w + 6); p.occ_b = ld_global_nc_u64(w + 7); p.state = ld_global_nc_u64(w + 8);
    if (planes) half_write_planes(p, planes + (size_t)t * (19 * 64), hl);
    if (!mask) continue;

    for (int i = hl; i < MASK_BIT_WORDS / 4; i += 16) reinterpret_cast<uint4*>(bits)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp(hmask);
#ifndef ENC_EXP_STORES_ONLY   // timing experiment (tools/scratch): the store pattern without the move logic
    // danger map (squares the opponent attacks) and checkers: one opposing piece per lane, OR-reduced over the 16 lanes
    u64 danger = 0, checkers = 0;
    {
      const u64 theirs = pos_them(p), own_kings = p.kings & pos_us(p);
      const u64 king_bb = own_kings ? sq_bb(msb(own_kings)) : 0;
      const int enemies = popcnt(theirs);
      for (int j = hl; j < enemies; j += 16) {
        const int e = nth_set_bit(theirs, j);
        const u64 a = attacks_from_lut(p, e, s_knight, s_king);
        danger |= a;
        if (a & king_bb) checkers |= sq_bb(e);
      }
#pragma unroll
      for (int o = 8; o; o >>= 1) {
        danger |= __shfl_xor_sync(hmask, danger, o);
        checkers |= __shfl_xor_sync(hmask, checkers, o);
      }
    }
    const LegalCtx c = make_legal_ctx(p, &checkers);       // the same for the 16 lanes
    // steps of the king: the k-th safe square goes to lane k (at most 8)
    {
      const u64 steps = c.king_cand & ~danger;
      if (hl < popcnt(steps)) set_mask_bit(bits, s_tab[c.king * 64 + nth_set_bit(steps, hl)]);
    }
    // every other own piece: lane hl takes pieces hl, hl + 16, ...
    const int pieces = popcnt(c.ours);
    for (int j = hl; j < pieces; j += 16) {
      const int from = nth_set_bit(c.ours, j);
      const bool pawn = (p.pawns >> from) & 1;
      u64 tg = piece_targets_given(p, c, from, pawn ? 0 : attacks_from_lut(p, from, s_knight, s_king));
      while (tg) {
        int to = lsb(tg);
        tg &= tg - 1;
        set_mask_bit(bits, s_tab[from * 64 + to]);
        if (pawn && ((to >> 3) == 0 || (to >> 3) == 7)) {  // + the three under-promotions (the queen shares the plain slot)
          for (int pr = PT_KNIGHT; pr <= PT_ROOK; ++pr) set_mask_bit(bits, policy_index(make_move(from, to, pr), c.us));
        }
      }
      if (pawn && ep_capture_legal(p, c, from)) set_mask_bit(bits, s_tab[from * 64 + c.ep]);
    }
    if (hl == 15 && !c.checkers) {                         // castling: the lane least likely to hold a piece
      int ksq = 0, to[2];
      // (an out-of-line cold path for the exact tests, by reference or by value, doubled the kernel time: keep it inline)
      int nc = legal_castling(p, c, &ksq, to, &danger);
      for (int i = 0; i < nc; ++i) set_mask_bit(bits, s_tab[ksq * 64 + to[i]]);
    }
#endif
    __syncwarp(hmask);
    // bits -> bytes: every lane expands 16 bits into one 16-byte store, 256 contiguous bytes per half warp and step
    uint4* row = reinterpret_cast<uint4*>(mask + (size_t)t * POLICY_SIZE);
    const u16* b16 = reinterpret_cast<const u16*>(bits);
#pragma unroll 1
    for (int i0 = 0; i0 < MASK_CHUNKS; i0 += 16) {         // 18 full steps and a 4-lane tail
      const int i = i0 + hl;
      if (i0 + 16 <= MASK_CHUNKS || i < MASK_CHUNKS) {
        u32 v = b16[i];
        const uint2 lo = s_spread[v & 255u], hi8 = s_spread[v >> 8];
        st_global_cs_u4(row + i, make_uint4(lo.x, lo.y, hi8.x, hi8.y));
      }
    }
    __syncwarp(hmask);
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
  if (!d_moves && !d_idx && !d_counts) {                   // planes and / or mask only: the half-warp-per-position kernel
    if (!d_planes && !d_mask) return 0;
    if (d_mask) {
      static bool tab_ready[64] = {};
      int dev = 0;
      M0_CUDA_TRY(cudaGetDevice(&dev));
      if (dev < 0 || dev >= 64 || !tab_ready[dev]) {       // built once per device, before the first use on ANY stream:
        build_move_index_tab_kernel<<<16, 256, 0, (cudaStream_t)stream>>>();
        int rc = m0_check_launch("build_move_index_tab_kernel");
        if (rc != 0) return rc;
        // later calls may come on other streams (the game recorder encodes on its side stream), so the one-time build is waited for
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing((cudaStream_t)stream, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
        if (cap == cudaStreamCaptureStatusNone) {
          M0_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
          if (dev >= 0 && dev < 64) tab_ready[dev] = true;
        }
      }
    }
    int blocks = (n + ENCW_POS_PER_BLOCK - 1) / ENCW_POS_PER_BLOCK;
    encode_mask_planes_kernel<<<blocks, ENCW_THREADS, 0, (cudaStream_t)stream>>>(d_pos, n, d_planes, d_mask);
    return m0_check_launch("m0_encode_positions");
  }
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
