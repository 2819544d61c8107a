/* matrix0_b200 -- C ABI of the B200-native self-play search engine.
 *
 * The reference (lukifer23/Matrix0) is pure Python and has no FFI; its seams for this path are the
 * Python duck-types listed in SURVEY.md section 8b.  This header is the boundary a binding for
 * those seams links against: plain pointers and sizes, no torch / C++ types.  Every function
 * returns 0 on success or a negative M0_ERR_* code; m0_last_error() then returns a thread-local,
 * library-owned message.  Unless stated otherwise all buffer pointers are DEVICE pointers supplied
 * by the caller (e.g. torch tensor.data_ptr()); the library never allocates memory the caller
 * sees, never synchronises the host unless the function name ends in _host or _sync, and launches
 * on the `stream` argument (a cudaStream_t passed as void*).  There is no CPU implementation
 * behind any entry point.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference).
 */
#ifndef MATRIX0_B200_H
#define MATRIX0_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M0_OK 0
#define M0_ERR_CUDA (-1)     /* a CUDA runtime call or launch failed */
#define M0_ERR_ARG (-2)      /* invalid argument */
#define M0_ERR_STATE (-3)    /* call sequence error (e.g. search step without begin) */
#define M0_ERR_CAPACITY (-4) /* node pool / table capacity exhausted */

#define M0_POSITION_WORDS 9  /* packed position: 8 bitboards + state word (csrc/chess_core.cuh) */
#define M0_RAW_WORDS 10      /* raw python-chess record, see m0_positions_pack */
#define M0_MAX_MOVES 256     /* row stride of move / index lists */
#define M0_POLICY_SIZE 4672  /* azchess/encoding.py:51 POLICY_SHAPE = (8, 8, 73) */
#define M0_PLANES 19         /* azchess/encoding.py:11 */

/* ---- library ------------------------------------------------------------------------------ */
const char* m0_last_error(void);
int m0_version(void);
int m0_device_count(void);            /* <= 0: the product cannot run (no CPU path) */
int m0_device_sm_count(int device);

/* ---- positions ------------------------------------------------------------------------------
 * Raw record per position (uint64[10]), the public fields of python-chess's chess.Board that the
 * reference reads (encoding.py:22-33, mcts.py:342):
 *   [0..5] pawns, knights, bishops, rooks, queens, kings   [6] occupied_co[WHITE]  [7] occupied_co[BLACK]
 *   [8] castling_rights (rook-square mask)   [9] turn | ep_square<<8 (255 = None) | halfmove_clock<<16 | fullmove_number<<32
 * m0_positions_pack applies Board.clean_castling_rights() and writes packed positions uint64[n][9]. */
int m0_positions_pack(const uint64_t* d_raw, int n, uint64_t* d_pos, void* stream);
/* Synthetic positions for benches/tests: position i = plies_i (hash(seed,i) % (max_plies+1)) uniformly
 * random legal moves from the start position -- batch form of azchess/utils/board.py:7-38. */
int m0_random_playouts(uint64_t* d_pos, int n, uint64_t seed, int max_plies, void* stream);

/* ---- encoding: azchess/encoding.py ------------------------------------------------------------
 * m0_encode_positions is the fused kernel; any output pointer may be NULL.
 *   d_planes float32[n][19][8][8]   = encode_board          (encoding.py:11-37, row = 7 - rank)
 *   d_mask   uint8[n][4672]         = MoveEncoder.get_legal_actions (encoding.py:243-253)
 *   d_moves  uint16[n][256]         = list(board.legal_moves) in python-chess generation order,
 *                                     packed from | to<<6 | promotion<<12 (mcts.py:140)
 *   d_idx    uint16[n][256]         = move_to_index of each move (encoding.py:113-150)
 *   d_counts int32[n]               = number of legal moves */
int m0_encode_positions(const uint64_t* d_pos, int n, float* d_planes, uint8_t* d_mask, uint16_t* d_moves,
                        uint16_t* d_idx, int32_t* d_counts, void* stream);
int m0_encode_planes(const uint64_t* d_pos, int n, float* d_planes, void* stream);
int m0_legal_mask(const uint64_t* d_pos, int n, uint8_t* d_mask, void* stream);
int m0_legal_moves(const uint64_t* d_pos, int n, uint16_t* d_moves, uint16_t* d_idx, int32_t* d_counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MATRIX0_B200_H */
