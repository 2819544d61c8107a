// Self-play game-loop state (selfplay_kernels.cu) -- mirrors the per-game locals of
// azchess/selfplay/internal.py:334-365.
#pragma once
#include "engine.cuh"

namespace m0 {

static constexpr int SP_WINDOW = 8;

struct SelfPlayParams {
  double temperature_start, temperature_end;  // internal.py:347-349
  double resign_threshold, resign_min_entropy, resign_value_margin;  // :352-360
  int temperature_moves, max_game_len, min_resign_plies, resign_window, resign_consecutive_bad, opening_random_plies;
  unsigned long long seed;
  int argmax_after_plies;  // >= 0: arena rule (arena.py:75-91): temperature_start while plies < this, then argmax; < 0: self-play schedule
  int reserved;
};

struct FinishedGame {
  int game;       // slot
  int plies;      // len(states)
  float z;        // result from White's point of view (internal.py:587-599)
  int reason;     // END_* (selfplay_kernels.cu)
  float avg_entropy;
};

struct SelfPlayState {
  const SelfPlayParams* params;
  int* ply;                 // [G] len(states)
  int* consec_bad;          // [G]
  int* recent_n;            // [G]
  int* ent_n;               // [G]
  int* ent_total;           // [G]
  float* recent_val;        // [G][SP_WINDOW]
  float* recent_ent;        // [G][SP_WINDOW]
  double* ent_sum;          // [G]
  double* last_value;       // [G]
  int* games_started;       // [G]
  u8* need_start;           // [G]
  FinishedGame* finished;   // ring [finished_cap]
  unsigned* finished_count; // total finished so far
  int finished_cap;
};

}  // namespace m0
