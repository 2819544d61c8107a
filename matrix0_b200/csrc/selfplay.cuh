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
  int low_visit_threshold; // internal.py:419-425: max visits below this -> temperature at least 0.8 (0 = off)
  // should_adjudicate_draw heuristics (draw.py:43-82), active when draw_enabled
  int draw_enabled, draw_min_plies, draw_window, draw_min_unique, draw_halfmove_cap, draw_material_threshold;
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
  u16* hist_move;           // [G][hist_cap] moves played so far (move_history, internal.py:366-379, :538)
  int* start_budget;        // [1] games that may still be started (>= 2^30: unlimited)
  const double* uniforms;   // [G] caller-supplied np.random draws for the next advance (NULL: device generator)
  FinishedGame* finished;   // ring [finished_cap]
  unsigned* finished_count; // total finished so far
  int finished_cap;
};

}  // namespace m0
