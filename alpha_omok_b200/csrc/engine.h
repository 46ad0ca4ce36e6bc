// Internal (non-ABI) definitions shared by tree.cu, tower.cu, rules.cu and engine.cu.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "alpha_omok_b200.h"

namespace ao {

constexpr int kMaxB = 15;
constexpr int kMaxA = 225;
constexpr int kRowsPad = 16;  // row masks per plane (>= kMaxB)

// ---- game status
enum : int32_t { ST_FREE = 0, ST_SEARCH = 1, ST_WAIT_NN = 2, ST_SEARCH_DONE = 3, ST_FINISHED = 4, ST_ERROR = 5 };

// child-slot codes
constexpr int32_t CH_UNVISITED = -1;  // terminal children: -(1 + win_index)  (<= -2)

struct __align__(16) Game {
  // root position
  uint16_t rows_b[kRowsPad];
  uint16_t rows_w[kRowsPad];
  uint8_t moves[kMaxA + 3];  // action of ply 1..n_moves at moves[0..n_moves)
  int32_t n_moves;
  int32_t last1, last2;      // last / second-last move at the root (-1 none)
  // search state
  int32_t status;
  int32_t arena;             // 0/1: which half of the tree storage is live
  int32_t root_node;         // slot offset of the root's child block, CH_UNVISITED, or terminal code
  uint32_t root_n;
  float root_w;
  uint32_t slot_count;       // bump pointer in the live arena
  int32_t sims_done, sims_target;
  int32_t is_real_root;
  int32_t auto_play;
  uint32_t rng_ctr, noise_draws, game_key;
  // pending leaf (between select and expand)
  uint16_t leaf_rows_b[kRowsPad];
  uint16_t leaf_rows_w[kRowsPad];
  int32_t leaf_depth;        // path length
  int32_t leaf_n_moves;
  int32_t nn_slot;
  uint32_t nn_ticket;        // ticket of the pending request in its weight set's queue (answered once head > ticket)
  int32_t nn_static;         // the pending request sits in the game's own slot (persistent kernel) and is answered
  int32_t winner;
  int32_t error;
  uint32_t nn_log_count;
  // arena mode (auto_play == 4), kept in the PLAYER slot of a match: side to move (0 player, 1 enemy), colours, index
  // of the match this slot is playing (eval_main.main plays N_MATCH matches one after the other, colours swapped)
  int32_t arena_cur, arena_player_black, arena_match;
  unsigned long long sims_total, nn_evals, terminal_sims, moves_played, games_finished;
};

// NN request: the five input planes of utils.get_state_pt as row bit-masks
struct __align__(16) LeafIn {
  uint16_t plane[4][kRowsPad];
  uint32_t colour;  // 1 iff black to move
  int32_t game;
  uint32_t pad[2];
};

struct TreeParams {
  int B, A, G;
  int num_mcts, noise, tau_thres, eval_mode, noise_mode;
  double c_puct, alpha;
  uint32_t seed_lo, seed_hi;
  uint32_t slot_cap;  // slots per arena
  uint32_t gc_cap;    // queue entries per game
  int nn_log_cap;
  int tape_rows;      // rows of gamma tape per game
  // storage
  Game* games;
  uint8_t* slot_act;
  uint2* slot_nw;    // {n, float_as_uint(w)}
  double* slot_p;
  int32_t* slot_child;
  uint32_t* path;    // [G][A+1] slot indices (absolute within the engine arrays)
  uint32_t* gc_old;  // [G][gc_cap] packed (L << 20 | old offset)
  uint32_t* gc_new;  // [G][gc_cap]
  uint32_t* rec_visits;  // [G][A][A]
  double* gamma_tape;    // [G][tape_rows][A] or null
  // NN exchange
  LeafIn* nn_in;         // [G]
  float* nn_policy;      // [G][A]
  float* nn_value;       // [G]
  // request queues, one per weight set: rings of nn_ring_mask + 1 slots inside nn_in / nn_policy / nn_value (set s at
  // offset s * (nn_ring_mask + 1)); the tree step issues tickets (nn_count = tails, monotonic), the tower serves them in
  // order and advances nn_head - possibly leaving a ragged tail of requests for the next round (NNQueue.defer)
  uint32_t* nn_count;    // [2] tickets issued
  uint32_t* nn_head;     // [2] tickets served
  uint32_t nn_ring_mask;
  int32_t* n_active;     // device counter
  float* nnlog_policy;   // [G][cap][A]
  float* nnlog_value;    // [G][cap]
  // continuous self-play (auto_play == 3): a slot whose episode ends packs its record at index (key - first_key) of
  // stream_out and takes the next unplayed key from the device counter until key_end is reached
  uint8_t* stream_out;         // [key_end - first_key][stream_rec_bytes]
  size_t stream_rec_bytes;
  uint32_t* stream_next_key;   // device counter
  uint32_t stream_first_key, stream_key_end;
  // two-kernel rounds: searches that are complete play their move (arg-max / sampling, win check, subtree compaction -
  // a few hundred microseconds of one warp) only in rounds with allow_moves set, i.e. every kMoveEvery-th round: once
  // games have drifted apart there would otherwise be a mover - and its latency - in EVERY round's tree step
  int allow_moves;
  // request slot = game slot instead of the next free slot of the round (the persistent self-play kernel evaluates game
  // g in a fixed CTA pass, see tower_stag.cu); 0 = dense packing through nn_count
  int static_slots;
  // arena (eval_main.py:204-333): matches [0, arena_M); side s (0 player, 1 enemy) of match m lives in game slot
  // s * arena_M + m with its own tree and decision stream; network requests of side s go to nn slots
  // [s * arena_M, ...) and are evaluated with weight set s.  0 = self-play / facade mode.
  int arena_M;
  int arena_matches_per_slot;   // consecutive matches a slot plays (colours swapped after each, streams continue)
  int arena_num_mcts[2];        // N_MCTS_PLAYER / N_MCTS_ENEMY (eval_main.py:33-34)
  int arena_kind[2];            // AO_SIDE_*: ZeroAgent | RandomAgent (agents.py:637-657) | PUCTAgent | UCTAgent (:263-634)
  int rollout_sims_per_round;   // PUCT / UCT sides run this many play-out simulations per lock-step round
  const double* log_table;      // log(k), k < log_table_n, as numpy computes it (UCT's exploration term)
  int log_table_n;
  uint32_t synth_salt[2];       // AO_EVAL_SYNTH: which synthetic "network" a side uses
};

// folded network parameters on the device (one weight set)
struct TowerWeights {
  const __half* conv_hi;   // packed UMMA B operands: stem [9][2][128][8] then 2*n_blocks x [9][16][128][8]
  const __half* conv_lo;   // low parts for the split (x3) mode, same layout (null in single-pass mode)
  const __half* conv_pair; // hi parts for CTA pairs: per layer and tap [2 cluster ranks][k-chunks][64 co][8]
  const __half* conv_pair_lo;  // low parts in the CTA-pair layout (split mode)
  const __half* conv_quad; // hi parts for clusters of four (tower_solo.cu): per layer [4 ranks][9 taps][k-chunks][32 co][8]
  const __half* conv_quad_x3;  // split mode: per layer [4 ranks][9 taps][k-chunks][32 co hi ; 32 co lo][8] (twice the size)
  const float* bias;       // [1 + 2*n_blocks][128]  BN-folded bias per conv layer
  const float* head_w;     // [3][128] policy c0, policy c1, value conv (BN scale folded)
  const float* head_b;     // [3]
  const float* pfc_wT;     // [2A][A]   policy FC transposed
  const float* pfc_b;      // [A]
  const float* vfc1_wT;    // [A][128]
  const float* vfc1_b;     // [128]
  const float* vfc2_w;     // [128]
  float vfc2_b;
  int n_layers;            // 1 + 2*n_blocks
#ifdef AO_PROBE           // probe build only (libalpha_omok_b200_probe.so): never part of the product library
  unsigned long long* dbg; // optional profiling counters of CTA 0 (null = off), see ao_tower_debug
  int xflags;              // timing experiments (env AO_TOWER_XFLAGS, results become wrong): see tower_stag_kernel
#endif
};

// Instrumentation hooks of the tower kernels: compiled in only with -DAO_PROBE.
#ifdef AO_PROBE
#define AO_DBG(...) __VA_ARGS__
#define AO_XFLAG(W, bit) (((W).xflags & (bit)) != 0)
#else
#define AO_DBG(...)
#define AO_XFLAG(W, bit) false
#endif

// Request queue of one weight set as the tower sees it.  tail == nullptr: no queue, serve slots [0, n_max).
struct NNQueue {
  const uint32_t* tail;  // tickets issued so far (device)
  uint32_t* head;        // tickets served so far (device); advanced by the last CTA of a tower launch to finish
  uint32_t* done;        // CTAs of the running launch that have finished
  uint32_t mask;         // ring capacity - 1
  int defer;             // leave a ragged last wave (< 60 % full) for the next round instead of paying a whole pass for it
};

// host-side launchers implemented in the .cu files
cudaError_t launch_tree_step(const TreeParams& p, const int32_t* game_ids, int n, int max_iters, cudaStream_t s);
cudaError_t launch_set_roots(const TreeParams& p, const int32_t* game_ids_dev, int n, const int16_t* roots_dev,
                             const int32_t* lens_dev, cudaStream_t s);
cudaError_t launch_export_roots(const TreeParams& p, const int32_t* game_ids_dev, int n, uint32_t* visits_dev,
                                double* priors_dev, int32_t* real_root_dev, cudaStream_t s);
cudaError_t launch_reset_games(const TreeParams& p, const int32_t* game_ids_dev, int n, const uint32_t* keys_dev,
                               int auto_play, cudaStream_t s);
cudaError_t launch_sum_counters(const TreeParams& p, int n, int n_running, unsigned long long* out5_dev, cudaStream_t s);
cudaError_t launch_reset_arena(const TreeParams& p, int n_slots, uint32_t first_key, cudaStream_t s);
cudaError_t launch_rollout_search(const TreeParams& p, int kind, int num_mcts, const int32_t* game_ids_dev, int n,
                                  const int16_t* roots_dev, const int32_t* lens_dev, uint32_t* visits_dev, float* w_dev,
                                  cudaStream_t s);
cudaError_t launch_pack_records(const TreeParams& p, int n, uint8_t* out, size_t bytes_per_game, cudaStream_t s);

cudaError_t launch_tower(const TowerWeights& w, int B, int precision, const LeafIn* in, const NNQueue& q,
                         int n_max, float* policy, float* value, int num_sms, cudaStream_t s);
cudaError_t launch_tower_stag(const TowerWeights& w, int B, const LeafIn* in, const NNQueue& q, int n_max,
                              float* policy, float* value, int num_sms, cudaStream_t s);
cudaError_t launch_selfplay_persist(const TowerWeights& w, int B, const TreeParams& p, int n_games, int rounds,
                                    int num_sms, uint32_t* cta_pos, cudaStream_t s);
cudaError_t launch_selfplay_solo(const TowerWeights& w, int B, int precision, const TreeParams& p, int n_games, int rounds,
                                 cudaStream_t s);
bool solo_supports(int B, int precision);
int solo_max_games(int num_sms);
cudaError_t tower_configure(int B, int precision);
cudaError_t launch_pack_states(const float* states_dev, int n, int B, int inplanes, LeafIn* out, int* bad_flag_dev,
                               cudaStream_t s);

cudaError_t launch_check_win(const int8_t* boards_dev, int n, int B, uint8_t* out_dev, cudaStream_t s);
cudaError_t launch_encode_state(const int16_t* ids_dev, const int32_t* lens_dev, int n, int B, float* out_dev,
                                cudaStream_t s);
cudaError_t launch_legal_actions(const int16_t* ids_dev, const int32_t* lens_dev, int n, int B, int16_t* out_dev,
                                 cudaStream_t s);

}  // namespace ao
