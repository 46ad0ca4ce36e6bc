/* alpha_omok_b200 - C ABI of the B200-native self-play hot path (MCTS + rules + policy-value network inference).
 *
 * The reference (reinforcement-learning-kr/alpha_omok) has no FFI: its boundary is duck-typed Python.  Every entry
 * point below names the reference interface it replaces (file:line relative to 2_AlphaOmok/).  The Python facades in
 * alpha_omok_b200/{agents,utils,model}.py and alpha_omok_b200/env/ bind these through ctypes and keep the reference's
 * names, argument meaning and return values.  All pointers are HOST pointers unless the name says `_dev`;
 * buffers are caller-owned; every function returns 0 on success or a negative error code, and ao_last_error()
 * returns a description of the last failure on the calling thread.  One engine = one CUDA device + one stream;
 * an engine is not thread-safe.
 */
#ifndef ALPHA_OMOK_B200_H
#define ALPHA_OMOK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ao_engine ao_engine;

enum { AO_EVAL_PVNET = 0, AO_EVAL_SYNTH = 1 };    /* synthetic hash "network": exact floats, used by parity tests */
enum { AO_NOISE_DEVICE = 0, AO_NOISE_TAPE = 1 };  /* Dirichlet gammas: on-device Philox generator, or host tape    */
enum { AO_SIDE_ZERO = 0,     /* agents.ZeroAgent (network-guided PUCT search)                    eval_main.py:85-101 */
       AO_SIDE_RANDOM = 1,   /* agents.RandomAgent (agents.py:637-657)                            eval_main.py:70-72  */
       AO_SIDE_PUCT = 2,     /* agents.PUCTAgent: uniform prior + random play-outs (:263-453)      eval_main.py:73-75  */
       AO_SIDE_UCT = 3 };    /* agents.UCTAgent: UCB1 + random play-outs (:456-634)                eval_main.py:76-78  */
enum { AO_NN_FP16 = 0,       /* fp16 operands, one MMA per k-step, issued by CTA pairs (tcgen05 cta_group::2), the
                                two tiles of a CTA staggered so that epilogues hide behind MMAs (tower_stag.cu)    */
       AO_NN_FP16X3 = 1,     /* hi/lo split operands, 3 MMAs per k-step (1e-4 on trained nets), CTA pairs       */
       AO_NN_FP16_1CTA = 2,  /* one CTA per MMA (cta_group::1), MMA and epilogue in lock-step; kept for comparison */
       AO_NN_FP16_LOCKSTEP = 3 };/* CTA pairs with MMA and epilogue in lock-step (the round-1 v5 kernel); comparison */

/* Mirrors the module-level constants of main.py:26-45 / eval_main.py:22-51 and ZeroAgent.__init__ (agents.py:39-53). */
typedef struct ao_config {
  int32_t device;       /* CUDA ordinal                                             */
  int32_t board_size;   /* BOARD_SIZE: 9 (env_small) or 15 (env_regular); 5..15     */
  int32_t inplanes;     /* IN_PLANES = 5                                            */
  int32_t planes;       /* OUT_PLANES = 128 (only 128 is implemented)               */
  int32_t n_blocks;     /* N_BLOCKS = 10                                            */
  int32_t num_mcts;     /* N_MCTS (400 self-play / 800 arena)                       */
  int32_t noise;        /* ZeroAgent(noise=...)                                     */
  int32_t tau_thres;    /* TAU_THRES = 6                                            */
  int32_t max_games;    /* concurrent game slots resident in HBM                    */
  int32_t node_cap;     /* expanded nodes per game tree arena (0 = default 2048)    */
  int32_t eval_mode;    /* AO_EVAL_*                                                */
  int32_t noise_mode;   /* AO_NOISE_*                                               */
  int32_t nn_precision; /* AO_NN_*                                                  */
  int32_t nn_log_cap;   /* per-game capacity of the NN-output log (0 = off)         */
  double c_puct;        /* 5 (agents.py:48)                                         */
  double alpha;         /* 10 / B^2 (agents.py:47); <= 0 selects that default        */
  uint64_t seed;        /* key of the per-game Philox decision streams              */
  void* stream;         /* cudaStream_t to launch on, or NULL to create one         */
} ao_config;

const char* ao_last_error(void);
int ao_engine_create(const ao_config* cfg, ao_engine** out);
int ao_engine_destroy(ao_engine* h);

/* Agent.model = PVNet(...) / load_state_dict (main.py:81,356-358; eval_main.py:91-101).  fp32 tensors straight from
 * state_dict(): names[i] is the state_dict key, ptrs[i] its data, numel[i] its element count.  BatchNorm (eval mode,
 * eps 1e-5) is folded here; missing `num_batches_tracked` keys are fine. */
int ao_load_weights(ao_engine* h, int n_tensors, const char* const* names, const float* const* ptrs,
                    const int64_t* numel);

/* Second weight set: eval_main.Evaluator.set_agents builds TWO networks, player and enemy (eval_main.py:87-101).
 * set 0 = Agent.model / the player (what ao_load_weights fills), set 1 = the arena's enemy. */
int ao_load_weights_set(ao_engine* h, int set, int n_tensors, const char* const* names, const float* const* ptrs,
                        const int64_t* numel);

/* ZeroAgent.reset() + a fresh GameState (agents.py:55-58, main.py:136).  game_keys[i] selects the decision stream. */
int ao_games_reset(ao_engine* h, const int32_t* game_ids, int n, const uint32_t* game_keys);

/* Parity protocol: raw Gamma(alpha,1) variates consumed by the Dirichlet draws of one game,
 * tape[n_draws][A] float64.  Only read when noise_mode == AO_NOISE_TAPE. */
int ao_set_gamma_tape(ao_engine* h, int game_id, const double* tape, int n_draws);

/* ZeroAgent.get_pi minus the final visit -> pi arithmetic (agents.py:60-76): runs num_mcts (+1 on a real root)
 * simulations for each listed game from the given root ID.  roots[i*(A+1) .. ] holds the ID tuple (leading 0
 * included), root_lens[i] its length.  Outputs (any may be NULL): visits[n][A] = child n, priors[n][A] = child p
 * (noise-mixed, float64), is_real_root[n]. */
int ao_search(ao_engine* h, const int32_t* game_ids, int n, const int16_t* roots, const int32_t* root_lens,
              uint32_t* visits, double* priors, int32_t* is_real_root);

/* ZeroAgent.get_pv (agents.py:252-260) / PVNet.forward (model.py:97-104) on explicit states:
 * states float32 [n][inplanes][B][B] with {0,1} entries (utils.get_state_pt layout) -> p [n][A], v [n]. */
int ao_nn_forward(ao_engine* h, const float* states, int n, float* p, float* v);
int ao_nn_forward_set(ao_engine* h, int set, const float* states, int n, float* p, float* v);

/* Batched twin of main.self_play (main.py:122-250): every game slot plays one episode on the device
 * (get_pi -> get_action -> env.step -> root advance), decisions from the per-game stream.
 *   ao_selfplay_begin : reset all slots [0, n_games) with game keys first_key + g.
 *   ao_selfplay_begin_mode(recycle=1): a slot whose episode ends immediately starts the next one (key += max_games)
 *                       - steady-state throughput runs; per-episode records are then not kept.
 *   ao_selfplay_rounds: run up to `rounds` lock-step rounds (select -> NN -> expand/backup [-> move]) and return
 *                       totals in out[8]: [0] simulations completed, [1] games still running, [2] NN evaluations,
 *                       [3] games in error (tree arena overflow), [4] moves played, [5] episodes finished,
 *                       [6] terminal-leaf simulations, [7] reserved.
 *   ao_selfplay_fetch : moves[n][A] (int16, -1 padded), n_moves[n], winners[n] (0 running,1,2,3),
 *                       visits[n][A][A] uint32 per ply (may be NULL). */
int ao_selfplay_begin(ao_engine* h, int n_games, uint32_t first_key);
int ao_selfplay_begin_mode(ao_engine* h, int n_games, uint32_t first_key, int recycle);
int ao_selfplay_rounds(ao_engine* h, int rounds, uint64_t* out8);
int ao_selfplay_fetch(ao_engine* h, int n_games, int16_t* moves, int32_t* n_moves, int8_t* winners,
                      uint32_t* visits);
/* Continuous self-play: n_slots concurrent games play the episodes with decision-stream keys [first_key, first_key +
 * n_episodes); a slot whose episode ends packs its record on the device and takes the next unplayed key (main.py:132
 * `for episode in range(n_selfplay)` with more episodes than game slots).  Drive it with ao_selfplay_rounds until no
 * game is running; ao_selfplay_stream_records_dev then returns the DEVICE slab of n_episodes records in ao_records_dev's
 * layout, index = key - first_key.  Episodes are functions of their key alone: the records equal those of
 * ao_selfplay_begin runs with the same keys.  Needs noise_mode == AO_NOISE_DEVICE (or noise off). */
int ao_selfplay_stream_begin(ao_engine* h, int n_slots, uint32_t first_key, int n_episodes);
int ao_selfplay_stream_records_dev(ao_engine* h, void** dev_ptr, size_t* bytes_per_game, int* n_episodes);
/* ao_selfplay_rounds with CUDA events around every launch: summed device milliseconds of the tree-step kernels and of
 * the tower kernels over the `rounds` rounds (bench.py's roofline numbers). rounds <= 4096. */
int ao_selfplay_rounds_timed(ao_engine* h, int rounds, uint64_t* out8, float* tree_ms, float* tower_ms);
/* Switch the tower's operand mode (AO_NN_*) at run time (per weight set). */
int ao_set_nn_precision(ao_engine* h, int mode);
int ao_set_nn_precision_set(ao_engine* h, int set, int mode);

/* eval_main.main's match loop (eval_main.py:204-333; Evaluator.get_action :153-170) on the device.  Every slot plays
 * `matches_per_slot` consecutive matches like one run of eval_main.main: player (weight set 0, n_mcts_player sims) vs
 * enemy (weight set 1 with n_mcts_enemy sims); player_kind / enemy_kind = AO_SIDE_* choose the agent class of a side
 * as eval_main.Evaluator.set_agents does by name (ZeroAgent, RandomAgent, PUCTAgent, UCTAgent); per ply
 * get_pi(root_id, tau=0) -> argmax_onehot -> root_id = mover.root_id + (action,) -> env.step; each side keeps its own
 * tree, so the other side's next root is a reused root (possibly never visited, n == 0) or a real root
 * (agents.py:82-111); colours swap after every match (the player is black first in even slots), both agents are reset,
 * their decision streams (keys first_key + 2*slot + side) run on.  Needs max_games >= 2 * n_slots (one game slot per
 * side).  Drive it with ao_selfplay_rounds (out[1] = slots still playing, out[5] = matches finished); with
 * keep_records, ao_selfplay_stream_records_dev returns n_slots * matches_per_slot records (ao_records_dev layout,
 * index = slot * matches_per_slot + match; visits[ply] = the mover's root visit counts, zeros for a RandomAgent ply;
 * the pad byte after `winner` holds 1 when the player was black). n_mcts_* <= 0 selects ao_config.num_mcts. */
int ao_arena_begin(ao_engine* h, int n_slots, uint32_t first_key, int matches_per_slot, int player_kind,
                   int enemy_kind, int keep_records, int n_mcts_player, int n_mcts_enemy);

/* PUCTAgent / UCTAgent.get_pi(root_id, board, turn, tau) minus the final arg-max (agents.py:283-296, 461-476): every
 * call starts a fresh tree, runs num_mcts + 1 simulations (uniform prior or UCB1 selection, one uniformly random
 * play-out per non-root leaf, agents.py:319-432 / 503-613) and returns visits[n][A] = n(child) and w[n][A] = w(child)
 * (q = w / n; w is integer-valued).  kind = AO_SIDE_PUCT | AO_SIDE_UCT; num_mcts <= 0 selects ao_config.num_mcts. */
int ao_rollout_search(ao_engine* h, int kind, const int32_t* game_ids, int n, const int16_t* roots,
                      const int32_t* root_lens, int num_mcts, uint32_t* visits, float* w);
/* table[k] = log(k) for k < n (k = 0 unused) as the caller's numpy computes it: UCT compares its exploration terms for
 * equality, so the device takes the logarithms from the host instead of its own libm.  Needed before AO_SIDE_UCT. */
int ao_set_log_table(ao_engine* h, const double* table, int n);
/* Number of kernels this engine has launched so far (bench.py's gpu_launches). */
int ao_launch_count(ao_engine* h, uint64_t* out);

/* NN-output log of one game (parity protocol: the oracle replays these floats): policy[count][A], value[count]. */
int ao_get_nn_log(ao_engine* h, int game_id, float* policy, float* value, int32_t capacity, int32_t* count);

/* Replay records on the device for the multi-GPU all-gather: a slab of fixed-size records
 * {int16 n_moves, int8 winner, pad, int16 moves[A], uint32 visits[A][A]} per game; pointer valid until destroy. */
int ao_records_dev(ao_engine* h, void** dev_ptr, size_t* bytes_per_game);
int ao_records_pack(ao_engine* h, int n_games);

/* utils.augment_dataset + the (state, pi, z) assembly of main.self_play on the device (utils.py:226-239,
 * main.py:155-227): record slab (device, n_games records, e.g. the all-gathered one) -> float32 training tensors
 * states_dev [8*plies][5][B][B], pi_dev [8*plies][A], z_dev [8*plies] (DEVICE pointers, caller-allocated, capacity in
 * samples; pass NULL outputs to only count). *n_samples_out (host) = 8 * total plies. `stream`: cudaStream_t or NULL. */
int ao_augment_records_dev(const void* slab_dev, int n_games, int board_size, int tau_thres, float* states_dev,
                           float* pi_dev, float* z_dev, long long capacity_samples, long long* n_samples_out,
                           void* stream);

/* rep_memory = deque(maxlen=MEMORY_SIZE); rep_memory.extend(utils.augment_dataset(cur_memory, BOARD_SIZE))
 * (main.py:66,250) on the device: the augmented samples of the record slab go straight into a caller-allocated ring
 * ring_states_dev [cap][5][B][B], ring_pi_dev [cap][A], ring_z_dev [cap] (DEVICE).  *head_io / *len_io (HOST) are the
 * deque's state - logical item i (0 = oldest) lives in slot (head + i) % cap - and are updated with deque semantics
 * (the oldest items fall out).  *n_samples_out (host) = samples appended by this call. */
int ao_replay_extend_dev(const void* slab_dev, int n_games, int board_size, int tau_thres, float* ring_states_dev,
                         float* ring_pi_dev, float* ring_z_dev, long long ring_cap, long long* head_io,
                         long long* len_io, long long* n_samples_out, void* stream);
/* train_memory = random.sample(rep_memory, k) (main.py:263-264) as a device gather: idx_dev[k] (DEVICE, int64) are the
 * logical deque indices drawn on the host; out_* (DEVICE) receive the samples in that order. Asynchronous on `stream`. */
int ao_replay_gather_dev(const float* ring_states_dev, const float* ring_pi_dev, const float* ring_z_dev,
                         long long ring_cap, long long head, const long long* idx_dev, long long k, int board_size,
                         float* out_states_dev, float* out_pi_dev, float* out_z_dev, void* stream);

int ao_synchronize(ao_engine* h);

/* Stateless unit-test entry points (device 0 unless an engine was created).
 * utils.check_win (utils.py:30-59): boards int8 [n][B*B] (+1 black, -1 white) -> out uint8 [n]. */
int ao_check_win(const int8_t* boards, int n, int board_size, uint8_t* out);
/* utils.get_state_pt (utils.py:139-168): ids int16 [n][A+1] (-1 padded) -> float32 [n][5][B][B]. */
int ao_encode_state(const int16_t* ids, const int32_t* lens, int n, int board_size, float* out);
/* utils.legal_actions (utils.py:22-27) incl. the CPython set-order regime: -> int16 [n][A] (-1 padded). */
int ao_legal_actions(const int16_t* ids, const int32_t* lens, int n, int board_size, int16_t* out);
#ifdef __cplusplus
}
#endif
#endif /* ALPHA_OMOK_B200_H */
