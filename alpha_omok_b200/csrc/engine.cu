// Host side of the C ABI (include/alpha_omok_b200.h): engine life cycle, BN folding + UMMA weight packing,
// the lock-step round loop (tree step -> tower -> tree step ...), and the host<->device staging of the facade calls.
#include <cuda_fp16.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "engine.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define AO_CUDA(x)                                                                           \
  do {                                                                                       \
    cudaError_t e__ = (x);                                                                   \
    if (e__ != cudaSuccess) return fail(-100 - (int)e__, "%s: %s", #x, cudaGetErrorString(e__)); \
  } while (0)

// Every entry point runs with the engine's device current and restores the caller's device on the way out (an engine
// created for GPU k inside a process whose current device is j must neither fail nor silently switch the caller to k).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

template <class T>
cudaError_t dalloc(T** p, size_t count) {
  return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}

}  // namespace

struct ao_engine {
  ao_config cfg;
  int B, A, G;
  int num_sms;
  bool own_stream;
  cudaStream_t stream;
  ao::TreeParams tp;
  // two weight sets: 0 = Agent.model (self-play, facades, the arena's player), 1 = the arena's enemy (eval_main.py:87-101)
  struct WeightSet {
    ao::TowerWeights tw;
    bool loaded;
    int precision;  // AO_NN_*
    __half *d_conv_hi, *d_conv_lo, *d_conv_pair, *d_conv_pair_lo, *d_conv_quad, *d_conv_quad_x3;
    float *d_bias, *d_head_w, *d_head_b, *d_pfc_wT, *d_pfc_b, *d_vfc1_wT, *d_vfc1_b, *d_vfc2_w;
  } ws[2];
  std::vector<void*> allocs;
  // staging
  int32_t *d_ids, *d_lens, *d_real_root;
  int16_t* d_roots;
  uint32_t *d_keys, *d_visits;
  double* d_priors;
  unsigned long long* d_counters;
  uint8_t* d_records;
  size_t rec_bytes;
  int32_t* h_pinned;  // [0] n_active
  int selfplay_games;
  // continuous self-play: record slab of the episodes of one stream run (resized on demand), next-key counter
  uint8_t* d_stream;
  size_t stream_capacity;      // episodes the slab can hold
  uint32_t* d_stream_next;
  int stream_episodes;         // episodes of the current stream run (0 = none)
  // CUDA graph of kGraphRounds lock-step rounds (memsets + tree step + tower), re-captured whenever a kernel argument
  // changes; ao_selfplay_rounds replays it instead of issuing 4 stream operations per round
  cudaGraphExec_t round_graph;
  ao::TreeParams graph_tp;
  ao::TowerWeights graph_tw[2];
  int graph_n, graph_max_iters, graph_precision[2];
  unsigned long long graph_launches_per_round;
  bool graph_disabled;
  // optional per-kernel timing (bench roofline): event pairs recorded around every launch of a timed call
  bool timing;
  std::vector<cudaEvent_t> ev;  // [tree_begin, tree_end(=tower_begin), tower_end] per round
  unsigned long long launches;  // kernels launched by this engine (bench's gpu_launches)
  uint32_t ring_cap;            // request-ring capacity per weight set (power of two >= max_games)
  uint32_t* d_tower_done;       // [2] finished-CTA counters of the running tower launches
  bool defer_tail;              // towers leave a ragged last wave for the next round (AO_NO_DEFER=1 disables)
  unsigned long long round_no;  // two-kernel rounds issued so far (move rounds: round_no % kMoveEvery == 0)
  bool free_run;                // persistent kernel cycles through each CTA's games with full passes (AO_NO_FREERUN=1: static map)
  uint32_t* d_cta_pos;          // [512] cycle position per CTA
  float* d_fwd_states;          // ao_nn_forward staging (lazily allocated)
  int* d_fwd_bad;
  float* d_wsum;                // ao_rollout_search: w of the root's children
  int log_table_n;
  // persistent self-play kernel (tower_stag.cu, PERSIST): allowed unless AO_NO_PERSIST is set; `persist_pending` =
  // the running games wait in ST_WAIT_NN with their requests in nn_in[game] (static slots) and no answer computed yet
  bool persist_allowed, persist_pending;
  bool solo_allowed;            // a few games: cluster-of-four kernel (AO_NO_SOLO=1: always the CTA-pair kernel)
  long long last_running;       // running games after the last self-play call (-1: unknown = all)
  cudaEvent_t pev[3];
};

namespace {

template <class T>
int ealloc(ao_engine* h, T** p, size_t count) {
  T* q = nullptr;
  cudaError_t e = dalloc(&q, count ? count : 1);
  if (e != cudaSuccess) return fail(-2, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
  h->allocs.push_back(q);
  *p = q;
  return 0;
}

constexpr int kGraphRounds = 16;
constexpr int kMoveEvery = kGraphRounds;  // a graph replay = one move round followed by 15 plain rounds
constexpr int kPollEvery = 4;             // ao_search: host polls of the running-games counter

ao::NNQueue queue_of(const ao_engine* h, int set, bool defer) {
  ao::NNQueue q;
  q.tail = h->tp.nn_count + set;
  q.head = h->tp.nn_head + set;
  q.done = h->d_tower_done + set;
  q.mask = h->tp.nn_ring_mask;
  q.defer = defer && h->defer_tail ? 1 : 0;
  return q;
}
// all game slots are being reset: requests still queued belong to nobody - mark them served
cudaError_t drop_queued_requests(ao_engine* h) {
  return cudaMemcpyAsync(h->tp.nn_head, h->tp.nn_count, 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream);
}
ao::NNQueue no_queue() {
  ao::NNQueue q;
  memset(&q, 0, sizeof q);
  return q;
}

// one lock-step round: tree step (consume the answers that have arrived, select next leaves, queue their requests) then
// the tower(s) on the queued requests
int run_round(ao_engine* h, const int32_t* ids_dev, int n, int max_iters, int timed_slot = -1) {
  AO_CUDA(cudaMemsetAsync(h->tp.n_active, 0, sizeof(int32_t), h->stream));
  if (timed_slot >= 0) AO_CUDA(cudaEventRecord(h->ev[3 * timed_slot + 0], h->stream));
  ao::TreeParams tp = h->tp;
  // every kMoveEvery-th round is a move round (the synthetic evaluator plays whole games per launch: always)
  tp.allow_moves = (h->cfg.eval_mode == AO_EVAL_SYNTH || h->round_no % kMoveEvery == 0) ? 1 : 0;
  h->round_no += 1;
  AO_CUDA(ao::launch_tree_step(tp, ids_dev, n, max_iters, h->stream));
  if (timed_slot >= 0) AO_CUDA(cudaEventRecord(h->ev[3 * timed_slot + 1], h->stream));
  h->launches += 1;
  if (h->cfg.eval_mode == AO_EVAL_PVNET) {
    const int M = h->tp.arena_M;
    if (!(M > 0 && h->tp.arena_kind[0] != AO_SIDE_ZERO)) {
      AO_CUDA(ao::launch_tower(h->ws[0].tw, h->B, h->ws[0].precision, h->tp.nn_in, queue_of(h, 0, true), n,
                               h->tp.nn_policy, h->tp.nn_value, h->num_sms, h->stream));
      h->launches += 1;
    }
    if (M > 0 && h->tp.arena_kind[1] == AO_SIDE_ZERO) {  // the enemy's queue: second ring, weight set 1
      const size_t cap = h->ring_cap;
      AO_CUDA(ao::launch_tower(h->ws[1].tw, h->B, h->ws[1].precision, h->tp.nn_in + cap, queue_of(h, 1, true), n,
                               h->tp.nn_policy + cap * h->A, h->tp.nn_value + cap, h->num_sms, h->stream));
      h->launches += 1;
    }
  }
  if (timed_slot >= 0) AO_CUDA(cudaEventRecord(h->ev[3 * timed_slot + 2], h->stream));
  return 0;
}


// `rounds` lock-step rounds over the first n game slots; whole multiples of kGraphRounds are replayed from a CUDA graph
// (captured from run_round itself, so both paths launch exactly the same work), the rest is issued directly.
int run_rounds(ao_engine* h, int n, int max_iters, int rounds) {
  int rc;
  int done = 0;
  const bool graphable = !h->graph_disabled && rounds >= 2 * kGraphRounds;
  if (graphable) {
    bool stale = h->round_graph == nullptr || h->graph_n != n || h->graph_max_iters != max_iters ||
                 memcmp(&h->graph_tp, &h->tp, sizeof h->tp) != 0;
    for (int k = 0; k < 2; ++k)
      stale = stale || h->graph_precision[k] != h->ws[k].precision || memcmp(&h->graph_tw[k], &h->ws[k].tw, sizeof(ao::TowerWeights)) != 0;
    if (stale) {
      if (h->round_graph) cudaGraphExecDestroy(h->round_graph);
      h->round_graph = nullptr;
      // one direct round first: per-kernel attributes (dynamic smem opt-in) are set outside the capture
      const unsigned long long launches_before_direct = h->launches;
      if ((rc = run_round(h, nullptr, n, max_iters)) != 0) return rc;
      ++done;
      cudaGraph_t graph = nullptr;
      h->graph_launches_per_round = h->launches - launches_before_direct;
      const unsigned long long launches0 = h->launches;
      const unsigned long long round_no0 = h->round_no;
      h->round_no = 0;  // the captured sequence starts with a move round
      cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
      if (e == cudaSuccess) {
        for (int r = 0; r < kGraphRounds && rc == 0; ++r) rc = run_round(h, nullptr, n, max_iters);
        e = cudaStreamEndCapture(h->stream, &graph);
      }
      h->round_no = round_no0;
      h->launches = launches0;  // nothing ran during the capture
      if (e == cudaSuccess && rc == 0 && graph) e = cudaGraphInstantiate(&h->round_graph, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (e != cudaSuccess || rc != 0 || !h->round_graph) {  // fall back to direct launches for good
        cudaGetLastError();
        h->round_graph = nullptr;
        h->graph_disabled = true;
      } else {
        memcpy(&h->graph_tp, &h->tp, sizeof h->tp);
        for (int k = 0; k < 2; ++k) {
          memcpy(&h->graph_tw[k], &h->ws[k].tw, sizeof(ao::TowerWeights));
          h->graph_precision[k] = h->ws[k].precision;
        }
        h->graph_n = n;
        h->graph_max_iters = max_iters;
      }
    }
    for (; h->round_graph && done < rounds && h->round_no % kMoveEvery != 0; ++done)  // line up with the move rounds
      if ((rc = run_round(h, nullptr, n, max_iters)) != 0) return rc;
    while (h->round_graph && rounds - done >= kGraphRounds) {
      AO_CUDA(cudaGraphLaunch(h->round_graph, h->stream));
      h->launches += h->graph_launches_per_round * kGraphRounds;
      h->round_no += kGraphRounds;
      done += kGraphRounds;
    }
  }
  for (; done < rounds; ++done)
    if ((rc = run_round(h, nullptr, n, max_iters)) != 0) return rc;
  return 0;
}

// totals over the running self-play games, or over both sides of the running arena matches (running = matches)
cudaError_t launch_sum_selfplay(ao_engine* h) {
  const int M = h->tp.arena_M;
  if (M > 0) {  // sides live in slots [0, n) and [M, M + n): sum over [0, M + n), count running matches in [0, n)
    return ao::launch_sum_counters(h->tp, M + h->selfplay_games, h->selfplay_games, h->d_counters, h->stream);
  }
  return ao::launch_sum_counters(h->tp, h->selfplay_games, h->selfplay_games, h->d_counters, h->stream);
}

// a few games and a tower mode the cluster-of-four kernel exists for (single pass; split precision on 9x9)
bool solo_usable(const ao_engine* h, int n) {
  return h->solo_allowed && n >= 1 && n <= ao::solo_max_games(h->num_sms) && ao::solo_supports(h->B, h->ws[0].precision);
}

// ---- persistent self-play kernel: state transitions
// usable: PVNet evaluator, single-pass fp16 tower (tower_stag.cu; a few games: tower_solo.cu, which also has the split
// mode), plain self-play (no arena), at least a few rounds
// and (nearly) all game slots still playing: the persistent kernel evaluates every slot every round (fixed game -> pass
// map), the two-kernel path packs the live requests densely, which wins once a batch played to the end thins out
bool persist_usable(const ao_engine* h, int rounds) {
  return h->persist_allowed && h->cfg.eval_mode == AO_EVAL_PVNET &&
         (h->ws[0].precision == AO_NN_FP16 || solo_usable(h, h->selfplay_games)) && h->tp.arena_M == 0 &&
         h->selfplay_games > 0 && rounds >= 4 &&
         (h->last_running < 0 || h->last_running * 16 >= (long long)h->selfplay_games * 15);
}
// two-kernel state (answers pending in nn_policy[gm->nn_slot], or nothing pending) -> persistent state: one tree step
// with static request slots consumes what is pending and leaves every running game's next request in nn_in[game]
int enter_persist(ao_engine* h, int max_iters, int n = -1) {
  if (h->persist_pending) return 0;
  if (n < 0) n = h->selfplay_games;
  // answer whatever is still queued (a deferred tail of the last two-kernel round), without deferring again
  AO_CUDA(ao::launch_tower(h->ws[0].tw, h->B, h->ws[0].precision, h->tp.nn_in, queue_of(h, 0, false), n, h->tp.nn_policy,
                           h->tp.nn_value, h->num_sms, h->stream));
  h->launches += 1;
  h->tp.static_slots = 1;
  AO_CUDA(cudaMemsetAsync(h->d_cta_pos, 0, 512 * sizeof(uint32_t), h->stream));
  AO_CUDA(cudaMemsetAsync(h->tp.n_active, 0, sizeof(int32_t), h->stream));
  AO_CUDA(ao::launch_tree_step(h->tp, nullptr, n, max_iters, h->stream));
  h->launches += 1;
  h->persist_pending = true;
  return 0;
}
// persistent state -> two-kernel state: one plain tower launch over all game slots answers the pending requests
// (gm->nn_slot == game slot), after which tree_step_kernel consumes them as usual.  `drop`: the pending requests are
// about to be discarded anyway (games reset / new roots).
int leave_persist(ao_engine* h, bool drop) {
  if (h->persist_pending && !drop) {
    AO_CUDA(ao::launch_tower(h->ws[0].tw, h->B, h->ws[0].precision, h->tp.nn_in, no_queue(), h->selfplay_games, h->tp.nn_policy,
                             h->tp.nn_value, h->num_sms, h->stream));
    h->launches += 1;
  }
  h->persist_pending = false;
  h->tp.static_slots = 0;
  return 0;
}

// all rounds of a call in one launch: a few games -> one cluster of four CTAs per game (tower_solo.cu, latency-bound
// end), otherwise the CTA-pair kernel (tower_stag.cu, PERSIST)
cudaError_t launch_persist_any(ao_engine* h, int n, int rounds) {
  if (solo_usable(h, n))
    return ao::launch_selfplay_solo(h->ws[0].tw, h->B, h->ws[0].precision, h->tp, n, rounds, h->stream);
  return ao::launch_selfplay_persist(h->ws[0].tw, h->B, h->tp, n, rounds, h->num_sms, h->free_run ? h->d_cta_pos : nullptr, h->stream);
}

int poll_active(ao_engine* h, int* active) {
  AO_CUDA(cudaMemcpyAsync(h->h_pinned, h->tp.n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  *active = h->h_pinned[0];
  return 0;
}

int require_weights(ao_engine* h, int set = 0) {
  if (h->cfg.eval_mode == AO_EVAL_PVNET && !h->ws[set].loaded)
    return fail(-3, "no weights loaded%s: call ao_load_weights%s before searching with the PVNet evaluator",
                set ? " for weight set 1 (the arena's enemy)" : "", set ? "_set" : "");
  return 0;
}

}  // namespace

extern "C" const char* ao_last_error(void) { return g_err.c_str(); }

extern "C" int ao_engine_create(const ao_config* cfg, ao_engine** out) {
  if (!cfg || !out) return fail(-1, "null argument");
  if (cfg->board_size < 5 || cfg->board_size > ao::kMaxB) return fail(-1, "board_size %d unsupported (5..15)", cfg->board_size);
  if (cfg->inplanes != 5) return fail(-1, "inplanes must be 5 (history*2+1), got %d", cfg->inplanes);
  if (cfg->planes != 128) return fail(-1, "planes must be 128, got %d", cfg->planes);
  if (cfg->n_blocks < 1 || cfg->n_blocks > 10) return fail(-1, "n_blocks must be in 1..10, got %d", cfg->n_blocks);
  if (cfg->max_games < 1) return fail(-1, "max_games must be >= 1");
  if (cfg->num_mcts < 1) return fail(-1, "num_mcts must be >= 1");
  if (cfg->eval_mode == AO_EVAL_PVNET && cfg->board_size != 9 && cfg->board_size != 15)
    return fail(-1, "the PVNet tower kernel is built for board_size 9 and 15 only");

  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev == 0)
    return fail(-4, "no CUDA device: alpha_omok_b200 has no CPU fallback (%s)", cudaGetErrorString(e0));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(-1, "device %d out of range 0..%d", cfg->device, ndev - 1);
  DeviceGuard guard(cfg->device);
  cudaDeviceProp prop;
  AO_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(-4, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);

  ao_engine* h = new ao_engine();
  memset(&h->tp, 0, sizeof h->tp);
  memset(h->ws, 0, sizeof h->ws);
  h->ws[0].precision = h->ws[1].precision = cfg->nn_precision;
  h->cfg = *cfg;
  h->B = cfg->board_size;
  h->A = h->B * h->B;
  h->G = cfg->max_games;
  h->num_sms = prop.multiProcessorCount;
#ifdef AO_PROBE  // probe build only: run the tower kernels on fewer SMs (what a cluster shape that strands SMs would start from)
  if (getenv("AO_SM_LIMIT") && atoi(getenv("AO_SM_LIMIT")) >= 2 && atoi(getenv("AO_SM_LIMIT")) < h->num_sms)
    h->num_sms = atoi(getenv("AO_SM_LIMIT")) & ~1;
#endif
  h->selfplay_games = 0;
  h->d_stream = nullptr; h->stream_capacity = 0; h->d_stream_next = nullptr; h->stream_episodes = 0;
  h->round_graph = nullptr; h->graph_n = -1; h->graph_disabled = getenv("AO_NO_GRAPH") != nullptr;
  h->timing = false;
  h->launches = 0;
  h->h_pinned = nullptr;
  h->d_fwd_states = nullptr; h->d_fwd_bad = nullptr;
  h->d_wsum = nullptr; h->log_table_n = 0;
  h->persist_allowed = getenv("AO_NO_PERSIST") == nullptr;
  h->solo_allowed = getenv("AO_NO_SOLO") == nullptr;
  h->persist_pending = false;
  h->last_running = -1;
  h->pev[0] = h->pev[1] = h->pev[2] = nullptr;
  if (cfg->stream) {
    h->stream = reinterpret_cast<cudaStream_t>(cfg->stream);
    h->own_stream = false;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail(-2, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    h->own_stream = true;
  }
  const int A = h->A, G = h->G;
  const int node_cap = cfg->node_cap > 0 ? cfg->node_cap : 2048;
  ao::TreeParams& tp = h->tp;
  tp.B = h->B; tp.A = A; tp.G = G;
  tp.num_mcts = cfg->num_mcts; tp.noise = cfg->noise ? 1 : 0; tp.tau_thres = cfg->tau_thres;
  tp.eval_mode = cfg->eval_mode; tp.noise_mode = cfg->noise_mode;
  tp.c_puct = cfg->c_puct > 0 ? cfg->c_puct : 5.0;
  tp.alpha = cfg->alpha > 0 ? cfg->alpha : 10.0 / A;
  tp.seed_lo = (uint32_t)cfg->seed; tp.seed_hi = (uint32_t)(cfg->seed >> 32);
  const uint64_t slot_cap = (uint64_t)node_cap * A;
  if (slot_cap >= (1u << 20)) { delete h; return fail(-1, "node_cap*A = %llu must stay below 2^20", (unsigned long long)slot_cap); }
  tp.slot_cap = (uint32_t)slot_cap;
  tp.gc_cap = (uint32_t)(slot_cap / 4 + 1024);
  tp.nn_log_cap = cfg->nn_log_cap;
  tp.tape_rows = cfg->noise_mode == AO_NOISE_TAPE ? A + 2 : 0;
  const size_t slots = (size_t)G * 2 * slot_cap;
  int rc = 0;
#define EA(p, cnt) if ((rc = ealloc(h, &(p), (cnt))) != 0) { ao_engine_destroy(h); return rc; }
  EA(tp.games, (size_t)G);
  EA(tp.slot_act, slots);
  EA(tp.slot_nw, slots);
  EA(tp.slot_p, slots);
  EA(tp.slot_child, slots);
  EA(tp.path, (size_t)G * (A + 1));
  EA(tp.gc_old, (size_t)G * tp.gc_cap);
  EA(tp.gc_new, (size_t)G * tp.gc_cap);
  EA(tp.rec_visits, (size_t)G * A * A);
  if (tp.tape_rows) { EA(tp.gamma_tape, (size_t)G * tp.tape_rows * A); }
  h->ring_cap = 1u;
  while (h->ring_cap < (uint32_t)G) h->ring_cap <<= 1;
  tp.nn_ring_mask = h->ring_cap - 1u;
  h->defer_tail = getenv("AO_NO_DEFER") == nullptr;
  EA(tp.nn_in, (size_t)2 * h->ring_cap);
  EA(tp.nn_policy, (size_t)2 * h->ring_cap * A);
  EA(tp.nn_value, (size_t)2 * h->ring_cap);
  EA(tp.nn_count, 2);
  EA(tp.nn_head, 2);
  EA(h->d_tower_done, 2);
  EA(h->d_cta_pos, 512);
  h->free_run = getenv("AO_NO_FREERUN") == nullptr;
  h->round_no = 0;
  tp.allow_moves = 1;
  EA(tp.n_active, 1);
  if (tp.nn_log_cap > 0) {
    EA(tp.nnlog_policy, (size_t)G * tp.nn_log_cap * A);
    EA(tp.nnlog_value, (size_t)G * tp.nn_log_cap);
  }
  EA(h->d_ids, (size_t)G);
  EA(h->d_lens, (size_t)G);
  EA(h->d_real_root, (size_t)G);
  EA(h->d_roots, (size_t)G * (A + 1));
  EA(h->d_keys, (size_t)G);
  EA(h->d_visits, (size_t)G * A);
  EA(h->d_priors, (size_t)G * A);
  EA(h->d_counters, 8);
  h->rec_bytes = ((4 + (size_t)A * 2 + 3) & ~(size_t)3) + (size_t)A * A * 4;
  EA(h->d_records, (size_t)G * h->rec_bytes);
#undef EA
  if (cudaMallocHost(reinterpret_cast<void**>(&h->h_pinned), 64) != cudaSuccess) { ao_engine_destroy(h); return fail(-2, "cudaMallocHost failed"); }
  cudaError_t e = cudaMemsetAsync(tp.games, 0, (size_t)G * sizeof(ao::Game), h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(tp.nn_count, 0, 2 * sizeof(uint32_t), h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(tp.nn_head, 0, 2 * sizeof(uint32_t), h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(h->d_tower_done, 0, 2 * sizeof(uint32_t), h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(tp.nn_in, 0, (size_t)2 * h->ring_cap * sizeof(ao::LeafIn), h->stream);
  if (e == cudaSuccess && tp.gamma_tape) e = cudaMemsetAsync(tp.gamma_tape, 0, (size_t)G * tp.tape_rows * A * sizeof(double), h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { ao_engine_destroy(h); return fail(-2, "engine init: %s", cudaGetErrorString(e)); }
  *out = h;
  return 0;
}

extern "C" int ao_engine_destroy(ao_engine* h) {
  if (!h) return 0;
  DeviceGuard guard(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  if (h->d_stream) cudaFree(h->d_stream);
  if (h->round_graph) cudaGraphExecDestroy(h->round_graph);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  for (cudaEvent_t e : h->pev) if (e) cudaEventDestroy(e);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

extern "C" int ao_synchronize(ao_engine* h) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  AO_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------------- weights
extern "C" int ao_load_weights(ao_engine* h, int n_tensors, const char* const* names, const float* const* ptrs,
                               const int64_t* numel) {
  return ao_load_weights_set(h, 0, n_tensors, names, ptrs, numel);
}

extern "C" int ao_load_weights_set(ao_engine* h, int set, int n_tensors, const char* const* names,
                                   const float* const* ptrs, const int64_t* numel) {
  if (!h) return fail(-1, "null engine");
  if (set < 0 || set > 1) return fail(-1, "weight set %d out of range 0..1", set);
  DeviceGuard guard(h->cfg.device);
  ao_engine::WeightSet* W = &h->ws[set];
  std::map<std::string, std::pair<const float*, int64_t>> sd;
  for (int i = 0; i < n_tensors; ++i) sd[names[i]] = {ptrs[i], numel[i]};
  const int A = h->A, C = 128, nb = h->cfg.n_blocks, CI = h->cfg.inplanes;
  auto get = [&](const std::string& k, int64_t want) -> const float* {
    auto it = sd.find(k);
    if (it == sd.end()) { fail(-5, "state_dict key '%s' missing", k.c_str()); return nullptr; }
    if (it->second.second != want) { fail(-5, "state_dict key '%s' has %lld elements, expected %lld", k.c_str(), (long long)it->second.second, (long long)want); return nullptr; }
    return it->second.first;
  };
  struct BN { const float *g, *b, *m, *v; };
  auto get_bn = [&](const std::string& p, int c, BN& bn) {
    bn.g = get(p + ".weight", c); bn.b = get(p + ".bias", c); bn.m = get(p + ".running_mean", c); bn.v = get(p + ".running_var", c);
    return bn.g && bn.b && bn.m && bn.v;
  };
  const int n_layers = 1 + 2 * nb;
  const size_t stem_halves = (size_t)9 * 16 * C, res_halves = (size_t)9 * C * C;
  const size_t total_halves = stem_halves + (size_t)(n_layers - 1) * res_halves;
  std::vector<__half> hi(total_halves), lo(total_halves), pr(total_halves), prl(total_halves), qd(total_halves), qx(2 * total_halves);
  std::vector<float> bias((size_t)n_layers * C);
  size_t off = 0;
  for (int l = 0; l < n_layers; ++l) {
    std::string wname, bnname;
    int ci_n = C;
    if (l == 0) { wname = "conv1.weight"; bnname = "bn1"; ci_n = CI; }
    else {
      const int blk = (l - 1) / 2, which = (l - 1) % 2 + 1;
      wname = "layers." + std::to_string(blk) + ".conv" + std::to_string(which) + ".weight";
      bnname = "layers." + std::to_string(blk) + ".bn" + std::to_string(which);
    }
    const float* w = get(wname, (int64_t)C * ci_n * 9);
    BN bn;
    if (!w || !get_bn(bnname, C, bn)) return -5;
    const int kpad = l == 0 ? 16 : C;
    for (int co = 0; co < C; ++co) {
      const float s = bn.g[co] / std::sqrt(bn.v[co] + 1e-5f);
      bias[(size_t)l * C + co] = bn.b[co] - bn.m[co] * s;
      for (int ti = 0; ti < 9; ++ti)  // packed tap order: centre first (its MMA has no disabled rows), then row-major
        for (int ci = 0; ci < kpad; ++ci) {
          const int t = ti;  // storage slot
          const int tap = ti == 0 ? 4 : (ti <= 4 ? ti - 1 : ti);
          const float v = ci < ci_n ? w[((size_t)co * ci_n + ci) * 9 + tap] * s : 0.f;
          const __half vh = __float2half_rn(v);
          const size_t idx = off + ((size_t)t * (kpad / 8) + ci / 8) * C * 8 + (size_t)co * 8 + (ci % 8);
          hi[idx] = vh;
          lo[idx] = __float2half_rn(v - __half2float(vh));
          // CTA-pair layout: per tap [rank = co / 64][k-chunk][co % 64][8]
          const size_t idx_pair = off + (((size_t)t * 2 + co / 64) * (kpad / 8) + ci / 8) * 64 * 8 + (size_t)(co % 64) * 8 + (ci % 8);
          pr[idx_pair] = vh;
          prl[idx_pair] = lo[idx];
          // cluster-of-four layout (tower_solo.cu): per layer [rank = co / 32][tap][k-chunk][co % 32][8]
          const size_t idx_quad = off + (((size_t)(co / 32) * 9 + t) * (kpad / 8) + ci / 8) * 32 * 8 + (size_t)(co % 32) * 8 + (ci % 8);
          qd[idx_quad] = vh;
          // split mode of the same kernel: a tap's image is [32 rows hi ; 32 rows lo] of this rank's channels
          const size_t idx_x3 = 2 * off + (((size_t)(co / 32) * 9 + t) * (kpad / 8) + ci / 8) * 64 * 8 + (size_t)(co % 32) * 8 + (ci % 8);
          qx[idx_x3] = vh;
          qx[idx_x3 + 32 * 8] = lo[idx];
        }
    }
    off += l == 0 ? stem_halves : res_halves;
  }
  // heads
  std::vector<float> head_w(3 * C), head_b(3), pfc_wT((size_t)2 * A * A), vfc1_wT((size_t)A * C);
  {
    const float* pw = get("policy_head.policy_head.weight", 2 * C);
    const float* vw = get("value_head.value_head.weight", C);
    BN pb, vb;
    if (!pw || !vw || !get_bn("policy_head.policy_bn", 2, pb) || !get_bn("value_head.value_bn", 1, vb)) return -5;
    for (int c = 0; c < 2; ++c) {
      const float s = pb.g[c] / std::sqrt(pb.v[c] + 1e-5f);
      head_b[c] = pb.b[c] - pb.m[c] * s;
      for (int ci = 0; ci < C; ++ci) head_w[c * C + ci] = pw[c * C + ci] * s;
    }
    const float s = vb.g[0] / std::sqrt(vb.v[0] + 1e-5f);
    head_b[2] = vb.b[0] - vb.m[0] * s;
    for (int ci = 0; ci < C; ++ci) head_w[2 * C + ci] = vw[ci] * s;
  }
  const float* pfc_w = get("policy_head.policy_fc.weight", (int64_t)A * 2 * A);
  const float* pfc_b = get("policy_head.policy_fc.bias", A);
  const float* v1w = get("value_head.value_fc1.weight", (int64_t)C * A);
  const float* v1b = get("value_head.value_fc1.bias", C);
  const float* v2w = get("value_head.value_fc2.weight", C);
  const float* v2b = get("value_head.value_fc2.bias", 1);
  if (!pfc_w || !pfc_b || !v1w || !v1b || !v2w || !v2b) return -5;
  for (int o = 0; o < A; ++o)
    for (int f = 0; f < 2 * A; ++f) pfc_wT[(size_t)f * A + o] = pfc_w[(size_t)o * 2 * A + f];
  for (int j = 0; j < C; ++j)
    for (int p = 0; p < A; ++p) vfc1_wT[(size_t)p * C + j] = v1w[(size_t)j * A + p];

  if (!W->loaded) {
    int rc;
#define EA(p, cnt) if ((rc = ealloc(h, &(p), (cnt))) != 0) return rc;
    EA(W->d_conv_hi, total_halves);
    EA(W->d_conv_lo, total_halves);
    EA(W->d_conv_pair, total_halves);
    EA(W->d_conv_pair_lo, total_halves);
    EA(W->d_conv_quad, total_halves);
    EA(W->d_conv_quad_x3, 2 * total_halves);
    EA(W->d_bias, bias.size());
    EA(W->d_head_w, head_w.size());
    EA(W->d_head_b, 4);
    EA(W->d_pfc_wT, pfc_wT.size());
    EA(W->d_pfc_b, (size_t)A);
    EA(W->d_vfc1_wT, vfc1_wT.size());
    EA(W->d_vfc1_b, (size_t)C);
    EA(W->d_vfc2_w, (size_t)C);
#undef EA
  }
  AO_CUDA(cudaStreamSynchronize(h->stream));
  AO_CUDA(cudaMemcpy(W->d_conv_hi, hi.data(), total_halves * 2, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_conv_lo, lo.data(), total_halves * 2, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_conv_pair, pr.data(), total_halves * 2, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_conv_pair_lo, prl.data(), total_halves * 2, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_conv_quad, qd.data(), total_halves * 2, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_conv_quad_x3, qx.data(), total_halves * 4, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_head_w, head_w.data(), head_w.size() * 4, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_head_b, head_b.data(), 3 * 4, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_pfc_wT, pfc_wT.data(), pfc_wT.size() * 4, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_pfc_b, pfc_b, (size_t)A * 4, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_vfc1_wT, vfc1_wT.data(), vfc1_wT.size() * 4, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_vfc1_b, v1b, (size_t)C * 4, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(W->d_vfc2_w, v2w, (size_t)C * 4, cudaMemcpyHostToDevice));
  ao::TowerWeights& tw = W->tw;
#ifdef AO_PROBE
  tw.xflags = getenv("AO_TOWER_XFLAGS") ? atoi(getenv("AO_TOWER_XFLAGS")) : 0;
#endif
  tw.conv_hi = W->d_conv_hi; tw.conv_lo = W->d_conv_lo; tw.conv_pair = W->d_conv_pair; tw.conv_pair_lo = W->d_conv_pair_lo; tw.conv_quad = W->d_conv_quad; tw.conv_quad_x3 = W->d_conv_quad_x3; tw.bias = W->d_bias; tw.head_w = W->d_head_w; tw.head_b = W->d_head_b;
  tw.pfc_wT = W->d_pfc_wT; tw.pfc_b = W->d_pfc_b; tw.vfc1_wT = W->d_vfc1_wT; tw.vfc1_b = W->d_vfc1_b; tw.vfc2_w = W->d_vfc2_w;
  tw.vfc2_b = v2b[0];
  tw.n_layers = n_layers;
  W->loaded = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------------- games
extern "C" int ao_games_reset(ao_engine* h, const int32_t* game_ids, int n, const uint32_t* game_keys) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (n < 0 || n > h->G) return fail(-1, "n out of range");
  for (int i = 0; i < n; ++i)
    if (game_ids[i] < 0 || game_ids[i] >= h->G) return fail(-1, "game id %d out of range", game_ids[i]);
  h->tp.arena_M = 0;
  leave_persist(h, true);
  AO_CUDA(cudaMemcpyAsync(h->d_ids, game_ids, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  if (game_keys) AO_CUDA(cudaMemcpyAsync(h->d_keys, game_keys, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  AO_CUDA(ao::launch_reset_games(h->tp, h->d_ids, n, game_keys ? h->d_keys : nullptr, 0, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int ao_set_gamma_tape(ao_engine* h, int game_id, const double* tape, int n_draws) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (!h->tp.gamma_tape) return fail(-1, "engine was not created with noise_mode = AO_NOISE_TAPE");
  if (game_id < 0 || game_id >= h->G) return fail(-1, "game id out of range");
  if (n_draws > h->tp.tape_rows) n_draws = h->tp.tape_rows;
  AO_CUDA(cudaMemcpy(h->tp.gamma_tape + (size_t)game_id * h->tp.tape_rows * h->A, tape, (size_t)n_draws * h->A * sizeof(double),
                     cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int ao_search(ao_engine* h, const int32_t* game_ids, int n, const int16_t* roots, const int32_t* root_lens,
                         uint32_t* visits, double* priors, int32_t* is_real_root) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (n < 1 || n > h->G) return fail(-1, "n = %d out of range 1..%d", n, h->G);
  int rc = require_weights(h);
  if (rc) return rc;
  const int A = h->A;
  std::vector<uint8_t> seen((size_t)h->G, 0);  // one warp per listed slot: a slot listed twice would race on its tree
  for (int i = 0; i < n; ++i) {
    if (game_ids[i] < 0 || game_ids[i] >= h->G) return fail(-1, "game id %d out of range", game_ids[i]);
    if (seen[game_ids[i]]) return fail(-1, "game id %d listed twice", game_ids[i]);
    seen[game_ids[i]] = 1;
    if (root_lens[i] < 1 || root_lens[i] > A) return fail(-1, "root id length %d out of range 1..%d", root_lens[i], A);
  }
  h->tp.arena_M = 0;
  leave_persist(h, true);
  AO_CUDA(cudaMemcpyAsync(h->d_ids, game_ids, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  AO_CUDA(cudaMemcpyAsync(h->d_lens, root_lens, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  AO_CUDA(cudaMemcpyAsync(h->d_roots, roots, (size_t)n * (A + 1) * 2, cudaMemcpyHostToDevice, h->stream));
  AO_CUDA(ao::launch_set_roots(h->tp, h->d_ids, n, h->d_roots, h->d_lens, h->stream));
  h->launches += 3;  // set_roots + export_roots + sum_counters
  const bool synth = h->cfg.eval_mode == AO_EVAL_SYNTH;
  const int max_iters = synth ? (1 << 30) : 64;
  int active = 1, rounds = 0;
  const int blind = synth ? 0 : h->cfg.num_mcts;  // at least this many rounds are needed before anyone can finish
  bool identity = true;  // the persistent kernel walks the slots [0, n): usable when the caller's ids are exactly those
  for (int i = 0; i < n && identity; ++i) identity = game_ids[i] == i;
  if (identity && h->persist_allowed && !synth && (h->ws[0].precision == AO_NN_FP16 || solo_usable(h, n))) {
    // the whole search in one launch of the persistent kernel (tower + fused tree step): a search needs num_mcts (+1)
    // simulations and every round completes at least one per game; finished games simply stop asking
    if ((rc = enter_persist(h, max_iters, n)) != 0) return rc;
    int todo = h->cfg.num_mcts + 1;
    while (active > 0) {
      AO_CUDA(launch_persist_any(h, n, todo));
      h->launches += 1;
      rounds += todo;
      AO_CUDA(ao::launch_sum_counters(h->tp, n, n, h->d_counters, h->stream));
      unsigned long long cc[8];
      AO_CUDA(cudaMemcpyAsync(cc, h->d_counters, sizeof cc, cudaMemcpyDeviceToHost, h->stream));
      AO_CUDA(cudaStreamSynchronize(h->stream));
      active = (int)cc[1];
      todo = 4;
      if (rounds > 4 * (h->cfg.num_mcts + 2) + 64) return fail(-6, "search did not converge after %d rounds", rounds);
    }
    leave_persist(h, true);
  }
  while (active > 0) {
    if ((rc = run_round(h, h->d_ids, n, max_iters)) != 0) return rc;
    ++rounds;
    // rounds after a search has finished are no-ops for that game (the tree step returns at once, the tower sees no
    // request), so the host only looks every few rounds instead of synchronising after each one
    if (rounds >= blind && ((rounds - blind) % kPollEvery == 0) && (rc = poll_active(h, &active)) != 0) return rc;
    if (rounds > 4 * (h->cfg.num_mcts + 2) + 64) return fail(-6, "search did not converge after %d rounds", rounds);
  }
  AO_CUDA(ao::launch_export_roots(h->tp, h->d_ids, n, h->d_visits, h->d_priors, h->d_real_root, h->stream));
  if (visits) AO_CUDA(cudaMemcpyAsync(visits, h->d_visits, (size_t)n * A * 4, cudaMemcpyDeviceToHost, h->stream));
  if (priors) AO_CUDA(cudaMemcpyAsync(priors, h->d_priors, (size_t)n * A * 8, cudaMemcpyDeviceToHost, h->stream));
  if (is_real_root) AO_CUDA(cudaMemcpyAsync(is_real_root, h->d_real_root, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  AO_CUDA(ao::launch_sum_counters(h->tp, h->G, h->G, h->d_counters, h->stream));
  unsigned long long c[8];
  AO_CUDA(cudaMemcpyAsync(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  if (c[3] != 0) return fail(-7, "%llu game tree(s) overflowed their arena (raise node_cap)", c[3]);
  return 0;
}

extern "C" int ao_nn_forward(ao_engine* h, const float* states, int n, float* p, float* v) {
  return ao_nn_forward_set(h, 0, states, n, p, v);
}

extern "C" int ao_nn_forward_set(ao_engine* h, int set, const float* states, int n, float* p, float* v) {
  if (!h) return fail(-1, "null engine");
  if (set < 0 || set > 1) return fail(-1, "weight set %d out of range 0..1", set);
  DeviceGuard guard(h->cfg.device);
  if (!h->ws[set].loaded) return fail(-3, "no weights loaded");
  if (h->B != 9 && h->B != 15) return fail(-1, "tower kernel supports board_size 9 and 15");
  const int A = h->A, C = h->cfg.inplanes;
  const int chunk = h->G;
  if (!h->d_fwd_states) {  // staging of the dense input planes: allocated once, lives as long as the engine
    int rc0;
    if ((rc0 = ealloc(h, &h->d_fwd_states, (size_t)chunk * C * A)) != 0) return rc0;
    if ((rc0 = ealloc(h, &h->d_fwd_bad, 1)) != 0) return rc0;
  }
  float* d_states = h->d_fwd_states;
  int* d_bad = h->d_fwd_bad;
  AO_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream));
  int rc = 0;
  for (int o = 0; o < n && rc == 0; o += chunk) {
    const int m = n - o < chunk ? n - o : chunk;
    cudaError_t e = cudaMemcpyAsync(d_states, states + (size_t)o * C * A, (size_t)m * C * A * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = ao::launch_pack_states(d_states, m, h->B, C, h->tp.nn_in, d_bad, h->stream);
    if (e == cudaSuccess) e = ao::launch_tower(h->ws[set].tw, h->B, h->ws[set].precision, h->tp.nn_in, no_queue(), m, h->tp.nn_policy, h->tp.nn_value, h->num_sms, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p + (size_t)o * A, h->tp.nn_policy, (size_t)m * A * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(v + o, h->tp.nn_value, (size_t)m * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail(-100 - (int)e, "ao_nn_forward: %s", cudaGetErrorString(e));
  }
  int bad = 0;
  if (rc == 0 && cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(-2, "memcpy failed");
  if (rc == 0 && bad) rc = fail(-8, "states must be {0,1}-valued planes with a constant colour plane (utils.get_state_pt)");
  return rc;
}

// ---------------------------------------------------------------------------------------------------- self-play
extern "C" int ao_selfplay_begin_mode(ao_engine* h, int n_games, uint32_t first_key, int recycle);
extern "C" int ao_selfplay_begin(ao_engine* h, int n_games, uint32_t first_key) {
  return ao_selfplay_begin_mode(h, n_games, first_key, 0);
}
extern "C" int ao_selfplay_begin_mode(ao_engine* h, int n_games, uint32_t first_key, int recycle) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (n_games < 1 || n_games > h->G) return fail(-1, "n_games = %d out of range 1..%d", n_games, h->G);
  int rc = require_weights(h);
  if (rc) return rc;
  h->tp.arena_M = 0;
  leave_persist(h, true);
  AO_CUDA(drop_queued_requests(h));
  std::vector<uint32_t> keys(n_games);
  for (int i = 0; i < n_games; ++i) keys[i] = first_key + (uint32_t)i;
  AO_CUDA(cudaMemcpyAsync(h->d_keys, keys.data(), (size_t)n_games * 4, cudaMemcpyHostToDevice, h->stream));
  AO_CUDA(ao::launch_reset_games(h->tp, nullptr, n_games, h->d_keys, recycle ? 2 : 1, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  h->selfplay_games = n_games;
  h->last_running = -1;
  return 0;
}

// Continuous self-play (the caller side of main.self_play for more episodes than game slots): `n_slots`
// games run concurrently; a slot whose episode ends packs its record on the device and starts the episode with the
// next unplayed decision-stream key until `n_episodes` keys [first_key, first_key + n_episodes) are handed out.  Every
// episode is a function of its key alone, so the records equal those of ao_selfplay_begin runs with the same keys, and
// short games no longer leave their slot idle while the longest game of the batch finishes.
extern "C" int ao_selfplay_stream_begin(ao_engine* h, int n_slots, uint32_t first_key, int n_episodes) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (n_episodes < 1) return fail(-1, "n_episodes = %d must be positive", n_episodes);
  if (n_slots > n_episodes) n_slots = n_episodes;
  if (n_slots < 1 || n_slots > h->G) return fail(-1, "n_slots = %d out of range 1..%d", n_slots, h->G);
  if (h->cfg.noise && h->cfg.noise_mode == AO_NOISE_TAPE)
    return fail(-1, "continuous self-play needs the device noise generator (gamma tapes are per slot, not per key)");
  int rc = require_weights(h);
  if (rc) return rc;
  if ((size_t)n_episodes > h->stream_capacity) {
    if (h->d_stream) cudaFree(h->d_stream);
    h->d_stream = nullptr;
    h->stream_capacity = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&h->d_stream), (size_t)n_episodes * h->rec_bytes);
    if (e != cudaSuccess) return fail(-2, "cudaMalloc of the stream record slab (%zu bytes) failed: %s", (size_t)n_episodes * h->rec_bytes, cudaGetErrorString(e));
    h->stream_capacity = (size_t)n_episodes;
  }
  if (!h->d_stream_next && (rc = ealloc(h, &h->d_stream_next, 1)) != 0) return rc;
  h->tp.arena_M = 0;
  leave_persist(h, true);
  AO_CUDA(drop_queued_requests(h));
  AO_CUDA(cudaMemsetAsync(h->d_stream, 0, (size_t)n_episodes * h->rec_bytes, h->stream));
  const uint32_t next = first_key + (uint32_t)n_slots;
  AO_CUDA(cudaMemcpyAsync(h->d_stream_next, &next, 4, cudaMemcpyHostToDevice, h->stream));
  h->tp.stream_out = h->d_stream;
  h->tp.stream_rec_bytes = h->rec_bytes;
  h->tp.stream_next_key = h->d_stream_next;
  h->tp.stream_first_key = first_key;
  h->tp.stream_key_end = first_key + (uint32_t)n_episodes;
  std::vector<uint32_t> keys(n_slots);
  for (int i = 0; i < n_slots; ++i) keys[i] = first_key + (uint32_t)i;
  AO_CUDA(cudaMemcpyAsync(h->d_keys, keys.data(), (size_t)n_slots * 4, cudaMemcpyHostToDevice, h->stream));
  AO_CUDA(ao::launch_reset_games(h->tp, nullptr, n_slots, h->d_keys, 3, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  h->selfplay_games = n_slots;
  h->stream_episodes = n_episodes;
  h->last_running = -1;
  return 0;
}

// Record slab of the current / last stream run: n_episodes records of ao_records_dev's layout, index = key - first_key
// (complete once ao_selfplay_rounds reports 0 running games).  The pointer stays valid until the next stream_begin.
extern "C" int ao_selfplay_stream_records_dev(ao_engine* h, void** dev_ptr, size_t* bytes_per_game, int* n_episodes) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (h->stream_episodes <= 0) return fail(-1, "call ao_selfplay_stream_begin first");
  AO_CUDA(cudaStreamSynchronize(h->stream));
  if (dev_ptr) *dev_ptr = h->d_stream;
  if (bytes_per_game) *bytes_per_game = h->rec_bytes;
  if (n_episodes) *n_episodes = h->stream_episodes;
  return 0;
}

// Arena (eval_main.main, eval_main.py:204-333) on the device: n_slots concurrent series of `matches_per_slot` matches,
// the player (weight set 0) against the enemy (weight set 1, or a RandomAgent), each side with its own tree and
// decision stream, colours swapped after every match.  Driven by ao_selfplay_rounds like self-play.
extern "C" int ao_arena_begin(ao_engine* h, int n_slots, uint32_t first_key, int matches_per_slot, int player_kind,
                              int enemy_kind, int keep_records, int n_mcts_player, int n_mcts_enemy) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (n_slots < 1 || 2 * n_slots > h->G) return fail(-1, "n_slots = %d needs max_games >= %d (one game slot per side), have %d", n_slots, 2 * n_slots, h->G);
  if (matches_per_slot < 1) return fail(-1, "matches_per_slot must be positive");
  if (player_kind < AO_SIDE_ZERO || player_kind > AO_SIDE_UCT || enemy_kind < AO_SIDE_ZERO || enemy_kind > AO_SIDE_UCT)
    return fail(-1, "agent kinds must be AO_SIDE_ZERO / RANDOM / PUCT / UCT");
  int rc;
  if (player_kind == AO_SIDE_ZERO && (rc = require_weights(h, 0)) != 0) return rc;
  if (enemy_kind == AO_SIDE_ZERO && (rc = require_weights(h, 1)) != 0) return rc;
  const int mcts_p = n_mcts_player > 0 ? n_mcts_player : h->cfg.num_mcts, mcts_e = n_mcts_enemy > 0 ? n_mcts_enemy : h->cfg.num_mcts;
  if ((player_kind == AO_SIDE_UCT && h->log_table_n < mcts_p + 2) || (enemy_kind == AO_SIDE_UCT && h->log_table_n < mcts_e + 2))
    return fail(-1, "UCT sides need ao_set_log_table with at least num_mcts + 2 entries");
  const size_t n_rec = keep_records ? (size_t)n_slots * (size_t)matches_per_slot : 0;
  if (n_rec > h->stream_capacity) {
    if (h->d_stream) cudaFree(h->d_stream);
    h->d_stream = nullptr;
    h->stream_capacity = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&h->d_stream), n_rec * h->rec_bytes);
    if (e != cudaSuccess) return fail(-2, "cudaMalloc of the match record slab (%zu bytes) failed: %s", n_rec * h->rec_bytes, cudaGetErrorString(e));
    h->stream_capacity = n_rec;
  }
  if (n_rec) AO_CUDA(cudaMemsetAsync(h->d_stream, 0, n_rec * h->rec_bytes, h->stream));
  leave_persist(h, true);
  AO_CUDA(drop_queued_requests(h));
  ao::TreeParams& tp = h->tp;
  tp.stream_out = n_rec ? h->d_stream : nullptr;
  tp.stream_rec_bytes = h->rec_bytes;
  tp.arena_M = n_slots;
  tp.arena_matches_per_slot = matches_per_slot;
  tp.arena_num_mcts[0] = mcts_p;
  tp.arena_num_mcts[1] = mcts_e;
  tp.arena_kind[0] = player_kind;
  tp.arena_kind[1] = enemy_kind;
  tp.rollout_sims_per_round = 8;
  tp.synth_salt[0] = 0u;
  tp.synth_salt[1] = 1u;
  AO_CUDA(ao::launch_reset_arena(tp, n_slots, first_key, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  h->selfplay_games = n_slots;
  h->stream_episodes = (int)n_rec;
  return 0;
}

// log(k) for k in [0, n) exactly as the caller's numpy computes it: UCTAgent's exploration term sqrt(2 log(sum n) / n)
// (agents.py:556) is compared for equality between children, so the device must not use a different libm.
extern "C" int ao_set_log_table(ao_engine* h, const double* table, int n) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (!table || n < 2) return fail(-1, "log table needs at least 2 entries");
  double* d = nullptr;
  int rc = ealloc(h, &d, (size_t)n);
  if (rc) return rc;
  AO_CUDA(cudaStreamSynchronize(h->stream));
  AO_CUDA(cudaMemcpy(d, table, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  h->tp.log_table = d;
  h->tp.log_table_n = n;
  h->log_table_n = n;
  return 0;
}

// PUCTAgent / UCTAgent.get_pi (agents.py:283-296 / 461-476) minus the final arg-max: a fresh search of num_mcts + 1
// simulations with random play-outs for each listed slot from the given root ID.
extern "C" int ao_rollout_search(ao_engine* h, int kind, const int32_t* game_ids, int n, const int16_t* roots,
                                 const int32_t* root_lens, int num_mcts, uint32_t* visits, float* w) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (kind != AO_SIDE_PUCT && kind != AO_SIDE_UCT) return fail(-1, "kind must be AO_SIDE_PUCT or AO_SIDE_UCT");
  if (n < 1 || n > h->G) return fail(-1, "n = %d out of range 1..%d", n, h->G);
  if (num_mcts < 1) num_mcts = h->cfg.num_mcts;
  if (kind == AO_SIDE_UCT && h->log_table_n < num_mcts + 2) return fail(-1, "UCT needs ao_set_log_table with at least num_mcts + 2 entries");
  const int A = h->A;
  std::vector<uint8_t> seen((size_t)h->G, 0);
  for (int i = 0; i < n; ++i) {
    if (game_ids[i] < 0 || game_ids[i] >= h->G) return fail(-1, "game id %d out of range", game_ids[i]);
    if (seen[game_ids[i]]) return fail(-1, "game id %d listed twice", game_ids[i]);
    seen[game_ids[i]] = 1;
    if (root_lens[i] < 1 || root_lens[i] > A) return fail(-1, "root id length %d out of range 1..%d", root_lens[i], A);
  }
  h->tp.arena_M = 0;
  leave_persist(h, true);
  if (!h->d_wsum) {
    int rc = ealloc(h, &h->d_wsum, (size_t)h->G * A);
    if (rc) return rc;
  }
  AO_CUDA(cudaMemcpyAsync(h->d_ids, game_ids, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  AO_CUDA(cudaMemcpyAsync(h->d_lens, root_lens, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  AO_CUDA(cudaMemcpyAsync(h->d_roots, roots, (size_t)n * (A + 1) * 2, cudaMemcpyHostToDevice, h->stream));
  AO_CUDA(ao::launch_rollout_search(h->tp, kind, num_mcts, h->d_ids, n, h->d_roots, h->d_lens, h->d_visits, h->d_wsum, h->stream));
  h->launches += 1;
  if (visits) AO_CUDA(cudaMemcpyAsync(visits, h->d_visits, (size_t)n * A * 4, cudaMemcpyDeviceToHost, h->stream));
  if (w) AO_CUDA(cudaMemcpyAsync(w, h->d_wsum, (size_t)n * A * 4, cudaMemcpyDeviceToHost, h->stream));
  AO_CUDA(ao::launch_sum_counters(h->tp, h->G, h->G, h->d_counters, h->stream));
  unsigned long long c[8];
  AO_CUDA(cudaMemcpyAsync(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  if (c[3] != 0) return fail(-7, "%llu game tree(s) overflowed their arena (raise node_cap)", c[3]);
  return 0;
}

extern "C" int ao_selfplay_rounds(ao_engine* h, int rounds, uint64_t* out5) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (h->selfplay_games <= 0) return fail(-1, "call ao_selfplay_begin first");
  const bool synth = h->cfg.eval_mode == AO_EVAL_SYNTH;
  const int max_iters = synth ? (1 << 30) : 64;
  int rc;
  if (persist_usable(h, rounds)) {
    // ONE launch for all `rounds` rounds: tower and tree step fused in the persistent kernel (tower_stag.cu)
    if ((rc = enter_persist(h, max_iters)) != 0) return rc;
    AO_CUDA(launch_persist_any(h, h->selfplay_games, rounds));
    h->launches += 1;
  } else {
    if ((rc = leave_persist(h, false)) != 0) return rc;
    if ((rc = run_rounds(h, h->selfplay_games, max_iters, rounds)) != 0) return rc;
  }
  AO_CUDA(launch_sum_selfplay(h));
  h->launches += 1;
  unsigned long long c[8];
  AO_CUDA(cudaMemcpyAsync(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  h->last_running = (long long)c[1];
  if (out5) for (int i = 0; i < 8; ++i) out5[i] = c[i];
  return 0;
}

// Same as ao_selfplay_rounds but with CUDA events around every kernel launch: returns the summed device time of the
// tree-step kernels and of the tower kernels (ms) - the per-kernel numbers behind bench.py's roofline object.
extern "C" int ao_selfplay_rounds_timed(ao_engine* h, int rounds, uint64_t* out5, float* tree_ms, float* tower_ms) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (h->selfplay_games <= 0) return fail(-1, "call ao_selfplay_begin first");
  if (rounds < 1 || rounds > 4096) return fail(-1, "rounds out of range 1..4096");
  if (persist_usable(h, rounds)) {
    // persistent kernel: tree_ms = the one stand-alone tree step that enters the persistent state (0 when already in
    // it), tower_ms = the persistent kernel itself (tower + fused tree steps of all `rounds` rounds)
    for (int i = 0; i < 3; ++i)
      if (!h->pev[i]) AO_CUDA(cudaEventCreate(&h->pev[i]));
    const int max_iters_p = 64;
    int rc;
    AO_CUDA(cudaEventRecord(h->pev[0], h->stream));
    if ((rc = enter_persist(h, max_iters_p)) != 0) return rc;
    AO_CUDA(cudaEventRecord(h->pev[1], h->stream));
    AO_CUDA(launch_persist_any(h, h->selfplay_games, rounds));
    h->launches += 1;
    AO_CUDA(cudaEventRecord(h->pev[2], h->stream));
    AO_CUDA(launch_sum_selfplay(h));
    h->launches += 1;
    unsigned long long c[8];
    AO_CUDA(cudaMemcpyAsync(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    AO_CUDA(cudaStreamSynchronize(h->stream));
    float a = 0.f, b = 0.f;
    AO_CUDA(cudaEventElapsedTime(&a, h->pev[0], h->pev[1]));
    AO_CUDA(cudaEventElapsedTime(&b, h->pev[1], h->pev[2]));
    if (tree_ms) *tree_ms = a;
    if (tower_ms) *tower_ms = b;
    h->last_running = (long long)c[1];
    if (out5) for (int i = 0; i < 8; ++i) out5[i] = c[i];
    return 0;
  }
  {
    int rc0 = leave_persist(h, false);
    if (rc0) return rc0;
  }
  while ((int)h->ev.size() < 3 * rounds) {
    cudaEvent_t e;
    AO_CUDA(cudaEventCreate(&e));
    h->ev.push_back(e);
  }
  const int max_iters = h->cfg.eval_mode == AO_EVAL_SYNTH ? (1 << 30) : 64;
  int rc;
  for (int r = 0; r < rounds; ++r)
    if ((rc = run_round(h, nullptr, h->selfplay_games, max_iters, r)) != 0) return rc;
  AO_CUDA(launch_sum_selfplay(h));
  h->launches += 1;
  unsigned long long c[8];
  AO_CUDA(cudaMemcpyAsync(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  float t_tree = 0.f, t_tower = 0.f;
  for (int r = 0; r < rounds; ++r) {
    float a = 0.f, b = 0.f;
    AO_CUDA(cudaEventElapsedTime(&a, h->ev[3 * r], h->ev[3 * r + 1]));
    AO_CUDA(cudaEventElapsedTime(&b, h->ev[3 * r + 1], h->ev[3 * r + 2]));
    t_tree += a;
    t_tower += b;
  }
  if (tree_ms) *tree_ms = t_tree;
  if (tower_ms) *tower_ms = t_tower;
  h->last_running = (long long)c[1];
  if (out5) for (int i = 0; i < 8; ++i) out5[i] = c[i];
  return 0;
}

#ifdef AO_PROBE
// Cycle counters of CTA 0 of the tower kernel, accumulated over all launches since enabling:
// [0] MMA-issuer total, [1] MMA waits for the epilogue (operand ready), [2] MMA waits for weights (TMA ring),
// [3] launches, [4] epilogue thread 0 total, [5] epilogue waits for the accumulators, [6] heads.
extern "C" int ao_tower_debug(ao_engine* h, int enable, uint64_t* out8) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  AO_CUDA(cudaStreamSynchronize(h->stream));
  if (out8 && h->ws[0].tw.dbg) AO_CUDA(cudaMemcpy(out8, h->ws[0].tw.dbg, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (enable && !h->ws[0].tw.dbg) {
    unsigned long long* p = nullptr;
    int rc = ealloc(h, &p, 8);
    if (rc) return rc;
    h->ws[0].tw.dbg = p;
  }
  if (h->ws[0].tw.dbg) AO_CUDA(cudaMemset(h->ws[0].tw.dbg, 0, 8 * sizeof(uint64_t)));
  if (!enable) h->ws[0].tw.dbg = nullptr;
  return 0;
}
#endif  // AO_PROBE

// Switch the tower's operand mode at run time (AO_NN_*); the facades use it to pick the cheapest mode that meets the
// 1e-4 contract for the loaded weights.
extern "C" int ao_set_nn_precision(ao_engine* h, int mode) { return ao_set_nn_precision_set(h, 0, mode); }

extern "C" int ao_set_nn_precision_set(ao_engine* h, int set, int mode) {
  if (!h) return fail(-1, "null engine");
  if (set < 0 || set > 1) return fail(-1, "weight set %d out of range 0..1", set);
  DeviceGuard guard(h->cfg.device);
  if (mode != AO_NN_FP16 && mode != AO_NN_FP16X3 && mode != AO_NN_FP16_1CTA && mode != AO_NN_FP16_LOCKSTEP) return fail(-1, "unknown nn_precision %d", mode);
  if (set == 0 && mode != h->ws[0].precision) leave_persist(h, false);
  AO_CUDA(cudaStreamSynchronize(h->stream));
  h->ws[set].precision = mode;
  if (set == 0) h->cfg.nn_precision = mode;
  return 0;
}

extern "C" int ao_launch_count(ao_engine* h, uint64_t* out) {
  if (!h || !out) return fail(-1, "null argument");
  *out = h->launches;
  return 0;
}

extern "C" int ao_selfplay_fetch(ao_engine* h, int n_games, int16_t* moves, int32_t* n_moves, int8_t* winners,
                                 uint32_t* visits) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (n_games < 1 || n_games > h->G) return fail(-1, "n_games out of range");
  const int A = h->A;
  AO_CUDA(ao::launch_pack_records(h->tp, n_games, h->d_records, h->rec_bytes, h->stream));
  std::vector<uint8_t> buf((size_t)n_games * h->rec_bytes);
  AO_CUDA(cudaMemcpyAsync(buf.data(), h->d_records, buf.size(), cudaMemcpyDeviceToHost, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  const size_t voff = (4 + (size_t)A * 2 + 3) & ~(size_t)3;
  for (int g = 0; g < n_games; ++g) {
    const uint8_t* rec = buf.data() + (size_t)g * h->rec_bytes;
    const int16_t* hdr = reinterpret_cast<const int16_t*>(rec);
    if (n_moves) n_moves[g] = hdr[0];
    if (winners) winners[g] = (int8_t)rec[2];
    if (moves) memcpy(moves + (size_t)g * A, hdr + 2, (size_t)A * 2);
    if (visits) memcpy(visits + (size_t)g * A * A, rec + voff, (size_t)A * A * 4);
  }
  return 0;
}

extern "C" int ao_get_nn_log(ao_engine* h, int game_id, float* policy, float* value, int32_t capacity, int32_t* count) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (h->tp.nn_log_cap <= 0) return fail(-1, "engine was created with nn_log_cap = 0");
  if (game_id < 0 || game_id >= h->G) return fail(-1, "game id out of range");
  ao::Game g;
  AO_CUDA(cudaStreamSynchronize(h->stream));
  AO_CUDA(cudaMemcpy(&g, h->tp.games + game_id, sizeof g, cudaMemcpyDeviceToHost));
  int c = (int)g.nn_log_count;
  if (count) *count = c;
  if (c > h->tp.nn_log_cap) return fail(-9, "NN log of game %d overflowed (%d > %d)", game_id, c, h->tp.nn_log_cap);
  if (c > capacity) c = capacity;
  if (policy) AO_CUDA(cudaMemcpy(policy, h->tp.nnlog_policy + (size_t)game_id * h->tp.nn_log_cap * h->A, (size_t)c * h->A * 4, cudaMemcpyDeviceToHost));
  if (value) AO_CUDA(cudaMemcpy(value, h->tp.nnlog_value + (size_t)game_id * h->tp.nn_log_cap, (size_t)c * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ao_records_dev(ao_engine* h, void** dev_ptr, size_t* bytes_per_game) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (dev_ptr) *dev_ptr = h->d_records;
  if (bytes_per_game) *bytes_per_game = h->rec_bytes;
  return 0;
}

extern "C" int ao_records_pack(ao_engine* h, int n_games) {
  if (!h) return fail(-1, "null engine");
  DeviceGuard guard(h->cfg.device);
  if (n_games < 1 || n_games > h->G) return fail(-1, "n_games out of range");
  AO_CUDA(ao::launch_pack_records(h->tp, n_games, h->d_records, h->rec_bytes, h->stream));
  AO_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------------- stateless
namespace {
int need_gpu() {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(-4, "no CUDA device: alpha_omok_b200 has no CPU fallback (%s)", cudaGetErrorString(e));
  return 0;
}
}  // namespace

extern "C" int ao_check_win(const int8_t* boards, int n, int board_size, uint8_t* out) {
  int rc = need_gpu();
  if (rc) return rc;
  if (board_size < 5 || board_size > ao::kMaxB) return fail(-1, "board_size unsupported");
  if (n <= 0) return 0;
  const size_t A = (size_t)board_size * board_size;
  int8_t* d_b = nullptr;
  uint8_t* d_o = nullptr;
  AO_CUDA(dalloc(&d_b, (size_t)n * A));
  AO_CUDA(dalloc(&d_o, (size_t)n));
  AO_CUDA(cudaMemcpy(d_b, boards, (size_t)n * A, cudaMemcpyHostToDevice));
  AO_CUDA(ao::launch_check_win(d_b, n, board_size, d_o, 0));
  AO_CUDA(cudaMemcpy(out, d_o, (size_t)n, cudaMemcpyDeviceToHost));
  cudaFree(d_b);
  cudaFree(d_o);
  return 0;
}

extern "C" int ao_encode_state(const int16_t* ids, const int32_t* lens, int n, int board_size, float* out) {
  int rc = need_gpu();
  if (rc) return rc;
  if (board_size < 5 || board_size > ao::kMaxB) return fail(-1, "board_size unsupported");
  if (n <= 0) return 0;
  const size_t A = (size_t)board_size * board_size;
  int16_t* d_i = nullptr;
  int32_t* d_l = nullptr;
  float* d_o = nullptr;
  AO_CUDA(dalloc(&d_i, (size_t)n * (A + 1)));
  AO_CUDA(dalloc(&d_l, (size_t)n));
  AO_CUDA(dalloc(&d_o, (size_t)n * 5 * A));
  AO_CUDA(cudaMemcpy(d_i, ids, (size_t)n * (A + 1) * 2, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(d_l, lens, (size_t)n * 4, cudaMemcpyHostToDevice));
  AO_CUDA(ao::launch_encode_state(d_i, d_l, n, board_size, d_o, 0));
  AO_CUDA(cudaMemcpy(out, d_o, (size_t)n * 5 * A * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_i); cudaFree(d_l); cudaFree(d_o);
  return 0;
}

extern "C" int ao_legal_actions(const int16_t* ids, const int32_t* lens, int n, int board_size, int16_t* out) {
  int rc = need_gpu();
  if (rc) return rc;
  if (board_size < 5 || board_size > ao::kMaxB) return fail(-1, "board_size unsupported");
  if (n <= 0) return 0;
  const size_t A = (size_t)board_size * board_size;
  int16_t *d_i = nullptr, *d_o = nullptr;
  int32_t* d_l = nullptr;
  AO_CUDA(dalloc(&d_i, (size_t)n * (A + 1)));
  AO_CUDA(dalloc(&d_l, (size_t)n));
  AO_CUDA(dalloc(&d_o, (size_t)n * A));
  AO_CUDA(cudaMemcpy(d_i, ids, (size_t)n * (A + 1) * 2, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMemcpy(d_l, lens, (size_t)n * 4, cudaMemcpyHostToDevice));
  AO_CUDA(ao::launch_legal_actions(d_i, d_l, n, board_size, d_o, 0));
  AO_CUDA(cudaMemcpy(out, d_o, (size_t)n * A * 2, cudaMemcpyDeviceToHost));
  cudaFree(d_i); cudaFree(d_l); cudaFree(d_o);
  return 0;
}
