// Kernels of the tree search and their launchers; the device functions live in tree_device.cuh.
#include "tree_device.cuh"

namespace ao {

namespace {

// One warp advances one game until it needs a network evaluation (or finishes its search / game).
template <int MAXJ>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
tree_step_kernel(TreeParams P, const int32_t* __restrict__ game_ids, int n, int max_iters) {
  __shared__ WarpSmemStore s_warp[kWarpsPerBlock];
  const int wib = threadIdx.x >> 5;
  const int w = blockIdx.x * kWarpsPerBlock + wib;
  if (w >= n) return;
  WarpSmem ws = s_warp[wib].ref();
  const bool running = tree_step_game<MAXJ>(P, game_ids ? game_ids[w] : w, &ws, (int)(threadIdx.x & 31), max_iters, false, 0.f);
  if ((threadIdx.x & 31) == 0 && running) atomicAdd(P.n_active, 1);
}

// ZeroAgent.reset() + new GameState
__global__ void reset_games_kernel(TreeParams P, const int32_t* __restrict__ ids, int n, const uint32_t* __restrict__ keys,
                                   int auto_play) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int game = ids ? ids[i] : i;
  Game* gm = &P.games[game];
  for (int r = 0; r < kRowsPad; ++r) {
    gm->rows_b[r] = 0; gm->rows_w[r] = 0; gm->leaf_rows_b[r] = 0; gm->leaf_rows_w[r] = 0;
  }
  gm->n_moves = 0; gm->last1 = -1; gm->last2 = -1;
  gm->status = auto_play ? ST_SEARCH : ST_FREE;
  gm->arena = 0; gm->root_node = CH_UNVISITED; gm->root_n = 0u; gm->root_w = 0.f; gm->slot_count = 0u;
  gm->sims_done = 0; gm->sims_target = P.num_mcts + 1; gm->is_real_root = 1; gm->auto_play = auto_play;
  gm->rng_ctr = 0u; gm->noise_draws = 0u; gm->game_key = keys ? keys[i] : (uint32_t)game;
  gm->leaf_depth = 0; gm->leaf_n_moves = 0; gm->nn_slot = 0; gm->nn_ticket = 0u; gm->nn_static = 0; gm->winner = 0; gm->error = 0; gm->nn_log_count = 0u;
  gm->sims_total = 0ull; gm->nn_evals = 0ull; gm->terminal_sims = 0ull; gm->moves_played = 0ull;
  gm->games_finished = 0ull;
}

// Facade root handling == `_init_mcts` (agents.py:82-103): is the requested root ID inside the stored tree?
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
set_roots_kernel(TreeParams P, const int32_t* __restrict__ ids, int n, const int16_t* __restrict__ roots,
                 const int32_t* __restrict__ lens) {
  __shared__ WarpSmemStore s_warp[kWarpsPerBlock];
  const int wib = threadIdx.x >> 5;
  const int w = blockIdx.x * kWarpsPerBlock + wib;
  if (w >= n) return;
  const int game = ids[w];
  Game* gm = &P.games[game];
  WarpSmem ws = s_warp[wib].ref();
  Ctx c{P, gm, &ws, (int)(threadIdx.x & 31), game, 0, make_uint2(P.seed_lo, P.seed_hi)};
  Regs g;
  load_regs(c, g);
  c.abase = arena_base(P, game, g.arena);
  const int16_t* rid = roots + (size_t)w * (P.A + 1);
  const int m = lens[w] - 1;  // plies in the requested root
  bool in_tree = gm->status != ST_FREE && gm->status != ST_ERROR && m >= g.n_moves;
  if (in_tree)
    for (int i = 0; i < g.n_moves; ++i)
      if ((int)gm->moves[i] != (int)rid[1 + i]) in_tree = false;
  if (in_tree) {
    while (g.n_moves < m) {
      if (g.root_node < 0) {  // nothing stored below an unexpanded / terminal root
        in_tree = false;
        break;
      }
      if (!advance_root(c, g, (int)rid[1 + g.n_moves])) {
        in_tree = false;
        break;
      }
    }
  }
  if (!in_tree) {  // real root: fresh tree at this position
    g.rb = 0u; g.rw = 0u;
    for (int i = 0; i < m; ++i) {
      const int a = rid[1 + i];
      if (c.lane == a / P.B) {
        if ((i & 1) == 0) g.rb |= 1u << (a % P.B);
        else g.rw |= 1u << (a % P.B);
      }
      if (c.lane == 0) gm->moves[i] = (uint8_t)a;
    }
    g.n_moves = m;
    g.last1 = m >= 1 ? rid[m] : -1;
    g.last2 = m >= 2 ? rid[m - 1] : -1;
    g.root_node = CH_UNVISITED; g.root_n = 0u; g.root_w = 0.f; g.slot_count = 0u;
    g.sims_target = P.num_mcts + 1;
  } else {
    g.sims_target = P.num_mcts;
    if (P.noise && g.root_node >= 0) remix_root_noise(c, g.root_node, P.A - g.n_moves, g.noise_draws);
  }
  g.sims_done = 0;
  if (g.status != ST_ERROR || !in_tree) g.status = ST_SEARCH;  // a compaction-queue overflow stays reported
  if (c.lane == 0) {
    gm->is_real_root = in_tree ? 0 : 1;
    gm->auto_play = 0;
  }
  __syncwarp();
  store_regs<1>(c, g);
}

// PUCTAgent / UCTAgent.get_pi minus the final arg-max (agents.py:283-296 / 461-476): a fresh search of num_mcts + 1
// simulations from the given root ID in the listed slots; visits[a] = n(child a), w[a] = w(child a) (q = w / n).
template <int MAXJ>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rollout_search_kernel(TreeParams P, int kind, int num_mcts, const int32_t* __restrict__ ids, int n,
                      const int16_t* __restrict__ roots, const int32_t* __restrict__ lens, uint32_t* __restrict__ visits,
                      float* __restrict__ wsum) {
  __shared__ WarpSmemStore s_warp[kWarpsPerBlock];
  const int wib = threadIdx.x >> 5;
  const int w = blockIdx.x * kWarpsPerBlock + wib;
  if (w >= n) return;
  const int game = ids[w];
  Game* gm = &P.games[game];
  WarpSmem ws = s_warp[wib].ref();
  Ctx c{P, gm, &ws, (int)(threadIdx.x & 31), game, 0, make_uint2(P.seed_lo, P.seed_hi)};
  Regs g;
  load_regs(c, g);
  c.abase = arena_base(P, game, g.arena);
  const int16_t* rid = roots + (size_t)w * (P.A + 1);
  const int m = lens[w] - 1;
  g.rb = 0u; g.rw = 0u;
  for (int i = 0; i < m; ++i) {
    const int a = rid[1 + i];
    if (c.lane == a / P.B) {
      if ((i & 1) == 0) g.rb |= 1u << (a % P.B);
      else g.rw |= 1u << (a % P.B);
    }
    if (c.lane == 0) gm->moves[i] = (uint8_t)a;
  }
  g.n_moves = m;
  g.last1 = m >= 1 ? rid[m] : -1;
  g.last2 = m >= 2 ? rid[m - 1] : -1;
  rollout_fresh_tree(g);
  g.sims_done = 0;
  g.sims_target = num_mcts + 1;
  g.status = ST_SEARCH_DONE;
  bool ok = true;
  while (ok && g.sims_done < g.sims_target) ok = rollout_sim<MAXJ>(c, g, kind);
  if (!ok) {
    g.status = ST_ERROR;
    if (c.lane == 0) gm->error = 1;
  }
  for (int a = c.lane; a < P.A; a += 32) {
    visits[(size_t)w * P.A + a] = 0u;
    wsum[(size_t)w * P.A + a] = 0.f;
  }
  __syncwarp();
  if (g.root_node >= 0) {
    const size_t base = c.abase + (size_t)g.root_node;
    const int L = P.A - g.n_moves;
    for (int i = c.lane; i < L; i += 32) {
      const int a = P.slot_act[base + i];
      const uint2 nw = P.slot_nw[base + i];
      visits[(size_t)w * P.A + a] = nw.x;
      wsum[(size_t)w * P.A + a] = __uint_as_float(nw.y);
    }
  }
  if (c.lane == 0) {
    gm->is_real_root = 1;
    gm->auto_play = 0;
    gm->sims_total += (unsigned long long)g.sims_done;
  }
  __syncwarp();
  store_regs<1>(c, g);
}

// visit[a] = n(child a), policy[a] = p(child a)  (agents.py:64-73)
__global__ void export_roots_kernel(TreeParams P, const int32_t* __restrict__ ids, int n, uint32_t* __restrict__ visits,
                                    double* __restrict__ priors, int32_t* __restrict__ real_root) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const int game = ids[w];
  const Game* gm = &P.games[game];
  for (int a = lane; a < P.A; a += 32) {
    if (visits) visits[(size_t)w * P.A + a] = 0u;
    if (priors) priors[(size_t)w * P.A + a] = 0.0;
  }
  __syncwarp();
  if (gm->root_node >= 0) {
    const size_t base = arena_base(P, game, gm->arena) + (size_t)gm->root_node;
    const int L = P.A - gm->n_moves;
    for (int i = lane; i < L; i += 32) {
      const int a = P.slot_act[base + i];
      if (visits) visits[(size_t)w * P.A + a] = P.slot_nw[base + i].x;
      if (priors) priors[(size_t)w * P.A + a] = P.slot_p[base + i];
    }
  }
  if (lane == 0 && real_root) real_root[w] = gm->is_real_root;
}

// eval_main.main's start (eval_main.py:213-229): both agents empty, the player is black in even slots' first match
__global__ void reset_arena_kernel(TreeParams P, int n_slots, uint32_t first_key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n_slots) return;
  const int side = i >= n_slots ? 1 : 0, m = i - side * n_slots;
  Game* gm = &P.games[side * P.arena_M + m];
  for (int r = 0; r < kRowsPad; ++r) {
    gm->rows_b[r] = 0; gm->rows_w[r] = 0; gm->leaf_rows_b[r] = 0; gm->leaf_rows_w[r] = 0;
  }
  gm->n_moves = 0; gm->last1 = -1; gm->last2 = -1;
  gm->status = ST_SEARCH;
  gm->arena = 0; gm->root_node = CH_UNVISITED; gm->root_n = 0u; gm->root_w = 0.f; gm->slot_count = 0u;
  gm->sims_done = 0; gm->sims_target = P.arena_num_mcts[side] + 1; gm->is_real_root = 1; gm->auto_play = 4;
  gm->rng_ctr = 0u; gm->noise_draws = 0u; gm->game_key = first_key + 2u * (uint32_t)m + (uint32_t)side;
  gm->leaf_depth = 0; gm->leaf_n_moves = 0; gm->nn_slot = 0; gm->nn_ticket = 0u; gm->nn_static = 0; gm->winner = 0; gm->error = 0; gm->nn_log_count = 0u;
  gm->sims_total = 0ull; gm->nn_evals = 0ull; gm->terminal_sims = 0ull; gm->moves_played = 0ull;
  gm->games_finished = 0ull;
  gm->arena_match = 0;
  gm->arena_player_black = (m & 1) == 0 ? 1 : 0;
  gm->arena_cur = gm->arena_player_black ? 0 : 1;
}

__global__ void sum_counters_kernel(TreeParams P, int n, int n_running, unsigned long long* out) {
  unsigned long long sims = 0, running = 0, evals = 0, errors = 0, moves = 0, fin = 0, term = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Game* gm = &P.games[i];
    sims += gm->sims_total;
    evals += gm->nn_evals;
    moves += gm->moves_played;
    fin += gm->games_finished;
    term += gm->terminal_sims;
    running += (i < n_running && (gm->status == ST_SEARCH || gm->status == ST_WAIT_NN)) ? 1ull : 0ull;
    errors += gm->status == ST_ERROR ? 1ull : 0ull;
  }
  atomicAdd(&out[0], sims);
  atomicAdd(&out[1], running);
  atomicAdd(&out[2], evals);
  atomicAdd(&out[3], errors);
  atomicAdd(&out[4], moves);
  atomicAdd(&out[5], fin);
  atomicAdd(&out[6], term);
}

// replay record slab: {int16 n_moves, int8 winner, int8 pad, int16 moves[A], (pad to 4) uint32 visits[A][A]}
__global__ void pack_records_kernel(TreeParams P, int n, uint8_t* out, size_t bytes_per_game) {
  const int game = blockIdx.x;
  if (game >= n) return;
  const Game* gm = &P.games[game];
  uint8_t* rec = out + (size_t)game * bytes_per_game;
  int16_t* hdr = reinterpret_cast<int16_t*>(rec);
  if (threadIdx.x == 0) {
    hdr[0] = (int16_t)gm->n_moves;
    rec[2] = (uint8_t)gm->winner;
    rec[3] = 0;
  }
  int16_t* mv = hdr + 2;
  for (int i = threadIdx.x; i < P.A; i += blockDim.x) mv[i] = i < gm->n_moves ? (int16_t)gm->moves[i] : (int16_t)-1;
  const size_t voff = (4 + (size_t)P.A * 2 + 3) & ~(size_t)3;
  if (threadIdx.x == 0)
    for (size_t i = 4 + (size_t)P.A * 2; i < voff; ++i) rec[i] = 0;  // alignment padding: records compare byte for byte
  uint32_t* vis = reinterpret_cast<uint32_t*>(rec + voff);
  const uint32_t* src = P.rec_visits + (size_t)game * P.A * P.A;
  for (int i = threadIdx.x; i < P.A * P.A; i += blockDim.x) vis[i] = (i / P.A) < gm->n_moves ? src[i] : 0u;
}

}  // namespace

cudaError_t launch_tree_step(const TreeParams& p, const int32_t* game_ids, int n, int max_iters, cudaStream_t s) {
  const int blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (p.A <= 96) tree_step_kernel<3><<<blocks, kWarpsPerBlock * 32, 0, s>>>(p, game_ids, n, max_iters);
  else tree_step_kernel<8><<<blocks, kWarpsPerBlock * 32, 0, s>>>(p, game_ids, n, max_iters);
  return cudaGetLastError();
}
cudaError_t launch_set_roots(const TreeParams& p, const int32_t* ids, int n, const int16_t* roots, const int32_t* lens,
                             cudaStream_t s) {
  const int blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  set_roots_kernel<<<blocks, kWarpsPerBlock * 32, 0, s>>>(p, ids, n, roots, lens);
  return cudaGetLastError();
}
cudaError_t launch_export_roots(const TreeParams& p, const int32_t* ids, int n, uint32_t* visits, double* priors,
                                int32_t* real_root, cudaStream_t s) {
  export_roots_kernel<<<(n * 32 + 127) / 128, 128, 0, s>>>(p, ids, n, visits, priors, real_root);
  return cudaGetLastError();
}
cudaError_t launch_reset_games(const TreeParams& p, const int32_t* ids, int n, const uint32_t* keys, int auto_play,
                               cudaStream_t s) {
  reset_games_kernel<<<(n + 127) / 128, 128, 0, s>>>(p, ids, n, keys, auto_play);
  return cudaGetLastError();
}
cudaError_t launch_sum_counters(const TreeParams& p, int n, int n_running, unsigned long long* out5, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out5, 0, 8 * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  sum_counters_kernel<<<32, 128, 0, s>>>(p, n, n_running, out5);
  return cudaGetLastError();
}
cudaError_t launch_rollout_search(const TreeParams& p, int kind, int num_mcts, const int32_t* ids, int n,
                                  const int16_t* roots, const int32_t* lens, uint32_t* visits, float* w, cudaStream_t s) {
  const int blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (p.A <= 96) rollout_search_kernel<3><<<blocks, kWarpsPerBlock * 32, 0, s>>>(p, kind, num_mcts, ids, n, roots, lens, visits, w);
  else rollout_search_kernel<8><<<blocks, kWarpsPerBlock * 32, 0, s>>>(p, kind, num_mcts, ids, n, roots, lens, visits, w);
  return cudaGetLastError();
}
cudaError_t launch_reset_arena(const TreeParams& p, int n_slots, uint32_t first_key, cudaStream_t s) {
  reset_arena_kernel<<<(2 * n_slots + 127) / 128, 128, 0, s>>>(p, n_slots, first_key);
  return cudaGetLastError();
}
cudaError_t launch_pack_records(const TreeParams& p, int n, uint8_t* out, size_t bytes_per_game, cudaStream_t s) {
  pack_records_kernel<<<n, 128, 0, s>>>(p, n, out, bytes_per_game);
  return cudaGetLastError();
}

}  // namespace ao
