// Device side of the tree search (included by tree.cu and by the persistent self-play kernel in tower_stag.cu).
//
// Warp-per-game PUCT search over ID-keyed trees resident in HBM (the replacement of agents.py:60-239) plus the
// per-move step of main.py:150-196 (pi, action sampling, env.step, root advance with subtree compaction) and the
// arena's match step (eval_main.py:243-283).
//
// Tree storage (SoA, per game two arenas of slot_cap slots that ping-pong at every root advance):
//   a node that has been expanded owns a contiguous block of L child slots (L = empty cells = A - stones, so the
//   block length is implied by depth); slot i = {act u8, n u32, w f32, p f64, child i32}. `child` is the arena-
//   relative offset of the child's own block, CH_UNVISITED, or -(1+win_index) for a terminal child.
//   The statistics of a node live in its parent's slot (the root's in Game.root_n / root_w), which is equivalent
//   to the reference's dict-of-nodes (agents.py:87-91,206-210).
// Arithmetic follows SURVEY appendix A.2: n exact, w/q float32 (round-to-nearest, no FMA), p/u/q+u float64.
#pragma once
#include <stdio.h>

#include "engine.h"
#include "rules.cuh"

namespace ao {
namespace {


// Per-warp scratch in shared memory, by reference: the caller chooses where the arrays live (tree.cu: one WarpSmemStore
// per warp; the persistent self-play kernel of tower_stag.cu: a compact layout sized for its board inside its own smem).
struct WarpSmem {
  double* dbuf;            // [A] priors indexed by action / pi / cdf
  double* dbuf2;           // [A] Dirichlet noise indexed by child position
  float* pol;              // [A] NN policy of the leaf being expanded
  uint8_t* order;          // [256] child order (actions) of the node being expanded (+ resize parking area)
  int16_t* table;          // [128] CPython set emulation scratch
  uint16_t (*rows)[32];    // [2][32] row-mask scratch
};
struct WarpSmemStore {
  double dbuf[kMaxA + 7];
  double dbuf2[kMaxA + 7];
  float pol[kMaxA + 7];
  uint8_t order[256];
  int16_t table[128];
  uint16_t rows[2][32];
  __device__ __forceinline__ WarpSmem ref() { return WarpSmem{dbuf, dbuf2, pol, order, table, rows}; }
};

constexpr int kWarpsPerBlock = 4;

// probe build: cycle counters of the phases of tree_step_game (lane 0; one copy per translation unit, printed by the
// solo kernel at the end of a launch when ao_tower_debug is on)
#ifdef AO_PROBE
__device__ unsigned long long g_tree_dbg[12];
__device__ int g_tree_dbg_on;
#define AO_TDBG_BEGIN long long tdbg_t = g_tree_dbg_on ? clock64() : 0;
#define AO_TDBG(k)                                                      \
  if (g_tree_dbg_on && (threadIdx.x & 31) == 0) {                        \
    const long long tdbg_n = clock64();                                 \
    atomicAdd(&g_tree_dbg[k], (unsigned long long)(tdbg_n - tdbg_t));   \
    tdbg_t = tdbg_n;                                                    \
  }
#else
#define AO_TDBG_BEGIN
#define AO_TDBG(k)
#endif

struct Ctx {
  const TreeParams& P;
  Game* gm;
  WarpSmem* sm;
  int lane, game;
  size_t abase;  // first slot of the live arena
  uint2 key;
};

__device__ __forceinline__ size_t arena_base(const TreeParams& P, int game, int arena) {
  return ((size_t)game * 2 + (size_t)arena) * (size_t)P.slot_cap;
}

// one block of the decision stream (oracle.DecisionStream._block)
__device__ __forceinline__ uint4 draw_block(Ctx& c, uint32_t& rng_ctr) {
  const uint4 r = philox4x32(make_uint4(rng_ctr, 0u, c.gm->game_key, 0u), c.key);
  ++rng_ctr;
  return r;
}
__device__ __forceinline__ int draw_choice(Ctx& c, uint32_t& rng_ctr, int k) {
  if (k <= 1) return 0;
  const uint4 r = draw_block(c, rng_ctr);
  return (int)(((unsigned long long)r.x * (unsigned long long)k) >> 32);
}
__device__ __forceinline__ double draw_uniform53(Ctx& c, uint32_t& rng_ctr) {
  const uint4 r = draw_block(c, rng_ctr);
  return __ddiv_rn(__dadd_rn(__dmul_rn((double)(r.x >> 5), 67108864.0), (double)(r.y >> 6)), 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------------- Dirichlet
// Gamma(alpha,1), alpha < 1: Marsaglia-Tsang on alpha+1, boosted by U^(1/alpha). Stream 1 of the game's Philox key.
__device__ double device_gamma(const Ctx& c, uint32_t draw, uint32_t idx, double alpha) {
  const double d = alpha + 1.0 - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
  for (uint32_t attempt = 0; attempt < 64; ++attempt) {
    const uint4 r = philox4x32(make_uint4(draw, (idx << 8) | attempt, c.gm->game_key, 1u), c.key);
    const double u1 = ((double)r.x + 0.5) * (1.0 / 4294967296.0), u2 = ((double)r.y + 0.5) * (1.0 / 4294967296.0);
    const double u3 = ((double)r.z + 0.5) * (1.0 / 4294967296.0), u4 = ((double)r.w + 0.5) * (1.0 / 4294967296.0);
    const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    double v = 1.0 + cc * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (log(u3) < 0.5 * x * x + d - d * v + d * log(v)) {
      const double g = d * v * exp(log(u4) / alpha);
      return g > 1e-300 ? g : 1e-300;
    }
  }
  return 1e-300;
}

// eta[0..L) -> sm->dbuf2 ; consumes one Dirichlet draw (none when L == 0, like np.random.dirichlet([]))
__device__ void draw_dirichlet(Ctx& c, int L, uint32_t& noise_draws) {
  if (L == 0) return;
  const TreeParams& P = c.P;
  double* eta = c.sm->dbuf2;
  if (P.noise_mode == AO_NOISE_TAPE) {
    const uint32_t row = noise_draws < (uint32_t)P.tape_rows ? noise_draws : (uint32_t)P.tape_rows - 1u;
    const double* g = P.gamma_tape + ((size_t)c.game * P.tape_rows + row) * P.A;
    for (int i = c.lane; i < L; i += 32) eta[i] = g[i];
  } else {
    for (int i = c.lane; i < L; i += 32) eta[i] = device_gamma(c, noise_draws, (uint32_t)i, P.alpha);
  }
  ++noise_draws;
  __syncwarp();
  double inv = 0.0;
  if (c.lane == 0) {
    double acc = 0.0;
    for (int i = 0; i < L; ++i) acc = __dadd_rn(acc, eta[i]);
    inv = __ddiv_rn(1.0, acc);
  }
  inv = __shfl_sync(kFull, inv, 0);
  for (int i = c.lane; i < L; i += 32) eta[i] = __dmul_rn(eta[i], inv);
  __syncwarp();
}

// p_i <- 0.75 p_i + 0.25 eta_i on the root's existing children (agents.py:95-103)
__device__ void remix_root_noise(Ctx& c, int32_t root_node, int L, uint32_t& noise_draws) {
  draw_dirichlet(c, L, noise_draws);
  double* sp = c.P.slot_p + c.abase + (size_t)root_node;
  for (int i = c.lane; i < L; i += 32)
    sp[i] = __dadd_rn(__dmul_rn(0.75, sp[i]), __dmul_rn(0.25, c.sm->dbuf2[i]));
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------- synthetic NN
__device__ void synth_eval(Ctx& c, const uint8_t* root_moves, int n_root, int depth, float& value, uint32_t salt) {
  // FNV-1a style fold of the move sequence (root moves then the path actions), then Philox per action.
  unsigned long long hh = 0xCBF29CE484222325ull;
  for (int i = 0; i < n_root; ++i) hh = (hh ^ (unsigned long long)(root_moves[i] + 1)) * 0x100000001B3ull;
  const uint32_t* path = c.P.path + (size_t)c.game * (c.P.A + 1);
  for (int i = 0; i < depth; ++i) {
    const unsigned a = c.P.slot_act[path[i]];
    hh = (hh ^ (unsigned long long)(a + 1)) * 0x100000001B3ull;
  }
  const uint32_t lo = (uint32_t)hh, hi = (uint32_t)(hh >> 32);
  const uint2 k = make_uint2(0x5EEDu, 0x0A0Au + salt);
  for (int a = c.lane; a < c.P.A; a += 32) {
    const uint4 r = philox4x32(make_uint4((uint32_t)a, 0u, lo, hi), k);
    c.sm->pol[a] = (float)((r.x >> 8) + 1u) * 5.9604644775390625e-08f;  // 2^-24, exact
  }
  const uint4 r = philox4x32(make_uint4(0xFFFFu, 0u, lo, hi), k);
  value = __fadd_rn((float)(r.y >> 8) * 1.1920928955078125e-07f, -1.0f);  // 2^-23, exact
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------- search state in registers
struct Regs {
  int32_t status, root_node, sims_done, sims_target, n_moves, last1, last2, arena;
  uint32_t root_n, slot_count, rng_ctr, noise_draws;
  float root_w;
  uint32_t rb, rw;  // this lane's root row masks
};

// Expansion (agents.py:180-212) of the pending leaf with sm->pol / value, then backup (agents.py:223-239).
// leaf rows are in sm->rows. Returns false on arena overflow.
__device__ bool expand_and_backup(Ctx& c, Regs& g, int depth, float value, bool terminal) {
  const TreeParams& P = c.P;
  uint32_t* path = P.path + (size_t)c.game * (P.A + 1);
  float delta;
  AO_TDBG_BEGIN
  if (!terminal) {
    WarpSmem* sm = c.sm;
    // occupancy rows -> legal actions in reference child order
    if (c.lane < 32) sm->rows[0][c.lane] = (uint16_t)(sm->rows[0][c.lane] | sm->rows[1][c.lane]);
    __syncwarp();
    const int L = legal_order(sm->rows[0], P.B, P.A, sm->order, sm->table, c.lane);
    AO_TDBG(2)
    // prior_prob = zeros(A); prior_prob[legal] = policy[legal]; prior_prob /= prior_prob.sum()   (agents.py:183-189)
    for (int a = c.lane; a < P.A; a += 32) {
      const bool legal = ((sm->rows[0][a / P.B] >> (a % P.B)) & 1u) == 0u;
      sm->dbuf[a] = legal ? (double)sm->pol[a] : 0.0;
    }
    __syncwarp();
    const double S = np_pairwise_sum(sm->dbuf, P.A, c.lane);
    AO_TDBG(3)
    const bool is_root = depth == 0;
    const bool mix = is_root && P.noise;
    if (mix) draw_dirichlet(c, L, g.noise_draws);
    if (g.slot_count + (uint32_t)L > P.slot_cap) return false;
    const uint32_t off = g.slot_count;
    g.slot_count += (uint32_t)L;
    const size_t base = c.abase + off;
    for (int i = c.lane; i < L; i += 32) {
      const int a = sm->order[i];
      double pr = __ddiv_rn(sm->dbuf[a], S);
      if (mix) pr = __dadd_rn(__dmul_rn(0.75, pr), __dmul_rn(0.25, sm->dbuf2[i]));
      P.slot_act[base + i] = (uint8_t)a;
      P.slot_nw[base + i] = make_uint2(0u, 0u);
      P.slot_p[base + i] = pr;
      P.slot_child[base + i] = CH_UNVISITED;
    }
    if (is_root) g.root_node = (int32_t)off;
    else if (c.lane == 0) P.slot_child[path[depth - 1]] = (int32_t)off;
    delta = -value;
    AO_TDBG(4)
  } else {
    delta = 1.0f;  // reward 1.0 for wins and draws alike (agents.py:216-221)
  }
  // backup: leaf slot gets +delta, alternating up to and including the root.  Every level of the path is a slot of its
  // own, so the lanes take the levels in parallel: one round trip to memory for the whole path instead of two per level
  // (the arithmetic per slot - one float add of +-delta - is the sequential loop's).
  for (int i0 = 0; i0 < depth; i0 += 32) {
    const int i = i0 + c.lane;
    if (i < depth) {
      const uint32_t s = path[i];
      uint2 nw = P.slot_nw[s];
      nw.x += 1u;
      nw.y = __float_as_uint(__fadd_rn(__uint_as_float(nw.y), ((depth - 1 - i) & 1) ? -delta : delta));
      P.slot_nw[s] = nw;
    }
  }
  g.root_w = __fadd_rn(g.root_w, (depth & 1) ? -delta : delta);  // the same value in every lane
  g.root_n += 1u;
  g.sims_done += 1;
  __syncwarp();
  AO_TDBG(5)
  return true;
}

// Copy the subtree whose root block sits at `old_root` (arena-relative, block length L0) from the live arena into
// the other one (breadth first), flip arenas. Returns the new root offset (0).
__device__ void compact_subtree(Ctx& c, Regs& g, int32_t old_root, int L0) {
  const TreeParams& P = c.P;
  const size_t src = c.abase;
  const size_t dst = arena_base(P, c.game, g.arena ^ 1);
  uint32_t* q_old = P.gc_old + (size_t)c.game * P.gc_cap;
  uint32_t* q_new = P.gc_new + (size_t)c.game * P.gc_cap;
  if (c.lane == 0) {
    q_old[0] = ((uint32_t)L0 << 20) | (uint32_t)old_root;
    q_new[0] = 0u;
  }
  __syncwarp();
  uint32_t head = 0, tail = 1, dcount = (uint32_t)L0;
  bool overflow = false;
  while (head < tail) {
    const uint32_t packed = q_old[head];
    const uint32_t noff = q_new[head];
    ++head;
    const int L = (int)(packed >> 20);
    const uint32_t ooff = packed & 0xFFFFFu;
    const int Lc = L - 1;
    for (int b = 0; b < L; b += 32) {
      const int i = b + c.lane;
      int32_t ch = CH_UNVISITED;
      if (i < L) {
        ch = P.slot_child[src + ooff + i];
        P.slot_act[dst + noff + i] = P.slot_act[src + ooff + i];
        P.slot_nw[dst + noff + i] = P.slot_nw[src + ooff + i];
        P.slot_p[dst + noff + i] = P.slot_p[src + ooff + i];
      }
      const bool has = i < L && ch >= 0;
      const unsigned bal = __ballot_sync(kFull, has);
      if (has) {
        const uint32_t rank = __popc(bal & ((1u << c.lane) - 1u));
        const uint32_t no = dcount + rank * (uint32_t)Lc;
        if (tail + rank < P.gc_cap) {
          q_old[tail + rank] = ((uint32_t)Lc << 20) | (uint32_t)ch;
          q_new[tail + rank] = no;
        }
        ch = (int32_t)no;
      }
      if (i < L) P.slot_child[dst + noff + i] = ch;
      const uint32_t cnt = __popc(bal);
      if (tail + cnt > P.gc_cap) overflow = true;  // more expanded nodes than queue entries: report, never corrupt
      tail += cnt;
      dcount += cnt * (uint32_t)Lc;
    }
    __syncwarp();
    if (overflow) break;
  }
  if (overflow) {
    g.status = ST_ERROR;
    if (c.lane == 0) c.gm->error = 2;
  }
  g.arena ^= 1;
  c.abase = dst;
  g.slot_count = dcount;
}

// Advance the root by one action: stones, stats, subtree compaction (tree reuse of the reference's persistent dict).
// Returns false when the new root is not "in the tree" (only possible in facade mode).
// keep_tree = false (RandomAgent / PUCT / UCT sides, which never reuse a tree): only the stone is placed.
__device__ bool advance_root(Ctx& c, Regs& g, int action, bool keep_tree = true) {
  const TreeParams& P = c.P;
  const int L = P.A - g.n_moves;  // children of the current root
  int32_t child = CH_UNVISITED;
  uint32_t cn = 0;
  float cw = 0.f;
  bool found = false;
  if (keep_tree && g.root_node >= 0) {
    const size_t base = c.abase + (size_t)g.root_node;
    for (int b = 0; b < L; b += 32) {
      const int i = b + c.lane;
      const bool hit = i < L && P.slot_act[base + i] == (uint8_t)action;
      const unsigned bal = __ballot_sync(kFull, hit);
      if (bal) {
        const int src_lane = __ffs(bal) - 1;
        uint2 nw = make_uint2(0u, 0u);
        int32_t ch = 0;
        if (hit) {
          nw = P.slot_nw[base + i];
          ch = P.slot_child[base + i];
        }
        cn = __shfl_sync(kFull, nw.x, src_lane);
        cw = __uint_as_float(__shfl_sync(kFull, nw.y, src_lane));
        child = __shfl_sync(kFull, ch, src_lane);
        found = true;
        break;
      }
    }
  }
  // place the stone on the root position
  const int y = action / P.B, x = action % P.B;
  const bool black = (g.n_moves & 1) == 0;
  if (c.lane == y) {
    if (black) g.rb |= 1u << x;
    else g.rw |= 1u << x;
  }
  if (c.lane == 0) c.gm->moves[g.n_moves] = (uint8_t)action;
  g.n_moves += 1;
  g.last2 = g.last1;
  g.last1 = action;
  if (!found) {
    g.root_node = CH_UNVISITED;
    g.root_n = 0u;
    g.root_w = 0.f;
    g.slot_count = 0u;
    return false;
  }
  g.root_n = cn;
  g.root_w = cw;
  if (child >= 0) {
    compact_subtree(c, g, child, L - 1);
    g.root_node = 0;
  } else {
    g.root_node = child;  // unvisited or terminal: nothing below it
    g.slot_count = 0u;
  }
  return true;
}

template <int MAXJ>
__device__ void store_regs(Ctx& c, const Regs& g) {
  Game* gm = c.gm;
  if (c.lane < kRowsPad) {
    gm->rows_b[c.lane] = (uint16_t)g.rb;
    gm->rows_w[c.lane] = (uint16_t)g.rw;
  }
  if (c.lane == 0) {
    gm->status = g.status;
    gm->root_node = g.root_node;
    gm->sims_done = g.sims_done;
    gm->sims_target = g.sims_target;
    gm->n_moves = g.n_moves;
    gm->last1 = g.last1;
    gm->last2 = g.last2;
    gm->arena = g.arena;
    gm->root_n = g.root_n;
    gm->root_w = g.root_w;
    gm->slot_count = g.slot_count;
    gm->rng_ctr = g.rng_ctr;
    gm->noise_draws = g.noise_draws;
  }
}

__device__ void load_regs(Ctx& c, Regs& g) {
  const Game* gm = c.gm;
  g.status = gm->status;
  g.root_node = gm->root_node;
  g.sims_done = gm->sims_done;
  g.sims_target = gm->sims_target;
  g.n_moves = gm->n_moves;
  g.last1 = gm->last1;
  g.last2 = gm->last2;
  g.arena = gm->arena;
  g.root_n = gm->root_n;
  g.root_w = gm->root_w;
  g.slot_count = gm->slot_count;
  g.rng_ctr = gm->rng_ctr;
  g.noise_draws = gm->noise_draws;
  g.rb = c.lane < kRowsPad ? gm->rows_b[c.lane] : 0u;
  g.rw = c.lane < kRowsPad ? gm->rows_w[c.lane] : 0u;
}

// End of a search in self-play mode: pi, action, env.step, root advance (main.py:150-196, agents.py:64-80).
template <int MAXJ>
__device__ void play_move(Ctx& c, Regs& g) {
  const TreeParams& P = c.P;
  WarpSmem* sm = c.sm;
  const int A = P.A;
  const int L = A - g.n_moves;
  // visit[a] = n(child a)
  for (int a = c.lane; a < A; a += 32) sm->dbuf[a] = 0.0;
  __syncwarp();
  uint32_t* rec = P.rec_visits + ((size_t)c.game * A + (size_t)g.n_moves) * A;
  for (int a = c.lane; a < A; a += 32) rec[a] = 0u;
  __syncwarp();
  if (g.root_node >= 0) {
    const size_t base = c.abase + (size_t)g.root_node;
    for (int i = c.lane; i < L; i += 32) {
      const int a = P.slot_act[base + i];
      const uint32_t n = P.slot_nw[base + i].x;
      sm->dbuf[a] = (double)n;
      rec[a] = n;
    }
  }
  __syncwarp();
  int action = 0;
  const int tau = g.n_moves < P.tau_thres ? 1 : 0;
  if (tau == 1) {
    // pi = visit / visit.sum(); action = np.random.choice(A, p=pi)  (utils.py:189-195, legacy RandomState.choice)
    double total = 0.0;  // exact: integer-valued
    for (int a = c.lane; a < A; a += 32) total += sm->dbuf[a];
#pragma unroll
    for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
    const double u = draw_uniform53(c, g.rng_ctr);
    if (c.lane == 0) {
      double acc = 0.0;
      for (int a = 0; a < A; ++a) {  // cdf = cumsum(pi) sequential
        acc = __dadd_rn(acc, __ddiv_rn(sm->dbuf[a], total));
        sm->dbuf2[a] = acc;
      }
      const double last = sm->dbuf2[A - 1];
      int idx = A;
      for (int a = 0; a < A; ++a) {
        if (__ddiv_rn(sm->dbuf2[a], last) > u) {  // searchsorted(cdf / cdf[-1], u, side='right')
          idx = a;
          break;
        }
      }
      action = idx < A ? idx : A - 1;
    }
    action = __shfl_sync(kFull, action, 0);
  } else {
    // pi = one-hot(argmax with uniform tie-break over ascending indices) (utils.py:198-205); get_action on a
    // one-hot pi returns that index but still consumes one uniform.
    double mx = 0.0;
    for (int a = c.lane; a < A; a += 32) mx = fmax(mx, sm->dbuf[a]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(kFull, mx, o));
    int K = 0;
    for (int b = 0; b < A; b += 32) {
      const int a = b + c.lane;
      K += __popc(__ballot_sync(kFull, a < A && sm->dbuf[a] == mx));
    }
    int r = draw_choice(c, g.rng_ctr, K);
    for (int b = 0; b < A; b += 32) {
      const int a = b + c.lane;
      const unsigned bal = __ballot_sync(kFull, a < A && sm->dbuf[a] == mx);
      const int cnt = __popc(bal);
      if (r < cnt) {
        action = b + (int)__fns(bal, 0, r + 1);
        break;
      }
      r -= cnt;
    }
    (void)draw_block(c, g.rng_ctr);  // the uniform consumed by get_action
  }
  __syncwarp();
  // env.step: place stone, check_win (env_small.py:155-196)
  const bool in_tree = advance_root(c, g, action);
  (void)in_tree;
  if (c.lane == 0) c.gm->moves_played += 1ull;
  const int win = check_win_rows(g.rb, g.rw, P.B, g.n_moves, sm->rows, c.lane);
  if (win != 0) {
    if (c.lane == 0) {
      c.gm->winner = win;
      c.gm->games_finished += 1ull;
    }
    const int mode = c.gm->auto_play;
    if (mode != 2 && mode != 3) {
      g.status = ST_FINISHED;
      return;
    }
    uint32_t next_key = c.gm->game_key + (uint32_t)P.G;  // bench mode (2): the slot's next decision-stream key
    if (mode == 3) {
      // continuous self-play: the finished episode's record goes to index (key - first_key) of the stream slab (same
      // layout as pack_records_kernel), then the slot takes the next unplayed key - or retires when none is left
      __syncwarp();  // lane 0's store of the last move (advance_root) must be visible to the lanes that copy the moves
      const uint32_t key = c.gm->game_key;
      uint8_t* out = P.stream_out + (size_t)(key - P.stream_first_key) * P.stream_rec_bytes;
      if (c.lane == 0) {
        reinterpret_cast<int16_t*>(out)[0] = (int16_t)g.n_moves;
        out[2] = (uint8_t)win;
        out[3] = 0;
      }
      int16_t* mv = reinterpret_cast<int16_t*>(out) + 2;
      for (int i = c.lane; i < A; i += 32) mv[i] = i < g.n_moves ? (int16_t)c.gm->moves[i] : (int16_t)-1;
      uint32_t* vis = reinterpret_cast<uint32_t*>(out + ((4 + (size_t)A * 2 + 3) & ~(size_t)3));
      const uint32_t* src = P.rec_visits + (size_t)c.game * A * A;
      const int n_words = g.n_moves * A;
      for (int i = c.lane; i < A * A; i += 32) vis[i] = i < n_words ? src[i] : 0u;
      uint32_t nk = 0u;
      if (c.lane == 0) nk = atomicAdd(P.stream_next_key, 1u);
      next_key = __shfl_sync(kFull, nk, 0);
      if (next_key >= P.stream_key_end) {
        g.status = ST_FINISHED;
        return;
      }
    }
    // recycle the slot: a fresh episode
    g.rb = 0u; g.rw = 0u;
    g.n_moves = 0; g.last1 = -1; g.last2 = -1;
    g.root_node = CH_UNVISITED; g.root_n = 0u; g.root_w = 0.f; g.slot_count = 0u;
    g.sims_done = 0; g.sims_target = P.num_mcts + 1;
    g.rng_ctr = 0u; g.noise_draws = 0u;
    if (c.lane == 0) {
      c.gm->game_key = next_key;
      c.gm->is_real_root = 1;
      c.gm->winner = 0;
    }
    __syncwarp();
    return;
  }
  // next search: reused root (agents.py:93-103) -> num_mcts sims, Dirichlet re-mixed into the existing priors
  g.sims_done = 0;
  g.sims_target = P.num_mcts;
  if (c.lane == 0) c.gm->is_real_root = 0;
  if (P.noise && g.root_node >= 0) remix_root_noise(c, g.root_node, P.A - g.n_moves, g.noise_draws);
}

#include "rollout.cuh"

// ---------------------------------------------------------------------------------------------- arena
__device__ __forceinline__ int arena_side(const TreeParams& P, int game) { return game >= P.arena_M ? 1 : 0; }

// pi = one-hot(argmax visits, uniform tie-break over ascending indices) (utils.py:198-205) from sm->dbuf
__device__ int argmax_tiebreak(Ctx& c, Regs& g, int A) {
  WarpSmem* sm = c.sm;
  double mx = __longlong_as_double(0xFFF0000000000000ll);  // -inf: UCT scores can all be negative
  for (int a = c.lane; a < A; a += 32) mx = fmax(mx, sm->dbuf[a]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(kFull, mx, o));
  int K = 0;
  for (int b = 0; b < A; b += 32) {
    const int a = b + c.lane;
    K += __popc(__ballot_sync(kFull, a < A && sm->dbuf[a] == mx));
  }
  int r = draw_choice(c, g.rng_ctr, K);
  int action = 0;
  for (int b = 0; b < A; b += 32) {
    const int a = b + c.lane;
    const unsigned bal = __ballot_sync(kFull, a < A && sm->dbuf[a] == mx);
    const int cnt = __popc(bal);
    if (r < cnt) {
      action = b + (int)__fns(bal, 0, r + 1);
      break;
    }
    r -= cnt;
  }
  return action;
}

// The side whose search just finished (or a RandomAgent side) moves: eval_main.py:243-283.
//   pi = get_pi(root_id, tau=0); action = argmax_onehot(pi)   (one tie-break draw when several maxima, no other draw)
//   root_id = mover.root_id + (action,); env.step; then the OTHER side's next get_pi(root_id) finds the new root in its
//   own tree (reused root, possibly with n == 0: agents.py:93-111) or not (real root).
// One warp owns the match, so it updates both sides' trees in turn and goes on searching as the other side.
// Returns true when the warp can go on (as the other side) within this launch.
__device__ bool arena_move(Ctx& c, Regs& g) {
  const TreeParams& P = c.P;
  WarpSmem* sm = c.sm;
  const int A = P.A, M = P.arena_M;
  const int side = arena_side(P, c.game);
  const int m = c.game - side * M;
  Game* pg = &P.games[m];  // player slot: match-level fields
  uint32_t* rec = P.rec_visits + ((size_t)m * A + (size_t)g.n_moves) * A;
  for (int a = c.lane; a < A; a += 32) {
    sm->dbuf[a] = 0.0;
    rec[a] = 0u;
  }
  __syncwarp();
  int action;
  const int kind = P.arena_kind[side];
  if (kind == SIDE_PUCT || kind == SIDE_UCT) {
    rollout_scores(c, g, kind, rec);
    action = argmax_tiebreak(c, g, A);
  } else if (kind == SIDE_RANDOM) {
    // RandomAgent.get_pi: uniform over the empty cells; argmax_onehot then draws one of them (ascending order)
    const int L = A - g.n_moves;
    int r = draw_choice(c, g.rng_ctr, L);
    action = 0;
    for (int b = 0; b < A; b += 32) {
      const int a = b + c.lane;
      const uint32_t occ = __shfl_sync(kFull, g.rb | g.rw, a < A ? a / P.B : 0);
      const bool empty = a < A && ((occ >> (a % P.B)) & 1u) == 0u;
      const unsigned bal = __ballot_sync(kFull, empty);
      const int cnt = __popc(bal);
      if (r < cnt) {
        action = b + (int)__fns(bal, 0, r + 1);
        break;
      }
      r -= cnt;
    }
    action = __shfl_sync(kFull, action, 0);
  } else {
    if (g.root_node >= 0) {
      const size_t base = c.abase + (size_t)g.root_node;
      const int L = A - g.n_moves;
      for (int i = c.lane; i < L; i += 32) {
        const int a = P.slot_act[base + i];
        const uint32_t n = P.slot_nw[base + i].x;
        sm->dbuf[a] = (double)n;
        rec[a] = n;
      }
    }
    __syncwarp();
    action = argmax_tiebreak(c, g, A);
  }
  __syncwarp();
  (void)advance_root(c, g, action, kind == SIDE_ZERO);  // mover.root_id + (action,)
  if (c.lane == 0) c.gm->moves_played += 1ull;
  const int win = check_win_rows(g.rb, g.rw, P.B, g.n_moves, sm->rows, c.lane);
  const int o = side ? m : M + m;  // the other side's slot
  Game* og = &P.games[o];
  if (win != 0) {
    __syncwarp();  // lane 0's store of the last move (advance_root) must be visible to the lanes that copy the moves
    const int k = pg->arena_match;
    if (P.stream_out) {  // record of the finished match: same layout as pack_records_kernel, index = slot * mps + k
      uint8_t* out = P.stream_out + ((size_t)m * P.arena_matches_per_slot + (size_t)k) * P.stream_rec_bytes;
      if (c.lane == 0) {
        reinterpret_cast<int16_t*>(out)[0] = (int16_t)g.n_moves;
        out[2] = (uint8_t)win;
        out[3] = (uint8_t)(pg->arena_player_black ? 1 : 0);
      }
      int16_t* mv = reinterpret_cast<int16_t*>(out) + 2;
      for (int i = c.lane; i < A; i += 32) mv[i] = i < g.n_moves ? (int16_t)c.gm->moves[i] : (int16_t)-1;
      uint32_t* vis = reinterpret_cast<uint32_t*>(out + ((4 + (size_t)A * 2 + 3) & ~(size_t)3));
      const uint32_t* src = P.rec_visits + (size_t)m * A * A;
      const int n_words = g.n_moves * A;
      for (int i = c.lane; i < A * A; i += 32) vis[i] = i < n_words ? src[i] : 0u;
    }
    __syncwarp();
    const bool more = k + 1 < P.arena_matches_per_slot;
    const int pb = pg->arena_player_black ^ 1;  // colours swap (eval_main.py:316); both agents are reset (:333)
    // the mover's side (registers) ...
    g.rb = 0u; g.rw = 0u;
    g.n_moves = 0; g.last1 = -1; g.last2 = -1;
    g.root_node = CH_UNVISITED; g.root_n = 0u; g.root_w = 0.f; g.slot_count = 0u;
    g.sims_done = 0; g.sims_target = P.arena_num_mcts[side] + 1;
    g.status = more ? ST_SEARCH : ST_FINISHED;
    // ... and the other side (global); the decision streams of both agents run on (np.random is never re-seeded)
    if (c.lane < kRowsPad) {
      og->rows_b[c.lane] = 0;
      og->rows_w[c.lane] = 0;
    }
    if (c.lane == 0) {
      c.gm->is_real_root = 1;
      og->n_moves = 0; og->last1 = -1; og->last2 = -1;
      og->root_node = CH_UNVISITED; og->root_n = 0u; og->root_w = 0.f; og->slot_count = 0u;
      og->sims_done = 0; og->sims_target = P.arena_num_mcts[side ^ 1] + 1;
      og->is_real_root = 1;
      og->status = more ? ST_SEARCH : ST_FINISHED;
      pg->winner = win;  // of the last finished match
      pg->games_finished += 1ull;
      pg->arena_match = k + 1;
      pg->arena_player_black = pb;
      pg->arena_cur = pb ? 0 : 1;  // black moves first
    }
    __syncwarp();
    return false;  // the next launch picks the side that opens the next match
  }
  // ---- hand the position over to the other side: its get_pi(root_id) (agents.py:82-103)
  const int mcts_other = P.arena_num_mcts[side ^ 1];
  store_regs<1>(c, g);
  __syncwarp();
  c.gm = og;
  c.game = o;
  load_regs(c, g);
  c.abase = arena_base(P, o, g.arena);
  const int okind = P.arena_kind[side ^ 1];
  const bool in_tree = advance_root(c, g, action, okind == SIDE_ZERO);
  g.sims_done = 0;
  g.sims_target = in_tree ? mcts_other : mcts_other + 1;  // PUCT / UCT: always a fresh tree, num_mcts + 1 simulations
  g.status = ST_SEARCH;
  if (c.lane == 0) {
    og->is_real_root = in_tree ? 0 : 1;
    pg->arena_cur = side ^ 1;
  }
  if (in_tree && P.noise && g.root_node >= 0) remix_root_noise(c, g.root_node, P.A - g.n_moves, g.noise_draws);
  __syncwarp();
  return true;
}

// One warp advances one game (arena: one match) until it needs a network evaluation or finishes its search / game:
// consume the pending network output (expand + backup), play moves when searches complete, select the next leaf and
// emit its request.  Returns true when the game is still running (searching or waiting for the network).
template <int MAXJ>
__device__ __forceinline__ bool tree_step_game(const TreeParams& P, int game0, WarpSmem* ws, int lane_, int max_iters, bool nn_in_smem,
                               float nn_value_smem) {
  const bool arena = P.arena_M > 0;
  if (arena) game0 = P.games[game0].arena_cur ? P.arena_M + game0 : game0;  // unit = match; slot of the side to move
  Ctx c{P, &P.games[game0], ws, lane_, game0, 0, make_uint2(P.seed_lo, P.seed_hi)};
  Regs g;
  AO_TDBG_BEGIN
  load_regs(c, g);
  if (g.status != ST_SEARCH && g.status != ST_WAIT_NN) return false;
  c.abase = arena_base(P, c.game, g.arena);
  const int lane = c.lane;
  WarpSmem* sm = c.sm;
  const bool auto_play = c.gm->auto_play != 0;
  // the pending leaf (only the first iteration can find the game waiting for the network): loaded together with the
  // registers above - one round trip to memory instead of three dependent ones
  const int pre_slot = c.gm->nn_slot, pre_depth = c.gm->leaf_depth, pre_static = c.gm->nn_static;
  const uint32_t pre_ticket = c.gm->nn_ticket, pre_log = c.gm->nn_log_count;
  const uint16_t pre_rb = lane_ < kRowsPad ? c.gm->leaf_rows_b[lane_] : (uint16_t)0;
  const uint16_t pre_rw = lane_ < kRowsPad ? c.gm->leaf_rows_w[lane_] : (uint16_t)0;
  // c.gm / c.game change when an arena match hands over to the other side: always go through them
  auto path = [&]() { return P.path + (size_t)c.game * (P.A + 1); };

  for (int it = 0; it < max_iters; ++it) {
    if (g.status == ST_WAIT_NN) {
      // -------- consume the network output for the pending leaf - once the tower has served its ticket (a ragged last
      // wave of requests may have been left for the next round, NNQueue.defer)
      const bool pre = it == 0;
      if (!nn_in_smem && !(pre ? pre_static : c.gm->nn_static)) {
        const int qnet = arena ? arena_side(P, c.game) : 0;
        if ((int32_t)(P.nn_head[qnet] - (pre ? pre_ticket : c.gm->nn_ticket)) <= 0) break;  // not answered yet: keep waiting
      }
      const int slot = pre ? pre_slot : c.gm->nn_slot;
      const int depth = pre ? pre_depth : c.gm->leaf_depth;
      // the network's answer: from the exchange buffers in HBM, or already in sm->pol (persistent self-play kernel)
      if (!nn_in_smem)
        for (int a = lane; a < P.A; a += 32) sm->pol[a] = P.nn_policy[(size_t)slot * P.A + a];
      const float value = nn_in_smem ? nn_value_smem : P.nn_value[slot];
      sm->rows[0][lane] = pre ? pre_rb : (lane < kRowsPad ? c.gm->leaf_rows_b[lane] : (uint16_t)0);
      sm->rows[1][lane] = pre ? pre_rw : (lane < kRowsPad ? c.gm->leaf_rows_w[lane] : (uint16_t)0);
      __syncwarp();
      if (P.nn_log_cap > 0) {
        const uint32_t k = pre ? pre_log : c.gm->nn_log_count;
        if (k < (uint32_t)P.nn_log_cap) {
          float* lp = P.nnlog_policy + ((size_t)c.game * P.nn_log_cap + k) * P.A;
          for (int a = lane; a < P.A; a += 32) lp[a] = sm->pol[a];
          if (lane == 0) P.nnlog_value[(size_t)c.game * P.nn_log_cap + k] = value;
        }
        __syncwarp();
        if (lane == 0) c.gm->nn_log_count = k + 1u;
      }
      g.status = ST_SEARCH;
      AO_TDBG(0)
      if (!expand_and_backup(c, g, depth, value, false)) {
        g.status = ST_ERROR;
        if (lane == 0) c.gm->error = 1;
        break;
      }
      if (lane == 0) c.gm->sims_total += 1ull;
      AO_TDBG(1)
      continue;
    }
    if (arena) {
      const int kind = P.arena_kind[arena_side(P, c.game)];
      if (kind == SIDE_PUCT || kind == SIDE_UCT) {
        // a play-out agent's turn: a slice of its search per lock-step round, so that the network sides of the other
        // matches are not held up by a whole search
        if (g.sims_done == 0) rollout_fresh_tree(g);
        bool ok = true;
        int ran = 0;
        for (; ok && ran < P.rollout_sims_per_round && g.sims_done < g.sims_target; ++ran)
          ok = rollout_sim<MAXJ>(c, g, kind);
        if (lane == 0) c.gm->sims_total += (unsigned long long)ran;
        if (!ok) {
          g.status = ST_ERROR;
          if (lane == 0) c.gm->error = 1;
          break;
        }
        if (g.sims_done < g.sims_target) break;  // to be continued next round
      }
      if (kind != SIDE_ZERO || g.sims_done >= g.sims_target) {
        if (!P.allow_moves) break;  // moves are played in the move rounds (TreeParams.allow_moves)
        if (!arena_move(c, g)) break;
        continue;
      }
    }
    if (g.sims_done >= g.sims_target) {
      if (!auto_play) {
        g.status = ST_SEARCH_DONE;
        break;
      }
      if (!P.allow_moves) break;  // moves are played in the move rounds (TreeParams.allow_moves)
      play_move<MAXJ>(c, g);
      if (g.status != ST_SEARCH) break;
      continue;
    }
    // -------- selection (agents.py:134-168)
    uint32_t rb = g.rb, rw = g.rw;
    int nm = g.n_moves;
    int depth = 0;
    int win = 0;
    bool need_eval = true;
    int walk_a1 = -1, walk_a2 = -1;  // actions of the last two levels of the walk (= slot_act[path[depth-1]], [depth-2])
    if (g.root_n > 0u) {
      int32_t node = g.root_node;
      if (node <= -2) {
        win = -(node + 1);
        need_eval = false;
      }
      while (node >= 0) {
        const int L = P.A - nm;
        const size_t base = c.abase + (size_t)node;
        uint32_t n_[MAXJ];
        float w_[MAXJ];
        double p_[MAXJ];
        int32_t ch_[MAXJ];
        uint32_t act_[MAXJ];
        uint32_t tot = 0;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          const int i = j * 32 + lane;
          n_[j] = 0u; w_[j] = 0.f; p_[j] = 0.0; ch_[j] = CH_UNVISITED; act_[j] = 0u;
          if (i < L) {
            const uint2 nw = P.slot_nw[base + i];
            n_[j] = nw.x;
            w_[j] = __uint_as_float(nw.y);
            p_[j] = P.slot_p[base + i];
            ch_[j] = P.slot_child[base + i];
            act_[j] = P.slot_act[base + i];
            tot += nw.x;
          }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
        const double sq = __dsqrt_rn((double)tot);
        double val[MAXJ];
        double best = -1.0e300;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          const int i = j * 32 + lane;
          val[j] = -1.0e300;
          if (i < L) {
            const float q = n_[j] > 0u ? __fdiv_rn(w_[j], (float)n_[j]) : 0.0f;
            const double u = __ddiv_rn(__dmul_rn(__dmul_rn(P.c_puct, p_[j]), sq), (double)(n_[j] + 1u));
            val[j] = __dadd_rn((double)q, u);
            best = fmax(best, val[j]);
          }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) best = fmax(best, __shfl_xor_sync(kFull, best, o));
        unsigned tie[MAXJ];
        int K = 0;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          tie[j] = __ballot_sync(kFull, (j * 32 + lane) < L && val[j] == best);
          K += __popc(tie[j]);
        }
        int r = draw_choice(c, g.rng_ctr, K);
        uint32_t s_n = 0, s_act = 0;
        int32_t s_ch = CH_UNVISITED;
        int s_idx = 0;
        bool done = false;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          if (!done) {
            const int cnt = __popc(tie[j]);
            if (r < cnt) {
              const int src = (int)__fns(tie[j], 0, r + 1);
              s_n = __shfl_sync(kFull, n_[j], src);
              s_act = __shfl_sync(kFull, act_[j], src);
              s_ch = __shfl_sync(kFull, ch_[j], src);
              s_idx = j * 32 + src;
              done = true;
            } else {
              r -= cnt;
            }
          }
        }
        if (lane == 0) path()[depth] = (uint32_t)(base + s_idx);
        ++depth;
        walk_a2 = walk_a1;
        walk_a1 = (int)s_act;
        const int y = (int)s_act / P.B, x = (int)s_act % P.B;
        if (lane == y) {
          if ((nm & 1) == 0) rb |= 1u << x;
          else rw |= 1u << x;
        }
        ++nm;
        if (s_n == 0u) break;  // unvisited child: this is the leaf
        node = s_ch;
        if (node <= -2) {
          win = -(node + 1);
          need_eval = false;
        }
      }
    }
    AO_TDBG(6)
    if (need_eval) win = check_win_rows(rb, rw, P.B, nm, sm->rows, lane);
    __syncwarp();
    AO_TDBG(7)
    if (win != 0) {
      // terminal leaf: mark, backup reward (no network call; the reference's call is discarded, agents.py:171-178)
      if (lane == 0) {
        if (depth > 0) P.slot_child[path()[depth - 1]] = -(1 + win);
        c.gm->terminal_sims += 1ull;
        c.gm->sims_total += 1ull;
      }
      if (depth == 0) g.root_node = -(1 + win);
      __syncwarp();
      expand_and_backup(c, g, depth, 0.f, true);
      AO_TDBG(8)
      continue;
    }
    // -------- non-terminal leaf: evaluate
    sm->rows[0][lane] = (uint16_t)rb;
    sm->rows[1][lane] = (uint16_t)rw;
    __syncwarp();
    if (P.eval_mode == AO_EVAL_SYNTH) {
      float value;
      synth_eval(c, c.gm->moves, g.n_moves, depth, value, arena ? P.synth_salt[arena_side(P, c.game)] : P.synth_salt[0]);
      if (lane == 0) {
        c.gm->nn_evals += 1ull;
        c.gm->sims_total += 1ull;
      }
      if (!expand_and_backup(c, g, depth, value, false)) {
        g.status = ST_ERROR;
        if (lane == 0) c.gm->error = 1;
        break;
      }
      continue;
    }
    // network request: the five planes of utils.get_state_pt (utils.py:139-168) as row masks
    {
      // arena: requests of side s are evaluated with weight set s and live in nn slots [s * M, ...)
      const int net = arena ? arena_side(P, c.game) : 0;
      int slot = c.game;  // static_slots: request slot = game slot (the persistent kernel's fixed game -> CTA pass map)
      uint32_t ticket = 0u;
      if (!P.static_slots) {  // next ticket of the weight set's queue; its ring slot
        if (lane == 0) ticket = atomicAdd(P.nn_count + net, 1u);
        ticket = __shfl_sync(kFull, ticket, 0);
        slot = (int)((ticket & P.nn_ring_mask) + (uint32_t)net * (P.nn_ring_mask + 1u));
      }
      // last two actions on the path to the leaf (or of the root position)
      int l1 = g.last1, l2 = g.last2;
      if (depth >= 1) {  // the walk above selected them: no need to chase path[] -> slot_act[] through memory again
        l2 = l1;
        l1 = walk_a1;
        if (depth >= 2) l2 = walk_a2;
      }
      const bool black_to_move = (nm & 1) == 0;
      uint32_t own = black_to_move ? rb : rw, opp = black_to_move ? rw : rb;
      uint32_t opp_prev = opp, own_prev = own;
      if (l1 >= 0 && lane == l1 / P.B) opp_prev &= ~(1u << (l1 % P.B));
      if (l2 >= 0 && lane == l2 / P.B) own_prev &= ~(1u << (l2 % P.B));
      LeafIn* in = &P.nn_in[slot];
      if (lane < kRowsPad) {
        in->plane[0][lane] = (uint16_t)own_prev;
        in->plane[1][lane] = (uint16_t)opp_prev;
        in->plane[2][lane] = (uint16_t)own;
        in->plane[3][lane] = (uint16_t)opp;
        c.gm->leaf_rows_b[lane] = (uint16_t)rb;
        c.gm->leaf_rows_w[lane] = (uint16_t)rw;
      }
      if (lane == 0) {
        in->colour = black_to_move ? 1u : 0u;
        in->game = c.game;
        c.gm->leaf_depth = depth;
        c.gm->leaf_n_moves = nm;
        c.gm->nn_slot = slot;
        c.gm->nn_ticket = ticket;
        c.gm->nn_static = P.static_slots ? 1 : 0;
        c.gm->nn_evals += 1ull;
      }
      g.status = ST_WAIT_NN;
      AO_TDBG(9)
      break;
    }
  }
  __syncwarp();
  store_regs<MAXJ>(c, g);
  AO_TDBG(10)
  return g.status == ST_SEARCH || g.status == ST_WAIT_NN;
}

}  // namespace
}  // namespace ao
