// PUCTAgent / UCTAgent (agents.py:263-634): pure-MCTS agents with uniformly random play-outs, the non-network sides
// eval_main.py:67-84 offers by name ('puct' / 'uct').  Included by tree.cu (inside its anonymous namespace, after the
// decision-stream and tree helpers): one warp runs the whole search of one position on the slot's tree storage.
//
// Reference semantics kept bit for bit: every get_pi starts a fresh tree (`_init_mcts` overwrites the root with n = 0,
// so the first simulation re-creates all children, agents.py:300-311 / 482-496), num_mcts + 1 simulations, children in
// ascending cell order (utils.valid_actions), all statistics Python floats = float64 (w is a sum of +-1 / 0 rewards,
// i.e. an integer, stored exactly in the float32 slot; q = w / n is formed in float64 when needed),
//   PUCT: u = ((c_puct * p) * sqrt(sum n)) / (n + 1), p = 1 / #children          (agents.py:354-363)
//   UCT : u = inf for n == 0, else sqrt((2 * log(sum n)) / n)                    (agents.py:551-559)
// ties broken by one decision-stream draw over the candidates in child order, a non-root leaf is evaluated by one
// uniformly random play-out (one draw per move, none when a single cell is left), reward from utils.get_reward for the
// side that moved into the leaf, +1 for a terminal leaf, 0 for the root; backup alternates the sign up to the root.
// log(k) comes from a table the host fills with numpy's own values (ao_set_log_table), so no libm difference can move
// an arg-max.
#pragma once

enum : int { SIDE_ZERO = 0, SIDE_RANDOM = 1, SIDE_PUCT = 2, SIDE_UCT = 3 };

// r-th (0-based) empty cell in ascending order; occ = this lane's row occupancy. Warp-uniform result.
__device__ __forceinline__ int nth_empty_cell(uint32_t occ, int B, int r, int lane) {
  const uint32_t rowmask = (1u << B) - 1u;
  const uint32_t free_bits = lane < B ? (~occ & rowmask) : 0u;
  const int e = __popc(free_bits);
  int incl = e;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  const int excl = incl - e;
  const bool mine = r >= excl && r < incl;
  const unsigned bal = __ballot_sync(kFull, mine);
  const int src = __ffs(bal) - 1;
  int cell = 0;
  if (mine) cell = lane * B + (int)__fns(free_bits, 0, r - excl + 1);
  return __shfl_sync(kFull, cell, src < 0 ? 0 : src);
}

// start a fresh search tree at the slot's current root position (`_init_mcts`)
__device__ __forceinline__ void rollout_fresh_tree(Regs& g) {
  g.root_node = CH_UNVISITED;
  g.root_n = 0u;
  g.root_w = 0.f;
  g.slot_count = 0u;
}

// One simulation (selection, expansion + play-out, backup). Returns false on tree-arena overflow.
template <int MAXJ>
__device__ bool rollout_sim(Ctx& c, Regs& g, int kind) {
  const TreeParams& P = c.P;
  WarpSmem* sm = c.sm;
  const int lane = c.lane, A = P.A, B = P.B;
  uint32_t* path = P.path + (size_t)c.game * (A + 1);
  uint32_t rb = g.rb, rw = g.rw;
  int nm = g.n_moves, depth = 0, win = 0;
  bool leaf_is_root = true;
  if (g.root_n > 0u) {
    int32_t node = g.root_node;
    while (true) {  // a node with n > 0: terminal, or expanded
      win = check_win_rows(rb, rw, B, nm, sm->rows, lane);
      if (win != 0 || node < 0) break;
      const int L = A - nm;
      const size_t base = c.abase + (size_t)node;
      uint32_t n_[MAXJ], act_[MAXJ];
      float w_[MAXJ];
      int32_t ch_[MAXJ];
      uint32_t tot = 0;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int i = j * 32 + lane;
        n_[j] = 0u; w_[j] = 0.f; ch_[j] = CH_UNVISITED; act_[j] = 0u;
        if (i < L) {
          const uint2 nw = P.slot_nw[base + i];
          n_[j] = nw.x;
          w_[j] = __uint_as_float(nw.y);
          ch_[j] = P.slot_child[base + i];
          act_[j] = P.slot_act[base + i];
          tot += nw.x;
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
      const double sq = __dsqrt_rn((double)tot);
      const double cp = __dmul_rn(P.c_puct, __ddiv_rn(1.0, (double)L));  // c_puct * p, p = 1 / len(actions)
      const double two_log = kind == SIDE_UCT && tot > 0u ? __dmul_rn(2.0, P.log_table[tot < (uint32_t)P.log_table_n ? tot : 0u]) : 0.0;
      const double inf = __longlong_as_double(0x7FF0000000000000ll);
      double val[MAXJ];
      double best = -inf;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int i = j * 32 + lane;
        val[j] = -inf;
        if (i < L) {
          const double q = n_[j] > 0u ? __ddiv_rn((double)w_[j], (double)n_[j]) : 0.0;
          double u;
          if (kind == SIDE_PUCT) u = __ddiv_rn(__dmul_rn(cp, sq), (double)(n_[j] + 1u));
          else u = n_[j] == 0u ? inf : __dsqrt_rn(__ddiv_rn(two_log, (double)n_[j]));
          val[j] = __dadd_rn(q, u);
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
      if (lane == 0) path[depth] = (uint32_t)(base + s_idx);
      ++depth;
      const int y = (int)s_act / B, x = (int)s_act % B;
      if (lane == y) {
        if ((nm & 1) == 0) rb |= 1u << x;
        else rw |= 1u << x;
      }
      ++nm;
      leaf_is_root = false;
      if (s_n == 0u) {
        win = check_win_rows(rb, rw, B, nm, sm->rows, lane);
        break;
      }
      node = s_ch;
    }
  }
  __syncwarp();
  float reward;
  if (win == 0) {
    // expansion: one child per empty cell, ascending (utils.valid_actions)
    const int L = A - nm;
    if (g.slot_count + (uint32_t)L > P.slot_cap) return false;
    const uint32_t off = g.slot_count;
    g.slot_count += (uint32_t)L;
    const size_t base = c.abase + off;
    const uint32_t occ = rb | rw;
    {  // lane y enumerates the empty cells of row y; an exclusive scan of the per-row counts gives their child indices
      const uint32_t rowmask = (1u << B) - 1u;
      const uint32_t free_bits = lane < B ? (~occ & rowmask) : 0u;
      const int e = __popc(free_bits);
      int incl = e;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      int idx = incl - e;  // child index of this row's first empty cell
      uint32_t fb = free_bits;
      while (fb) {
        const int x = __ffs(fb) - 1;
        fb &= fb - 1u;
        P.slot_act[base + idx] = (uint8_t)(lane * B + x);
        P.slot_nw[base + idx] = make_uint2(0u, 0u);
        P.slot_child[base + idx] = CH_UNVISITED;
        ++idx;
      }
    }
    if (leaf_is_root) g.root_node = (int32_t)off;
    else if (lane == 0) P.slot_child[path[depth - 1]] = (int32_t)off;
    if (leaf_is_root) {
      reward = 0.f;  // "root node don't simulation"
    } else {
      // uniformly random play-out from the leaf (agents.py:391-411)
      uint32_t sb = rb, sw = rw;
      int sn = nm, turn = nm & 1, w = 0;  // turn 0: black places
      while (true) {
        const int r = draw_choice(c, g.rng_ctr, A - sn);
        const int cell = nth_empty_cell(sb | sw, B, r, lane);
        if (lane == cell / B) {
          if (turn == 0) sb |= 1u << (cell % B);
          else sw |= 1u << (cell % B);
        }
        ++sn;
        w = check_win_rows(sb, sw, B, sn, sm->rows, lane);
        if (w != 0) break;
        turn ^= 1;
      }
      // utils.get_reward(win, leaf_id): turn = get_turn(leaf_id) = 0 iff black is to move at the leaf (nm even)
      const int leaf_turn = nm & 1;
      reward = w == 1 ? (leaf_turn == 1 ? 1.f : -1.f) : w == 2 ? (leaf_turn == 1 ? -1.f : 1.f) : 0.f;
    }
  } else {
    reward = 1.f;  // "terminal node don't expansion"
  }
  // backup: n += 1, w += reward * (-1)^count from the leaf up to the root (agents.py:421-432)
  if (lane == 0) {
    float d = reward;
    for (int i = depth - 1; i >= 0; --i) {
      uint2 nw = P.slot_nw[path[i]];
      nw.x += 1u;
      nw.y = __float_as_uint(__fadd_rn(__uint_as_float(nw.y), d));
      P.slot_nw[path[i]] = nw;
      d = -d;
    }
    g.root_w = __fadd_rn(g.root_w, d);
  }
  g.root_w = __shfl_sync(kFull, g.root_w, 0);
  g.root_n += 1u;
  g.sims_done += 1;
  __syncwarp();
  return true;
}

// root children -> sm->dbuf[a] = score the agent's get_pi maximises (PUCT: visits, agents.py:288-296; UCT: q with -inf
// for cells that are no children and 0 for unvisited children, agents.py:466-476); rec (optional) receives the visits
__device__ void rollout_scores(Ctx& c, const Regs& g, int kind, uint32_t* rec) {
  const TreeParams& P = c.P;
  const int A = P.A;
  const double ninf = __longlong_as_double(0xFFF0000000000000ll);
  for (int a = c.lane; a < A; a += 32) c.sm->dbuf[a] = kind == SIDE_UCT ? ninf : 0.0;
  __syncwarp();
  if (g.root_node >= 0) {
    const size_t base = c.abase + (size_t)g.root_node;
    const int L = A - g.n_moves;
    for (int i = c.lane; i < L; i += 32) {
      const int a = P.slot_act[base + i];
      const uint2 nw = P.slot_nw[base + i];
      if (rec) rec[a] = nw.x;
      c.sm->dbuf[a] = kind == SIDE_UCT ? (nw.x > 0u ? __ddiv_rn((double)__uint_as_float(nw.y), (double)nw.x) : 0.0)
                                       : (double)nw.x;
    }
  }
  __syncwarp();
}
