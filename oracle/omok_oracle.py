"""CPU ORACLE (test infrastructure, NOT product code) for the alpha_omok self-play hot path.

A from-scratch numpy/pure-Python restatement of the reference algorithm. Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this file; the product path (alpha_omok_b200/) never does.

PARITY PINNING: the reference has no tests or golden vectors of its own. This oracle is pinned
against the UNMODIFIED reference code imported from /root/reference/2_AlphaOmok in the build container:
`tests/golden/make_golden.py` runs reference agents.ZeroAgent / utils / env_small / model.PVNet with the random
decisions routed through the same counter-based decision stream (class DecisionStream below) and commits the resulting
visit-count vectors, winners and NN outputs under tests/golden/; `tests/test_oracle_golden.py` replays them here.

Reference citations (relative to /root/reference/2_AlphaOmok):
  agents.py:39-260   ZeroAgent (get_pi, _init_mcts, _mcts, _selection, _expansion_evaluation, _backup)
  utils.py:22-27     legal_actions (CPython set order)        utils.py:30-59    check_win
  utils.py:139-168   get_state_pt                             utils.py:171-179  get_board
  utils.py:189-205   get_action / argmax_onehot               model.py:13-104   PVNet
  env/env_small.py:106-199  GameState.step                    main.py:132-250   self_play loop

Arithmetic contract (numpy >= 2, NEP 50): n integer-valued, w/q float32, p float64,
u = ((5*p)*sqrt(sum_child_n))/(n+1) in float64, q+u and the max/== in float64, prior renormalisation by numpy's
pairwise float64 sum over the full A-vector.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------------------------
# Decision stream: Philox4x32-10, identical on host (here) and device (philox4x32 in csrc/rules.cuh, used by csrc/tree.cu).
# counter = (index, 0, game, stream) ; key = (seed_lo, seed_hi)
# --------------------------------------------------------------------------------------------------------------
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32(counter, key):
    c0, c1, c2, c3 = (int(x) & _MASK for x in counter)
    k0, k1 = (int(x) & _MASK for x in key)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


class DecisionStream:
    """Per-game stream of random decisions.

    choice(k)   : k == 1 -> 0 without consuming (numpy does not consume either); else one Philox block, (x0*k)>>32.
    choice_p(p) : numpy legacy RandomState.choice(p=...) restated: cdf = cumsum(p)/cdf[-1], one 53-bit uniform,
                  searchsorted(side='right').
    gamma_row(L): raw Gamma(alpha,1) variates for the next Dirichlet draw, read from a host-generated tape
                  (parity mode: the device reads the same tape). Dirichlet = g[:L] * (1/sequential_sum(g[:L])).
    """

    def __init__(self, seed: int, game: int, gamma_tape: np.ndarray | None = None):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.game = game
        self.ctr = 0
        self.noise_draws = 0
        self.gamma_tape = gamma_tape  # float64 [n_draws, A]

    def _block(self):
        out = philox4x32((self.ctr, 0, self.game, 0), self.key)
        self.ctr += 1
        return out

    def choice(self, k: int) -> int:
        if k <= 1:
            return 0
        return (self._block()[0] * k) >> 32

    def uniform53(self) -> float:
        b = self._block()
        return ((b[0] >> 5) * 67108864.0 + (b[1] >> 6)) / 9007199254740992.0

    def choice_p(self, p: np.ndarray) -> int:
        cdf = np.cumsum(np.asarray(p, np.float64))
        cdf = cdf / cdf[-1]
        u = self.uniform53()
        return int(np.searchsorted(cdf, u, side="right"))

    def dirichlet(self, L: int) -> np.ndarray:
        if L == 0:
            return np.zeros(0)
        if self.gamma_tape is None:
            raise RuntimeError("oracle DecisionStream needs a gamma tape for Dirichlet draws")
        g = self.gamma_tape[self.noise_draws, :L]
        self.noise_draws += 1
        acc = 0.0
        for x in g:  # sequential float64 sum (numpy legacy dirichlet does the same)
            acc = acc + float(x)
        inv = 1.0 / acc
        return np.asarray([float(x) * inv for x in g], np.float64)


def make_gamma_tape(seed: int, game: int, n_draws: int, A: int, alpha: float) -> np.ndarray:
    """Host-side Gamma(alpha,1) tape for one game (numpy Generator, PCG64; shipped to the device in parity mode)."""
    rng = np.random.Generator(np.random.PCG64([seed, game, 0x0A0C]))
    g = rng.standard_gamma(alpha, size=(n_draws, A))
    return np.maximum(g, np.finfo(np.float64).tiny)


# --------------------------------------------------------------------------------------------------------------
# Rules / encodings (utils.py)
# --------------------------------------------------------------------------------------------------------------
def get_turn(node_id) -> int:
    """utils.py:182-186 - 0 = black to move, 1 = white to move."""
    return 0 if len(node_id) % 2 == 1 else 1


def get_board(node_id, board_size: int) -> np.ndarray:
    """utils.py:171-179 - ID -> board (+1 black = odd ply, -1 white)."""
    b = np.zeros(board_size * board_size)
    for ply, a in enumerate(node_id[1:]):
        b[a] = 1.0 if ply % 2 == 0 else -1.0
    return b.reshape(board_size, board_size)


def check_win(board: np.ndarray, win_mark: int) -> int:
    """utils.py:30-59 - windows scanned row-major; inside a window black (rows/cols, diagonals) before white.
    0 playing, 1 black, 2 white, 3 draw (full board, no five). Overlines count."""
    B = board.shape[0]
    bl = (board > 0).astype(np.int64)
    wh = (board < 0).astype(np.int64)
    for r in range(B - win_mark + 1):
        for c in range(B - win_mark + 1):
            for colour, res in ((bl, 1), (wh, 2)):
                g = colour[r:r + win_mark, c:c + win_mark]
                if (g.sum(axis=1) == win_mark).any() or (g.sum(axis=0) == win_mark).any():
                    return res
                if np.trace(g) == win_mark or np.trace(g[::-1]) == win_mark:
                    return res
    if np.count_nonzero(board) == B * B:
        return 3
    return 0


def get_state_pt(node_id, board_size: int, channel_size: int) -> np.ndarray:
    """utils.py:139-168 - [C,B,B] float64. A length-C FIFO starts as C zero planes; ply 0 pushes two zero planes, every
    later ply pushes the cumulative stone plane of the colour that just moved; finally the colour plane is pushed
    (ones iff black is to move). The state is the last C pushes."""
    B = board_size
    fifo = [np.zeros((B, B)) for _ in range(channel_size)]
    cum = [np.zeros((B, B)), np.zeros((B, B))]  # black, white
    colour = 1.0
    for ply, a in enumerate(node_id):
        if ply == 0:
            fifo += [cum[0].copy(), cum[1].copy()]
            continue
        c = 0 if ply % 2 == 1 else 1
        cum[c][a // B, a % B] = 1.0
        fifo.append(cum[c].copy())
        colour = 0.0 if c == 0 else 1.0
    fifo.append(np.full((B, B), colour))
    return np.stack(fifo[-channel_size:])


# --- CPython 3.12 set iteration order, restated (Objects/setobject.c: set_add_entry / set_insert_clean /
#     set_table_resize / set_difference).  SURVEY.md appendix A.3.
_LINEAR_PROBES = 9
_PERTURB_SHIFT = 5


def _set_insert_clean(table, mask, key):
    perturb = key
    i = key & mask
    while True:
        probes = _LINEAR_PROBES if i + _LINEAR_PROBES <= mask else 0
        j = i
        while True:
            if table[j] is None:
                table[j] = key
                return
            if probes == 0:
                break
            probes -= 1
            j += 1
        perturb >>= _PERTURB_SHIFT
        i = (i * 5 + 1 + perturb) & mask


def cpython_set_difference_order(A: int, occupied) -> list:
    """Iteration order of `set(range(A)) - set(occupied)` on CPython 3.12 (utils.py:22-27)."""
    occ = set(occupied)
    if (A >> 2) > len(occ):
        # set_copy_and_difference: copy of the big table (slot == key) then discards -> ascending
        return [a for a in range(A) if a not in occ]
    mask = 7
    table = [None] * 8
    fill = 0
    for key in range(A):  # `so` iterates in slot order == ascending for the comprehension-built {0..A-1}
        if key in occ:
            continue
        # set_add_entry on a table without dummies and without equal keys == set_insert_clean probing
        _set_insert_clean(table, mask, key)
        fill += 1
        if fill * 5 >= mask * 3:
            minused = fill * 4 if fill <= 50000 else fill * 2
            newsize = 8
            while newsize <= minused:
                newsize <<= 1
            old = table
            table = [None] * newsize
            mask = newsize - 1
            for k in old:
                if k is not None:
                    _set_insert_clean(table, mask, k)
    return [k for k in table if k is not None]


def legal_actions(node_id, board_size: int) -> list:
    """utils.py:22-27 - empty cells in CPython-set iteration order (defines the child order)."""
    return cpython_set_difference_order(board_size * board_size, node_id[1:])


def np_pairwise_sum(a: np.ndarray) -> float:
    """numpy's float64 pairwise summation (numpy/_core/src/umath/loops_utils.h.src, PW_BLOCKSIZE 128), restated;
    this is what `prior_prob.sum()` (agents.py:189) executes. The device kernel implements exactly this order."""
    a = np.asarray(a, np.float64)
    n = a.shape[0]
    if n < 8:
        res = 0.0
        for x in a:
            res = res + float(x)
        return res
    if n <= 128:
        r = [float(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + float(a[i + j])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return np_pairwise_sum(a[:n2]) + np_pairwise_sum(a[n2:])


def argmax_onehot(pi: np.ndarray, stream: DecisionStream):
    """utils.py:198-205 - argmax over ascending indices with uniform tie-break."""
    idx = np.flatnonzero(pi == pi.max())
    a = int(idx[stream.choice(len(idx))])
    onehot = np.zeros(len(pi))
    onehot[a] = 1.0
    return onehot, a


def get_action(pi: np.ndarray, stream: DecisionStream):
    """utils.py:189-195 - sample an action from pi (one uniform consumed even for a one-hot pi)."""
    a = stream.choice_p(pi)
    onehot = np.zeros(len(pi))
    onehot[a] = 1.0
    return onehot, a


# --------------------------------------------------------------------------------------------------------------
# env.GameState.step (env/env_small.py:106-199, env_regular.py identical up to the size constant)
# --------------------------------------------------------------------------------------------------------------
class OracleGameState:
    def __init__(self, board_size: int, win_stones: int = 5):
        self.B = board_size
        self.win = win_stones
        self.init = False
        self.num_stones = 0
        self.gameboard = np.zeros((board_size, board_size))
        self.turn = 0
        self.black_win = self.white_win = self.count_draw = 0

    def step(self, onehot):
        if self.init:  # lazy reset after a finished game (env_small.py:108-117)
            self.num_stones = 0
            self.gameboard = np.zeros((self.B, self.B))
            self.turn = 0
            self.init = False
        valid = False
        a = 0
        if np.any(onehot):
            a = int(np.argmax(onehot))
            y, x = a // self.B, a % self.B
            valid = self.gameboard[y, x] == 0
            self.gameboard[y, x] = 1.0 if self.turn == 0 else -1.0  # occupied cells are overwritten (:161-176)
            self.turn ^= 1
            self.num_stones += 1
        w = check_win(self.gameboard, self.win)
        if w == 1:
            self.black_win += 1
        elif w == 2:
            self.white_win += 1
        elif w == 3:
            self.count_draw += 1
        self.init = w != 0
        return self.gameboard, bool(valid), w, self.turn, a


# --------------------------------------------------------------------------------------------------------------
# ZeroAgent MCTS (agents.py:39-260), restated over a node pool with per-parent child slots
# --------------------------------------------------------------------------------------------------------------
class _Node:
    __slots__ = ("acts", "n", "w", "p", "child")

    def __init__(self, acts, priors):
        self.acts = list(acts)
        self.n = [0] * len(acts)
        self.w = [np.float32(0.0)] * len(acts)
        self.p = [float(x) for x in priors]
        self.child = [-1] * len(acts)  # -1 unvisited, >=0 expanded node index, -(1+win_index) terminal


class OracleZeroAgent:
    """evaluate(moves_tuple) -> (policy float32[A], value float32) is the NN (agents.py:175-178)."""

    def __init__(self, board_size, num_mcts, evaluate, stream: DecisionStream, noise=True):
        self.B = board_size
        self.A = board_size * board_size
        self.num_mcts = num_mcts
        self.win_mark = 3 if board_size == 3 else 5
        self.alpha = 10 / self.A
        self.c_puct = 5
        self.noise = noise
        self.evaluate = evaluate
        self.stream = stream
        self.visit = np.zeros(self.A)
        self.policy = np.zeros(self.A)
        self.sims = 0
        self.terminal_sims = 0
        self.reset()

    def reset(self):  # agents.py:55-58
        self.nodes = []
        self.root_moves = None
        self.root_node = -1
        self.root_n = 0
        self.root_w = np.float32(0.0)
        self.is_real_root = True

    # ---- root handling == `root_id in self.tree` (agents.py:82-103)
    def _set_root(self, root_id):
        moves = tuple(root_id)
        in_tree = False
        if self.root_moves is not None and moves[:len(self.root_moves)] == self.root_moves:
            in_tree = True
            node, rn, rw = self.root_node, self.root_n, self.root_w
            for a in moves[len(self.root_moves):]:
                if node < 0:
                    in_tree = False
                    break
                nd = self.nodes[node]
                i = nd.acts.index(a)
                rn, rw, node = nd.n[i], nd.w[i], nd.child[i]
            if in_tree:
                self.root_node, self.root_n, self.root_w = node, rn, rw
        if not in_tree:
            self.nodes = []
            self.root_node, self.root_n, self.root_w = -1, 0, np.float32(0.0)
        self.root_moves = moves
        self.is_real_root = not in_tree
        if in_tree and self.noise and self.root_node >= 0:  # re-mix noise into the existing priors (:95-103)
            nd = self.nodes[self.root_node]
            eta = self.stream.dirichlet(len(nd.acts))
            for i in range(len(nd.acts)):
                nd.p[i] = 0.75 * nd.p[i] + 0.25 * float(eta[i])

    def _simulate(self):
        moves = list(self.root_moves)
        path = []
        node = self.root_node
        win = 0
        if self.root_n > 0:
            if node <= -2:  # terminal root: `_selection` returns it immediately
                win = -(node + 1)
            while node >= 0:
                nd = self.nodes[node]
                total_n = float(sum(nd.n))
                sq = np.sqrt(total_n)
                best, ties = None, []
                for i in range(len(nd.acts)):
                    q = (nd.w[i] / np.float32(nd.n[i])) if nd.n[i] > 0 else 0.0
                    u = self.c_puct * np.float64(nd.p[i]) * sq / (nd.n[i] + 1)
                    v = np.float64(q) + u
                    if best is None or v > best:
                        best, ties = v, [i]
                    elif v == best:
                        ties.append(i)
                i = ties[self.stream.choice(len(ties))]
                path.append((node, i))
                moves.append(nd.acts[i])
                if nd.n[i] == 0:
                    win = check_win(get_board(tuple(moves), self.B), self.win_mark)
                    node = -1
                    break
                node = nd.child[i]
                if node <= -2:
                    win = -(node + 1)
        else:
            win = check_win(get_board(tuple(moves), self.B), self.win_mark)
        leaf_moves = tuple(moves)
        # ---- expansion + evaluation (agents.py:170-221); the reference also evaluates terminal leaves and discards
        if win == 0:
            policy, value = self.evaluate(leaf_moves)
            policy = np.asarray(policy, np.float32)
            value = np.float32(value)
            acts = legal_actions(leaf_moves, self.B)
            prior = np.zeros(self.A)
            for a in acts:
                prior[a] = policy[a]
            prior /= prior.sum()
            pri = [prior[a] for a in acts]
            is_root = len(path) == 0
            if self.noise and is_root:
                eta = self.stream.dirichlet(len(acts))
                pri = [0.75 * pri[i] + 0.25 * float(eta[i]) for i in range(len(acts))]
            self.nodes.append(_Node(acts, pri))
            new_idx = len(self.nodes) - 1
            if is_root:
                self.root_node = new_idx
            else:
                pn, pi_ = path[-1]
                self.nodes[pn].child[pi_] = new_idx
            delta = -value  # float32
        else:
            self.terminal_sims += 1
            if path:
                pn, pi_ = path[-1]
                self.nodes[pn].child[pi_] = -(1 + win)
            else:
                self.root_node = -(1 + win)
            delta = np.float32(1.0)
        # ---- backup (agents.py:223-239): leaf gets +delta, alternating sign up to and including the root
        sign = 1
        for pn, pi_ in reversed(path):
            nd = self.nodes[pn]
            nd.n[pi_] += 1
            nd.w[pi_] = np.float32(nd.w[pi_] + (delta if sign > 0 else -delta))
            sign = -sign
        self.root_n += 1
        self.root_w = np.float32(self.root_w + (delta if sign > 0 else -delta))
        self.sims += 1

    @property
    def root_id(self):  # agents.py:83 (`self.root_id = root_id`)
        return self.root_moves

    def get_pi(self, root_id, tau):  # agents.py:60-80
        self._set_root(root_id)
        n_sims = self.num_mcts + 1 if self.is_real_root else self.num_mcts
        for _ in range(n_sims):
            self._simulate()
        visit = np.zeros(self.A)
        policy = np.zeros(self.A)
        if self.root_node >= 0:
            nd = self.nodes[self.root_node]
            for i, a in enumerate(nd.acts):
                visit[a] = nd.n[i]
                policy[a] = nd.p[i]
        self.visit, self.policy = visit, policy
        pi = visit / visit.sum()
        if tau == 0:
            # `host_stream` (optional): where draws made OUTSIDE `_mcts` come from - tests of the Python facades set it,
            # because there the search draws happen on the device and this one in numpy's global generator
            pi, _ = argmax_onehot(pi, getattr(self, "host_stream", None) or self.stream)
        return pi


def self_play_game(board_size, num_mcts, evaluate, stream, tau_thres=6, noise=True, max_moves=None, inplanes=5):
    """main.py:132-250 for one episode. Returns dict(moves, visits[int64 per move], pis, winner, records)."""
    agent = OracleZeroAgent(board_size, num_mcts, evaluate, stream, noise=noise)
    env = OracleGameState(board_size)
    root_id = (0,)
    win_index = 0
    t = 0
    visits, pis, states = [], [], []
    while win_index == 0 and (max_moves is None or t < max_moves):
        tau = 1 if t < tau_thres else 0
        pi = agent.get_pi(root_id, tau)
        visits.append(agent.visit.astype(np.int64))
        pis.append(pi)
        states.append(root_id)
        onehot, a = get_action(pi, stream)
        root_id = root_id + (a,)
        _, _, win_index, _, _ = env.step(onehot)
        t += 1
    z_black = {1: 1.0, 2: -1.0}.get(win_index, 0.0)
    records = []
    for ply, (sid, pi) in enumerate(zip(states, pis)):  # chronological, black/white interleaved (main.py:219-227)
        z = z_black if ply % 2 == 0 else -z_black
        records.append((sid, pi, z))
    return dict(moves=list(root_id[1:]), visits=visits, pis=pis, winner=win_index, records=records,
                sims=agent.sims, terminal_sims=agent.terminal_sims)


def augment_dataset(memory, board_size):
    """utils.py:226-239 - 8-fold dihedral augmentation of (state[C,B,B], pi[A], z)."""
    out = []
    for s, pi, z in memory:
        for k in range(4):
            s_r = np.rot90(s, k, axes=(1, 2)).copy()
            p_r = np.rot90(pi.reshape(board_size, board_size), k)
            out.append((s_r, p_r.flatten().copy(), z))
            out.append((np.flip(s_r, 2).copy(), np.fliplr(p_r).flatten().copy(), z))
    return out


# --------------------------------------------------------------------------------------------------------------
# Arena (eval_main.py:54-188 Evaluator, :204-333 main): two agents with their own trees, tau = 0, colours swapped
# after every match.  Test infrastructure like everything else in this file.
# --------------------------------------------------------------------------------------------------------------
class OracleRandomAgent:
    """agents.py:637-657 - pi uniform over the empty cells; the move is then argmax_onehot's uniform tie-break."""

    def __init__(self, board_size, stream: DecisionStream):
        self.B, self.A = board_size, board_size * board_size
        self.stream = stream
        self.root_id = None
        self.visit = np.zeros(self.A)
        self.is_real_root = True

    def reset(self):
        self.root_id = None

    def get_pi(self, root_id, tau):
        self.root_id = tuple(root_id)
        empty = (get_board(self.root_id, self.B).reshape(-1) == 0).astype("float")
        return empty / empty.sum()


def arena_matches(board_size, player, enemy, n_match, forced=None, player_black_first=True):
    """eval_main.py:204-333 for `n_match` consecutive matches: the player is black in match 0 (turn 0, enemy_turn 1,
    :213-214), colours swap after every match (:316), both agents are reset after a match (:333).  Per ply
    (eval_main.py:243-283): mover.get_pi(root_id, tau=0) -> utils.argmax_onehot (a second, draw-free arg-max for
    ZeroAgents whose pi already is one-hot; THE uniform tie-break over the empty cells for a RandomAgent) -> root_id =
    mover.root_id + (action,) -> env.step -> opponent.del_parents (memory only).
    `player` / `enemy`: OracleZeroAgent or OracleRandomAgent, each drawing from its own DecisionStream.
    Returns one dict per match: moves, visits [ply][A] of the mover's search, mover ('player'/'enemy') per ply,
    real_root per ply, winner (1 black / 2 white / 3 draw), outcome ('player' / 'enemy' / 'draw').
    `forced` = {(match, ply): action} overrides the move played (test hook: pushes the opponent onto a reply its tree
    never visited); the mover's search and its random draws still happen.  `player_black_first=False` starts with the
    colours the reference has in its second match (the device arena does that in odd slots)."""
    out = []
    forced = dict(forced or {})
    env = OracleGameState(board_size)  # one env for all matches: lazy reset after a finished game (env_small.py:108-117)
    turn, enemy_turn = 0, (1 if player_black_first else 0)
    for m in range(n_match):
        root_id, win_index = (0,), 0
        rec = dict(moves=[], visits=[], movers=[], real_root=[])
        while win_index == 0:
            name, mover = ("player", player) if turn != enemy_turn else ("enemy", enemy)
            pi = mover.get_pi(root_id, 0)
            onehot, a = argmax_onehot(pi, getattr(mover, "host_stream", None) or mover.stream)
            if (m, len(rec["moves"])) in forced:
                a = int(forced[(m, len(rec["moves"]))])
                onehot = np.zeros(len(pi))
                onehot[a] = 1.0
            root_id = mover.root_id + (a,)
            _, _, win_index, turn, _ = env.step(onehot)
            rec["moves"].append(a)
            rec["visits"].append(np.asarray(mover.visit, np.int64).copy())
            rec["movers"].append(name)
            rec["real_root"].append(bool(mover.is_real_root))
        rec["winner"] = win_index
        # `turn` now belongs to the side that did NOT make the last move (eval_main.py:285-312)
        rec["outcome"] = "draw" if win_index == 3 else ("player" if turn == enemy_turn else "enemy")
        rec["player_black"] = enemy_turn == 1
        out.append(rec)
        enemy_turn = abs(enemy_turn - 1)
        turn = 0
        player.reset()
        enemy.reset()
    return out


# --------------------------------------------------------------------------------------------------------------
# PUCTAgent / UCTAgent (agents.py:263-634): pure-MCTS baselines with random rollouts, selectable as arena sides by
# string in eval_main.py:67-84.  Both rebuild their tree on every get_pi (`_init_mcts` overwrites the root with n = 0,
# so the first simulation re-creates all children), run num_mcts + 1 simulations, keep every statistic in Python
# floats (float64) and evaluate a non-root leaf by one uniformly random play-out.
# --------------------------------------------------------------------------------------------------------------
def valid_actions(board: np.ndarray) -> list:
    """utils.py:8-19 - empty cells in ascending index order."""
    return [int(i) for i in np.flatnonzero(board.reshape(-1) == 0)]


def get_reward(win_index: int, leaf_id) -> float:
    """utils.py:208-223."""
    turn = get_turn(leaf_id)
    if win_index == 1:
        return 1.0 if turn == 1 else -1.0
    if win_index == 2:
        return -1.0 if turn == 1 else 1.0
    return 0.0


class _RNode:
    __slots__ = ("acts", "n", "w", "child")

    def __init__(self, acts):
        self.acts = list(acts)
        self.n = [0.0] * len(acts)
        self.w = [0.0] * len(acts)
        self.child = [-1] * len(acts)


class OracleRolloutAgent:
    """kind = 'puct' (agents.py:263-453: uniform prior, u = c_puct * p * sqrt(sum n) / (n + 1), move = most visited
    child) or 'uct' (agents.py:456-634: u = inf for unvisited children else sqrt(2 ln(sum n) / n), move = child with the
    largest q, unvisited children counting as q = 0)."""

    def __init__(self, kind, board_size, num_mcts, stream: DecisionStream):
        assert kind in ("puct", "uct")
        self.kind, self.B, self.A, self.num_mcts = kind, board_size, board_size * board_size, num_mcts
        self.win_mark = 3 if board_size == 3 else 5
        self.c_puct = 5
        self.stream = stream
        self.root_id = None
        self.visit = np.zeros(self.A)
        self.q = np.zeros(self.A)
        self.is_real_root = True
        self.rollout_moves = 0

    def reset(self):
        self.root_id = None

    def _simulate(self, st):
        nodes = st["nodes"]
        moves = list(self.root_id)
        path = []
        node = st["root_node"]
        win = 0
        leaf_is_root = True
        if st["root_n"] > 0:
            while True:  # node has n > 0: it was expanded (or is terminal)
                win = check_win(get_board(tuple(moves), self.B), self.win_mark)
                if win != 0:
                    break
                nd = nodes[node]
                total_n = 0.0
                for x in nd.n:
                    total_n += x
                best, ties = None, []
                for i in range(len(nd.acts)):
                    q = nd.w[i] / nd.n[i] if nd.n[i] > 0 else 0.0
                    if self.kind == "puct":
                        u = self.c_puct * (1 / len(nd.acts)) * np.sqrt(total_n) / (nd.n[i] + 1)
                    else:
                        u = np.inf if nd.n[i] == 0 else np.sqrt(2 * np.log(total_n) / nd.n[i])
                    v = q + u
                    if best is None or v > best:
                        best, ties = v, [i]
                    elif v == best:
                        ties.append(i)
                i = ties[self.stream.choice(len(ties))]
                path.append((node, i))
                moves.append(nd.acts[i])
                leaf_is_root = False
                if nd.n[i] == 0:
                    win = check_win(get_board(tuple(moves), self.B), self.win_mark)
                    node = -1
                    break
                node = nd.child[i]
        leaf_id = tuple(moves)
        if win == 0:
            board = get_board(leaf_id, self.B)
            nodes.append(_RNode(valid_actions(board)))
            if leaf_is_root:
                st["root_node"] = len(nodes) - 1
                reward = 0.0  # "root node don't simulation"
            else:
                pn, pi_ = path[-1]
                nodes[pn].child[pi_] = len(nodes) - 1
                turn_sim = get_turn(leaf_id)
                while True:  # uniformly random play-out (agents.py:391-411)
                    acts = valid_actions(board)
                    a = acts[self.stream.choice(len(acts))]
                    board[a // self.B, a % self.B] = 1.0 if turn_sim == 0 else -1.0
                    self.rollout_moves += 1
                    w = check_win(board, self.win_mark)
                    if w == 0:
                        turn_sim = abs(turn_sim - 1)
                    else:
                        reward = get_reward(w, leaf_id)
                        break
        else:
            reward = 1.0  # "terminal node don't expansion"
        sign = 1.0
        for pn, pi_ in reversed(path):
            nodes[pn].n[pi_] += 1
            nodes[pn].w[pi_] += reward * sign
            sign = -sign
        st["root_n"] += 1
        st["root_w"] += reward * sign

    def get_pi(self, root_id, tau=0):
        """get_pi(root_id, board, turn, tau) of the reference: board and turn follow from root_id"""
        self.root_id = tuple(root_id)
        st = dict(nodes=[], root_node=-1, root_n=0.0, root_w=0.0)
        for _ in range(self.num_mcts + 1):
            self._simulate(st)
        visit = np.zeros(self.A)
        q = np.ones(self.A) * -np.inf
        if st["root_node"] >= 0:
            nd = st["nodes"][st["root_node"]]
            for i, a in enumerate(nd.acts):
                visit[a] = nd.n[i]
                q[a] = nd.w[i] / nd.n[i] if nd.n[i] > 0 else 0.0
        self.visit, self.q = visit, q
        score = visit if self.kind == "puct" else q
        idx = np.flatnonzero(score == score.max())
        pi = np.zeros(self.A)
        pi[idx[(getattr(self, "host_stream", None) or self.stream).choice(len(idx))]] = 1
        return pi
