"""Drop-in for 2_AlphaOmok/agents.py: `Agent`, `ZeroAgent` (same constructor, attributes and methods), plus the
batched twins that make the B200 path fast: `BatchedZeroAgent` (get_pi for thousands of games per call) and
`self_play` (main.py:122-250 for a whole batch of episodes, resident on the device).

The search itself (agents.py:60-239) runs in csrc/tree.cu + csrc/tower.cu; this file only marshals IDs in and
visit counts / priors out through the C ABI and finishes `get_pi` with the reference's own host arithmetic
(visit / visit.sum(), argmax_onehot with numpy's global RNG).
"""
from __future__ import annotations

import numpy as np

from . import _cabi, utils

PRINT_MCTS = False  # agents.py:11 (per-simulation printing is not reproduced; `message` is updated once per search)


class Agent(object):
    """agents.py:16-36"""

    def __init__(self, board_size):
        self.policy = np.zeros(board_size ** 2, "float")
        self.visit = np.zeros(board_size ** 2, "float")
        self.message = "Hello"

    def get_policy(self):
        return self.policy

    def get_visit(self):
        return self.visit

    def get_name(self):
        return type(self).__name__

    def get_message(self):
        return self.message

    def get_pv(self, root_id):
        return None, None


class _EngineOwner:
    """Shared plumbing: lazily create the CUDA engine and keep its weights in sync with `self.model`."""

    _engine = None
    _fingerprint = None

    def _make_engine(self, max_games, **kw):
        m = self.model
        n_blocks = getattr(m, "n_block", None)
        if n_blocks is None:  # any nn.Module with the reference's parameter names
            n_blocks = len({k.split(".")[1] for k in m.state_dict() if k.startswith("layers.")})
        if getattr(self, "nn_precision", "auto") != "auto":
            kw.setdefault("nn_precision", self.nn_precision)
        kw.setdefault("device", _cabi.default_device(m))
        return _cabi.Engine(board_size=self.board_size, num_mcts=self.num_mcts, max_games=max_games,
                            noise=self.noise, n_blocks=n_blocks, inplanes=self.inplanes, c_puct=self.c_puct,
                            alpha=self.alpha, **kw)

    def invalidate_weights(self):
        """Force a re-upload of `self.model`'s weights at the next search.  Needed only after in-place writes that bypass
        autograd's version counter (`p.data.mul_()`, EMA code, ...) or after swapping sub-modules: optimizer steps,
        `load_state_dict`, `.to()`, replaced Parameter objects and - for CPU-resident weights - `.data` writes that touch
        the first, middle or last tensor (sampled values) are detected automatically."""
        self._fingerprint = None
        self._slots = None

    _slots = None  # (id(model), [(module._parameters | module._buffers dict, name), ...])

    def _weight_tensors(self):
        """The live parameter / buffer tensors of `self.model`, looked up by name in every sub-module's own dict on every
        call (so a replaced Parameter object is seen) without re-walking the module tree (0.2-0.3 ms per search for
        state_dict() / parameters(), a tenth of a 40-simulation search).  Swapping whole sub-modules after the first
        search needs `invalidate_weights()`."""
        m = self.model
        if self._slots is None or self._slots[0] != id(m):
            slots = []
            for mod in m.modules():
                slots += [(mod._parameters, k) for k in mod._parameters]
                slots += [(mod._buffers, k) for k in mod._buffers]
            self._slots = (id(m), slots)
        return [d.get(k) for d, k in self._slots[1]]

    def _sync_weights(self):
        if self.model is None:
            raise RuntimeError("ZeroAgent.model is not set (assign a PVNet before searching, main.py:81)")
        # identify the weights by storage address and version counter (bumped by optimizer steps and load_state_dict);
        # `.data` writes use their own counter, so three CPU tensors also contribute a few sampled values (cheap, no sync)
        ts = [t for t in self._weight_tensors() if t is not None]
        probe = ()
        if ts and ts[0].device.type == "cpu":
            probe = tuple(float(t.detach().reshape(-1)[:4].double().sum())
                          for t in (ts[0], ts[len(ts) // 2], ts[-1]) if t.numel() and t.is_floating_point())
        fp = (id(self.model), probe) + tuple((t.data_ptr(), t._version) for t in ts)
        if fp != self._fingerprint:
            self._engine.load_state_dict(self.model.state_dict())
            self._fingerprint = fp
            if getattr(self, "nn_precision", "auto") == "auto":
                # cheapest tower mode that keeps policy / value within the 1e-4 contract for THESE weights
                self._engine.choose_nn_precision()


class ZeroAgent(Agent, _EngineOwner):
    """agents.py:39-260.  One game per agent, exactly the reference's surface; the tree lives in HBM."""

    def __init__(self, board_size, num_mcts, inplanes, noise=True, seed=0, engine_kwargs=None, nn_precision="auto"):
        super(ZeroAgent, self).__init__(board_size)
        self.nn_precision = nn_precision  # "auto": fp16 single pass unless the weights need the hi/lo split mode
        self.board_size = board_size
        self.num_mcts = num_mcts
        self.inplanes = inplanes
        self.win_mark = 3 if board_size == 3 else 5
        self.alpha = 10 / self.board_size ** 2
        self.c_puct = 5
        self.noise = noise
        self.root_id = None
        self.model = None
        self.is_real_root = True
        self._seed = seed
        self._engine_kwargs = dict(engine_kwargs or {})
        self._episode = 0

    @property
    def tree(self):
        raise AttributeError("the search tree is resident in GPU memory; use get_visit()/get_policy()")

    def _ensure_engine(self):
        if self._engine is None:
            self._engine = self._make_engine(1, seed=self._seed, **self._engine_kwargs)
            self._engine.games_reset([0], keys=[self._episode])
        self._sync_weights()
        return self._engine

    def reset(self):
        self.root_id = None
        self.is_real_root = True
        self._episode += 1
        if self._engine is not None:
            self._engine.games_reset([0], keys=[self._episode])

    def get_pi(self, root_id, tau):
        eng = self._ensure_engine()
        self.root_id = tuple(root_id)
        visits, priors, real = eng.search([0], [self.root_id])
        self.is_real_root = bool(real[0])
        self.visit = visits[0].astype("float")
        self.policy = priors[0]
        self.message = "simulation: {}\r".format(self.num_mcts + (1 if self.is_real_root else 0))
        pi = self.visit / self.visit.sum()
        if tau == 0:
            pi, _ = utils.argmax_onehot(pi)
        return pi

    def del_parents(self, root_id):
        """agents.py:241-250 prunes dict entries above the new root; here every root advance already compacts the
        reachable subtree into the other arena, so there is nothing left to delete."""
        return None

    def get_pv(self, root_id):
        eng = self._ensure_engine()
        state = utils.get_state_pt(root_id, self.board_size, self.inplanes)
        p, v = eng.nn_forward(state[None].astype(np.float32))
        return p[0], v[0]


class RandomAgent(Agent):
    """agents.py:637-657 - uniform pi over the empty cells (host side; arena opponent)"""

    def __init__(self, board_size):
        super(RandomAgent, self).__init__(board_size)
        self.board_size = board_size
        self.root_id = None

    def get_pi(self, root_id, board, turn, tau):
        self.root_id = root_id
        empty = (np.asarray(board).reshape(-1) == 0).astype("float")
        return empty / empty.sum()

    def reset(self):
        self.root_id = None

    def del_parents(self, root_id):
        return None


class _RolloutAgent(Agent):
    """Shared body of PUCTAgent / UCTAgent (agents.py:263-634): pure MCTS with uniformly random play-outs, no network.
    Same constructor and `get_pi(root_id, board, turn, tau)` as the reference; the search (a fresh tree and
    num_mcts + 1 simulations per call) runs in csrc/rollout.cuh through `ao_rollout_search`, the final arg-max with its
    tie-break is the reference's own numpy code."""

    kind = None

    def __init__(self, board_size, num_mcts, seed=0, engine_kwargs=None):
        super(_RolloutAgent, self).__init__(board_size)
        self.board_size = board_size
        self.num_mcts = num_mcts
        self.win_mark = 3 if board_size == 3 else 5
        self.c_puct = 5
        self.root_id = None
        self.board = None
        self.turn = None
        self.is_real_root = True
        self._seed = seed
        self._engine_kwargs = dict(engine_kwargs or {})
        self._engine = None
        self._episode = 0

    def _ensure_engine(self):
        if self._engine is None:
            kw = dict(self._engine_kwargs)
            kw.setdefault("device", _cabi.default_device())
            kw.setdefault("eval_mode", _cabi.AO_EVAL_SYNTH)  # no network on this path
            self._engine = _cabi.Engine(board_size=self.board_size, num_mcts=self.num_mcts, max_games=1, noise=False,
                                        seed=self._seed, node_cap=max(2048, self.num_mcts + 8), **kw)
            self._engine.games_reset([0], keys=[self._episode])
        return self._engine

    def reset(self):
        self.is_real_root = True
        self.root_id = None
        self.board = None
        self.turn = None
        self._episode += 1
        if self._engine is not None:
            self._engine.games_reset([0], keys=[self._episode])

    def _search(self, root_id, board, turn):
        self.root_id, self.board, self.turn = tuple(int(a) for a in root_id), board, turn
        visits, w = self._ensure_engine().rollout_search(self.kind, [0], [self.root_id], self.num_mcts)
        self.visit = visits[0].astype("float")
        self.message = "simulation: {}\r".format(self.num_mcts + 1)
        return self.visit, w[0].astype("float")

    def del_parents(self, root_id):
        return None  # every get_pi rebuilds the tree (`_init_mcts` overwrites the root): nothing is kept


class PUCTAgent(_RolloutAgent):
    """agents.py:263-453"""
    kind = "puct"

    def get_pi(self, root_id, board, turn, tau):
        visit, _ = self._search(root_id, board, turn)
        pi = np.zeros(self.board_size ** 2, "float")
        max_idx = np.argwhere(visit == visit.max())
        pi[max_idx[np.random.choice(len(max_idx))]] = 1     # agents.py:293-294
        return pi


class UCTAgent(_RolloutAgent):
    """agents.py:456-634"""
    kind = "uct"

    def get_pi(self, root_id, board, turn, tau):
        visit, w = self._search(root_id, board, turn)
        q = np.ones(self.board_size ** 2, "float") * -np.inf
        occupied = set(self.root_id[1:])
        for a in range(self.board_size ** 2):
            if a not in occupied:                            # the root's children (agents.py:469-471)
                q[a] = w[a] / visit[a] if visit[a] > 0 else 0.0
        pi = np.zeros(self.board_size ** 2, "float")
        max_idx = np.argwhere(q == q.max())
        pi[max_idx[np.random.choice(len(max_idx))]] = 1     # agents.py:473-474
        return pi


class BatchedZeroAgent(_EngineOwner):
    """get_pi for many independent games per call (slot g of the engine = game g).

    visits, priors = agent.search(root_ids)           # uint32 [n][A], float64 [n][A]
    pis = agent.get_pi(root_ids, taus)                # float64 [n][A], reference arithmetic per row
    """

    def __init__(self, board_size, num_mcts, inplanes, n_games, noise=True, seed=0, engine_kwargs=None,
                 nn_precision="auto"):
        self.nn_precision = nn_precision
        self.board_size, self.num_mcts, self.inplanes, self.noise = board_size, num_mcts, inplanes, noise
        self.n_games = n_games
        self.alpha = 10 / board_size ** 2
        self.c_puct = 5
        self.model = None
        self._seed = seed
        self._engine_kwargs = dict(engine_kwargs or {})
        self._ids = np.arange(n_games, dtype=np.int32)
        self.visit = np.zeros((n_games, board_size ** 2))
        self.policy = np.zeros((n_games, board_size ** 2))

    def _ensure_engine(self):
        if self._engine is None:
            self._engine = self._make_engine(self.n_games, seed=self._seed, **self._engine_kwargs)
            self._engine.games_reset(self._ids, keys=self._ids.astype(np.uint32))
        self._sync_weights()
        return self._engine

    def reset(self, first_key=0):
        if self._engine is not None:
            self._engine.games_reset(self._ids, keys=(self._ids + first_key).astype(np.uint32))

    def search(self, root_ids, game_ids=None):
        """root_ids[i] is searched in engine slot game_ids[i] (default: slot i); each slot keeps its own tree."""
        eng = self._ensure_engine()
        ids = self._ids[:len(root_ids)] if game_ids is None else np.ascontiguousarray(game_ids, np.int32)
        visits, priors, real = eng.search(ids, root_ids)
        self.visit, self.policy, self.is_real_root = visits.astype("float"), priors, real.astype(bool)
        return visits, priors

    def get_pi(self, root_ids, taus, game_ids=None):
        visits, _ = self.search(root_ids, game_ids)
        pis = visits / visits.sum(axis=1, keepdims=True)
        for i, tau in enumerate(np.broadcast_to(taus, (len(root_ids),))):
            if tau == 0:
                pis[i], _ = utils.argmax_onehot(pis[i])
        return pis


def self_play(model, n_selfplay, board_size=9, num_mcts=400, inplanes=5, tau_thres=6, noise=True, seed=0,
              first_key=0, rounds_per_call=256, engine=None, augment=False, max_slots=None):
    """Batched twin of main.self_play (main.py:122-250): `n_selfplay` episodes played concurrently on the device.

    Returns (cur_memory, result): cur_memory is the reference's list of (state [5,B,B] float64, pi [A] float64, z)
    in chronological black/white-interleaved order per episode; result = {'Black','White','Draw'} counts.
    """
    A = board_size * board_size
    own = engine is None
    slots = n_selfplay if max_slots is None else min(n_selfplay, max_slots)
    if own:
        n_blocks = getattr(model, "n_block", None) or len(
            {k.split(".")[1] for k in model.state_dict() if k.startswith("layers.")})
        engine = _cabi.Engine(board_size=board_size, num_mcts=num_mcts, max_games=slots, noise=noise,
                              tau_thres=tau_thres, n_blocks=n_blocks, inplanes=inplanes, seed=seed,
                              device=_cabi.default_device(model))
        engine.load_state_dict(model.state_dict())
    else:
        slots = min(slots, engine.G)
    stream = n_selfplay > slots   # more episodes than resident games: continuous mode (ao_selfplay_stream_begin)
    if stream:
        engine.selfplay_stream_begin(n_selfplay, n_slots=slots, first_key=first_key)
    else:
        engine.selfplay_begin(n_selfplay, first_key=first_key)
    st = engine.selfplay_rounds(rounds_per_call)
    while st["running"]:
        st = engine.selfplay_rounds(rounds_per_call)
    if st["errors"]:
        raise _cabi.AoError("%d game tree(s) overflowed their arena; raise node_cap" % st["errors"])
    if stream:
        from . import replay
        cur_memory, result = replay.decode_records(replay.device_stream_records(engine), board_size, tau_thres)
        if own:
            engine.close()
        return (utils.augment_dataset(cur_memory, board_size) if augment else cur_memory), result
    moves, n_moves, winners, visits = engine.selfplay_fetch(n_selfplay)
    cur_memory, result = [], {"Black": 0, "White": 0, "Draw": 0}
    for g in range(n_selfplay):
        k, w = int(n_moves[g]), int(winners[g])
        result["Black" if w == 1 else "White" if w == 2 else "Draw"] += 1
        z_black = 1.0 if w == 1 else -1.0 if w == 2 else 0.0
        ids = [(0,) + tuple(int(a) for a in moves[g, :t]) for t in range(k)]
        states = _cabi.encode_state_batch(ids, board_size).astype(np.float64)
        for t in range(k):
            v = visits[g, t].astype(np.float64)
            if t < tau_thres:
                pi = v / v.sum()
            else:  # one-hot of the tie-broken argmax == the action that was played (main.py:150-166)
                pi = np.zeros(A)
                pi[moves[g, t]] = 1.0
            cur_memory.append((states[t], pi, z_black if t % 2 == 0 else -z_black))
    if own:
        engine.close()
    if augment:
        cur_memory = utils.augment_dataset(cur_memory, board_size)
    return cur_memory, result
