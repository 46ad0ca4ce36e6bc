"""Batched arena: many concurrent matches between two agents, the B200 twin of eval_main.main's match loop
(eval_main.py:204-333; `Evaluator.get_action` :153-170), played entirely on the device (csrc/tree.cu arena_move).
Each side owns its own search tree (one engine slot per side and match) and its own network (two weight sets in one
engine), ZeroAgents search with noise=False and tau=0 and pick `utils.argmax_onehot(pi)`, colours alternate by match
parity (eval_main.py:233,316).  `elo` / `elo_sequence` restate the reference's rating update (eval_main.py:191-198,
285-312); the web dashboard feed stays out of scope."""
from __future__ import annotations

import numpy as np

from . import _cabi, agents, utils


def elo(player_elo, enemy_elo, p_winscore, e_winscore):
    """eval_main.py:191-198 (K = 32, logistic expectation on a 400-point scale)."""
    elo_diff = enemy_elo - player_elo
    ex_pw = 1 / (1 + 10 ** (elo_diff / 400))
    ex_ew = 1 / (1 + 10 ** (-elo_diff / 400))
    player_elo += 32 * (p_winscore - ex_pw)
    enemy_elo += 32 * (e_winscore - ex_ew)
    return player_elo, enemy_elo


def elo_sequence(outcomes, player_elo=1500, enemy_elo=1500):
    """Ratings after a sequence of match outcomes in match order ('player' | 'enemy' | 'draw'), starting from the
    reference's 1500 / 1500 (eval_main.py:216-217, 285-312).  Returns (player_elo, enemy_elo, result, winrate %)."""
    result = {"Player": 0, "Enemy": 0, "Draw": 0}
    for o in outcomes:
        if o == "draw":
            result["Draw"] += 1
            player_elo, enemy_elo = elo(player_elo, enemy_elo, 0.5, 0.5)
        elif o == "player":
            result["Player"] += 1
            player_elo, enemy_elo = elo(player_elo, enemy_elo, 1, 0)
        else:
            result["Enemy"] += 1
            player_elo, enemy_elo = elo(player_elo, enemy_elo, 0, 1)
    n = sum(result.values())
    winrate = (result["Player"] + 0.5 * result["Draw"]) / n * 100 if n else 0.0
    return player_elo, enemy_elo, result, winrate


def _n_blocks(model):
    return getattr(model, "n_block", None) or len({k.split(".")[1] for k in model.state_dict() if k.startswith("layers.")})


def decode_match_records(slab, board_size):
    """uint8 [n, record_bytes] (ao_arena_begin's record slab) -> list of dict(moves, visits [plies][A], winner,
    player_black, outcome 'player' | 'enemy' | 'draw' | None while unfinished)."""
    A = board_size * board_size
    buf = slab.cpu().numpy() if hasattr(slab, "cpu") else np.asarray(slab)
    voff = (4 + 2 * A + 3) & ~3
    out = []
    for rec in buf:
        k = int(rec[:2].view(np.int16)[0])
        w, pb = int(rec[2]), bool(rec[3])
        moves = [int(a) for a in rec[4:4 + 2 * A].view(np.int16)[:k]]
        visits = rec[voff:voff + 4 * A * A].view(np.uint32).reshape(A, A)[:k].copy()
        outcome = None if w == 0 else "draw" if w == 3 else ("player" if (w == 1) == pb else "enemy")
        out.append(dict(moves=moves, visits=visits, winner=w, player_black=pb, outcome=outcome))
    return out


def play_matches(player_model, enemy_model=None, n_matches=1024, board_size=9, num_mcts=800, inplanes=5, seed=0,
                 enemy="zero", matches_per_slot=1, first_key=0, num_mcts_enemy=None, nn_precision="auto",
                 rounds_per_call=256, max_rounds=None, engine_kwargs=None, return_records=False):
    """`n_matches` concurrent runs of eval_main.main's match loop (eval_main.py:204-333), each `matches_per_slot`
    matches long, entirely on the device (ao_arena_begin): the match loop, both sides' trees and both networks live in
    one engine, the two towers run back to back every round.
    player: ZeroAgent(player_model). enemy: 'zero' -> ZeroAgent(enemy_model), 'random' -> RandomAgent, 'puct' / 'uct'
    -> PUCTAgent / UCTAgent(num_mcts_enemy) (eval_main.py:67-84).
    `player_model` / `enemy_model`: nn.Module with the reference's parameter names, or a state_dict (then pass
    engine_kwargs=dict(n_blocks=...) if it is not a 10-block net).
    Returns dict(player_win, enemy_win, draw, black_win, white_win, plies, unfinished, player_elo, enemy_elo, winrate
    [, records])."""
    B = board_size
    kw = dict(engine_kwargs or {})

    def sd_of(m):
        return m if isinstance(m, dict) else m.state_dict()

    if "n_blocks" not in kw and not isinstance(player_model, dict):
        kw["n_blocks"] = _n_blocks(player_model)
    kw.setdefault("device", _cabi.default_device(None if isinstance(player_model, dict) else player_model))
    eng = _cabi.Engine(board_size=B, num_mcts=num_mcts, max_games=2 * n_matches, noise=False, inplanes=inplanes,
                       seed=seed, **kw)
    try:
        if kw.get("eval_mode", _cabi.AO_EVAL_PVNET) == _cabi.AO_EVAL_PVNET:
            eng.load_state_dict(sd_of(player_model), which=0)
            if nn_precision == "auto":
                eng.choose_nn_precision(which=0)
            else:
                eng.set_nn_precision(nn_precision, which=0)
            if enemy == "zero":
                eng.load_state_dict(sd_of(enemy_model), which=1)
                if nn_precision == "auto":
                    eng.choose_nn_precision(which=1)
                else:
                    eng.set_nn_precision(nn_precision, which=1)
        eng.arena_begin(n_matches, first_key=first_key, matches_per_slot=matches_per_slot, enemy_kind=enemy,
                        keep_records=True, n_mcts_player=num_mcts, n_mcts_enemy=num_mcts_enemy or num_mcts)
        st = eng.selfplay_rounds(rounds_per_call)
        rounds = rounds_per_call
        while st["running"] and (max_rounds is None or rounds < max_rounds):
            st = eng.selfplay_rounds(rounds_per_call)
            rounds += rounds_per_call
        if st["errors"]:
            raise _cabi.AoError("%d match tree(s) overflowed their arena; raise node_cap" % st["errors"])
        from . import replay
        recs = decode_match_records(replay.device_stream_records(eng), B)
        precisions = (eng.nn_precision, eng.nn_precision_enemy)
    finally:
        eng.close()
    res = dict(player_win=0, enemy_win=0, draw=0, black_win=0, white_win=0, plies=[len(r["moves"]) for r in recs],
               unfinished=0, sims=st["sims"], nn_precision=precisions)
    outcomes = []
    for r in recs:
        if r["outcome"] is None:
            res["unfinished"] += 1
            continue
        res["black_win"] += r["winner"] == 1
        res["white_win"] += r["winner"] == 2
        res[{"player": "player_win", "enemy": "enemy_win", "draw": "draw"}[r["outcome"]]] += 1
        outcomes.append(r["outcome"])
    # the reference plays its matches one after the other and updates the ratings after each (eval_main.py:285-312);
    # here: in record order (slot-major)
    res["player_elo"], res["enemy_elo"], _, res["winrate"] = elo_sequence(outcomes)
    if return_records:
        res["records"] = recs
    return res


# ---------------------------------------------------------------------------------------------------------------------
# Drop-in for eval_main.Evaluator / eval_main.main (eval_main.py:54-188, 204-333) without the Flask / pygame wiring:
# one match at a time through the reference-shaped single-game agents (each ZeroAgent = a batch of 1 on the device).
class Evaluator(object):
    """eval_main.py:54-188.  `set_agents(player, enemy, monitor)`: each argument is 'random' or a checkpoint path
    (-> ZeroAgent(noise=False) + PVNet with the reference's tolerant key-by-key load), 'puct' / 'uct' (the play-out
    agents of agents.py:263-634, searched on the device too); 'human' / 'web' are interactive I/O and raise
    NotImplementedError."""

    def __init__(self, board_size=9, n_mcts_player=800, n_mcts_enemy=800, n_mcts_monitor=800, n_blocks=10,
                 in_planes=5, out_planes=128, engine_kwargs=None):
        self.engine_kwargs = dict(engine_kwargs or {})
        self.board_size = board_size
        self.n_mcts = {"player": n_mcts_player, "enemy": n_mcts_enemy, "monitor": n_mcts_monitor}
        self.n_blocks, self.in_planes, self.out_planes = n_blocks, in_planes, out_planes
        self.player = self.enemy = self.monitor = None
        self.env = None

    def _make(self, role, path):
        import torch
        from . import model
        if path == "random":
            return agents.RandomAgent(self.board_size)
        if path == "puct":
            return agents.PUCTAgent(self.board_size, self.n_mcts[role], engine_kwargs=self.engine_kwargs)
        if path == "uct":
            return agents.UCTAgent(self.board_size, self.n_mcts[role], engine_kwargs=self.engine_kwargs)
        if path in ("human", "web"):
            raise NotImplementedError("agent '%s' (interactive pygame / Flask input) is outside the accelerated path" % path)
        agent = agents.ZeroAgent(self.board_size, self.n_mcts[role], self.in_planes, noise=False,
                                 engine_kwargs=self.engine_kwargs)
        agent.model = model.PVNet(self.n_blocks, self.in_planes, self.out_planes, self.board_size)
        state = agent.model.state_dict()
        loaded = path if isinstance(path, dict) else torch.load(path, map_location="cpu")
        for k, v in loaded.items():          # eval_main.py:95-101: keys the module does not know are ignored,
            if k in state:                   # keys the file lacks (num_batches_tracked) keep their defaults
                state[k] = v
        agent.model.load_state_dict(state)
        return agent

    def set_agents(self, model_path_a, model_path_b, model_path_m):
        from .env import env_regular, env_small
        game = env_small if self.board_size == 9 else env_regular
        self.env = game.GameState("text")
        self.player = self._make("player", model_path_a)
        self.enemy = self._make("enemy", model_path_b)
        self.monitor = self._make("monitor", model_path_m)

    def get_action(self, root_id, board, turn, enemy_turn):
        """eval_main.py:153-170"""
        mover = self.player if turn != enemy_turn else self.enemy
        if isinstance(mover, agents.ZeroAgent):
            pi = mover.get_pi(root_id, tau=0)
        else:
            pi = mover.get_pi(root_id, board, turn, tau=0)
            if mover is self.player:
                self.monitor.get_pi(root_id, tau=0)      # "for monitor" (eval_main.py:160-161)
        return utils.argmax_onehot(pi)

    def return_env(self):
        return self.env

    def reset(self):
        self.player.reset()
        self.enemy.reset()


def run_matches(evaluator, n_match=12, verbose=False, dashboard=None):
    """eval_main.main (eval_main.py:204-333): alternating colours, the opponent's tree pruned after every move
    (`del_parents`), ELO and win-rate bookkeeping.  `dashboard` (an `info.Dashboard`) receives the updates the
    reference makes to webapi's game_info / player_agent_info / enemy_agent_info at the same points of the loop, so a
    web front end polling `dashboard.periodic_status()` sees what the reference's would.
    Returns (result, player_elo, enemy_elo)."""
    B = evaluator.board_size
    env = evaluator.return_env()
    result = {"Player": 0, "Enemy": 0, "Draw": 0}
    turn, enemy_turn = 0, 1
    player_elo, enemy_elo = 1500, 1500
    gi = pi_ = ei = None
    if dashboard is not None:
        gi, pi_, ei = dashboard.game_info, dashboard.player_agent_info, dashboard.enemy_agent_info
        pi_.agent, ei.agent = evaluator.player, evaluator.enemy      # eval_main.py:206-207
        gi.enemy_turn, gi.game_status = enemy_turn, 0                # :221-222
    interactive = ()                                                 # HumanAgent / WebAgent: not built here
    for i in range(n_match):
        board = np.zeros([B, B])
        root_id, win_index, action_index = (0,), 0, None
        if gi is not None:
            gi.game_board, gi.game_status = board, 0                 # :230, :236
        while win_index == 0:
            if verbose:
                utils.render_str(board, B, action_index)
            p, v = evaluator.monitor.get_pv(root_id)
            action, action_index = evaluator.get_action(root_id, board, turn, enemy_turn)
            mover = evaluator.player if turn != enemy_turn else evaluator.enemy
            root_id = (mover.root_id if mover.root_id is not None else root_id) + (int(action_index),)
            board, _, win_index, turn, _ = env.step(action)
            if gi is not None:                                       # :259-262
                gi.game_board, gi.action_index, gi.win_index, gi.curr_turn = board, int(action_index), win_index, turn
            move = np.count_nonzero(board)
            if turn == enemy_turn:                                   # the player just moved (:266-277)
                if pi_ is not None:
                    src = evaluator.monitor if isinstance(evaluator.player, interactive) else evaluator.player
                    pi_.visit, pi_.p = src.get_visit(), src.get_policy()
                    pi_.add_value(move, v)
                evaluator.enemy.del_parents(root_id)
            else:                                                    # the enemy just moved (:279-283)
                if ei is not None:
                    ei.visit, ei.p = evaluator.enemy.get_visit(), evaluator.enemy.get_policy()
                    ei.add_value(move, v)
                evaluator.player.del_parents(root_id)
            if win_index != 0:
                if pi_ is not None:                                  # :286-290
                    pi_.clear_values()
                    ei.clear_values()
                    gi.game_status = win_index
                if win_index == 3:
                    result["Draw"] += 1
                    player_elo, enemy_elo = elo(player_elo, enemy_elo, 0.5, 0.5)
                elif turn == enemy_turn:      # the side that just moved (and won) was the player
                    result["Player"] += 1
                    player_elo, enemy_elo = elo(player_elo, enemy_elo, 1, 0)
                else:
                    result["Enemy"] += 1
                    player_elo, enemy_elo = elo(player_elo, enemy_elo, 0, 1)
                enemy_turn = abs(enemy_turn - 1)      # swap colours (eval_main.py:316)
                turn = 0
                if gi is not None:
                    gi.enemy_turn, gi.curr_turn = enemy_turn, turn   # :318-319
                evaluator.reset()
    return result, player_elo, enemy_elo
