"""Batched arena: many concurrent matches between two agents, the B200 twin of eval_main.main's match loop
(eval_main.py:204-333; `Evaluator.get_action` :153-170).  Each side owns its own search trees (one engine slot per
match), ZeroAgents search with noise=False and tau=0 and pick `utils.argmax_onehot(pi)`, colours alternate by match
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


def play_matches(player_model, enemy_model=None, n_matches=1024, board_size=9, num_mcts=800, inplanes=5, seed=0,
                 enemy="zero", max_plies=None):
    """player: ZeroAgent(player_model). enemy: 'zero' -> ZeroAgent(enemy_model), 'random' -> RandomAgent.
    Returns dict(player_win, enemy_win, draw, black_win, white_win, plies)."""
    B, A = board_size, board_size * board_size
    sides = {}
    sides["player"] = agents.BatchedZeroAgent(B, num_mcts, inplanes, n_matches, noise=False, seed=seed)
    sides["player"].model = player_model
    if enemy == "zero":
        sides["enemy"] = agents.BatchedZeroAgent(B, num_mcts, inplanes, n_matches, noise=False, seed=seed + 1)
        sides["enemy"].model = enemy_model
    roots = [(0,) for _ in range(n_matches)]
    boards = np.zeros((n_matches, A), np.int8)
    winner = np.zeros(n_matches, np.int32)  # 0 running
    player_is_black = (np.arange(n_matches) % 2) == 0
    ply = 0
    while (winner == 0).any() and (max_plies is None or ply < max_plies):
        black_to_move = ply % 2 == 0
        active = np.flatnonzero(winner == 0)
        for name in ("player", "enemy"):
            mine = active[(player_is_black[active] == black_to_move) == (name == "player")]
            if len(mine) == 0:
                continue
            if name == "enemy" and enemy == "random":
                acts = []
                for m in mine:  # RandomAgent.get_pi + argmax_onehot (agents.py:637-657, eval_main.py:166-168)
                    empty = (boards[m] == 0).astype("float")
                    acts.append(int(utils.argmax_onehot(empty / empty.sum())[1]))
            else:
                pis = sides[name].get_pi([roots[m] for m in mine], 0, game_ids=mine)
                acts = [int(np.argmax(pi)) for pi in pis]  # pi is already the tie-broken one-hot
            for m, a in zip(mine, acts):
                roots[m] = roots[m] + (a,)
                boards[m, a] = 1 if black_to_move else -1
        w = _cabi.check_win_batch(boards[active], B)
        winner[active] = w
        ply += 1
    res = dict(player_win=0, enemy_win=0, draw=0, black_win=int((winner == 1).sum()), white_win=int((winner == 2).sum()),
               plies=[len(r) - 1 for r in roots], unfinished=int((winner == 0).sum()))
    outcomes = []
    for m in range(n_matches):
        if winner[m] == 3:
            res["draw"] += 1
            outcomes.append("draw")
        elif winner[m] in (1, 2):
            black_won = winner[m] == 1
            mine = black_won == bool(player_is_black[m])
            res["player_win" if mine else "enemy_win"] += 1
            outcomes.append("player" if mine else "enemy")
    # the reference plays the matches one after the other and updates the ratings after each (eval_main.py:285-312)
    res["player_elo"], res["enemy_elo"], _, res["winrate"] = elo_sequence(outcomes)
    return res


# ---------------------------------------------------------------------------------------------------------------------
# Drop-in for eval_main.Evaluator / eval_main.main (eval_main.py:54-188, 204-333) without the Flask / pygame wiring:
# one match at a time through the reference-shaped single-game agents (each ZeroAgent = a batch of 1 on the device).
class Evaluator(object):
    """eval_main.py:54-188.  `set_agents(player, enemy, monitor)`: each argument is 'random' or a checkpoint path
    (-> ZeroAgent(noise=False) + PVNet with the reference's tolerant key-by-key load); 'puct' / 'uct' / 'human' / 'web'
    name agents that are out of this repository's scope and raise NotImplementedError."""

    def __init__(self, board_size=9, n_mcts_player=800, n_mcts_enemy=800, n_mcts_monitor=800, n_blocks=10,
                 in_planes=5, out_planes=128):
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
        if path in ("puct", "uct", "human", "web"):
            raise NotImplementedError("agent '%s' (rollout / interactive) is outside the accelerated path" % path)
        agent = agents.ZeroAgent(self.board_size, self.n_mcts[role], self.in_planes, noise=False)
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


def run_matches(evaluator, n_match=12, verbose=False):
    """eval_main.main (eval_main.py:204-333) minus the dashboard objects: alternating colours, the opponent's tree
    pruned after every move (`del_parents`), ELO and win-rate bookkeeping.  Returns (result, player_elo, enemy_elo)."""
    B = evaluator.board_size
    env = evaluator.return_env()
    result = {"Player": 0, "Enemy": 0, "Draw": 0}
    turn, enemy_turn = 0, 1
    player_elo, enemy_elo = 1500, 1500
    for i in range(n_match):
        board = np.zeros([B, B])
        root_id, win_index, action_index = (0,), 0, None
        while win_index == 0:
            if verbose:
                utils.render_str(board, B, action_index)
            evaluator.monitor.get_pv(root_id)
            action, action_index = evaluator.get_action(root_id, board, turn, enemy_turn)
            mover = evaluator.player if turn != enemy_turn else evaluator.enemy
            root_id = (mover.root_id if mover.root_id is not None else root_id) + (int(action_index),)
            board, _, win_index, turn, _ = env.step(action)
            (evaluator.enemy if turn == enemy_turn else evaluator.player).del_parents(root_id)
            if win_index != 0:
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
                evaluator.reset()
    return result, player_elo, enemy_elo
