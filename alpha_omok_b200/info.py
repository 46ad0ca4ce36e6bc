"""Dashboard feed objects of the arena: `AgentInfo` / `GameInfo` (info/agent_info.py, info/game_info.py)
and the two status payloads the reference's Flask routes return (webapi.py:28-76) as plain functions - no web server
here (Flask / JS frontend are out of scope), but a maintainer's `@route` can `jsonify` these dicts unchanged.
`arena.run_matches(..., dashboard=Dashboard(B))` keeps them up to date exactly where eval_main.main does
(eval_main.py:206-207, 221-222, 230, 236, 259-283, 286-290, 318-319)."""
from __future__ import annotations

import numpy as np

from . import agents


class AgentInfo:
    """info/agent_info.py:5-28"""

    def __init__(self, board_size):
        self.p = np.zeros([board_size, board_size])
        self.p_size = board_size * board_size
        self.visit = np.zeros([board_size, board_size])
        self.visit_size = board_size * board_size
        self.moves = []
        self.values = []
        self.agent = agents.Agent(board_size)

    def add_value(self, move, value):
        self.moves.append(move)
        self.values.append((value + 1.0) / 2.0 * 100.0)   # win probability in percent

    def clear_values(self):
        self.moves = []
        self.values = []


class GameInfo:
    """info/game_info.py:4-17"""

    def __init__(self, board_size):
        self.game_board = np.zeros([board_size, board_size])
        self.win_index = 0
        self.curr_turn = 0
        self.enemy_turn = 0
        self.action_index = -1
        self.message = "오목"
        self.player_agent_name = ""
        self.enemy_agent_name = ""
        self.enemy_action_index = -1
        self.game_status = 0


class Dashboard:
    """the three module-level objects of webapi.py:12-16"""

    def __init__(self, board_size):
        self.game_info = GameInfo(board_size)
        self.player_agent_info = AgentInfo(board_size)
        self.enemy_agent_info = AgentInfo(board_size)

    def periodic_status(self):
        return periodic_status(self.game_info, self.player_agent_info, self.enemy_agent_info)

    def prompt_status(self):
        return prompt_status(self.player_agent_info, self.enemy_agent_info)


def periodic_status(game_info, player_agent_info, enemy_agent_info):
    """payload of GET /periodic_status (webapi.py:28-64), polled every 500 ms by static/dashboard.js"""
    data = {"success": False}
    data["game_board_size"] = game_info.game_board.shape[0]
    data["game_board_values"] = game_info.game_board.reshape(game_info.game_board.size).astype(int).tolist()
    data["game_board_message"] = game_info.message
    data["action_index"] = game_info.action_index
    data["win_index"] = game_info.win_index
    data["curr_turn"] = game_info.curr_turn
    data["enemy_turn"] = game_info.enemy_turn
    data["player_agent_name"] = player_agent_info.agent.get_name()
    data["enemy_agent_name"] = enemy_agent_info.agent.get_name()
    for side, info in (("player", player_agent_info), ("enemy", enemy_agent_info)):
        data[side + "_agent_p_size"] = info.p_size
        data[side + "_agent_p_values"] = np.asarray(info.p).reshape(info.p_size).astype(float).tolist()
        data[side + "_agent_visit_size"] = info.visit_size
        data[side + "_agent_visit_values"] = np.asarray(info.visit).reshape(info.visit_size).astype(float).tolist()
    data["player_agent_moves"] = player_agent_info.moves
    data["player_agent_values"] = player_agent_info.values
    data["enemy_agent_moves"] = enemy_agent_info.moves
    data["enemy_agent_values"] = enemy_agent_info.values
    data["success"] = True
    return data


def prompt_status(player_agent_info, enemy_agent_info):
    """payload of GET /prompt_status (webapi.py:67-76), polled every 100 ms"""
    return {"success": True, "player_message": player_agent_info.agent.get_message(),
            "enemy_message": enemy_agent_info.agent.get_message()}
