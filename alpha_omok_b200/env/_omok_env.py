"""GameState for Omok: drop-in for env/env_small.py and env/env_regular.py (step / board params only; the pygame
rendering of the reference is out of scope).  The five-in-a-row test runs on the GPU (utils.check_win -> ao_check_win).
"""
from __future__ import annotations

import numpy as np

from .. import utils

WIN_STONES = 5


def make_env(board_size, name):
    class GameState:
        """env_small.py:60-199 (text mode). `step(onehot)` -> (gameboard, check_valid_pos, win_index, turn, action_index)"""

        def __init__(self, gamemode="text"):
            if gamemode != "text":
                raise NotImplementedError("only gamemode='text' is supported (no pygame window)")
            self.gamemode = gamemode
            self.init = False
            self.num_stones = 0
            self.gameboard = np.zeros([board_size, board_size])
            self.black_win = self.white_win = self.count_draw = 0
            self.turn = 0  # 0 black, 1 white

        def step(self, input_):
            if self.init:  # lazy reset after a finished game (env_small.py:108-117)
                self.num_stones, self.turn, self.init = 0, 0, False
                self.gameboard = np.zeros([board_size, board_size])
            check_valid_pos, action_index = False, 0
            if np.any(input_):
                action_index = np.argmax(input_)
                y, x = int(action_index / board_size), action_index % board_size
                check_valid_pos = self.gameboard[y, x] == 0
                # like the reference, an occupied cell is overwritten (env_small.py:161-176)
                self.gameboard[y, x] = 1 if self.turn == 0 else -1
                self.turn ^= 1
                self.num_stones += 1
            win_index = utils.check_win(self.gameboard, WIN_STONES)
            if win_index == 1:
                self.black_win += 1
            elif win_index == 2:
                self.white_win += 1
            elif win_index == 3:
                self.count_draw += 1
            self.init = win_index != 0  # env_small.py:299-316
            return self.gameboard, bool(check_valid_pos), win_index, self.turn, action_index

    def ReturnName():
        return name

    def Return_Num_Action():
        return board_size * board_size

    def Return_BoardParams():
        return board_size, board_size ** 2

    return GameState, ReturnName, Return_Num_Action, Return_BoardParams
