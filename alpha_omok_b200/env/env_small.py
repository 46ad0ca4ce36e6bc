"""9x9 Omok environment - drop-in for 2_AlphaOmok/env/env_small.py (GameState.step, Return_BoardParams, ...)."""
from ._omok_env import make_env

GAMEBOARD_SIZE = 9
WIN_STONES = 5
GameState, ReturnName, Return_Num_Action, Return_BoardParams = make_env(GAMEBOARD_SIZE, "mini_omok")
