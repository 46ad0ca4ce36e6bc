"""15x15 Omok environment - drop-in for 2_AlphaOmok/env/env_regular.py (GameState.step, Return_BoardParams, ...)."""
from ._omok_env import make_env

GAMEBOARD_SIZE = 15
WIN_STONES = 5
GameState, ReturnName, Return_Num_Action, Return_BoardParams = make_env(GAMEBOARD_SIZE, "regular_omok")
