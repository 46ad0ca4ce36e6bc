"""Play G complete self-play episodes on the device (BASELINE config 2 shape) and report games/s and sanity stats."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi
from alpha_omok_b200.model import seeded_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--board", type=int, default=9)
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--sims", type=int, default=400)
a = ap.parse_args()
eng = _cabi.Engine(board_size=a.board, num_mcts=a.sims, max_games=a.games, seed=7)
eng.load_state_dict(seeded_state_dict(0, 10, 5, 128, a.board))
eng.selfplay_begin(a.games)
t0 = time.time()
st = eng.selfplay_rounds(a.sims)
while st["running"]:
    st = eng.selfplay_rounds(a.sims)
dt = time.time() - t0
moves, n_moves, winners, _ = eng.selfplay_fetch(a.games, with_visits=False)
print(f"{a.games} games {a.board}x{a.board} @{a.sims} sims: {dt:.1f} s  -> {a.games/dt:.1f} games/s, {st['sims']/dt/1e6:.3f} M expansions/s (whole run incl. ragged tail)")
print(f"moves/game mean {n_moves.mean():.1f} min {n_moves.min()} max {n_moves.max()}; winners black {int((winners==1).sum())} white {int((winners==2).sum())} draw {int((winners==3).sum())}; "
      f"errors {st['errors']}; terminal sims {st['terminal_sims']} ({100*st['terminal_sims']/st['sims']:.2f} %)")
