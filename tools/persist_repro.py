"""small repro / sanitizer target for the persistent self-play kernel: n games of 9x9, a few rounds, compared with the
two-kernel path (AO_NO_PERSIST=1 in a second process)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alpha_omok_b200 import _cabi
from alpha_omok_b200.model import seeded_state_dict
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 2
B = int(sys.argv[4]) if len(sys.argv) > 4 else 9
eng = _cabi.Engine(board_size=B, num_mcts=6, max_games=n, n_blocks=nb, seed=3)
eng.load_state_dict(seeded_state_dict(0, nb, 5, 128, B))
eng.selfplay_begin(n, recycle=False)
st = eng.selfplay_rounds(rounds)
while st["running"]:          # every game to its end: complete episodes do not depend on the scheduling
    st = eng.selfplay_rounds(rounds)
moves, n_moves, winners, visits = eng.selfplay_fetch(n)
print("sims", st["sims"], "moves", st["moves"], "checksum", int(visits.astype(np.int64).sum()), int(moves.astype(np.int64).sum()),
      int(winners.astype(np.int64).sum()), flush=True)
eng.close()
