"""BASELINE config 5 shape: 1024 concurrent 9x9 matches, 800 sims/move, trained ZeroAgent (split precision) vs a
seed-1 random-init ZeroAgent, alternating colours. Reports wall time, expansions/s and the score."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import agents, arena, model
from alpha_omok_b200.model import seeded_state_dict

n, sims = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 800
np.random.seed(0)
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trained_9x9_180927.npz"))
player = model.PVNet(10, 5, 128, 9); player.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=False)
enemy = model.PVNet(10, 5, 128, 9); enemy.load_state_dict(seeded_state_dict(1, 10, 5, 128, 9), strict=False)
# nn_precision="auto": the trained player's engine picks the hi/lo split tower, the random-init enemy's the fp16 one
t0 = time.time()
res = arena.play_matches(player, enemy, n_matches=n, num_mcts=sims, seed=1)
dt = time.time() - t0
plies = sum(res["plies"])
print(f"{n} matches @{sims} sims: {dt:.1f} s, {plies} plies -> {plies*sims/dt/1e6:.3f} M expansions/s; "
      f"trained wins {res['player_win']}, random-init wins {res['enemy_win']}, draws {res['draw']}, unfinished {res['unfinished']}")
