"""BASELINE config 1 on the device: ONE 9x9 self-play game, 40 sims/move, random-init PVNet, driven move by move through
the reference-shaped facades (agents.ZeroAgent.get_pi / utils.get_action / env.step) exactly like main.py's loop - the
latency-bound end of the path (batch of one leaf per network call). The reference does ~75 sims/s/core here."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import agents, model, utils
sims = int(sys.argv[1]) if len(sys.argv) > 1 else 40
if len(sys.argv) > 2 and sys.argv[2] == "15":    # BASELINE config 3's board, one game
    from alpha_omok_b200.env import env_regular as game
else:
    from alpha_omok_b200.env import env_small as game
for rep in range(3):
    np.random.seed(rep)
    B = game.Return_BoardParams()[0]
    Agent = agents.ZeroAgent(B, sims, 5, noise=True)
    Agent.model = model.PVNet(10, 5, 128, B)
    if "trained" in sys.argv:   # the reference's shipped 9x9 checkpoint: the facade picks the hi/lo split tower for it
        import torch
        z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "trained_9x9_180927.npz"))
        Agent.model.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=False)
    else:
        Agent.model.load_state_dict(model.seeded_state_dict(0, 10, 5, 128, B), strict=False)
    Agent.model.eval()
    env = game.GameState("text")
    root_id, win_index, t, n_sims = (0,), 0, 0, 0
    Agent.get_pi(root_id, 1); Agent.reset()          # engine creation + weight upload outside the timed region
    t0 = time.time()
    while win_index == 0:
        pi = Agent.get_pi(root_id, 1 if t < 6 else 0)
        n_sims += sims + (1 if Agent.is_real_root else 0)
        action, action_index = utils.get_action(pi)
        root_id += (int(action_index),)
        _, _, win_index, _, _ = env.step(action)
        t += 1
    dt = time.time() - t0
    print(f"game {rep}: {t} moves, {n_sims} sims @{sims}/move in {dt * 1e3:.0f} ms -> {n_sims / dt:.0f} sims/s, {dt / t * 1e3:.1f} ms per move (winner {win_index})", flush=True)
