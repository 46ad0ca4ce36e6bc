"""One BASELINE configuration of the hot path for a fixed number of lock-step rounds - the command ncu wraps for the
per-kernel captures under profiles/ (tools/profile_r02.sh) and a quick throughput probe without bench.py's other legs.
    python tools/perf_leg.py selfplay9|selfplay15|trained9|arena|rollout [rounds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alpha_omok_b200 import _cabi
from alpha_omok_b200.model import seeded_state_dict

leg = sys.argv[1] if len(sys.argv) > 1 else "selfplay9"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 200


def trained():
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "trained_9x9_180927.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


if leg == "rollout":
    eng = _cabi.Engine(board_size=9, num_mcts=800, max_games=1024, noise=False, eval_mode=_cabi.AO_EVAL_SYNTH)
    ids = list(range(1024))
    eng.games_reset(ids, keys=ids)
    for kind in ("puct", "uct"):
        t0 = time.perf_counter()
        vis, w = eng.rollout_search(kind, ids, [(0,)] * 1024, 800)
        dt = time.perf_counter() - t0
        print("%s: 1024 searches x 801 play-out simulations in %.3f s = %.2f M simulations/s" % (kind, dt, 1024 * 801 / dt / 1e6))
    sys.exit(0)
if leg == "arena":
    eng = _cabi.Engine(board_size=9, num_mcts=800, max_games=2048, noise=False, seed=1)
    eng.load_state_dict(trained(), which=0)
    eng.load_state_dict(seeded_state_dict(1, 10, 5, 128, 9), which=1)
    eng.choose_nn_precision(which=0), eng.choose_nn_precision(which=1)
    eng.arena_begin(1024, first_key=0, matches_per_slot=1 << 20, keep_records=False)
else:
    B = 15 if leg == "selfplay15" else 9
    eng = _cabi.Engine(board_size=B, num_mcts=400, max_games=4096, seed=1)
    eng.load_state_dict(trained() if leg == "trained9" else seeded_state_dict(0, 10, 5, 128, B))
    if leg == "trained9":
        eng.choose_nn_precision()
    eng.selfplay_begin(4096, recycle=True)
st0 = eng.selfplay_rounds(32)
t0 = time.perf_counter()
st = eng.selfplay_rounds_timed(rounds)
dt = time.perf_counter() - t0
print("%s: %d rounds, %.0f expansions/s (wall), tower %.3f ms/round, tree %.3f ms/round" %
      (leg, rounds, (st["sims"] - st0["sims"]) / dt, st["tower_ms"] / rounds, st["tree_ms"] / rounds))
eng.close()
