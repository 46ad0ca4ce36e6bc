"""which scheduling variant plays a different episode? (debug companion of
tests/test_gpu_parity.py::test_persistent_kernel_and_deferred_tails_equal_the_plain_round_loop)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alpha_omok_b200 import _cabi
from alpha_omok_b200.model import seeded_state_dict
B, G, sims = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sd = seeded_state_dict(2, 2, 5, 128, B)
runs = {}
for name, env in (("default", {}), ("no_persist", {"AO_NO_PERSIST": "1"}), ("plain", {"AO_NO_PERSIST": "1", "AO_NO_DEFER": "1"}),
                  ("persist_static", {"AO_NO_FREERUN": "1"})):
    for k in ("AO_NO_PERSIST", "AO_NO_DEFER", "AO_NO_FREERUN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = _cabi.Engine(board_size=B, num_mcts=sims, max_games=G, seed=13, n_blocks=2)
    eng.load_state_dict(sd)
    eng.selfplay_begin(G, first_key=50)
    st = eng.selfplay_rounds(64)
    calls = 1
    while st["running"]:
        st = eng.selfplay_rounds(64)
        calls += 1
    runs[name] = eng.selfplay_fetch(G)
    print(name, "calls", calls, st, flush=True)
    eng.close()
ref = runs["plain"]
for name, r in runs.items():
    bad = [g for g in range(G) if not (np.array_equal(r[0][g], ref[0][g]) and np.array_equal(r[3][g], ref[3][g]) and r[2][g] == ref[2][g])]
    print(name, "differing games:", len(bad), bad[:10])
    for g in bad[:2]:
        k = min(int(r[1][g]), int(ref[1][g]))
        t = next((t for t in range(k) if r[0][g][t] != ref[0][g][t] or not np.array_equal(r[3][g][t], ref[3][g][t])), k)
        print("  game", g, "n_moves", int(r[1][g]), int(ref[1][g]), "first differing ply", t, "moves", r[0][g][max(0,t-1):t+2], ref[0][g][max(0,t-1):t+2],
              "visit sums", int(r[3][g][t].sum()) if t < k else None, int(ref[3][g][t].sum()) if t < k else None)
