"""Where the time of a ONE-game search goes (probe build, AO_USE_PROBE_LIB=1): cycle counters of CTA 0's MMA issuer and
epilogue thread 0 over a 400-simulation search through Engine.search (ao_search)."""
import os, sys, time
os.environ["AO_USE_PROBE_LIB"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alpha_omok_b200 import _cabi, model

sims = int(sys.argv[1]) if len(sys.argv) > 1 else 400
B = 9
eng = _cabi.Engine(board_size=B, num_mcts=sims, max_games=1, noise=True)
eng.load_state_dict(model.seeded_state_dict(0, 10, 5, 128, B))
eng.search([0], [(0,)])
eng.tower_debug(True)
t0 = time.time()
for rep in range(5):
    eng.search([0], [(0, 40 + rep)])
dt = (time.time() - t0) / 5
d = eng.tower_debug(False)
n = 5 * (sims + 1)
print("search of %d sims: %.2f ms wall = %.1f us/sim" % (sims, dt * 1e3, dt * 1e6 / (sims + 1)))
if os.environ.get("AO_NO_SOLO") is None:
    k = max(d[0], 1)
    print("solo kernel (cluster of four), cycles per simulation over %d passes: tower %.0f (of which the epilogue thread waits for the MMAs %.0f), "
          "heads %.0f, tree step %.0f, request turnaround %.0f; issuer 0: operands ready -> layer accumulated everywhere %.0f, waits for its operands (after the accumulators are free) %.0f"
          % (d[0], d[1] / k, d[5] / k, d[2] / k, d[3] / k, d[4] / k, d[6] / k, d[7] / k))
    sys.exit(0)
print("per simulation, cycles: mma warp total %.0f, waits for operands (epilogue / heads / tree / request) %.0f, waits for weights %.0f; "
      "epilogue thread total %.0f, waits for accumulators %.0f; launches %d"
      % (d[0] / n, d[1] / n, d[2] / n, d[4] / n, d[5] / n, d[3]))
