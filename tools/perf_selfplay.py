"""Quick device-side throughput probe: G concurrent self-play games, lock-step rounds, per-kernel times and the tower's
cycle counters.  The counters exist only in the probe library, which this tool therefore selects (AO_USE_PROBE_LIB)."""
import argparse
import os
import sys
import time

os.environ.setdefault("AO_USE_PROBE_LIB", "1")

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi  # noqa: E402
from alpha_omok_b200.model import seeded_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--board", type=int, default=9)
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--sims", type=int, default=400)
ap.add_argument("--rounds", type=int, default=100)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--node-cap", type=int, default=0)
ap.add_argument("--x3", action="store_true")
ap.add_argument("--single-cta", action="store_true")
ap.add_argument("--mode", type=int, default=-1, help="AO_NN_* (overrides --x3 / --single-cta)")
a = ap.parse_args()

eng = _cabi.Engine(board_size=a.board, num_mcts=a.sims, max_games=a.games, seed=1, node_cap=a.node_cap, nn_precision=a.mode if a.mode >= 0 else (1 if a.x3 else (2 if a.single_cta else 0)))
eng.load_state_dict(seeded_state_dict(0, 10, 5, 128, a.board))
eng.selfplay_begin(a.games)
prev = eng.selfplay_rounds(20)
eng.tower_debug(True)
for rep in range(a.reps):
    t0 = time.time()
    st = eng.selfplay_rounds_timed(a.rounds)
    dt = time.time() - t0
    ds, de = st["sims"] - prev["sims"], st["nn_evals"] - prev["nn_evals"]
    print(f"rep {rep}: {a.rounds} rounds wall {dt*1e3:.1f} ms  sims {ds} ({ds/dt/1e6:.3f} M/s)  evals {de}  "
          f"tree {st['tree_ms']:.1f} ms  tower {st['tower_ms']:.1f} ms  "
          f"tower TFLOP/s {de*(478.8e6 if a.board==9 else 1330.1e6)/st['tower_ms']/1e9:.1f}  running {st['running']} "
          f"moves {st['moves']} err {st['errors']}", flush=True)
    prev = st
d = eng.tower_debug(False)
if d[3]:
    print("tower CTA0 cycles/launch: mma_total %.0f  wait_epilogue %.0f  wait_weights %.0f | epi_total %.0f  wait_acc %.0f  heads %.0f"
          % (d[0] / d[3], d[1] / d[3], d[2] / d[3], d[4] / d[3], d[5] / d[3], d[6] / d[3]))
