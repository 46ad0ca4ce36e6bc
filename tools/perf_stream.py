"""Continuous self-play vs batch-at-a-time: E episodes of 9x9 @400 sims on G game slots (games/s incl. the ragged tail)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi
from alpha_omok_b200.model import seeded_state_dict
ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=4096)
ap.add_argument("--episodes", type=int, default=12288)
ap.add_argument("--sims", type=int, default=400)
a = ap.parse_args()
eng = _cabi.Engine(board_size=9, num_mcts=a.sims, max_games=a.slots, seed=7)
eng.load_state_dict(seeded_state_dict(0, 10, 5, 128, 9))
def drain():
    st = eng.selfplay_rounds(a.sims)
    while st["running"]:
        st = eng.selfplay_rounds(a.sims)
    return st
t0 = time.time()
eng.selfplay_stream_begin(a.episodes, first_key=0)
st = drain()
dt = time.time() - t0
print(f"continuous: {a.episodes} episodes on {a.slots} slots: {dt:.1f} s -> {a.episodes / dt:.1f} games/s, {st['sims'] / dt / 1e6:.3f} M expansions/s, errors {st['errors']}", flush=True)
t0 = time.time(); sims = 0
for b in range(a.episodes // a.slots):
    eng.selfplay_begin(a.slots, first_key=b * a.slots)
    sims += drain()["sims"]
dt = time.time() - t0
print(f"batches:    {a.episodes} episodes as {a.episodes // a.slots} batches of {a.slots}: {dt:.1f} s -> {a.episodes / dt:.1f} games/s, {sims / dt / 1e6:.3f} M expansions/s")
