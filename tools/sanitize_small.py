"""Small end-to-end run for compute-sanitizer (memcheck / synccheck / racecheck): tower forward in every operand mode
(9x9 and 15x15), a few self-play rounds with the real tower, the replay ring kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi
from alpha_omok_b200.model import seeded_state_dict

modes = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 2, 3]
for B in (9, 15):
    A = B * B
    sd = seeded_state_dict(0, 2, 5, 128, B)
    rs = np.random.RandomState(1)
    x = np.zeros((7, 5, B, B), np.float32)
    x[:, 2:4] = (rs.rand(7, 2, B, B) < 0.2)
    x[:, 0:2] = x[:, 2:4]
    ref = None
    for mode in modes:
        eng = _cabi.Engine(board_size=B, num_mcts=8, max_games=8, n_blocks=2, nn_precision=mode)
        eng.load_state_dict(sd)
        p, v = eng.nn_forward(x)
        assert np.isfinite(p).all() and abs(p.sum(1) - 1).max() < 1e-4
        if ref is not None:
            assert abs(p - ref).max() < 1e-4
        ref = p
        eng.selfplay_begin(8)
        st = eng.selfplay_rounds(30)
        assert st["errors"] == 0 and st["sims"] > 0
        eng.close()
        print("board", B, "mode", mode, "ok", flush=True)
import torch
from alpha_omok_b200 import replay
eng = _cabi.Engine(board_size=9, num_mcts=8, max_games=4, seed=6, eval_mode=_cabi.AO_EVAL_SYNTH)
eng.selfplay_begin(4)
st = eng.selfplay_rounds(1)
while st["running"]:
    st = eng.selfplay_rounds(1)
slab = replay.device_records(eng, 4).clone()
buf = replay.DeviceReplayBuffer(9, maxlen=500)
for _ in range(3):
    buf.extend_records(slab)
s, p, z = buf.sample(64)
torch.cuda.synchronize()
print("replay ring ok", len(buf), flush=True)
# round 2: the arena match loop (two weight sets, both towers per round), RandomAgent / PUCT / UCT sides, rollout search
sd0, sd1 = seeded_state_dict(0, 2, 5, 128, 9), seeded_state_dict(1, 2, 5, 128, 9)
eng = _cabi.Engine(board_size=9, num_mcts=12, max_games=8, n_blocks=2, noise=False, seed=3)
eng.load_state_dict(sd0, which=0)
eng.load_state_dict(sd1, which=1)
eng.set_nn_precision(_cabi.AO_NN_FP16X3, which=0)
for kind in ("zero", "random", "puct", "uct"):
    eng.arena_begin(4, first_key=0, matches_per_slot=2, enemy_kind=kind)
    st = eng.selfplay_rounds(40)
    n = 0
    while st["running"] and n < 60:
        st = eng.selfplay_rounds(40)
        n += 1
    assert st["errors"] == 0 and st["games_finished"] >= 1, st
    print("arena enemy", kind, "ok", st["games_finished"], "matches", flush=True)
for kind in ("puct", "uct"):
    vis, w = eng.rollout_search(kind, [0, 1], [(0,), (0, 40, 41)], 24)
    assert vis.sum(axis=1).tolist() == [24, 24]
    print("rollout", kind, "ok", flush=True)
eng.close()
# round 2, second half: the cluster-of-four kernel (tower_solo.cu) is what the 8-game self-play runs above and the
# searches below use; 40 games are beyond its range and run the CTA-pair persistent kernel (static game -> pass map)
for B in (9, 15):
    eng = _cabi.Engine(board_size=B, num_mcts=6, max_games=40, n_blocks=2, seed=4)
    eng.load_state_dict(seeded_state_dict(0, 2, 5, 128, B))
    vis, pri, real = eng.search([0, 1, 2], [(0,), (0, 3), (0, 3, 4)])
    assert all(real) and (vis.sum(axis=1) >= 5).all()
    vis, pri, real = eng.search([0], [(0, int(np.argmax(vis[0])))])
    assert vis.sum() >= 5
    eng.selfplay_begin(40)
    st = eng.selfplay_rounds(12)
    assert st["errors"] == 0 and st["sims"] > 0
    eng.close()
    print("board", B, "solo search + 40-game persistent kernel ok", flush=True)
