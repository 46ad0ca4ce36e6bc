"""The whole AlphaZero iteration of main.py (main.py:377-407) on one B200 with every stage on the device, timed:
self-play of G concurrent episodes -> records -> 8-fold augmentation into the replay ring -> sampled batches -> PyTorch
train steps -> re-folded weights for the next round."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import trainer

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--sims", type=int, default=400)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--memory", type=int, default=2_000_000)
ap.add_argument("--train-samples", type=int, default=65536)
a = ap.parse_args()
tr = trainer.Trainer(board_size=9, n_mcts=a.sims, n_selfplay=a.games, memory_size=a.memory, batch_size=32, seed=0)
for it in range(a.iters):
    torch.cuda.synchronize(); t0 = time.time()
    plies = tr.self_play(a.games)
    torch.cuda.synchronize(); t1 = time.time()
    line = (f"iter {it}: self-play {a.games} games @{a.sims} sims: {t1 - t0:.1f} s ({a.games / (t1 - t0):.0f} games/s, {plies} plies, "
            f"result {tr.result}, nn mode {tr._engine.nn_precision}), replay {len(tr.rep_memory)} samples")
    if it > 0:
        log = tr.train(max_samples=a.train_samples)
        torch.cuda.synchronize(); t2 = time.time()
        l = np.asarray(log)
        line += (f"; train {len(log)} steps of 32 in {t2 - t1:.1f} s ({len(log) / (t2 - t1):.0f} steps/s): loss {l[:50, 0].mean():.3f} -> "
                 f"{l[-50:, 0].mean():.3f} (v {l[-50:, 1].mean():.3f}, p {l[-50:, 2].mean():.3f})")
    print(line, flush=True)
    tr.reset_iter()
