"""Bandwidth of the device augmentation kernel on a full-size round (4096 finished 9x9 games)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi, replay
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = _cabi.Engine(board_size=9, num_mcts=32, max_games=G, seed=3, eval_mode=_cabi.AO_EVAL_SYNTH)
eng.selfplay_begin(G)
st = eng.selfplay_rounds(1)
while st["running"]:
    st = eng.selfplay_rounds(1)
slab = replay.device_records(eng, G).clone()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    states, pi, z = replay.augmented_tensors(slab, 9)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out_bytes = states.numel() * 4 + pi.numel() * 4 + z.numel() * 4
    in_bytes = slab.numel()
    print(f"rep {rep}: {states.shape[0]} samples from {G} games: {ms:.2f} ms incl. allocation + count pass, "
          f"{out_bytes/1e9:.2f} GB written + {in_bytes/1e9:.2f} GB read -> {(out_bytes+in_bytes)/ms/1e6:.0f} GB/s")
    del states, pi, z

# device replay ring (deque(maxlen) semantics) + random.sample gather
import random
buf = replay.DeviceReplayBuffer(9, maxlen=2_000_000)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = buf.extend_records(slab)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"ring extend rep {rep}: {n} samples, len {len(buf)}: {ms:.2f} ms -> {(n*1948+slab.numel())/ms/1e6:.0f} GB/s")
k = len(buf) // 2
idx = torch.as_tensor(buf.sample_indices(k), dtype=torch.int64, device='cuda')  # upload once: the kernel is what is timed
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    s, p, z = buf.gather(idx)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"gather rep {rep}: k={k} samples ({k*1948/1e9:.2f} GB read + written each): {ms:.2f} ms incl. output allocation "
          f"-> {2*k*1948/ms/1e6:.0f} GB/s")
    del s, p, z
