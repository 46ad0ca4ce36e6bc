"""Kernel mix of the batch-32 training step (main.py:286-305) on the device: torch.profiler over 20 graph replays of
trainer.GraphedTrainStep, kernels grouped by name."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from alpha_omok_b200 import model, trainer

dev = torch.device("cuda", 0)
B, A = 9, 81
rs = np.random.RandomState(0)
s = torch.from_numpy((rs.rand(32, 5, B, B) < 0.3).astype(np.float32)).to(dev)
pi = torch.softmax(torch.from_numpy(rs.randn(32, A).astype(np.float32)), -1).to(dev)
z = torch.from_numpy(rs.choice([-1.0, 0.0, 1.0], 32).astype(np.float32)).to(dev)
torch.manual_seed(0)
net = model.PVNet(10, 5, 128, B).to(dev)
opt = trainer.make_optimizer(net)
kind = sys.argv[1] if len(sys.argv) > 1 else "graph"
g = trainer.GraphedTrainStep(net, opt, 32, B) if kind == "graph" else None
def step():
    if g is not None:
        g.step(s, pi, z)
    else:
        trainer.train_step(net, opt, s, pi, z)
for _ in range(5):
    step()
torch.cuda.synchronize()
N = 20
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name][0] += 1
        agg[e.name][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print("%s step: %d kernels per step, %.1f us of kernel time per step" % (kind, sum(v[0] for v in agg.values()) / N, tot / N))
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%6.1f us/step %5.1f%% x%-4.0f %s" % (t / N, 100 * t / tot, n / N, name[:110]))
