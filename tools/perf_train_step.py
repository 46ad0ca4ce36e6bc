"""Training-step throughput at the reference's batch size (main.py:253-336, BATCH_SIZE = 32): eager PyTorch vs the
CUDA-graph replay of the same step (trainer.GraphedTrainStep)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alpha_omok_b200 import model, trainer

dev = torch.device("cuda", 0)
B, A, N = 9, 81, 32 * 512
rs = np.random.RandomState(0)
s = torch.from_numpy((rs.rand(N, 5, B, B) < 0.3).astype(np.float32)).to(dev)
pi = torch.softmax(torch.from_numpy(rs.randn(N, A).astype(np.float32)), -1).to(dev)
z = torch.from_numpy(rs.choice([-1.0, 0.0, 1.0], N).astype(np.float32)).to(dev)
for graphed in (False, True):
    torch.manual_seed(0)
    net = model.PVNet(10, 5, 128, B).to(dev)
    opt = trainer.make_optimizer(net)
    g = trainer.GraphedTrainStep(net, opt, 32, B) if graphed else None
    trainer.train_batches(net, opt, s[:32 * 16], pi[:32 * 16], z[:32 * 16], graphed=g)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    log = trainer.train_batches(net, opt, s, pi, z, graphed=g)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s: %d steps of 32 in %.3f s = %.0f steps/s (%.2f ms/step), loss %.4f -> %.4f"
          % ("graph" if graphed else "eager", len(log), dt, len(log) / dt, 1e3 * dt / len(log), log[0][0], log[-1][0]))
