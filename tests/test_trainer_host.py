"""Host-side widening steps on CPU: the training step against the reference's own arithmetic
(golden fixture generated from the unmodified reference model.py + the loop of main.py:286-305), checkpoint / dataset
file formats (main.py:339-365), gradient averaging over ranks (gloo world 2), ELO bookkeeping (eval_main.py:191-198)."""
import os
import pickle
import socket
import subprocess
import sys
from collections import deque

import numpy as np
import pytest
import torch

from helpers import load
from oracle import pvnet_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _small_net(fx):
    from alpha_omok_b200 import model
    net = model.PVNet(int(fx["n_block"]), 5, 128, int(fx["B"]))
    sd = pvnet_ref.make_state_dict(int(fx["sd_seed"]), int(fx["n_block"]), 5, 128, int(fx["B"]), bn_jitter=True)
    missing = net.load_state_dict(sd, strict=False)
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys)
    return net


def test_train_step_matches_reference_golden():
    """2 epochs x (32, 32, 16) batches: per-step losses and the trained weights equal the reference's (same torch build:
    the arithmetic is op-for-op the same, tolerance only covers thread-count dependent reduction order)."""
    from alpha_omok_b200 import trainer
    fx = load("train_9_small")
    torch.manual_seed(0)
    net = _small_net(fx)
    opt = trainer.make_optimizer(net)
    log = trainer.train_batches(net, opt, torch.from_numpy(fx["states"]), torch.from_numpy(fx["pis"]),
                                torch.from_numpy(fx["zs"]), batch_size=32, n_epochs=2)
    assert len(log) == 6
    np.testing.assert_allclose(np.asarray(log), fx["losses"], rtol=2e-5, atol=2e-6)
    sd = net.state_dict()
    np.testing.assert_allclose(sd["conv1.weight"].numpy(), fx["conv1_weight"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(sd["bn1.running_mean"].numpy(), fx["bn1_running_mean"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(sd["policy_head.policy_fc.bias"].numpy(), fx["policy_fc_bias"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(sd["value_head.value_fc2.weight"].numpy(), fx["value_fc2_weight"], rtol=1e-4, atol=2e-6)
    sums = [float(v.double().abs().sum()) for k, v in sd.items() if not k.endswith("num_batches_tracked")]
    np.testing.assert_allclose(sums, fx["abs_sums"], rtol=1e-5)


def test_checkpoint_and_dataset_formats(tmp_path):
    """file names, the partial-update load idiom for checkpoints without num_batches_tracked, step / start_iter parsed
    from the file name, dataset pickle readable as the reference's deque (main.py:339-365)"""
    from alpha_omok_b200 import trainer
    fx = load("train_9_small")
    net = _small_net(fx)
    path = trainer.save_model(net, 200, 1234, datetime_now="181001", data_dir=str(tmp_path))
    assert os.path.basename(path) == "181001_200_1234_step_model.pickle"
    assert trainer.parse_model_path(path) == (1234, 201)
    sd = torch.load(path)
    assert list(sd.keys()) == list(net.state_dict().keys())
    # a 2018-style checkpoint: no num_batches_tracked keys
    old = {k: v for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    old_path = str(tmp_path / "180927_9400_297233_step_model.pickle")
    torch.save(old, old_path)
    net2 = _small_net(fx)
    with torch.no_grad():
        for p in net2.parameters():
            p.add_(1.0)
    trainer.load_model(net2, old_path)
    for k, v in net.state_dict().items():
        assert torch.equal(v, net2.state_dict()[k]), k
    assert trainer.parse_model_path(old_path) == (297233, 9401)
    # dataset: reference-format deque of (state, pi, z)
    mem = deque([(fx["states"][i].astype(np.float64), fx["pis"][i], float(fx["zs"][i])) for i in range(5)], maxlen=30000)
    dpath = trainer.save_dataset(mem, 200, 1234, datetime_now="181001", data_dir=str(tmp_path))
    assert os.path.basename(dpath) == "181001_200_1234_step_dataset.pickle"
    with open(dpath, "rb") as f:
        back = deque(pickle.load(f), maxlen=30000)     # main.py:362-365
    assert len(back) == 5 and np.array_equal(back[3][0], mem[3][0]) and back[3][2] == mem[3][2]


def test_elo_matches_reference_formula():
    from alpha_omok_b200 import arena
    p, e = arena.elo(1500, 1500, 1, 0)
    assert (p, e) == (1516.0, 1484.0)
    p, e = arena.elo(p, e, 0.5, 0.5)
    assert abs(p - (1516 + 32 * (0.5 - 1 / (1 + 10 ** (-32 / 400))))) < 1e-12 and abs((p + e) - 3000) < 1e-9
    pe, ee, result, winrate = arena.elo_sequence(["player", "draw", "enemy", "player"])
    assert result == {"Player": 2, "Enemy": 1, "Draw": 1} and abs(winrate - 62.5) < 1e-12
    assert abs((pe + ee) - 3000) < 1e-9 and pe > 1500 > ee


DDP_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from alpha_omok_b200 import model, trainer
from oracle import pvnet_ref
from helpers import load
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
torch.set_num_threads(2)
fx = load("train_9_small")
def net_():
    net = model.PVNet(2, 5, 128, 9)
    net.load_state_dict(pvnet_ref.make_state_dict(3, 2, 5, 128, 9, bn_jitter=True), strict=False)
    return net
s, pi, z = torch.from_numpy(fx["states"][:32]), torch.from_numpy(fx["pis"][:32]).float(), torch.from_numpy(fx["zs"][:32]).float()
# sharded step: rank r trains on rows r::world of the global batch, gradients averaged with one all-reduce
net = net_(); opt = trainer.make_optimizer(net)
net.train()
trainer.train_step(net, opt, s[rank::world], pi[rank::world], z[rank::world])
# reference: the same per-rank gradients averaged by hand in one process
ref = net_(); ropt = trainer.make_optimizer(ref); ref.train()
grads = None
for r in range(world):
    tmp = net_(); tmp.train()
    p_b, v_b = tmp(s[r::world])
    loss = (v_b - z[r::world]).pow(2).mean() - (pi[r::world] * p_b.log()).sum(-1).mean()
    loss.backward()
    g = [p.grad.clone() for p in tmp.parameters()]
    grads = g if grads is None else [a + b for a, b in zip(grads, g)]
for p, g in zip(ref.parameters(), grads):
    p.grad = g / world
ropt.step()
for (k, a), b in zip(net.named_parameters(), ref.parameters()):
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (k, float((a - b).abs().max()))
# every rank ends with identical weights
flat = torch.cat([p.detach().flatten() for p in net.parameters()])
lst = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(lst, flat)
assert all(torch.equal(lst[0], t) for t in lst)
dist.barrier(); dist.destroy_process_group()
sys.stdout.write("rank%dok\n" % rank); sys.stdout.flush()
"""


def test_gradient_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "ddp_worker.py"
    script.write_text(DDP_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "rank0ok" in res.stdout and "rank1ok" in res.stdout, res.stdout


def test_seeded_weight_generator_matches_the_oracles():
    """bench.py / tools use the product package's generator for "random-init PVNet (numpy seed s)"; the tests use the
    oracle's: both must be the same function so that fixtures and bench weights agree"""
    from alpha_omok_b200 import model
    for kw in (dict(seed=0), dict(seed=3, n_block=2, bn_jitter=True), dict(seed=1, board_size=15, n_block=1, gain=2.0)):
        a = model.seeded_state_dict(**kw)
        b = pvnet_ref.make_state_dict(**kw)
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].dtype == torch.float32 and torch.equal(a[k], b[k]), k


def test_global_batch_sharding_gives_every_rank_the_same_batches():
    """rows of every global batch of 32 are dealt round-robin to the ranks, so all ranks run the same number
    of optimizer steps (one gradient all-reduce each) with equal row counts - also for a short last batch"""
    from alpha_omok_b200 import trainer
    for n, world in ((64, 2), (34, 2), (96, 4), (40, 8), (8, 8)):
        shards = [trainer.shard_global_batches(n, 32, r, world) for r in range(world)]
        sizes = [tuple(s) for _, s in shards]
        assert len(set(sizes)) == 1, (n, world, sizes)                       # same batches, same row counts
        assert sum(sizes[0]) * world == n
        allrows = sorted(r for rows, _ in shards for r in rows)
        assert allrows == list(range(n))                                    # every sampled row trained exactly once
        for b, cnt in enumerate(sizes[0]):                                  # batch b holds rows of global batch b only
            for rows, s in shards:
                off = sum(s[:b])
                assert all(b * 32 <= r < (b + 1) * 32 for r in rows[off:off + cnt])
    with pytest.raises(ValueError):
        trainer.shard_global_batches(64, 32, 0, 3)
    with pytest.raises(ValueError):
        trainer.shard_global_batches(33, 32, 0, 2)
