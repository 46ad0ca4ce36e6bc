"""Drop-in for the training half of 2_AlphaOmok/main.py: `train` (main.py:253-336),
`save_model / save_dataset / load_data` (main.py:339-365) and the iteration loop (main.py:377-407), wired to the
device-resident self-play engine and replay ring.

What runs where:
* self-play           -> csrc/tree.cu + csrc/tower_stag.cu (`ao_selfplay_*`), all games of an iteration concurrently;
* replay buffer       -> csrc/augment.cu (`DeviceReplayBuffer`: deque(maxlen) + augment_dataset + random.sample on the GPU);
* training step       -> PyTorch autograd on the same `model.PVNet` module, exactly the reference's arithmetic
                         (train-mode BatchNorm, loss = MSE(v, z) - sum(pi * log p), Adam lr 2e-4 eps 1e-6).  This is the
                         part of the reference that already is library code (torch.nn / torch.optim) and stays so;
* multi-GPU           -> one process per GPU: games and replay records shard by rank (`replay.allgather_records`),
                         gradients are averaged with one flat all-reduce per step over NCCL (gloo on CPU).
The per-batch arithmetic is pinned against the unmodified reference (tests/golden/train_9_small.npz,
tests/test_trainer_host.py::test_train_step_matches_reference_golden).
"""
from __future__ import annotations

import logging
import os
import pickle
import random
from datetime import datetime

import numpy as np
import torch
import torch.distributed as dist
from torch import optim

from . import model as model_mod, replay


def make_optimizer(net, lr=2e-4, l2=0.0):
    """main.py:85: optim.Adam(Agent.model.parameters(), lr=LR, weight_decay=L2, eps=1e-6)"""
    return optim.Adam(net.parameters(), lr=lr, weight_decay=l2, eps=1e-6)


def _allreduce_mean_grads(params, group=None):
    """Average the gradients over all ranks with ONE collective (flat bucket)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)


def train_step(net, optimizer, s_batch, pi_batch, z_batch, group=None):
    """One optimizer step of main.py:286-305 on float32 tensors already on the model's device.
    Returns (loss, v_loss, p_loss) as Python floats."""
    p_batch, v_batch = net(s_batch)
    v_loss = (v_batch - z_batch).pow(2).mean()
    p_loss = -(pi_batch * p_batch.log()).sum(dim=-1).mean()
    loss = v_loss + p_loss
    optimizer.zero_grad()
    loss.backward()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        _allreduce_mean_grads([p for g in optimizer.param_groups for p in g["params"]], group)
    optimizer.step()
    return loss.item(), v_loss.item(), p_loss.item()


class GraphedTrainStep:
    """main.py:286-305 (forward in train-mode BN, loss, backward, Adam step) for ONE fixed batch size, captured once
    into a CUDA graph and replayed: at batch 32 the eager step is ~400 tiny kernels and entirely launch-bound
    (5.7 ms/step on a B200 in round 1); replaying the same kernels from a graph removes the host from the loop.
    The arithmetic is the eager step's, kernel for kernel; only Adam's bias correction is evaluated on the device
    (`capturable=True`) instead of in Python floats.  Losses stay on the device until `losses()` is called.

    Capture needs warm-up iterations on real memory; they run on the static buffers and every bit of state they touch
    (parameters, BatchNorm statistics, Adam moments and step counters) is put back afterwards, so a graphed run performs
    exactly the optimizer steps an eager run performs."""

    def __init__(self, net, optimizer, batch_size, board_size, inplanes=5, capacity=4096, channels_last=True, fused_adam=True,
                 cudnn_benchmark=True):
        dev = next(net.parameters()).device
        assert dev.type == "cuda", "CUDA graphs need the model on a GPU"
        A = board_size * board_size
        self.net, self.opt, self.bs = net, optimizer, batch_size
        # (1) NHWC activations and weights: cuDNN's tensor-core convolutions are NHWC kernels; with NCHW tensors every
        # convolution call is wrapped in layout-conversion kernels (160 of the step's 582 kernels, 0.43 of its 2.0 ms).
        # Values, shapes and state_dict contents are unchanged - only the strides of the 4-D parameters.
        if channels_last:
            net.to(memory_format=torch.channels_last)
            for st in optimizer.state.values():
                for k, v in st.items():
                    if torch.is_tensor(v) and v.dim() == 4:
                        st[k] = v.contiguous(memory_format=torch.channels_last)
        # (2) one fused multi-tensor Adam kernel instead of the for-each implementation (its bias-correction terms alone
        # are two tiny kernels per parameter tensor when captured: 144 launches, 0.23 ms)
        for g in optimizer.param_groups:
            g["capturable"] = True
            if fused_adam:
                g["fused"], g["foreach"] = True, False
        for st in optimizer.state.values():
            if torch.is_tensor(st.get("step")) and st["step"].device != dev:
                st["step"] = st["step"].to(dev)
        self.s = torch.zeros((batch_size, inplanes, board_size, board_size), device=dev)
        if channels_last:
            self.s = self.s.contiguous(memory_format=torch.channels_last)
        self.pi = torch.full((batch_size, A), 1.0 / A, device=dev)
        self.z = torch.zeros((batch_size,), device=dev)
        self.out = torch.zeros(3, device=dev)
        self.log = torch.zeros((capacity, 3), device=dev)
        self.n = 0
        # (3) BatchNorm's num_batches_tracked += 1 is one tiny kernel per BN layer and step (momentum is fixed, so the
        # counter never enters the arithmetic): kept out of the graph, added in bulk by sync_counters() / losses()
        self._bn = [m for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)
                    and m.num_batches_tracked is not None]
        self._bn_counters = [m.num_batches_tracked for m in self._bn]
        self._pending_steps = 0
        for m in self._bn:
            m.num_batches_tracked = None
        model_snap = {k: v.detach().clone() for k, v in net.state_dict().items()}
        opt_snap = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                    for p, st in optimizer.state.items()}
        net.train()
        # (4) let cuDNN time its algorithms for these (static) shapes during the warm-up iterations: its heuristic picks
        # a 40 us non-tensor-core weight-gradient kernel for NHWC 3x3 convolutions at batch 32 (0.83 ms per step)
        bench_flag = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = cudnn_benchmark
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    optimizer.zero_grad(set_to_none=True)
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.graph):
                self._body()
        finally:
            torch.backends.cudnn.benchmark = bench_flag
        # undo the warm-up / capture iterations in place (the graph holds the addresses of all of these tensors)
        with torch.no_grad():
            for k, v in net.state_dict().items():
                v.copy_(model_snap[k])
            for p, st in optimizer.state.items():
                old = opt_snap.get(id(p))
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()
        for m, c in zip(self._bn, self._bn_counters):
            m.num_batches_tracked = c
        torch.cuda.synchronize(dev)

    def _body(self):
        p_batch, v_batch = self.net(self.s)
        v_loss = (v_batch - self.z).pow(2).mean()
        p_loss = -(self.pi * p_batch.log()).sum(dim=-1).mean()
        loss = v_loss + p_loss
        loss.backward()
        self.opt.step()
        self.out.copy_(torch.stack((loss.detach(), v_loss.detach(), p_loss.detach())))

    def step(self, s_batch, pi_batch, z_batch):
        assert s_batch.shape[0] == self.bs
        if self.n == self.log.shape[0]:
            self.log = torch.cat((self.log, torch.zeros_like(self.log)))
        self.s.copy_(s_batch)
        self.pi.copy_(pi_batch)
        self.z.copy_(z_batch)
        self.graph.replay()
        self.log[self.n].copy_(self.out)
        self.n += 1
        self._pending_steps += 1

    def sync_counters(self):
        """BatchNorm.num_batches_tracked of every layer += the steps replayed since the last call"""
        if self._pending_steps and self._bn_counters:
            torch._foreach_add_(self._bn_counters, self._pending_steps)
        self._pending_steps = 0

    def losses(self):
        """(loss, v_loss, p_loss) of every step since the last call, as Python floats (one device sync)"""
        self.sync_counters()
        out = [tuple(r) for r in self.log[:self.n].cpu().tolist()]
        self.n = 0
        return out


def train_batches(net, optimizer, states, pis, zs, batch_size=32, n_epochs=1, group=None, batch_sizes=None,
                  graphed=None):
    """main.py:253-336 for an already sampled train_memory given as tensors [N,5,B,B], [N,A], [N]:
    `DataLoader(train_memory, batch_size=BATCH_SIZE, shuffle=False)` = consecutive slices, a short last batch kept.
    `batch_sizes` (optional) gives the row count of every consecutive batch explicitly (multi-rank shards).
    `graphed`: a GraphedTrainStep for full batches (single-rank CUDA runs); other batch sizes run eagerly."""
    net.train()
    dev = next(net.parameters()).device
    states, pis, zs = states.to(dev).float(), pis.to(dev).float(), zs.to(dev).float()
    if batch_sizes is None:
        n = states.shape[0]
        batch_sizes = [min(batch_size, n - i) for i in range(0, n, batch_size)]
    assert sum(batch_sizes) == states.shape[0]
    log = []
    for _ in range(n_epochs):
        i = 0
        for b in batch_sizes:
            if graphed is not None and b == graphed.bs:
                graphed.step(states[i:i + b], pis[i:i + b], zs[i:i + b])
                log.append(None)  # filled in below, in order
            else:
                log.append(train_step(net, optimizer, states[i:i + b], pis[i:i + b], zs[i:i + b], group))
            i += b
    if graphed is not None:
        it = iter(graphed.losses())
        log = [x if x is not None else next(it) for x in log]
    return log


def shard_global_batches(n_rows, batch_size, rank, world):
    """Rows of the sampled train_memory this rank trains on, batch by batch: global batch i = rows
    [i*BS, (i+1)*BS) exactly as the reference's DataLoader cuts them (main.py:266-269), of which rank r takes every
    world-th row.  Every rank gets the SAME number of batches (one gradient all-reduce per batch on every rank) and the
    same number of rows in each of them, which needs BATCH_SIZE % world == 0 and n_rows % world == 0 (callers trim).
    Returns (row index list, per-batch row counts)."""
    if batch_size % world != 0:
        raise ValueError("BATCH_SIZE = %d must be a multiple of the world size %d" % (batch_size, world))
    if n_rows % world != 0:
        raise ValueError("sample count %d must be a multiple of the world size %d" % (n_rows, world))
    rows, sizes = [], []
    for i in range(0, n_rows, batch_size):
        mine = list(range(i, min(i + batch_size, n_rows)))[rank::world]
        rows += mine
        sizes.append(len(mine))
    return rows, sizes


def model_file(datetime_now, n_iter, step, data_dir="data"):
    return os.path.join(data_dir, "{}_{}_{}_step_model.pickle".format(datetime_now, n_iter, step))


def dataset_file(datetime_now, n_iter, step, data_dir="data"):
    return os.path.join(data_dir, "{}_{}_{}_step_dataset.pickle".format(datetime_now, n_iter, step))


def save_model(net, n_iter, step, datetime_now=None, data_dir="data"):
    """main.py:339-342: torch.save(state_dict) as data/{yymmdd}_{iter}_{step}_step_model.pickle"""
    path = model_file(datetime_now or datetime.now().strftime("%y%m%d"), n_iter, step, data_dir)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save(net.state_dict(), path)
    return path


def save_dataset(memory, n_iter, step, datetime_now=None, data_dir="data"):
    """main.py:345-348: pickle of the replay memory (a DeviceReplayBuffer is written as the reference's deque of
    (state, pi, z) tuples, so the file loads in the reference and vice versa)."""
    path = dataset_file(datetime_now or datetime.now().strftime("%y%m%d"), n_iter, step, data_dir)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    from collections import deque
    if isinstance(memory, replay.DeviceReplayBuffer):
        memory = deque(memory.to_list(), maxlen=memory.maxlen)
    with open(path, "wb") as f:
        pickle.dump(memory, f, pickle.HIGHEST_PROTOCOL)
    return path


def parse_model_path(model_path):
    """main.py:359-360: step and start_iter come from the file NAME ({date}_{iter}_{step}_step_model.pickle)."""
    name = os.path.basename(model_path)
    return int(name.split("_")[2]), int(name.split("_")[1]) + 1


def load_model(net, model_path):
    """main.py:356-358 / eval_main.py:95-101: partial update, so 2018 checkpoints without `num_batches_tracked` load."""
    state = net.state_dict()
    state.update(torch.load(model_path, map_location="cpu"))
    net.load_state_dict(state)
    return net


class Trainer:
    """The module-level state and functions of main.py as an object (names follow main.py:26-85)."""

    def __init__(self, board_size=9, n_mcts=400, tau_thres=6, seed=0, n_blocks=10, in_planes=5, out_planes=128,
                 n_selfplay=100, memory_size=30000, n_epochs=1, batch_size=32, lr=2e-4, l2=0.0, device=None,
                 data_dir="data", group=None, max_slots=4096, nn_precision="auto", graph_train=True):
        self.BOARD_SIZE, self.N_MCTS, self.TAU_THRES, self.SEED = board_size, n_mcts, tau_thres, seed
        self.N_BLOCKS, self.IN_PLANES, self.OUT_PLANES = n_blocks, in_planes, out_planes
        self.N_SELFPLAY, self.MEMORY_SIZE, self.N_EPOCHS, self.BATCH_SIZE = n_selfplay, memory_size, n_epochs, batch_size
        self.data_dir, self.group = data_dir, group
        self.max_slots = max_slots  # concurrent games resident in HBM; more episodes than that run in continuous mode
        # "auto": keep policy / value within 1e-4 of fp32 for whatever the weights are (trained nets then run the 3-MMA
        # split mode, ~2.7x slower); 0 (AO_NN_FP16): always single-pass fp16, ~2e-3 on trained nets - the usual
        # AlphaZero trade-off for self-play data, but outside this repository's parity contract
        self.nn_precision = nn_precision
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        random.seed(seed)           # main.py:59-63
        np.random.seed(seed)
        torch.manual_seed(seed)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.model = model_mod.PVNet(n_blocks, in_planes, out_planes, board_size).to(self.device)
        self.optimizer = make_optimizer(self.model, lr, l2)
        self.rep_memory = replay.DeviceReplayBuffer(board_size, maxlen=memory_size, tau_thres=tau_thres, device=self.device)
        self.step, self.start_iter, self.total_epoch = 0, 0, 0
        self.result = {"Black": 0, "White": 0, "Draw": 0}
        self._engine = None
        self._episodes = 0
        self.graph_train = graph_train  # replay the batch-32 training step from a CUDA graph (single rank)
        self._graphed = None

    # ---------------------------------------------------------------- self-play (main.py:122-250)
    def self_play(self, n_selfplay=None):
        """All `n_selfplay` episodes of this rank concurrently on the device; records of every rank are all-gathered
        and appended (8-fold augmented) to every rank's replay ring. Returns len(cur_memory) over all ranks."""
        from . import _cabi
        n = n_selfplay or self.N_SELFPLAY
        self.model.eval()
        slots = min(n, self.max_slots)
        if self._engine is None or self._engine.G < slots:
            if self._engine is not None:
                self._engine.close()
            self._engine = _cabi.Engine(board_size=self.BOARD_SIZE, num_mcts=self.N_MCTS, max_games=slots, noise=True,
                                        tau_thres=self.TAU_THRES, n_blocks=self.N_BLOCKS, inplanes=self.IN_PLANES,
                                        seed=self.SEED, device=self.device.index or 0)
        eng = self._engine
        eng.load_state_dict(self.model.state_dict())
        if self.nn_precision == "auto":
            eng.choose_nn_precision()
        else:
            eng.set_nn_precision(int(self.nn_precision))
        # per-game decision-stream keys are global episode numbers: independent of how games are sharded over ranks
        first_key = self._episodes + self.rank * n
        if n > slots:   # a slot that finishes its episode takes the next unplayed key (ao_selfplay_stream_begin)
            eng.selfplay_stream_begin(n, n_slots=slots, first_key=first_key)
        else:
            eng.selfplay_begin(n, first_key=first_key)
        self._episodes += n * self.world
        st = eng.selfplay_rounds(256)
        while st["running"]:
            st = eng.selfplay_rounds(256)
        if st["errors"]:
            raise _cabi.AoError("%d game tree(s) overflowed their arena; raise node_cap" % st["errors"])
        local = replay.device_stream_records(eng) if n > slots else replay.device_records(eng, n)
        slab = replay.allgather_records(local, self.group)
        winners = slab[:, 2]
        for name, code in (("Black", 1), ("White", 2), ("Draw", 3)):
            self.result[name] += int((winners == code).sum())
        self.rep_memory.extend_records(slab)
        return self.rep_memory.cur_len

    # ---------------------------------------------------------------- training (main.py:253-336)
    def train(self, n_epochs=None, max_samples=None, strict_reference=False):
        """main.py:253-336.  The reference trains on BATCH_SIZE * len(cur_memory) samples (main.py:263-264) and
        `random.sample` raises when that exceeds the replay memory - which it does as soon as an iteration plays
        thousands of concurrent episodes instead of one.  `strict_reference=True` keeps that behaviour; by default the
        count is clamped to the memory size (and to `max_samples` if given)."""
        k = self.BATCH_SIZE * self.rep_memory.cur_len          # BATCH_SIZE * len(cur_memory), main.py:263-264
        if not strict_reference:
            k = min(k, len(self.rep_memory), max_samples if max_samples is not None else k)
        if self.world > 1:
            k -= k % self.world                                # every global batch then splits evenly over the ranks
        idx = self.rep_memory.sample_indices(k)                # every rank draws the same indices (same seed) ...
        rows, sizes = shard_global_batches(len(idx), self.BATCH_SIZE, self.rank, self.world)
        idx = [idx[r] for r in rows]                           # ... and trains on its rows of every global batch
        if self.world > 1:
            # belt and braces: a rank with a different batch count would dead-lock in the gradient all-reduce
            cnt = torch.tensor([len(sizes), -len(sizes)],
                               device=self.device if dist.get_backend(self.group) == "nccl" else "cpu")
            dist.all_reduce(cnt, op=dist.ReduceOp.MAX, group=self.group)
            if cnt[0].item() != -cnt[1].item():
                raise RuntimeError("ranks disagree on the number of training batches")
        s, pi, z = self.rep_memory.gather(idx)
        if self.graph_train and self.world == 1 and self.device.type == "cuda" and self._graphed is None:
            self._graphed = GraphedTrainStep(self.model, self.optimizer, self.BATCH_SIZE, self.BOARD_SIZE, self.IN_PLANES)
        log = train_batches(self.model, self.optimizer, s, pi, z, self.BATCH_SIZE // self.world,
                            n_epochs or self.N_EPOCHS, self.group, batch_sizes=sizes, graphed=self._graphed)
        self.step += len(log)
        self.total_epoch += n_epochs or self.N_EPOCHS
        if log:
            logging.warning("{:2} Epoch Loss: {:.4f}   Loss_V: {:.4f}   Loss_P: {:.4f}".format(
                self.total_epoch, *np.mean(np.asarray(log), axis=0)))
        return log

    # ---------------------------------------------------------------- checkpoints (main.py:339-365)
    def save(self, n_iter, datetime_now=None):
        if self.rank != 0:
            return None, None
        return (save_model(self.model, n_iter, self.step, datetime_now, self.data_dir),
                save_dataset(self.rep_memory, n_iter, self.step, datetime_now, self.data_dir))

    def load_data(self, model_path=None, dataset_path=None):
        if model_path:
            load_model(self.model, model_path)
            self.step, self.start_iter = parse_model_path(model_path)
        if dataset_path:
            with open(dataset_path, "rb") as f:
                memory = pickle.load(f)
            self.rep_memory = replay.DeviceReplayBuffer(self.BOARD_SIZE, maxlen=self.MEMORY_SIZE,
                                                        tau_thres=self.TAU_THRES, device=self.device)
            self.rep_memory.extend_list(memory)

    def reset_iter(self):
        self.result = {"Black": 0, "White": 0, "Draw": 0}
        self.total_epoch = 0
        self.rep_memory.cur_len = 0

    # ---------------------------------------------------------------- iteration loop (main.py:377-407)
    def run(self, total_iter, save_every=100, n_selfplay_later=1, max_samples=None):
        """main.py:386-407. The reference fills the buffer with N_SELFPLAY episodes in iteration 0 and plays
        N_SELFPLAY = 1 episode per iteration afterwards (main.py:397-400); `n_selfplay_later` is that second number -
        on a B200 thousands of concurrent episodes per iteration cost about the same wall time as one."""
        for n_iter in range(self.start_iter, total_iter):
            if n_iter > 0:
                self.self_play(n_selfplay_later)
                self.train(max_samples=max_samples)
            else:
                self.self_play(self.N_SELFPLAY)
            if n_iter % save_every == 0:
                self.save(n_iter + save_every)
            self.reset_iter()


def main(argv=None):
    """`python -m alpha_omok_b200.trainer` = main.py's `__main__` block (main.py:377-407) with its module constants as flags."""
    import argparse
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--board-size", type=int, default=9, choices=(9, 15))          # env_small / env_regular
    ap.add_argument("--n-mcts", type=int, default=400)
    ap.add_argument("--n-selfplay", type=int, default=100, help="episodes of iteration 0 (N_SELFPLAY)")
    ap.add_argument("--n-selfplay-later", type=int, default=1, help="episodes per later iteration (main.py:398)")
    ap.add_argument("--memory-size", type=int, default=30000)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--total-iter", type=int, default=10000000)
    ap.add_argument("--save-every", type=int, default=100)
    ap.add_argument("--max-slots", type=int, default=4096)
    ap.add_argument("--nn-precision", default="auto", help="auto | 0 (fp16 single pass) | 1 (hi/lo split)")
    ap.add_argument("--max-train-samples", type=int, default=None)
    ap.add_argument("--model-path", default=None)
    ap.add_argument("--dataset-path", default=None)
    ap.add_argument("--data-dir", default="data")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    os.makedirs("logs", exist_ok=True)
    logging.basicConfig(filename="logs/log_{}.txt".format(datetime.now().strftime("%y%m%d")), level=logging.WARNING)
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    tr = Trainer(board_size=a.board_size, n_mcts=a.n_mcts, n_selfplay=a.n_selfplay, memory_size=a.memory_size,
                 batch_size=a.batch_size, seed=a.seed, data_dir=a.data_dir, max_slots=a.max_slots,
                 nn_precision=a.nn_precision if a.nn_precision == "auto" else int(a.nn_precision))
    tr.load_data(a.model_path, a.dataset_path)
    tr.run(a.total_iter, save_every=a.save_every, n_selfplay_later=a.n_selfplay_later, max_samples=a.max_train_samples)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
