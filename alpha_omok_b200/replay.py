"""Replay records of a self-play round and their exchange between ranks.

Games are sharded over ranks with no data-path collective; the ONE exchange step of the path is the all-gather of
the (state, pi, z) records into every rank's replay buffer at the end of a round (main.py:250 `rep_memory.extend`).
Records travel in the compact fixed-size form the device packs (csrc/tree.cu pack_records_kernel):
    int16 n_moves | int8 winner | int8 pad | int16 moves[A] | pad to 4 | uint32 visits[A][A]
(state planes and pi are re-derived from moves / visits by `decode_records`), as one padded slab per rank, so a plain
all_gather_into_tensor over NCCL (NVLink / NVSwitch) is enough - no all-gather-v.  torch.distributed is plumbing here.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi


def record_bytes(A: int) -> int:
    return ((4 + 2 * A + 3) & ~3) + 4 * A * A


class _DevSlab:
    """__cuda_array_interface__ view of the engine-owned record slab (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def device_records(engine, n_games) -> torch.Tensor:
    """Pack the finished games of `engine` and return a uint8 CUDA tensor [n_games, record_bytes] aliasing the slab."""
    engine.records_pack(n_games)
    ptr, bpg = engine.records_dev()
    t = torch.as_tensor(_DevSlab(ptr, n_games * bpg), device=torch.device("cuda", engine.device))
    return t.view(n_games, bpg)


def device_stream_records(engine) -> torch.Tensor:
    """uint8 CUDA tensor [n_episodes, record_bytes] aliasing the record slab of a finished continuous self-play run
    (`Engine.selfplay_stream_begin`), episode i = decision-stream key first_key + i."""
    ptr, bpg, n = engine.stream_records_dev()
    t = torch.as_tensor(_DevSlab(ptr, n * bpg), device=torch.device("cuda", engine.device))
    return t.view(n, bpg)


def allgather_records(local: torch.Tensor, group=None) -> torch.Tensor:
    """[n_local, bytes] uint8 on every rank -> [world * n_local, bytes], rank-major (rank r's games at r*n_local...)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local.clone()
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def decode_records(slab, board_size: int, tau_thres: int = 6, with_states: bool = True):
    """uint8 [n, record_bytes] (torch or numpy) -> the reference's cur_memory list of (state, pi, z) (main.py:201-227)
    plus the {'Black','White','Draw'} result counts."""
    A = board_size * board_size
    buf = slab.cpu().numpy() if isinstance(slab, torch.Tensor) else np.asarray(slab)
    voff = (4 + 2 * A + 3) & ~3
    memory, result = [], {"Black": 0, "White": 0, "Draw": 0}
    for rec in buf:
        k = int(rec[:2].view(np.int16)[0])
        w = int(rec[2])
        moves = rec[4:4 + 2 * A].view(np.int16)
        visits = rec[voff:voff + 4 * A * A].view(np.uint32).reshape(A, A)
        result["Black" if w == 1 else "White" if w == 2 else "Draw"] += 1
        z_black = 1.0 if w == 1 else -1.0 if w == 2 else 0.0
        ids = [(0,) + tuple(int(a) for a in moves[:t]) for t in range(k)]
        states = _cabi.encode_state_batch(ids, board_size).astype(np.float64) if with_states and k else [None] * k
        for t in range(k):
            v = visits[t].astype(np.float64)
            if t < tau_thres:
                pi = v / v.sum()
            else:
                pi = np.zeros(A)
                pi[moves[t]] = 1.0
            memory.append((states[t] if with_states else ids[t], pi, z_black if t % 2 == 0 else -z_black))
    return memory, result


def shard_games(n_total: int, rank: int, world: int):
    """game g -> rank g % world (per-game decision-stream keys are independent of the world size)."""
    return list(range(rank, n_total, world))


def augmented_tensors(slab: torch.Tensor, board_size: int, tau_thres: int = 6):
    """Record slab on the GPU (uint8 [n, record_bytes], e.g. the output of allgather_records) -> the 8-fold augmented
    training set as CUDA float32 tensors (states [N,5,B,B], pi [N,A], z [N]), N = 8 * plies, in the order of
    `utils.augment_dataset(cur_memory)` (main.py:250).  One kernel, nothing leaves the device."""
    import ctypes as C
    assert slab.is_cuda and slab.dtype == torch.uint8 and slab.is_contiguous()
    A = board_size * board_size
    n = slab.shape[0]
    cnt = C.c_longlong(0)
    lib = _cabi.lib()
    stream = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.ao_augment_records_dev(slab.data_ptr(), n, board_size, tau_thres, None, None, None, 0, C.byref(cnt),
                                           stream))
    N = cnt.value
    states = torch.empty((N, 5, board_size, board_size), dtype=torch.float32, device=slab.device)
    pi = torch.empty((N, A), dtype=torch.float32, device=slab.device)
    z = torch.empty((N,), dtype=torch.float32, device=slab.device)
    if N:
        _cabi.check(lib.ao_augment_records_dev(slab.data_ptr(), n, board_size, tau_thres, states.data_ptr(),
                                               pi.data_ptr(), z.data_ptr(), N, C.byref(cnt), stream))
    return states, pi, z


class DeviceReplayBuffer:
    """`rep_memory = deque(maxlen=MEMORY_SIZE)` of main.py:66 kept on the GPU as float32 rings (states [M,5,B,B],
    pi [M,A], z [M]); logical item i (0 = oldest) lives in slot (head + i) % M.

    * `extend_records(slab)`  = `rep_memory.extend(utils.augment_dataset(cur_memory, BOARD_SIZE))` (main.py:250): the
      record slab (this rank's, or the all-gathered one) is decoded, augmented 8-fold and written into the ring by one
      kernel (csrc/augment.cu); `cur_len` is then `len(cur_memory)` of that round (plies before augmentation).
    * `sample(k)`             = `random.sample(rep_memory, k)` (main.py:263-264): the indices come from Python's
      `random` module exactly as the reference draws them (sampling `range(len)` consumes the generator like sampling
      the deque), the rows are gathered on the device.
    * `to_list()` / `extend_list()` speak the reference's pickle format of save_dataset / load_data (main.py:345-365).
    """

    def __init__(self, board_size: int, maxlen: int = 30000, tau_thres: int = 6, device=None):
        self.B, self.A, self.maxlen, self.tau_thres = board_size, board_size * board_size, int(maxlen), tau_thres
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.states = torch.zeros((self.maxlen, 5, board_size, board_size), dtype=torch.float32, device=dev)
        self.pi = torch.zeros((self.maxlen, self.A), dtype=torch.float32, device=dev)
        self.z = torch.zeros((self.maxlen,), dtype=torch.float32, device=dev)
        self.head, self.len, self.cur_len = 0, 0, 0

    def __len__(self):
        return self.len

    def extend_records(self, slab: torch.Tensor) -> int:
        import ctypes as C
        assert slab.is_cuda and slab.dtype == torch.uint8 and slab.is_contiguous()
        head, ln, n = C.c_longlong(self.head), C.c_longlong(self.len), C.c_longlong(0)
        stream = torch.cuda.current_stream().cuda_stream
        _cabi.check(_cabi.lib().ao_replay_extend_dev(slab.data_ptr(), slab.shape[0], self.B, self.tau_thres,
                                                     self.states.data_ptr(), self.pi.data_ptr(), self.z.data_ptr(),
                                                     self.maxlen, C.byref(head), C.byref(ln), C.byref(n), stream))
        self.head, self.len = head.value, ln.value
        self.cur_len = n.value // 8
        return n.value

    def sample_indices(self, k: int, rng=None):
        import random as _random
        return (rng or _random).sample(range(self.len), k)  # ValueError if k > len, like the reference

    def gather(self, indices):
        """logical deque indices -> (states [k,5,B,B], pi [k,A], z [k]) CUDA float32, in that order."""
        k = len(indices)
        dev = self.states.device
        if isinstance(indices, torch.Tensor):
            idx = indices.to(device=dev, dtype=torch.int64).contiguous()
        else:
            idx = torch.as_tensor(np.asarray(indices, np.int64)).to(dev, non_blocking=False)
        if k and (int(idx.min()) < 0 or int(idx.max()) >= self.len):
            raise IndexError("replay index out of range")
        s = torch.empty((k, 5, self.B, self.B), dtype=torch.float32, device=dev)
        p = torch.empty((k, self.A), dtype=torch.float32, device=dev)
        z = torch.empty((k,), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        _cabi.check(_cabi.lib().ao_replay_gather_dev(self.states.data_ptr(), self.pi.data_ptr(), self.z.data_ptr(),
                                                     self.maxlen, self.head, idx.data_ptr(), k, self.B, s.data_ptr(),
                                                     p.data_ptr(), z.data_ptr(), stream))
        return s, p, z

    def sample(self, k: int, rng=None):
        return self.gather(self.sample_indices(k, rng))

    def to_list(self):
        """The deque's content as the reference's list of (state f64 [5,B,B], pi f64 [A], z) (save_dataset)."""
        order = (torch.arange(self.len, device=self.states.device) + self.head) % self.maxlen
        s = self.states[order].cpu().numpy().astype(np.float64)
        p = self.pi[order].cpu().numpy().astype(np.float64)
        z = self.z[order].cpu().numpy().astype(np.float64)
        return [(s[i], p[i], float(z[i])) for i in range(self.len)]

    def extend_list(self, memory):
        """deque.extend of host (state, pi, z) tuples (load_data's pickle, main.py:362-365)."""
        memory = list(memory)[-self.maxlen:]
        if not memory:
            return
        dev = self.states.device
        s = torch.as_tensor(np.stack([np.asarray(m[0], np.float32) for m in memory])).to(dev)
        p = torch.as_tensor(np.stack([np.asarray(m[1], np.float32) for m in memory])).to(dev)
        z = torch.as_tensor(np.asarray([m[2] for m in memory], np.float32)).to(dev)
        n = len(memory)
        slots = (torch.arange(n, device=dev) + self.head + self.len) % self.maxlen
        self.states[slots], self.pi[slots], self.z[slots] = s, p, z
        over = self.len + n - self.maxlen
        if over > 0:
            self.head = (self.head + over) % self.maxlen
        self.len = min(self.len + n, self.maxlen)
