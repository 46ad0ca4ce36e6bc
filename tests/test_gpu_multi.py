"""2-GPU test of the one exchange step of the path: NCCL all-gather of replay records (skipped on 1-GPU boxes)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]

WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from alpha_omok_b200 import _cabi, replay
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G = 6
eng = _cabi.Engine(board_size=9, num_mcts=16, max_games=G, seed=5, device=local, eval_mode=_cabi.AO_EVAL_SYNTH)
eng.selfplay_begin(G, first_key=rank * G)          # per-game keys independent of the world size
st = eng.selfplay_rounds(1)
while st["running"]:
    st = eng.selfplay_rounds(1)
local_slab = replay.device_records(eng, G)
allrec = replay.allgather_records(local_slab)
assert allrec.shape[0] == world * G
assert torch.equal(allrec[rank * G:(rank + 1) * G], local_slab)
# the same games played on ONE engine with keys 0..world*G-1 must give the same records (results do not depend on #GPUs)
if rank == 0:
    ref = _cabi.Engine(board_size=9, num_mcts=16, max_games=world * G, seed=5, device=local, eval_mode=_cabi.AO_EVAL_SYNTH)
    ref.selfplay_begin(world * G, first_key=0)
    st = ref.selfplay_rounds(1)
    while st["running"]:
        st = ref.selfplay_rounds(1)
    assert torch.equal(replay.device_records(ref, world * G), allrec)
    mem, result = replay.decode_records(allrec, 9)
    assert sum(result.values()) == world * G and len(mem) > 0
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank%dok\n" % rank)
sys.stdout.flush()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_replay_allgather_two_gpus(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "rank0ok" in res.stdout and "rank1ok" in res.stdout, res.stdout


TRAINER_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from alpha_omok_b200 import trainer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tr = trainer.Trainer(board_size=9, n_mcts=16, n_blocks=2, n_selfplay=8, memory_size=6000, batch_size=32, seed=0,
                     device="cuda:%d" % local)
n0 = tr.self_play(8)                       # 8 episodes per rank, records of both ranks in every replay ring
assert sum(tr.result.values()) == 8 * world and len(tr.rep_memory) == min(6000, 8 * n0)
# every rank holds the same replay memory ...
flat = torch.cat([tr.rep_memory.states.flatten(), tr.rep_memory.pi.flatten(), tr.rep_memory.z])
lst = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(lst, flat)
assert all(torch.equal(lst[0], t) for t in lst)
tr.reset_iter()
n1 = tr.self_play(2)
log = tr.train()                           # rank r trains on rows r::world of every batch, one gradient all-reduce per step
assert len(log) == n1 and all(np.isfinite(l).all() for l in log)
# ... and the same weights after training (BatchNorm running statistics are per rank by design)
w = torch.cat([p.detach().flatten() for p in tr.model.parameters()])
lst = [torch.empty_like(w) for _ in range(world)]
dist.all_gather(lst, w)
assert all(torch.equal(lst[0], t) for t in lst)
n2 = tr.self_play(2)                       # next round with the trained weights
assert n2 > 0
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank%dok\n" % rank)
sys.stdout.flush()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_trainer_two_gpus_nccl(tmp_path):
    """the whole iteration (sharded device self-play -> NCCL all-gather of records -> device replay ring -> sharded
    batches -> gradient all-reduce) on 2 GPUs: identical replay memories and identical weights on both ranks"""
    script = tmp_path / "trainer_worker.py"
    script.write_text(TRAINER_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "rank0ok" in res.stdout and "rank1ok" in res.stdout, res.stdout
