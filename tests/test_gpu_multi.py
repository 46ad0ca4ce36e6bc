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
