"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares; host-side logic; loud failure
without a GPU; replay record exchange over gloo (world_size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import omok_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_library_exports_every_declared_symbol():
    from alpha_omok_b200 import _cabi
    lib = _cabi.lib()
    hdr = open(os.path.join(ROOT, "include", "alpha_omok_b200.h")).read()
    declared = set(re.findall(r"\b(ao_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert getattr(lib, name) is not None, name
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)


def test_product_library_carries_no_probe_code():
    """experiment switches and probes live in libalpha_omok_b200_probe.so only (include/alpha_omok_b200_probe.h): the
    product library exports none of the probe entry points and does not even contain the AO_TOWER_XFLAGS string"""
    from alpha_omok_b200 import _build, _cabi
    lib = _cabi.lib()
    for name in _cabi.PROBE_EXPORTS:
        assert not hasattr(lib, name), name
    blob = open(_build.LIB_PATH, "rb").read()
    assert b"AO_TOWER_XFLAGS" not in blob and b"ao_umma" not in blob
    hdr = open(os.path.join(ROOT, "include", "alpha_omok_b200_probe.h")).read()
    assert set(re.findall(r"\b(ao_[a-z0-9_]+)\s*\(", hdr)) == set(_cabi.PROBE_EXPORTS)
    probe = _cabi.probe_lib()      # builds it on demand; exports product + probe symbols
    for name in _cabi.EXPORTS + _cabi.PROBE_EXPORTS:
        assert getattr(probe, name) is not None, name


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    from alpha_omok_b200 import _cabi, utils
    with pytest.raises(_cabi.AoError, match="no CUDA device"):
        _cabi.Engine(board_size=9, num_mcts=4, max_games=1)
    with pytest.raises(_cabi.AoError, match="no CPU fallback"):
        utils.check_win(np.zeros((9, 9)), 5)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "alpha_omok_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_host_helpers():
    from alpha_omok_b200 import _cabi, utils
    ids, lens = _cabi.pad_ids([(0,), (0, 5, 7)], 81)
    assert ids.shape == (2, 82) and list(lens) == [1, 3] and ids[1, 2] == 7 and ids[0, 1] == -1
    assert utils.get_turn((0,)) == O.get_turn((0,)) == 0 and utils.get_turn((0, 3)) == O.get_turn((0, 3)) == 1
    rs = np.random.RandomState(0)
    mem = [(rs.rand(5, 9, 9), rs.rand(81), 1.0)]
    a, b = utils.augment_dataset(mem, 9), O.augment_dataset(mem, 9)
    assert len(a) == len(b) == 8
    for (s1, p1, z1), (s2, p2, z2) in zip(a, b):
        assert np.array_equal(s1, s2) and np.array_equal(p1, p2) and z1 == z2
    np.random.seed(3)
    x = utils.get_action(np.eye(81)[12])
    assert x[1] == 12 and x[0][12] == 1
    assert utils.argmax_onehot(np.eye(81)[30])[1] == 30


def test_model_container_is_state_dict_compatible():
    from alpha_omok_b200 import model
    from oracle import pvnet_ref
    net = model.PVNet(2, 5, 128, 9)
    sd = pvnet_ref.make_state_dict(3, 2, 5, 128, 9, bn_jitter=True)
    missing = net.load_state_dict(sd, strict=False)
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys) and not missing.unexpected_keys
    net.train()  # training path = plain torch (out of hot-path scope); eval / no_grad goes to the CUDA tower
    x = torch.zeros(2, 5, 9, 9)
    x[:, 4] = 1
    with torch.enable_grad():
        net.eval()
        p, v = net(x)
    pr, vr = pvnet_ref.pvnet_forward(sd, x)
    assert torch.allclose(p, pr, atol=1e-6) and torch.allclose(v, vr, atol=1e-6)


def test_record_codec_roundtrip():
    from alpha_omok_b200 import replay
    B, A = 9, 81
    rb = replay.record_bytes(A)
    rs = np.random.RandomState(0)
    slab = np.zeros((3, rb), np.uint8)
    voff = (4 + 2 * A + 3) & ~3
    truth = []
    for g in range(3):
        k = 7 + g
        mv = rs.permutation(A)[:k].astype(np.int16)
        vis = np.zeros((A, A), np.uint32)
        for t in range(k):
            vis[t, rs.permutation(A)[:10]] = rs.randint(1, 50, 10)
            vis[t, mv[t]] += 1
        slab[g, :2] = np.asarray([k], np.int16).view(np.uint8)
        slab[g, 2] = g + 1
        slab[g, 4:4 + 2 * A] = np.concatenate([mv, -np.ones(A - k, np.int16)]).view(np.uint8)
        slab[g, voff:] = vis.view(np.uint8).reshape(-1)
        truth.append((k, mv, vis))
    mem, result = replay.decode_records(slab, B, tau_thres=6, with_states=False)
    assert result == {"Black": 1, "White": 1, "Draw": 1}
    assert len(mem) == sum(k for k, _, _ in truth)
    sid, pi, z = mem[0]
    assert sid == (0,) and abs(pi.sum() - 1) < 1e-12 and z == 1.0
    sid, pi, z = mem[6]  # ply 6 of game 0: one-hot target, white... ply index 6 is black's 4th move
    assert pi[truth[0][1][6]] == 1.0 and pi.sum() == 1.0 and z == 1.0
    assert mem[1][2] == -1.0


GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from alpha_omok_b200 import replay
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
rb = replay.record_bytes(81)
n_local = 3
local = torch.full((n_local, rb), rank + 1, dtype=torch.uint8)
local[:, 0] = torch.arange(n_local, dtype=torch.uint8)
out = replay.allgather_records(local)
assert out.shape == (world * n_local, rb)
for r in range(world):
    blk = out[r * n_local:(r + 1) * n_local]
    assert (blk[:, 1:] == r + 1).all() and (blk[:, 0] == torch.arange(n_local, dtype=torch.uint8)).all()
games = [replay.shard_games(10, r, world) for r in range(world)]
assert sorted(sum(games, [])) == list(range(10))
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank%dok\n" % rank)
sys.stdout.flush()
"""


def test_allgather_records_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "rank0ok" in res.stdout and "rank1ok" in res.stdout, res.stdout


def test_replay_entry_points_reject_bad_arguments_without_touching_the_gpu():
    """argument validation of the replay-ring C ABI happens before any CUDA call (null / out-of-range -> -1)"""
    import ctypes as C
    from alpha_omok_b200 import _cabi
    L = _cabi.lib()
    n = C.c_longlong(0)
    assert L.ao_augment_records_dev(None, 1, 9, 6, None, None, None, 0, C.byref(n), None) == -1
    assert L.ao_augment_records_dev(1, 0, 9, 6, None, None, None, 0, C.byref(n), None) == -1
    assert L.ao_augment_records_dev(1, 1, 16, 6, None, None, None, 0, C.byref(n), None) == -1
    head, ln = C.c_longlong(0), C.c_longlong(0)
    assert L.ao_replay_extend_dev(None, 1, 9, 6, 1, 1, 1, 10, C.byref(head), C.byref(ln), C.byref(n), None) == -1
    head.value = 10                                   # head must be < cap
    assert L.ao_replay_extend_dev(1, 1, 9, 6, 1, 1, 1, 10, C.byref(head), C.byref(ln), C.byref(n), None) == -1
    head.value, ln.value = 0, 11                      # len must be <= cap
    assert L.ao_replay_extend_dev(1, 1, 9, 6, 1, 1, 1, 10, C.byref(head), C.byref(ln), C.byref(n), None) == -1
    assert L.ao_replay_gather_dev(None, 1, 1, 10, 0, 1, 1, 9, 1, 1, 1, None) == -1
    assert L.ao_replay_gather_dev(1, 1, 1, 10, 10, 1, 1, 9, 1, 1, 1, None) == -1
    assert L.ao_replay_gather_dev(1, 1, 1, 10, 0, 1, 0, 9, 1, 1, 1, None) == 0    # k = 0: nothing to do


def test_aux_utils_match_reference_golden(capsys):
    """render_str / valid_actions / get_reward / get_state_tf (host helpers off the hot path) vs outputs of the
    unmodified reference (tests/golden/utils_aux.json)"""
    import json
    from alpha_omok_b200 import utils
    with open(os.path.join(ROOT, "tests", "golden", "utils_aux.json")) as f:
        cases = json.load(f)
    assert len(cases) == 24
    for c in cases:
        B = c["B"]
        board = np.zeros((B, B))
        for t, cell in enumerate(c["cells"]):
            board[cell // B, cell % B] = 1 if t % 2 == 0 else -1
        capsys.readouterr()
        utils.render_str(board, B, c["last"])
        assert capsys.readouterr().out == c["render"]
        assert [[list(a[0]), a[1]] for a in utils.valid_actions(board)] == c["valid"]
        rid = (0,) + tuple(c["cells"])
        assert [utils.get_reward(w, rid) for w in (0, 1, 2, 3)] == c["reward"]
        for t in (0, 1):
            st = utils.get_state_tf(rid, t, B, 5)
            assert float((st * np.arange(1, 6)).sum()) == c["state_tf_sum"][t]
        if c["state_tf"] is not None:
            assert np.array_equal(utils.get_state_tf(rid, 0, B, 5), np.asarray(c["state_tf"]))


def test_weights_are_uploaded_once_per_change_not_once_per_search():
    """the facades re-fold / re-upload the network only when its parameters changed: state_dict() returns fresh alias
    tensors on every call, so the fingerprint must not depend on their id()"""
    from alpha_omok_b200 import agents, model

    class FakeEngine:
        uploads = probes = 0

        def load_state_dict(self, sd):
            self.uploads += 1

        def choose_nn_precision(self):
            self.probes += 1

    agent = agents.ZeroAgent(9, 8, 5)
    agent.model = model.PVNet(2, 5, 128, 9)
    agent._engine = FakeEngine()
    for _ in range(3):
        agent._sync_weights()
    assert agent._engine.uploads == 1 and agent._engine.probes == 1
    opt = torch.optim.SGD(agent.model.parameters(), lr=0.1)
    agent.model.train()
    p, v = agent.model(torch.zeros(2, 5, 9, 9))
    (p.sum() + v.sum()).backward()
    opt.step()
    agent._sync_weights()
    agent._sync_weights()
    assert agent._engine.uploads == 2
    agent.model.load_state_dict(model.seeded_state_dict(1, 2, 5, 128, 9), strict=False)
    agent._sync_weights()
    assert agent._engine.uploads == 3
    other = model.PVNet(2, 5, 128, 9)       # main.py:81 style re-assignment of Agent.model
    agent.model = other
    agent._sync_weights()
    assert agent._engine.uploads == 4
    with torch.no_grad():                    # `.data` write (own version counter) on CPU weights: seen through the sampled values
        other.conv1.weight.data.mul_(1.5)
    agent._sync_weights()
    assert agent._engine.uploads == 5
    other.conv1.weight = torch.nn.Parameter(other.conv1.weight.detach().clone())  # replaced Parameter object
    agent._sync_weights()
    agent._sync_weights()
    assert agent._engine.uploads == 6
    agent.invalidate_weights()
    agent._sync_weights()
    assert agent._engine.uploads == 7
    agent._engine = None                     # do not let __del__ paths touch the fake


def test_dashboard_feed_payloads_match_reference():
    """info.AgentInfo / GameInfo / periodic_status / prompt_status against the payloads the reference's own objects and
    Flask routes produce for the same state (tests/golden/dashboard_feed.json, webapi.py:28-76, info/*.py)"""
    import json
    from alpha_omok_b200 import agents, info
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mg", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    src = open(os.path.join(ROOT, "tests", "golden", "make_golden.py")).read()
    ns = {"np": np}
    start = src.index("def fill_dashboard(")
    exec(src[start:src.index("\n\n\n", start)], ns)       # the very function that filled the reference's objects
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "dashboard_feed.json")))
    d = info.Dashboard(9)

    class ZeroAgent(agents.Agent):      # get_name() is the class name shown in the dashboard (webapi.py:41-42)
        pass

    ns["fill_dashboard"](d.game_info, d.player_agent_info, d.enemy_agent_info, 9, ZeroAgent(9), agents.RandomAgent(9))
    assert d.periodic_status() == want["periodic_status"]
    assert d.prompt_status() == want["prompt_status"]
    d.player_agent_info.clear_values()
    assert d.periodic_status() == want["after_clear"]
