"""Runs the UNMODIFIED reference (oracle/_ref, see oracle/make_ref.py) on the host CPU: baseline timing for bench.py.
Test / baseline infrastructure, NOT product code; nothing under alpha_omok_b200/ imports this.

Two protocols:
  config1(...)      BASELINE config 1 verbatim: np.random.seed(0); torch.manual_seed(0); model.PVNet(10,5,128,9) with
                    its default init; agents.ZeroAgent(9, 40, 5, noise=True); env_small.GameState('text'); the loop of
                    main.py:144-248 with TAU_THRES = 6 and printing off, torch threads = all cores.  main.py itself
                    cannot be imported (it opens log files, builds the model and seeds at import time, main.py:21-85),
                    so its 50-line loop is restated here around the reference's own objects.
  SelfPlayWorker    one single-threaded self-play game at the bench's own setting (400 sims/move), advanced one
                    `get_pi` at a time: N of these in N processes = the "whole host" aggregate.
When oracle/_ref is absent the same two protocols run on the oracle port (kind "port").
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "MANIFEST.json"))


def import_reference():
    """agents / model / utils / env of the unmodified reference (CPU: call with CUDA hidden from the process)."""
    import warnings
    warnings.filterwarnings("ignore", message="Creating a tensor from a list of numpy.ndarrays")  # agents.py:175
    for p in (os.path.join(REF, "2_AlphaOmok"), os.path.join(REF, "stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import agents
    import model
    import utils
    from env import env_regular, env_small
    agents.PRINT_MCTS = False
    return agents, model, utils, env_small, env_regular


class _PortAgent:
    """the oracle port behind the reference's ZeroAgent surface (fallback when oracle/_ref is missing)"""

    def __init__(self, board, sims, seed):
        import torch
        from oracle import omok_oracle as O
        from oracle import pvnet_ref
        self.O, self.board, self.sims = O, board, sims
        sd = pvnet_ref.make_state_dict(0, 10, 5, 128, board)

        def evaluate(moves):
            x = torch.from_numpy(O.get_state_pt(moves, board, 5).astype(np.float32))[None]
            p, v = pvnet_ref.pvnet_forward(sd, x)
            return p[0].numpy(), v[0].item()

        A = board * board
        self.stream = O.DecisionStream(1234, seed, O.make_gamma_tape(1234, seed, A + 2, A, 10 / A))
        self.agent = O.OracleZeroAgent(board, sims, evaluate, self.stream, noise=True)

    def get_pi(self, root_id, tau):
        pi = self.agent.get_pi(root_id, tau)
        self.is_real_root = self.agent.is_real_root
        return pi

    def reset(self):
        self.agent.reset()


class SelfPlayWorker:
    """one self-play game of main.py:144-248, one move (= one get_pi of num_mcts (+1) simulations) per step()"""

    def __init__(self, board, sims, seed, threads=1):
        import torch
        torch.set_num_threads(threads)
        self.board, self.sims = board, sims
        self.kind = "reference" if available() else "port"
        np.random.seed(seed)
        torch.manual_seed(seed)
        if self.kind == "reference":
            agents, model, utils, env_small, env_regular = import_reference()
            self.utils = utils
            self.game = env_small if board == 9 else env_regular
            self.agent = agents.ZeroAgent(board, sims, 5, noise=True)
            self.agent.model = model.PVNet(10, 5, 128, board)
        else:
            from oracle import omok_oracle as O
            self.O = O
            self.agent = _PortAgent(board, sims, seed)
        self._new_game()

    def _new_game(self):
        self.root_id, self.t = (0,), 0
        if self.kind == "reference":
            self.env = self.game.GameState("text")
        else:
            self.env = self.O.OracleGameState(self.board)

    def step(self):
        """-> (simulations run, seconds, game finished)"""
        t0 = time.perf_counter()
        tau = 1 if self.t < 6 else 0
        pi = self.agent.get_pi(self.root_id, tau)
        n = self.sims + (1 if self.agent.is_real_root else 0)
        if self.kind == "reference":
            action, idx = self.utils.get_action(pi)
        else:
            action, idx = self.O.get_action(pi, self.agent.stream)
        self.root_id += (int(idx),)
        _, _, win, _, _ = self.env.step(action)
        self.t += 1
        done = win != 0
        if done:
            self.agent.reset()
            self._new_game()
        return n, time.perf_counter() - t0, done


def config1(threads, sims=40, seed=0, board=9):
    """BASELINE config 1 verbatim. Returns dict(sims, seconds, moves, winner, visit_sha256_16, kind, threads)."""
    import torch
    torch.set_num_threads(threads)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if not available():
        w = SelfPlayWorker(board, sims, seed, threads)
        n_tot, t_tot, moves, done = 0, 0.0, 0, False
        while not done:
            n, t, done = w.step()
            n_tot, t_tot, moves = n_tot + n, t_tot + t, moves + 1
        return dict(sims=n_tot, seconds=t_tot, moves=moves, winner=None, visit_sha256_16=None, kind="port",
                    threads=threads)
    agents, model, utils, env_small, env_regular = import_reference()
    game = env_small if board == 9 else env_regular
    Agent = agents.ZeroAgent(board, sims, 5, noise=True)
    Agent.model = model.PVNet(10, 5, 128, board)
    Agent.model.eval()
    env = game.GameState("text")
    root_id, win_index, time_steps, n_sims = (0,), 0, 0, 0
    h = hashlib.sha256()
    t0 = time.perf_counter()
    while win_index == 0:                                   # main.py:144
        tau = 1 if time_steps < 6 else 0                    # main.py:150-153 (TAU_THRES = 6)
        pi = Agent.get_pi(root_id, tau)                     # main.py:155
        n_sims += sims + (1 if Agent.is_real_root else 0)
        h.update(Agent.visit.astype(np.int64).tobytes())
        utils.get_state_pt(root_id, board, 5)               # main.py:159 (sample collection)
        action, action_index = utils.get_action(pi)         # main.py:170
        root_id += (action_index,)                          # main.py:171
        _, _, win_index, _, _ = env.step(action)            # main.py:196
        time_steps += 1
    seconds = time.perf_counter() - t0
    Agent.reset()                                           # main.py:248
    return dict(sims=n_sims, seconds=seconds, moves=time_steps, winner=int(win_index),
                visit_sha256_16=h.hexdigest()[:16], kind="reference", threads=threads)


def _config1_entry(conn, threads, sims, seed, board):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    os.environ["OMP_NUM_THREADS"] = str(threads)   # torchrun exports OMP_NUM_THREADS=1 to its ranks and their children
    conn.send(config1(threads, sims, seed, board))


def _worker_entry(conn, board, sims, seed):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""   # the reference picks 'cuda' when it sees one (agents.py:12-13): CPU path only
    os.environ["OMP_NUM_THREADS"] = "1"
    w = SelfPlayWorker(board, sims, seed, threads=1)
    conn.send(("ready", w.kind))
    while True:
        msg = conn.recv()
        if msg == "stop":
            return
        conn.send(w.step())


class HostPool:
    """`workers` single-thread processes, each playing its own self-play game with the reference; step() advances every
    game by one move and returns (total simulations, max seconds over workers, games finished)."""

    def __init__(self, board, sims, workers):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        self.procs, self.conns = [], []
        for i in range(workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker_entry, args=(b, board, sims, i), daemon=True)
            p.start()
            self.procs.append(p)
            self.conns.append(a)
        self.kind = [c.recv() for c in self.conns][0][1]

    def step(self):
        for c in self.conns:
            c.send("go")
        res = [c.recv() for c in self.conns]
        return sum(r[0] for r in res), max(r[1] for r in res), sum(bool(r[2]) for r in res)

    def close(self):
        for c in self.conns:
            try:
                c.send("stop")
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)


def run_config1_subprocess(threads, sims=40, seed=0, board=9, timeout=150.0):
    """config1() in a fresh process.  Bounded: a host whose cores are busy elsewhere (other ranks of a multi-GPU run
    spinning in their CUDA / NCCL waits) can slow an all-core OpenMP run down by two orders of magnitude - after
    `timeout` seconds the process is killed and TimeoutError raised."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    a, b = ctx.Pipe()
    p = ctx.Process(target=_config1_entry, args=(b, threads, sims, seed, board), daemon=True)
    p.start()
    if not a.poll(timeout):
        p.kill()
        p.join(timeout=10)
        raise TimeoutError("config 1 on the host did not finish within %.0f s" % timeout)
    out = a.recv()
    p.join(timeout=10)
    return out
