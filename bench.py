#!/usr/bin/env python
"""bench.py - MCTS node-expansions/s of the self-play hot path (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N --steps K --warmup W]           our arm   (one process per GPU under torchrun for N>1)
  python bench.py --impl reference [...]                    the reference arm: the CPU oracle port of the same path

Workload (config[1] of BASELINE.json): 4096 concurrent 9x9 self-play games per GPU, 400 sims/move, seed-0
random-init PVNet(10 blocks, 128 planes), Dirichlet noise on, tau threshold 6, synthetic (empty-board) starts.
A STEP = `sims` lock-step rounds (select -> PVNet tower -> expand/backup, moves played on the device when a search
completes) = one move's worth of search for every game.  Finished episodes are recycled in place so the batch stays
full.  `value` = simulations completed by all ranks / max-over-ranks device time (CUDA events on the launch stream).
`e2e` = the same metric through the reference-facing call (BatchedZeroAgent.get_pi -> ao_search) with root IDs in
pinned host memory copied H2D and visit counts copied D2H inside the timed region, every step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_EXPANSION = {9: 478_800_004, 15: 1_330_129_156}  # one PVNet forward, 2*MAC, padded taps
TOWER_DRAM_BYTES_PER_LAUNCH = {(9, 4096): 6_716_160}  # ncu capture of bench.py itself, round 1 v7 (profiles/r01_tower_stag_kernel_v7_ncu_full_summary.txt)
METRIC = "MCTS node-expansions/sec, 9x9 Omok self-play @400 sims/move"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--board", type=int, default=9)
    ap.add_argument("--games", type=int, default=4096, help="concurrent games per GPU")
    ap.add_argument("--sims", type=int, default=400)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-episodes", action="store_true", help="skip the games/s leg (all games played to the end)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- CPU oracle timing
def _cpu_worker(args):
    """One single-threaded worker: the oracle port (oracle/) of one game's search on torch CPU fp32."""
    board, sims, seconds, widx = args
    import torch
    torch.set_num_threads(1)
    from oracle import omok_oracle as O
    from oracle import pvnet_ref
    A = board * board
    sd = pvnet_ref.make_state_dict(0, 10, 5, 128, board)

    def evaluate(moves):
        x = torch.from_numpy(O.get_state_pt(moves, board, 5).astype(np.float32))[None]
        p, v = pvnet_ref.pvnet_forward(sd, x)
        return p[0].numpy(), v[0].item()

    stream = O.DecisionStream(1234, widx, O.make_gamma_tape(1234, widx, 4, A, 10 / A))
    agent = O.OracleZeroAgent(board, sims, evaluate, stream, noise=True)
    agent._set_root((0,))
    evaluate((0,))  # warm-up (page-in, thread pools)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds and n < sims + 1:
        agent._simulate()
        n += 1
    return n, time.perf_counter() - t0


def cpu_oracle_throughput(board, sims, seconds, workers):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        res = pool.map(_cpu_worker, [(board, sims, seconds, i) for i in range(workers)])
    total = sum(n for n, _ in res)
    elapsed = max(t for _, t in res)
    return total / elapsed, total, elapsed


def workload_config(a, world):
    B, G, S = a.board, a.games, a.sims
    return {"workload": f"{G} parallel {B}x{B} self-play games per GPU, {S} sims/move, PVNet 10x128 random-init (numpy seed 0)",
            "games_per_gpu": G, "sims_per_move": S, "rounds_per_step": S,
            "parallelism": f"dp{world} (games sharded, no data-path collective)",
            "l2": "inputs larger than L2: per-GPU tree storage %.1f GB; 5.9 MB of fp16 weights are L2-resident by design"
                  % (G * 2 * 2048 * B * B * 21 / 1e9)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_step = max(2.0, min(20.0, 150.0 / max(1, a.steps + a.warmup)))
    for _ in range(a.warmup):
        cpu_oracle_throughput(a.board, a.sims, min(per_step, 3.0), workers)
    tot, t = 0, 0.0
    for _ in range(a.steps):
        v, n, el = cpu_oracle_throughput(a.board, a.sims, per_step, workers)
        tot += n
        t += el
    value = tot / t
    sample = (f"{workers} single-thread workers, each running the oracle port of ZeroAgent's simulation loop "
              f"(first move of a {a.board}x{a.board} game, {a.sims} sims budget, torch CPU fp32 PVNet) for {per_step:.0f} s per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "expansions/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t / max(1, a.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, a.gpus),
        "cpu_baseline": {"value": value, "unit": "expansions/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "expansions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from alpha_omok_b200 import _cabi
    from alpha_omok_b200.model import seeded_state_dict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line: everything libraries print to fd 1 meanwhile (e.g. NCCL's version banner,
    # which ignores NCCL_DEBUG_FILE on some builds) is sent to stderr; the JSON goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, G, S = a.board, a.games, a.sims
    A = B * B
    stream = torch.cuda.Stream()
    eng = _cabi.Engine(board_size=B, num_mcts=S, max_games=G, seed=1000 + rank, device=local, stream=stream.cuda_stream)
    eng.load_state_dict(seeded_state_dict(0, 10, 5, 128, B))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident self-play (value)
    eng.selfplay_begin(G, first_key=rank * G, recycle=True)
    st0 = eng.selfplay_rounds(S) if a.warmup > 0 else eng.selfplay_rounds(0)
    for _ in range(max(0, a.warmup - 1)):
        st0 = eng.selfplay_rounds(S)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tree_ms = tower_ms = 0.0
    with torch.cuda.stream(stream):
        ev0.record(stream)
        st = st0
        for _ in range(a.steps):
            st = eng.selfplay_rounds_timed(S)
            tree_ms += st["tree_ms"]
            tower_ms += st["tower_ms"]
        ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    sims = st["sims"] - st0["sims"]
    evals = st["nn_evals"] - st0["nn_evals"]
    moves = st["moves"] - st0["moves"]
    games_done = st["games_finished"] - st0["games_finished"]
    if st["errors"]:
        raise SystemExit(f"bench.py: {st['errors']} game tree(s) overflowed their arena")

    # ---------------- end to end through the reference-facing call (e2e)
    e2e = None
    if not a.no_e2e:
        ids = np.arange(G, dtype=np.int32)
        roots_t = torch.full((G, A + 1), -1, dtype=torch.int16).pin_memory()
        lens_t = torch.ones(G, dtype=torch.int32).pin_memory()
        vis_t = torch.zeros((G, A), dtype=torch.int32).pin_memory()
        real_t = torch.zeros(G, dtype=torch.int32).pin_memory()
        roots, lens, vis, real = roots_t.numpy(), lens_t.numpy(), vis_t.numpy().view(np.uint32), real_t.numpy()
        roots[:, 0] = 0
        eng.games_reset(ids, keys=(ids + rank * G).astype(np.uint32))
        e_sims = 0
        e_ms = 0.0
        n_e2e_steps = max(1, min(a.steps, 3))
        n_e2e_warm = 1 if a.warmup > 0 else 0
        for k in range(n_e2e_warm + n_e2e_steps):
            timed = k >= n_e2e_warm
            if timed:
                barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                eng.search_raw(ids, roots, lens, vis, None, real)  # H2D roots -> 400(+1) sims/game -> D2H visits
                e1.record(stream)
            torch.cuda.synchronize()
            if timed:
                e_ms += e0.elapsed_time(e1)
                e_sims += G * S + int(real.sum())  # a real root runs num_mcts + 1 simulations (agents.py:107-111)
            # host side of the loop (main.py:170-171): play the most visited move, extend the root IDs
            best = vis.argmax(axis=1)
            roots[ids, lens] = best.astype(np.int16)
            lens += 1
        e2e = {"sims": e_sims, "ms": e_ms, "h2d": int(roots.nbytes + lens.nbytes + ids.nbytes), "d2h": int(vis.nbytes + real.nbytes),
               "steps": n_e2e_steps}

    # ---------------- games/s: every game of the batch played to its end (ragged tail included)
    full = None
    if not a.no_full_episodes:
        eng.selfplay_begin(G, first_key=rank * G + 7 * G, recycle=False)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            f0.record(stream)
            fs = eng.selfplay_rounds(S)
            while fs["running"]:
                fs = eng.selfplay_rounds(S)
            f1.record(stream)
        torch.cuda.synchronize()
        full = {"ms": f0.elapsed_time(f1), "games": G, "moves": fs["moves"], "sims": fs["sims"], "errors": fs["errors"]}
        # the one exchange step of the path: all-gather of this round's replay records into every rank
        from alpha_omok_b200 import replay
        for timed in (False, True):                        # one untimed pass: allocator / communicator warm-up
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            local = replay.device_records(eng, G)          # pack kernel on the engine stream (synchronised inside)
            gathered = replay.allgather_records(local)     # NCCL all_gather_into_tensor of the fixed-size slabs
            g1.record()
            torch.cuda.synchronize()
            if not timed:
                del gathered
        full.update(allgather_ms=g0.elapsed_time(g1), allgather_bytes=int(gathered.numel()),
                    allgather_ok=bool(torch.equal(gathered[rank * G:(rank + 1) * G], local)))
        del gathered

    # ---------------- reduce over ranks
    t = torch.tensor([ms, float(sims), float(evals), float(moves), float(games_done), tower_ms, tree_ms,
                      float(launches), e2e["ms"] if e2e else 0.0, float(e2e["sims"]) if e2e else 0.0,
                      full["ms"] if full else 0.0, float(full["games"]) if full else 0.0,
                      float(full["moves"]) if full else 0.0, float(full["errors"]) if full else 0.0,
                      full["allgather_ms"] if full else 0.0, float(full["allgather_ok"]) if full else 1.0],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_max, e_ms_max, f_ms_max, ag_ms_max = tmax[0].item(), tmax[8].item(), tmax[10].item(), tmax[14].item()
    else:
        ms_max, e_ms_max, f_ms_max, ag_ms_max = t[0].item(), t[8].item(), t[10].item(), t[14].item()
    tot = t.cpu().numpy()

    if rank == 0:
        sustained, burst, peak_src = peaks()
        value = tot[1] / (ms_max * 1e-3)
        tower_ms_rank = tot[5] / world
        achieved = (tot[2] / world) * FLOP_PER_EXPANSION[B] / (tower_ms_rank * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "expansions/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_max / max(1, a.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 (tensor-core inputs, f32 accumulate); tree: f32 w/q, f64 p/u",
            "data": "synthetic",
            "config": workload_config(a, world),
            "moves_per_s": tot[3] / (ms_max * 1e-3), "games_finished_in_window": int(tot[4]),
            "gpu_launches": int(tot[7]),
            "roofline": {"bound": "tensor", "kernel": "tower_stag_kernel", "achieved": achieved, "peak": sustained,
                         "unit": "TFLOP/s", "frac": achieved / sustained, "peak_source": peak_src + ", bf16 sustained",
                         "frac_of_burst": achieved / burst,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch (4096 leaves), from the committed
                         # ncu --set full capture profiles/r01_tower_stag_kernel_v7_ncu_full_summary.txt
                         "traffic": TOWER_DRAM_BYTES_PER_LAUNCH.get((B, G)),
                         "kernel_share_of_step": tot[5] / (tot[5] + tot[6]) if tot[5] + tot[6] > 0 else None,
                         "tower_ms_per_launch": tower_ms_rank / (a.steps * S), "tree_ms_per_launch": tot[6] / world / (a.steps * S),
                         "flop_per_expansion": FLOP_PER_EXPANSION[B]},
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = {"value": tot[9] / (e_ms_max * 1e-3), "unit": "expansions/s", "h2d_bytes_per_step": e2e["h2d"],
                           "d2h_bytes_per_step": e2e["d2h"], "steps": e2e["steps"],
                           "api": "Engine.search_raw == BatchedZeroAgent.get_pi (ao_search), pinned host buffers"}
        if full:
            line["selfplay_games_per_s"] = tot[11] / (f_ms_max * 1e-3)
            line["full_episodes"] = {"games": int(tot[11]), "seconds": f_ms_max * 1e-3, "moves_per_game": tot[12] / tot[11],
                                     "tree_overflows": int(tot[13])}
            line["replay_allgather"] = {"ms": ag_ms_max, "bytes_gathered_per_rank": full["allgather_bytes"],
                                        "records": "pack kernel + NCCL all_gather_into_tensor of fixed-size record slabs",
                                        "verified_own_shard": bool(tot[15] == world)}
        if not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            workers = max(1, min(cores, 64))
            v, n, el = cpu_oracle_throughput(B, S, a.cpu_seconds, workers)
            line["cpu_baseline"] = {"value": v, "unit": "expansions/s", "cores": workers, "kind": "port",
                                    "sample": f"{workers} single-thread workers x {a.cpu_seconds:.0f} s of the oracle port's simulation loop "
                                              f"(first move, {B}x{B}, {S} sims budget, torch CPU fp32 PVNet); host has {cores} cpus"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
