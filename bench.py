#!/usr/bin/env python
"""bench.py - MCTS node-expansions/s of the self-play hot path (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N --steps K --warmup W]           our arm   (one process per GPU under torchrun for N>1)
  python bench.py --impl reference [...]                    the reference arm: the UNMODIFIED reference (oracle/_ref,
                                                            assembled by oracle/make_ref.py) on all host cores

Workload (config[1] of BASELINE.json): 4096 concurrent 9x9 self-play games per GPU, 400 sims/move, seed-0
random-init PVNet(10 blocks, 128 planes), Dirichlet noise on, tau threshold 6, synthetic (empty-board) starts.
A STEP = `sims` lock-step rounds (select -> PVNet tower -> expand/backup, moves played on the device when a search
completes) = one move's worth of search for every game.  Finished episodes are recycled in place so the batch stays
full.  `value` = simulations completed by all ranks / max-over-ranks device time (CUDA events on the launch stream).
`e2e` = the same metric through the reference-facing call (BatchedZeroAgent.get_pi -> ao_search) with root IDs in
pinned host memory copied H2D and visit counts copied D2H inside the timed region, every step.
At N = 1 the line also carries `legs`: the other BASELINE configs measured the same way (config 3: 4096 games of 15x15;
config 5: the arena, 1024 concurrent matches at 800 sims/move, shipped trained checkpoint vs a random-init net, both
networks in one engine; 4096-game self-play with the trained checkpoint = the hi/lo split tower), each with its own
roofline numbers, and `single_game` = BASELINE config 1 itself on the device: ONE game driven move by move through
agents.ZeroAgent.get_pi (40 and 400 sims/move; the cluster-of-four kernel of tower_solo.cu), wall clock.  `cpu_baseline` = the unmodified reference on the host: BASELINE config 1 verbatim (one 40-sims/move
game, all torch threads) and the all-core aggregate at the bench's own 400 sims/move.
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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE tower launch, `ncu --set full` captures of round 2
# (profiles/r02_kernels_ncu_full_summary.txt); key = (board, leaves per launch, tower mode)
TOWER_DRAM_BYTES_PER_LAUNCH = {(9, 4096, "fp16"): 6_716_416, (15, 4096, "fp16"): 7_138_560, (9, 4096, "split"): 12_632_064}
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
    ap.add_argument("--no-legs", action="store_true", help="skip the config 3 / config 5 / trained-net legs (N = 1 only)")
    ap.add_argument("--leg-steps", type=int, default=2, help="timed steps per leg")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- CPU reference timing
def host_workers():
    cores = os.cpu_count() or 1
    return cores, max(1, min(cores, 64))


def cpu_aggregate(board, sims, steps, warmup, workers):
    """`workers` single-thread processes, each playing its own self-play game with the reference's own agents /
    model / env (oracle/ref_runner.py); one step = one move (= one get_pi of `sims` (+1) simulations) per worker.
    Returns (expansions/s over the timed steps, total sims, seconds, kind, per-step seconds)."""
    from oracle import ref_runner
    pool = ref_runner.HostPool(board, sims, workers)
    try:
        for _ in range(warmup):
            pool.step()
        tot, secs, per = 0, 0.0, []
        for _ in range(steps):
            n, t, _ = pool.step()
            tot += n
            secs += t
            per.append(t)
    finally:
        pool.close()
    return tot / secs, tot, secs, pool.kind, per


def cpu_baseline_block(board, sims, cores, workers, agg_steps=2, with_config1=True):
    from oracle import ref_runner
    v, n, secs, kind, _ = cpu_aggregate(board, sims, agg_steps, 0, workers)
    out = {"value": v, "unit": "expansions/s", "cores": workers, "kind": kind,
           "sample": f"{workers} single-thread processes, each playing its own {board}x{board} self-play game at {sims} sims/move "
                     f"with the {'unmodified reference (agents.ZeroAgent + model.PVNet on torch CPU fp32 + env)' if kind == 'reference' else 'oracle port'}: "
                     f"{agg_steps} moves per process = {n} simulations in {secs:.1f} s; host has {cores} cpus"}
    c1 = None
    if with_config1:
        try:
            c1 = ref_runner.run_config1_subprocess(workers)
        except Exception as e:  # bounded (ref_runner): the aggregate above stands on its own
            out["config1"] = {"error": repr(e)}
    if c1 is not None:
        out["config1"] = {"what": "BASELINE config 1 verbatim: one 9x9 self-play game, 40 sims/move, seeds 0, random-init "
                                  "PVNet(10,5,128,9), main.py:144-248 loop, torch threads = cores",
                          "sims_per_s": c1["sims"] / c1["seconds"], "games_per_s": 1.0 / c1["seconds"], "sims": c1["sims"],
                          "moves": c1["moves"], "winner": c1["winner"], "seconds": c1["seconds"], "threads": c1["threads"],
                          "visit_sha256_16": c1["visit_sha256_16"], "kind": c1["kind"]}
    return out


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
    cores, workers = host_workers()
    value, tot, secs, kind, per = cpu_aggregate(a.board, a.sims, a.steps, a.warmup, workers)
    sample = (f"{workers} single-thread processes, each playing its own {a.board}x{a.board} self-play game at {a.sims} sims/move with the "
              f"{'unmodified reference (oracle/_ref: agents.ZeroAgent, model.PVNet on torch CPU fp32, env)' if kind == 'reference' else 'oracle port'}; "
              f"one step = one move (one get_pi) per process; host has {cores} cpus")
    cb = {"value": value, "unit": "expansions/s", "cores": workers, "kind": kind, "sample": sample}
    try:
        from oracle import ref_runner
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            raise RuntimeError("config 1 verbatim is timed at N = 1 only")
        c1 = ref_runner.run_config1_subprocess(workers)
        cb["config1"] = {"sims_per_s": c1["sims"] / c1["seconds"], "games_per_s": 1.0 / c1["seconds"], "sims": c1["sims"],
                         "moves": c1["moves"], "winner": c1["winner"], "threads": c1["threads"],
                         "visit_sha256_16": c1["visit_sha256_16"], "kind": c1["kind"]}
    except Exception as e:  # the aggregate above is the line's value; config 1 is extra evidence
        cb["config1"] = {"error": repr(e)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "expansions/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * secs / max(1, a.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, a.gpus),
        "cpu_baseline": cb,
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


# ---------------------------------------------------------------------------------------------- legs
def _trained_state_dict():
    """the reference's shipped 9x9 checkpoint (tests/golden/trained_9x9_180927.npz = data/180927_9400_297233_step_model.pickle)"""
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", "trained_9x9_180927.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _timed_steps(eng, rounds, steps, warmup, stream, torch):
    """`warmup` untimed + `steps` timed steps of `rounds` lock-step rounds; device time from CUDA events on the launch
    stream, tower / tree time from the per-launch events of ao_selfplay_rounds_timed."""
    st0 = None
    for _ in range(max(1, warmup)):
        st0 = eng.selfplay_rounds(rounds)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tower_ms = tree_ms = 0.0
    with torch.cuda.stream(stream):
        e0.record(stream)
        st = st0
        for _ in range(steps):
            st = eng.selfplay_rounds_timed(rounds)
            tower_ms += st["tower_ms"]
            tree_ms += st["tree_ms"]
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if st["errors"]:
        raise SystemExit(f"bench.py: {st['errors']} game tree(s) overflowed their arena")
    return {"ms": ms, "sims": st["sims"] - st0["sims"], "evals": st["nn_evals"] - st0["nn_evals"],
            "moves": st["moves"] - st0["moves"], "finished": st["games_finished"] - st0["games_finished"],
            "tower_ms": tower_ms, "tree_ms": tree_ms, "launches": steps * rounds}


def _leg_line(name, workload, board, m, steps, rounds, passes_per_eval=1, traffic=None):
    sustained, burst, peak_src = peaks()
    achieved = m["evals"] * FLOP_PER_EXPANSION[board] / (m["tower_ms"] * 1e-3) / 1e12 if m["tower_ms"] > 0 else None
    return {"workload": workload, "value": m["sims"] / (m["ms"] * 1e-3), "unit": "expansions/s", "steps": steps,
            "rounds_per_step": rounds, "ms_per_step": m["ms"] / steps, "moves_per_s": m["moves"] / (m["ms"] * 1e-3),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                         "frac": achieved / sustained if achieved else None, "frac_of_burst": achieved / burst if achieved else None,
                         "peak_source": peak_src + ", bf16 sustained", "flop_per_expansion": FLOP_PER_EXPANSION[board],
                         "mma_passes_per_k_step": passes_per_eval,
                         "raw_mma_tflops": achieved * passes_per_eval if achieved else None,
                         "tower_ms_per_round": m["tower_ms"] / (steps * rounds), "tree_ms_per_round": m["tree_ms"] / (steps * rounds),
                         "kernel_share_of_step": m["tower_ms"] / (m["tower_ms"] + m["tree_ms"]) if m["tower_ms"] else None,
                         "traffic": traffic}}


def run_legs(a, local, stream, _cabi):
    import torch
    from alpha_omok_b200.model import seeded_state_dict
    legs = {}
    steps = max(1, a.leg_steps)
    # ---- config 3: 4096 parallel 15x15 self-play games, 400 sims/move, random-init PVNet(10,5,128,15)
    G, S = a.games, a.sims
    eng = _cabi.Engine(board_size=15, num_mcts=S, max_games=G, seed=2000, device=local, stream=stream.cuda_stream)
    eng.load_state_dict(seeded_state_dict(0, 10, 5, 128, 15))
    eng.selfplay_begin(G, first_key=0, recycle=True)
    m = _timed_steps(eng, S, steps, 1, stream, torch)
    legs["board15"] = _leg_line("board15", f"BASELINE config 3: {G} parallel 15x15 self-play games, {S} sims/move, PVNet 10x128 random-init, "
                                "single-pass fp16 tower (tower_stag_kernel<15>)", 15, m, steps, S,
                                traffic=TOWER_DRAM_BYTES_PER_LAUNCH.get((15, G, "fp16")))
    eng.close()
    # ---- trained-net self-play: the shipped checkpoint needs the hi/lo split tower (3 MMAs per k-step) for 1e-4
    eng = _cabi.Engine(board_size=9, num_mcts=S, max_games=G, seed=2001, device=local, stream=stream.cuda_stream)
    eng.load_state_dict(_trained_state_dict())
    mode = eng.choose_nn_precision()
    eng.selfplay_begin(G, first_key=0, recycle=True)
    m = _timed_steps(eng, S, steps, 1, stream, torch)
    legs["trained_selfplay"] = _leg_line("trained_selfplay", f"{G} parallel 9x9 self-play games, {S} sims/move, the reference's shipped trained "
                                         f"checkpoint; tower mode picked by the 1e-4 probe: {'fp16 hi/lo split (AO_NN_FP16X3)' if mode == 1 else 'single-pass fp16'}",
                                         9, m, steps, S, passes_per_eval=3 if mode == 1 else 1,
                                         traffic=TOWER_DRAM_BYTES_PER_LAUNCH.get((9, G, "split" if mode == 1 else "fp16")))
    eng.close()
    # ---- config 5: arena, 1024 concurrent matches, 800 sims/move, trained (split tower) vs random-init (fp16 tower)
    M, SA = 1024, 800
    eng = _cabi.Engine(board_size=9, num_mcts=SA, max_games=2 * M, noise=False, seed=2002, device=local, stream=stream.cuda_stream)
    eng.load_state_dict(_trained_state_dict(), which=0)
    eng.load_state_dict(seeded_state_dict(1, 10, 5, 128, 9), which=1)
    pm, em = eng.choose_nn_precision(which=0), eng.choose_nn_precision(which=1)
    eng.arena_begin(M, first_key=0, matches_per_slot=1 << 20, keep_records=False)   # steady state: slots keep playing
    m = _timed_steps(eng, SA, steps, 1, stream, torch)
    leg = _leg_line("arena", f"BASELINE config 5: eval_main arena, {M} concurrent 9x9 matches, {SA} sims/move, noise off, tau 0; player = shipped "
                    f"trained checkpoint (tower mode {pm}), enemy = random-init PVNet seed 1 (tower mode {em}); both networks in one "
                    "engine, two tower launches per round", 9, m, steps, SA)
    # all 1024 matches played to the end: winners
    from alpha_omok_b200 import arena as arena_mod, replay
    eng.arena_begin(M, first_key=5000, matches_per_slot=1, keep_records=True)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        f0.record(stream)
        st = eng.selfplay_rounds(SA)
        while st["running"]:
            st = eng.selfplay_rounds(SA)
        f1.record(stream)
    torch.cuda.synchronize()
    recs = arena_mod.decode_match_records(replay.device_stream_records(eng), 9)
    outc = [r["outcome"] for r in recs]
    leg["full_matches"] = {"matches": M, "seconds": f0.elapsed_time(f1) * 1e-3, "expansions_per_s": st["sims"] / (f0.elapsed_time(f1) * 1e-3),
                           "player_win": outc.count("player"), "enemy_win": outc.count("enemy"), "draw": outc.count("draw"),
                           "mean_plies": float(np.mean([len(r["moves"]) for r in recs])), "tree_overflows": st["errors"]}
    legs["arena"] = leg
    eng.close()
    legs["single_game"] = single_game_leg(local)
    return legs


def single_game_leg(device):
    """BASELINE config 1 on the device: ONE 9x9 self-play game driven move by move through the reference-shaped facade
    (agents.ZeroAgent.get_pi -> utils.get_action -> env.step, main.py:144-196), batch of one leaf per network call - the
    latency-bound end of the path (tower_solo.cu: one game on a cluster of four CTAs).  Wall-clock around whole games,
    host <-> device traffic of every get_pi included."""
    import torch
    from alpha_omok_b200 import agents, model, utils
    from alpha_omok_b200.env import env_small as game
    out = {"workload": "BASELINE config 1: one 9x9 self-play game through ZeroAgent.get_pi (ao_search), random-init PVNet 10x128 (numpy seed 0), "
                       "noise on, tau 1 for 6 plies; sims/move = 40 (config 1) and 400 (the headline's search size); trained_*: the same "
                       "with the reference's shipped checkpoint (hi/lo split tower, tower_mode 1)",
           "kernel": "tower_solo_kernel<9, X3>: one cluster of four CTAs per game, whole search in one launch", "unit": "simulations/s"}
    nets = {}
    net = model.PVNet(10, 5, 128, 9)
    net.load_state_dict(model.seeded_state_dict(0, 10, 5, 128, 9), strict=False)
    nets[""] = net.eval()
    net = model.PVNet(10, 5, 128, 9)    # the reference's shipped checkpoint: the facade picks the hi/lo split tower for it
    net.load_state_dict(_trained_state_dict(), strict=False)
    nets["trained_"] = net.eval()
    for prefix, net in nets.items():
      for sims in (40, 400):
        best = None
        for rep in range(3 if not prefix else 2):
            np.random.seed(rep)
            agent = agents.ZeroAgent(9, sims, 5, noise=True, engine_kwargs={"device": device})
            agent.model = net
            env = game.GameState("text")
            root_id, win_index, t, n_sims = (0,), 0, 0, 0
            agent.get_pi(root_id, 1)          # engine creation + weight upload outside the timed region
            agent.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            while win_index == 0:
                pi = agent.get_pi(root_id, 1 if t < 6 else 0)
                n_sims += sims + (1 if agent.is_real_root else 0)
                action, action_index = utils.get_action(pi)
                root_id += (int(action_index),)
                _, _, win_index, _, _ = env.step(action)
                t += 1
            dt = time.perf_counter() - t0
            mode = agent._engine.nn_precision
            agent._engine.close()
            r = {"sims_per_s": n_sims / dt, "ms_per_move": 1e3 * dt / t, "us_per_sim": 1e6 * dt / n_sims, "moves": t, "sims": n_sims,
                 "tower_mode": int(mode)}
            if best is None or r["sims_per_s"] > best["sims_per_s"]:
                best = r
        out["%ssims%d" % (prefix, sims)] = best
    out["value"] = out["sims40"]["sims_per_s"]
    return out


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
            own = replay.device_records(eng, G)            # pack kernel on the engine stream (synchronised inside)
            gathered = replay.allgather_records(own)       # NCCL all_gather_into_tensor of the fixed-size slabs
            g1.record()
            torch.cuda.synchronize()
            if not timed:
                del gathered
        full.update(allgather_ms=g0.elapsed_time(g1), allgather_bytes=int(gathered.numel()),
                    allgather_ok=bool(torch.equal(gathered[rank * G:(rank + 1) * G], own)))
        del gathered

    eng.close()
    # ---------------- the other BASELINE configs (N = 1 only; extra keys, the headline above is unchanged)
    legs = None
    if world == 1 and not a.no_legs:
        legs = run_legs(a, local, stream, _cabi)

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
            "roofline": {"bound": "tensor",
                         "kernel": "tower_stag_kernel<9, PERSIST> - the persistent self-play kernel: PVNet tower of all leaves "
                                   "of all rounds of a step in ONE launch, the tree step of every pass fused in (head warps); "
                                   "with AO_NO_PERSIST=1: tower_stag_kernel<9> per round + tree_step_kernel",
                         "achieved": achieved, "peak": sustained,
                         "unit": "TFLOP/s", "frac": achieved / sustained, "peak_source": peak_src + ", bf16 sustained",
                         "frac_of_burst": achieved / burst,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch (4096 leaves), from the committed
                         # ncu --set full capture profiles/r02_kernels_ncu_full_summary.txt
                         "traffic": TOWER_DRAM_BYTES_PER_LAUNCH.get((B, G, "fp16")),
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
        if legs is not None:
            line["legs"] = legs
        if not a.no_cpu_baseline and world == 1:   # the host baseline is timed on rank 0 at N = 1 only (idle host cores)
            cores, workers = host_workers()
            line["cpu_baseline"] = cpu_baseline_block(B, S, cores, workers)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
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
