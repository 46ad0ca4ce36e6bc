"""GPU parity tests of the arena path (BASELINE config 5; eval_main.py:54-188 Evaluator, :204-333 main): the device
match loop, the facade route through ao_search and the drop-in Evaluator, each compared ply by ply - visit-count
vectors, moves, winners - with fixtures produced by the UNMODIFIED reference's eval_main.main
(tests/golden/make_golden.py gen_arena) or with the oracle those fixtures pin."""
import numpy as np
import pytest

from helpers import load, synth_eval
from oracle import omok_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cabi():
    from alpha_omok_b200 import _cabi
    return _cabi


def _run(eng, rounds=64):
    st = eng.selfplay_rounds(rounds)
    while st["running"]:
        st = eng.selfplay_rounds(rounds)
    assert st["errors"] == 0
    return st


@pytest.mark.parametrize("name", ["arena_9_synth_s30", "arena_9_synth_s200", "arena_9_random_enemy_s40"])
def test_device_arena_reproduces_eval_main_golden(cabi, name):
    """one arena slot = one run of eval_main.main: consecutive matches, colours swapped, both agents' decision streams
    running on; every ply's visits / move and every winner equal the reference's, including the reused roots the
    searching side had never visited (n == 0) that make up most of these fixtures"""
    from alpha_omok_b200 import arena, replay
    fx = load(name)
    B, sims, n_match = int(fx["B"]), int(fx["sims"]), int(fx["n_match"])
    assert int(fx["n_unvisited_reused_roots"]) > 0
    eng = cabi.Engine(board_size=B, num_mcts=sims, max_games=2, noise=False, seed=int(fx["seed"]),
                      eval_mode=cabi.AO_EVAL_SYNTH)
    eng.arena_begin(1, first_key=0, matches_per_slot=n_match, enemy_random=str(fx["enemy_kind"]) == "random")
    st = _run(eng)
    assert st["games_finished"] == n_match
    recs = arena.decode_match_records(replay.device_stream_records(eng), B)
    assert len(recs) == n_match
    for m, r in enumerate(recs):
        assert r["moves"] == [int(a) for a in fx[f"moves{m}"]], m
        assert np.array_equal(r["visits"], fx[f"visits{m}"].astype(np.uint32)), m
        assert r["winner"] == int(fx[f"winner{m}"]) and r["outcome"] == str(fx[f"outcome{m}"]), m
        assert r["player_black"] == (m % 2 == 0)
    eng.close()


def _oracle_side(B, sims, seed, key, salt, random_agent=False):
    st = O.DecisionStream(seed, key)
    if random_agent:
        return O.OracleRandomAgent(B, st)
    return O.OracleZeroAgent(B, sims, lambda mv: synth_eval(mv, B * B, salt), st, noise=False)


@pytest.mark.parametrize("enemy_random", [False, True])
def test_device_arena_many_slots_vs_oracle(cabi, enemy_random):
    """24 concurrent slots x 2 matches, each slot checked bit-exactly against the oracle's eval_main loop on the slot's own
    decision streams (keys first_key + 2*slot + side); odd slots start with the player as white"""
    from alpha_omok_b200 import arena, replay
    B, sims, seed, S, MPS, K0 = 9, 20, 123, 24, 2, 10
    eng = cabi.Engine(board_size=B, num_mcts=sims, max_games=2 * S, noise=False, seed=seed, eval_mode=cabi.AO_EVAL_SYNTH)
    eng.arena_begin(S, first_key=K0, matches_per_slot=MPS, enemy_random=enemy_random)
    st = _run(eng)
    assert st["games_finished"] == S * MPS and st["running"] == 0
    recs = arena.decode_match_records(replay.device_stream_records(eng), B)
    eng.close()
    for slot in range(0, S, 5):
        ora = O.arena_matches(B, _oracle_side(B, sims, seed, K0 + 2 * slot, 0),
                              _oracle_side(B, sims, seed, K0 + 2 * slot + 1, 1, enemy_random), MPS,
                              player_black_first=(slot % 2 == 0))
        for k in range(MPS):
            r, o = recs[slot * MPS + k], ora[k]
            assert r["moves"] == o["moves"] and r["winner"] == o["winner"] and r["outcome"] == o["outcome"], (slot, k)
            assert np.array_equal(r["visits"], np.stack(o["visits"]).astype(np.uint32)), (slot, k)
            assert r["player_black"] == o["player_black"]


def test_device_arena_with_root_noise_vs_oracle(cabi):
    """noise=True in the arena (not the reference's default, but `ZeroAgent(noise=...)` is a constructor argument): a
    reused root the side never visited is expanded WITH root noise because leaf_id == root_id (agents.py:191-204), a
    visited reused root gets the re-mix (agents.py:95-103)"""
    from alpha_omok_b200 import arena, replay
    B, A, sims, seed = 9, 81, 40, 9
    eng = cabi.Engine(board_size=B, num_mcts=sims, max_games=2, noise=True, seed=seed, eval_mode=cabi.AO_EVAL_SYNTH,
                      noise_mode=cabi.AO_NOISE_TAPE)
    tapes = [O.make_gamma_tape(seed, k, A + 2, A, 10 / A) for k in range(2)]
    eng.set_gamma_tape(0, tapes[0])  # slot 0 = player side of match 0, slot M + 0 = 1 = enemy side
    eng.set_gamma_tape(1, tapes[1])
    eng.arena_begin(1, first_key=0, matches_per_slot=1)
    _run(eng)
    rec = arena.decode_match_records(replay.device_stream_records(eng), B)[0]
    eng.close()
    sides = [O.OracleZeroAgent(B, sims, lambda mv, s=s: synth_eval(mv, A, s), O.DecisionStream(seed, s, tapes[s]),
                               noise=True) for s in range(2)]
    o = O.arena_matches(B, sides[0], sides[1], 1)[0]
    assert rec["moves"] == o["moves"] and rec["winner"] == o["winner"]
    assert np.array_equal(rec["visits"], np.stack(o["visits"]).astype(np.uint32))


def test_facade_search_unvisited_reply_vs_oracle(cabi):
    """the same situations through ao_search (= ZeroAgent.get_pi of the drop-in agents): two agents with their own
    trees, replies forced onto cells the searching side never visited (reused root with n == 0, agents.py:93-111) and
    onto visited ones; the oracle's handling of exactly this is pinned by arena_9_synth_s60_forced.npz"""
    B, A, sims, seed = 9, 81, 60, 44
    eng = cabi.Engine(board_size=B, num_mcts=sims, max_games=2, noise=False, seed=seed, eval_mode=cabi.AO_EVAL_SYNTH)
    eng.games_reset([0, 1], keys=[0, 1])
    ora = [O.OracleZeroAgent(B, sims, lambda mv: synth_eval(mv, A), O.DecisionStream(seed, k), noise=False)
           for k in range(2)]
    root, n0 = (0,), 0
    for ply in range(14):
        s = ply % 2
        vis, pri, real = eng.search([s], [root])
        ora[s].get_pi(root, 1)
        assert np.array_equal(vis[0], ora[s].visit.astype(np.uint32)), ply
        assert np.array_equal(pri[0], ora[s].policy), ply
        assert bool(real[0]) == ora[s].is_real_root, ply
        n0 += (not real[0]) and int(vis[0].sum()) == sims - 1
        if ply % 3 == 1:   # a cell nobody looked at: the highest empty one
            a = max(c for c in range(A) if c not in root[1:] and vis[0][c] == 0)
        else:
            a = int(np.argmax(vis[0]))
        root = root + (a,)
        if O.check_win(O.get_board(root, B), 5):
            break
    assert n0 >= 3
    eng.close()


def test_play_matches_device_arena_trained_vs_random_init(cabi):
    """BASELINE config 5 at toy size through arena.play_matches: trained checkpoint (hi/lo split tower) against a
    random-init net (single-pass fp16), both weight sets in one engine; the trained side wins every match and the
    records are consistent (winner = utils.check_win of the final position, players alternate)"""
    import torch
    from alpha_omok_b200 import arena, model
    from oracle import pvnet_ref
    z = load("trained_9x9_180927")
    player = model.PVNet(10, 5, 128, 9)
    player.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=False)
    enemy = model.PVNet(10, 5, 128, 9)
    enemy.load_state_dict(pvnet_ref.make_state_dict(1, 10, 5, 128, 9), strict=False)
    res = arena.play_matches(player, enemy, n_matches=16, num_mcts=48, seed=2, return_records=True)
    assert res["unfinished"] == 0 and res["player_win"] == 16, res
    assert res["nn_precision"] == (cabi.AO_NN_FP16X3, cabi.AO_NN_FP16)
    for i, r in enumerate(res["records"]):
        assert r["player_black"] == (i % 2 == 0)
        assert O.check_win(O.get_board((0,) + tuple(r["moves"]), 9), 5) == r["winner"]
        assert O.check_win(O.get_board((0,) + tuple(r["moves"][:-1]), 9), 5) == 0
        assert (r["visits"].sum(axis=1) >= 47).all()


def test_evaluator_dropin_vs_oracle_arena(cabi, monkeypatch):
    """arena.Evaluator / arena.run_matches (the drop-in for eval_main.Evaluator / eval_main.main built from the single-
    game ZeroAgent facade): two matches, every ply's visits and every move equal the oracle's eval_main loop.  The
    facade draws argmax_onehot's tie-break from numpy's global generator like the reference; the test routes that draw
    and the oracle's to one shared host stream."""
    from alpha_omok_b200 import agents, arena
    B, A, sims, seed = 9, 81, 50, 0
    host = O.DecisionStream(777, 0)
    monkeypatch.setattr(np.random, "choice", lambda a, size=None, replace=True, p=None: host.choice(int(a)))
    ev = arena.Evaluator(board_size=B, n_mcts_player=sims, n_mcts_enemy=sims, n_mcts_monitor=sims, n_blocks=2,
                         engine_kwargs=dict(eval_mode=cabi.AO_EVAL_SYNTH))
    from oracle import pvnet_ref
    sd = pvnet_ref.make_state_dict(3, 2, 5, 128, B)
    ev.set_agents(sd, sd, sd)
    log = []
    for name in ("player", "enemy"):
        ag = getattr(ev, name)
        orig = ag.get_pi

        def get_pi(root_id, tau, ag=ag, orig=orig, name=name):
            pi = orig(root_id, tau)
            log.append((name, tuple(int(x) for x in root_id), ag.visit.astype(np.int64).copy(), bool(ag.is_real_root)))
            return pi

        ag.get_pi = get_pi
    result, pe, ee = arena.run_matches(ev, n_match=2)
    got_host_draws = host.ctr

    class Side(O.OracleZeroAgent):  # facade agents take decision-stream key = episode number (bumped by reset())
        def __init__(self, hs):
            super().__init__(B, sims, lambda mv: synth_eval(mv, A), O.DecisionStream(seed, 0), noise=False)
            self.host_stream = hs
            self.episode = 0

        def reset(self):
            super().reset()
            if hasattr(self, "episode"):  # not during __init__
                self.episode += 1
                self.stream = O.DecisionStream(seed, self.episode)

    hs = O.DecisionStream(777, 0)
    p, e = Side(hs), Side(hs)
    ora = O.arena_matches(B, p, e, 2)
    flat = [(mv, v, rr, who) for m in ora for mv, v, rr, who in zip(m["moves"], m["visits"], m["real_root"], m["movers"])]
    assert len(flat) == len(log)
    for i, ((who, root, vis, real), (mv, v, rr, owho)) in enumerate(zip(log, flat)):
        assert who == owho and np.array_equal(vis, v) and real == rr, i
    assert hs.ctr == got_host_draws
    outcomes = [m["outcome"] for m in ora]
    assert result == {"Player": outcomes.count("player"), "Enemy": outcomes.count("enemy"), "Draw": outcomes.count("draw")}
    assert (pe, ee) == arena.elo_sequence(outcomes)[:2]


# ------------------------------------------------------------------------------------------------ PUCT / UCT agents
def test_rollout_agents_reproduce_reference_golden(cabi):
    """PUCTAgent / UCTAgent.get_pi of the UNMODIFIED reference (rollout_agents_s40.npz: 9x9 early / mid-game and 15x15
    late positions): the device search - fresh tree, num_mcts + 1 simulations, UCB / PUCT selection in float64, one
    uniformly random play-out per leaf - gives the same child visit counts and w sums and consumes the same number of
    decision-stream blocks (checked through the next draw)"""
    fx = load("rollout_agents_s40")
    sims, seed = int(fx["sims"]), int(fx["seed"])
    for i in range(int(fx["n_cases"])):
        kind, B = str(fx[f"kind{i}"]), int(fx[f"B{i}"])
        root = tuple(int(a) for a in fx[f"root{i}"])
        eng = cabi.Engine(board_size=B, num_mcts=sims, max_games=1, noise=False, seed=seed, eval_mode=cabi.AO_EVAL_SYNTH)
        eng.games_reset([0], keys=[i])
        vis, w = eng.rollout_search(kind, [0], [root], sims)
        assert np.array_equal(vis[0], fx[f"visits{i}"].astype(np.uint32)), (i, kind, B)
        assert np.array_equal(w[0], fx[f"w{i}"]), (i, kind, B)
        # a second search in the same slot continues the slot's decision stream exactly where the reference's would
        ora = O.OracleRolloutAgent(kind, B, sims, O.DecisionStream(seed, i))
        ora.stream.ctr = int(fx[f"draws{i}"]) - (1 if _final_tiebreak_drew(fx, i, kind) else 0)
        ora.get_pi(root)
        vis2, _ = eng.rollout_search(kind, [0], [root], sims)
        assert np.array_equal(vis2[0], ora.visit.astype(np.uint32)), (i, kind, B)
        eng.close()


def _final_tiebreak_drew(fx, i, kind):
    """did get_pi's closing arg-max draw from the stream (more than one maximum)? The device search stops before it."""
    vis, w = fx[f"visits{i}"].astype(np.float64), fx[f"w{i}"].astype(np.float64)
    if kind == "puct":
        score = vis
    else:
        root = set(int(a) for a in fx[f"root{i}"][1:])
        score = np.full(len(vis), -np.inf)
        for a in range(len(vis)):
            if a not in root:
                score[a] = w[a] / vis[a] if vis[a] > 0 else 0.0
    return int((score == score.max()).sum()) > 1


@pytest.mark.parametrize("name", ["arena_9_puct_enemy_s30", "arena_9_uct_enemy_s30"])
def test_device_arena_with_rollout_enemy_reproduces_eval_main_golden(cabi, name):
    """eval_main.main with enemy = 'puct' / 'uct' (eval_main.py:73-78), run unmodified: the device arena slices the
    play-out agent's search over the lock-step rounds and still reproduces every ply"""
    from alpha_omok_b200 import arena, replay
    fx = load(name)
    B, sims, n_match, kind = int(fx["B"]), int(fx["sims"]), int(fx["n_match"]), str(fx["enemy_kind"])
    eng = cabi.Engine(board_size=B, num_mcts=sims, max_games=2, noise=False, seed=int(fx["seed"]),
                      eval_mode=cabi.AO_EVAL_SYNTH)
    eng.arena_begin(1, first_key=0, matches_per_slot=n_match, enemy_kind=kind)
    st = _run(eng)
    assert st["games_finished"] == n_match
    recs = arena.decode_match_records(replay.device_stream_records(eng), B)
    for m, r in enumerate(recs):
        assert r["moves"] == [int(a) for a in fx[f"moves{m}"]], m
        assert np.array_equal(r["visits"], fx[f"visits{m}"].astype(np.uint32)), m
        assert r["winner"] == int(fx[f"winner{m}"]) and r["outcome"] == str(fx[f"outcome{m}"]), m
    eng.close()


def test_evaluator_accepts_rollout_agents_by_name(cabi, monkeypatch):
    """Evaluator.set_agents('puct' | 'uct') (eval_main.py:73-78) through the drop-in facade agents: one match each against
    the oracle's eval_main loop with the facades' host draws routed to a shared stream"""
    from alpha_omok_b200 import agents, arena
    from oracle import pvnet_ref
    B, A, sims = 9, 81, 24
    sd = pvnet_ref.make_state_dict(3, 2, 5, 128, B)
    for kind in ("puct", "uct"):
        host = O.DecisionStream(555, 1)
        monkeypatch.setattr(np.random, "choice", lambda a, size=None, replace=True, p=None, host=host: host.choice(int(a)))
        ev = arena.Evaluator(board_size=B, n_mcts_player=sims, n_mcts_enemy=sims, n_mcts_monitor=sims, n_blocks=2,
                             engine_kwargs=dict(eval_mode=cabi.AO_EVAL_SYNTH))
        ev.set_agents(sd, kind, sd)
        assert type(ev.enemy).__name__ == ("PUCTAgent" if kind == "puct" else "UCTAgent")
        moves = []
        orig = ev.get_action

        def get_action(*a, orig=orig, moves=moves):
            action, idx = orig(*a)
            moves.append(int(idx))
            return action, idx

        ev.get_action = get_action
        result, _, _ = arena.run_matches(ev, n_match=1)
        hs = O.DecisionStream(555, 1)
        p = O.OracleZeroAgent(B, sims, lambda mv: synth_eval(mv, A), O.DecisionStream(0, 0), noise=False)
        p.host_stream = hs
        e = O.OracleRolloutAgent(kind, B, sims, O.DecisionStream(0, 0))
        e.host_stream = hs
        o = O.arena_matches(B, p, e, 1)[0]
        assert moves == o["moves"], kind
        assert sum(result.values()) == 1
