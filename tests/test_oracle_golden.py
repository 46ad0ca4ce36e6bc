"""CPU tests: the oracle restatement replayed against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py). No GPU, no /root/reference."""
import numpy as np
import pytest
import torch

from helpers import load, oracle_game_from_fixture, synth_eval, unpad_id
from oracle import omok_oracle as O
from oracle import pvnet_ref


@pytest.mark.parametrize("B", [9, 15])
def test_rules_golden(B):
    fx = load("rules")
    A = B * B
    for b, w in zip(fx[f"boards{B}"], fx[f"wins{B}"]):
        assert O.check_win(b.astype(np.float64), 5) == int(w)
    states = np.unpackbits(fx[f"states{B}"], axis=1)[:, :5 * A].reshape(-1, 5, B, B)
    for row, st, la in zip(fx[f"ids{B}"], states, fx[f"legal{B}"]):
        mv = unpad_id(row)
        assert np.array_equal(O.get_state_pt(mv, B, 5), st.astype(np.float64))
        assert O.legal_actions(mv, B) == [int(a) for a in la if a >= 0]


def test_legal_actions_matches_this_cpython():
    """the set-order restatement agrees with the running interpreter's own set (what the reference executes)"""
    rs = np.random.RandomState(0)
    for B in (9, 15):
        A = B * B
        for s in list(range(0, A, 7)) + list(range(A - 25, A)):
            mv = (0,) + tuple(int(x) for x in rs.permutation(A)[:s])
            assert O.legal_actions(mv, B) == list({a for a in range(A)} - set(mv[1:]))


def test_pairwise_sum_matches_numpy():
    rs = np.random.RandomState(1)
    for n in (3, 8, 81, 128, 129, 225):
        for _ in range(300):
            v = (rs.rand(n).astype(np.float32) ** 6).astype(np.float64) * (rs.rand(n) < 0.7)
            assert O.np_pairwise_sum(v) == v.sum()


@pytest.mark.parametrize("name", ["nn_9_init", "nn_9_jitter", "nn_15_init", "nn_9_small"])
def test_pvnet_ref_golden(name):
    fx = load(name)
    B = int(fx["B"])
    sd = pvnet_ref.make_state_dict(int(fx["seed"]), int(fx["n_block"]), 5, 128, B, bn_jitter=bool(fx["jitter"]))
    x = torch.from_numpy(np.stack([O.get_state_pt(unpad_id(r), B, 5) for r in fx["ids"]]).astype(np.float32))
    p, v = pvnet_ref.pvnet_forward(sd, x)
    assert np.abs(p.numpy() - fx["p"]).max() < 2e-6
    assert np.abs(v.numpy() - fx["v"]).max() < 2e-6


@pytest.mark.parametrize("name", ["mcts_9_synth_s40", "mcts_9_synth_nonoise", "mcts_15_synth_s50", "mcts_9_synth_s400",
                                  "mcts_15_synth_long"])
def test_mcts_synth_golden(name):
    fx = load(name)
    ora = oracle_game_from_fixture(fx)
    assert ora["moves"] == [int(m) for m in fx["moves"]]
    assert ora["winner"] == int(fx["winner"])
    assert np.array_equal(np.asarray(ora["visits"]), fx["visits"])


@pytest.mark.parametrize("name", ["mcts_9_pvnet_s40", "mcts_9_trained_s40"])
def test_mcts_pvnet_golden_nn_replay(name):
    """reference game driven by the real PVNet (random-init; the shipped trained checkpoint): the oracle replays the
    logged NN outputs and must reproduce it"""
    fx = load(name)
    it = iter(range(len(fx["nn_value"])))

    def evaluate(mv):
        k = next(it)
        assert len(mv) == int(fx["nn_leaf_len"][k])
        return fx["nn_policy"][k], fx["nn_value"][k]

    ora = oracle_game_from_fixture(fx, evaluate)
    assert ora["moves"] == [int(m) for m in fx["moves"]]
    assert ora["winner"] == int(fx["winner"])
    assert np.array_equal(np.asarray(ora["visits"]), fx["visits"])


@pytest.mark.parametrize("name", ["arena_9_synth_s30", "arena_9_synth_s200", "arena_9_synth_s60_forced",
                                  "arena_9_random_enemy_s40"])
def test_arena_golden(name):
    """eval_main.main run unmodified (two agents with own trees, tau = 0, colours swapped, forced replies onto unvisited
    cells in one fixture): the oracle's arena loop reproduces every ply"""
    fx = load(name)
    B, A, sims, seed, n_match = int(fx["B"]), int(fx["B"]) ** 2, int(fx["sims"]), int(fx["seed"]), int(fx["n_match"])
    forced = {(int(m), int(p)): int(a) for m, p, a in fx["forced"]}
    player = O.OracleZeroAgent(B, sims, lambda mv: synth_eval(mv, A, 0), O.DecisionStream(seed, 0), noise=False)
    if str(fx["enemy_kind"]) == "random":
        enemy = O.OracleRandomAgent(B, O.DecisionStream(seed, 1))
    else:
        enemy = O.OracleZeroAgent(B, sims, lambda mv: synth_eval(mv, A, 1), O.DecisionStream(seed, 1), noise=False)
    ora = O.arena_matches(B, player, enemy, n_match, forced=forced)
    n0 = 0
    for m, o in enumerate(ora):
        assert o["moves"] == [int(a) for a in fx[f"moves{m}"]]
        assert np.array_equal(np.stack(o["visits"]), fx[f"visits{m}"])
        assert o["winner"] == int(fx[f"winner{m}"]) and o["outcome"] == str(fx[f"outcome{m}"])
        assert [x == "player" for x in o["movers"]] == [bool(x) for x in fx[f"player_mover{m}"]]
        assert o["real_root"] == [bool(x) for x in fx[f"real_root{m}"]]
        n0 += sum(1 for v, r in zip(o["visits"], o["real_root"]) if not r and v.sum() == sims - 1)
    assert n0 >= int(fx["n_unvisited_reused_roots"]) > 0


def test_late_roots_golden():
    fx = load("search_15_late_roots")
    B, A, sims, seed = int(fx["B"]), int(fx["B"]) ** 2, int(fx["sims"]), int(fx["seed"])
    for g in range(int(fx["n_roots"])):
        agent = O.OracleZeroAgent(B, sims, lambda mv: synth_eval(mv, A), O.DecisionStream(seed, g, fx[f"tape{g}"]))
        for step, row in enumerate(fx[f"roots{g}"]):
            agent.get_pi(unpad_id(row), 1)
            assert np.array_equal(agent.visit, fx[f"visits{g}"][step])
            assert np.array_equal(agent.policy, fx[f"priors{g}"][step])


def test_decision_stream_properties():
    s = O.DecisionStream(7, 3)
    assert s.choice(1) == 0 and s.ctr == 0  # no consumption for a single candidate
    xs = [s.choice(5) for _ in range(2000)]
    assert set(xs) == {0, 1, 2, 3, 4}
    pi = np.zeros(81)
    pi[17] = 1.0
    assert all(O.DecisionStream(7, g).choice_p(pi) == 17 for g in range(50))
    # Philox4x32-10 known-answer vectors (Random123 kat_vectors)
    assert O.philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert O.philox4x32((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)


def test_env_step_semantics():
    env = O.OracleGameState(9)
    onehot = np.zeros(81)
    onehot[40] = 1
    board, valid, win, turn, a = env.step(onehot)
    assert valid and win == 0 and turn == 1 and a == 40 and board[4, 4] == 1
    board, valid, win, turn, a = env.step(onehot)  # occupied: flagged invalid but overwritten (env_small.py:161-176)
    assert not valid and board[4, 4] == -1 and turn == 0


def test_reference_copy_reproduces_config1_and_the_oracle():
    """oracle/_ref (the unmodified reference assembled by oracle/make_ref.py - what bench.py times as the CPU baseline)
    plays BASELINE config 1 to the survey's golden result (39 moves, black wins, 1561 simulations, visit-count hash
    ceddffce5bfb2897) and its files are byte-identical to the recorded upstream hashes"""
    from oracle import make_ref, ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref not assembled (needs /root/reference: python oracle/make_ref.py)")
    assert make_ref.verify()
    r = ref_runner.run_config1_subprocess(threads=4)
    assert r["kind"] == "reference"
    assert (r["moves"], r["winner"], r["sims"]) == (39, 1, 1561)
    assert r["visit_sha256_16"] == "ceddffce5bfb2897"


def test_rollout_agents_golden():
    """PUCTAgent / UCTAgent of the unmodified reference (agents.py:263-634): visits, w sums, move and stream consumption"""
    fx = load("rollout_agents_s40")
    sims, seed = int(fx["sims"]), int(fx["seed"])
    for i in range(int(fx["n_cases"])):
        kind, B = str(fx[f"kind{i}"]), int(fx[f"B{i}"])
        root = tuple(int(a) for a in fx[f"root{i}"])
        ag = O.OracleRolloutAgent(kind, B, sims, O.DecisionStream(seed, i))
        pi = ag.get_pi(root)
        assert np.array_equal(ag.visit, fx[f"visits{i}"]) and int(np.argmax(pi)) == int(fx[f"move{i}"])
        assert ag.stream.ctr == int(fx[f"draws{i}"])


@pytest.mark.parametrize("name", ["arena_9_puct_enemy_s30", "arena_9_uct_enemy_s30"])
def test_arena_rollout_enemy_golden(name):
    fx = load(name)
    B, A, sims, seed, n_match = 9, 81, int(fx["sims"]), int(fx["seed"]), int(fx["n_match"])
    player = O.OracleZeroAgent(B, sims, lambda mv: synth_eval(mv, A, 0), O.DecisionStream(seed, 0), noise=False)
    enemy = O.OracleRolloutAgent(str(fx["enemy_kind"]), B, sims, O.DecisionStream(seed, 1))
    for m, o in enumerate(O.arena_matches(B, player, enemy, n_match)):
        assert o["moves"] == [int(a) for a in fx[f"moves{m}"]]
        assert np.array_equal(np.stack(o["visits"]), fx[f"visits{m}"])
        assert o["winner"] == int(fx[f"winner{m}"]) and o["outcome"] == str(fx[f"outcome{m}"])
