"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

The reference (/root/reference/2_AlphaOmok: agents.py, utils.py, model.py, env/env_small.py, env/env_regular.py) is
imported as-is behind a 2-file pygame stub. Its random decisions (np.random.choice / np.random.dirichlet, call sites
agents.py:97,163,194 and utils.py:192,202) are routed through oracle.omok_oracle.DecisionStream so that the oracle
and the device can consume the very same decisions. While generating, every fixture is cross-checked against the
oracle restatement (assertion failure = oracle bug), then written as small .npz files which the CPU tests
(tests/test_oracle_golden.py) and the GPU tests replay without /root/reference.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("ALPHA_OMOK_REF", "/root/reference/2_AlphaOmok")

from oracle import omok_oracle as O  # noqa: E402
from oracle import pvnet_ref  # noqa: E402


FLASK_STUB = """
class Blueprint:
    def __init__(self, *a, **k): pass
    def route(self, *a, **k): return lambda f: f
class Flask(Blueprint):
    def register_blueprint(self, *a, **k): pass
    def run(self, *a, **k): pass
def render_template(*a, **k): return ""
def jsonify(*a, **k): return a[0] if a else k
class _Req:
    args = {}
request = _Req()
"""


def import_reference():
    stub = tempfile.mkdtemp(prefix="pygame_stub_")
    os.makedirs(os.path.join(stub, "pygame"))
    with open(os.path.join(stub, "pygame", "__init__.py"), "w") as f:
        f.write("")
    with open(os.path.join(stub, "pygame", "locals.py"), "w") as f:
        f.write("QUIT = 12\n")
    with open(os.path.join(stub, "flask.py"), "w") as f:   # eval_main.py:14 / webapi.py:1-2 (dashboard wiring only)
        f.write(FLASK_STUB)
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import agents, model, utils  # noqa: E401
    from env import env_regular, env_small
    import eval_main
    agents.PRINT_MCTS = False
    return types.SimpleNamespace(agents=agents, model=model, utils=utils, env_small=env_small, env_regular=env_regular,
                                 eval_main=eval_main)


# ------------------------------------------------------------------------------------------------ synthetic NN
def synth_eval(moves, A, salt=0):
    """Hash 'network': exact float32 outputs, identical here, in the oracle tests and in csrc (EVAL_SYNTH).
    `salt` selects a different 'network' (the arena's two sides)."""
    hh = 0xCBF29CE484222325
    for m in moves[1:]:
        hh = ((hh ^ (int(m) + 1)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    lo, hi = hh & 0xFFFFFFFF, hh >> 32
    pol = np.empty(A, np.float32)
    for a in range(A):
        w = O.philox4x32((a, 0, lo, hi), (0x5EED, 0x0A0A + salt))
        pol[a] = np.float32(((w[0] >> 8) + 1) * 2.0 ** -24)
    w = O.philox4x32((0xFFFF, 0, lo, hi), (0x5EED, 0x0A0A + salt))
    val = np.float32((w[1] >> 8) * 2.0 ** -23 - 1.0)
    return pol, val


class StubModel:
    """nn.Module-like object for agents.ZeroAgent.model (agents.py:173-178)."""

    def __init__(self, fn):
        self.fn = fn
        self.cur_leaf = None
        self.log = []

    def eval(self):
        return self

    def __call__(self, x):
        p, v = self.fn(self.cur_leaf, x)
        self.log.append((self.cur_leaf, np.asarray(p, np.float32), np.float32(v)))
        return torch.from_numpy(np.asarray(p, np.float32))[None], torch.from_numpy(np.asarray([v], np.float32))


class RngPatch:
    def __init__(self):
        self.stream = None

    def choice(self, a, size=None, replace=True, p=None):
        if p is None:
            return self.stream.choice(int(a))
        return self.stream.choice_p(np.asarray(p))

    def dirichlet(self, alpha, size=None):
        return self.stream.dirichlet(len(alpha))


PATCH = RngPatch()


def ref_self_play(R, B, sims, stream, model_stub, noise, tau_thres, max_moves, inplanes=5):
    """main.py:132-250 (one episode) driven with the reference's own agents/utils/env objects."""
    PATCH.stream = stream
    env_mod = R.env_small if B == 9 else R.env_regular
    agent = R.agents.ZeroAgent(B, sims, inplanes, noise=noise)
    agent.model = model_stub
    orig = agent._expansion_evaluation

    def wrapped(leaf_id, win_index):
        model_stub.cur_leaf = leaf_id
        return orig(leaf_id, win_index)

    agent._expansion_evaluation = wrapped
    env = env_mod.GameState("text")
    root_id, win_index, t = (0,), 0, 0
    visits, pis = [], []
    while win_index == 0 and (max_moves is None or t < max_moves):
        tau = 1 if t < tau_thres else 0
        pi = agent.get_pi(root_id, tau)
        visits.append(agent.visit.astype(np.int64))
        pis.append(pi.copy())
        action, action_index = R.utils.get_action(pi)
        root_id += (int(action_index),)
        _, _, win_index, _, _ = env.step(action)
        t += 1
    return dict(moves=list(root_id[1:]), visits=visits, pis=pis, winner=int(win_index))


def gen_mcts_game(R, name, B, sims, seed, game, noise, tau_thres, max_moves, nn_kind, sd=None):
    A = B * B
    tape = O.make_gamma_tape(seed, game, A + 2, A, 10 / A)
    if nn_kind == "synth":
        fn = lambda leaf, x: synth_eval(leaf, A)  # noqa: E731
    else:
        net = R.model.PVNet(pvnet_ref.n_blocks_of(sd), 5, 128, B)
        net.load_state_dict(sd, strict=False)
        net.eval()

        def fn(leaf, x):
            with torch.no_grad():
                p, v = net(x)
            return p[0].numpy(), v[0].item()
    stub = StubModel(fn)
    ref = ref_self_play(R, B, sims, O.DecisionStream(seed, game, tape), stub, noise, tau_thres, max_moves)
    # ---- cross-check the oracle restatement on the same decisions and the same NN outputs
    log = {}
    for leaf, p, v in stub.log:
        log[leaf] = (p, v)
    ora = O.self_play_game(B, sims, lambda mv: log[mv], O.DecisionStream(seed, game, tape), tau_thres=tau_thres,
                           noise=noise, max_moves=max_moves)
    assert ora["moves"] == ref["moves"], (name, ora["moves"], ref["moves"])
    assert ora["winner"] == ref["winner"]
    for a, b in zip(ora["visits"], ref["visits"]):
        assert np.array_equal(a, b), name
    for a, b in zip(ora["pis"], ref["pis"]):
        assert np.array_equal(a, b), name
    out = dict(B=B, sims=sims, seed=seed, game=game, noise=int(noise), tau_thres=tau_thres,
               max_moves=-1 if max_moves is None else max_moves, nn_kind=nn_kind,
               moves=np.asarray(ref["moves"], np.int32), winner=ref["winner"],
               visits=np.asarray(ref["visits"], np.int32), gamma_tape=tape[:len(ref["moves"]) + 2])
    if nn_kind != "synth":  # ship the NN outputs of the non-terminal expansions, in call order
        keep = [(leaf, p, v) for leaf, p, v in stub.log
                if R.utils.check_win(R.utils.get_board(leaf, B), 5) == 0]
        out["nn_policy"] = np.stack([p for _, p, _ in keep])
        out["nn_value"] = np.asarray([v for _, _, v in keep], np.float32)
        out["nn_leaf_len"] = np.asarray([len(leaf) for leaf, _, _ in keep], np.int32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {len(ref['moves'])} moves, winner {ref['winner']}, {len(stub.log)} evals  OK (oracle == reference)")


def gen_arena(R, name, sims, seed, n_match, enemy_kind="zero", forced=None):
    """eval_main.main (eval_main.py:204-333) run UNMODIFIED (only `evaluator.set_agents`, which would read checkpoint
    paths, is replaced by pre-built agents): two reference ZeroAgent(noise=False) sides with their own trees (or a
    ZeroAgent against the reference RandomAgent), `get_pi(tau=0)` -> `argmax_onehot` -> `del_parents`, colours swapped
    after every match.  Each agent draws from its own decision stream (key = side), each side has its own synthetic
    network (salt = side).  `forced` = {(match, ply): action}: the move actually played is overridden (the searching
    side's choice is discarded) - used to push the opponent onto a reply its tree never visited."""
    import contextlib
    import io
    E = R.eval_main
    B, A = 9, 81
    forced = dict(forced or {})
    streams = {"player": O.DecisionStream(seed, 0), "enemy": O.DecisionStream(seed, 1)}
    log = []          # (agent name, root_id, visits, is_real_root)
    actions = []      # action index per ply
    boundaries = []   # len(actions) at every evaluator.reset()

    def make(side):
        if side == "enemy" and enemy_kind == "random":
            agent = R.agents.RandomAgent(B)
        elif side == "enemy" and enemy_kind in ("puct", "uct"):
            agent = (R.agents.PUCTAgent if enemy_kind == "puct" else R.agents.UCTAgent)(B, sims)
        else:
            agent = R.agents.ZeroAgent(B, sims, 5, noise=False)
            salt = 0 if side == "player" else 1
            stub = StubModel(lambda leaf, x, salt=salt: synth_eval(leaf, A, salt))
            agent.model = stub
            orig_ee = agent._expansion_evaluation

            def wrapped_ee(leaf_id, win_index, orig_ee=orig_ee, stub=stub):
                stub.cur_leaf = leaf_id
                return orig_ee(leaf_id, win_index)

            agent._expansion_evaluation = wrapped_ee
        orig = agent.get_pi

        def get_pi(*a, side=side, agent=agent, orig=orig, **k):
            PATCH.stream = streams[side]
            pi = orig(*a, **k)
            zero = isinstance(agent, R.agents.ZeroAgent)
            vis = np.zeros(A, np.int64)
            if zero:
                vis = agent.visit.astype(np.int64).copy()
            elif isinstance(agent, (R.agents.PUCTAgent, R.agents.UCTAgent)):
                for act in agent.tree[agent.root_id]["child"]:
                    vis[act] = agent.tree[agent.root_id + (act,)]["n"]
            log.append((side, tuple(a[0]), vis, bool(agent.is_real_root) if zero else True))
            return pi

        agent.get_pi = get_pi
        return agent

    ev = E.evaluator
    ev.player, ev.enemy = make("player"), make("enemy")
    ev.monitor = R.agents.ZeroAgent(B, sims, 5, noise=False)
    ev.monitor.model = StubModel(lambda leaf, x: (np.full(A, 1 / A, np.float32), np.float32(0)))
    ev.env = R.env_small.GameState("text")
    ev.set_agents = lambda *a: None
    orig_get_action = ev.get_action
    match_no = [0]

    def get_action(root_id, board, turn, enemy_turn):
        action, idx = orig_get_action(root_id, board, turn, enemy_turn)
        key = (match_no[0], len(actions) - (boundaries[-1] if boundaries else 0))
        if key in forced:
            idx = forced[key]
            action = np.zeros(A)
            action[idx] = 1
        actions.append(int(idx))
        return action, idx

    ev.get_action = get_action
    orig_reset = ev.reset

    def reset():
        boundaries.append(len(actions))
        match_no[0] += 1
        orig_reset()

    ev.reset = reset
    E.N_MATCH = n_match
    with contextlib.redirect_stdout(io.StringIO()):
        E.main()
    assert len(boundaries) == n_match and len(log) == len(actions)

    # ---- cross-check the oracle restatement (same streams, same synthetic networks)
    def omake(side):
        st = O.DecisionStream(seed, 0 if side == "player" else 1)
        if side == "enemy" and enemy_kind == "random":
            return O.OracleRandomAgent(B, st)
        if side == "enemy" and enemy_kind in ("puct", "uct"):
            return O.OracleRolloutAgent(enemy_kind, B, sims, st)
        salt = 0 if side == "player" else 1
        return O.OracleZeroAgent(B, sims, lambda mv, salt=salt: synth_eval(mv, A, salt), st, noise=False)

    ora = O.arena_matches(B, omake("player"), omake("enemy"), n_match, forced=forced)
    out = dict(B=B, sims=sims, seed=seed, n_match=n_match, enemy_kind=enemy_kind,
               forced=np.asarray([[m, p, a] for (m, p), a in sorted(forced.items())], np.int32).reshape(-1, 3))
    lo = 0
    n0 = 0
    for m, hi in enumerate(boundaries):
        mv = actions[lo:hi]
        vis = np.stack([log[i][2] for i in range(lo, hi)])
        movers = [log[i][0] for i in range(lo, hi)]
        real = [log[i][3] for i in range(lo, hi)]
        winner = R.utils.check_win(R.utils.get_board((0,) + tuple(mv), B), 5)
        o = ora[m]
        assert o["moves"] == mv and o["winner"] == winner and o["movers"] == movers and o["real_root"] == real, (name, m)
        assert np.array_equal(np.stack(o["visits"]), vis), (name, m)
        # reused roots the searching side had never visited (agents.py:93-111 with n == 0): num_mcts - 1 child visits
        n0 += sum(1 for i in range(len(mv)) if (movers[i] != "enemy" or enemy_kind == "zero")
                  and not real[i] and vis[i].sum() == sims - 1)
        out[f"moves{m}"] = np.asarray(mv, np.int32)
        out[f"visits{m}"] = vis.astype(np.int32)
        out[f"player_mover{m}"] = np.asarray([x == "player" for x in movers], np.int8)
        out[f"real_root{m}"] = np.asarray(real, np.int8)
        out[f"winner{m}"] = winner
        out[f"outcome{m}"] = o["outcome"]
        lo = hi
    assert n0 > 0 or enemy_kind in ("puct", "uct"), "no reused root with n == 0 in this golden"
    out["n_unvisited_reused_roots"] = n0
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {n_match} matches, plies {[int(b) for b in np.diff([0] + boundaries)]}, "
          f"{n0} reused roots with n == 0  OK (oracle == reference eval_main.main)")


def gen_rollout_agents(R, name, sims, seed):
    """PUCTAgent / UCTAgent.get_pi (agents.py:263-634) of the unmodified reference on a handful of 9x9 and 15x15 positions:
    visit counts and w sums of the root's children, the move, and the number of decision-stream blocks consumed (every
    selection tie-break and every play-out move is one draw).  Cross-checked against oracle.OracleRolloutAgent."""
    import contextlib
    import io
    rs = np.random.RandomState(seed)
    out = dict(sims=sims, seed=seed)
    cases = []
    for kind, cls in (("puct", R.agents.PUCTAgent), ("uct", R.agents.UCTAgent)):
        for B, lo, hi in ((9, 0, 12), (9, 30, 60), (15, 100, 180)):
            A = B * B
            while True:
                k = int(rs.randint(lo, hi))
                root = (0,) + tuple(int(x) for x in rs.permutation(A)[:k])
                if R.utils.check_win(R.utils.get_board(root, B), 5) == 0:
                    break
            key = len(cases)
            PATCH.stream = O.DecisionStream(seed, key)
            agent = cls(B, sims)
            with contextlib.redirect_stdout(io.StringIO()):
                pi = agent.get_pi(root, R.utils.get_board(root, B), R.utils.get_turn(root), 0)
            vis, w = np.zeros(A, np.int64), np.zeros(A)
            for a in agent.tree[root]["child"]:
                vis[a], w[a] = agent.tree[root + (a,)]["n"], agent.tree[root + (a,)]["w"]
            ora = O.OracleRolloutAgent(kind, B, sims, O.DecisionStream(seed, key))
            po = ora.get_pi(root)
            assert np.array_equal(pi, po) and np.array_equal(vis, ora.visit) and PATCH.stream.ctr == ora.stream.ctr, (kind, B)
            i = len(cases)
            cases.append((kind, B))
            out[f"kind{i}"], out[f"B{i}"] = kind, B
            out[f"root{i}"] = np.asarray(root, np.int16)
            out[f"visits{i}"], out[f"w{i}"] = vis.astype(np.int32), w.astype(np.float32)
            out[f"move{i}"], out[f"draws{i}"] = int(np.argmax(pi)), PATCH.stream.ctr
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {len(cases)} searches ({sims} sims)  OK (oracle == reference)")


def fill_dashboard(gi, pa, ea, B, player_agent, enemy_agent):
    """deterministic contents for the dashboard objects (used identically by tests/test_cabi_and_host.py)"""
    rs = np.random.RandomState(99)
    board = np.zeros(B * B)
    cells = rs.permutation(B * B)[:23]
    board[cells[0::2]], board[cells[1::2]] = 1, -1
    gi.game_board = board.reshape(B, B)
    gi.win_index, gi.curr_turn, gi.enemy_turn, gi.action_index, gi.game_status = 0, 1, 1, int(cells[-1]), 0
    pa.agent, ea.agent = player_agent, enemy_agent
    pa.visit, pa.p = rs.randint(0, 50, B * B).astype("float"), rs.dirichlet(np.ones(B * B))
    ea.visit, ea.p = rs.randint(0, 50, B * B).astype("float"), rs.dirichlet(np.ones(B * B))
    for mv, v in ((1, 0.25), (3, -0.5), (5, 0.875)):
        pa.add_value(mv, v)
    for mv, v in ((2, -0.125), (4, 0.0)):
        ea.add_value(mv, v)
    player_agent.message, enemy_agent.message = "simulation: 800\r", "Hello"


def gen_dashboard(R):
    """payloads of the reference's /periodic_status and /prompt_status routes (webapi.py:28-76) for a fixed state of its
    GameInfo / AgentInfo objects (info/*.py) - the only thing the web dashboard ever reads from the arena"""
    import json
    import webapi
    B = 9
    fill_dashboard(webapi.game_info, webapi.player_agent_info, webapi.enemy_agent_info, B,
                   R.agents.ZeroAgent(B, 800, 5, noise=False), R.agents.RandomAgent(B))
    out = {"periodic_status": webapi.periodic_status(), "prompt_status": webapi.prompt_status()}
    webapi.player_agent_info.clear_values()
    out["after_clear"] = webapi.periodic_status()
    with open(os.path.join(HERE, "dashboard_feed.json"), "w") as f:
        json.dump(out, f)
    print("dashboard_feed.json OK", len(out["periodic_status"]), "keys")


def pick_forced_replies(sims, seed, plies=(4, 9)):
    """{(0, ply): action}: at the given plies of match 0 the move played becomes the highest empty cell instead of the
    searcher's choice - with 60 simulations over ~75 children a cell neither side's tree has ever visited, so the
    opponent's next root is a reused root with n == 0 and the mover's own next root an unvisited child as well"""
    B, A = 9, 81
    forced = {}
    for p in plies:
        def mk(side):
            return O.OracleZeroAgent(B, sims, lambda mv, s=side: synth_eval(mv, A, s), O.DecisionStream(seed, side),
                                     noise=False)
        m = O.arena_matches(B, mk(0), mk(1), 1, forced=forced)[0]
        assert len(m["moves"]) > p + 2, "match too short for the forced plies"
        empty = [c for c in range(A) if c not in m["moves"][:p]]
        forced[(0, p)] = max(c for c in empty if c != m["moves"][p])
    return forced


def find_long_15x15_game(sims, seed0, min_plies):
    """scan decision-stream seeds with the oracle for a 15x15 self-play game that enters the CPython-set child-order
    regime (>= 149 stones); the reference then replays the seed found"""
    A = 225
    for seed in range(seed0, seed0 + 400):
        tape = O.make_gamma_tape(seed, 0, A + 2, A, 10 / A)
        g = O.self_play_game(15, sims, lambda mv: synth_eval(mv, A), O.DecisionStream(seed, 0, tape))
        if len(g["moves"]) >= min_plies:
            return seed, len(g["moves"])
    raise RuntimeError("no long 15x15 game found")


def gen_late_roots(R, name, B, sims, seed, n_roots, lo, hi):
    """ZeroAgent.get_pi (agents.py:60-132) of the unmodified reference on late-game roots - stone counts in the regime
    where `legal_actions` comes out in CPython's hash-table order (63-79 stones on 9x9, 149-223 on 15x15),
    which decides the child a tie-break index and a Dirichlet component refer to.  Every root is searched, then the
    position two plies deeper (own most-visited move + most-visited reply: a reused, re-noised root)."""
    A = B * B
    rs = np.random.RandomState(seed)
    roots = []
    while len(roots) < n_roots:
        k = int(rs.randint(lo, hi))
        mv = (0,) + tuple(int(x) for x in rs.permutation(A)[:k])
        if R.utils.check_win(R.utils.get_board(mv, B), 5) == 0:
            assert R.utils.legal_actions(mv, B) != sorted(R.utils.legal_actions(mv, B))
            roots.append(mv)
    out = dict(B=B, sims=sims, seed=seed, n_roots=n_roots)
    for g, root in enumerate(roots):
        tape = O.make_gamma_tape(seed, g, 4, A, 10 / A)
        PATCH.stream = O.DecisionStream(seed, g, tape)
        agent = R.agents.ZeroAgent(B, sims, 5, noise=True)
        stub = StubModel(lambda leaf, x: synth_eval(leaf, A))
        agent.model = stub
        orig = agent._expansion_evaluation

        def wrapped(leaf_id, win_index, orig=orig, stub=stub):
            stub.cur_leaf = leaf_id
            return orig(leaf_id, win_index)

        agent._expansion_evaluation = wrapped
        ora = O.OracleZeroAgent(B, sims, lambda mv: synth_eval(mv, A), O.DecisionStream(seed, g, tape), noise=True)
        cur, searched, vis, pri, real = root, [], [], [], []
        for step in range(2):
            agent.get_pi(cur, 1)
            ora.get_pi(cur, 1)
            assert np.array_equal(agent.visit, ora.visit) and np.array_equal(agent.policy, ora.policy), (name, g, step)
            assert agent.is_real_root == ora.is_real_root
            searched.append(np.asarray(cur + (-1,) * (A + 1 - len(cur)), np.int16))
            vis.append(agent.visit.astype(np.int32))
            pri.append(agent.policy.copy())
            real.append(int(agent.is_real_root))
            a = int(np.argmax(agent.visit))
            sub = agent.tree[cur + (a,)]["child"]
            if not sub:
                break
            b = max(sub, key=lambda c: agent.tree[cur + (a, c)]["n"])   # first maximum in child order
            cur = cur + (a, int(b))
            if R.utils.check_win(R.utils.get_board(cur, B), 5) != 0:
                break
        out[f"roots{g}"] = np.stack(searched)
        out[f"visits{g}"] = np.stack(vis)
        out[f"priors{g}"] = np.stack(pri)
        out[f"real{g}"] = np.asarray(real, np.int8)
        out[f"tape{g}"] = tape
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {n_roots} late roots ({lo}-{hi} stones), {sims} sims  OK (oracle == reference)")


def gen_round2(R, parts=("arena", "trained", "long15", "late", "rollout", "dashboard")):
    """round-2 fixtures: arena (eval_main.main), trained-checkpoint self-play, a full-length 15x15 game, late roots"""
    if "arena" in parts:
        gen_arena(R, "arena_9_synth_s30", sims=30, seed=41, n_match=2)
        gen_arena(R, "arena_9_synth_s200", sims=200, seed=42, n_match=2)
        gen_arena(R, "arena_9_synth_s60_forced", sims=60, seed=44, n_match=1, forced=pick_forced_replies(60, 44))
        gen_arena(R, "arena_9_random_enemy_s40", sims=40, seed=43, n_match=2, enemy_kind="random")
    if "trained" in parts:
        z = np.load(os.path.join(HERE, "trained_9x9_180927.npz"))
        sd = {k: torch.from_numpy(z[k]) for k in z.files}
        gen_mcts_game(R, "mcts_9_trained_s40", 9, 40, seed=16, game=0, noise=True, tau_thres=6, max_moves=None,
                      nn_kind="pvnet", sd=sd)
    if "long15" in parts:
        # a whole 15x15 self-play game that runs into the >= 149-stone regime: 2 simulations per move keep the play
        # close to uniform (games of two searching players end long before), the search code runs all the same
        seed, plies = find_long_15x15_game(sims=2, seed0=100, min_plies=152)
        print("15x15 long game: seed", seed, "plies", plies)
        gen_mcts_game(R, "mcts_15_synth_long", 15, 2, seed=seed, game=0, noise=True, tau_thres=6, max_moves=None,
                      nn_kind="synth")
    if "late" in parts:
        gen_late_roots(R, "search_15_late_roots", 15, 60, seed=51, n_roots=4, lo=149, hi=215)
    if "dashboard" in parts:
        gen_dashboard(R)
    if "rollout" in parts:
        gen_rollout_agents(R, "rollout_agents_s40", sims=40, seed=61)
        gen_arena(R, "arena_9_puct_enemy_s30", sims=30, seed=45, n_match=2, enemy_kind="puct")
        gen_arena(R, "arena_9_uct_enemy_s30", sims=30, seed=46, n_match=2, enemy_kind="uct")


def gen_rules(R):
    rs = np.random.RandomState(1234)
    out = {}
    for B in (9, 15):
        A = B * B
        boards, wins = [], []
        for k in range(300):
            fill = rs.randint(0, A + 1)
            cells = rs.permutation(A)[:fill]
            b = np.zeros(A)
            b[cells] = rs.choice([-1.0, 1.0], size=fill)
            if k % 3 == 0 and fill:  # plant lines (fives, overlines, both colours)
                for _ in range(rs.randint(1, 3)):
                    L = rs.randint(4, 8)
                    dy, dx = [(0, 1), (1, 0), (1, 1), (1, -1)][rs.randint(4)]
                    y0, x0 = rs.randint(B), rs.randint(B)
                    col = rs.choice([-1.0, 1.0])
                    for i in range(L):
                        y, x = y0 + i * dy, x0 + i * dx
                        if 0 <= y < B and 0 <= x < B:
                            b[y * B + x] = col
            if k % 10 == 9:  # full boards (draw candidates)
                b = rs.choice([-1.0, 1.0], size=A)
            b = b.reshape(B, B)
            w = R.utils.check_win(b, 5)
            assert O.check_win(b, 5) == w
            boards.append(b.astype(np.int8))
            wins.append(w)
        out[f"boards{B}"] = np.stack(boards)
        out[f"wins{B}"] = np.asarray(wins, np.int8)
        # IDs: states, legal-action order at every stone count
        ids, states, legal = [], [], []
        for s in list(range(0, A)) + list(range(A - 20, A)):
            mv = (0,) + tuple(int(x) for x in rs.permutation(A)[:s])
            la = R.utils.legal_actions(mv, B)
            assert O.legal_actions(mv, B) == la, (B, s)
            st = R.utils.get_state_pt(mv, B, 5)
            assert np.array_equal(O.get_state_pt(mv, B, 5), st)
            assert np.array_equal(O.get_board(mv, B), R.utils.get_board(mv, B))
            assert O.get_turn(mv) == R.utils.get_turn(mv)
            ids.append(np.asarray(mv + (-1,) * (A + 1 - len(mv)), np.int16))
            states.append(np.packbits(st.astype(np.uint8).reshape(-1)))
            legal.append(np.asarray(la + [-1] * (A - len(la)), np.int16))
        out[f"ids{B}"] = np.stack(ids)
        out[f"states{B}"] = np.stack(states)
        out[f"legal{B}"] = np.stack(legal)
    # numpy pairwise-sum restatement vs ndarray.sum
    for n in (81, 225):
        for _ in range(2000):
            v = (rs.rand(n).astype(np.float32) ** 8).astype(np.float64) * (rs.rand(n) < 0.8)
            assert O.np_pairwise_sum(v) == v.sum()
    np.savez_compressed(os.path.join(HERE, "rules.npz"), **out)
    print("rules.npz OK")


def gen_nn(R):
    rs = np.random.RandomState(7)
    for B, n_block, jitter, seed, name in ((9, 10, False, 0, "nn_9_init"), (9, 10, True, 1, "nn_9_jitter"),
                                            (15, 10, False, 0, "nn_15_init"), (9, 2, True, 3, "nn_9_small")):
        A = B * B
        sd = pvnet_ref.make_state_dict(seed, n_block, 5, 128, B, bn_jitter=jitter)
        net = R.model.PVNet(n_block, 5, 128, B)
        missing = net.load_state_dict(sd, strict=False)
        assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
        net.eval()
        ids = []
        for _ in range(24):
            s = rs.randint(0, min(A, 60))
            ids.append((0,) + tuple(int(x) for x in rs.permutation(A)[:s]))
        x = torch.tensor(np.stack([R.utils.get_state_pt(i, B, 5) for i in ids])).float()
        with torch.no_grad():
            p, v = net(x)
        p2, v2 = pvnet_ref.pvnet_forward(sd, x)
        assert torch.allclose(p, p2, atol=1e-6) and torch.allclose(v, v2, atol=1e-6)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), B=B, n_block=n_block, jitter=int(jitter), seed=seed,
                            ids=np.stack([np.asarray(i + (-1,) * (A + 1 - len(i)), np.int16) for i in ids]),
                            p=p.numpy(), v=v.numpy())
        print(name, "OK  max|p-p_ref|", float((p - p2).abs().max()))


def gen_train(R):
    """main.py:253-305 on a fixed train_memory: the reference's PVNet (train-mode BN), loss and Adam, restated from
    main.py:286-305 because main.py itself cannot be imported (import-time side effects)."""
    from torch import optim
    from torch.utils.data import DataLoader
    B, n_block, A = 9, 2, 81
    rs = np.random.RandomState(21)
    sd = pvnet_ref.make_state_dict(3, n_block, 5, 128, B, bn_jitter=True)
    net = R.model.PVNet(n_block, 5, 128, B)
    net.load_state_dict(sd, strict=False)
    optimizer = optim.Adam(net.parameters(), lr=2e-4, weight_decay=0, eps=1e-6)   # main.py:85
    train_memory = []
    for _ in range(80):  # 2 full batches of 32 + a short one of 16
        k = rs.randint(0, 60)
        rid = (0,) + tuple(int(x) for x in rs.permutation(A)[:k])
        pi = rs.dirichlet(np.full(A, 0.3))
        train_memory.append((R.utils.get_state_pt(rid, B, 5), pi, float(rs.randint(-1, 2))))
    dataloader = DataLoader(train_memory, batch_size=32, shuffle=False, pin_memory=False)   # main.py:266-269
    net.train()
    losses = []
    for epoch in range(2):
        for i, (s, pi, z) in enumerate(dataloader):
            s_batch, pi_batch, z_batch = s.float(), pi.float(), z.float()
            p_batch, v_batch = net(s_batch)
            v_loss = (v_batch - z_batch).pow(2).mean()
            p_loss = -(pi_batch * p_batch.log()).sum(dim=-1).mean()
            loss = v_loss + p_loss
            losses.append((loss.item(), v_loss.item(), p_loss.item()))
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
    out_sd = net.state_dict()
    np.savez_compressed(os.path.join(HERE, "train_9_small.npz"), B=B, n_block=n_block, sd_seed=3,
                        states=np.stack([m[0] for m in train_memory]).astype(np.float32),
                        pis=np.stack([m[1] for m in train_memory]), zs=np.asarray([m[2] for m in train_memory]),
                        losses=np.asarray(losses, np.float64),
                        conv1_weight=out_sd["conv1.weight"].numpy(), bn1_running_mean=out_sd["bn1.running_mean"].numpy(),
                        policy_fc_bias=out_sd["policy_head.policy_fc.bias"].numpy(),
                        value_fc2_weight=out_sd["value_head.value_fc2.weight"].numpy(),
                        abs_sums=np.asarray([float(v.double().abs().sum()) for k, v in out_sd.items()
                                             if not k.endswith("num_batches_tracked")]))
    print("train_9_small OK  losses", losses[0], "->", losses[-1])


def gen_trained_checkpoint():
    """the reference's shipped trained network (data/180927_9400_297233_step_model.pickle, 121 fp32 tensors of
    PVNet(10,5,128,9) without num_batches_tracked keys) re-saved as .npz: the 'trained' side of BASELINE config 5 and the
    realistic-weights numerics fixture.  Weights are data, not code; nothing else is taken over."""
    src = os.path.join(REF, "data", "180927_9400_297233_step_model.pickle")
    sd = torch.load(src, map_location="cpu")
    np.savez_compressed(os.path.join(HERE, "trained_9x9_180927.npz"), **{k: v.numpy() for k, v in sd.items()})
    print("trained_9x9_180927.npz OK", len(sd), "tensors")


def gen_utils_aux(R):
    """the reference's non-hot-path utils (render_str, valid_actions, get_reward, get_state_tf) on random boards"""
    import contextlib, io, json
    rs = np.random.RandomState(5)
    cases = []
    for B in (9, 15):
        for trial in range(12):
            k = int(rs.randint(0, B * B)) if trial else 0
            cells = [int(c) for c in rs.permutation(B * B)[:k]]
            board = np.zeros((B, B))
            for t, c in enumerate(cells):
                board[c // B, c % B] = 1 if t % 2 == 0 else -1
            last = cells[-1] if k else None
            if trial % 5 == 4:
                last = int(rs.randint(0, B * B))
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                R.utils.render_str(board, B, last)
            rid = (0,) + tuple(cells)
            cases.append(dict(B=B, cells=cells, last=last, render=buf.getvalue(),
                              valid=[[list(a[0]), a[1]] for a in R.utils.valid_actions(board)],
                              reward=[R.utils.get_reward(w, rid) for w in (0, 1, 2, 3)],
                              state_tf_sum=[float((R.utils.get_state_tf(rid, t, B, 5) * np.arange(1, 6)).sum()) for t in (0, 1)],
                              state_tf=R.utils.get_state_tf(rid, 0, B, 5).astype(int).tolist() if trial < 3 else None))
    with open(os.path.join(HERE, "utils_aux.json"), "w") as f:
        json.dump(cases, f)
    print("utils_aux.json OK", len(cases))


def main():
    R = import_reference()
    np.random.choice = PATCH.choice
    np.random.dirichlet = PATCH.dirichlet
    if "--only-train" in sys.argv:
        gen_train(R)
        return
    if "--only-utils-aux" in sys.argv:
        gen_utils_aux(R)
        return
    if "--only-trained" in sys.argv:
        gen_trained_checkpoint()
        return
    if "--only-round2" in sys.argv:
        parts = [a[len("--parts="):].split(",") for a in sys.argv if a.startswith("--parts=")]
        gen_round2(R, *parts)
        return
    gen_rules(R)
    gen_nn(R)
    gen_train(R)
    gen_utils_aux(R)
    gen_trained_checkpoint()
    gen_mcts_game(R, "mcts_9_synth_s40", 9, 40, seed=11, game=0, noise=True, tau_thres=6, max_moves=None, nn_kind="synth")
    gen_mcts_game(R, "mcts_9_synth_s400", 9, 400, seed=12, game=3, noise=True, tau_thres=6, max_moves=None, nn_kind="synth")
    gen_mcts_game(R, "mcts_9_synth_nonoise", 9, 60, seed=13, game=1, noise=False, tau_thres=0, max_moves=None, nn_kind="synth")
    gen_mcts_game(R, "mcts_15_synth_s50", 15, 50, seed=14, game=2, noise=True, tau_thres=6, max_moves=40, nn_kind="synth")
    sd = pvnet_ref.make_state_dict(0, 10, 5, 128, 9)
    gen_mcts_game(R, "mcts_9_pvnet_s40", 9, 40, seed=15, game=0, noise=True, tau_thres=6, max_moves=None, nn_kind="pvnet", sd=sd)
    gen_round2(R)


if __name__ == "__main__":
    main()
