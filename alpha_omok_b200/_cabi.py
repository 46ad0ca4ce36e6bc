"""ctypes binding of include/alpha_omok_b200.h. No compute happens in Python: every hot-path call below lands in the
CUDA library. The library is built in-tree by alpha_omok_b200._build; loading fails loudly if it is missing or if
no B200 is present when a compute entry point is called (there is no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

AO_EVAL_PVNET, AO_EVAL_SYNTH = 0, 1
AO_NOISE_DEVICE, AO_NOISE_TAPE = 0, 1
AO_NN_FP16, AO_NN_FP16X3, AO_NN_FP16_1CTA, AO_NN_FP16_LOCKSTEP = 0, 1, 2, 3
AO_SIDE_ZERO, AO_SIDE_RANDOM, AO_SIDE_PUCT, AO_SIDE_UCT = 0, 1, 2, 3
SIDE_KINDS = {"zero": AO_SIDE_ZERO, "random": AO_SIDE_RANDOM, "puct": AO_SIDE_PUCT, "uct": AO_SIDE_UCT}


class AoConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("board_size", C.c_int32), ("inplanes", C.c_int32), ("planes", C.c_int32),
        ("n_blocks", C.c_int32), ("num_mcts", C.c_int32), ("noise", C.c_int32), ("tau_thres", C.c_int32),
        ("max_games", C.c_int32), ("node_cap", C.c_int32), ("eval_mode", C.c_int32), ("noise_mode", C.c_int32),
        ("nn_precision", C.c_int32), ("nn_log_cap", C.c_int32), ("c_puct", C.c_double), ("alpha", C.c_double),
        ("seed", C.c_uint64), ("stream", C.c_void_p),
    ]


EXPORTS = [  # every symbol include/alpha_omok_b200.h declares (checked by tests/test_cabi_and_host.py)
    "ao_last_error", "ao_engine_create", "ao_engine_destroy", "ao_load_weights", "ao_games_reset",
    "ao_set_gamma_tape", "ao_search", "ao_nn_forward", "ao_selfplay_begin", "ao_selfplay_begin_mode",
    "ao_selfplay_stream_begin", "ao_selfplay_stream_records_dev", "ao_selfplay_rounds", "ao_selfplay_rounds_timed",
    "ao_launch_count", "ao_set_nn_precision", "ao_selfplay_fetch", "ao_get_nn_log", "ao_records_dev",
    "ao_records_pack", "ao_augment_records_dev", "ao_replay_extend_dev", "ao_replay_gather_dev", "ao_synchronize",
    "ao_check_win", "ao_encode_state", "ao_legal_actions",
    "ao_load_weights_set", "ao_nn_forward_set", "ao_set_nn_precision_set", "ao_arena_begin", "ao_rollout_search",
    "ao_set_log_table",
]
PROBE_EXPORTS = ["ao_tower_debug", "ao_umma_probe", "ao_umma_probe_masked", "ao_umma_rate"]  # alpha_omok_b200_probe.h

_lib = None


class AoError(RuntimeError):
    pass


def _bind(L):
    L.ao_last_error.restype = C.c_char_p
    vp, i32, u32 = C.c_void_p, C.c_int32, C.c_uint32
    L.ao_engine_create.argtypes = [C.POINTER(AoConfig), C.POINTER(vp)]
    L.ao_engine_destroy.argtypes = [vp]
    L.ao_load_weights.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(C.c_int64)]
    L.ao_load_weights_set.argtypes = [vp, i32, i32, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(C.c_int64)]
    L.ao_nn_forward_set.argtypes = [vp, i32, vp, i32, vp, vp]
    L.ao_set_nn_precision_set.argtypes = [vp, i32, i32]
    L.ao_arena_begin.argtypes = [vp, i32, u32, i32, i32, i32, i32, i32, i32]
    L.ao_rollout_search.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, vp]
    L.ao_set_log_table.argtypes = [vp, vp, i32]
    L.ao_games_reset.argtypes = [vp, vp, i32, vp]
    L.ao_set_gamma_tape.argtypes = [vp, i32, vp, i32]
    L.ao_search.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.ao_nn_forward.argtypes = [vp, vp, i32, vp, vp]
    L.ao_selfplay_begin.argtypes = [vp, i32, u32]
    L.ao_selfplay_begin_mode.argtypes = [vp, i32, u32, i32]
    L.ao_selfplay_rounds.argtypes = [vp, i32, vp]
    L.ao_selfplay_stream_begin.argtypes = [vp, i32, u32, i32]
    L.ao_selfplay_stream_records_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(i32)]
    L.ao_selfplay_rounds_timed.argtypes = [vp, i32, vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.ao_set_nn_precision.argtypes = [vp, i32]
    L.ao_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.ao_selfplay_fetch.argtypes = [vp, i32, vp, vp, vp, vp]
    L.ao_get_nn_log.argtypes = [vp, i32, vp, vp, i32, C.POINTER(i32)]
    L.ao_records_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.ao_records_pack.argtypes = [vp, i32]
    L.ao_augment_records_dev.argtypes = [vp, i32, i32, i32, vp, vp, vp, C.c_longlong, C.POINTER(C.c_longlong), vp]
    ll, pll = C.c_longlong, C.POINTER(C.c_longlong)
    L.ao_replay_extend_dev.argtypes = [vp, i32, i32, i32, vp, vp, vp, ll, pll, pll, pll, vp]
    L.ao_replay_gather_dev.argtypes = [vp, vp, vp, ll, ll, vp, ll, i32, vp, vp, vp, vp]
    L.ao_synchronize.argtypes = [vp]
    L.ao_check_win.argtypes = [vp, i32, i32, vp]
    L.ao_encode_state.argtypes = [vp, vp, i32, i32, vp]
    L.ao_legal_actions.argtypes = [vp, vp, i32, i32, vp]
    for name in EXPORTS:
        if name != "ao_last_error":
            getattr(L, name).restype = C.c_int
    return L


def lib():
    """Load (building if necessary) libalpha_omok_b200.so - the product library.

    Profiling tools set AO_USE_PROBE_LIB=1 to run the same Python API on libalpha_omok_b200_probe.so (the sources
    compiled with -DAO_PROBE: tower cycle counters + AO_TOWER_XFLAGS experiments); tests and bench never do."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("AO_USE_PROBE_LIB") == "1":
        _lib = probe_lib()
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        path = _build.build()
    _lib = _bind(C.CDLL(path))
    return _lib


_probe = None


def probe_lib():
    """libalpha_omok_b200_probe.so (include/alpha_omok_b200_probe.h): product entry points + instrumentation."""
    global _probe
    if _probe is None:
        path = _build.build(probe=True)  # stamp-checked: rebuilt only when the sources changed
        L = _bind(C.CDLL(path))
        vp, i32 = C.c_void_p, C.c_int32
        L.ao_tower_debug.argtypes = [vp, i32, vp]
        L.ao_umma_probe.argtypes = [vp, i32, vp, vp, vp, i32, i32, vp]
        L.ao_umma_probe_masked.argtypes = [vp, i32, vp, vp, vp, i32, i32, vp, vp]
        for name in PROBE_EXPORTS:
            getattr(L, name).restype = C.c_int
        _probe = L
    return _probe


def default_device(model=None):
    """CUDA ordinal an engine created on behalf of `model` should live on: the device of the model's parameters when
    they are on a GPU, otherwise torch's current device (so a torchrun rank that called torch.cuda.set_device(local_rank)
    gets its own GPU), otherwise 0."""
    try:
        import torch
        if model is not None:
            for p in model.parameters():
                if p.device.type == "cuda":
                    return p.device.index if p.device.index is not None else torch.cuda.current_device()
                break
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def check(rc):
    if rc != 0:
        raise AoError(f"alpha_omok_b200 C-ABI error {rc}: {lib().ao_last_error().decode()}")


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pad_ids(root_ids, A):
    """list of ID tuples -> (int16 [n][A+1] padded with -1, int32 lens)."""
    n = len(root_ids)
    ids = np.full((n, A + 1), -1, np.int16)
    lens = np.empty(n, np.int32)
    for i, r in enumerate(root_ids):
        ids[i, :len(r)] = r
        lens[i] = len(r)
    return ids, lens


class Engine:
    """Thin RAII wrapper over ao_engine (one CUDA device + stream)."""

    def __init__(self, board_size=9, num_mcts=400, max_games=1, noise=True, tau_thres=6, n_blocks=10, inplanes=5,
                 planes=128, seed=0, device=0, node_cap=0, eval_mode=AO_EVAL_PVNET, noise_mode=AO_NOISE_DEVICE,
                 nn_precision=AO_NN_FP16, nn_log_cap=0, c_puct=5.0, alpha=0.0, stream=None):
        self.B, self.A, self.G = board_size, board_size * board_size, max_games
        self.num_mcts = num_mcts
        self.device = int(device)
        self.nn_precision = self.nn_precision_enemy = nn_precision
        cfg = AoConfig(device, board_size, inplanes, planes, n_blocks, num_mcts, int(bool(noise)), tau_thres,
                       max_games, node_cap, eval_mode, noise_mode, nn_precision, nn_log_cap, float(c_puct),
                       float(alpha), seed, stream)
        self._h = C.c_void_p()
        self._lib = lib()  # keep the library alive for __del__ at interpreter shutdown
        check(self._lib.ao_engine_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and getattr(self, "_lib", None) is not None:
            self._lib.ao_engine_destroy(h)
            self._h = None

    __del__ = close

    def load_state_dict(self, state_dict, which=0):
        """fp32 tensors straight from nn.Module.state_dict() (BN folded inside the library).
        which = 1 fills the second weight set (the arena's enemy, eval_main.py:87-101)."""
        names, arrs = [], []
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked"):
                continue
            a = v.detach().cpu().float().contiguous().numpy() if hasattr(v, "detach") else np.asarray(v, np.float32)
            names.append(k.encode())
            arrs.append(np.ascontiguousarray(a, np.float32))
        n = len(names)
        c_names = (C.c_char_p * n)(*names)
        c_ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        c_numel = (C.c_int64 * n)(*[a.size for a in arrs])
        check(lib().ao_load_weights_set(self._h, which, n, c_names, c_ptrs, c_numel))

    def set_nn_precision(self, mode, which=0):
        check(lib().ao_set_nn_precision_set(self._h, which, mode))
        if which == 0:
            self.nn_precision = mode
        else:
            self.nn_precision_enemy = mode

    def choose_nn_precision(self, tol=5e-5, n_probe=48, seed=0, which=0):
        """Pick the cheapest tower mode whose outputs agree with the hi/lo-split mode within `tol` on a set of probe
        positions (the split mode is within 1e-4 of fp32 even on trained nets). Random-init nets stay on
        the single-pass fp16 mode; trained nets switch to AO_NN_FP16X3.  A heuristic, not a proof: the probe is half
        uniformly scattered stones and half clustered ones (each stone next to an earlier one, like real games); `tol`
        is half the contract.  Force a mode with nn_precision=AO_NN_* where the guarantee matters."""
        rs = np.random.RandomState(seed)
        states = np.zeros((n_probe, 5, self.B, self.B), np.float32)
        B = self.B
        for i in range(n_probe):  # legal-looking positions: k stones, alternating colours
            k = int(rs.randint(0, max(1, self.A - 20)))
            if i % 2 == 0:
                cells = rs.permutation(self.A)[:k]
            else:  # clustered: every stone within distance 2 of an earlier one
                cells, taken = [], set()
                c = int(rs.randint(0, self.A))
                for _ in range(k):
                    for _try in range(30):
                        if cells:
                            b = cells[int(rs.randint(0, len(cells)))]
                            y, x = b // B + int(rs.randint(-2, 3)), b % B + int(rs.randint(-2, 3))
                            c = y * B + x if 0 <= y < B and 0 <= x < B else -1
                        if c >= 0 and c not in taken:
                            break
                    else:
                        free = [q for q in range(self.A) if q not in taken]
                        c = free[int(rs.randint(0, len(free)))]
                    cells.append(c)
                    taken.add(c)
                cells = np.asarray(cells, np.int64)
            own, opp = cells[k % 2::2], cells[(k + 1) % 2::2]
            states[i, 2].flat[own] = 1
            states[i, 3].flat[opp] = 1
            states[i, 0], states[i, 1] = states[i, 2], states[i, 3]
            if len(own):
                states[i, 0].flat[own[-1]] = 0
            if len(opp):
                states[i, 1].flat[opp[-1]] = 0
            states[i, 4] = 1.0 if k % 2 == 0 else 0.0
        self.set_nn_precision(AO_NN_FP16X3, which)
        p3, v3 = self.nn_forward(states, which)
        self.set_nn_precision(AO_NN_FP16, which)
        p1, v1 = self.nn_forward(states, which)
        err = max(float(np.abs(p1 - p3).max()), float(np.abs(v1 - v3).max()))
        if err > tol:
            self.set_nn_precision(AO_NN_FP16X3, which)
        return self.nn_precision if which == 0 else self.nn_precision_enemy

    def games_reset(self, game_ids, keys=None):
        ids = np.ascontiguousarray(game_ids, np.int32)
        k = None if keys is None else np.ascontiguousarray(keys, np.uint32)
        check(lib().ao_games_reset(self._h, ptr(ids), len(ids), ptr(k)))

    def set_gamma_tape(self, game_id, tape):
        t = np.ascontiguousarray(tape, np.float64)
        assert t.shape[1] == self.A
        check(lib().ao_set_gamma_tape(self._h, game_id, ptr(t), t.shape[0]))

    def search(self, game_ids, root_ids):
        """-> visits uint32 [n][A], priors float64 [n][A], is_real_root int32 [n]"""
        ids = np.ascontiguousarray(game_ids, np.int32)
        roots, lens = pad_ids(root_ids, self.A)
        n = len(ids)
        visits = np.empty((n, self.A), np.uint32)
        priors = np.empty((n, self.A), np.float64)
        real = np.empty(n, np.int32)
        check(lib().ao_search(self._h, ptr(ids), n, ptr(roots), ptr(lens), ptr(visits), ptr(priors), ptr(real)))
        return visits, priors, real

    def search_raw(self, ids, roots, lens, visits, priors=None, real=None):
        check(lib().ao_search(self._h, ptr(ids), len(ids), ptr(roots), ptr(lens), ptr(visits), ptr(priors), ptr(real)))

    def nn_forward(self, states, which=0):
        s = np.ascontiguousarray(states, np.float32)
        n = s.shape[0]
        p = np.empty((n, self.A), np.float32)
        v = np.empty(n, np.float32)
        check(lib().ao_nn_forward_set(self._h, which, ptr(s), n, ptr(p), ptr(v)))
        return p, v

    def arena_begin(self, n_slots, first_key=0, matches_per_slot=1, enemy_random=False, keep_records=True,
                    n_mcts_player=0, n_mcts_enemy=0, player_kind="zero", enemy_kind=None):
        """eval_main.main's match loop on the device (ao_arena_begin); drive with selfplay_rounds().
        player_kind / enemy_kind: 'zero' | 'random' | 'puct' | 'uct' (eval_main.py:67-84)."""
        if enemy_kind is None:
            enemy_kind = "random" if enemy_random else "zero"
        pk, ek = SIDE_KINDS[player_kind], SIDE_KINDS[enemy_kind]
        if AO_SIDE_UCT in (pk, ek):
            self.ensure_log_table(max(n_mcts_player or self.num_mcts, n_mcts_enemy or self.num_mcts) + 2)
        check(lib().ao_arena_begin(self._h, n_slots, first_key, matches_per_slot, pk, ek, int(bool(keep_records)),
                                   n_mcts_player, n_mcts_enemy))

    def ensure_log_table(self, n):
        """UCT's log(sum n) values, computed by THIS numpy one scalar at a time like agents.py:556 does"""
        if getattr(self, "_log_n", 0) >= n:
            return
        n = max(n, 1024)
        with np.errstate(divide="ignore"):
            tab = np.asarray([float(np.log(float(k))) for k in range(n)], np.float64)
        tab[0] = 0.0
        check(lib().ao_set_log_table(self._h, ptr(tab), n))
        self._log_n = n

    def rollout_search(self, kind, game_ids, root_ids, num_mcts=0):
        """PUCTAgent / UCTAgent search (fresh tree, num_mcts + 1 simulations with random play-outs)
        -> visits uint32 [n][A], w float32 [n][A] of the root's children"""
        k = SIDE_KINDS[kind]
        if k == AO_SIDE_UCT:
            self.ensure_log_table((num_mcts or self.num_mcts) + 2)
        ids = np.ascontiguousarray(game_ids, np.int32)
        roots, lens = pad_ids(root_ids, self.A)
        n = len(ids)
        visits = np.empty((n, self.A), np.uint32)
        w = np.empty((n, self.A), np.float32)
        check(lib().ao_rollout_search(self._h, k, ptr(ids), n, ptr(roots), ptr(lens), num_mcts, ptr(visits), ptr(w)))
        return visits, w

    def selfplay_begin(self, n_games, first_key=0, recycle=False):
        check(lib().ao_selfplay_begin_mode(self._h, n_games, first_key, int(recycle)))

    def selfplay_stream_begin(self, n_episodes, n_slots=None, first_key=0):
        """continuous self-play: `n_slots` (default: all) concurrent games work through `n_episodes` episodes"""
        check(lib().ao_selfplay_stream_begin(self._h, self.G if n_slots is None else n_slots, first_key, n_episodes))

    def stream_records_dev(self):
        p, b, n = C.c_void_p(), C.c_size_t(), C.c_int32()
        check(lib().ao_selfplay_stream_records_dev(self._h, C.byref(p), C.byref(b), C.byref(n)))
        return p.value, b.value, n.value

    @staticmethod
    def _counters(out, **extra):
        d = dict(sims=int(out[0]), running=int(out[1]), nn_evals=int(out[2]), errors=int(out[3]), moves=int(out[4]),
                 games_finished=int(out[5]), terminal_sims=int(out[6]))
        d.update(extra)
        return d

    def selfplay_rounds(self, rounds):
        out = np.zeros(8, np.uint64)
        check(lib().ao_selfplay_rounds(self._h, rounds, ptr(out)))
        return self._counters(out)

    def selfplay_rounds_timed(self, rounds):
        out = np.zeros(8, np.uint64)
        a, b = C.c_float(0), C.c_float(0)
        check(lib().ao_selfplay_rounds_timed(self._h, rounds, ptr(out), C.byref(a), C.byref(b)))
        return self._counters(out, tree_ms=a.value, tower_ms=b.value)

    def tower_debug(self, enable=True):
        """probe build only (AO_USE_PROBE_LIB=1): cycle counters of CTA 0 of the tower kernel"""
        if not hasattr(lib(), "ao_tower_debug"):
            raise AoError("ao_tower_debug exists only in libalpha_omok_b200_probe.so (set AO_USE_PROBE_LIB=1)")
        out = np.zeros(8, np.uint64)
        check(lib().ao_tower_debug(self._h, int(enable), ptr(out)))
        return [int(x) for x in out]

    def launch_count(self):
        n = C.c_uint64(0)
        check(lib().ao_launch_count(self._h, C.byref(n)))
        return n.value

    def selfplay_fetch(self, n_games, with_visits=True):
        moves = np.empty((n_games, self.A), np.int16)
        n_moves = np.empty(n_games, np.int32)
        winners = np.empty(n_games, np.int8)
        visits = np.empty((n_games, self.A, self.A), np.uint32) if with_visits else None
        check(lib().ao_selfplay_fetch(self._h, n_games, ptr(moves), ptr(n_moves), ptr(winners), ptr(visits)))
        return moves, n_moves, winners, visits

    def nn_log(self, game_id, capacity):
        pol = np.empty((capacity, self.A), np.float32)
        val = np.empty(capacity, np.float32)
        cnt = C.c_int32(0)
        check(lib().ao_get_nn_log(self._h, game_id, ptr(pol), ptr(val), capacity, C.byref(cnt)))
        return pol[:cnt.value], val[:cnt.value]

    def records_dev(self):
        p = C.c_void_p()
        b = C.c_size_t()
        check(lib().ao_records_dev(self._h, C.byref(p), C.byref(b)))
        return p.value, b.value

    def records_pack(self, n_games):
        check(lib().ao_records_pack(self._h, n_games))

    def synchronize(self):
        check(lib().ao_synchronize(self._h))


def check_win_batch(boards, board_size):
    b = np.ascontiguousarray(boards, np.int8).reshape(-1, board_size * board_size)
    out = np.empty(b.shape[0], np.uint8)
    check(lib().ao_check_win(ptr(b), b.shape[0], board_size, ptr(out)))
    return out


def encode_state_batch(root_ids, board_size):
    A = board_size * board_size
    ids, lens = pad_ids(root_ids, A)
    out = np.empty((len(root_ids), 5, board_size, board_size), np.float32)
    check(lib().ao_encode_state(ptr(ids), ptr(lens), len(root_ids), board_size, ptr(out)))
    return out


def legal_actions_batch(root_ids, board_size):
    A = board_size * board_size
    ids, lens = pad_ids(root_ids, A)
    out = np.empty((len(root_ids), A), np.int16)
    check(lib().ao_legal_actions(ptr(ids), ptr(lens), len(root_ids), board_size, ptr(out)))
    return out
