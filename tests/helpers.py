"""Shared helpers for the tests (oracle-side: test infrastructure only)."""
import os

import numpy as np

from oracle import omok_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def unpad_id(row):
    row = [int(x) for x in row]
    return tuple(row[:row.index(-1)]) if -1 in row else tuple(row)


def synth_eval(moves, A, salt=0):
    """Same hash 'network' as tests/golden/make_golden.py and csrc/tree.cu (AO_EVAL_SYNTH); salt = arena side."""
    hh = 0xCBF29CE484222325
    for m in moves[1:]:
        hh = ((hh ^ (int(m) + 1)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    lo, hi = hh & 0xFFFFFFFF, hh >> 32
    pol = np.empty(A, np.float32)
    for a in range(A):
        w = O.philox4x32((a, 0, lo, hi), (0x5EED, 0x0A0A + salt))
        pol[a] = np.float32(((w[0] >> 8) + 1) * 2.0 ** -24)
    w = O.philox4x32((0xFFFF, 0, lo, hi), (0x5EED, 0x0A0A + salt))
    return pol, np.float32((w[1] >> 8) * 2.0 ** -23 - 1.0)


def oracle_game_from_fixture(fx, evaluate=None):
    B, A = int(fx["B"]), int(fx["B"]) ** 2
    mm = int(fx["max_moves"])
    stream = O.DecisionStream(int(fx["seed"]), int(fx["game"]), fx["gamma_tape"])
    if evaluate is None:
        evaluate = lambda mv: synth_eval(mv, A)  # noqa: E731
    return O.self_play_game(B, int(fx["sims"]), evaluate, stream, tau_thres=int(fx["tau_thres"]),
                            noise=bool(fx["noise"]), max_moves=None if mm < 0 else mm)
