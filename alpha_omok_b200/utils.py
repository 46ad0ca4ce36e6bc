"""Drop-in for 2_AlphaOmok/utils.py (the functions on the self-play path, SURVEY 8a: a10-a13, a15, a16).

Same names, arguments and return types as the reference.  The rule / encoding functions run on the GPU through the
C ABI (ao_check_win, ao_encode_state, ao_legal_actions - one warp per board; the batched forms live in _cabi);
`get_action` / `argmax_onehot` draw from numpy's global RNG exactly like the reference (utils.py:189-205) so that a
seeded host script keeps its random stream.
"""
from __future__ import annotations

import numpy as np

from . import _cabi


def get_turn(node_id):
    """utils.py:182-186 - 0: black to move, 1: white to move"""
    return (len(node_id) + 1) % 2


def legal_actions(node_id, board_size):
    """utils.py:22-27 - empty cells in the reference's (CPython set) order = child order of the search tree"""
    la = _cabi.legal_actions_batch([tuple(node_id)], board_size)[0]
    return [int(a) for a in la if a >= 0]


def check_win(board, win_mark):
    """utils.py:30-59 - 0 playing / 1 black / 2 white / 3 draw.  Omok only (win_mark 5)."""
    if win_mark != 5:
        raise ValueError("alpha_omok_b200 implements five-in-a-row only (win_mark == 5)")
    b = np.asarray(board)
    return int(_cabi.check_win_batch(np.sign(b).astype(np.int8)[None], b.shape[0])[0])


def get_state_pt(node_id, board_size, channel_size):
    """utils.py:139-168 - [C,B,B] float64 input planes of the network for the position `node_id`"""
    if channel_size != 5:
        raise ValueError("alpha_omok_b200 implements IN_PLANES == 5")
    return _cabi.encode_state_batch([tuple(node_id)], board_size)[0].astype(np.float64)


def get_board(node_id, board_size):
    """utils.py:171-179 - [B,B] float64 board, +1 black / -1 white"""
    st = _cabi.encode_state_batch([tuple(node_id)], board_size)[0]
    own, opp = st[2].astype(np.float64), st[3].astype(np.float64)
    return own - opp if st[4, 0, 0] == 1.0 else opp - own


def get_action(pi):
    """utils.py:189-195 - sample an action index from pi (numpy global RNG), return (one-hot, index)"""
    pi = np.asarray(pi)
    idx = np.random.choice(len(pi), p=pi)
    onehot = np.zeros(len(pi))
    onehot[idx] = 1
    return onehot, idx


def argmax_onehot(pi):
    """utils.py:198-205 - argmax with uniform random tie-break (numpy global RNG), return (one-hot, index)"""
    pi = np.asarray(pi)
    best = np.flatnonzero(pi == pi.max())
    idx = best[np.random.choice(len(best))]
    onehot = np.zeros(len(pi))
    onehot[idx] = 1
    return onehot, idx


def augment_dataset(memory, board_size):
    """utils.py:226-239 - 8-fold dihedral augmentation of (state, pi, z) records (host side; SURVEY 8f 'next')"""
    out = []
    for s, pi, z in memory:
        grid = np.asarray(pi).reshape(board_size, board_size)
        for k in range(4):
            s_k, p_k = np.rot90(s, k, axes=(1, 2)), np.rot90(grid, k)
            out.append((s_k.copy(), p_k.flatten().copy(), z))
            out.append((s_k[:, :, ::-1].copy(), p_k[:, ::-1].flatten().copy(), z))
    return out
