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


# ---------------------------------------------------------------------------------------------------------------------
# Host-side helpers of the reference's utils module that are NOT on the hot path (printing, the rollout agents' helpers,
# a TensorFlow-era encoder) - provided so that `import utils` keeps working for main.py / eval_main.py unchanged.
ALPHABET = ' A B C D E F G H I J K L M N O P Q R S'


def valid_actions(board):
    """utils.py:8-19 - [[(row, col), flat index], ...] of the empty cells, row-major"""
    b = np.asarray(board)
    n = len(b)
    return [[(int(i), int(j)), int(i * n + j)] for i, j in zip(*np.nonzero(b == 0))]


def get_reward(win_index, leaf_id):
    """utils.py:208-223 - +1 / -1 from the point of view of the side to move at the leaf, 0 for draw / running"""
    if win_index not in (1, 2):
        return 0.
    mover_lost = (win_index == 1) == (get_turn(leaf_id) == 1)
    return 1. if mover_lost else -1.


def render_str(board, board_size, action_index):
    """utils.py:62-102 - text board: ' .', ' O' (black), ' X' (white); the last move is bracketed '(O)' and the cell to
    its right drops its leading blank so that the columns stay aligned; footer with the move count."""
    b = np.asarray(board)
    count = int(np.count_nonzero(b))
    last = None if action_index is None else (int(action_index) // board_size, int(action_index) % board_size)
    if count > 0 and last is None:
        raise NameError("render_str needs the last action once stones are on the board")  # as the reference does
    lines = ['', '  ' + ALPHABET[:board_size * 2]]
    glyph = {0: '.', 1: 'O', -1: 'X'}
    for i in range(board_size):
        row = '{:2}'.format(i + 1)
        for j in range(board_size):
            g = glyph[int(b[i][j])]
            if last is not None and (i, j) == last and g != '.':
                row += '(' + g + ')'
            elif last is not None and (i, j) == (last[0], last[1] + 1) and (g != '.' or count > 0):
                row += g
            else:
                row += ' ' + g
        lines.append(row + ' ')
    footer = '  ' + '-' * (board_size - 6) + '  MOVE: {:2}  '.format(count) + '-' * (board_size - 6)
    print('\n'.join(lines) + '\n' + footer)


def get_state_tf(id, turn, board_size, channel_size):
    """utils.py:105-136 - channels-last encoder kept from the reference's TensorFlow days (unused by the PyTorch path):
    channel c < channel_size-1 holds the cumulative stones of the colour that moved c plies before the end of `id`,
    the last channel is the colour-to-move flag."""
    state = np.zeros([board_size, board_size, channel_size])
    planes = [np.zeros([board_size, board_size]), np.zeros([board_size, board_size])]  # even / odd positions of id
    n = len(id)
    for i in range(n):
        if i != 0:
            planes[i % 2][int(id[i] / board_size), int(id[i] % board_size)] = 1
        if n - i < channel_size:
            state[:, :, n - i - 1] = planes[i % 2]
    state[:, :, channel_size - 1] = 1 if turn == 0 else 0
    return state
